"""CPU tests of the validation hold-out: the row draw and the oracle's restatement of
early stopping."""
import numpy as np

import ydf_b200
from oracle import oracle as O
from tests.util import synth


def test_split_mask_is_the_mt19937_draw():
    """ExtractValidationDataset (gradient_boosted_trees.cc:2731-2738): row r trains iff
    uniform_real_distribution<float>(mt19937(seed)) > ratio.  libstdc++ draws one 32-bit word per float:
    generate_canonical = word * 2^-32 (clamped below 1), so numpy's MT19937 stream reproduces it."""
    n, seed, ratio = 5000, 123456, 0.1
    m = ydf_b200.validation_split_mask(seed, n, ratio)
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)       # init_genrand(seed) = std::mt19937(seed)
    words = bg.random_raw(n).astype(np.float64)
    u = (words * 2.0 ** -32).astype(np.float32)
    u = np.minimum(u, np.nextafter(np.float32(1), np.float32(0)))
    np.testing.assert_array_equal(m, u > np.float32(ratio))
    assert 0.08 < 1 - m.mean() < 0.12
    assert ydf_b200.validation_split_mask(seed, 100, 0.0).all()
    # the oracle draws the same rows
    bins, nb, na, y = synth(n, 3, seed=1, bins=16)
    cfg = O.default_config(num_trees=2, max_depth=3)
    r = O.gbt_train_validated(bins, nb, na, y, cfg, ratio)
    np.testing.assert_array_equal(r["in_training"], m)


def test_oracle_early_stopping_policies():
    """EarlyStopping::Update / ShouldStop (early_stopping.cc:30-62) and FinalizeModelWithValidationDataset
    (gradient_boosted_trees.cc:212-272) on a noisy problem that overfits quickly."""
    rng = np.random.default_rng(0)
    n = 3000
    bins = rng.integers(0, 32, size=(6, n)).astype(np.uint8)
    y = ((bins[0] > 15) ^ (rng.random(n) < 0.35)).astype(np.int32) + 1
    nb, na = [32] * 6, [0] * 6
    base = dict(num_trees=120, max_depth=6, shrinkage=0.3, min_examples=2)
    inc = O.gbt_train_validated(bins, nb, na, y, O.default_config(early_stopping=2, early_stopping_num_trees_look_ahead=10,
                                                                  early_stopping_initial_iteration=3, **base), 0.2)
    k = inc["num_entries"]
    vl = inc["valid_loss"]
    best = 3 + int(np.argmin(vl[3:]))                      # first minimum from the initial iteration on (strict <)
    assert inc["early_stopping_triggered"] and k < 120      # stopped early
    assert len(inc["trees"]) == best + 1 and k == best + 1 + 10
    assert abs(inc["validation_loss"] - vl[best]) == 0
    full = O.gbt_train_validated(bins, nb, na, y, O.default_config(early_stopping=1, early_stopping_initial_iteration=3, **base), 0.2)
    assert full["num_entries"] == 120 and len(full["trees"]) == 3 + int(np.argmin(full["valid_loss"][3:])) + 1
    none = O.gbt_train_validated(bins, nb, na, y, O.default_config(early_stopping=0, **base), 0.2)
    assert len(none["trees"]) == 120 and not none["early_stopping_triggered"]
    assert none["validation_loss"] == none["valid_loss"][-1]
    # too few trees for early stopping: everything is kept, the last loss is reported
    few = O.gbt_train_validated(bins, nb, na, y, O.default_config(early_stopping=2, early_stopping_initial_iteration=10,
                                                                  **dict(base, num_trees=8)), 0.2)
    assert len(few["trees"]) == 8 and not few["early_stopping_triggered"] and few["validation_loss"] == few["valid_loss"][-1]
    # the trees are those of a plain run on the training rows
    tr = inc["in_training"]
    plain = O.gbt_train(bins[:, tr], nb, na, y[tr], O.default_config(**base), 5)
    for a, b in zip(plain["trees"], inc["trees"][:5]):
        assert a.tobytes() == b.tobytes()


def test_hold_out_and_early_stopping_against_a_reference_run():
    """test_data/model/adult_binary_class_gbdt_v2 is `ydf.GradientBoostedTreesLearner(label="income")` trained by the
    reference on adult_train.csv with default hyper-parameters (fixture tests/golden/ydf_adult_gbdt_v2_head.npz).  Its
    header and logs pin, against a REAL reference run:
      * the hold-out draw: ygg_validation_split_mask(123456, 22792, 0.1) keeps exactly the 20533 rows the reference
        trained on, and the held-out / kept rows have the reference's class shares (its first log entry is the accuracy
        of the still-constant model on each part);
      * the initial prediction = log-odds of the KEPT rows only (gradient_boosted_trees.cc:1288-1296);
      * the early-stopping bookkeeping: the model keeps best_num_trees = argmin(validation loss) + 1 trees, training
        went on for exactly early_stopping_num_trees_look_ahead = 30 more iterations, Header.validation_loss is the
        best loss (early_stopping.cc:30-62, gradient_boosted_trees.cc:212-272) — what ygg_gbt_train replays."""
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref = np.load(os.path.join(G, "ydf_adult_gbdt_v2_head.npz"))
    y = np.load(os.path.join(G, "adult_numerical.npz"))["train_income"]          # 1 = ">50K"
    mask = ydf_b200.validation_split_mask(123456, len(y), 0.1)
    assert len(y) == 22792 and int(mask.sum()) == int(ref["root_num_examples"]) == 20533
    assert np.float32(np.mean(y[~mask] == 0)) == ref["first_validation_accuracy"]
    assert np.float32(np.mean(y[mask] == 0)) == ref["first_training_accuracy"]
    ratio = np.mean(y[mask] == 1)
    assert np.float32(np.log(ratio / (1 - ratio))) == ref["initial_prediction"]
    assert abs(O.initial_prediction(0, y[mask].astype(np.int32) + 1) - float(ref["initial_prediction"])) == 0
    assert int(ref["num_log_entries"]) - int(ref["num_trees"]) == 30 == ydf_b200.default_config().early_stopping_num_trees_look_ahead
    assert int(ref["best_validation_loss_entry"]) + 1 == int(ref["num_trees"]) == 163
    assert int(ref["last_number_of_trees"]) == int(ref["num_log_entries"])


def test_regression_hold_out_against_a_reference_run():
    """The same for squared error: test_data/model/abalone_regression_gbdt_v2 (PYDF defaults on abalone.csv, 4177 rows).
    The reference trained on the 3771 rows the mask keeps; its initial prediction is their mean label
    (loss_imp_mean_square_error.cc:56-88); the root of tree 0 stores sum and sum of squares of the first gradients
    g = y - initial prediction (float, :113-118; squares taken in float, utils/distribution.h:66-71); 45 trees are kept,
    30 more iterations were trained."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ydf_abalone_gbdt_v2_head.npz"))
    y = ref["rings"].astype(np.float32)
    mask = ydf_b200.validation_split_mask(123456, len(y), 0.1)
    assert len(y) == 4177 and int(mask.sum()) == int(ref["root_num_examples"]) == 3771 and int(ref["loss"]) == 2
    init = O.initial_prediction(1, y[mask])
    assert np.float32(init) == ref["initial_prediction"]
    g = (y[mask] - np.float32(init)).astype(np.float32)
    assert abs(float(g.astype(np.float64).sum()) - float(ref["root_sum"])) < 1e-6
    assert abs(float((g * g).astype(np.float64).sum()) - float(ref["root_sum_squares"])) < 1e-6 * float(ref["root_sum_squares"])
    assert int(ref["num_trees"]) == int(ref["best_validation_loss_entry"]) + 1 == 45
    assert int(ref["num_log_entries"]) == 45 + 30
    # the oracle's root statistics for the same rows
    bins = np.zeros((1, int(mask.sum())), np.uint8)
    t = O.gbt_train(bins, [2], [0], y[mask], O.default_config(loss=1, num_trees=1, max_depth=2), 1)["trees"][0]
    assert abs(t[0]["stat"][0] - float(ref["root_sum"])) < 1e-6 and t[0]["stat"][2] == 3771
    assert abs(t[0]["stat"][1] - float(ref["root_sum_squares"])) < 1e-6 * float(ref["root_sum_squares"])
