"""GPU parity of the validation hold-out and early stopping against the oracle's
restatement of the learner loop (oracle_gbt_train_validated)."""
import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees, synth_mixed

pytestmark = pytest.mark.gpu


def _oracle_cfg(cfg):
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    return o


def _noisy(n, seed, task):
    bins, nb, na, ft, y = synth_mixed(n, 4, [6, 40], seed=seed, task=task)
    rng = np.random.default_rng(seed)
    if task == "binary":
        flip = rng.random(n) < 0.3
        y = np.where(flip, 3 - y, y).astype(np.int32)
    else:
        y = (y + rng.normal(scale=2.0, size=n)).astype(np.float32)
    return bins, nb, na, ft, y


def test_split_rows_is_a_stable_gather():
    bins, nb, na, ft, y = _noisy(30000, 1, "binary")
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    m = ydf_b200.validation_split_mask(7, bins.shape[1], 0.25)
    a, b = ds.split_rows(m)
    assert a.n_rows == int(m.sum()) and b.n_rows == int((~m).sum())
    for f in range(bins.shape[0]):
        np.testing.assert_array_equal(a.get_bins(f), bins[f, m])
        np.testing.assert_array_equal(b.get_bins(f), bins[f, ~m])


@pytest.mark.parametrize("policy", [2, 1, 0])
@pytest.mark.parametrize("loss", [0, 1])
def test_early_stopping_matches_oracle(loss, policy):
    n = 6000
    bins, nb, na, ft, y = _noisy(n, 3 + loss, "binary" if loss == 0 else "regression")
    kw = dict(loss=loss, num_trees=90, max_depth=6, shrinkage=0.3, min_examples=2, early_stopping=policy,
              early_stopping_num_trees_look_ahead=12, early_stopping_initial_iteration=4)
    cfg = ydf_b200.default_config(**kw)
    O.set_stable_category_sort(True)
    try:
        ref = O.gbt_train_validated(bins, nb, na, y, _oracle_cfg(cfg), 0.2, num_threads=4, feature_type=ft)
    finally:
        O.set_stable_category_sort(False)
    full = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    m = ydf_b200.validation_split_mask(cfg.random_seed, n, 0.2)
    np.testing.assert_array_equal(m, ref["in_training"])
    tr, va = full.split_rows(m)
    gbt = ydf_b200.Gbt(tr, cfg)
    gbt.set_labels(y[m])
    gbt.set_validation(va, y[~m])
    gbt.train(cfg.num_trees)
    assert gbt.num_iterations() == ref["num_entries"]
    assert gbt.num_trees() == len(ref["trees"])
    if policy == 2:
        assert gbt.num_iterations() < cfg.num_trees, "the problem is meant to stop early"
    for i in range(gbt.num_trees()):
        errs = compare_trees(gbt.get_tree(i), ref["trees"][i])
        assert not errs, (i, errs[:5])
    for i in range(gbt.num_iterations()):
        tl, _ = gbt.train_loss(i)
        vl, vs = gbt.validation_loss(i)
        assert abs(tl - ref["train_loss"][i]) <= 1e-5 * abs(ref["train_loss"][i])
        assert abs(vl - ref["valid_loss"][i]) <= 1e-5 * abs(ref["valid_loss"][i])
        assert abs(vs - ref["valid_secondary"][i]) <= 1e-5 * max(1.0, abs(ref["valid_secondary"][i]))
    fv, trig = gbt.final_validation()
    assert trig == ref["early_stopping_triggered"]
    assert abs(fv - ref["validation_loss"]) <= 1e-5 * abs(ref["validation_loss"])
    if policy != 0:
        with pytest.raises(ydf_b200.YggError):
            gbt.step()      # finalized


def test_learner_defaults_hold_out_and_stop(tmp_path):
    """Reference defaults: validation_ratio=0.1, early_stopping=LOSS_INCREASE (look-ahead 30)."""
    rng = np.random.default_rng(5)
    n = 8000
    x = rng.normal(size=(n, 3)).astype(np.float32)
    lab = np.where((x[:, 0] > 0) ^ (rng.random(n) < 0.3), "a", "b")
    data = {"x0": x[:, 0], "x1": x[:, 1], "x2": x[:, 2], "y": lab}
    learner = ydf_b200.GradientBoostedTreesLearner(label="y", discretize_numerical_columns=True, num_trees=200,
                                                   shrinkage=0.3, max_depth=6)
    model = learner.train(data)
    logs = model.training_logs
    assert model.early_stopping_triggered and model.num_trees() < 200
    assert len(logs) == model.num_trees() + 30              # stopped 30 trees after the best validation loss
    best = min(range(10, len(logs)), key=lambda i: (logs[i]["validation_loss"], i))
    assert model.num_trees() == best + 1 and model.validation_loss == logs[best]["validation_loss"]
    model.save(str(tmp_path / "m"))
    back = ydf_b200.model_io.read_ydf_model(str(tmp_path / "m"))
    assert back["num_trees"] == model.num_trees() and back["early_stopping_triggered"]
    assert abs(back["validation_loss"] - model.validation_loss) < 1e-7
    assert len(back["training_logs"]) == len(logs) and "validation_loss" in back["training_logs"][0]
    # an explicit validation dataset instead of the hold-out
    m2 = learner.train({k: v[:6000] for k, v in data.items()}, valid={k: v[6000:] for k, v in data.items()})
    assert m2.validation_loss is not None and len(m2.training_logs) >= m2.num_trees()


@pytest.mark.parametrize("loss", [0, 2])
def test_device_predictions_match_a_host_walk_of_the_trees(loss):
    """ygg_gbt_predict (ComputePredictions, gradient_boosted_trees.cc:2872-2930): raw scores of the trained model on the
    training rows (== the boosting state the engine keeps) and on another dataset (== a numpy walk of the fetched trees)."""
    from tests.util import predict_raw, synth_mixed
    bins, nb, na, ft, y = synth_mixed(40000, 5, [6, 40], seed=9)
    K = 3 if loss == 2 else 1
    labels = y if loss == 0 else (bins[0].astype(np.int32) % 3) + 1
    ds = ydf_b200.Dataset(bins[:, :30000], nb, na, feature_types=ft)
    other = ydf_b200.Dataset(bins[:, 30000:], nb, na, feature_types=ft)
    gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=loss, num_classes=K if loss == 2 else 0, max_depth=5, num_trees=6))
    gbt.set_labels(labels[:30000])
    gbt.train(6)
    np.testing.assert_allclose(gbt.predict(ds), gbt.get_predictions(), rtol=0, atol=1e-6)
    trees = [gbt.get_tree(i) for i in range(gbt.num_trees())]
    got = gbt.predict(other)
    for k in range(K):
        want = predict_raw(trees[k::K], gbt.initial_prediction(), bins[:, 30000:])
        np.testing.assert_allclose(got[:, k] if K > 1 else got, want, rtol=0, atol=1e-6)
