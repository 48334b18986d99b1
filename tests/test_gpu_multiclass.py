"""GPU parity of the multinomial log-likelihood loss (K trees per iteration) against the
oracle's restatement of loss_imp_multinomial.cc and of the K-tree loop (gradient_boosted_trees.cc:1490-1511)."""
import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees, first_divergence, synth_mixed

pytestmark = pytest.mark.gpu


def _oracle_cfg(cfg):
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    return o


def _data(n, K, seed):
    bins, nb, na, ft, y = synth_mixed(n, 5, [6, 30], seed=seed, task="regression")
    rng = np.random.default_rng(seed)
    edges = np.quantile(y, np.linspace(0, 1, K + 1)[1:-1])
    cls = np.searchsorted(edges, y).astype(np.int32)
    flip = rng.random(n) < 0.1
    cls[flip] = rng.integers(0, K, size=int(flip.sum()))
    return bins, nb, na, ft, cls + 1


@pytest.mark.parametrize("K,hess", [(3, 0), (5, 1), (2, 0)])
def test_boosting_loop_matches_oracle(K, hess):
    n, iters = 30000, 8
    bins, nb, na, ft, y = _data(n, K, 50 + K)
    cfg = ydf_b200.default_config(loss=2, num_classes=K, num_trees=iters, max_depth=5, use_hessian_gain=hess)
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    gbt = ydf_b200.Gbt(ds, cfg)
    gbt.set_labels(y)
    gbt.train(iters)
    O.set_stable_category_sort(True)
    O.set_hessian_buckets_double(bool(hess))
    try:
        ref = O.gbt_train_mc(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, feature_type=ft)
    finally:
        O.set_stable_category_sort(False)
        O.set_hessian_buckets_double(False)
    assert gbt.num_trees() == iters * K and gbt.num_iterations() == iters
    got = [gbt.get_tree(i) for i in range(iters * K)]
    # node sums: a 1-ulp difference of a leaf value shifts the gradients of ALL its rows the same way in the
    # following trees, and K trees per iteration feed every class score -> 1e-7 per row instead of 2e-8
    # at iteration 0 every gradient is 1 - 1/K or -1/K: bucket sums cancel exactly in double but not after the
    # 24-bit quantisation, which may move categories that are ABSENT from a node across its split
    t, errs = first_divergence(got, ref["trees"], stat_atol_per_row=1e-7, present_in=bins)
    assert t is None, (t, errs[:8])
    for i in range(iters):
        l, a = gbt.train_loss(i)
        assert abs(l - ref["loss"][i]) <= 1e-5 * abs(ref["loss"][i]) and abs(a - ref["secondary"][i]) <= 1e-5
    np.testing.assert_allclose(gbt.get_predictions(), ref["predictions"], rtol=0, atol=2e-5)
    assert gbt.initial_prediction() == 0.0


def test_error_behaviour():
    bins, nb, na, ft, y = _data(2000, 3, 1)
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    with pytest.raises(ydf_b200.YggError):
        ydf_b200.Gbt(ds, ydf_b200.default_config(loss=2, num_classes=1))
    gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=2, num_classes=3, num_trees=2))
    bad = y.copy()
    bad[7] = 4
    with pytest.raises(ydf_b200.YggError) as e:
        gbt.set_labels(bad)
    assert "Expected value between 1 and 3" in str(e.value)      # loss_imp_multinomial.cc:86-89


def test_learner_multiclass_with_validation(tmp_path):
    rng = np.random.default_rng(2)
    n = 12000
    x = rng.normal(size=(n, 3)).astype(np.float32)
    kind = np.where(x[:, 0] > 0.5, "c", np.where(x[:, 1] > 0, "b", "a"))
    noise = rng.random(n) < 0.15
    kind[noise] = rng.choice(["a", "b", "c"], size=int(noise.sum()))
    data = {"x0": x[:, 0], "x1": x[:, 1], "x2": x[:, 2], "y": kind}
    learner = ydf_b200.GradientBoostedTreesLearner(label="y", discretize_numerical_columns=True, num_trees=60,
                                                   shrinkage=0.3, max_depth=4, early_stopping_num_trees_look_ahead=9)
    model = learner.train(data)
    assert model.loss == "MULTINOMIAL_LOG_LIKELIHOOD" and model.num_trees_per_iter() == 3
    assert model.num_trees() % 3 == 0 and model.early_stopping_triggered and model.num_trees() < 180
    logs = model.training_logs
    best = min(range(10, len(logs)), key=lambda i: (logs[i]["validation_loss"], i))
    assert model.num_trees() == (best + 1) * 3                  # EarlyStopping counts trees: look-ahead 9 = 3 iterations
    assert len(logs) == best + 1 + 3
    p = model.predict(data)
    assert p.shape == (n, 3) and np.allclose(p.sum(axis=1), 1, atol=1e-5)
    assert model.evaluate(data)["accuracy"] > 0.8
    model.save(str(tmp_path / "m"))
    back = ydf_b200.model_io.read_ydf_model(str(tmp_path / "m"))
    assert back["loss"] == 3 and back["num_trees_per_iter"] == 3 and back["num_trees"] == model.num_trees()
    assert back["initial_predictions"] == [0.0, 0.0, 0.0]
    assert back["training_logs"][0]["number_of_trees"] == 3
