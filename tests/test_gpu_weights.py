"""Example weights: ygg_gbt_set_weights_f32 against the oracle's restatement of the reference's
weighted path — LabelNumericalBucket<weighted=true> (splitter_accumulator.h:1552-1560), SetLeafValueWithNewtonRaphsonStep<true>
(loss_utils.cc:81-89), the weighted losses / initial predictions (loss_imp_binomial.cc:65-99, :204-234;
loss_imp_mean_square_error.cc:56-88; metric/metric.cc:2097-2115).  Structure and counts exact, scores 1e-5 relative,
leaves 1e-5 absolute, losses 1e-5 relative; node statistics (weighted sum, sum of squares, WEIGHT sum) 1e-6 relative.
"""
import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees, synth, synth_mixed

pytestmark = pytest.mark.gpu


def _oracle_cfg(cfg):
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    return o


def _weights(kind, y, n, seed=3):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(0.1, 3.0, n).astype(np.float32)
    if kind == "class":       # the usual use: re-balance the classes
        return np.where(np.asarray(y) == 2, 3.0, 1.0).astype(np.float32)
    if kind == "heavy":       # a long tail, some zero weights, max far above the mean
        w = rng.lognormal(0.0, 1.5, n).astype(np.float32)
        w[rng.random(n) < 0.05] = 0.0
        return w
    raise ValueError(kind)


def _train_both(bins, nb, na, y, w, iters, ft=None, **kw):
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    cfg = ydf_b200.default_config(num_trees=iters, **kw)
    gbt = ydf_b200.Gbt(ds, cfg)
    gbt.set_weights(w)
    gbt.set_labels(y)
    gbt.train(iters)
    O.set_weights(w)
    try:
        ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, feature_type=ft)
    finally:
        O.set_weights(None)
    return gbt, ref


@pytest.mark.parametrize("loss,kind,extra", [
    (1, "uniform", {}),
    (0, "class", {}),
    (0, "heavy", {}),
    (1, "heavy", dict(min_examples=20)),
    (0, "uniform", dict(subsample=0.7)),
    (1, "uniform", dict(sibling_subtraction=0)),
])
def test_weighted_training_matches_oracle(loss, kind, extra):
    n, iters = 40000, 10
    bins, nb, na, y = synth(n, 10, seed=31, task="binary" if loss == 0 else "regression", bins=128)
    w = _weights(kind, y, n)
    gbt, ref = _train_both(bins, nb, na, y, w, iters, loss=loss, max_depth=6, **extra)
    for i in range(iters):
        got, want = gbt.get_tree(i), ref["trees"][i]
        # node sums are 31-bit fixed point relative to the iteration's max |w*g|: with a long-tailed weight distribution
        # that scale, and with it the absolute resolution per row, grows
        scale = max(1.0, float(w.max()) * (1.0 if loss == 0 else float(np.abs(y - y.mean()).max())))
        errs = compare_trees(got, want, stat_atol_per_row=2e-8 * scale)
        assert not errs, (i, errs[:6])
        l, s = gbt.train_loss(i)
        assert abs(l - ref["loss"][i]) <= 1e-5 * abs(ref["loss"][i]), (i, l, ref["loss"][i])
        assert abs(s - ref["secondary"][i]) <= 1e-5, (i, s, ref["secondary"][i])
    # the root's count statistic is the weight sum of the rows the tree was trained on
    if "subsample" not in extra:
        assert abs(gbt.get_tree(0)[0]["stat"][2] - float(np.sum(w.astype(np.float64)))) <= 1e-6 * float(w.sum())
    np.testing.assert_allclose(gbt.get_predictions(), ref["predictions"], rtol=0, atol=2e-5)
    assert len(gbt.get_tree(0)) > 15


def test_weighted_categorical_features_match_oracle():
    bins, nb, na, ft, y = synth_mixed(30000, 5, [12, 40, 200], seed=9, task="regression")
    w = _weights("uniform", y, 30000, seed=8)
    O.set_stable_category_sort(True)
    try:
        gbt, ref = _train_both(bins, nb, na, y, w, 6, ft=ft, loss=1, max_depth=5)
    finally:
        O.set_stable_category_sort(False)
    cat = 0
    for i in range(6):
        got, want = gbt.get_tree(i), ref["trees"][i]
        errs = compare_trees(got, want)
        assert not errs, (i, errs[:6])
        cat += int((got["condition_type"] == 1).sum())
    assert cat > 0


def test_unit_weights_equal_the_unweighted_run():
    """weights == 1 take the weighted kernels and must give the unweighted trees (same integers, same scores)."""
    bins, nb, na, y = synth(30000, 8, seed=4, task="binary", bins=64)
    ds = ydf_b200.Dataset(bins, nb, na)
    cfg = ydf_b200.default_config(loss=0, max_depth=6, num_trees=5)
    a = ydf_b200.Gbt(ds, cfg)
    a.set_labels(y)
    a.train(5)
    b = ydf_b200.Gbt(ds, cfg)
    b.set_weights(np.ones(30000, np.float32))
    b.set_labels(y)
    b.train(5)
    for i in range(5):
        ta, tb = a.get_tree(i), b.get_tree(i)
        for k in ("feature", "threshold_bin", "num_examples", "num_pos_examples"):
            assert np.array_equal(ta[k], tb[k]), (i, k)
        np.testing.assert_allclose(ta["split_score"], tb["split_score"], rtol=1e-6)
        np.testing.assert_allclose(ta["leaf_value"], tb["leaf_value"], rtol=0, atol=1e-7)
        assert abs(a.train_loss(i)[0] - b.train_loss(i)[0]) <= 1e-6


def test_integer_weights_equal_repeated_rows():
    """Domain property: weight k == the row repeated k times (min_examples = 1 so that row counts do not matter)."""
    n = 20000
    bins, nb, na, y = synth(n, 6, seed=12, task="regression", bins=32)
    k = np.random.default_rng(1).integers(1, 4, n)
    idx = np.repeat(np.arange(n), k)
    cfg = ydf_b200.default_config(loss=1, max_depth=5, num_trees=4, min_examples=1)
    ds = ydf_b200.Dataset(bins, nb, na)
    a = ydf_b200.Gbt(ds, cfg)
    a.set_weights(k.astype(np.float32))
    a.set_labels(y)
    a.train(4)
    ds2 = ydf_b200.Dataset(np.ascontiguousarray(bins[:, idx]), nb, na)
    b = ydf_b200.Gbt(ds2, cfg)
    b.set_labels(y[idx])
    b.train(4)
    for i in range(4):
        ta, tb = a.get_tree(i), b.get_tree(i)
        assert np.array_equal(ta["feature"], tb["feature"]) and np.array_equal(ta["threshold_bin"], tb["threshold_bin"]), i
        np.testing.assert_allclose(ta["leaf_value"], tb["leaf_value"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(ta["split_score"], tb["split_score"], rtol=1e-5)
        assert abs(a.train_loss(i)[0] - b.train_loss(i)[0]) <= 1e-5 * b.train_loss(i)[0]


def test_weighted_validation_and_early_stopping_match_oracle():
    n = 30000
    bins, nb, na, y = synth(n, 8, seed=17, task="binary", bins=64)
    w = _weights("uniform", y, n, seed=5)
    cfg = ydf_b200.default_config(loss=0, max_depth=4, num_trees=25, validation_ratio=0.1,
                                  early_stopping_num_trees_look_ahead=5, early_stopping_initial_iteration=3)
    O.set_weights(w)
    try:
        ref = O.gbt_train_validated(bins, nb, na, y, _oracle_cfg(cfg), 0.1, num_threads=4)
    finally:
        O.set_weights(None)
    tr = ref["in_training"]
    ds = ydf_b200.Dataset(np.ascontiguousarray(bins[:, tr]), nb, na)
    vds = ydf_b200.Dataset(np.ascontiguousarray(bins[:, ~tr]), nb, na)
    gbt = ydf_b200.Gbt(ds, cfg)
    gbt.set_weights(w[tr])
    gbt.set_labels(y[tr])
    gbt.set_validation(vds, y[~tr], weights=w[~tr])
    gbt.train(25)
    assert gbt.num_iterations() == ref["num_entries"]
    assert gbt.num_trees() == len(ref["trees"])
    for i in range(ref["num_entries"]):
        vl, vs = gbt.validation_loss(i)
        assert abs(vl - ref["valid_loss"][i]) <= 1e-5 * abs(ref["valid_loss"][i]), (i, vl, ref["valid_loss"][i])
        assert abs(vs - ref["valid_secondary"][i]) <= 1e-5
        assert abs(gbt.train_loss(i)[0] - ref["train_loss"][i]) <= 1e-5 * abs(ref["train_loss"][i])
    fv, trig = gbt.final_validation()
    assert abs(fv - ref["validation_loss"]) <= 1e-5 * abs(ref["validation_loss"])
    assert trig == ref["early_stopping_triggered"]


def test_weights_are_refused_where_not_implemented():
    bins, nb, na, y = synth(5000, 4, seed=2, task="binary", bins=16)
    ds = ydf_b200.Dataset(bins, nb, na)
    w = np.ones(5000, np.float32)
    g = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=0, use_hessian_gain=1))
    with pytest.raises(ydf_b200.YggError, match="variance gain"):
        g.set_weights(w)
    g = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=0))
    g.set_labels(y)
    with pytest.raises(ydf_b200.YggError, match="before the labels"):
        g.set_weights(w)
    g = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=0))
    bad = w.copy()
    bad[7] = -1.0
    with pytest.raises(ydf_b200.YggError, match="negative"):
        g.set_weights(bad)
    with pytest.raises(ydf_b200.YggError, match="sum of the weights"):
        g.set_weights(np.zeros(5000, np.float32))


def test_learner_weights_column():
    """GradientBoostedTreesLearner(weights="w"): the column is not a feature, follows its rows into the hold-out, and
    re-balancing the classes moves the decision threshold the way the weighted loss says."""
    rng = np.random.default_rng(0)
    n = 30000
    x0, x1 = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    lab = np.where(x0 + 0.5 * x1 + 1.2 * rng.normal(size=n) > 1.0, "pos", "neg")   # ~25 % positives, noisy
    w = np.where(lab == "pos", 4.0, 1.0).astype(np.float32)
    data = {"x0": x0, "x1": x1, "y": lab, "w": w}
    kw = dict(label="y", discretize_numerical_columns=True, num_trees=30, max_depth=4, validation_ratio=0.1)
    plain = ydf_b200.GradientBoostedTreesLearner(**kw).train({k: v for k, v in data.items() if k != "w"})
    weighted = ydf_b200.GradientBoostedTreesLearner(weights="w", **kw).train(data)
    assert [c.name for c in weighted.data_spec.columns] == ["x0", "x1"]
    pp, pw = plain.predict(data), weighted.predict(data)
    assert (pw > 0.5).mean() > (pp > 0.5).mean() + 0.05      # up-weighted positives are predicted more often
    recall = lambda p: ((p > 0.5) & (lab == "pos")).sum() / (lab == "pos").sum()
    assert recall(pw) > recall(pp) + 0.1
    assert all(np.isfinite(e["validation_loss"]) for e in weighted.training_logs)
    with pytest.raises(NotImplementedError):
        ydf_b200.GradientBoostedTreesLearner(label="y", weights="w", use_hessian_gain=True)


@pytest.mark.parametrize("loss,alpha,beta,shuffle", [(0, 0.2, 0.1, 0), (1, 0.2, 0.1, 0), (1, 0.5, 0.0, 0), (1, 0.0, 0.3, 0), (0, 0.3, 0.2, 2)])
def test_goss_matches_oracle(loss, alpha, beta, shuffle):
    """Gradient-based one-side sampling (SampleTrainingExamplesWithGoss, gradient_boosted_trees.cc:2958-3007): the
    ceil(alpha n) rows of largest |g| plus each other row with probability beta and weight (1 - alpha) / beta, one engine
    word per row of the tail in sorted order (after the candidate shuffles of the previous tree when those are replayed).
    The oracle's sampler is pinned on the reference's KAT; rows with EQUAL |g| are ordered by row index on both sides here
    (oracle.set_goss_stable_sort: the reference's std::sort leaves them in its standard library's order) — at iteration 0 of
    the binomial loss every row has one of two values, so the whole first sample depends on it."""
    n, iters = 40000, 8
    bins, nb, na, y = synth(n, 10, seed=41, task="binary" if loss == 0 else "regression", bins=64)
    ds = ydf_b200.Dataset(bins, nb, na)
    cfg = ydf_b200.default_config(num_trees=iters, loss=loss, max_depth=6, goss_alpha=alpha, goss_beta=beta,
                                  candidate_shuffle=shuffle, split_jobs_draw_seeds=int(shuffle != 0))
    gbt = ydf_b200.Gbt(ds, cfg)
    gbt.set_labels(y)
    gbt.train(iters)
    O.set_goss_stable_sort(True)
    try:
        ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, shuffle_candidates=shuffle)
    finally:
        O.set_goss_stable_sort(False)
    amp = (1 - alpha) / beta if beta > 0 else 1.0
    # The sample is a function of the ORDER of 40 000 floats: a gradient that differs in its last bit between the device's
    # and glibc's expf (or through a leaf value that differs by 1e-8) can swap two neighbours around the cut-off and change
    # the sample by a row.  The first trees are compared node by node; over the whole run the losses stay together.
    strict = 4
    for i in range(iters):
        got, want = gbt.get_tree(i), ref["trees"][i]
        assert got[0]["num_examples"] == want[0]["num_examples"] < n      # the root holds the sample (its SIZE is the draws')
        l, s = gbt.train_loss(i)
        if i < strict:
            scale = max(1.0, amp * (1.0 if loss == 0 else float(np.abs(y - y.mean()).max())))
            errs = compare_trees(got, want, stat_atol_per_row=2e-8 * scale)
            assert not errs, (i, errs[:5])
            assert abs(l - ref["loss"][i]) <= 1e-5 * abs(ref["loss"][i]) and abs(s - ref["secondary"][i]) <= 1e-5
        else:
            assert abs(l - ref["loss"][i]) <= 5e-3 * abs(ref["loss"][i]), (i, l, ref["loss"][i])
    assert np.abs(gbt.get_predictions() - ref["predictions"]).mean() < 5e-3
    # expected sample size: alpha n + beta (1 - alpha) n
    assert abs(gbt.get_tree(0)[0]["num_examples"] - (alpha + beta * (1 - alpha)) * n) < 0.02 * n


def test_goss_is_refused_where_not_implemented():
    bins, nb, na, y = synth(5000, 4, seed=2, task="binary", bins=16)
    ds = ydf_b200.Dataset(bins, nb, na)
    for kw, pat in ((dict(use_hessian_gain=1), "variance gain"), (dict(subsample=0.5), "alternative sampling"),
                    (dict(loss=2, num_classes=3), "multinomial")):
        with pytest.raises(ydf_b200.YggError, match=pat):
            ydf_b200.Gbt(ds, ydf_b200.default_config(goss_alpha=0.2, goss_beta=0.1, **kw))
    g = ydf_b200.Gbt(ds, ydf_b200.default_config(goss_alpha=0.2, goss_beta=0.1))
    with pytest.raises(ydf_b200.YggError, match="GOSS"):
        g.set_weights(np.ones(5000, np.float32))


@pytest.mark.parametrize("K,validated", [(3, False), (4, True)])
def test_weighted_multinomial_matches_oracle(K, validated):
    """Example weights with the multinomial loss (loss_imp_multinomial.cc:238-256: loss -= w log(...), accuracy by weight; K trees
    per iteration, each grown on its class plane of w*g / w*h with its own fixed-point scale)."""
    n, iters = 24000, 6
    bins, nb, na, ft, yr = synth_mixed(n, 5, [6, 30], seed=61, task="regression")
    edges = np.quantile(yr, np.linspace(0, 1, K + 1)[1:-1])
    y = (np.searchsorted(edges, yr) + 1).astype(np.int32)
    w = _weights("uniform", y, n, seed=7)
    cfg = ydf_b200.default_config(loss=2, num_classes=K, num_trees=iters, max_depth=5,
                                  validation_ratio=0.2 if validated else 0.0, early_stopping_initial_iteration=2,
                                  early_stopping_num_trees_look_ahead=3 * K)
    O.set_stable_category_sort(True)
    O.set_weights(w)
    try:
        if validated:
            ref = O.gbt_train_validated(bins, nb, na, y, _oracle_cfg(cfg), 0.2, num_threads=4, feature_type=ft)
        else:
            ref = O.gbt_train_mc(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, feature_type=ft)
    finally:
        O.set_weights(None)
        O.set_stable_category_sort(False)
    if validated:
        tr = ref["in_training"]
        ds = ydf_b200.Dataset(np.ascontiguousarray(bins[:, tr]), nb, na, feature_types=ft)
        vds = ydf_b200.Dataset(np.ascontiguousarray(bins[:, ~tr]), nb, na, feature_types=ft)
        gbt = ydf_b200.Gbt(ds, cfg)
        gbt.set_weights(w[tr])
        gbt.set_labels(y[tr])
        gbt.set_validation(vds, y[~tr], weights=w[~tr])
        train_bins = bins[:, tr]
    else:
        ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
        gbt = ydf_b200.Gbt(ds, cfg)
        gbt.set_weights(w)
        gbt.set_labels(y)
        train_bins = bins
    gbt.train(iters)
    from tests.util import first_divergence
    got = [gbt.get_tree(i) for i in range(gbt.num_trees())]
    assert len(got) == len(ref["trees"])
    t, errs = first_divergence(got, ref["trees"], stat_atol_per_row=1e-7, present_in=train_bins)
    assert t is None, (t, errs[:6])
    for i in range(gbt.num_iterations()):
        l, a = gbt.train_loss(i)
        want_l = ref["train_loss"][i] if validated else ref["loss"][i]
        assert abs(l - want_l) <= 1e-5 * abs(want_l), (i, l, want_l)
        if validated:
            vl, va = gbt.validation_loss(i)
            assert abs(vl - ref["valid_loss"][i]) <= 1e-5 * abs(ref["valid_loss"][i]) and abs(va - ref["valid_secondary"][i]) <= 1e-5
        else:
            assert abs(a - ref["secondary"][i]) <= 1e-5
