"""GPU parity tests for categorical features: the CART categorical splitter —
buckets sorted by label mean / hessian priority, scanned like a numerical feature, positive set =
the buckets after the best boundary (splitter_scanner.h:1823-1826, splitter_accumulator.h:391-411,
:1492-1494, :1797-1804) — against the CPU oracle in stable-sort mode (equal sort keys ordered by
category index, see oracle header).  Same bar as test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees, first_divergence, synth_mixed

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _stable_sort():
    O.set_stable_category_sort(True)
    yield
    O.set_stable_category_sort(False)


def _oracle_cfg(cfg):
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    return o


def _mk(bins, nb, na, ft, **kw):
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    cfg = ydf_b200.default_config(**kw)
    return ds, ydf_b200.Gbt(ds, cfg), cfg


def test_kat_categorical_split():
    """decision_tree_test.cc:1208-1297 (FindBestCategoricalSplitCartNumericalLabels, unweighted):
    attributes {2,3,0,1,NA,NA} with the NA replacement 1, labels {1,1,0,0,1,0}, 4 categories ->
    positive set {2,3}, 2 positive rows, na_value false, score 0.125."""
    attr = np.array([[2, 3, 0, 1, 1, 1]], np.uint8)  # NA already replaced by most_frequent_value = 1
    g = np.array([1, 1, 0, 0, 1, 0], np.float32)
    ds, gbt, cfg = _mk(attr, [4], [1], [1], loss=1, min_examples=1, max_depth=2)
    t = gbt.train_tree_on_gradients(g)
    assert len(t) == 3 and t[0]["condition_type"] == 1 and t[0]["cat_mask"][0] == 0b1100
    assert t[0]["num_pos_examples"] == 2 and t[0]["na_value"] == 0
    assert abs(t[0]["split_score"] - 0.125) < 1e-6
    want = O.train_tree(attr, [4], [1], g, None, _oracle_cfg(cfg), feature_type=[1])
    assert not compare_trees(t, want)
    # a constant categorical attribute cannot be split (kInvalidAttribute)
    ds, gbt, cfg = _mk(np.ones((1, 6), np.uint8), [4], [1], [1], loss=1, min_examples=1, max_depth=2)
    assert len(gbt.train_tree_on_gradients(g)) == 1


CASES = [
    dict(n=20000, f_num=3, cats=[5, 17], kw=dict(loss=1, max_depth=6)),
    dict(n=20000, f_num=3, cats=[5, 17], kw=dict(loss=1, max_depth=6, use_hessian_gain=1)),
    dict(n=50000, f_num=4, cats=[3, 40, 256], kw=dict(loss=0, max_depth=8)),
    dict(n=50000, f_num=4, cats=[3, 40, 256], kw=dict(loss=0, max_depth=8, use_hessian_gain=1,
                                                      l2_regularization=0.5, l2_regularization_categorical=1.0)),
    dict(n=30011, f_num=0, cats=[2, 9, 100], kw=dict(loss=0, max_depth=5, min_examples=50,
                                                     in_split_min_examples_check=0)),
    dict(n=9000, f_num=1, cats=[64], kw=dict(loss=1, max_depth=9, min_examples=1)),
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_tree_on_gradients_matches_oracle(case):
    n, kw = case["n"], case["kw"]
    task = "binary" if kw["loss"] == 0 else "regression"
    bins, nb, na, ft, y = synth_mixed(n, case["f_num"], case["cats"], seed=17, task=task)
    ds, gbt, cfg = _mk(bins, nb, na, ft, **kw)
    rng = np.random.default_rng(5)
    if kw["loss"] == 0:
        p = 1 / (1 + np.exp(-rng.normal(size=n)))
        g = ((y == 2) - p).astype(np.float32)
        h = (p * (1 - p)).astype(np.float32)
    else:
        g = (y - y.mean() + 0.1 * rng.normal(size=n)).astype(np.float32)
        h = np.ones(n, np.float32)
    got = gbt.train_tree_on_gradients(g, h)
    if kw.get("use_hessian_gain"):
        O.set_hessian_buckets_double(True)  # see test_gpu_parity.py: the f32 buckets are compared loosely
    try:
        want = O.train_tree(bins, nb, na, g, h, _oracle_cfg(cfg), num_threads=4, feature_type=ft)
    finally:
        O.set_hessian_buckets_double(False)
    errs = compare_trees(got, want)
    assert not errs, errs[:10]
    assert (got["condition_type"] == 1).sum() > 0 and len(got) > 3


@pytest.mark.parametrize("sib", [0, 1])
@pytest.mark.parametrize("loss,hess", [(0, 0), (0, 1), (1, 0)])
def test_gbt_loop_matches_oracle(loss, hess, sib):
    n, iters = 40000, 15
    bins, nb, na, ft, y = synth_mixed(n, 5, [4, 12, 33, 200], seed=23, task="binary" if loss == 0 else "regression")
    ds, gbt, cfg = _mk(bins, nb, na, ft, loss=loss, use_hessian_gain=hess, max_depth=6, num_trees=iters,
                       sibling_subtraction=sib)
    gbt.set_labels(y)
    gbt.train(iters)
    if hess:
        O.set_hessian_buckets_double(True)
    try:
        ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, feature_type=ft)
    finally:
        O.set_hessian_buckets_double(False)
    got = [gbt.get_tree(i) for i in range(iters)]
    t, errs = first_divergence(got, ref["trees"])
    assert t is None, (t, errs[:10])
    assert sum(int((g["condition_type"] == 1).sum()) for g in got) > 10
    for i in range(iters):
        l, _ = gbt.train_loss(i)
        assert abs(l - ref["loss"][i]) <= 1e-5 * abs(ref["loss"][i])
    np.testing.assert_allclose(gbt.get_predictions(), ref["predictions"], rtol=0, atol=2e-5)


def test_sibling_subtraction_and_reruns_are_bit_identical():
    bins, nb, na, ft, y = synth_mixed(60000, 4, [7, 90, 256], seed=31)
    out = []
    for sib in (0, 1, 1):
        ds, gbt, cfg = _mk(bins, nb, na, ft, max_depth=7, num_trees=5, sibling_subtraction=sib)
        gbt.set_labels(y)
        gbt.train(5)
        out.append([gbt.get_tree(i).tobytes() for i in range(5)] + [gbt.get_predictions().tobytes()])
    assert out[0] == out[1] == out[2]


def test_adult_all_features_matches_oracle():
    """BASELINE configs[0] on all 14 Adult features (6 discretized numerical + 8 categorical)."""
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    num, cat = np.load(os.path.join(G, "adult_numerical.npz")), np.load(os.path.join(G, "adult_categorical.npz"))
    from ydf_b200 import dataspec
    cols = []
    for c in ["age", "fnlwgt", "education_num", "capital_gain", "capital_loss", "hours_per_week"]:
        cols.append((dataspec.infer_column(c, num[f"train_{c}"].astype(np.float32)), num[f"train_{c}"].astype(np.float32)))
    for c in ["workclass", "education", "marital_status", "occupation", "relationship", "race", "sex", "native_country"]:
        v = cat[f"strings_{c}"][cat[f"train_{c}"]]
        cols.append((dataspec.infer_categorical_column(c, v), v))
    bins = np.stack([c.encode(v) for c, v in cols])
    nb, na = [c.num_bins for c, _ in cols], [c.na_bin for c, _ in cols]
    ft = [c.feature_type for c, _ in cols]
    y = num["train_income"].astype(np.int32) + 1
    for hess in (0, 1):
        ds, gbt, cfg = _mk(bins, nb, na, ft, num_trees=30, max_depth=6, use_hessian_gain=hess)
        gbt.set_labels(y)
        gbt.train(30)
        O.set_hessian_buckets_double(bool(hess))
        try:
            ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), 30, num_threads=4, feature_type=ft)
        finally:
            O.set_hessian_buckets_double(False)
        got = [gbt.get_tree(i) for i in range(30)]
        # pure nodes: +-1e-16 rounding noise of the reference's variance arithmetic (see prune_noise_splits)
        t, errs = first_divergence(got, ref["trees"], prune_noise=None if hess else 1e-12)
        assert t is None, (hess, t, errs[:10])
        l, acc = gbt.train_loss(29)
        assert abs(l - ref["loss"][29]) <= 1e-5 * ref["loss"][29] and acc > 0.86


def test_learner_with_string_columns(tmp_path):
    rng = np.random.default_rng(3)
    n = 20000
    color = rng.choice(["red", "green", "blue", "teal", "rare"], size=n, p=[0.4, 0.3, 0.2, 0.0998, 0.0002])
    shape = rng.choice(["sq", "tri", ""], size=n, p=[0.5, 0.4, 0.1])  # "" = missing
    x = rng.normal(size=n).astype(np.float32)
    eff = {"red": 1.0, "green": -1.0, "blue": 0.3, "teal": -0.2, "rare": 0.0}
    margin = np.array([eff[c] for c in color]) + 0.8 * x + np.where(shape == "tri", 0.7, 0.0)
    data = {"color": color, "shape": shape, "x": x,
            "y": np.where(margin + 0.3 * rng.normal(size=n) > 0, "yes", "no")}
    learner = ydf_b200.GradientBoostedTreesLearner(
        label="y", discretize_numerical_columns=True, validation_ratio=0.0, early_stopping="NONE",
        num_trees=20, max_depth=4)
    model = learner.train(data)
    spec = model.data_spec
    assert spec.columns[0].vocabulary == ["<OOD>", "red", "green", "blue", "teal"]  # "rare" (< 5) is pruned
    # PYDF leaves most_frequent_value at 0: missing strings are counted with <OOD> (dataspec.py header)
    assert spec.columns[1].na_bin == 0 and spec.columns[1].num_missing > 0
    assert model.evaluate(data)["accuracy"] > 0.9
    assert any((t["condition_type"] == 1).any() for t in model.trees)
    # the saved model directory holds Contains conditions over the dictionary indices
    model.save(str(tmp_path / "m"))
    back = ydf_b200.model_io.read_ydf_model(str(tmp_path / "m"))
    assert back["columns"][1]["vocabulary"]["red"] == 1 and back["columns"][1]["number_of_unique_values"] == 5
    flat = np.concatenate(model.trees)
    cat_nodes = [nd for nd in back["nodes"] if "positive_categories" in nd]
    assert len(cat_nodes) == int((flat["condition_type"] == 1).sum())
    first = flat[flat["condition_type"] == 1][0]
    want = [c for c in range(256) if first["cat_mask"][c >> 5] >> (c & 31) & 1]
    assert cat_nodes[0]["positive_categories"] == want
