"""Randomised differential test: engine == oracle over seeded random configurations of everything the engine accepts
(rows, features, bins, categorical cardinalities, losses, gains, regularisation, depth, min_examples, subsample, weights,
sibling subtraction).  Each case trains 4 trees on both sides and compares every node with the bars of test_gpu_parity.py.

The cases are seeds, not hand-picked inputs: `python tests/test_gpu_fuzz.py 200` runs more of them and prints the failures.
Known, documented sources of legitimate divergence are normalised the same way as in the fixed tests: hessian-gain buckets
are compared against the oracle's exact (double) buckets, pure nodes are pruned on both sides (prune_noise_splits), category
buckets with equal keys use the stable order on both sides, absent categories are dropped from the masks.
"""
import sys

import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees, drop_absent_categories, synth_mixed

pytestmark = pytest.mark.gpu


def _oracle_cfg(cfg):
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    return o


def _l1(v, l1):
    return np.sign(v) * max(0.0, abs(v) - l1)


def prune_noise(tree, kw):
    """Collapses the splits that exist only by rounding noise (pre-order in, pre-order out; returns the tree and the number
    of collapsed splits).  On a PURE node — every row the same gradient and hessian — the exact gain of every split is 0:
    the reference's double arithmetic leaves +-1e-16 relative, the engine's integer sums give exactly 0 (variance gain,
    unweighted) or the same kind of noise, so whether and where such a node "splits" is a coin flip on both sides
    (DESIGN.md §6).  Variance gain: score <= 1e-12; hessian gain: gain over the parent's term <= 1e-6 of that term."""
    out, dropped = [], [0]

    def is_noise(nd):
        if not kw["use_hessian_gain"]:
            return nd["split_score"] <= 1e-12
        l2 = kw["l2_regularization_categorical"] if nd["condition_type"] == 1 else kw["l2_regularization"]
        parent = _l1(float(nd["stat"][0]), kw["l1_regularization"]) ** 2 / (float(nd["stat"][1]) + l2)
        gain = float(nd["split_score"]) - (0.0 if kw["hessian_split_score_subtract_parent"] else parent)
        return gain <= 1e-6 * max(parent, 1e-30)

    def walk(i):
        nd = tree[i].copy()
        me = len(out)
        out.append(nd)
        if nd["feature"] >= 0 and is_noise(nd):
            dropped[0] += 1
            for k, v in (("feature", -1), ("threshold_bin", 0), ("na_value", 0), ("split_score", 0.0),
                         ("num_pos_examples", 0), ("condition_type", 0), ("neg_child", -1), ("pos_child", -1)):
                nd[k] = v
            nd["cat_mask"] = 0
            if "threshold_value" in tree.dtype.names:
                nd["threshold_value"] = 0
        elif nd["feature"] >= 0:
            nd["neg_child"] = len(out)
            walk(int(tree[i]["neg_child"]))
            nd["pos_child"] = len(out)
            walk(int(tree[i]["pos_child"]))
        out[me] = nd

    walk(0)
    return np.array(out, dtype=tree.dtype), dropped[0]


def draw_case(seed):
    r = np.random.default_rng(1000 + seed)
    n = int(r.choice([300, 2000, 9000, 30000]))
    f_num = int(r.integers(0, 9))
    n_cat = int(r.integers(0, 4)) if f_num > 0 else int(r.integers(1, 4))
    cats = [int(r.choice([2, 3, 7, 24, 25, 60, 255, 256])) for _ in range(n_cat)]
    loss = int(r.integers(0, 2))
    hess = int(r.integers(0, 2))
    kw = dict(loss=loss, use_hessian_gain=hess, max_depth=int(r.integers(2, 10)), min_examples=int(r.choice([1, 5, 40])),
              in_split_min_examples_check=int(r.integers(0, 2)), l1_regularization=float(r.choice([0.0, 0.0, 0.3])),
              l2_regularization=float(r.choice([0.0, 0.0, 1.0])), l2_regularization_categorical=float(r.choice([0.0, 1.0])),
              shrinkage=float(r.choice([0.1, 0.3])), sibling_subtraction=int(r.integers(0, 2)),
              subsample=float(r.choice([1.0, 1.0, 0.6])),
              hessian_split_score_subtract_parent=int(hess and r.integers(0, 2)))
    weights = bool(r.integers(0, 2)) and not hess
    return dict(seed=seed, n=n, f_num=f_num, cats=cats, bins=int(r.choice([4, 16, 64, 255])), weights=weights, kw=kw)


def run_case(c, iters=4):
    kw = c["kw"]
    task = "binary" if kw["loss"] == 0 else "regression"
    bins, nb, na, ft, y = synth_mixed(c["n"], c["f_num"], c["cats"], seed=c["seed"], task=task, bins=c["bins"])
    w = np.random.default_rng(c["seed"]).uniform(0.2, 2.5, c["n"]).astype(np.float32) if c["weights"] else None
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    cfg = ydf_b200.default_config(num_trees=iters, **kw)
    gbt = ydf_b200.Gbt(ds, cfg)
    if w is not None:
        gbt.set_weights(w)
    gbt.set_labels(y)
    gbt.train(iters)
    got = [gbt.get_tree(i) for i in range(iters)]
    got_loss = [gbt.train_loss(i)[0] for i in range(iters)]
    pred = gbt.get_predictions()
    gbt.close()
    ds.close()
    O.set_stable_category_sort(True)
    O.set_hessian_buckets_double(bool(kw["use_hessian_gain"]))
    O.set_weights(w)
    try:
        ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, feature_type=ft)
    finally:
        O.set_weights(None)
        O.set_hessian_buckets_double(False)
        O.set_stable_category_sort(False)
    errs = []
    rtol = 1e-5
    noise = 0
    # the rows each tree was trained on (SampleTrainingExamples: one engine word per row and iteration, no other consumer here)
    stream = O.Rng(cfg.random_seed) if kw["subsample"] < 1.0 else None
    for i in range(iters):
        a, na_ = prune_noise(got[i], kw)
        b, nb_ = prune_noise(ref["trees"][i], kw)
        trained = bins
        if stream is not None:
            words = np.array([stream.next() for _ in range(c["n"])], dtype=np.uint64)
            trained = bins[:, (words.astype(np.float32) / np.float32(4294967296.0)) < np.float32(kw["subsample"])]
        if c["cats"]:   # categories without a TRAINING row in the node have empty buckets; where those go is not compared
            a, b = drop_absent_categories(a, trained), drop_absent_categories(b, trained)
        # (sums of gradients: predictions agree to 2e-5, so gradients to ~5e-6 per row in the worst case)
        e = compare_trees(a, b, score_rtol=rtol, stat_atol_per_row=2e-7)
        if e:
            errs.append((i, e[:4]))
            break   # later trees inherit the divergence
        noise += na_ + nb_
        if noise:
            break   # noise splits route rows to children whose leaves differ as soon as l2 > 0: nothing to compare after them
        if abs(got_loss[i] - ref["loss"][i]) > 1e-5 * abs(ref["loss"][i]):
            errs.append((i, f"loss {got_loss[i]} vs {ref['loss'][i]}"))
    # (with a row sample, rows outside it may carry categories no training row of their node has: their side is not defined)
    if not errs and not noise and not (stream is not None and c["cats"]) and np.abs(pred - ref["predictions"]).max() > 2e-5:
        errs.append(("predictions", float(np.abs(pred - ref["predictions"]).max())))
    return errs


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_matches_oracle(seed):
    c = draw_case(seed)
    errs = run_case(c)
    assert not errs, (c, errs)


if __name__ == "__main__":
    bad = 0
    for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 100):
        c = draw_case(s)
        try:
            e = run_case(c)
        except Exception as ex:   # noqa: BLE001
            e = [("exception", repr(ex))]
        if e:
            bad += 1
            print("FAIL", c, e, flush=True)
    print("failures:", bad)
