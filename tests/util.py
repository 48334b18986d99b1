"""Shared helpers for the parity tests: synthetic data (SURVEY.md §8d generator), tree comparison."""
import numpy as np


def synth(n, f, seed=1234, informative=None, task="binary", bins=255, sample=100000):
    """X ~ N(0,1) [f, n]; y = 1[sum_j w_j x_j + 0.5 x0 x1 + 0.3 sin(3 x2) + eps > 0] (binary) or the
    margin itself (regression).  Returns (bins uint8 [f, n], num_bins, na_bin, labels)."""
    import ydf_b200
    rng = np.random.default_rng(seed)
    informative = informative or min(10, f)
    w = rng.normal(size=informative)
    margin = np.zeros(n, dtype=np.float64)
    out = np.empty((f, n), dtype=np.uint8)
    nb, na = [], []
    cols = {}
    for j in range(f):
        x = rng.normal(size=n).astype(np.float32)
        if j < informative:
            margin += w[j] * x
        if j < 3:
            cols[j] = x
        b, mean = ydf_b200.discretize_boundaries(x[:sample], bins, 3)
        nbin = len(b) + 1
        nab = int(np.searchsorted(b, np.float32(mean), side="right"))
        out[j] = ydf_b200.discretize_encode(x, b, nab)
        nb.append(nbin)
        na.append(nab)
    if f >= 3:
        margin += 0.5 * cols[0] * cols[1] + 0.3 * np.sin(3 * cols[2])
    margin += rng.normal(scale=0.5, size=n)
    if task == "binary":
        y = (margin > 0).astype(np.int32) + 1
    else:
        y = margin.astype(np.float32)
    return out, np.array(nb, np.int32), np.array(na, np.int32), y


def compare_trees(a, b, score_rtol=1e-5, leaf_atol=1e-5, stat_rtol=1e-6, stat_atol_per_row=2e-8):
    """a, b: node arrays (pre-order).  Returns a list of mismatch strings (empty = parity)."""
    errs = []
    if len(a) != len(b):
        return [f"node count {len(a)} != {len(b)}"]
    for i, (x, y) in enumerate(zip(a, b)):
        for k in ("feature", "threshold_bin", "na_value", "depth", "neg_child", "pos_child",
                  "num_examples", "num_pos_examples"):
            if x[k] != y[k]:
                errs.append(f"node {i}: {k} {x[k]} != {y[k]}")
        if abs(float(x["split_score"]) - float(y["split_score"])) > score_rtol * max(1e-30, abs(float(y["split_score"]))):
            errs.append(f"node {i}: split_score {x['split_score']} vs {y['split_score']}")
        if abs(float(x["leaf_value"]) - float(y["leaf_value"])) > leaf_atol:
            errs.append(f"node {i}: leaf_value {x['leaf_value']} vs {y['leaf_value']}")
        for s in range(3):
            d = abs(float(x["stat"][s]) - float(y["stat"][s]))
            # sums over n rows: 1e-6 relative, or 2e-8 per row (fixed-point resolution / 1-ulp
            # differences of the per-row gradients between devices)
            if d > max(stat_rtol * abs(float(y["stat"][s])), stat_atol_per_row * float(y["num_examples"]), 1e-9):
                errs.append(f"node {i}: stat[{s}] {x['stat'][s]} vs {y['stat'][s]}")
    return errs


def first_divergence(trees_a, trees_b, **kw):
    for t, (a, b) in enumerate(zip(trees_a, trees_b)):
        e = compare_trees(a, b, **kw)
        if e:
            return t, e
    return None, []
