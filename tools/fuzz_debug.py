"""Debug helper: one fuzz case (tests/test_gpu_fuzz.py), tree by tree: predictions and node differences."""
import sys
import numpy as np
import ydf_b200
from oracle import oracle as O
from tests.test_gpu_fuzz import draw_case, _oracle_cfg
from tests.util import synth_mixed, compare_trees

seed = int(sys.argv[1])
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = draw_case(seed)
print(c)
kw = c["kw"]
task = "binary" if kw["loss"] == 0 else "regression"
bins, nb, na, ft, y = synth_mixed(c["n"], c["f_num"], c["cats"], seed=c["seed"], task=task, bins=c["bins"])
w = np.random.default_rng(c["seed"]).uniform(0.2, 2.5, c["n"]).astype(np.float32) if c["weights"] else None
O.set_stable_category_sort(True)
O.set_hessian_buckets_double(bool(kw["use_hessian_gain"]))
for k in range(1, iters + 1):
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    cfg = ydf_b200.default_config(num_trees=k, **kw)
    gbt = ydf_b200.Gbt(ds, cfg)
    if w is not None:
        gbt.set_weights(w)
    gbt.set_labels(y)
    gbt.train(k)
    pred = gbt.get_predictions()
    O.set_weights(w)
    ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), k, num_threads=4, feature_type=ft)
    O.set_weights(None)
    d = np.abs(pred - ref["predictions"])
    print("iters", k, "init", gbt.initial_prediction(), "max |dpred|", d.max(), "mean", d.mean(), "rows > 1e-6:", int((d > 1e-6).sum()))
    got, want = gbt.get_tree(k - 1), ref["trees"][k - 1]
    e = compare_trees(got, want)
    print("  nodes", len(got), len(want), "errs", e[:8])
    if len(got) == len(want):
        dl = np.abs(got["leaf_value"] - want["leaf_value"])
        print("  max leaf diff", dl.max(), "at", int(dl.argmax()), got[int(dl.argmax())], want[int(dl.argmax())])
    else:
        for i, (a, b) in enumerate(zip(got, want)):
            if a["feature"] != b["feature"] or a["threshold_bin"] != b["threshold_bin"]:
                print("  first structural difference at", i, a, b)
                break
