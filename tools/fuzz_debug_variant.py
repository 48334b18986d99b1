import sys
import numpy as np
import ydf_b200
from oracle import oracle as O
from tests.test_gpu_fuzz import draw_variant, _oracle_cfg
from tests.util import synth_mixed

seed = int(sys.argv[1])
c = draw_variant(seed)
print(c)
kw = dict(c["kw"])
task = "regression" if kw["loss"] == 1 else "binary"
bins, nb, na, ft, y = synth_mixed(c["n"], c["f_num"], c["cats"], seed=c["seed"], task=task, bins=c["bins"])
bins = np.concatenate([bins, bins[:1]]); nb, na, ft = np.append(nb, nb[0]), np.append(na, na[0]), np.append(ft, ft[0])
w = np.random.default_rng(c["seed"]).uniform(0.2, 2.5, c["n"]).astype(np.float32) if c["weights"] else None
iters = 2
cfg = ydf_b200.default_config(num_trees=iters, **kw)
O.set_stable_category_sort(True); O.set_hessian_buckets_double(bool(kw["use_hessian_gain"])); O.set_weights(w)
ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), iters, num_threads=4, shuffle_candidates=kw["candidate_shuffle"], feature_type=ft)
O.set_weights(None)
ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
gbt = ydf_b200.Gbt(ds, cfg)
if w is not None: gbt.set_weights(w)
gbt.set_labels(y); gbt.train(iters)
print("tie stats (renamed, unresolved):", gbt.tie_stats())
for t in range(iters):
    a, b = gbt.get_tree(t), ref["trees"][t]
    print("tree", t, "nodes", len(a), len(b))
    for i in range(min(len(a), len(b))):
        if a[i]["feature"] != b[i]["feature"] or a[i]["threshold_bin"] != b[i]["threshold_bin"]:
            print("  first diff at node", i, "depth", a[i]["depth"], "\n   engine", a[i], "\n   oracle", b[i])
            # equal scores among features at this node? show a few nodes before
            break
