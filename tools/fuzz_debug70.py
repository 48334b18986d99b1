import numpy as np
import ydf_b200
from oracle import oracle as O
from tests.test_gpu_fuzz import draw_case, _oracle_cfg
from tests.util import synth_mixed

c = draw_case(70)
kw = c["kw"]
bins, nb, na, ft, y = synth_mixed(c["n"], c["f_num"], c["cats"], seed=70, task="binary", bins=c["bins"])
N = c["n"]
cfg = ydf_b200.default_config(num_trees=2, **kw)
ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
gbt = ydf_b200.Gbt(ds, cfg); gbt.set_labels(y); gbt.train(2)
O.set_stable_category_sort(True); O.set_hessian_buckets_double(True)
ref = O.gbt_train(bins, nb, na, y, _oracle_cfg(cfg), 2, num_threads=4, feature_type=ft)
got, want = gbt.get_tree(1), ref["trees"][1]
# the sample of iteration 1 (second block of N draws)
r = O.Rng(cfg.random_seed)
words = np.array([r.next() for _ in range(2 * N)], dtype=np.uint64)
u = (words.astype(np.float32) / np.float32(4294967296.0))
sel = u[N:] < np.float32(kw["subsample"])
print("selected", sel.sum(), "root n", got[0]["num_examples"], want[0]["num_examples"])
a, b = got[0], want[0]
print("feature", a["feature"], b["feature"], "score", a["split_score"], b["split_score"], "npos", a["num_pos_examples"], b["num_pos_examples"])
f = int(a["feature"])
diff = []
for cat in range(256):
    ba = (int(a["cat_mask"][cat >> 5]) >> (cat & 31)) & 1
    bb = (int(b["cat_mask"][cat >> 5]) >> (cat & 31)) & 1
    if ba != bb:
        rows = np.nonzero(bins[f] == cat)[0]
        diff.append((cat, ba, bb, len(rows), int(sel[rows].sum())))
print("differing categories (cat, engine, oracle, rows, sampled rows):", diff)
print("na_bin", na[f], "num_bins", nb[f])
