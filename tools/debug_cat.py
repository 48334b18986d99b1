"""Debug aid: dumps the GPU trees of two categorical parity cases to gpurun_out/cat_debug.npz."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ydf_b200
from tests.util import synth_mixed
out = {}
bins, nb, na, ft, y = synth_mixed(40000, 5, [4, 12, 33, 200], seed=23, task="binary")
ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(loss=0, use_hessian_gain=1, max_depth=6, num_trees=2, sibling_subtraction=0))
gbt.set_labels(y); gbt.train(2)
out["loop_t0"] = gbt.get_tree(0); out["loop_t1"] = gbt.get_tree(1)
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
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
ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=1, max_depth=6))
gbt.set_labels(y); gbt.train(1)
out["adult_t0"] = gbt.get_tree(0)
np.savez("gpurun_out/cat_debug.npz", **out)
print("ok")
