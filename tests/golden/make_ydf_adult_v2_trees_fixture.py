"""Builds tests/golden/ydf_adult_gbdt_v2_trees.npz: every node of all 163 trees of the reference's golden model
test_data/model/adult_binary_class_gbdt_v2 (PYDF defaults on adult_train.csv; see make_ydf_adult_v2_head_fixture.py),
in the model's pre-order (node, negative subtree, positive subtree), the dictionaries and per-column
most_frequent_value of its PYDF-made dataspec, and its training log.  Run in the authoring container
(/root/reference mounted)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

m = model_io.read_ydf_model("/root/reference/yggdrasil_decision_forests/test_data/model/adult_binary_class_gbdt_v2")
nodes, cols = m["nodes"], m["columns"]


def skip(i):
    return i + 1 if "attribute" not in nodes[i] else skip(skip(i + 1))


tree_first, i = [], 0
while i < len(nodes):
    tree_first.append(i)
    i = skip(i)
assert len(tree_first) == m["num_trees"] == 163
tree0 = nodes
K = len(tree0)
feature = np.full(K, -1, np.int32)
threshold = np.full(K, np.nan, np.float32)
mask = np.zeros(K, np.uint64)
n_pos = np.zeros(K, np.int64)
score = np.zeros(K, np.float32)
na_value = np.zeros(K, bool)
for i, nd in enumerate(tree0):
    if "attribute" not in nd:
        continue
    feature[i], n_pos[i], score[i], na_value[i] = nd["attribute"], nd["n_pos"], nd["split_score"], nd["na_value"]
    assert nd["n_cond"] == nd["n"]
    if "positive_categories" in nd:
        assert max(nd["positive_categories"]) < 64  # native_country: 41 values
        mask[i] = sum(1 << c for c in nd["positive_categories"])
    else:
        threshold[i] = nd["higher_threshold"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_adult_gbdt_v2_trees.npz")
logs = m["training_logs"]
vocab = {f"vocabulary_{c['name']}": np.array(sorted(c["vocabulary"], key=c["vocabulary"].get)) for c in cols if c["type"] == 4}
np.savez_compressed(OUT, tree_first=np.array(tree_first, np.int32), initial_prediction=np.float32(m["initial_predictions"][0]),
         log_num_trees=np.array([e["number_of_trees"] for e in logs], np.int32),
         log_training_loss=np.array([e["training_loss"] for e in logs], np.float32),
         log_training_accuracy=np.array([e["training_secondary"] for e in logs], np.float32),
         log_validation_loss=np.array([e["validation_loss"] for e in logs], np.float32),
         log_validation_accuracy=np.array([e["validation_secondary"] for e in logs], np.float32),
         **vocab, column_names=np.array([c["name"] for c in cols]), column_types=np.array([c["type"] for c in cols], np.int32),
         most_frequent_value=np.array([c.get("most_frequent_value", -1) for c in cols], np.int32),
         feature=feature, threshold=threshold, positive_mask=mask, n=np.array([nd["n"] for nd in tree0], np.int64),
         n_pos=n_pos, split_score=score, na_value=na_value,
         value=np.array([nd["top_value"] for nd in tree0], np.float32))
z = np.load(OUT)
print(OUT, K, "nodes;", {k: z[k].shape for k in z.files})
