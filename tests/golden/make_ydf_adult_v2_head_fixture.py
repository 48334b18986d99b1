"""Builds tests/golden/ydf_adult_gbdt_v2_head.npz: a few numbers of the reference's golden model
test_data/model/adult_binary_class_gbdt_v2 = `ydf.GradientBoostedTreesLearner(label="income").train(adult_train.csv)`
with every hyper-parameter at its default (so: 10 % validation hold-out drawn from mt19937(123456), early stopping).
Run in the authoring container (/root/reference mounted)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

m = model_io.read_ydf_model("/root/reference/yggdrasil_decision_forests/test_data/model/adult_binary_class_gbdt_v2")
log0 = m["training_logs"][0]
root = m["nodes"][0]
root_col = m["columns"][root["attribute"]]


def skip(i):  # index after the subtree rooted at i (pre-order: node, negative subtree, positive subtree)
    return i + 1 if "attribute" not in m["nodes"][i] else skip(skip(i + 1))


child = m["nodes"][skip(1)]          # positive child of the root: the rows with relationship in the positive set
child_col = m["columns"][child["attribute"]]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_adult_gbdt_v2_head.npz")
np.savez(OUT, num_trees=m["num_trees"], num_log_entries=len(m["training_logs"]),
         initial_prediction=np.float32(m["initial_predictions"][0]), root_num_examples=m["nodes"][0]["n_cond"],
         validation_loss=np.float32(m["validation_loss"]),
         first_training_accuracy=np.float32(log0["training_secondary"]),
         first_validation_accuracy=np.float32(log0["validation_secondary"]),
         last_number_of_trees=m["training_logs"][-1]["number_of_trees"],
         root_feature=root_col["name"], root_positive_categories=np.array(root["positive_categories"], np.int32),
         root_num_pos=root["n_pos"], root_split_score=np.float32(root["split_score"]), root_na_value=bool(root["na_value"]),
         root_vocabulary=np.array(sorted(root_col["vocabulary"], key=root_col["vocabulary"].get)),
         child_feature=child_col["name"], child_positive_categories=np.array(child["positive_categories"], np.int32),
         child_num_examples=child["n_cond"], child_num_pos=child["n_pos"], child_split_score=np.float32(child["split_score"]),
         child_na_value=bool(child["na_value"]),
         best_validation_loss_entry=int(np.argmin([e["validation_loss"] for e in m["training_logs"]])))
print(OUT, dict(np.load(OUT)))
