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
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_adult_gbdt_v2_head.npz")
np.savez(OUT, num_trees=m["num_trees"], num_log_entries=len(m["training_logs"]),
         initial_prediction=np.float32(m["initial_predictions"][0]), root_num_examples=m["nodes"][0]["n_cond"],
         validation_loss=np.float32(m["validation_loss"]),
         first_training_accuracy=np.float32(log0["training_secondary"]),
         first_validation_accuracy=np.float32(log0["validation_secondary"]),
         last_number_of_trees=m["training_logs"][-1]["number_of_trees"],
         best_validation_loss_entry=int(np.argmin([e["validation_loss"] for e in m["training_logs"]])))
print(OUT, dict(np.load(OUT)))
