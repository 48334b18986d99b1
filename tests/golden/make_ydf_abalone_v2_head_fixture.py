"""Builds tests/golden/ydf_abalone_gbdt_v2_head.npz: the Rings column of the reference's abalone.csv and a few numbers of
its golden model test_data/model/abalone_regression_gbdt_v2 = `ydf.GradientBoostedTreesLearner(label="Rings",
task=REGRESSION).train(abalone.csv)` with default hyper-parameters.  Run in the authoring container."""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

R = "/root/reference/yggdrasil_decision_forests/test_data"
m = model_io.read_ydf_model(os.path.join(R, "model", "abalone_regression_gbdt_v2"))
rings = pd.read_csv(os.path.join(R, "dataset", "abalone.csv"))["Rings"].to_numpy().astype(np.int16)
root = m["nodes"][0]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_abalone_gbdt_v2_head.npz")
np.savez_compressed(OUT, rings=rings, num_trees=m["num_trees"], num_log_entries=len(m["training_logs"]), loss=m["loss"],
                    initial_prediction=np.float32(m["initial_predictions"][0]), root_num_examples=root["n_cond"],
                    root_sum=root["distribution"][0], root_sum_squares=root["distribution"][1],
                    validation_loss=np.float32(m["validation_loss"]),
                    best_validation_loss_entry=int(np.argmin([e["validation_loss"] for e in m["training_logs"]])))
print(OUT, os.path.getsize(OUT))
