"""Builds tests/golden/ydf_adult_gbdt.npz: the files of the reference's golden model
test_data/model/adult_binary_class_gbdt (68 trees, NUMERICAL + CATEGORICAL conditions) as raw bytes, the education_num
column of adult_test as the strings the model's dataspec uses, and the reference's golden predictions for adult_test
(test_data/prediction/adult_test_binary_class_gbdt.csv, written by the reference's own `predict` CLI).
Run in the authoring container, where /root/reference is mounted."""
import os

import numpy as np
import pandas as pd

R = "/root/reference/yggdrasil_decision_forests/test_data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_adult_gbdt.npz")
out = {}
for f in ("header.pb", "data_spec.pb", "gradient_boosted_trees_header.pb", "nodes-00000-of-00001"):
    out["file_" + f] = np.frombuffer(open(os.path.join(R, "model", "adult_binary_class_gbdt", f), "rb").read(), dtype=np.uint8)
pred = pd.read_csv(os.path.join(R, "prediction", "adult_test_binary_class_gbdt.csv"))
out["golden_p_positive"] = pred[">50K"].to_numpy().astype(np.float32)
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT), len(pred))
