"""Builds tests/golden/ydf_8bits_gbdt.npz: the files of the reference's golden model
test_data/model/8bits_numerical_binary_class_gbdt — a GBT the REFERENCE trained on DISCRETIZED_NUMERICAL features
(binomial loss, variance gain, 10 trees) — as raw bytes.  Its training set is not in the reference tree, but every node
stores the statistics the splitter worked with, which is what tests/test_oracle_kat.py checks the formulas against.
Run in the authoring container, where /root/reference is mounted."""
import os

import numpy as np

R = "/root/reference/yggdrasil_decision_forests/test_data/model/8bits_numerical_binary_class_gbdt"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_8bits_gbdt.npz")
out = {"file_" + f: np.frombuffer(open(os.path.join(R, f), "rb").read(), dtype=np.uint8)
       for f in ("header.pb", "data_spec.pb", "gradient_boosted_trees_header.pb", "nodes-00000-of-00001")}
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT))
