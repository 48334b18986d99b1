"""Builds tests/golden/ydf_iris_gbdt_v2.npz: the files of the reference's golden model
test_data/model/iris_multi_class_gbdt_v2 (a default PYDF GradientBoostedTreesLearner on iris.csv: multinomial loss,
3 trees per iteration, gzip blob sequence) as raw bytes.  Run in the authoring container (/root/reference mounted)."""
import os

import numpy as np

R = "/root/reference/yggdrasil_decision_forests/test_data/model/iris_multi_class_gbdt_v2"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_iris_gbdt_v2.npz")
out = {"file_" + f: np.frombuffer(open(os.path.join(R, f), "rb").read(), dtype=np.uint8)
       for f in ("header.pb", "data_spec.pb", "gradient_boosted_trees_header.pb", "nodes-00000-of-00001")}
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT))
