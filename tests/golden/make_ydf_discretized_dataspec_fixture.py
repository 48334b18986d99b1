"""Builds tests/golden/ydf_adult_discretized_dataspec.npz: the DISCRETIZED_NUMERICAL boundaries the REFERENCE computed on
adult_train.csv (GenDiscretizedBoundaries, 255 bins), as stored in the dataspec of its golden model
test_data/model/adult_binary_class_rf_discret_numerical.  Run in the authoring container (/root/reference mounted)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

cols, n = model_io.read_data_spec("/root/reference/yggdrasil_decision_forests/test_data/model/"
                                  "adult_binary_class_rf_discret_numerical/data_spec.pb")
out = {"created_num_rows": np.int64(n)}
for c in cols:
    if "boundaries" in c:
        out[f"boundaries_{c['name']}"] = c["boundaries"].astype(np.float32)
        out[f"mean_{c['name']}"] = np.float64(c["mean"])
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ydf_adult_discretized_dataspec.npz")
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT), sorted(out))
