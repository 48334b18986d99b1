"""Builds tests/golden/adult_categorical.npz: the eight string columns of the reference's
adult_train.csv / adult_test.csv (UCI Adult) as integer codes into a per-column list of raw strings
("" = missing), plus the dictionary the reference itself inferred for each column, read from the
data_spec.pb of its golden model test_data/model/adult_binary_class_gbdt (trained on adult_train.csv
with the default guide: min_vocab_frequency 5, max_vocab_count 2000).  The latter pins
dataspec.infer_categorical_column (dataset/data_spec_inference.cc:277-441).
Run in the authoring container, where /root/reference is mounted."""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

R = "/root/reference/yggdrasil_decision_forests/test_data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adult_categorical.npz")
CAT = ["workclass", "education", "marital_status", "occupation", "relationship", "race", "sex", "native_country"]
out = {}
frames = {s: pd.read_csv(os.path.join(R, "dataset", f"adult_{s}.csv"), dtype=str, keep_default_na=False)
          for s in ("train", "test")}
for c in CAT:
    raw = sorted(set(frames["train"][c]) | set(frames["test"][c]))
    code = {k: i for i, k in enumerate(raw)}
    out[f"strings_{c}"] = np.array(raw, dtype=str)
    for s in ("train", "test"):
        out[f"{s}_{c}"] = frames[s][c].map(code).to_numpy().astype(np.uint8)
model = model_io.read_ydf_model(os.path.join(R, "model", "adult_binary_class_gbdt"))
for col in model["columns"]:
    if col["name"] in CAT:
        vocab = col["vocabulary"]
        keys = sorted(vocab, key=lambda k: vocab[k])
        assert [vocab[k] for k in keys] == list(range(len(keys)))
        out[f"ref_vocab_{col['name']}"] = np.array(keys, dtype=str)
        out[f"ref_mfv_{col['name']}"] = np.int32(col["most_frequent_value"])
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT))
