"""Builds tests/golden/adult_numerical.npz from the reference's adult_train.csv / adult_test.csv
(UCI Adult): the six numerical columns and the income label.  Categorical features are outside
the accelerated path this round, so BASELINE config 1 is exercised on the
numerical columns.  Run in the authoring container, where /root/reference is mounted."""
import os

import numpy as np
import pandas as pd

D = "/root/reference/yggdrasil_decision_forests/test_data/dataset"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adult_numerical.npz")
NUM = ["age", "fnlwgt", "education_num", "capital_gain", "capital_loss", "hours_per_week"]
out = {}
for split in ("train", "test"):
    df = pd.read_csv(os.path.join(D, f"adult_{split}.csv"))
    for c in NUM:
        out[f"{split}_{c}"] = df[c].to_numpy().astype(np.int32)
    out[f"{split}_income"] = (df["income"].astype(str).str.strip() == ">50K").to_numpy().astype(np.uint8)
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT), {k: v.shape for k, v in out.items() if k.endswith("age")})
