"""Builds tests/golden/adult_cxx_test_folds.npz: the train / test folds of the reference's C++ acceptance tests on Adult
(`GradientBoostedTreesOnAdult`, learner/gradient_boosted_trees/gradient_boosted_trees_test.cc:553-569): adult.csv,
`dataset_sampling_ = 0.2`, `split_train_ratio_ = 0.5`, drawn by utils/test_utils.cc's DeterministicBinomial (:100-127,
:505-600), encoded with the dataspec the test infers from ALL rows of the file with
`detect_numerical_as_discretized_numerical` (255-bin boundaries, C++ dictionary rule).  Only the 6513 sampled rows are
stored, as bucket / dictionary codes.  Run in the authoring container (/root/reference mounted)."""
import csv
import os
import sys
from collections import Counter

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import dataspec  # noqa: E402

rows = list(csv.DictReader(open("/root/reference/yggdrasil_decision_forests/test_data/dataset/adult.csv")))
names = list(rows[0].keys())
NUM = ["age", "fnlwgt", "education_num", "capital_gain", "capital_loss", "hours_per_week"]
cols, enc, cols16, enc16 = [], [], [], []
for name in names[:-1]:
    raw = [r[name] for r in rows]
    if name in NUM:
        v = np.array([float(x) if x != "" else np.nan for x in raw], np.float32)
        c = dataspec.infer_column(name, v)
        c16 = dataspec.infer_column(name, v, maximum_num_bins=16)   # BaseAggressiveDiscretizedNumerical (:1208-1229)
    else:
        v = np.array(raw, dtype=object)
        c = c16 = dataspec.infer_categorical_column(name, v, front_end=dataspec.FRONT_END_CPP)
    cols.append(c)
    enc.append(c.encode(v))
    cols16.append(c16)
    enc16.append(c16.encode(v))
bins, bins16 = np.stack(enc), np.stack(enc16)
count = Counter(r["income"] for r in rows)
classes = sorted(count, key=lambda k: (count[k], k), reverse=True)   # C++ dictionary rule: ["<=50K", ">50K"]
y = np.array([classes.index(r["income"]) + 1 for r in rows], np.int32)


class DeterministicBinomial:
    def __init__(self, rate):
        self.rate, self.pos, self.total = np.float32(rate), 0, 0

    def sample(self):
        if self.total == 0:
            self.total += 1
            if self.rate == 1:
                self.pos += 1
                return True
            return False
        if self.pos > float(self.rate * np.float32(self.total)):
            self.total += 1
            return False
        self.pos += 1
        self.total += 1
        return True


sampling, split = DeterministicBinomial(0.2), DeterministicBinomial(0.5)
train, test = [], []
for i in range(len(rows)):
    if sampling.sample():
        (train if split.sample() else test).append(i)
train, test = np.array(train), np.array(test)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adult_cxx_test_folds.npz")
np.savez_compressed(OUT, feature_names=np.array(names[:-1]), num_bins=np.array([c.num_bins for c in cols], np.int32),
                    na_bin=np.array([c.na_bin for c in cols], np.int32),
                    feature_type=np.array([int(c.feature_type) for c in cols], np.int32), classes=np.array(classes),
                    train_rows=train.astype(np.int32), test_rows=test.astype(np.int32),
                    train_bins=bins[:, train], test_bins=bins[:, test], train_labels=y[train], test_labels=y[test],
                    num_bins16=np.array([c.num_bins for c in cols16], np.int32),
                    na_bin16=np.array([c.na_bin for c in cols16], np.int32),
                    train_bins16=bins16[:, train], test_bins16=bins16[:, test])
print(OUT, os.path.getsize(OUT), len(train), len(test))
