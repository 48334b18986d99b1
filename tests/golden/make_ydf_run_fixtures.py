"""Builds tests/golden/ydf_run_*.npz: complete training runs of the reference, as stored in its golden models
test_data/model/{adult_binary_class,iris_multi_class,abalone_regression}_gbdt_v2 (PYDF) and in the goldens of its C++
tests test_data/golden/gbt_{iris,iris_hessian,adult_subsampling,abalone} (`ExpectEqualGoldenModel`).  The PYDF ones:  Each is
`ydf.GradientBoostedTreesLearner(label=...).train(<csv>)` with every hyper-parameter at its default (10 % validation
hold-out drawn from mt19937(123456), early stopping, exact numerical splits, Contains conditions).  A fixture holds
every node of every tree in the model's pre-order (node, negative subtree, positive subtree), the dictionaries and
most_frequent_value of the PYDF-made dataspec, the training log, and — for the two small datasets — the CSV columns
(Adult's columns are in adult_numerical.npz / adult_categorical.npz).  Run in the authoring container
(/root/reference mounted); consumed by tests/reference_replay.py."""
import csv
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ydf_b200 import model_io  # noqa: E402

R = "/root/reference/yggdrasil_decision_forests/test_data"
HERE = os.path.dirname(os.path.abspath(__file__))


def node_count(nd):
    """Training rows in the node: NodeRegressorOutput.distribution.count, or sum_weights of the hessian statistics."""
    if nd.get("n") is not None:
        return int(nd["n"])
    if "hessian_stats" in nd:
        return int(round(nd["hessian_stats"][2]))
    return int(nd["n_cond"])


class DeterministicBinomial:
    """utils/test_utils.cc:100-127: the tester's sampler (first draw false unless rate == 1, then keeps the running
    share of positives at the rate; the product rate * total is taken in float)."""

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


def tester_folds(n_rows, dataset_sampling):
    """(training fold, test fold) of utils::TrainAndTestTester (test_utils.cc:505-600): down-sampling, then a 50 % split."""
    sampling, split = DeterministicBinomial(dataset_sampling), DeterministicBinomial(0.5)
    train, test = [], []
    for i in range(n_rows):
        if sampling.sample():
            (train if split.sample() else test).append(i)
    return train, test


def build(model, out, csv_name=None, cxx_test=None):
    """cxx_test: the model is a golden of the reference's C++ tests (test_data/golden/<model>), trained by
    utils::TrainAndTestTester on its training fold with one thread; dict(dataset_sampling, use_hessian_gain, subsample,
    max_depth).  Only the training fold's rows are stored."""
    m = model_io.read_ydf_model(f"{R}/{'golden' if cxx_test is not None else 'model'}/{model}")
    nodes, cols = m["nodes"], m["columns"]

    def skip(i):
        return i + 1 if "attribute" not in nodes[i] else skip(skip(i + 1))

    tree_first, i = [], 0
    while i < len(nodes):
        tree_first.append(i)
        i = skip(i)
    assert len(tree_first) == m["num_trees"]
    K = len(nodes)
    feature = np.full(K, -1, np.int32)
    threshold = np.full(K, np.nan, np.float32)
    mask = np.zeros(K, np.uint64)
    n_pos = np.zeros(K, np.int64)
    score = np.zeros(K, np.float32)
    na_value = np.zeros(K, bool)
    for i, nd in enumerate(nodes):
        if "attribute" not in nd:
            continue
        feature[i], n_pos[i], score[i], na_value[i] = nd["attribute"], nd["n_pos"], nd["split_score"], nd["na_value"]
        assert nd["n"] is None or nd["n_cond"] == nd["n"]   # hessian-gain nodes carry (sum g, sum h, n) instead of n
        if "positive_categories" in nd:
            assert max(nd["positive_categories"]) < 64  # Adult's native_country: 41 values
            mask[i] = sum(1 << c for c in nd["positive_categories"])
        else:
            threshold[i] = nd["higher_threshold"]
    logs = m["training_logs"]
    extra = {f"vocabulary_{c['name']}": np.array(sorted(c["vocabulary"], key=c["vocabulary"].get))
             for c in cols if c["type"] == 4}
    run = dict(front_end="pydf", single_thread=0, use_hessian_gain=0, subsample=1.0, max_depth=6)
    if csv_name:
        rows = list(csv.DictReader(open(f"{R}/dataset/{csv_name}")))
        if cxx_test is not None:
            run.update(front_end="cpp", single_thread=1, use_hessian_gain=int(cxx_test.get("use_hessian_gain", 0)),
                       subsample=float(cxx_test.get("subsample", 1.0)), max_depth=int(cxx_test.get("max_depth", 6)))
            fold, test_fold = tester_folds(len(rows), cxx_test.get("dataset_sampling", 1.0))
            extra["fold_rows"] = np.array(fold, np.int32)
            extra["csv_num_rows"] = np.int64(len(rows))
            if cxx_test.get("with_test_fold"):   # the fold the tester evaluates on (golden metric values)
                for c in cols:
                    v = [rows[i][c["name"]] for i in test_fold]
                    extra[f"test_{c['name']}"] = (np.array(v) if c["type"] == 4
                                                  else np.array([float(x) if x != "" else np.nan for x in v], np.float32))
            rows = [rows[i] for i in fold]
        for c in cols:
            v = [r[c["name"]] for r in rows]
            extra[f"data_{c['name']}"] = (np.array(v) if c["type"] == 4
                                          else np.array([float(x) if x != "" else np.nan for x in v], np.float32))
    path = os.path.join(HERE, out)
    np.savez_compressed(
        path, run_front_end=run["front_end"], run_single_thread=run["single_thread"],
        run_use_hessian_gain=run["use_hessian_gain"], run_subsample=np.float32(run["subsample"]), run_max_depth=run["max_depth"],
        loss=m["loss"], task=m["task"], label_col_idx=m["label_col_idx"], num_trees_per_iter=m["num_trees_per_iter"],
        tree_first=np.array(tree_first, np.int32), initial_predictions=np.array(m["initial_predictions"], np.float32),
        validation_loss=np.float32(m["validation_loss"]),
        log_num_trees=np.array([e["number_of_trees"] for e in logs], np.int32),
        log_training_loss=np.array([e["training_loss"] for e in logs], np.float32),
        log_training_secondary=np.array([e["training_secondary"] for e in logs], np.float32),
        log_validation_loss=np.array([e["validation_loss"] for e in logs], np.float32),
        log_validation_secondary=np.array([e["validation_secondary"] for e in logs], np.float32),
        column_names=np.array([c["name"] for c in cols]), column_types=np.array([c["type"] for c in cols], np.int32),
        most_frequent_value=np.array([c.get("most_frequent_value", -1) for c in cols], np.int32),
        feature=feature, threshold=threshold, positive_mask=mask, n=np.array([node_count(nd) for nd in nodes], np.int64),
        n_pos=n_pos, split_score=score, na_value=na_value, value=np.array([nd["top_value"] for nd in nodes], np.float32),
        **extra)
    print(path, os.path.getsize(path), "bytes;", m["num_trees"], "trees,", K, "nodes,", len(logs), "log entries")


build("adult_binary_class_gbdt_v2", "ydf_run_adult_v2.npz")
build("iris_multi_class_gbdt_v2", "ydf_run_iris_v2.npz", "iris.csv")
build("abalone_regression_gbdt_v2", "ydf_run_abalone_v2.npz", "abalone.csv")
# goldens of the reference's C++ tests (gradient_boosted_trees_test.cc): one thread, C++ dataspec inference, tester folds
build("gbt_iris", "ydf_run_cxx_iris.npz", "iris.csv", cxx_test=dict())                                        # :1737-1743
build("gbt_iris_hessian", "ydf_run_cxx_iris_hessian.npz", "iris.csv", cxx_test=dict(use_hessian_gain=1))      # :1752-1761
build("gbt_adult_subsampling", "ydf_run_cxx_adult_subsampling.npz", "adult.csv",                              # :592-636
      cxx_test=dict(dataset_sampling=0.2, subsample=0.9, max_depth=4, with_test_fold=True))
build("gbt_abalone", "ydf_run_cxx_abalone.npz", "abalone.csv", cxx_test=dict())                               # :1630-1635


def build_header_only(model, out):
    """gbt_adult_base (GradientBoostedTreesOnAdult.Base, :571-590) stores its nodes in the reference's older record format,
    which this repo does not read; its header still holds the complete training log."""
    from ydf_b200.model_io import pb_decode, _one
    g = pb_decode(open(f"{R}/golden/{model}/gradient_boosted_trees_header.pb", "rb").read())
    logs = [dict((("number_of_trees", "training_loss", "training_secondary", "validation_loss", "validation_secondary")[f - 1], v)
                 for f, _, v in pb_decode(e) if 1 <= f <= 5) for ff, _, e in pb_decode(_one(g, 8, b"")) if ff == 1]
    path = os.path.join(HERE, out)
    np.savez_compressed(
        path, num_trees=_one(g, 2), validation_loss=np.float32(_one(g, 6)),
        initial_predictions=np.array([v for f, _, v in g if f == 4], np.float32),
        log_num_trees=np.array([e["number_of_trees"] for e in logs], np.int32),
        log_training_loss=np.array([e["training_loss"] for e in logs], np.float32),
        log_training_secondary=np.array([e["training_secondary"] for e in logs], np.float32),
        log_validation_loss=np.array([e["validation_loss"] for e in logs], np.float32),
        log_validation_secondary=np.array([e["validation_secondary"] for e in logs], np.float32))
    print(path, os.path.getsize(path), "bytes;", len(logs), "log entries")


build_header_only("gbt_adult_base", "ydf_run_cxx_adult_base_logs.npz")
