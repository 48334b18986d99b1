"""Builds tests/golden/ydf_run_{adult,iris,abalone}_v2.npz: complete training runs of the reference, as stored in its
golden models test_data/model/{adult_binary_class,iris_multi_class,abalone_regression}_gbdt_v2.  Each is
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


def build(model, out, csv_name=None):
    m = model_io.read_ydf_model(f"{R}/model/{model}")
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
        assert nd["n_cond"] == nd["n"]
        if "positive_categories" in nd:
            assert max(nd["positive_categories"]) < 64  # Adult's native_country: 41 values
            mask[i] = sum(1 << c for c in nd["positive_categories"])
        else:
            threshold[i] = nd["higher_threshold"]
    logs = m["training_logs"]
    extra = {f"vocabulary_{c['name']}": np.array(sorted(c["vocabulary"], key=c["vocabulary"].get))
             for c in cols if c["type"] == 4}
    if csv_name:
        rows = list(csv.DictReader(open(f"{R}/dataset/{csv_name}")))
        for c in cols:
            v = [r[c["name"]] for r in rows]
            extra[f"data_{c['name']}"] = np.array(v) if c["type"] == 4 else np.array([float(x) for x in v], np.float32)
    path = os.path.join(HERE, out)
    np.savez_compressed(
        path, loss=m["loss"], task=m["task"], label_col_idx=m["label_col_idx"], num_trees_per_iter=m["num_trees_per_iter"],
        tree_first=np.array(tree_first, np.int32), initial_predictions=np.array(m["initial_predictions"], np.float32),
        validation_loss=np.float32(m["validation_loss"]),
        log_num_trees=np.array([e["number_of_trees"] for e in logs], np.int32),
        log_training_loss=np.array([e["training_loss"] for e in logs], np.float32),
        log_training_secondary=np.array([e["training_secondary"] for e in logs], np.float32),
        log_validation_loss=np.array([e["validation_loss"] for e in logs], np.float32),
        log_validation_secondary=np.array([e["validation_secondary"] for e in logs], np.float32),
        column_names=np.array([c["name"] for c in cols]), column_types=np.array([c["type"] for c in cols], np.int32),
        most_frequent_value=np.array([c.get("most_frequent_value", -1) for c in cols], np.int32),
        feature=feature, threshold=threshold, positive_mask=mask, n=np.array([nd["n"] for nd in nodes], np.int64),
        n_pos=n_pos, split_score=score, na_value=na_value, value=np.array([nd["top_value"] for nd in nodes], np.float32),
        **extra)
    print(path, os.path.getsize(path), "bytes;", m["num_trees"], "trees,", K, "nodes,", len(logs), "log entries")


build("adult_binary_class_gbdt_v2", "ydf_run_adult_v2.npz")
build("iris_multi_class_gbdt_v2", "ydf_run_iris_v2.npz", "iris.csv")
build("abalone_regression_gbdt_v2", "ydf_run_abalone_v2.npz", "abalone.csv")
