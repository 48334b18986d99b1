"""CPU tests: the categorical dictionary rule and the Contains conditions of the model writer."""
import os

import numpy as np

import ydf_b200
from ydf_b200 import dataspec, model_io

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_toy_dictionary_kat():
    """dataset/data_spec_inference_test.cc:300-360 (toy.csv, min_vocab_frequency 2): Cat_1 = A,B,A,C ->
    {<OOD>: 2, A: 2}, most_frequent_value 1 (a tie with <OOD> goes to the first real item);
    Cat_2 = A,NA,B,NA -> {<OOD>: 2} only, most_frequent_value 0, two missing."""
    c1 = dataspec.infer_categorical_column("Cat_1", np.array(["A", "B", "A", "C"]), min_vocab_frequency=2)
    assert (c1.vocabulary, c1.counts, c1.na_bin, c1.num_bins) == (["<OOD>", "A"], [2, 2], 1, 2)
    c2 = dataspec.infer_categorical_column("Cat_2", np.array(["A", "", "B", ""]), min_vocab_frequency=2)
    assert (c2.vocabulary, c2.counts, c2.na_bin, c2.num_missing) == (["<OOD>"], [2], 0, 2)
    assert c1.encode(np.array(["A", "Z", "", None], dtype=object)).tolist() == [1, 0, 1, 1]


def test_equal_counts_are_ordered_by_key_descending():
    # std::greater<std::pair<int64, std::string>> (data_spec_inference.cc:346-347)
    c = dataspec.infer_categorical_column("c", np.array(["a", "b", "c", "b", "a", "c", "d"]), min_vocab_frequency=1)
    assert c.vocabulary == ["<OOD>", "c", "b", "a", "d"]


def test_pydf_front_end_dictionary_rule():
    """port/python/ydf/dataset/dataset.cc:402-455: equal counts by key ASCENDING, max_vocab_count -1 = no limit and
    0 = only <OOD>; :510-564 never sets most_frequent_value, so missing strings are encoded as <OOD>."""
    v = np.array(["a", "b", "c", "b", "a", "c", "d", ""])
    c = dataspec.infer_categorical_column("c", v, min_vocab_frequency=1, front_end=dataspec.FRONT_END_PYDF)
    assert (c.vocabulary, c.counts, c.na_bin, c.num_missing) == (["<OOD>", "a", "b", "c", "d"], [0, 2, 2, 2, 1], 0, 1)
    assert c.encode(np.array(["d", "", "zz"])).tolist() == [4, 0, 0]
    c = dataspec.infer_categorical_column("c", v, min_vocab_frequency=2, max_vocab_count=2, front_end=dataspec.FRONT_END_PYDF)
    assert (c.vocabulary, c.counts) == (["<OOD>", "a", "b"], [3, 2, 2])
    c = dataspec.infer_categorical_column("c", v, min_vocab_frequency=1, max_vocab_count=-1, front_end=dataspec.FRONT_END_PYDF)
    assert len(c.vocabulary) == 5
    c = dataspec.infer_categorical_column("c", v, min_vocab_frequency=1, max_vocab_count=0, front_end=dataspec.FRONT_END_PYDF)
    assert (c.vocabulary, c.counts) == (["<OOD>"], [7])
    # the dictionaries of a PYDF-trained reference model (golden adult_binary_class_gbdt_v2; native_country has five
    # pairs of equal counts) are checked in tests/reference_replay.py::replay


def test_adult_dictionaries_match_the_reference_model():
    """The dictionaries the reference inferred for its golden Adult model (fixture generator:
    tests/golden/make_adult_categorical_fixture.py)."""
    z = np.load(os.path.join(G, "adult_categorical.npz"))
    for c in ["workclass", "education", "marital_status", "occupation", "relationship", "race", "sex",
              "native_country"]:
        col = dataspec.infer_categorical_column(c, z[f"strings_{c}"][z[f"train_{c}"]])
        assert col.vocabulary == list(z[f"ref_vocab_{c}"]), c
        assert col.na_bin == int(z[f"ref_mfv_{c}"]), c


def test_contains_conditions_round_trip(tmp_path):
    """SetPositiveAttributeSetOfCategoricalContainsCondition (learner/decision_tree/utils.cc:31-63):
    bitmap when ceil(num_values / 8) <= 4 * |positive set|, else a sorted vector."""
    big = dataspec.CategoricalColumn("big", ["<OOD>"] + [f"v{i}" for i in range(199)], [0] * 200, 200, 1)
    small = dataspec.CategoricalColumn("small", ["<OOD>", "a", "b"], [0, 5, 5], 3, 1)
    spec = dataspec.DataSpec(columns=[big, small], label="y", task="REGRESSION", num_rows=10)

    def mask(cs):
        m = [0] * 8
        for c in cs:
            m[c >> 5] |= 1 << (c & 31)
        return tuple(m)

    t = np.zeros(5, dtype=ydf_b200.NODE_DTYPE)
    t[0] = (0, 0, 0, 1, 1, 2, 0.5, 0.0, 10, 4, (0, 0, 10), 1, 0, mask([3, 150]))        # vector: 25 B > 8 B
    t[1] = (-1, 0, 0, 2, -1, -1, 0, -0.1, 6, 0, (0, 0, 6), 0, 0, (0,) * 8)
    t[2] = (1, 0, 1, 2, 3, 4, 0.25, 0.0, 4, 2, (0, 0, 4), 1, 0, mask([1]))               # bitmap: 1 B <= 4 B
    t[3] = (-1, 0, 0, 3, -1, -1, 0, 0.2, 2, 0, (0, 0, 2), 0, 0, (0,) * 8)
    t[4] = (-1, 0, 0, 3, -1, -1, 0, 0.3, 2, 0, (0, 0, 2), 0, 0, (0,) * 8)
    m = ydf_b200.GradientBoostedTreesModel(spec, [t], 0.0, "SQUARED_ERROR")
    m.save(str(tmp_path / "m"))
    raw = [model_io.pb_decode(r) for r in model_io.read_blob_sequence(str(tmp_path / "m" / "nodes-00000-of-00001"))]

    def cond_kind(node):
        nc = model_io.pb_decode(model_io._one(node, 3))
        return [f for f, _, _ in model_io.pb_decode(model_io._one(nc, 3))]

    assert cond_kind(raw[0]) == [4] and cond_kind(raw[2]) == [5]   # contains_condition / contains_bitmap_condition
    back = model_io.read_ydf_model(str(tmp_path / "m"))
    assert back["nodes"][0]["positive_categories"] == [3, 150]
    assert back["nodes"][2]["positive_categories"] == [1] and back["nodes"][2]["na_value"] is True
    assert back["columns"][2]["vocabulary"] == {"<OOD>": 0, "a": 1, "b": 2}
    # numpy traversal of the model follows the masks
    bins = np.array([[3, 150, 7, 3], [0, 0, 1, 1]], np.uint8)
    np.testing.assert_allclose(m._raw(bins), np.array([0.2, 0.2, -0.1, 0.3], np.float32))


def test_lossless_buckets():
    """dataspec.infer_column_lossless: one bucket per distinct value, boundaries strictly above the lower neighbour even
    for adjacent floats, NaN = missing -> a bucket of its own at the mean (the exact splitter imputes the mean and may cut on
    either side of the missing rows), None beyond 255 distinct values."""
    a = np.float32(1.0)
    b = np.nextafter(a, np.float32(2.0))
    v = np.array([3.0, a, b, np.nan, 3.0, -2.5, 7.0], np.float32)
    c = dataspec.infer_column_lossless("x", v)
    # distinct values -2.5, a, b, 3, 7 and the mean 2.08.. of the six present rows as a sixth "value"
    assert c.num_bins == 6 and c.num_missing == 1 and len(c.boundaries) == 5
    assert c.boundaries[1] == b          # (a + b) / 2 rounds to a: the boundary moves up to b
    codes = c.encode(v)
    assert c.na_bin == 3 and codes.tolist() == [4, 1, 2, 3, 4, 0, 5]
    assert c.na_bin == int(np.searchsorted(c.boundaries, np.float32(c.mean), side="right"))
    # both cuts around the missing rows exist: "b | NA" (threshold 3) and "NA | 3.0" (threshold 4)
    assert (codes >= 3).tolist() == [True, False, False, True, True, False, True]
    assert (codes >= 4).tolist() == [True, False, False, False, True, False, True]
    # no missing values: no extra bucket
    assert dataspec.infer_column_lossless("x", v[~np.isnan(v)]).num_bins == 5
    # values outside the statistics sample still get their own bucket
    assert dataspec.infer_column_lossless("x", np.array([1, 2, 3, 4, 5], np.float32), max_rows=2).num_bins == 5
    assert dataspec.infer_column_lossless("x", np.arange(255, dtype=np.float32)).num_bins == 255
    assert dataspec.infer_column_lossless("x", np.arange(256, dtype=np.float32)) is None
