"""The reference's known-answer tests for GenDiscretizedBoundaries (dataset/data_spec_test.cc:567-684),
restated against the library's host rule (csrc/ygg_dataspec.cc) — the rule the GPU binning path
(csrc/ygg_binning.cu) is then held to bit for bit in tests/test_gpu_binning.py."""
import numpy as np

import ydf_b200

G = ydf_b200.gen_discretized_boundaries
I5 = (np.arange(1, 6), np.ones(5))
I100 = (np.arange(1, 101), np.ones(100))
f32 = np.float32


def _eq(got, want):
    np.testing.assert_array_equal(got, np.array(want, dtype=np.float32))


def test_gen_discretized_boundaries():                       # data_spec_test.cc:567-620
    _eq(G(*I5, 4, 1), [1.5, 2.5, 3.5])
    _eq(G(*I100, 4, 1), [25.5, 50.5, 75.5])
    _eq(G(*I5, 10, 1), [1.5, 2.5, 3.5, 4.5])
    _eq(G(*I100, 10, 1), [10.5, 20.5, 30.5, 40.5, 50.5, 60.5, 70.5, 80.5, 90.5])
    _eq(G(*I5, 1000, 3), [3.5])
    _eq(G(*I100, 1000, 15), [15.5, 30.5, 45.5, 60.5, 75.5, 90.5])


def test_corner_cases():                                     # data_spec_test.cc:622-646
    _eq(G([], [], 10, 1), [])
    _eq(G(*I5, 1, 1), [])
    # 1 - 1 special - 1 in bounds wraps in the reference's size_t arithmetic: every candidate is kept
    _eq(G(*I5, 1, 1, [2.0]), [1.5, np.nextafter(f32(2), f32(1)), np.nextafter(f32(2), f32(3)), 2.5, 3.5, 4.5])


def test_special_values():                                   # data_spec_test.cc:648-684
    _eq(G(*I5, 5, 1, [0.0]), [np.nextafter(f32(0), f32(1)), 1.5, 2.5, 3.5])
    _eq(G(*I5, 6, 1, [2.0]), [1.5, np.nextafter(f32(2), f32(1)), np.nextafter(f32(2), f32(3)), 2.5, 3.5])
    _eq(G(*I5, 6, 1, [2.5]), [1.5, np.nextafter(f32(2.5), f32(2)), np.nextafter(f32(2.5), f32(3)), 3.5])
    _eq(G(*I5, 5, 1, [5.0]), [1.5, 2.5, 3.5, np.nextafter(f32(5), f32(4))])


def test_column_entry_uses_the_same_rule():
    """ygg_discretize_boundaries = sort + unique + the rule above with the special values {0, mean}."""
    rng = np.random.default_rng(3)
    v = np.round(rng.normal(size=5000) * 20).astype(np.float32)
    got, mean = ydf_b200.discretize_boundaries(v, 64, 3)
    u, c = np.unique(v, return_counts=True)
    _eq(got, G(u, c, 64, 3, [0.0, np.float32(mean)]))


def test_boundaries_of_a_reference_made_dataspec():
    """The reference's golden model adult_binary_class_rf_discret_numerical was trained on adult_train.csv with
    DISCRETIZED_NUMERICAL columns; its dataspec stores the boundaries GenDiscretizedBoundaries produced (73 / 254 / 100 /
    65 / 82 of them) and the column means.  The host rule of this repo (csrc/ygg_dataspec.cc, which the GPU binning of
    csrc/ygg_binning.cu is bit-identical to: tests/test_gpu_binning.py) reproduces every boundary bit for bit."""
    import os
    import ydf_b200
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref = np.load(os.path.join(G, "ydf_adult_discretized_dataspec.npz"))
    data = np.load(os.path.join(G, "adult_numerical.npz"))
    assert int(ref["created_num_rows"]) == 22792
    sizes = {}
    for name in ("age", "fnlwgt", "capital_gain", "capital_loss", "hours_per_week"):
        col = ydf_b200.dataspec.infer_column(name, data[f"train_{name}"].astype(np.float32))
        want = ref[f"boundaries_{name}"]
        assert col.boundaries.dtype == np.float32 and np.array_equal(col.boundaries, want), name
        assert abs(col.mean - float(ref[f"mean_{name}"])) <= 1e-9 * abs(col.mean)
        sizes[name] = len(want)
    assert sizes == {"age": 73, "fnlwgt": 254, "capital_gain": 100, "capital_loss": 65, "hours_per_week": 82}


def test_numpy_restatement_matches_the_product_rule():
    """oracle/binning.py (what bench.py's reference arm bins its data with, so that it never loads the product
    library) gives the product's host rule bit for bit: both bench arms see identical bytes."""
    from oracle import binning as B
    rng = np.random.default_rng(0)
    cases = []
    for n in (10, 100, 5000, 100000):
        cases.append(rng.normal(size=n).astype(np.float32))
        cases.append(np.round(rng.normal(size=n) * 3).astype(np.float32))
        cases.append(np.where(rng.random(n) < 0.4, 0, rng.exponential(size=n)).astype(np.float32))
        cases.append(rng.integers(0, 300, size=n).astype(np.float32))
        x = rng.normal(size=n).astype(np.float32)
        x[rng.random(n) < 0.1] = np.nan
        cases.append(x)
        cases.append(np.concatenate([np.full(n // 2, 1.5, np.float32), rng.normal(size=n - n // 2).astype(np.float32),
                                     np.full(n // 3, -2.25, np.float32)]))
    for x in cases:
        for max_bins, min_obs in ((255, 3), (256, 3), (16, 1), (64, 10), (4, 3)):
            want, mean = ydf_b200.discretize_boundaries(x, max_bins, min_obs)
            got, mean2 = B.discretize_boundaries(x, max_bins, min_obs)
            np.testing.assert_array_equal(got, want)
            assert mean == mean2
            na_bin = int(np.searchsorted(want, np.float32(mean), side="right"))
            np.testing.assert_array_equal(B.discretize_encode(x[:2000], got, na_bin), ydf_b200.discretize_encode(x[:2000], want, na_bin))
    # the reference's KATs (dataset/data_spec_test.cc:567-684) through the restatement as well
    b = B.gen_discretized_boundaries(np.array([1, 2, 3, 4], np.float32), [10, 10, 10, 10], 255, 3, ())
    np.testing.assert_array_equal(b, np.array([1.5, 2.5, 3.5], np.float32))
