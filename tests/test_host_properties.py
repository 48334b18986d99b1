"""Property tests (hypothesis) of the host-side rules of the library: the binning rule against the numpy
restatement of GenDiscretizedBoundaries in test_capi_cpu.py, the encode rule, the categorical dictionary."""
import numpy as np
from hypothesis import given, settings, strategies as st

import ydf_b200
from ydf_b200 import dataspec
from tests.test_capi_cpu import _np_boundaries

finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, width=32)


@settings(max_examples=60, deadline=None)
@given(values=st.lists(st.one_of(finite, st.sampled_from([0.0, -0.0, 1.0, 2.5, float("nan")])), min_size=2, max_size=400),
       max_bins=st.integers(min_value=4, max_value=64), min_obs=st.integers(min_value=1, max_value=6),
       round_to=st.sampled_from([None, 0, 1]))
def test_boundaries_match_the_numpy_restatement(values, max_bins, min_obs, round_to):
    v = np.array(values, dtype=np.float32)
    if round_to is not None:
        v = np.round(v, round_to).astype(np.float32)       # heavy ties
    if np.isnan(v).all():
        return
    got, mean = ydf_b200.discretize_boundaries(v, max_bins, min_obs)
    want, wmean = _np_boundaries(v, max_bins, min_obs)
    np.testing.assert_array_equal(got, want)
    assert np.all(np.diff(got) > 0) or len(got) <= 1        # strictly increasing boundaries
    na_bin = int(np.searchsorted(got, np.float32(mean), side="right"))
    enc = ydf_b200.discretize_encode(v, got, na_bin)
    ref = np.searchsorted(got, v, side="right")
    ref[np.isnan(v)] = na_bin
    np.testing.assert_array_equal(enc, ref.astype(np.uint8))
    # monotone: a larger value never gets a smaller bin
    ok = ~np.isnan(v)
    order = np.argsort(v[ok], kind="stable")
    assert np.all(np.diff(enc[ok][order].astype(int)) >= 0)


@settings(max_examples=60, deadline=None)
@given(keys=st.lists(st.sampled_from(["a", "b", "c", "dd", "e", "", "zz", "<OOD>"]), min_size=1, max_size=200),
       min_freq=st.integers(min_value=1, max_value=6))
def test_dictionary_rule_properties(keys, min_freq):
    col = dataspec.infer_categorical_column("c", np.array(keys, dtype=object), min_vocab_frequency=min_freq)
    assert col.vocabulary[0] == "<OOD>" and len(set(col.vocabulary)) == len(col.vocabulary)
    # counts are non-increasing over the real items; equal counts are ordered by key, descending
    items = list(zip(col.counts[1:], col.vocabulary[1:]))
    assert items == sorted(items, key=lambda kv: (kv[0], kv[1].encode()), reverse=True)
    assert all(c >= min_freq for c in col.counts[1:])
    assert sum(col.counts) + col.num_missing == len(keys)
    enc = col.encode(np.array(keys, dtype=object))
    for k, e in zip(keys, enc):
        if k == "":
            assert e == col.na_bin
        elif k in col.vocabulary[1:]:
            assert col.vocabulary[e] == k
        else:
            assert e == 0


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(min_value=0, max_value=2**32 - 1), n=st.integers(min_value=1, max_value=3000),
       ratio=st.floats(min_value=0.0, max_value=1.0))
def test_split_mask_is_a_prefix_consistent_draw(seed, n, ratio):
    m = ydf_b200.validation_split_mask(seed, n, np.float32(ratio))
    # one mt19937 draw per row in row order: the mask of a shorter dataset is a prefix
    np.testing.assert_array_equal(ydf_b200.validation_split_mask(seed, max(1, n // 2), np.float32(ratio)), m[:max(1, n // 2)])
    if ratio == 0.0:
        assert m.all()


@settings(max_examples=60, deadline=None)
@given(keys=st.lists(st.sampled_from(["a", "b", "c", "dd", "e", "", "zz"]), min_size=1, max_size=200),
       min_freq=st.integers(min_value=1, max_value=6), max_vocab=st.sampled_from([-1, 0, 1, 3, 2000]))
def test_pydf_dictionary_rule_properties(keys, min_freq, max_vocab):
    """FRONT_END_PYDF (port/python/ydf/dataset/dataset.cc:402-455): counts descending, equal counts by key ASCENDING,
    max_vocab_count -1 = unlimited and 0 = <OOD> only, missing strings encoded as <OOD> (most_frequent_value stays 0)."""
    col = dataspec.infer_categorical_column("c", np.array(keys, dtype=object), min_vocab_frequency=min_freq,
                                            max_vocab_count=max_vocab, front_end=dataspec.FRONT_END_PYDF)
    assert col.vocabulary[0] == "<OOD>" and col.na_bin == 0
    items = list(zip(col.counts[1:], col.vocabulary[1:]))
    assert items == sorted(items, key=lambda kv: (-kv[0], kv[1].encode()))
    assert all(c >= min_freq for c in col.counts[1:])
    if max_vocab >= 0:
        assert len(items) <= max_vocab
    assert sum(col.counts) + col.num_missing == len(keys)
    enc = col.encode(np.array(keys, dtype=object))
    for k, e in zip(keys, enc):
        assert (col.vocabulary[e] == k) if k in col.vocabulary[1:] else (e == 0)


@settings(max_examples=60, deadline=None)
@given(values=st.lists(st.one_of(finite, st.sampled_from([0.0, 1.0, 2.5, float("nan")])), min_size=1, max_size=300),
       round_to=st.sampled_from([None, 0, 1]))
def test_lossless_buckets_properties(values, round_to):
    """dataspec.infer_column_lossless: distinct values <-> buckets one to one, order preserved, NaN -> bucket of the mean."""
    v = np.array(values, dtype=np.float32)
    if round_to is not None:
        v = np.round(v, round_to).astype(np.float32)
    present = v[~np.isnan(v)]
    col = dataspec.infer_column_lossless("x", v)
    distinct = np.unique(present)
    if len(distinct) == 0 or len(distinct) > 255:
        assert col is None
        return
    has_na = bool(np.isnan(v).any())
    # with missing values the mean is a value of its own (the exact splitter's imputation), unless it already is one
    values = np.unique(np.append(distinct, np.float32(col.mean))) if has_na else distinct
    assert col.num_bins == len(values) and np.all(np.diff(col.boundaries) > 0)
    enc = col.encode(v)
    ok = ~np.isnan(v)
    np.testing.assert_array_equal(enc[ok], np.searchsorted(values, v[ok]).astype(np.uint8))   # bucket = rank of the value
    assert np.all(enc[~ok] == col.na_bin)
    if has_na:
        assert values[col.na_bin] == np.float32(col.mean)   # the NA rows have their own bucket, between their neighbours
    assert col.na_bin == int(np.searchsorted(col.boundaries, np.float32(col.mean), side="right"))
