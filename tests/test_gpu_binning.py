"""On-GPU binning (csrc/ygg_binning.cu) against the host rule (csrc/ygg_dataspec.cc,
pinned by the reference's KATs in tests/test_binning_kat.py): boundaries, mean, NA bin and every encoded
bin must be identical, on distributions that exercise each branch of the rule."""
import numpy as np
import pytest

import ydf_b200
from oracle import oracle as O
from tests.util import compare_trees

pytestmark = pytest.mark.gpu


def _columns(n, rng):
    x = rng.normal(size=n).astype(np.float32)
    cols = {
        "normal": x,
        "with_nan": np.where(rng.random(n) < 0.07, np.nan, x).astype(np.float32),
        "integers": rng.integers(-40, 300, size=n).astype(np.float32),             # heavy ties, > max_bins values
        "few_values": rng.integers(0, 12, size=n).astype(np.float32),               # direct branch
        "skewed": np.concatenate([np.zeros(n // 3), rng.exponential(size=n // 3),
                                  np.full(n - 2 * (n // 3), 7.5)]).astype(np.float32),  # large candidates
        "zipf": np.minimum(rng.zipf(1.3, size=n), 5000).astype(np.float32),         # many large + long tail
        "constant": np.full(n, 3.25, np.float32),
        "neg_zero": np.where(rng.random(n) < 0.5, -0.0, 0.0).astype(np.float32) + (rng.random(n) < 0.1) * x,
        "all_nan_but_few": np.where(np.arange(n) < 5, x, np.nan).astype(np.float32),
        "huge_range": (x * 1e30).astype(np.float32),
    }
    return cols


@pytest.mark.parametrize("n", [1, 37, 8192, 100003])
@pytest.mark.parametrize("max_bins,min_obs", [(255, 3), (256, 3), (16, 3), (4, 1), (64, 50)])
def test_matches_host_rule(n, max_bins, min_obs):
    rng = np.random.default_rng(1000 + n)
    cols = _columns(n, rng)
    b = ydf_b200.DatasetBuilder(n, len(cols))
    want = {}
    for f, (name, v) in enumerate(cols.items()):
        bounds, mean, na, miss = b.add_numerical(f, v, max_bins, min_obs)
        wb, wmean = ydf_b200.discretize_boundaries(v, max_bins, min_obs)
        np.testing.assert_array_equal(bounds, wb, err_msg=name)
        assert mean == pytest.approx(wmean, rel=1e-14, abs=1e-300), name
        assert np.float32(mean) == np.float32(wmean), name
        wna = int(np.searchsorted(wb, np.float32(wmean), side="right"))
        assert na == wna and miss == int(np.isnan(v).sum()), name
        want[f] = ydf_b200.discretize_encode(v, wb, wna)
    ds = b.finish()
    for f, name in enumerate(cols):
        np.testing.assert_array_equal(ds.get_bins(f), want[f], err_msg=name)


def test_async_ring_matches_synchronous_calls():
    """More columns than lanes in flight; collected out of order."""
    rng = np.random.default_rng(77)
    n, k = 50000, 8
    cols = [np.where(rng.random(n) < 0.02, np.nan, rng.normal(size=n) * (j + 1)).astype(np.float32) for j in range(k)]
    b = ydf_b200.DatasetBuilder(n, k)
    for j in range(k):
        b.add_numerical_async(j, cols[j], 128, 3)
    got = {j: b.get_numerical(j) for j in (5, 0, 7, 1, 2, 3, 4, 6)}
    ds = b.finish()
    for j in range(k):
        wb, wmean = ydf_b200.discretize_boundaries(cols[j], 128, 3)
        np.testing.assert_array_equal(got[j][0], wb)
        np.testing.assert_array_equal(ds.get_bins(j), ydf_b200.discretize_encode(cols[j], wb, got[j][2]))


def test_statistics_on_a_row_prefix():
    """max_num_scanned_rows_to_compute_statistics: boundaries from the first rows, every row encoded."""
    rng = np.random.default_rng(5)
    v = rng.normal(size=50000).astype(np.float32)
    v[40000:] += 3
    b = ydf_b200.DatasetBuilder(len(v), 1)
    bounds, mean, na, miss = b.add_numerical(0, v, 255, 3, n_stats_rows=10000)
    wb, wmean = ydf_b200.discretize_boundaries(v[:10000], 255, 3)
    np.testing.assert_array_equal(bounds, wb)
    np.testing.assert_array_equal(b.finish().get_bins(0), ydf_b200.discretize_encode(v, wb, na))


def test_mixed_builder_trains_like_the_host_binned_dataset():
    """Builder (GPU-binned numerical + host-encoded categorical) -> same trees as Dataset(host bins), and
    both equal to the oracle."""
    rng = np.random.default_rng(9)
    n = 30000
    x = [rng.normal(size=n).astype(np.float32) for _ in range(4)]
    x[1][rng.random(n) < 0.05] = np.nan
    cat = rng.integers(0, 9, size=n).astype(np.uint8)
    y = ((x[0] + 0.5 * np.nan_to_num(x[1]) + (cat % 3 == 0) + 0.3 * rng.normal(size=n)) > 0.3).astype(np.int32) + 1
    b = ydf_b200.DatasetBuilder(n, 5)
    host_bins, nb, na = [], [], []
    for f in range(4):
        bounds, mean, nab, _ = b.add_numerical(f, x[f], 64, 3)
        host_bins.append(ydf_b200.discretize_encode(x[f], bounds, nab))
        nb.append(len(bounds) + 1)
        na.append(nab)
    b.add_bins(4, cat, 9, 1, feature_type=1)
    host_bins.append(cat); nb.append(9); na.append(1)
    ds_gpu = b.finish()
    ft = [0, 0, 0, 0, 1]
    ds_host = ydf_b200.Dataset(np.stack(host_bins), nb, na, feature_types=ft)
    cfg = ydf_b200.default_config(num_trees=5, max_depth=5)
    trees = []
    for ds in (ds_gpu, ds_host):
        g = ydf_b200.Gbt(ds, cfg)
        g.set_labels(y)
        g.train(5)
        trees.append([g.get_tree(i) for i in range(5)])
    assert [t.tobytes() for t in trees[0]] == [t.tobytes() for t in trees[1]]
    o = O.default_config()
    for k, _ in cfg._fields_:
        if k != "reserved":
            setattr(o, k, getattr(cfg, k))
    O.set_stable_category_sort(True)
    try:
        ref = O.gbt_train(np.stack(host_bins), nb, na, y, o, 5, num_threads=4, feature_type=ft)
    finally:
        O.set_stable_category_sort(False)
    for a, r in zip(trees[0], ref["trees"]):
        assert not compare_trees(a, r)


def test_builder_errors():
    b = ydf_b200.DatasetBuilder(100, 2)
    b.add_numerical(0, np.zeros(100, np.float32))
    with pytest.raises(ydf_b200.YggError):
        b.finish()                                            # feature 1 never added
    with pytest.raises(ydf_b200.YggError):
        b.add_numerical(1, np.zeros(100, np.float32), maximum_num_bins=300)
    with pytest.raises(ydf_b200.YggError):
        b.add_bins(1, np.zeros(100, np.uint8), 4, 7)          # na_bin outside [0, num_bins)
    b.close()
