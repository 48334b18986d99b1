"""numpy restatement of the reference's binning rule for DISCRETIZED_NUMERICAL columns (TEST INFRASTRUCTURE).

Used where the product library must not be loaded: bench.py's `--impl reference` arm prepares its data with this
module, and tests/test_binning_kat.py holds it against the product's host rule (csrc/ygg_dataspec.cc) and the
reference's KATs, so both bench arms see byte-identical bins.

Follows, relative to /root/reference/yggdrasil_decision_forests/:
  GenDiscretizedBoundaries                 dataset/data_spec.cc:854-986
  AddBucket                                dataset/data_spec.cc:77-107
  FinalizeComputeSpecDiscretizedNumerical  dataset/data_spec_inference.cc:226-250 (special values 0 and the mean)
  NumericalToDiscretizedNumerical          dataset/data_spec.cc:1006-1018 (upper_bound)
"""
import numpy as np


def _add_bucket(v, bounds):
    """AddBucket (data_spec.cc:77-107): a one-value bin [v - ulp, v + ulp]."""
    v = np.float32(v)
    lo = np.nextafter(v, np.float32(v - np.float32(1)), dtype=np.float32)
    hi = np.nextafter(v, np.float32(v + np.float32(1)), dtype=np.float32)
    if not bounds:
        return [lo, hi]
    kept = [b for b in bounds if not (lo <= b <= hi)]
    if not kept:   # undefined in the reference (min_element of an empty vector); one special bucket stays
        return [lo, hi]
    mn, mx = min(kept), max(kept)
    if mn < hi:
        kept.append(lo)
    if mx > lo:
        kept.append(hi)
    return kept


def gen_discretized_boundaries(values, counts, maximum_num_bins, min_obs_in_bins, special_values=()):
    """values: sorted unique float32 candidates, counts: their multiplicities -> sorted float32 boundaries."""
    values = np.asarray(values, np.float32)
    counts = np.asarray(counts, np.int64)
    nc = len(values)
    special = [np.float32(s) for s in special_values]
    in_bounds = sum(1 for s in special if nc and values[0] < s < values[-1])
    reserved = int(maximum_num_bins) - len(special) - in_bounds
    unlimited = reserved < 0          # the reference's size_t arithmetic wraps (data_spec.cc:889-896)
    max_bins = max(1, reserved)
    max_boundaries = max_bins - 1
    bounds = []
    mids = ((values[:-1] + values[1:]) / np.float32(2)).astype(np.float32) if nc > 1 else np.zeros(0, np.float32)
    if not unlimited and nc > max_bins:
        total = int(counts.sum())
        max_bins = max(1, min(max_bins, total // int(min_obs_in_bins)))
        large = total // max_bins
        is_large = counts >= large
        remaining_bins = max_bins - int(is_large.sum())
        remaining = total - int(counts[is_large].sum())
        if remaining_bins < 1:
            remaining_bins = 1
        cur_large = remaining // remaining_bins
        cum = np.concatenate([[0], np.cumsum(counts)])            # cum[i] = rows before candidate i
        cum_small = np.concatenate([[0], np.cumsum(np.where(is_large, 0, counts))])
        small_total = int(cum_small[-1])
        large_idx = np.flatnonzero(is_large)
        i = 0          # next candidate to consume
        start = 0      # first candidate of the running bin
        made = 0
        last = nc - 1  # the loop visits candidates 0 .. nc-2
        while i < last:
            # the cut falls on the first candidate j >= i with one of (data_spec.cc:935-941):
            #   (a) is_large[j]   (b) running >= cur_large   (c) is_large[j+1] and running >= max(1, cur_large // 2)
            ja = int(large_idx[np.searchsorted(large_idx, i)]) if np.searchsorted(large_idx, i) < len(large_idx) else nc
            jb = int(np.searchsorted(cum, cum[start] + cur_large, side="left")) - 1   # cum[j+1] - cum[start] >= cur_large
            jb = max(jb, i)
            j = min(ja, jb)
            # (c): a large candidate right after j' <= j with enough running count
            k = np.searchsorted(large_idx, i + 1)
            if k < len(large_idx):
                jc = int(large_idx[k]) - 1
                if jc < j and jc >= i and cum[jc + 1] - cum[start] >= max(1, cur_large // 2):
                    j = jc
            if j >= last:
                break
            bounds.append(mids[j])
            made += 1
            if made >= max_boundaries:
                break
            if not is_large[j]:
                remaining_bins = max(1, remaining_bins - 1)
                cur_large = (small_total - int(cum_small[j + 1])) // remaining_bins
            i = j + 1
            start = j + 1
    else:
        running = 0
        for j in range(nc - 1):
            running += int(counts[j])
            if running >= min_obs_in_bins:
                bounds.append(mids[j])
                running = 0
    for s in special:
        bounds = _add_bucket(s, bounds)
    return np.sort(np.asarray(bounds, np.float32))


def discretize_boundaries(values, maximum_num_bins=255, min_obs_in_bins=3):
    """-> (boundaries float32, mean float64) of one column; NaN = missing."""
    v = np.asarray(values, np.float32)
    v = v[~np.isnan(v)]
    mean = float(np.sum(v, dtype=np.longdouble) / len(v)) if len(v) else 0.0
    u, c = np.unique(v, return_counts=True)
    return gen_discretized_boundaries(u, c, maximum_num_bins, min_obs_in_bins, (0.0, np.float32(mean))), mean


def discretize_encode(values, boundaries, na_bin):
    v = np.asarray(values, np.float32)
    out = np.searchsorted(np.asarray(boundaries, np.float32), v, side="right").astype(np.uint8)
    out[np.isnan(v)] = na_bin
    return out
