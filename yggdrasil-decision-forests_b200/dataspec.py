"""Column semantics and binning of the engine's input (DISCRETIZED_NUMERICAL and CATEGORICAL columns).

Mirrors what PYDF does when `discretize_numerical_columns=True`
(port/python/ydf/dataset/dataset.cc:192-316 -> dataset/data_spec.cc:854-1018): per column a sorted
boundary vector, `bin = upper_bound(boundaries, x)`, NA replaced by the bin of the column mean.
String columns become CATEGORICAL with the reference's dictionary rule
(dataset/data_spec_inference.cc:277-441): items sorted by (count, key) descending, items rarer than
min_vocab_frequency folded into index 0 (<OOD>), NA replaced by the column's most_frequent_value.  That field differs
by front end: the C++ inference (CSV / CLI path, data_spec_inference.cc:436-441) sets it to the most frequent item,
while PYDF's in-memory path builds the column spec itself (port/python/ydf/dataset/dataset.cc:510-564) and never sets
it, so a PYDF-trained model carries most_frequent_value = 0 and missing strings go to the <OOD> bucket (pinned on the
reference's golden model adult_binary_class_gbdt_v2, tests/test_oracle_kat.py).  `na_replacement` selects the rule.
"""
import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi


@dataclasses.dataclass
class DiscretizedColumn:
    name: str
    boundaries: np.ndarray  # float32, sorted
    mean: float
    num_bins: int
    na_bin: int
    num_missing: int = 0
    num_values: int = 0
    feature_type = _capi.FEATURE_DISCRETIZED_NUMERICAL

    def encode(self, values) -> np.ndarray:
        return _capi.discretize_encode(np.asarray(values, dtype=np.float32), self.boundaries, self.na_bin)


@dataclasses.dataclass
class CategoricalColumn:
    name: str
    vocabulary: List[str]     # index -> key; vocabulary[0] == "<OOD>"
    counts: List[int]
    num_bins: int             # number_of_unique_values
    na_bin: int               # most_frequent_value
    num_missing: int = 0
    num_values: int = 0
    feature_type = _capi.FEATURE_CATEGORICAL

    def encode(self, values) -> np.ndarray:
        index = {k: i for i, k in enumerate(self.vocabulary) if i > 0}
        keys, na = _categorical_keys(values)
        out = np.fromiter((index.get(k, 0) for k in keys), dtype=np.uint8, count=len(keys))
        out[na] = self.na_bin
        return out


def _categorical_keys(values):
    """-> (list of str keys, NA mask).  Missing: None, NaN, or the empty string (data_spec_inference.cc:716)."""
    v = np.asarray(values, dtype=object)
    keys, na = [], np.zeros(len(v), dtype=bool)
    for i, x in enumerate(v):
        if x is None or (isinstance(x, float) and x != x):
            na[i] = True
            keys.append("")
            continue
        k = x.decode() if isinstance(x, bytes) else str(x)
        if k == "":
            na[i] = True
        keys.append(k)
    return keys, na


NA_MOST_FREQUENT = "most_frequent"   # C++ dataspec inference (CSV / CLI front end)
NA_OUT_OF_DICTIONARY = "pydf"        # PYDF in-memory front end: most_frequent_value left at 0 = <OOD>


def infer_categorical_column(name: str, values, min_vocab_frequency: int = 5, max_vocab_count: int = 2000,
                             max_rows: Optional[int] = None, na_replacement: str = NA_MOST_FREQUENT) -> CategoricalColumn:
    if na_replacement not in (NA_MOST_FREQUENT, NA_OUT_OF_DICTIONARY):
        raise ValueError(f"na_replacement: {na_replacement!r}")
    keys, na = _categorical_keys(values if max_rows is None else values[:max_rows])
    raw: Dict[str, int] = {}
    for k, is_na in zip(keys, na):
        if not is_na:
            raw[k] = raw.get(k, 0) + 1
    ood = raw.pop("<OOD>", 0)
    # std::greater<std::pair<int64, std::string>>: count, then key (byte order), both descending
    items = sorted(((c, k.encode()) for k, c in raw.items()), reverse=True)
    while items and items[-1][0] < min_vocab_frequency:
        ood += items.pop()[0]
    if max_vocab_count > 0 and len(items) > max_vocab_count:
        ood += sum(c for c, _ in items[max_vocab_count:])
        items = items[:max_vocab_count]
    vocabulary = ["<OOD>"] + [k.decode() for _, k in items]
    counts = [ood] + [c for c, _ in items]
    if len(vocabulary) > 256:
        raise NotImplementedError(
            f"column {name!r}: {len(vocabulary)} categories do not fit the engine's uint8 bins "
            "(raise min_vocab_frequency or lower max_vocab_count)")
    # the first non-OOD item with the highest count, unless <OOD> is strictly more frequent
    most_frequent = 1 if (len(counts) > 1 and counts[1] >= counts[0]) else 0
    if na_replacement == NA_OUT_OF_DICTIONARY:
        most_frequent = 0
    return CategoricalColumn(name=name, vocabulary=vocabulary, counts=counts, num_bins=len(vocabulary),
                             na_bin=most_frequent, num_missing=int(na.sum()), num_values=len(keys))


@dataclasses.dataclass
class DataSpec:
    columns: List  # DiscretizedColumn | CategoricalColumn
    label: str
    task: str
    label_classes: Optional[List] = None   # classification: class 1, class 2 (index 0 is OOD)
    label_mean: float = 0.0
    label_sd: float = 0.0
    label_min: float = 0.0
    label_max: float = 0.0
    num_rows: int = 0

    @property
    def feature_names(self) -> List[str]:
        return [c.name for c in self.columns]


def as_columns(ds) -> Dict[str, np.ndarray]:
    """Accepts a dict of arrays or a pandas DataFrame (the two in-memory inputs PYDF takes)."""
    if isinstance(ds, dict):
        return {k: np.asarray(v) for k, v in ds.items()}
    try:
        import pandas as pd
        if isinstance(ds, pd.DataFrame):
            return {c: ds[c].to_numpy() for c in ds.columns}
    except ImportError:
        pass
    raise TypeError("dataset must be a dict of numpy arrays or a pandas DataFrame")


def infer_column(name: str, values, maximum_num_bins: int = 255, min_obs_in_bins: int = 3,
                 max_rows: Optional[int] = None) -> DiscretizedColumn:
    v = np.asarray(values, dtype=np.float32)
    sample = v if (max_rows is None or len(v) <= max_rows) else v[:max_rows]
    boundaries, mean = _capi.discretize_boundaries(sample, maximum_num_bins, min_obs_in_bins)
    if len(boundaries) + 1 > 256:
        raise ValueError(f"column {name!r}: {len(boundaries) + 1} bins do not fit the engine's uint8 bins")
    # NumericalToDiscretizedNumerical(mean): training.cc:917-922
    na_bin = int(np.searchsorted(boundaries, np.float32(mean), side="right"))
    return DiscretizedColumn(name=name, boundaries=boundaries, mean=float(mean),
                             num_bins=len(boundaries) + 1, na_bin=na_bin,
                             num_missing=int(np.isnan(v).sum()), num_values=len(v))


def encode_features(cols: Dict[str, np.ndarray], columns: Sequence[DiscretizedColumn]) -> np.ndarray:
    n = len(next(iter(cols.values())))
    out = np.empty((len(columns), n), dtype=np.uint8)
    for i, c in enumerate(columns):
        out[i] = c.encode(cols[c.name])
    return out
