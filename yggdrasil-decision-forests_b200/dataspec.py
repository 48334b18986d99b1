"""Column semantics and binning of the engine's input (DISCRETIZED_NUMERICAL columns).

Mirrors what PYDF does when `discretize_numerical_columns=True`
(port/python/ydf/dataset/dataset.cc:192-316 -> dataset/data_spec.cc:854-1018): per column a sorted
boundary vector, `bin = upper_bound(boundaries, x)`, NA replaced by the bin of the column mean.
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

    def encode(self, values) -> np.ndarray:
        return _capi.discretize_encode(np.asarray(values, dtype=np.float32), self.boundaries, self.na_bin)


@dataclasses.dataclass
class DataSpec:
    columns: List[DiscretizedColumn]
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
