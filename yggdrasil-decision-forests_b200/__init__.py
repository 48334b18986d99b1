"""ygg_b200 — B200-native GBT histogram split-finding engine behind YDF's learner surface.

Import as `import ydf_b200` (repo-root shim) — the directory name carries a hyphen.
"""
from ._capi import (Comm, Dataset, DatasetBuilder, gen_discretized_boundaries, validation_split_mask, Gbt, YggError, NODE_DTYPE, default_config, device_count,
                    discretize_boundaries, discretize_encode, lib, feature_shard,
                    merge_shard_best, SHARD_BEST_DTYPE)
from .learner import GradientBoostedTreesLearner, Task
from .model import GradientBoostedTreesModel
from . import dataspec
from . import model_io

__all__ = ["Comm", "Dataset", "DatasetBuilder", "gen_discretized_boundaries", "validation_split_mask", "Gbt", "YggError", "NODE_DTYPE", "default_config", "device_count",
           "discretize_boundaries", "discretize_encode", "lib", "feature_shard", "merge_shard_best",
           "SHARD_BEST_DTYPE", "GradientBoostedTreesLearner",
           "Task", "GradientBoostedTreesModel", "dataspec"]
