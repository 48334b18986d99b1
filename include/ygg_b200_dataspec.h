/*
 * ygg_b200_dataspec.h — host-side binning helpers of libygg_b200.so (no GPU needed).
 * They implement the rule that produces the engine's input contract (uint8 bins); the reference
 * does this inside dataspec inference and VerticalDataset population.
 */
#ifndef YGG_B200_DATASPEC_H_
#define YGG_B200_DATASPEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GenDiscretizedBoundaries (dataset/data_spec.cc:854-986) with the special values {0, mean}
 * (FinalizeComputeSpecDiscretizedNumerical, dataset/data_spec_inference.cc:226-250).
 * values: n floats, NaN = missing.  Writes the sorted boundaries (at most maximum_num_bins - 1)
 * and the column mean of the non-missing values.  Returns 0 or YGG_ERR_INVALID_ARGUMENT. */
int ygg_discretize_boundaries(const float* values, int64_t n, int32_t maximum_num_bins,
                              int32_t min_obs_in_bins, float* out_boundaries, int32_t capacity,
                              int32_t* out_num_boundaries, double* out_mean);

/* NumericalToDiscretizedNumerical (dataset/data_spec.cc:1006-1018): bin = upper_bound(boundaries, x);
 * missing values are folded into na_bin (see ygg_dataset_create). */
int ygg_discretize_encode(const float* values, int64_t n, const float* boundaries,
                          int32_t num_boundaries, int32_t na_bin, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* YGG_B200_DATASPEC_H_ */
