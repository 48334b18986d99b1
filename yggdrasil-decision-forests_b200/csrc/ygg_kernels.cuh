// ygg_kernels.cuh — hand-written sm_100a kernels of the GBT histogram split finder.
//
// Kernel inventory, one boosting iteration = pred/grad -> quantise -> per level
// {hist, scan, select, partition, node_stats}:
//   k_pred_grad   UpdatePredictions + Loss + UpdateGradients, fused (gradient_boosted_trees.cc:1445,
//                 :1544, :1575; loss_imp_binomial.cc:124-144, loss_imp_mean_square_error.cc:96-120)
//   k_quantize    fixed-point encoding of g/h, root statistics (loss_utils.cc:49-132 for the root)
//   k_hist        FillExampleBucketSet for all open nodes x features of one level
//                 (splitter_scanner.h:859-909) — the HBM-bound hot kernel
//   k_scan        ScanSplits<bucket_interpolation=true> (splitter_scanner.h:931-1101)
//   k_select      FindBestConditionConcurrentManager's ordered arg-max (training.cc:1728-1746) + NodeTrain
//                 bookkeeping (training.cc:4865-5049)
//   k_partition   SplitExamplesInPlace (training.cc:5243-5305) as a node-id relabel + child statistics
//   k_node_stats  SetLeafValueWithNewtonRaphsonStep (loss_utils.cc:49-132)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ygg_device.cuh"
#include "ygg_hist.cuh"

namespace ygg {

// ---------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Power of two P with |g| <= P for all rows, from the float bits of max|g|.
__device__ __forceinline__ float pow2_cover(unsigned int max_bits) {
  if (max_bits == 0u) return 1.f;
  const int e = static_cast<int>(max_bits >> 23) - 127;
  const bool exact = (max_bits & 0x7FFFFFu) == 0u;
  return exact ? exp2f(static_cast<float>(e)) : exp2f(static_cast<float>(e + 1));
}

struct GradParams {
  int64_t n;
  float* pred;
  const uint8_t* label_u8;   // binomial: 1 if the row's class is the positive one ("2")
  const float* label_f32;    // regression target
  const uint16_t* node_of_row;
  const NodeRec* pending_tree;  // tree whose leaves are still to be added to pred (or null)
  float* g;
  float* h;
  DeviceState* st;
  int compute_grad;
};

// expf / logf evaluated in double and rounded once: within the reference's glibc (<1 ulp,
// correctly rounded for all but ~1e-3 of inputs) far more often than the 2-ulp device expf.
__device__ __forceinline__ float exp_rn(float x) { return static_cast<float>(exp(static_cast<double>(x))); }
__device__ __forceinline__ float log_rn(float x) { return static_cast<float>(log(static_cast<double>(x))); }

template <int LOSS>
__global__ void __launch_bounds__(256) k_pred_grad(GradParams p) {
  double loss = 0;
  unsigned long long correct = 0;
  float gmax = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.n; r += stride) {
    float pred = p.pred[r];
    if (p.pending_tree != nullptr) {
      // UpdatePredictionWithSingleUnivariateTree (loss_utils.cc:214-229): the leaf of a row is its
      // final node id, no traversal needed.
      pred += p.pending_tree[p.node_of_row[r]].leaf_value;
      p.pred[r] = pred;
      if (LOSS == 0) {
        // loss_imp_binomial.cc:204-234, float arithmetic as in the reference.
        const float label = p.label_u8[r] ? 1.f : 0.f;
        const float term = 2 * (label * pred - log_rn(1.f + exp_rn(pred)));
        loss -= term;
        const bool predicted_pos = pred > 0.f;
        correct += (predicted_pos == (p.label_u8[r] != 0)) ? 1ull : 0ull;
      } else {
        const float d = p.label_f32[r] - pred;  // metric/metric.cc:2173-2199
        loss += d * d;
      }
    }
    if (p.compute_grad) {
      float g, h;
      if (LOSS == 0) {
        const float label = p.label_u8[r] ? 1.f : 0.f;
        const float proba = 1.f / (1.f + exp_rn(-pred));
        g = label - proba;
        h = proba * (1 - proba);
      } else {
        g = p.label_f32[r] - pred;
        h = 1.f;
      }
      p.g[r] = g;
      if (LOSS == 0) p.h[r] = h;
      gmax = fmaxf(gmax, fabsf(g));
    }
  }
  // block reduction
  loss = warp_sum_f64(loss);
  correct = warp_sum_u64(correct);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  __shared__ double s_loss[8];
  __shared__ unsigned long long s_cor[8];
  __shared__ float s_gmax[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_loss[w] = loss; s_cor[w] = correct; s_gmax[w] = gmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) { loss += s_loss[i]; correct += s_cor[i]; gmax = fmaxf(gmax, s_gmax[i]); }
    if (p.pending_tree != nullptr) {
      atomicAdd(&p.st->loss_sum, loss);
      atomicAdd(&p.st->correct, correct);
    }
    if (p.compute_grad) atomicMax(&p.st->gmax_bits, __float_as_uint(gmax));
  }
}

// ---------------------------------------------------------------------------------------------
struct QuantParams {
  int64_t n;
  int64_t n_pad;
  const float* g;
  const float* h;          // null for squared error (h == 1)
  uint32_t* q24;           // [n_pad]  biased 24-bit quantised gradient of every row
  uint32_t* hq24;          // [n_pad]  24-bit quantised hessian (hessian histogram only, else null)
  uint2* act;              // root level active lists (dense): (q24 | slot 0, row offset)
  uint32_t* act_h;
  int32_t* act_count;
  uint16_t* node_of_row;
  DeviceState* st;
  int root_candidate;
  float h_pow2;
};

__device__ __forceinline__ uint32_t quant_biased(float v, float scale, uint32_t bias, uint32_t vmax) {
  // rint(v * scale) + bias, clamped to [0, vmax]; scale is a power of two so v*scale is exact.
  const float t = rintf(v * scale) + static_cast<float>(bias);
  return static_cast<uint32_t>(fminf(fmaxf(t, 0.f), static_cast<float>(vmax)));
}
__device__ __forceinline__ uint32_t quant_biased_d(float v, double scale, uint32_t bias, uint32_t vmax) {
  const double t = rint(static_cast<double>(v) * scale) + static_cast<double>(bias);
  return static_cast<uint32_t>(fmin(fmax(t, 0.0), static_cast<double>(vmax)));
}

__global__ void __launch_bounds__(256) k_quantize(QuantParams p) {
  const float P = pow2_cover(p.st->gmax_bits);
  const float qscale = static_cast<float>(1u << (kQBits - 1)) / P;     // 2^23 / P
  const double sscale = static_cast<double>(1u << (kSBits - 1)) / P;   // 2^30 / P
  const double s2scale = static_cast<double>(1u << kSBits) / (static_cast<double>(P) * P);  // g^2 in [0, P^2]
  const double hscale = static_cast<double>(1u << kSBits) / p.h_pow2;  // h in [0, h_pow2]
  const float hqscale = static_cast<float>(1u << kQBits) / p.h_pow2;
  unsigned long long sg = 0, sh = 0, sg2 = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.n_pad; r += stride) {
    if ((r & (kBlockRows - 1)) == 0) {
      // one thread per block publishes the block's active-row count: every real row at the root
      const int64_t left = p.n - r;
      p.act_count[r / kBlockRows] = p.root_candidate ? static_cast<int32_t>(left < 0 ? 0 : (left > kBlockRows ? kBlockRows : left)) : 0;
    }
    if (r < p.n) {
      const float g = p.g[r];
      const uint32_t q = quant_biased(g, qscale, kQBias, kQMax);
      p.q24[r] = q;
      p.act[r] = make_uint2(q, static_cast<uint32_t>(r & (kBlockRows - 1)));  // slot 0
      p.node_of_row[r] = 0;
      sg += quant_biased_d(g, sscale, kSBias, 0x7FFFFFFFu);
      sg2 += quant_biased_d(g * g, s2scale, 0u, 0x7FFFFFFFu);  // float product, as loss_utils.cc:94
      if (p.h != nullptr) {
        const float h = p.h[r];
        sh += quant_biased_d(h, hscale, 0u, 0x7FFFFFFFu);
        if (p.hq24 != nullptr) {
          const uint32_t hq = static_cast<uint32_t>(fminf(rintf(h * hqscale), static_cast<float>(kQMax)));
          p.hq24[r] = hq;
          p.act_h[r] = hq;
        }
      }
    }
  }
  sg = warp_sum_u64(sg); sh = warp_sum_u64(sh); sg2 = warp_sum_u64(sg2);
  __shared__ unsigned long long s[3][8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s[0][w] = sg; s[1][w] = sh; s[2][w] = sg2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) { sg += s[0][i]; sh += s[1][i]; sg2 += s[2][i]; }
    atomicAdd(&p.st->root_sg, sg);
    atomicAdd(&p.st->root_sh, sh);
    atomicAdd(&p.st->root_sg2, sg2);
    if (blockIdx.x == 0) { p.st->g_pow2 = P; p.st->h_pow2 = p.h_pow2; }
  }
}

// ---------------------------------------------------------------------------------------------
// k_scan: one CTA of 256 threads (thread b = bin b) per (family, feature).
struct ScanParams {
  int level;
  const LevelDesc* levels;
  const Family* families;     // families of this level
  NodeRec* nodes;
  int f_begin, f_count;
  const int32_t* num_bins;    // per dataset feature
  const int32_t* na_bin;
  unsigned long long* hist_sum;        // this level  [nodes][f_count][256]
  uint32_t* hist_cnt;
  unsigned long long* hist_hsum;
  const unsigned long long* phist_sum; // parent level
  const uint32_t* phist_cnt;
  const unsigned long long* phist_hsum;
  Candidate* cand;            // [level nodes][f_count]
  const DeviceState* st;
  int min_num_obs;
  int use_hessian;
  int has_h;                  // 0: h == 1 for every row, the hessian sum of a bin is its count
  int subtract_parent;
  double l1, l2;
  int write_derived;          // 0 on the last level (the derived histogram is never a parent)
};

__device__ __forceinline__ double l1_threshold_d(double v, double l1) {
  if (l1 == 0.0) return v;
  const double len = fmax(0.0, fabs(v) - l1);
  return v > 0 ? len : -len;
}

// Block-wide inclusive scan of (count, sum, hsum) over 256 threads.
struct Scan3 { long long c; long long s; long long h; };
__device__ __forceinline__ Scan3 block_inclusive_scan(Scan3 v, Scan3* s_warp /*[8]*/, Scan3* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long c = __shfl_up_sync(0xffffffffu, v.c, o);
    const long long s = __shfl_up_sync(0xffffffffu, v.s, o);
    const long long h = __shfl_up_sync(0xffffffffu, v.h, o);
    if (lane >= o) { v.c += c; v.s += s; v.h += h; }
  }
  if (lane == 31) s_warp[w] = v;
  __syncthreads();
  Scan3 off{0, 0, 0}, tot{0, 0, 0};
  for (int i = 0; i < 8; i++) {
    if (i < w) { off.c += s_warp[i].c; off.s += s_warp[i].s; off.h += s_warp[i].h; }
    tot.c += s_warp[i].c; tot.s += s_warp[i].s; tot.h += s_warp[i].h;
  }
  __syncthreads();
  v.c += off.c; v.s += off.s; v.h += off.h;
  *total = tot;
  return v;
}

// Scans one node's 256-bin histogram held one bin per thread; thread 0 writes the Candidate.
__device__ void scan_node(const ScanParams& p, const NodeRec& node, int f_global, long long cnt,
                          long long sq /*unbiased quantised sum*/, long long hq, Candidate* out) {
  __shared__ Scan3 s_warp[8];
  __shared__ double s_best_score[8];
  __shared__ int s_best_b[8];
  __shared__ int s_interp[8];
  const int b = threadIdx.x;
  const int B = p.num_bins[f_global];
  Scan3 tot;
  const Scan3 inc = block_inclusive_scan(Scan3{cnt, sq, hq}, s_warp, &tot);
  const double ginv = static_cast<double>(p.st->g_pow2) / static_cast<double>(1u << (kQBits - 1));
  const double hinv = static_cast<double>(p.st->h_pow2) / static_cast<double>(1u << kQBits);
  const long long n_neg = inc.c, n_pos = tot.c - inc.c;
  bool valid = (b <= B - 2) && (n_pos >= p.min_num_obs) && (n_neg >= p.min_num_obs);
  double score = 0.0;
  double min_score = 0.0;
  if (!p.use_hessian) {
    // Variance reduction (V0 - V_pos - V_neg) / c0 (splitter_scanner.h:911-926,
    // splitter_accumulator.h:1517-1519).  With consistent sums the sum-of-squares terms cancel and
    // the numerator is the between-group sum of squares n_pos*n_neg/c0 * (mean_pos - mean_neg)^2,
    // evaluated from the integer sums as d^2 / (n_pos*n_neg*c0) with d = S_pos*n_neg - S_neg*n_pos:
    // non-negative by construction and exactly 0 for equal means.
    const double c0 = static_cast<double>(tot.c);
    if (valid) {
      const double np_ = static_cast<double>(n_pos), nn_ = static_cast<double>(n_neg);
      const double d = (static_cast<double>(tot.s - inc.s) * nn_ - static_cast<double>(inc.s) * np_) * ginv;
      score = (d / np_) * (d / nn_) / (c0 * c0);
    }
  } else {
    // splitter_accumulator.h:755-773 (Score), :1706-1727 (parent / minimum score).
    const double g0 = l1_threshold_d(node.stat[0], p.l1);
    const double parent_full = g0 * g0 / (node.stat[1] + p.l2);
    const double parent_score = p.subtract_parent ? parent_full : 0.0;
    min_score = p.subtract_parent ? 0.0 : parent_full;
    if (valid) {
      const double gn = l1_threshold_d(static_cast<double>(inc.s) * ginv, p.l1);
      const double gp = l1_threshold_d(static_cast<double>(tot.s - inc.s) * ginv, p.l1);
      const double hn = fmax(static_cast<double>(inc.h) * hinv, kMinHessianForNewtonStep) + p.l2;
      const double hp = fmax(static_cast<double>(tot.h - inc.h) * hinv, kMinHessianForNewtonStep) + p.l2;
      score = gp * gp / hp + gn * gn / hn - parent_score;
    }
  }
  // best_score starts at max(condition.split_score (0), MinimumScore()) and needs strict '>'.
  valid = valid && (score > min_score) && (score > 0.0 || p.use_hessian);
  // arg-max with the lowest bin on ties (sequential strict '>' keeps the first maximum).
  double bs = valid ? score : -1.0;
  int bb = valid ? b : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int ob = __shfl_xor_sync(0xffffffffu, bb, o);
    if (os > bs || (os == bs && ob < bb)) { bs = os; bb = ob; }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_best_score[w] = bs; s_best_b[w] = bb; }
  __syncthreads();
  bs = s_best_score[0]; bb = s_best_b[0];
  for (int i = 1; i < 8; i++) {
    if (s_best_score[i] > bs || (s_best_score[i] == bs && s_best_b[i] < bb)) { bs = s_best_score[i]; bb = s_best_b[i]; }
  }
  const bool found = bb != 0x7fffffff;
  // Bucket interpolation (splitter_scanner.h:993-1000, :1076-1086): first non-empty bucket after the
  // best one that the sequential scan visits (indices <= B-2).
  int cand_i = (found && b > bb && b <= B - 2 && cnt > 0) ? b : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand_i = min(cand_i, __shfl_xor_sync(0xffffffffu, cand_i, o));
  if (lane == 0) s_interp[w] = cand_i;
  __syncthreads();
  int interp = s_interp[0];
  for (int i = 1; i < 8; i++) interp = min(interp, s_interp[i]);
  // n_pos at the best boundary lives in thread bb.
  __shared__ long long s_npos;
  if (found && b == bb) s_npos = n_pos;
  __syncthreads();
  if (threadIdx.x == 0) {
    Candidate c;
    c.found = found ? 1 : 0;
    c.score = found ? static_cast<float>(bs) : 0.f;
    int idx = bb;
    if (found && interp != 0x7fffffff && interp != bb + 1) idx = (bb + interp) / 2;
    c.thr = found ? idx + 1 : 0;
    c.n_pos = found ? static_cast<int32_t>(s_npos) : 0;
    *out = c;
  }
  __syncthreads();
}

template <bool HESS>
__global__ void __launch_bounds__(256) k_scan(ScanParams p) {
  const LevelDesc lv = p.levels[p.level];
  const int fam_idx = blockIdx.x;
  if (fam_idx >= lv.num_families) return;
  const int fl = blockIdx.y;  // local feature
  const Family fam = p.families[fam_idx];
  const int f_global = p.f_begin + fl;
  const int b = threadIdx.x;
  const NodeRec direct = p.nodes[fam.direct];
  const size_t od = (static_cast<size_t>(fam.direct - lv.first_node) * p.f_count + fl) * kMaxBins + b;
  const long long cnt_d = p.hist_cnt[od];
  const unsigned long long sum_d = p.hist_sum[od];
  // hessian sums in units of h_pow2 * 2^-24; with h == 1 (h_pow2 = 1) a row contributes 2^24
  const unsigned long long hs_d =
      HESS ? (p.has_h ? p.hist_hsum[od] : (static_cast<unsigned long long>(cnt_d) << kQBits)) : 0ull;
  if (direct.candidate) {
    scan_node(p, direct, f_global, cnt_d,
              static_cast<long long>(sum_d) - cnt_d * static_cast<long long>(kQBias),
              static_cast<long long>(hs_d),
              &p.cand[static_cast<size_t>(fam.direct - lv.first_node) * p.f_count + fl]);
  }
  if (fam.derived >= 0) {
    const NodeRec derived = p.nodes[fam.derived];
    if (derived.candidate) {
      const LevelDesc plv = p.levels[p.level - 1];
      const size_t op = (static_cast<size_t>(fam.parent - plv.first_node) * p.f_count + fl) * kMaxBins + b;
      const long long cnt_x = static_cast<long long>(p.phist_cnt[op]) - cnt_d;
      const unsigned long long sum_x = p.phist_sum[op] - sum_d;
      const unsigned long long hs_x =
          HESS ? (p.has_h ? p.phist_hsum[op] - hs_d : (static_cast<unsigned long long>(cnt_x) << kQBits)) : 0ull;
      if (p.write_derived) {
        const size_t ox = (static_cast<size_t>(fam.derived - lv.first_node) * p.f_count + fl) * kMaxBins + b;
        p.hist_cnt[ox] = static_cast<uint32_t>(cnt_x);
        p.hist_sum[ox] = sum_x;
        if (HESS && p.has_h) p.hist_hsum[ox] = hs_x;
      }
      scan_node(p, derived, f_global, cnt_x,
                static_cast<long long>(sum_x) - cnt_x * static_cast<long long>(kQBias),
                static_cast<long long>(hs_x),
                &p.cand[static_cast<size_t>(fam.derived - lv.first_node) * p.f_count + fl]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_select_local: per node of the level, ordered arg-max over this shard's features
// (FindBestConditionConcurrentManager, training.cc:1728-1746: float scores, strict '>', candidate
// order = feature index order).
struct SelectParams {
  int level;
  LevelDesc* levels;
  Family* next_families;
  int32_t* next_slot_node;
  NodeRec* nodes;
  const Candidate* cand;
  int f_begin, f_count;
  const int32_t* na_bin;
  ShardBest* shard_best;       // [world][max level nodes] (this rank writes its row; exchange fills the rest)
  int rank, world, max_level_nodes;
  int min_examples, max_depth;
  int sibling_subtraction;
  int max_slots;               // capacity of one histogram pass
  DeviceState* st;
  int max_nodes;
};

__global__ void k_select_local(SelectParams p) {
  const LevelDesc lv = p.levels[p.level];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < lv.num_nodes; j += gridDim.x * blockDim.x) {
    ShardBest best{0.f, -1, 0, 0};
    if (p.nodes[lv.first_node + j].candidate) {
      float best_score = 0.f;  // NodeCondition.split_score default
      for (int fl = 0; fl < p.f_count; fl++) {
        const Candidate c = p.cand[static_cast<size_t>(j) * p.f_count + fl];
        if (c.found && c.score > best_score) {
          best_score = c.score;
          best = ShardBest{c.score, p.f_begin + fl, c.thr, c.n_pos};
        }
      }
    }
    p.shard_best[static_cast<size_t>(p.rank) * p.max_level_nodes + j] = best;
  }
}

// k_select_global: merges the shards' bests in rank order (== global feature order), applies the
// split to the node table, creates the children and lays out the next level.  One CTA.
__global__ void k_select_global(SelectParams p) {
  const LevelDesc lv = p.levels[p.level];
  __shared__ int s_first_child;
  for (int j = threadIdx.x; j < lv.num_nodes; j += blockDim.x) {
    NodeRec& nd = p.nodes[lv.first_node + j];
    ShardBest best{0.f, -1, 0, 0};
    if (nd.candidate) best = merge_shard_bests(p.shard_best, p.world, p.max_level_nodes, j);
    if (best.feature >= 0 && best.n_pos > 0 && best.n_pos < nd.n) {
      nd.feature = best.feature;
      nd.thr = best.thr;
      nd.na_value = (p.na_bin[best.feature] >= best.thr) ? 1 : 0;  // na_bin > thr - 1
      nd.score = best.score;
      nd.n_pos = best.n_pos;
    } else {
      nd.feature = -1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // Serial layout of the next level (<= a few thousand nodes).
    int next = lv.first_node + lv.num_nodes;
    const int first_next = next;
    int slots = 0, fams = 0;
    for (int j = 0; j < lv.num_nodes; j++) {
      NodeRec& nd = p.nodes[lv.first_node + j];
      if (nd.feature < 0) continue;
      if (next + 2 > p.max_nodes) { p.st->error_flag = 2; nd.feature = -1; continue; }
      const int pos = next, neg = next + 1;
      next += 2;
      nd.pos_child = pos;
      nd.neg_child = neg;
      NodeRec& cp = p.nodes[pos];
      NodeRec& cn = p.nodes[neg];
      cp = NodeRec{};
      cn = NodeRec{};
      cp.parent = cn.parent = lv.first_node + j;
      cp.depth = cn.depth = nd.depth + 1;
      cp.feature = cn.feature = -1;
      cp.pos_child = cp.neg_child = cn.pos_child = cn.neg_child = -1;
      cp.sibling = neg; cn.sibling = pos;
      cp.n = nd.n_pos;
      cn.n = nd.n - nd.n_pos;
      cp.slot = cn.slot = -1;
      // NodeTrain stop tests (training.cc:4909-4914).
      cp.candidate = (cp.n >= p.min_examples && cp.depth < p.max_depth) ? 1 : 0;
      cn.candidate = (cn.n >= p.min_examples && cn.depth < p.max_depth) ? 1 : 0;
      if (cp.candidate || cn.candidate) {
        if (p.sibling_subtraction) {
          // accumulate the smaller child from rows, derive the other one (exact integer subtraction)
          const bool pos_small = cp.n <= cn.n;
          const int small = pos_small ? pos : neg, large = pos_small ? neg : pos;
          if (slots < p.max_slots) {
            p.nodes[small].slot = slots;
            p.next_slot_node[slots] = small;
            p.next_families[fams++] = Family{lv.first_node + j, small, large};
            slots++;
          } else {
            p.st->error_flag = 3;
          }
        } else {
          for (int c = 0; c < 2; c++) {
            const int id = c == 0 ? pos : neg;
            if (!p.nodes[id].candidate) continue;
            if (slots < p.max_slots) {
              p.nodes[id].slot = slots;
              p.next_slot_node[slots] = id;
              p.next_families[fams++] = Family{-1, id, -1};
              slots++;
            } else {
              p.st->error_flag = 3;
            }
          }
        }
      }
    }
    LevelDesc nl;
    nl.first_node = first_next;
    nl.num_nodes = next - first_next;
    nl.num_slots = slots;
    nl.num_families = fams;
    p.levels[p.level + 1] = nl;
    p.st->num_nodes = next;
    s_first_child = first_next;
  }
  (void)s_first_child;
}

// ---------------------------------------------------------------------------------------------
// k_partition: SplitExamplesInPlace (training.cc:5243-5305 -> decision_tree.cc:957-1012) on the
// node-id representation.  One CTA iteration = one block of 8192 rows (16 consecutive rows per
// thread).  Every row of a split node is relabelled with its child id; the rows whose child is
// histogrammed at the next level are compacted, in row order, into the block's active list
// (block-wide exclusive scan = the stable scatter of example indices); the children's exact
// statistics are accumulated on the way.
struct PartParams {
  int64_t n;
  int n_blocks;
  int level;
  const LevelDesc* levels;
  NodeRec* nodes;
  const uint8_t* bins;
  int64_t n_pad;
  uint16_t* node_of_row;
  const uint32_t* q24;
  const uint32_t* hq24;   // hessian histogram only
  uint2* act;
  uint32_t* act_h;
  int32_t* act_count;
  const float* g;
  const float* h;   // null: h == 1
  const DeviceState* st;
  int smem_children;          // capacity of the shared accumulators (children of this level)
  int smem_children_private;  // capacity with one accumulator copy per lane
};

// Shared accumulators per child: cnt, g_lo, g_hi, h_lo, h_hi, g2_lo, g2_hi.
constexpr int kPartWords = 7;
constexpr int kPartThreads = 512;
constexpr int kPartRowsPerThread = kBlockRows / kPartThreads;  // 16

__device__ __forceinline__ void add64_smem(uint32_t* lo, uint32_t* hi, uint32_t v) {
  const uint32_t old = atomicAdd(lo, v);
  if (old + v < old) atomicAdd(hi, 1u);
}

__global__ void __launch_bounds__(kPartThreads) k_partition(PartParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ int s_warp_tot[kPartThreads / 32];
  const LevelDesc lv = p.levels[p.level];
  const LevelDesc nl = p.levels[p.level + 1];
  const int n_children = nl.num_nodes;
  // Accumulator layout: lane-private copies ([child][word][lane], bank == lane: no conflicts even
  // when a whole warp feeds the same two children, as at the top levels) while they fit, else one
  // shared copy, else global atomics.
  const bool use_priv = n_children <= p.smem_children_private;
  const bool use_smem = use_priv || n_children <= p.smem_children;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int copies = use_priv ? 32 : 1;
  if (use_smem) {
    for (int i = threadIdx.x; i < n_children * kPartWords * copies; i += blockDim.x) smem[i] = 0u;
    __syncthreads();
  }
  const float P = p.st->g_pow2;
  const double sscale = static_cast<double>(1u << (kSBits - 1)) / P;
  const double s2scale = static_cast<double>(1u << kSBits) / (static_cast<double>(P) * P);
  const double hscale = static_cast<double>(1u << kSBits) / p.st->h_pow2;
  for (int blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
    const int64_t base = static_cast<int64_t>(blk) * kBlockRows + static_cast<int64_t>(threadIdx.x) * kPartRowsPerThread;
    uint32_t out_info[kPartRowsPerThread];
    uint32_t active_mask = 0;
    if (nl.num_nodes > 0) {
#pragma unroll
      for (int j = 0; j < kPartRowsPerThread; j++) {
        const int64_t r = base + j;
        out_info[j] = 0u;
        if (r >= p.n) continue;
        const int node = p.node_of_row[r];
        if (node < lv.first_node) continue;       // row sits in a finished leaf
        const NodeRec& nd = p.nodes[node];
        if (nd.feature < 0) continue;             // node became a leaf at this level
        const uint32_t b = p.bins[static_cast<int64_t>(nd.feature) * p.n_pad + r];
        // EvalConditionDiscretizedHigher (decision_tree.cc:724-743); NA already folded into na_bin.
        const int child = (static_cast<int>(b) >= nd.thr) ? nd.pos_child : nd.neg_child;
        p.node_of_row[r] = static_cast<uint16_t>(child);
        const int slot = p.nodes[child].slot;
        if (slot >= 0) {
          active_mask |= 1u << j;
          out_info[j] = p.q24[r] | (static_cast<uint32_t>(slot) << 24);
        }
        const float g = p.g[r];
        const uint32_t qg = quant_biased_d(g, sscale, kSBias, 0x7FFFFFFFu);
        const uint32_t qg2 = quant_biased_d(g * g, s2scale, 0u, 0x7FFFFFFFu);
        const uint32_t qh = p.h ? quant_biased_d(p.h[r], hscale, 0u, 0x7FFFFFFFu) : 0u;
        const int c = child - nl.first_node;
        if (use_smem) {
          // word w of child c lives at (c*kPartWords + w) * copies + (private ? lane : 0)
          uint32_t* a = smem + static_cast<size_t>(c) * kPartWords * copies + (use_priv ? lane : 0);
          atomicAdd(&a[0], 1u);
          add64_smem(&a[1 * copies], &a[2 * copies], qg);
          if (p.h) add64_smem(&a[3 * copies], &a[4 * copies], qh);
          add64_smem(&a[5 * copies], &a[6 * copies], qg2);
        } else {
          NodeRec& cn = p.nodes[child];
          atomicAdd(&cn.sg, static_cast<unsigned long long>(qg));
          if (p.h) atomicAdd(&cn.sh, static_cast<unsigned long long>(qh));
          atomicAdd(&cn.sg2, static_cast<unsigned long long>(qg2));
        }
      }
    }
    // Block-wide exclusive scan of the per-thread active counts (thread order == row order).
    const int mine = __popc(active_mask);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int offset = incl - mine;
    int total = 0;
    for (int w = 0; w < kPartThreads / 32; w++) {
      if (w < warp) offset += s_warp_tot[w];
      total += s_warp_tot[w];
    }
    __syncthreads();
    if (threadIdx.x == 0) p.act_count[blk] = total;
    const int64_t obase = static_cast<int64_t>(blk) * kBlockRows;
#pragma unroll
    for (int j = 0; j < kPartRowsPerThread; j++) {
      if (active_mask & (1u << j)) {
        p.act[obase + offset] = make_uint2(out_info[j], static_cast<uint32_t>(threadIdx.x * kPartRowsPerThread + j));
        if (p.hq24 != nullptr) p.act_h[obase + offset] = p.hq24[base + j];
        offset++;
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int c = threadIdx.x; c < n_children; c += blockDim.x) {
      unsigned long long w[kPartWords];
#pragma unroll
      for (int k = 0; k < kPartWords; k++) {
        unsigned long long t = 0;
        const uint32_t* a = smem + (static_cast<size_t>(c) * kPartWords + k) * copies;
        for (int l = 0; l < copies; l++) t += a[(l + threadIdx.x) & (copies - 1)];  // staggered: fewer bank conflicts
        w[k] = t;
      }
      if (w[0] == 0ull) continue;
      NodeRec& cn = p.nodes[nl.first_node + c];
      atomicAdd(&cn.sg, (w[2] << 32) + w[1]);
      if (p.h) atomicAdd(&cn.sh, (w[4] << 32) + w[3]);
      atomicAdd(&cn.sg2, (w[6] << 32) + w[5]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_node_stats: fixed-point sums -> the doubles the reference stores, and the Newton leaf value.
// SetLeafValueWithNewtonRaphsonStep<false> (loss_utils.cc:49-132).
struct StatsParams {
  int level;            // nodes of level `level` are finalised (level 0: the root)
  const LevelDesc* levels;
  NodeRec* nodes;
  DeviceState* st;
  int use_hessian, logit_loss, has_h;
  float shrinkage, clamp;
  double l1, l2;
  int root_n_is;        // unused
  int64_t n_rows;
  int min_examples, max_depth;
};

__global__ void k_node_stats(StatsParams p) {
  if (p.level == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    // Root record (NodeTrain on the root, training.cc:4880-4894).
    NodeRec root{};
    root.parent = -1; root.depth = 1; root.feature = -1; root.pos_child = root.neg_child = -1;
    root.sibling = -1;
    root.n = p.n_rows;
    root.sg = p.st->root_sg; root.sh = p.st->root_sh; root.sg2 = p.st->root_sg2;
    root.candidate = (p.n_rows >= p.min_examples && 1 < p.max_depth) ? 1 : 0;
    root.slot = root.candidate ? 0 : -1;
    p.nodes[0] = root;
  }
  __syncthreads();
  const LevelDesc lv = p.levels[p.level];
  const double P = p.st->g_pow2;
  const double ginv = P / static_cast<double>(1u << (kSBits - 1));
  const double g2inv = P * P / static_cast<double>(1u << kSBits);
  const double hinv = static_cast<double>(p.st->h_pow2) / static_cast<double>(1u << kSBits);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < lv.num_nodes; j += gridDim.x * blockDim.x) {
    NodeRec& nd = p.nodes[lv.first_node + j];
    const double n = static_cast<double>(nd.n);
    const double sum_g = (static_cast<double>(static_cast<long long>(nd.sg)) - n * static_cast<double>(kSBias)) * ginv;
    double sum_h = p.has_h ? static_cast<double>(nd.sh) * hinv : n;
    const double sum_g2 = static_cast<double>(nd.sg2) * g2inv;
    if (sum_h <= kMinHessianForNewtonStep) sum_h = kMinHessianForNewtonStep;
    if (p.use_hessian) { nd.stat[0] = sum_g; nd.stat[1] = sum_h; nd.stat[2] = n; }
    else { nd.stat[0] = sum_g; nd.stat[1] = sum_g2; nd.stat[2] = n; }
    const double numerator = l1_threshold_d(sum_g, p.l1);
    const double denominator = sum_h + p.l2;
    float value = static_cast<float>(static_cast<double>(p.shrinkage) * numerator / denominator);
    if (p.logit_loss) value = fminf(fmaxf(value, -p.clamp), p.clamp);
    nd.leaf_value = value;
  }
}

// Resets the per-iteration scalars and the level table.
__global__ void k_begin_iteration(DeviceState* st, LevelDesc* levels, Family* fam0, int32_t* slot_node0,
                                  int root_candidate) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->gmax_bits = 0u;
    st->root_sg = st->root_sh = st->root_sg2 = 0ull;
    st->num_nodes = 1;
    levels[0] = LevelDesc{0, 1, root_candidate ? 1 : 0, root_candidate ? 1 : 0};
    fam0[0] = Family{-1, 0, -1};
    slot_node0[0] = 0;
  }
}

__global__ void k_reset_loss(DeviceState* st) {
  st->loss_sum = 0.0;
  st->correct = 0ull;
}

// ---------------------------------------------------------------------------------------------
// Debug / seam kernels.

// SplitExamplesInPlace as a standalone stable partition of a row-id list (single CTA per 2048-row
// tile + decoupled offsets are overkill for a test seam: two-pass count/scatter with a global scan).
__global__ void k_partition_count(const uint8_t* col, const uint32_t* rows, int64_t n, int thr,
                                  uint32_t* block_pos_counts) {
  __shared__ uint32_t s[8];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool pos = i < n && col[rows[i]] >= thr;
  const uint32_t bal = __ballot_sync(0xffffffffu, pos);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < 8; k++) t += s[k];
    block_pos_counts[blockIdx.x] = t;
  }
}

__global__ void k_partition_scatter(const uint8_t* col, const uint32_t* rows, int64_t n, int thr,
                                    const uint32_t* block_pos_offsets, uint32_t total_pos,
                                    uint32_t* out) {
  __shared__ uint32_t s[8];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool in = i < n;
  const uint32_t row = in ? rows[i] : 0u;
  const bool pos = in && col[row] >= thr;
  const uint32_t bal = __ballot_sync(0xffffffffu, pos);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = __popc(bal);
  __syncthreads();
  uint32_t warp_off = 0;
  for (int k = 0; k < w; k++) warp_off += s[k];
  const uint32_t pos_before = block_pos_offsets[blockIdx.x] + warp_off + __popc(bal & ((1u << lane) - 1u));
  if (in) {
    if (pos) out[pos_before] = row;
    else out[total_pos + (static_cast<uint32_t>(i) - pos_before)] = row;
  }
}

}  // namespace ygg
