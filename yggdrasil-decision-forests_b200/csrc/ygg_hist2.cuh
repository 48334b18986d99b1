// ygg_hist2.cuh — k_hist2: FillExampleBucketSet (learner/decision_tree/splitter_scanner.h:859-909) for the SHALLOW
// levels of a tree (<= 4 histogram slots), with the lanes of a warp mapped to FEATURES instead of rows.
//
// Why a second kernel.  k_hist (ygg_hist.cuh) gives every lane its own row of ONE feature; the 32 bins a warp
// updates at once are random, so its shared-memory atomics hit random banks: ~3.45 wavefronts per ATOMS, which — not
// the number of atomic instructions — is what bounds it (profiles/k_hist_ncu_r01.md: 1.9e8 atomic wavefronts for
// 5.4e7 ATOMS; tools/hist_loop_bench.cu: 3.96 elements/clk/SM).  Here a warp takes ONE row (two / four rows) and its
// 32 (16 / 8) lanes take 32 consecutive features of it; the histogram of the work item is laid out
// [slot][bin][feature], so lane == bank for every atomic: one wavefront, 8.2 elements/clk/SM in the same benchmark.
// The layout needs S * 256 * FL * 8 bytes of shared memory for FL features at once, which is why it stops at S = 4.
//
// Data in HBM:
//   bins4[group][row]       u32: the bins of features 4*group .. 4*group+3 of one row (feature 4*group in the low
//                           byte) — the interleaved copy of the matrix that k_interleave4 builds once per dataset.
//                           A tile (R rows x GQ groups) is GQ contiguous runs of R words: one TMA bulk copy each.
//   act / act_count         the compacted active rows of every 8192-row block, as for k_hist
//   act_sub[block][8]       the number of active rows before each 1024-row sub-tile of the block (k_partition)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ygg_device.cuh"
#include "ygg_hist.cuh"

namespace ygg {

constexpr int kSubRows = 1024;                        // granularity of act_sub and of the tiles
constexpr int kSubPerBlock = kBlockRows / kSubRows;   // 8
constexpr int kHist2Stages = 2;
constexpr int kHist2Threads = 1024;
constexpr int kHist2TilePad = 4;                      // words between the groups of a tile: bank = 4 * group + row

struct Hist2Params {
  const uint32_t* bins4;      // [groups of the dataset][n_pad]
  int64_t n_pad, n;
  const uint32_t* q24;        // root: quantised gradient of every row
  const uint2* act;
  const int32_t* act_count;
  const int32_t* act_sub;     // [n_blocks][8]
  int n_blocks;
  int f_begin, f_count;       // features histogrammed by this rank
  int g_begin, n_groups;      // interleave groups covering them
  int S, T;                   // slots of the level; sub-tiles (1024 rows) per tile
  int chunk_blocks, level;
  const LevelDesc* levels;
  unsigned long long* hist_sum;
  uint32_t* hist_cnt;
  int f_chunk;
  long long chunk_stride;
};

// FL = features per work item = lanes per row (32, 16 or 8).
// Shared memory: [histogram: S*256 bins x 2 words x FL][tiles: stages x (FL/4 groups x (R + 4) words + root: R words of q24)]
//                [entry staging: 32 warps x 32 entries x 8 B][mbarriers]
__host__ __device__ inline size_t hist2_stage_bytes(int FL, int T, bool root) {
  const size_t R = static_cast<size_t>(T) * kSubRows;
  return (FL / 4) * (R + kHist2TilePad) * 4 + (root ? R * 4 : 0);
}
__host__ __device__ inline size_t hist2_smem_bytes(int FL, int S, int T, bool root) {
  const size_t hist = 2ull * S * kMaxBins * FL * 4;
  return hist + kHist2Stages * hist2_stage_bytes(FL, T, root) + (kHist2Threads / 32) * 32 * 8 + 2 * kHist2Stages * 8 + 16;
}

// Builds the interleaved copy: 4 consecutive rows x 4 features per thread (byte transposes in registers).
__global__ void __launch_bounds__(256) k_interleave4(const uint8_t* __restrict__ bins, int64_t n_pad, int F, uint32_t* __restrict__ out) {
  const int g = blockIdx.y;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r4 < n_pad / 4; r4 += stride) {
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int f = 4 * g + k;
      c[k] = f < F ? *reinterpret_cast<const uint32_t*>(bins + static_cast<int64_t>(f) * n_pad + 4 * r4) : 0u;
    }
    // c[k] = bytes of feature k for rows 0..3 -> w[r] = bytes of row r for features 0..3
    const uint32_t t0 = __byte_perm(c[0], c[1], 0x5140), t1 = __byte_perm(c[0], c[1], 0x7362);   // (f0r0 f1r0 f0r1 f1r1), rows 2,3
    const uint32_t u0 = __byte_perm(c[2], c[3], 0x5140), u1 = __byte_perm(c[2], c[3], 0x7362);
    uint4 w;
    w.x = __byte_perm(t0, u0, 0x5410);
    w.y = __byte_perm(t0, u0, 0x7632);
    w.z = __byte_perm(t1, u1, 0x5410);
    w.w = __byte_perm(t1, u1, 0x7632);
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(g) * n_pad + 4 * r4) = w;
  }
}

// The inner loop is bound by the number of shared-memory instructions the SM retires (~0.78 per clock, measured with
// tools/hist_loop_bench.cu) and by plain instruction issue, so it is pared down to, per warp instruction (= one row x
// FL features): one broadcast LDS.128 of a pre-digested entry, one LDS.32 of the row's group word, PRMT, two adds and
// the two atomics — no predicates: the tail of a batch is padded with entries whose addends are zero, and lanes of
// features outside the shard accumulate into columns that are never flushed.
template <int FL, bool ROOT>
__global__ void __launch_bounds__(kHist2Threads, 1) k_hist2(Hist2Params p) {
  constexpr int RL = 32 / FL;          // rows per warp instruction
  constexpr int GQ = FL / 4;           // interleave groups per work item
  constexpr int U = 4;                 // warp instructions in flight per inner iteration
  constexpr uint32_t kBinStride = 2u * FL * 4u;   // bytes between consecutive bins: [bin][plane][feature]
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ int s_off[kHistMaxChunkBlocks * (kSubPerBlock + 1) + 1];
  const LevelDesc lv = p.levels[p.level];
  if (lv.num_slots == 0) return;
  const int S = p.S;
  const int R = p.T * kSubRows;                       // rows per tile
  const int tiles_per_block = kSubPerBlock / p.T;
  const uint32_t hist_words = 2u * static_cast<uint32_t>(S) * kMaxBins * FL;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t s_hist = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  asm volatile("mov.u32 %0, %0;" : "+r"(s_hist));   // keep the window base in a register (see k_hist)
  const uint32_t s_tiles = s_hist + hist_words * 4u;
  const uint32_t group_bytes = static_cast<uint32_t>(R + kHist2TilePad) * 4u;
  const uint32_t stage_bytes = static_cast<uint32_t>(hist2_stage_bytes(FL, p.T, ROOT));
  const uint32_t s_entries = s_tiles + kHist2Stages * stage_bytes;
  const uint32_t s_full = s_entries + (kHist2Threads / 32) * 32 * 8;
  const uint32_t s_empty = s_full + kHist2Stages * 8;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int fi = lane % FL, rsel = lane / FL;
  const uint32_t byte_sel = 0x4440u | static_cast<uint32_t>(fi & 3);         // PRMT: byte (fi & 3) of the group word, zero-extended
  const uint32_t lane_tile = static_cast<uint32_t>(fi >> 2) * group_bytes;   // this lane's group inside a stage
  const uint32_t lane_hist = s_hist + static_cast<uint32_t>(fi) * 4u;
  const uint32_t my_entries = s_entries + static_cast<uint32_t>(warp) * 256u;
  const int n_fsets = (p.n_groups + GQ - 1) / GQ;
  const int n_chunks = (p.n_blocks + p.chunk_blocks - 1) / p.chunk_blocks;
  const int64_t n_items = static_cast<int64_t>(n_chunks) * n_fsets;

  if (tid == 0) {
    for (int s = 0; s < kHist2Stages; s++) {
      mbar_init(s_full + 8 * s, 1);
      mbar_init(s_empty + 8 * s, kHist2Threads / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t produced = 0, consumed = 0;

  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int chunk = static_cast<int>(item / n_fsets);
    const int fset = static_cast<int>(item - static_cast<int64_t>(chunk) * n_fsets);
    const int g0 = fset * GQ;                                  // first group of the set, relative to g_begin
    const int gcount = min(GQ, p.n_groups - g0);
    const int b0 = chunk * p.chunk_blocks;
    const int nb = min(b0 + p.chunk_blocks, p.n_blocks) - b0;
    const int n_tiles = nb * tiles_per_block;

    {
      uint4* z = reinterpret_cast<uint4*>(hist);
      const int n4 = static_cast<int>(hist_words / 4u);
      for (int i = tid; i < n4; i += kHist2Threads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // active-row offsets of the chunk's sub-tiles: s_off[b * 9 + k], k = 0..7 sub-tile starts, 8 = block total
    if (!ROOT) {
      for (int i = tid; i < nb * (kSubPerBlock + 1); i += kHist2Threads) {
        const int b = i / (kSubPerBlock + 1), k = i - b * (kSubPerBlock + 1);
        s_off[i] = k < kSubPerBlock ? p.act_sub[static_cast<int64_t>(b0 + b) * kSubPerBlock + k] : p.act_count[b0 + b];
      }
    }
    __syncthreads();

    auto issue = [&](int t) {   // producer (thread 0): the TMA copies of tile t of this item
      const uint32_t s = produced % kHist2Stages;
      const uint32_t use = produced / kHist2Stages;
      if (use > 0) mbar_wait(s_empty + 8 * s, (use - 1) & 1u);
      mbar_expect_tx(s_full + 8 * s, static_cast<uint32_t>(gcount + (ROOT ? 1 : 0)) * R * 4u);
      const int blk = b0 + t / tiles_per_block, st = t % tiles_per_block;
      const int64_t row0 = static_cast<int64_t>(blk) * kBlockRows + static_cast<int64_t>(st) * R;
      for (int gl = 0; gl < gcount; gl++)
        tma_bulk_g2s(s_tiles + s * stage_bytes + gl * group_bytes, p.bins4 + static_cast<int64_t>(p.g_begin + g0 + gl) * p.n_pad + row0,
                     static_cast<uint32_t>(R) * 4u, s_full + 8 * s);
      if (ROOT)   // the quantised gradients of the tile's rows (zero past n: k_quantize never writes the padding)
        tma_bulk_g2s(s_tiles + s * stage_bytes + GQ * group_bytes, p.q24 + row0, static_cast<uint32_t>(R) * 4u, s_full + 8 * s);
      produced++;
    };
    if (tid == 0 && n_tiles > 0) issue(0);

    // Entries of tile t: [e0, e1) of its block's active list (ROOT: the rows of the tile themselves).
    auto tile_range = [&](int t, int* e0, int* e1, int64_t* base, int* row0_in_block) {
      const int b = t / tiles_per_block, st = t % tiles_per_block;
      *row0_in_block = st * R;
      *base = static_cast<int64_t>(b0 + b) * kBlockRows;
      if (ROOT) {
        const int64_t left = p.n - (*base + *row0_in_block);
        *e0 = 0;
        *e1 = static_cast<int>(left < 0 ? 0 : (left > R ? R : left));
      } else {
        const int* o = s_off + b * (kSubPerBlock + 1);
        *e0 = o[st * p.T];
        *e1 = o[st * p.T + p.T];   // (st + 1) * T == 8 -> the block total
      }
    };
    // One batch = 32 entries of a warp, as 4 runs of 8 consecutive entries: run (r * 4 + lane / 8) * 32 + warp of the
    // tile's list, so that the 32 warps share a tile's entries 8 at a time (a tile of 2048 rows holds ~1000 active
    // rows: whole 32-entry batches would leave warps idle).  Entry = (q24 | slot << 24, row offset in the TILE);
    // y = kNoEntry past the end.
    constexpr uint32_t kNoEntry = 0xFFFFFFFFu;
    auto load_batch = [&](int t, int r) -> uint2 {
      int e0, e1, row0;
      int64_t base;
      tile_range(t, &e0, &e1, &base, &row0);
      const int e = e0 + (((r * 4 + (lane >> 3)) * (kHist2Threads / 32) + warp) << 3) + (lane & 7);
      if (e >= e1) return make_uint2(0u, kNoEntry);
      if (ROOT) return make_uint2(__ldg(p.q24 + base + row0 + e), static_cast<uint32_t>(e));
      const uint2 v = __ldg(p.act + base + e);
      return make_uint2(v.x, v.y - static_cast<uint32_t>(row0));
    };

    if constexpr (ROOT && FL == 32) {
      // Root: every row is active and sits in slot 0.  Four consecutive rows per step: one LDS.128 brings their four
      // gradients (the same 16 bytes for every lane), one LDS.128 their four group words (the 4 lanes of a group share
      // them: 8 x 16 bytes = one wavefront), then one returning atomic per row and feature (carry -> plane 1).
      for (int t = 0; t < n_tiles; t++) {
        if (tid == 0 && t + 1 < n_tiles) issue(t + 1);
        const uint32_t s = consumed % kHist2Stages;
        const uint32_t parity = (consumed / kHist2Stages) & 1u;
        mbar_wait(s_full + 8 * s, parity);
        const uint32_t tile = s_tiles + s * stage_bytes + lane_tile;
        const uint32_t qtile = s_tiles + s * stage_bytes + GQ * group_bytes;
#pragma unroll 2
        for (int m = warp; m < R / 4; m += kHist2Threads / 32) {
          uint32_t q[4], w[4];
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(qtile + m * 16u));
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(tile + m * 16u));
          uint32_t a[4], old[4];
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] = __byte_perm(w[k], 0u, byte_sel) * kBinStride + lane_hist;
#pragma unroll
          for (int k = 0; k < 4; k++) old[k] = smem_add(a[k], q[k]);
          bool carry = false;
#pragma unroll
          for (int k = 0; k < 4; k++) carry |= (old[k] + q[k] < old[k]);
          if (carry) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (old[k] + q[k] < old[k]) smem_red(a[k] + FL * 4u, 1u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + 8 * s);
        consumed++;
      }
    } else {
    uint2 cur = n_tiles > 0 ? load_batch(0, 0) : make_uint2(0u, kNoEntry);
    for (int t = 0; t < n_tiles; t++) {
      if (tid == 0 && t + 1 < n_tiles) issue(t + 1);
      const uint32_t s = consumed % kHist2Stages;
      const uint32_t parity = (consumed / kHist2Stages) & 1u;
      int e0, e1, row0;
      int64_t base;
      tile_range(t, &e0, &e1, &base, &row0);
      const int n_ent = e1 - e0;
      mbar_wait(s_full + 8 * s, parity);
      const uint32_t tile = s_tiles + s * stage_bytes + lane_tile;
      int r = 0;
      while (true) {
        const bool more = (r + 1) * kHist2Threads < n_ent;   // a round covers 32 warps x 32 entries
        const uint2 nxt = more ? load_batch(t, r + 1) : (t + 1 < n_tiles ? load_batch(t + 1, 0) : make_uint2(0u, kNoEntry));
        // entries of a batch are ordered by lane: the loop stops after the last valid one
        const int jmax = 32 - __clz(static_cast<int>(__ballot_sync(0xffffffffu, cur.y != kNoEntry)));
        if (jmax > 0) {
          // stage this lane's entry for the whole warp: (q24 | slot << 24, byte offset of the row in a tile group);
          // past the end: a zero gradient on row 0 (its count | coarse addend is masked below)
          {
            const bool has = cur.y != kNoEntry;
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(my_entries + lane * 8u), "r"(has ? cur.x : 0u),
                         "r"(has ? cur.y * 4u : 0x80000000u)
                         : "memory");
          }
          __syncwarp();
#pragma unroll 1
          for (int j = 0; j < jmax; j += RL * U) {
            uint32_t ex[U], ey[U], w[U];
#pragma unroll
            for (int u = 0; u < U; u++)
              asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ex[u]), "=r"(ey[u]) : "r"(my_entries + static_cast<uint32_t>(j + u * RL + rsel) * 8u));
#pragma unroll
            for (int u = 0; u < U; u++) w[u] = smem_ld_u32(tile + (ey[u] & 0x7FFFFFFFu));
#pragma unroll
            for (int u = 0; u < U; u++) {
              const uint32_t q = ex[u] & kQMax;
              const uint32_t a = __byte_perm(w[u], 0u, byte_sel) * kBinStride + (lane_hist + (ex[u] >> 24) * (kMaxBins * kBinStride));
              if (ROOT) {   // plane 0 = low word of the sum, plane 1 = carries
                const uint32_t old = smem_add(a, q);
                if (old + q < old) smem_red(a + FL * 4u, 1u);
              } else {      // kHistPacked words: plane 0 = count | coarse sum, plane 1 = sum mod 2^32
                // (q >> 18) << 13 | 1, and 0 for a padding entry (bit 31 of its offset word is set)
                smem_red(a, (((ex[u] >> (kPackedCoarseShift - kPackedCntBits)) & (0x3Fu << kPackedCntBits)) | 1u) & ~static_cast<uint32_t>(static_cast<int32_t>(ey[u]) >> 31));
                smem_red(a + FL * 4u, q);
              }
            }
          }
          __syncwarp();
        }
        cur = nxt;
        if (!more) break;
        r++;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + 8 * s);
      consumed++;
    }
    }
    __syncthreads();
    // Flush: word (slot * 256 + bin) * 2 * FL + plane * FL + feature.  Consecutive threads take consecutive bins of ONE
    // feature, so that the global atomics of a warp fall into one 256-byte run (the shared-memory reads are
    // bank-conflicted instead; they are 16 K words per item).
    const int used = lv.num_slots * kMaxBins;
    for (int i = tid; i < used * FL; i += kHist2Threads) {
      const int ffi = i / used, sb = i - ffi * used;
      const int f = 4 * (p.g_begin + g0) + ffi;
      if (f < p.f_begin || f >= p.f_begin + p.f_count) continue;
      const uint32_t w0 = hist[sb * 2 * FL + ffi], w1 = hist[sb * 2 * FL + FL + ffi];
      size_t oc;
      const size_t o = slot_hist_offset(sb >> 8, f - p.f_begin, sb & 0xFF, p.f_chunk, p.chunk_stride, &oc);
      if (ROOT) {   // the counts are precomputed (k_root_counts)
        const unsigned long long sum = (static_cast<unsigned long long>(w1) << 32) + w0;
        if (sum != 0ull) atomicAdd(&p.hist_sum[o], sum);
      } else if (w0 != 0u) {   // kHistPacked words (ygg_hist.cuh)
        const unsigned long long b64 = static_cast<unsigned long long>(w0 >> kPackedCntBits) << kPackedCoarseShift;
        atomicAdd(&p.hist_sum[o], b64 + static_cast<uint32_t>(w1 - static_cast<uint32_t>(b64)));
        atomicAdd(&p.hist_cnt[oc], w0 & kPackedMaxUpdates);
      }
    }
    __syncthreads();
  }
}

}  // namespace ygg
