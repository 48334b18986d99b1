// ygg_hist.cuh — k_hist, the hot kernel: FillExampleBucketSet (learner/decision_tree/
// splitter_scanner.h:859-909) for every open node x feature of one tree level.
//
// Data in HBM:
//   bins[f][row]            one byte per value, column-major, row stride n_pad (multiple of 8192)
//   act_info[block][k]      per 8192-row block, the compacted list of ACTIVE rows of this level
//                           (rows whose node's histogram is accumulated from rows, i.e. not a leaf
//                           and not the sibling derived by subtraction): q24 | slot << 24
//   act_ridx[block][k]      the row's offset inside its block (uint16)
//   act_count[block]        number of active rows of the block
// The lists are written by k_quantize (root) / k_partition (deeper levels), in ascending row order.
//
// Work item = (chunk of consecutive row blocks) x (group of G consecutive features).  Per item a
// CTA zeroes G*S*256 shared-memory bins, streams the chunk block by block — the G x 8192-byte bins
// tile of each block is staged into shared memory by the TMA engine (cp.async.bulk, mbarrier
// pipeline, kStages deep) — and flushes its non-empty bins to the 64-bit global histogram.
//
// Shared-memory bin = two 32-bit words updated with native ATOMS.ADD (the only shared-memory
// atomic add sm_100a executes natively; 64-bit and float adds compile to CAS loops):
//   word0 = count (bits 0..19) + carries of the sum (bits 20..31), word1 = low 32 bits of sum(q24).
// A chunk has < 2^20 rows, so count < 2^20 and carries <= count * 2^24 / 2^32 < 2^12: the pair is
// an exact 44-bit sum.  A carry is detected from the value ATOMS.ADD returns; it happens once per
// ~256 updates of a bin and is handled off the hot path.
//
// Because every lane of a warp works on an ACTIVE row, the inner loop has no predication and no
// divergence, and sibling subtraction halves the number of shared-memory atomics, which — not HBM
// bandwidth — is what bounds this kernel (profiles/atoms_microbench_r01.txt).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ygg_device.cuh"

namespace ygg {

#ifndef YGG_HIST_THREADS
#define YGG_HIST_THREADS 512
#endif
#ifndef YGG_HIST_UNROLL
#define YGG_HIST_UNROLL 2
#endif
constexpr int kHistThreads = YGG_HIST_THREADS;
constexpr int kBlockRows = 8192;                 // rows per block (TMA tile = G x 8192 bytes)
constexpr int kHistStages = 3;
constexpr int kHistCntBits = 20;
constexpr int kHistMaxChunkBlocks = ((1 << kHistCntBits) - 1) / kBlockRows;  // 127 blocks = 1,040,384 rows
constexpr int kHistUnroll = YGG_HIST_UNROLL;                   // active rows per thread per inner iteration

struct HistParams {
  const uint8_t* bins;
  int64_t n_pad;
  const uint32_t* act_info;
  const uint32_t* act_h;      // hq24 per active row (hessian histogram only)
  const uint16_t* act_ridx;
  const int32_t* act_count;
  int n_blocks;
  int f_begin;        // first feature (dataset index) of this shard
  int f_count;        // features in this shard
  int G;              // features per work item
  int S;              // shared-memory slots (>= slots used at this level)
  int chunk_blocks;   // row blocks per work item (<= kHistMaxChunkBlocks)
  int level;
  const LevelDesc* levels;
  const int32_t* slot_node;   // [S] node id owning slot s at this level
  unsigned long long* hist_sum;   // [level nodes][f_count][256]
  uint32_t* hist_cnt;
  unsigned long long* hist_hsum;  // hessian histogram only
};

// ---- PTX wrappers: mbarrier + TMA bulk copy (SASS: SYNCS / UBLKCP) and shared-memory atomics ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void smem_red(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;\n" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t smem_ld_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}

// Shared-memory layout (dynamic):  [hist planes][kHistStages x G x 8192 B tiles][mbarriers]
__host__ __device__ inline size_t hist_smem_bytes(int G, int S, bool hess) {
  const size_t planes = hess ? 4 : 2;
  return planes * G * static_cast<size_t>(S) * kMaxBins * 4 + static_cast<size_t>(kHistStages) * G * kBlockRows +
         2 * kHistStages * 8 + 16;
}

template <bool HESS>
__global__ void __launch_bounds__(kHistThreads, 1) k_hist(HistParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const LevelDesc lv = p.levels[p.level];
  if (lv.num_slots == 0) return;
  const int S = p.S;
  const int G = p.G;
  const int words_per_feature = S * kMaxBins;
  constexpr int planes = HESS ? 4 : 2;
  const int plane_words = G * words_per_feature;  // plane layout [plane][G][S][256]: cnt | lo | hlo | hhi
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);
  const uint32_t s_hist = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t plane_bytes = static_cast<uint32_t>(plane_words) * 4u;
  const uint32_t s_tiles = s_hist + static_cast<uint32_t>(planes) * plane_bytes;
  const uint32_t stage_bytes = static_cast<uint32_t>(G) * kBlockRows;
  const uint32_t s_full = s_tiles + kHistStages * stage_bytes;  // kHistStages full barriers
  const uint32_t s_empty = s_full + kHistStages * 8;            // kHistStages empty barriers

  const int tid = threadIdx.x;
  const int n_fgroups = (p.f_count + G - 1) / G;
  const int n_chunks = (p.n_blocks + p.chunk_blocks - 1) / p.chunk_blocks;
  const int64_t n_items = static_cast<int64_t>(n_chunks) * n_fgroups;

  if (tid == 0) {
    for (int s = 0; s < kHistStages; s++) {
      mbar_init(s_full + 8 * s, 1);
      mbar_init(s_empty + 8 * s, kHistThreads / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // Pipeline bookkeeping persists across work items: `produced` / `consumed` count tiles.
  uint32_t produced = 0, consumed = 0;

  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int chunk = static_cast<int>(item / n_fgroups);
    const int fg = static_cast<int>(item - static_cast<int64_t>(chunk) * n_fgroups);
    const int f0 = fg * G;
    const int gcount = min(G, p.f_count - f0);
    const int b0 = chunk * p.chunk_blocks;
    const int b1 = min(b0 + p.chunk_blocks, p.n_blocks);
    const int nb = b1 - b0;

    {
      uint4* z = reinterpret_cast<uint4*>(hist);
      const int n4 = planes * plane_words / 4;
      for (int i = tid; i < n4; i += kHistThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    // producer (thread 0): issue the TMA copies of tile `t` of this item
    auto issue = [&](int t) {
      const uint32_t s = produced % kHistStages;
      const uint32_t use = produced / kHistStages;
      if (use > 0) mbar_wait(s_empty + 8 * s, (use - 1) & 1u);  // consumers released the slot
      mbar_expect_tx(s_full + 8 * s, static_cast<uint32_t>(gcount) * kBlockRows);
      const uint8_t* src = p.bins + static_cast<int64_t>(p.f_begin + f0) * p.n_pad + static_cast<int64_t>(b0 + t) * kBlockRows;
      for (int gi = 0; gi < gcount; gi++)
        tma_bulk_g2s(s_tiles + s * stage_bytes + gi * kBlockRows, src + static_cast<int64_t>(gi) * p.n_pad, kBlockRows, s_full + 8 * s);
      produced++;
    };
    if (tid == 0) {
      const int pre = min(nb, kHistStages - 1);
      for (int t = 0; t < pre; t++) issue(t);
    }

    for (int t = 0; t < nb; t++) {
      if (tid == 0 && t + kHistStages - 1 < nb) issue(t + kHistStages - 1);
      const uint32_t s = consumed % kHistStages;
      const uint32_t parity = (consumed / kHistStages) & 1u;
      const int blk = b0 + t;
      const int n_act = p.act_count[blk];
      const uint32_t* info_p = p.act_info + static_cast<int64_t>(blk) * kBlockRows;
      const uint32_t* h_p = HESS ? p.act_h + static_cast<int64_t>(blk) * kBlockRows : nullptr;
      const uint16_t* ridx_p = p.act_ridx + static_cast<int64_t>(blk) * kBlockRows;
      mbar_wait(s_full + 8 * s, parity);
      const uint32_t tile = s_tiles + s * stage_bytes;
      for (int e0 = tid; e0 < n_act; e0 += kHistThreads * kHistUnroll) {
        uint32_t info[kHistUnroll], ridx[kHistUnroll], hq[kHistUnroll];
        bool ok[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; u++) {
          const int e = e0 + u * kHistThreads;
          ok[u] = e < n_act;
          info[u] = ok[u] ? __ldg(info_p + e) : 0u;
          ridx[u] = ok[u] ? __ldg(ridx_p + e) : 0u;
          hq[u] = (HESS && ok[u]) ? __ldg(h_p + e) : 0u;
        }
        if (ok[kHistUnroll - 1]) {
          // fast path: all kHistUnroll rows valid, no predication anywhere
          for (int gi = 0; gi < gcount; gi++) {
            const uint32_t fbase = s_hist + static_cast<uint32_t>(gi * words_per_feature) * 4u;
            const uint32_t tbase = tile + gi * kBlockRows;
            uint32_t addr[kHistUnroll], old[kHistUnroll], hold[kHistUnroll];
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) {
              const uint32_t b = smem_ld_u8(tbase + ridx[u]);
              addr[u] = fbase + ((((info[u] >> 24) << 8) | b) << 2);
            }
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) {
              smem_red(addr[u], 1u);
              old[u] = smem_add(addr[u] + plane_bytes, info[u] & kQMax);
              if (HESS) hold[u] = smem_add(addr[u] + 2u * plane_bytes, hq[u]);
            }
            bool carry = false;
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) {
              const uint32_t q = info[u] & kQMax;
              carry |= (old[u] + q < old[u]);
              if (HESS) carry |= (hold[u] + hq[u] < hold[u]);
            }
            if (carry) {
#pragma unroll
              for (int u = 0; u < kHistUnroll; u++) {
                const uint32_t q = info[u] & kQMax;
                if (old[u] + q < old[u]) smem_red(addr[u], 1u << kHistCntBits);
                if (HESS) {
                  if (hold[u] + hq[u] < hold[u]) smem_red(addr[u] + 3u * plane_bytes, 1u);
                }
              }
            }
          }
        } else {
          // tail of the block's active list
          for (int gi = 0; gi < gcount; gi++) {
            const uint32_t fbase = s_hist + static_cast<uint32_t>(gi * words_per_feature) * 4u;
            const uint32_t tbase = tile + gi * kBlockRows;
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) {
              if (!ok[u]) continue;
              const uint32_t b = smem_ld_u8(tbase + ridx[u]);
              const uint32_t a = fbase + ((((info[u] >> 24) << 8) | b) << 2);
              const uint32_t q = info[u] & kQMax;
              smem_red(a, 1u);
              const uint32_t o = smem_add(a + plane_bytes, q);
              if (o + q < o) smem_red(a, 1u << kHistCntBits);
              if (HESS) {
                const uint32_t ho = smem_add(a + 2u * plane_bytes, hq[u]);
                if (ho + hq[u] < ho) smem_red(a + 3u * plane_bytes, 1u);
              }
            }
          }
        }
      }
      // release the tile: one arrival per warp
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(s_empty + 8 * s);
      consumed++;
    }
    __syncthreads();
    // Flush non-empty bins to the global 64-bit histogram.
    const int used = lv.num_slots * kMaxBins;
    const uint32_t* s_cnt = hist;
    const uint32_t* s_lo = hist + plane_words;
    const uint32_t* s_hlo = hist + 2 * plane_words;
    const uint32_t* s_hhi = hist + 3 * plane_words;
    for (int gi = 0; gi < gcount; gi++) {
      const int f_local = f0 + gi;
      for (int i = tid; i < used; i += kHistThreads) {
        const uint32_t c = s_cnt[gi * words_per_feature + i];
        if (c != 0u) {
          const int s = i >> 8, b = i & 0xFF;
          const int j = p.slot_node[s] - lv.first_node;
          const size_t o = (static_cast<size_t>(j) * p.f_count + f_local) * kMaxBins + b;
          const unsigned long long sum =
              (static_cast<unsigned long long>(c >> kHistCntBits) << 32) + s_lo[gi * words_per_feature + i];
          atomicAdd(&p.hist_sum[o], sum);
          atomicAdd(&p.hist_cnt[o], c & ((1u << kHistCntBits) - 1u));
          if (HESS) {
            const unsigned long long hsum =
                (static_cast<unsigned long long>(s_hhi[gi * words_per_feature + i]) << 32) +
                s_hlo[gi * words_per_feature + i];
            atomicAdd(&p.hist_hsum[o], hsum);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace ygg
