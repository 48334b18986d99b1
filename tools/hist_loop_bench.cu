// hist_loop_bench.cu — the inner loop of k_hist in isolation, to decide between the two lane mappings:
//   MODE 0  row per lane     : 32 lanes = 32 rows of ONE feature; bins come from a byte tile (LDS.U8), the two words of a
//                              bin live in planes [cnt][sum] -> banks are random (what k_hist does today)
//   MODE 1  feature per lane : 32 lanes = 32 features of ONE row; tile is [row][32 features], histogram is
//                              [bin][32 features] -> bank == lane for the tile read and for both atomics
//   MODE 2  feature per lane, 16 features x 2 rows per warp instruction ([bin][16 features], 2 slots)
//   MODE 3  8 features x 4 rows ; MODE 4  4 features x 8 rows ; MODE 5  2 features x 16 rows
// Two non-returning atomics per element (kHistPacked addends).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kTile = 8192;   // bytes of bins staged in shared memory

template <int MODE>
__global__ void __launch_bounds__(1024) loop(const uint8_t* bins, uint32_t* out, int iters, int hist_words, long long* cycles) {
  extern __shared__ uint32_t sm[];
  uint32_t* hist = sm;                         // 2 planes of hist_words
  uint8_t* tile = reinterpret_cast<uint8_t*>(sm + 2 * hist_words);
  for (int i = threadIdx.x; i < 2 * hist_words; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < kTile; i += blockDim.x) tile[i] = bins[(blockIdx.x * kTile + i) % (1 << 20)];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = 0x800000u + threadIdx.x * 977u;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int u = 0; u < 8; u++) {
      uint32_t a;
      if (MODE == 0) {
        const uint32_t b = tile[(it * 8 + u) * 1024 % kTile + threadIdx.x];       // consecutive rows of one feature
        a = b;                                                                      // S = 1
      } else {
        constexpr int F = MODE == 1 ? 32 : MODE == 2 ? 16 : MODE == 3 ? 8 : MODE == 4 ? 4 : 2;   // features per warp instruction
        constexpr int R = 32 / F;                                                                // rows per warp instruction
        const int row = ((it * 8 + u) * 32 + warp) * R + lane / F;
        const uint32_t b = tile[(row * F + (lane % F)) % kTile];                                  // [row][F] tile
        const uint32_t slot = (row * 7u) % R;                                                     // R slots: same footprint
        a = ((slot << 8) | b) * F + (lane % F);
      }
      atomicAdd(&hist[a], (((q >> 18) & 0x3Fu) << 13) | 1u);
      atomicAdd(&hist[hist_words + a], q);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = hist[threadIdx.x % hist_words];
}

// MODE 6: the lean k_hist2 loop — per warp instruction ONE broadcast LDS.128 of a pre-digested entry
// (tile byte offset, q, packed addend, slot base), one LDS.32 of the row's group word, PRMT, two address adds and two
// REDs into interleaved planes ([bin][2][32 features]): 8 instructions and 4 shared-memory operations per 32 elements.
__global__ void __launch_bounds__(1024) lean(const uint8_t* bins, uint32_t* out, int iters, long long* cycles) {
  extern __shared__ uint32_t sm[];
  uint32_t* hist = sm;                                   // 256 bins x 2 planes x 32 features
  uint32_t* tile = sm + 256 * 64;                        // 8 groups x (2048 + 4) words
  uint4* stage = reinterpret_cast<uint4*>(tile + 8 * 2052);   // 32 warps x 32 entries
  for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < 8 * 2052; i += blockDim.x) tile[i] = reinterpret_cast<const uint32_t*>(bins)[(blockIdx.x * 977 + i) % (1 << 18)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // entry of this lane: row offset (bytes) in the tile, q, addend, slot base (bytes)
  const uint32_t row = (threadIdx.x * 2 + (threadIdx.x * 7 % 2)) % 2048;
  const uint32_t q = 0x800000u + threadIdx.x * 977u;
  stage[warp * 32 + lane] = make_uint4(row * 4u, q, (((q >> 18) & 0x3Fu) << 13) | 1u, 0u);
  __syncthreads();
  const uint32_t s_hist = static_cast<uint32_t>(__cvta_generic_to_shared(hist)) + lane * 4u;
  const uint32_t s_tile = static_cast<uint32_t>(__cvta_generic_to_shared(tile)) + (lane >> 2) * 2052u * 4u;
  const uint32_t s_stage = static_cast<uint32_t>(__cvta_generic_to_shared(stage)) + warp * 512u;
  const uint32_t sel = 0x4440u | (lane & 3);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint4 e[4];
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e[u].x), "=r"(e[u].y), "=r"(e[u].z), "=r"(e[u].w) : "r"(s_stage + ((it * 8 + h * 4 + u) & 31) * 16u));
#pragma unroll
      for (int u = 0; u < 4; u++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[u]) : "r"(s_tile + e[u].x));
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t b = __byte_perm(w[u], 0u, sel);
        const uint32_t a = (b << 8) + (s_hist + e[u].w);
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(e[u].z) : "memory");
        asm volatile("red.shared.add.u32 [%0+128], %1;" ::"r"(a), "r"(e[u].y) : "memory");
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = hist[threadIdx.x];
}

void run_lean(const uint8_t* bins) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms, iters = 2000;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, grid * 1024 * 4); cudaMalloc(&cyc, grid * 8);
  const size_t smem = 256 * 64 * 4 + 8 * 2052 * 4 + 1024 * 16;
  cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lean<<<grid, 1024, smem>>>(bins, out, 10, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  lean<<<grid, 1024, smem>>>(bins, out, iters, cyc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
  const double elems = 1024.0 * iters * 8;
  printf("%-44s                     %.3f ms  elems/clk/SM=%.2f  Gelem/s=%.1f  err=%s\n", "lean loop: LDS.128 + LDS + PRMT + 2 RED", ms, elems / avg,
         elems * grid / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
void run(const char* name, int hist_words, const uint8_t* bins) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms, iters = 2000;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, grid * 1024 * 4); cudaMalloc(&cyc, grid * 8);
  const size_t smem = 2 * hist_words * 4 + kTile;
  cudaFuncSetAttribute(loop<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  loop<MODE><<<grid, 1024, smem>>>(bins, out, 10, hist_words, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  loop<MODE><<<grid, 1024, smem>>>(bins, out, iters, hist_words, cyc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
  const double elems = 1024.0 * iters * 8;
  printf("%-44s hist=%6d words  %.3f ms  elems/clk/SM=%.2f  Gelem/s=%.1f  err=%s\n", name, hist_words, ms, elems / avg,
         elems * grid / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  uint8_t* hbins = new uint8_t[1 << 20];
  uint32_t x = 12345;
  for (int i = 0; i < (1 << 20); i++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; hbins[i] = (x >> 11) & 0xFF; }
  uint8_t* bins; cudaMalloc(&bins, 1 << 20); cudaMemcpy(bins, hbins, 1 << 20, cudaMemcpyHostToDevice);
  run<0>("row per lane, random bins, S=1", 256, bins);
  run<1>("feature per lane, [bin][32]", 256 * 32, bins);
  run<2>("16 features x 2 rows, [slot][bin][16]", 256 * 32, bins);
  run<3>("8 features x 4 rows, [slot][bin][8]", 256 * 32, bins);
  run<4>("4 features x 8 rows, [slot][bin][4]", 256 * 32, bins);
  run<5>("2 features x 16 rows, [slot][bin][2]", 256 * 32, bins);
  run_lean(bins);
  return 0;
}
