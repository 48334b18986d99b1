// atoms_bench.cu — shared-memory atomic throughput on B200, to size the histogram kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_bench atoms_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters, int nbins, long long* cycles) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 2 * nbins; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  uint32_t acc = 0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x = x * 1664525u + 1013904223u;
      uint32_t a;
      if (MODE == 0) a = (lane + 32 * ((x >> 8) % (nbins / 32)));     // conflict-free banks
      else a = (x >> 8) % nbins;                                        // random bins
      const uint32_t q = x & 0xFFFFFF;
      if (MODE == 0 || MODE == 1) {            // one non-returning atomic
        atomicAdd(&sm[a], 1u);
      } else if (MODE == 2) {                  // one returning atomic
        acc += atomicAdd(&sm[a], q);
      } else if (MODE == 3) {                  // count + sum with carry (the histogram update)
        atomicAdd(&sm[a], 1u);
        const uint32_t old = atomicAdd(&sm[nbins + a], q);
        if (old + q < old) atomicAdd(&sm[a], 1u << 24);
      } else if (MODE == 4) {                  // two non-returning atomics
        atomicAdd(&sm[a], 1u);
        atomicAdd(&sm[nbins + a], q);
      } else if (MODE == 5) {                  // interleaved layout: (cnt, lo) adjacent words
        atomicAdd(&sm[2 * a], 1u);
        atomicAdd(&sm[2 * a + 1], q);
      } else if (MODE == 6) {                  // half the lanes inactive (sibling subtraction, no compaction)
        if (x & 0x80000000u) { atomicAdd(&sm[a], 1u); atomicAdd(&sm[nbins + a], q); }
      } else if (MODE == 7) {                  // plain LDS+STS read-modify-write (no atomic; upper bound of smem path)
        sm[a] += 1u;
      } else if (MODE == 8) {                  // kHistPacked: two non-returning atomics with general addends
        atomicAdd(&sm[a], (((q >> 18) & 0x3Fu) << 13) | 1u);   // count | coarse sum
        atomicAdd(&sm[nbins + a], q);                          // sum mod 2^32
      } else if (MODE == 9) {                  // north_star's "warp-reduced per-bin accumulation": match.any + redux,
        const uint32_t m = __match_any_sync(0xffffffffu, a);   // one pair of atomics per distinct bin of the warp
        const uint32_t s = __reduce_add_sync(m, q);
        if (lane == __ffs(m) - 1) {
          atomicAdd(&sm[a], static_cast<uint32_t>(__popc(m)));
          atomicAdd(&sm[nbins + a], s);
        }
      } else if (MODE >= 11 && MODE <= 15) {
        // TRULY random bins per lane (the LCG above gives the lanes of a warp an arithmetic progression, i.e. almost
        // conflict-free banks): murmur3 finaliser of (thread, iteration)
        uint32_t hsh = x ^ (threadIdx.x * 0x9E3779B9u);
        hsh ^= hsh >> 16; hsh *= 0x85EBCA6Bu; hsh ^= hsh >> 13; hsh *= 0xC2B2AE35u; hsh ^= hsh >> 16;
        const uint32_t bin = hsh & 255u;
        uint32_t a0;
        if (MODE == 11) a0 = hsh % nbins;                          // row-per-lane: bank = random
        else if (MODE == 12) a0 = (bin << 5) | lane;               // feature-per-lane, 32 features: bank = lane
        else if (MODE == 13) a0 = (((hsh >> 8) & 1u) << 12) | (bin << 4) | (lane & 15);   // 16 features x 2 rows, 2 slots
        else if (MODE == 14) a0 = (((hsh >> 8) & 3u) << 11) | (bin << 3) | (lane & 7);    // 8 features x 4 rows, 4 slots
        else a0 = (((hsh >> 8) & 15u) << 9) | (bin << 1) | (lane & 1);                    // 2 features x 16 rows, 16 slots
        atomicAdd(&sm[a0], (((q >> 18) & 0x3Fu) << 13) | 1u);
        atomicAdd(&sm[nbins + a0], q);
      } else if (MODE == 10) {                 // one 64-bit shared atomic (compiles to a CAS loop on sm_100a)
        atomicAdd(reinterpret_cast<unsigned long long*>(sm) + a, (static_cast<unsigned long long>(q) << 20) | 1ull);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x % nbins];
}

template <int MODE>
void run(const char* name, int nbins, int ctas_per_sm) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * ctas_per_sm, iters = 2000;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
  const size_t smem = 2 * nbins * 4;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<MODE><<<grid, 256, smem>>>(out, 10, nbins, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  bench<MODE><<<grid, 256, smem>>>(out, iters, nbins, cyc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
  const double elems_per_cta = 256.0 * iters * 8;
  printf("%-34s bins=%6d cta/sm=%d  %.3f ms  elems/clk/SM=%.2f  (cycles/CTA %.0f)  Gelem/s=%.1f  err=%s\n", name, nbins, ctas_per_sm, ms,
         elems_per_cta * ctas_per_sm / avg, avg, elems_per_cta * grid / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  for (int cps : {1, 2, 4}) {
    run<0>("1 atomic, conflict-free", 256, cps);
    run<1>("1 atomic, random", 256, cps);
    run<1>("1 atomic, random", 8192, cps);
    run<2>("1 returning atomic, random", 8192, cps);
    run<3>("cnt+sum+carry, random", 256, cps);
    run<3>("cnt+sum+carry, random", 8192, cps);
    run<4>("2 non-returning, planar", 8192, cps);
    run<5>("2 non-returning, interleaved", 8192, cps);
    run<6>("2 atomics, half lanes active", 8192, cps);
    run<7>("LDS+STS rmw (non-atomic)", 8192, cps);
    run<8>("2 non-returning, general addends", 256, cps);
    run<8>("2 non-returning, general addends", 8192, cps);
    run<9>("match_any + redux, then 2 atomics", 256, cps);
    run<9>("match_any + redux, then 2 atomics", 8192, cps);
    run<10>("1 x 64-bit atomic (CAS loop)", 8192, cps);
    run<11>("2 RED, hashed bins, row per lane", 8192, cps);
    run<12>("2 RED, hashed bins, [bin][32 feat]", 8192, cps);
    run<13>("2 RED, hashed bins, [bin][16 feat]", 8192, cps);
    run<14>("2 RED, hashed bins, [bin][8 feat]", 8192, cps);
    run<15>("2 RED, hashed bins, [bin][2 feat]", 8192, cps);
  }
  return 0;
}
