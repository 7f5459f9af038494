// sector_peak.cu -- random 32-byte-sector read throughput of one B200 ("measure an L2
// random-sector peak microbenchmark"; BASELINE north_star: "the probe-bound roofline").
//
// A hash probe of the imputation path reads ONE 32-byte sector at a pseudo-random address of the
// label's hash region (csrc/grimb_tables.h: load_sector, one LDG.E.256).  A copy kernel moves whole
// 128-byte lines from open DRAM pages; a probe stream cannot, so the measured copy bandwidth
// (MEASURED_PEAKS.json hbm_gbs) is an upper bound the probe kernel can never reach.  This program
// measures what the memory system sustains for exactly the probe's access shape -- independent
// random 32-byte loads, nothing else -- for table sizes from L2-resident to far beyond L2, so
// k_fast_probe's sector rate can be placed against it.  Measurement tool only: not linked into the
// product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/sector_peak.bin tools/sector_peak.cu
//   tools/sector_peak.bin            -> one JSON line per (table size, loads in flight per thread)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // murmur3 finaliser (the table hash's mixer)
  x ^= x >> 16;
  x *= 0x85ebca6bu;
  x ^= x >> 13;
  x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

// one whole sector with one 256-bit load, the way load_sector does it
__device__ __forceinline__ uint32_t load32B(const uint8_t* p) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "l"(p));
  return r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
}

// Every thread issues `iters` rounds of ILP independent sector loads.  Addresses depend only on the
// thread id and the round (never on loaded data), like the 2^L probes of a subject.
template <int ILP>
__global__ void __launch_bounds__(256) k_sectors(const uint8_t* __restrict__ tab, uint32_t sector_mask, int iters,
                                                 uint32_t seed, unsigned int* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  uint32_t ctr = seed + tid * 0x9E3779B9u;
  for (int it = 0; it < iters; ++it) {
    uint32_t v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      ctr += 0x632BE5ABu;
      const uint64_t sec = mix32(ctr) & sector_mask;
      v[j] = load32B(tab + (sec << 5));
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= v[j];
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1u);   // keeps the loads alive; the table is zero-filled
}

// sequential whole-line reads with the same load instruction: the streaming ceiling beside it
__global__ void __launch_bounds__(256) k_stream(const uint8_t* __restrict__ tab, uint64_t n_sectors, unsigned int* sink) {
  uint32_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sectors; i += stride) acc ^= load32B(tab + (i << 5));
  if (acc == 0x12345678u) atomicAdd(sink, 1u);
}

template <int ILP>
static float run(const uint8_t* tab, uint32_t mask, int grid, int iters, unsigned int* sink, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int w = 0; w < 3; ++w) k_sectors<ILP><<<grid, 256>>>(tab, mask, iters, 17u + w, sink);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    k_sectors<ILP><<<grid, 256>>>(tab, mask, iters, 1000u + r, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return best;
}

int main(int argc, char** argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, dev));
  const int sms = pr.multiProcessorCount;
  unsigned int* sink;
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(sink, 0, 4));
  const uint64_t sizes[] = {1ull << 20, 1ull << 24, 1ull << 26, 1ull << 27, 1ull << 28, 1ull << 30, 1ull << 32, 1ull << 34};
  const uint64_t max_bytes = sizes[sizeof(sizes) / sizeof(sizes[0]) - 1];
  uint8_t* tab;
  CK(cudaMalloc(&tab, max_bytes));
  CK(cudaMemset(tab, 0, max_bytes));
  CK(cudaDeviceSynchronize());
  const int grid = sms * 8;                 // 8 CTAs of 256 threads per SM: all 64 warps resident
  const uint64_t threads = (uint64_t)grid * 256;
  for (uint64_t bytes : sizes) {
    const uint32_t mask = (uint32_t)((bytes >> 5) - 1);
    for (int ilp : {1, 2, 4}) {
      // about 2^27 sector loads (4 GiB of sectors) per launch
      const int iters = (int)(((1ull << 27) / threads + ilp - 1) / ilp);
      float ms = ilp == 1 ? run<1>(tab, mask, grid, iters, sink, 10)
                 : ilp == 2 ? run<2>(tab, mask, grid, iters, sink, 10)
                            : run<4>(tab, mask, grid, iters, sink, 10);
      const double loads = (double)threads * iters * ilp;
      printf("{\"kind\": \"random_sector\", \"gpu\": \"%s\", \"sms\": %d, \"table_bytes\": %llu, \"loads_in_flight_per_thread\": %d, "
             "\"threads\": %llu, \"sector_loads\": %.0f, \"ms\": %.4f, \"gsectors_per_s\": %.3f, \"gbs\": %.1f}\n",
             pr.name, sms, (unsigned long long)bytes, ilp, (unsigned long long)threads, loads, ms, loads / ms * 1e-6,
             loads * 32.0 / ms * 1e-6);
      fflush(stdout);
    }
  }
  {   // streaming read of 4 GiB with the same instruction
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const uint64_t nsec = (1ull << 32) >> 5;
    float best = 1e30f;
    for (int r = 0; r < 8; ++r) {
      CK(cudaEventRecord(a));
      k_stream<<<grid, 256>>>(tab, nsec, sink);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (r >= 2 && ms < best) best = ms;
    }
    printf("{\"kind\": \"stream_read\", \"gpu\": \"%s\", \"bytes\": %llu, \"ms\": %.4f, \"gbs\": %.1f}\n", pr.name,
           (unsigned long long)(nsec << 5), best, (double)(nsec << 5) / best * 1e-6);
  }
  CK(cudaFree(tab));
  CK(cudaFree(sink));
  return 0;
}
