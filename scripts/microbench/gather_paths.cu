// gather_paths.cu — standalone micro-benchmark (NOT part of libsla_b200.so): how many random 8-byte gathers per second can
// an SM issue from an L2-resident table, through each of the paths sm_100a offers?  DESIGN.md §3.1 derives the cfg-2 (#>)
// ceiling from ONE L1TEX -> crossbar request per SM-cycle (a request = one (warp instruction, 128-byte line) pair); this
// program measures that port directly and probes the ways around it:
//   ldg        ld.global.nc.L1::no_allocate f64, one random line per lane              (the path spmv_tile_kernel uses)
//   ldg_pair   the same, but lanes 2k / 2k+1 read the same 128-byte line              (requests are per LINE: expect ~2x)
//   ldg_quad   four lanes per line                                                    (expect ~4x until another limit bites)
//   ldgsts     cp.async.ca.shared.global 8 B per lane                                 (same L1TEX port expected)
//   bulk16     cp.async.bulk.shared::cluster.global 16 B per lane + mbarrier          (TMA / UBLKCP path: own port?)
// Output: one JSON object with gathers/s and gathers per SM-cycle for every variant.
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_paths gather_paths.cu
// run:    ./gather_paths [table_MB=32] [gathers_M=256]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int THREADS = 128;
constexpr int PER = 8;                      // gathers per thread, as in spmv_tile_kernel
constexpr int TILE = THREADS * PER;

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// share = lanes per 128-byte line (1, 2 or 4): lanes of a group get the same line, different doubles inside it
__global__ void make_idx(int* idx, int64_t n, int64_t table, int share) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t group = i / share;
    const int64_t line = (int64_t)(splitmix64((uint64_t)group * 0x9E37ull + 12345u) % (uint64_t)(table / 16));
    idx[i] = (int)(line * 16 + (i % share) * (16 / share));
  }
}
__global__ void fill_table(double* x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = (double)(i & 1023) * 0.5;
}

__device__ __forceinline__ double ldg_na(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_stream(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

__global__ void __launch_bounds__(THREADS) k_ldg(const int* __restrict__ idx, const double* __restrict__ x, double* out) {
  const int64_t base = (int64_t)blockIdx.x * TILE;
  int c[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * THREADS + threadIdx.x);
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < PER; ++it) acc += ldg_na(x + c[it]);
  if (acc == -1.0) out[0] = acc;            // keep the loads alive
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(THREADS) k_ldgsts(const int* __restrict__ idx, const double* __restrict__ x, double* out) {
  __shared__ double buf[TILE];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  int c[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * THREADS + threadIdx.x);
#pragma unroll
  for (int it = 0; it < PER; ++it)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(buf + it * THREADS + threadIdx.x)), "l"(x + c[it]) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < PER; ++it) acc += buf[it * THREADS + threadIdx.x];
  if (acc == -1.0) out[0] = acc;
}

// every lane issues its own 16-byte bulk copy (the aligned pair holding its double); one mbarrier per CTA counts the bytes
__global__ void __launch_bounds__(THREADS) k_bulk16(const int* __restrict__ idx, const double* __restrict__ x, double* out) {
  __shared__ alignas(16) double buf[TILE * 2];
  __shared__ alignas(8) uint64_t bar;
  const int64_t base = (int64_t)blockIdx.x * TILE;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(TILE * 16) : "memory");
  __syncthreads();
  int c[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * THREADS + threadIdx.x);
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const double* src = x + (c[it] & ~1);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf + 2 * (it * THREADS + threadIdx.x))), "l"(src), "r"(16), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < PER; ++it) acc += buf[2 * (it * THREADS + threadIdx.x) + (c[it] & 1)];
  if (acc == -1.0) out[0] = acc;
}

template <class K>
static double time_kernel(K kernel, const int* idx, const double* x, double* out, int64_t gathers, int reps) {
  const unsigned grid = (unsigned)(gathers / TILE);
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  for (int r = 0; r < 2; ++r) kernel<<<grid, THREADS>>>(idx, x, out);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) kernel<<<grid, THREADS>>>(idx, x, out);
  CHECK(cudaEventRecord(e1));
  CHECK(cudaEventSynchronize(e1));
  float ms = 0;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  return (double)ms / reps;
}

int main(int argc, char** argv) {
  const int64_t table_mb = argc > 1 ? atoll(argv[1]) : 32;
  const int64_t gathers = ((argc > 2 ? atoll(argv[2]) : 256) << 20) / TILE * TILE;
  const int64_t table = table_mb << 17;     // doubles
  int dev = 0, sms = 0, khz = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  double *x = nullptr, *out = nullptr;
  int* idx = nullptr;
  CHECK(cudaMalloc(&x, sizeof(double) * table));
  CHECK(cudaMalloc(&out, sizeof(double)));
  CHECK(cudaMalloc(&idx, sizeof(int) * gathers));
  fill_table<<<sms * 8, 256>>>(x, table);
  printf("{\"table_MB\": %lld, \"gathers\": %lld, \"sms\": %d, \"sm_clock_GHz\": %.3f", (long long)table_mb, (long long)gathers, sms, khz / 1e6);
  const char* names[5] = {"ldg", "ldg_pair", "ldg_quad", "ldgsts", "bulk16"};
  for (int v = 0; v < 5; ++v) {
    const int share = v == 1 ? 2 : v == 2 ? 4 : 1;
    make_idx<<<sms * 8, 256>>>(idx, gathers, table, share);
    CHECK(cudaDeviceSynchronize());
    double ms;
    if (v <= 2) ms = time_kernel(k_ldg, idx, x, out, gathers, 5);
    else if (v == 3) ms = time_kernel(k_ldgsts, idx, x, out, gathers, 5);
    else ms = time_kernel(k_bulk16, idx, x, out, gathers, 5);
    const double gps = gathers / (ms * 1e-3);
    printf(", \"%s\": {\"ms\": %.4f, \"Ggathers_per_s\": %.1f, \"per_sm_cycle\": %.3f}", names[v], ms, gps / 1e9, gps / (sms * (khz * 1e3)));
  }
  printf("}\n");
  return 0;
}
