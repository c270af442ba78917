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
#include <string.h>

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
    if (share == 0) { idx[i] = (int)(splitmix64((uint64_t)i * 0x9E37ull + 777u) % (uint64_t)table); continue; }   // any double of the table
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


// texture path: tex1Dfetch<int2> over the same linear table (TEX pipe of L1TEX instead of the LSU pipe)
__global__ void __launch_bounds__(THREADS) k_tex(const int* __restrict__ idx, cudaTextureObject_t tex, double* out) {
  const int64_t base = (int64_t)blockIdx.x * TILE;
  int c[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * THREADS + threadIdx.x);
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int2 v = tex1Dfetch<int2>(tex, c[it]);
    acc += __hiloint2double(v.y, v.x);
  }
  if (acc == -1.0) out[0] = acc;
}

// half the gathers through LDG, half through TEX (do the two pipes add up?)
__global__ void __launch_bounds__(THREADS) k_mix(const int* __restrict__ idx, const double* __restrict__ x, cudaTextureObject_t tex, double* out) {
  const int64_t base = (int64_t)blockIdx.x * TILE;
  int c[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * THREADS + threadIdx.x);
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    if (it & 1) { const int2 v = tex1Dfetch<int2>(tex, c[it]); acc += __hiloint2double(v.y, v.x); }
    else acc += ldg_na(x + c[it]);
  }
  if (acc == -1.0) out[0] = acc;
}

// shared-memory window: every CTA stages SLICE doubles of the table (coalesced 16-byte loads) and serves `tiles` tiles of
// random 8-byte gathers from it — the banded-family design question: what does a gather cost once x sits in shared memory?
constexpr int SM_THREADS = 1024;
__global__ void __launch_bounds__(SM_THREADS) k_smem(const int* __restrict__ idx, const double* __restrict__ x, double* out,
                                                     int slice, int tiles_per_cta) {
  extern __shared__ __align__(16) double win[];
  const double2* src = reinterpret_cast<const double2*>(x + (size_t)blockIdx.x * slice);
  for (int i = threadIdx.x; i < slice / 2; i += SM_THREADS) reinterpret_cast<double2*>(win)[i] = src[i];
  __syncthreads();
  double acc = 0.0;
  const int mask = slice - 1;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int64_t base = ((int64_t)blockIdx.x * tiles_per_cta + t) * (SM_THREADS * PER);
    int c[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) c[it] = ldg_stream(idx + base + it * SM_THREADS + threadIdx.x);
#pragma unroll
    for (int it = 0; it < PER; ++it) acc += win[c[it] & mask];
  }
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

  {
    make_idx<<<sms * 8, 256>>>(idx, gathers, table, 1);
    CHECK(cudaDeviceSynchronize());
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x; rd.res.linear.desc = cudaCreateChannelDesc<int2>();
    rd.res.linear.sizeInBytes = sizeof(double) * table;
    cudaTextureDesc td; memset(&td, 0, sizeof(td)); td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess) {
      const unsigned grid = (unsigned)(gathers / TILE);
      cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
      for (int v = 0; v < 2; ++v) {
        for (int r = 0; r < 2; ++r) { if (v == 0) k_tex<<<grid, THREADS>>>(idx, tex, out); else k_mix<<<grid, THREADS>>>(idx, x, tex, out); }
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) { if (v == 0) k_tex<<<grid, THREADS>>>(idx, tex, out); else k_mix<<<grid, THREADS>>>(idx, x, tex, out); }
        CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        float ms = 0; CHECK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        const double gps = gathers / (ms * 1e-3);
        printf(", \"%s\": {\"ms\": %.4f, \"Ggathers_per_s\": %.1f, \"per_sm_cycle\": %.3f}", v == 0 ? "tex_int2" : "ldg_tex_mix", ms, gps / 1e9, gps / (sms * (khz * 1e3)));
      }
    } else { cudaGetLastError(); printf(", \"tex_int2\": null"); }
    // shared-memory window, 8 K / 16 K doubles (64 / 128 KB) per CTA; indices uniform over the window (random banks)
    make_idx<<<sms * 8, 256>>>(idx, gathers, table, 0);
    CHECK(cudaDeviceSynchronize());
    for (int slice_k = 8; slice_k <= 16; slice_k *= 2) {
      const int slice = slice_k << 10;
      const size_t smem = sizeof(double) * slice;
      CHECK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int tiles = 1; tiles <= 16; tiles *= 4) {
        const int64_t per_cta = (int64_t)tiles * SM_THREADS * PER;
        int64_t ctas = gathers / per_cta;
        if ((int64_t)ctas * slice > table) ctas = table / slice;
        cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
        k_smem<<<(unsigned)ctas, SM_THREADS, smem>>>(idx, x, out, slice, tiles);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) k_smem<<<(unsigned)ctas, SM_THREADS, smem>>>(idx, x, out, slice, tiles);
        CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        float ms = 0; CHECK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        const double g = (double)ctas * per_cta, gps = g / (ms * 1e-3);
        printf(", \"smem%dk_t%d\": {\"ms\": %.4f, \"Ggathers_per_s\": %.1f, \"per_sm_cycle\": %.3f, \"fill_bytes_per_gather\": %.2f}", slice_k, tiles, ms, gps / 1e9,
               gps / (sms * (khz * 1e3)), 8.0 * slice / per_cta);
      }
    }
  }
  printf("}\n");
  return 0;
}
