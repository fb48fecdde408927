// fp64_pipes.cu — micro-benchmark of the FP64 pipes the blocked-WY kernel depends on (B200, sm_100a):
// DFMA, DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64), 64-bit SHFL, and the dependent-issue
// latency of each.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_pipes tools/fp64_pipes.cu
// Output: per-SM throughput in FMA/clk (DFMA, DMMA) or warp-instructions/clk (SHFL) and latencies in cycles.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP, int SHAPE>
__global__ void k_dmma(double* out, int iters, long long* cyc) {
  double d2[ILP][2], d4[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) { d2[i][0] = d2[i][1] = 0; d4[i][0] = d4[i][1] = d4[i][2] = d4[i][3] = 0; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1e-3 * (threadIdx.x * 3 + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (SHAPE == 0) dmma884(d2[i], a[0], b[0]);
      if (SHAPE == 1) dmma1684(d4[i], a[0], a[1], b[0]);
      if (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; dmma1688(d4[i], aa, b[0], b[1]); }
      if (SHAPE == 3) dmma16816(d4[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += d2[i][0] + d2[i][1] + d4[i][0] + d4[i][1] + d4[i][2] + d4[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_shfl(double* out, int iters, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (it & 15));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// shared-memory broadcast/LDS.128 throughput: every lane reads the same 16 bytes (broadcast) or its own
template <int MODE>
__global__ void k_lds(double* out, int iters, long long* cyc) {
  __shared__ double2 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2(i, i + 0.5);
  __syncthreads();
  double s0 = 0, s1 = 0;
  int idx = MODE == 0 ? 0 : (threadIdx.x & 31);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double2 v = buf[(idx + i * 32 + it) & 1023];
      s0 += v.x; s1 += v.y;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(double) * 148 * 1024 * 8));
  CK(cudaMallocManaged(&cyc, sizeof(long long)));
  const int iters = 4096;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int nsm = p.multiProcessorCount;
#define RUN(label, kern, threads, ctas_per_sm, work_per_thread_iter, unit)                                   \
  do {                                                                                                       \
    kern<<<nsm * ctas_per_sm, threads>>>(out, iters, cyc);                                                   \
    CK(cudaDeviceSynchronize());                                                                             \
    kern<<<nsm * ctas_per_sm, threads>>>(out, iters, cyc);                                                   \
    CK(cudaDeviceSynchronize());                                                                             \
    double per_sm = (double)(work_per_thread_iter) * iters * threads * ctas_per_sm / (double)*cyc;           \
    printf("%-44s %8.2f %s per clk per SM   (%lld cycles, %.1f cyc/iter)\n", label, per_sm, unit, *cyc, (double)*cyc / iters); \
  } while (0)

  // latency: 1 warp, ILP 1
  RUN("DFMA latency (1 warp, ILP1): FMA", (k_dfma<1>), 32, 1, 1, "FMA");
  RUN("DFMA 1 warp ILP8", (k_dfma<8>), 32, 1, 8, "FMA");
  RUN("DFMA 4 warps ILP8", (k_dfma<8>), 128, 1, 8, "FMA");
  RUN("DFMA 16 warps ILP8", (k_dfma<8>), 512, 1, 8, "FMA");
  RUN("DFMA 32 warps ILP8", (k_dfma<8>), 1024, 1, 8, "FMA");
  RUN("DMMA m8n8k4 latency (1 warp, ILP1)", (k_dmma<1, 0>), 32, 1, 8, "FMA");
  RUN("DMMA m8n8k4 1 warp ILP8", (k_dmma<8, 0>), 32, 1, 64, "FMA");
  RUN("DMMA m8n8k4 4 warps ILP8", (k_dmma<8, 0>), 128, 1, 64, "FMA");
  RUN("DMMA m8n8k4 16 warps ILP8", (k_dmma<8, 0>), 512, 1, 64, "FMA");
  RUN("DMMA m8n8k4 32 warps ILP4", (k_dmma<4, 0>), 1024, 1, 32, "FMA");
  RUN("DMMA m16n8k4 latency (1 warp, ILP1)", (k_dmma<1, 1>), 32, 1, 16, "FMA");
  RUN("DMMA m16n8k4 16 warps ILP4", (k_dmma<4, 1>), 512, 1, 64, "FMA");
  RUN("DMMA m16n8k8 latency (1 warp, ILP1)", (k_dmma<1, 2>), 32, 1, 32, "FMA");
  RUN("DMMA m16n8k8 16 warps ILP4", (k_dmma<4, 2>), 512, 1, 128, "FMA");
  RUN("DMMA m16n8k16 latency (1 warp, ILP1)", (k_dmma<1, 3>), 32, 1, 64, "FMA");
  RUN("DMMA m16n8k16 16 warps ILP4", (k_dmma<4, 3>), 512, 1, 256, "FMA");
  RUN("SHFL.64 latency (1 warp, ILP1)", (k_shfl<1>), 32, 1, 1.0 / 32, "warp-shfl64");
  RUN("SHFL.64 16 warps ILP8", (k_shfl<8>), 512, 1, 8.0 / 32, "warp-shfl64");
  RUN("LDS.128 broadcast 16 warps", (k_lds<0>), 512, 1, 8.0 / 32, "warp-LDS128");
  RUN("LDS.128 per-lane 16 warps", (k_lds<1>), 512, 1, 8.0 / 32, "warp-LDS128");
  return 0;
}
