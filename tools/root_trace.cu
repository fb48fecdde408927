// root_trace.cu — cycle trace of the block-angular TSQR root kernel (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DQRK_ROOT_TRACE -I qrkit_b200/csrc -I include -o tools/root_trace tools/root_trace.cu
// usage: tools/root_trace [count]      (count partial triangles, default 444 = one per resident CTA of K1)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "angular.cuh"
using namespace qrk;

template <int TPB>
int run(int count, int ld) {
  constexpr int M2 = 5, N = Tri<M2>::N;
  std::vector<double> h((size_t)count * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = synth_value(7, i / N, i % N, 0, 0.5, 5.0);
  double *tris, *out, *root; int *root_i, *perm_tail;
  cudaMalloc(&tris, h.size() * 8); cudaMalloc(&out, N * 8); cudaMalloc(&root, (M2 * M2 + 3 * M2) * 8);
  cudaMalloc(&root_i, (M2 + 1) * 4); cudaMalloc(&perm_tail, M2 * 4);
  cudaMemcpy(tris, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  AngularXchg xc;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(e0);
    angular_root_kernel<M2, TPB, false><<<1, TPB>>>(tris, count, ld, 1, out, root, root_i, 0, perm_tail, 0, xc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long t[16]; cudaMemcpyFromSymbol(t, g_root_trace, sizeof(t));
    const char* names[] = {"start", "triangles loaded (+ folds of later rounds)", "CTA merge done", "ColPiv start", "ColPiv done", "stores issued"};
    if (it < 2) continue;
    printf("%d threads, %d triangles (%s), event time %.2f us\n", TPB, count, ld ? "component-major" : "triangle-major", ms * 1e3);
    for (int i = 0; i < 6; i++) printf("  %7lld cycles  %s\n", t[i] - t[0], names[i]);
  }
  return 0;
}

int main(int argc, char** argv) {
  const int count = argc > 1 ? atoi(argv[1]) : 444;
  run<512>(count, count);                       // (random data: only the timing means something)
  run<256>(count, count);
  run<128>(count, count);                       // the shipped configuration for M2 <= 5 (4 triangles per lane)
  run<128>(count, 0);
  run<32>(count < 128 ? count : 128, count < 128 ? count : 128);
  return 0;
}
