// wy_trace.cu — cycle trace of one CTA of the blocked-WY kernel under a full-GPU load (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DQRK_WY_TRACE -I qrkit_b200/csrc -I include -o tools/wy_trace tools/wy_trace.cu
// usage: tools/wy_trace R C [nblocks]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bd_wy.cuh"
using namespace qrk;

template <int MR, int W>
void run(int r, int c, int nb) {
  const size_t smem = (wy_smem_bytes(r, c) + 2047) / 2048 * 2048;
  double *A, *P, *tau, *b, *x;
  cudaMalloc(&A, sizeof(double) * (size_t)nb * r * c); cudaMalloc(&P, sizeof(double) * (size_t)nb * r * c);
  cudaMalloc(&tau, sizeof(double) * (size_t)nb * c); cudaMalloc(&b, sizeof(double) * (size_t)nb * r); cudaMalloc(&x, sizeof(double) * (size_t)nb * c);
  std::vector<double> h((size_t)nb * r * c);
  for (size_t i = 0; i < h.size(); i++) h[i] = synth_value(1, i / (r * c), i % r, (i / r) % c, 0.5, 5.0);
  cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(b, 0, sizeof(double) * (size_t)nb * r);
  BlockIndex bi{}; bi.ur = r; bi.uc = c;
  auto k = bd_wy_factor_kernel<MR, W, true>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32 * W, smem);
  for (int it = 0; it < 2; it++) {
    int zero = 0; cudaMemcpyToSymbol(g_wy_trace_n, &zero, sizeof(int));
    k<<<nb, 32 * W, smem>>>(bi, nullptr, A, P, tau, b, x);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  }
  int n; cudaMemcpyFromSymbol(&n, g_wy_trace_n, sizeof(int));
  std::vector<long long> t(4096); cudaMemcpyFromSymbol(t.data(), g_wy_trace, sizeof(long long) * 4096);
  printf("block %dx%d MR=%d W=%d smem=%zu occupancy=%d CTAs/SM, %d trace records\n", r, c, MR, W, smem, occ, n);
  const char* names[] = {"panel:start", "panel:loaded", "panel:steps done", "tiles:start", "tile:done", "panel:T done", "apply:start", "apply:end", "phase:arrive", "phase:leave", "epilogue", "staged", "  steps: acc0 (row panel: partial dots | team panel: barrier A)", "  steps: acc1 (reduce-scatter | loads + dot + shuffles)", "  steps: acc2 (smem gather | scalar chain)", "  steps: acc3 (scalar chain | barrier B)", "  steps: acc4 (update+T | update + publication)"};
  long long t0 = t[1];
  for (int i = 0; i < n && i < 2048; i++) if (t[2 * i] / 16 < 12) t0 = t[2 * i + 1] < t0 ? t[2 * i + 1] : t0;
  for (int i = 0; i < n && i < 2048; i++) {
    const int tag = (int)(t[2 * i] / 16), w = (int)(t[2 * i] % 16);
    printf("%8lld  w%d  %s\n", tag >= 12 ? t[2 * i + 1] : t[2 * i + 1] - t0, w, names[tag]);
  }
}

int main(int argc, char** argv) {
  const int r = argc > 1 ? atoi(argv[1]) : 128, c = argc > 2 ? atoi(argv[2]) : 64, nb = argc > 3 ? atoi(argv[3]) : 148 * 6;
  const int mr = wy_mr(r, c), w = wy_warps(r, c);
  if (mr == 4 && w == 4) run<4, 4>(r, c, nb);
  else if (mr == 2 && w == 4) run<2, 4>(r, c, nb);
  else if (mr == 1 && w == 2) run<1, 2>(r, c, nb);
  else if (mr == 1 && w == 4) run<1, 4>(r, c, nb);
  else if (mr == 1 && w == 1) run<1, 1>(r, c, nb);
  else if (mr == 2 && w == 1) run<2, 1>(r, c, nb);
  else if (mr == 2 && w == 2) run<2, 2>(r, c, nb);
  else if (mr == 4 && w == 2) run<4, 2>(r, c, nb);
  else if (mr == 4 && w == 1) run<4, 1>(r, c, nb);
  else if (mr == 3 && w == 2) run<3, 2>(r, c, nb);
  else if (mr == 3 && w == 4) run<3, 4>(r, c, nb);
  else printf("unsupported combination\n");
  return 0;
}
