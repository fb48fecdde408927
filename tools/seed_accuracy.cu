// Measures the relative error of the MUFU.RCP64H / MUFU.RSQ64H seeds (rcp.approx.ftz.f64, rsqrt.approx.ftz.f64)
// and of the Newton-refined fast_rcp / fast_rsqrt of common.cuh against the IEEE results.
#include <cstdio>
#include <cmath>
#include "../qrkit_b200/csrc/common.cuh"
__global__ void k(double* out, int n) {
  double m_rcp = 0, m_rsq = 0, m_frcp = 0, m_frsq = 0, m_fsqrt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double x = qrk::synth_value(12345, i, 0, 0, 0.0, 1.0);
    x = ldexp(1.0 + x, (i % 200) - 100);
    double r0, y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double r = 1.0 / x, y = 1.0 / sqrt(x);
    m_rcp = fmax(m_rcp, fabs(r0 - r) / r);
    m_rsq = fmax(m_rsq, fabs(y0 - y) / y);
    double s;
    const double fy = qrk::fast_rsqrt(x, s);
    m_frcp = fmax(m_frcp, fabs(qrk::fast_rcp(x) - r) / r);
    m_frsq = fmax(m_frsq, fabs(fy - y) / y);
    m_fsqrt = fmax(m_fsqrt, fabs(s - sqrt(x)) / sqrt(x));
  }
  double* o = out + 5 * (blockIdx.x * blockDim.x + threadIdx.x);
  o[0] = m_rcp; o[1] = m_rsq; o[2] = m_frcp; o[3] = m_frsq; o[4] = m_fsqrt;
}
// householder_scalars against IEEE sqrt / division on the same inputs: max relative error of beta, inv, tau
__global__ void k2(double* out, int n) {
  double m_beta = 0, m_inv = 0, m_tau = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double x0 = qrk::synth_value(777, i, 0, 0, -1.0, 1.0);
    double tl = qrk::synth_value(778, i, 0, 0, 0.0, 1.0);
    x0 = ldexp(x0, (i % 120) - 60);
    tl = ldexp(tl, ((i / 7) % 240) - 120);            // tail norms from far below to far above |x0|
    double beta, inv, tau;
    qrk::householder_scalars(x0, tl, false, beta, inv, tau);
    const double nrm = sqrt(fma(x0, x0, tl));
    const double eb = (x0 >= 0.0) ? -nrm : nrm, ei = 1.0 / (x0 - eb), et = (eb - x0) / eb;
    if (tl > DBL_MIN) {
      m_beta = fmax(m_beta, fabs(beta - eb) / fabs(eb));
      m_inv = fmax(m_inv, fabs(inv - ei) / fabs(ei));
      m_tau = fmax(m_tau, fabs(tau - et) / fabs(et));
    }
  }
  double* o = out + 5 * (blockIdx.x * blockDim.x + threadIdx.x);
  o[0] = m_beta; o[1] = m_inv; o[2] = m_tau; o[3] = 0; o[4] = 0;
}
int main() {
  const int T = 148 * 256;
  double* d; cudaMalloc(&d, T * 5 * sizeof(double));
  k<<<148, 256>>>(d, 1 << 24);
  static double h[T * 5];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double m[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < T; i++) for (int j = 0; j < 5; j++) m[j] = fmax(m[j], h[5 * i + j]);
  printf("seed rcp rel err %.3e (%.1f bits)  seed rsqrt %.3e (%.1f bits)\n", m[0], -log2(m[0]), m[1], -log2(m[1]));
  printf("fast_rcp %.3e  fast_rsqrt %.3e  fast sqrt %.3e  (eps = %.3e)\n", m[2], m[3], m[4], 2.22e-16);
  k2<<<148, 256>>>(d, 1 << 24);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int j = 0; j < 5; j++) m[j] = 0;
  for (int i = 0; i < T; i++) for (int j = 0; j < 5; j++) m[j] = fmax(m[j], h[5 * i + j]);
  printf("householder_scalars vs IEEE: beta %.3e  inv %.3e  tau %.3e\n", m[0], m[1], m[2]);
  return 0;
}
