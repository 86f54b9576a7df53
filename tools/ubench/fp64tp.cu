// FP64 pipe throughput per SM: independent chains, all warps busy.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* cyc, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-6, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
  double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
  const double m = 1.0000001, c = 1e-9;
  int cnt = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
    if (OP == 1) { a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c; }
    if (OP == 2) { cnt += (a0 > a1) + (a1 > a2) + (a2 > a3) + (a3 > a4) + (a4 > a5) + (a5 > a6) + (a6 > a7) + (a7 > a0); a0 += c; }
    if (OP == 3) { a0 = 1.0 / a0 + 1.0; a1 = 1.0 / a1 + 1.0; a2 = 1.0 / a2 + 1.0; a3 = 1.0 / a3 + 1.0; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + cnt;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* nm, int ops_per_iter, double* o, long long* t) {
  const int iters = 4096, threads = 512;
  k<OP><<<148, threads>>>(o, t, iters);
  k<OP><<<148, threads>>>(o, t, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
  double per_clk = (double)iters * ops_per_iter * threads / (double)h;
  printf("%-10s %8.2f lane-ops/clk/SM  (%lld cycles)\n", nm, per_clk, h);
}
int main() {
  double* o; long long* t;
  cudaMalloc(&o, 148 * 512 * 8); cudaMalloc(&t, 8);
  run<0>("DFMA", 8, o, t);
  run<1>("DADD", 8, o, t);
  run<2>("DSETP", 8, o, t);
  run<3>("1/x+1", 4, o, t);
  return 0;
}
