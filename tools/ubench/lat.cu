// Latency micro-benchmarks for the matcher's critical path (single warp, dependent chains).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fit(double jc, double jm, double ac, double am, double lc, double lm, double rc, double rm) {
  if (ac + jc > lc) return 0.0;
  if (am + jm > lm) return 0.0;
  double cpu_fit = ((jc + ac) + rc) / (lc + rc);
  double mem_fit = ((jm + am) + rm) / (lm + rm);
  return (cpu_fit + mem_fit) / 2.0;
}
__global__ void k(double* out, long long* t, volatile int* flag) {
  __shared__ double sm[64];
  __shared__ volatile int sflag;
  const int lane = threadIdx.x;
  sm[lane] = 1.0 + lane; sm[32 + lane] = 3.0 + lane;
  sflag = 0;
  __syncthreads();
  const int N = 256;
  double x = 1.0 + lane * 1e-3, y = 3.0;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) x = y / (x + 1.0);  // dependent div + add
  long long t1 = clock64();
  double z = x;
  for (int i = 0; i < N; i++) z = z * 0.999 + 1.0;  // dependent dmul + dadd (no fma: -fmad=false)
  long long t2 = clock64();
  double f = z;
  for (int i = 0; i < N; i++) f = fit(1.0, 512.0 + f, 2.0, 1024.0, 64.0, 262144.0, 8.0, 4096.0);
  long long t3 = clock64();
  unsigned u = (unsigned)lane + (unsigned)(f * 0.0);
  for (int i = 0; i < N; i++) u = __reduce_max_sync(0xffffffffu, u + lane) & 0xffff;
  long long t4 = clock64();
  int idx = lane;
  double acc = 0;
  for (int i = 0; i < N; i++) { double v = sm[idx & 63]; idx = (int)v + i; acc += v; }
  long long t5 = clock64();
  for (int i = 0; i < N; i++) { __threadfence_block(); sflag = i; }
  long long t6 = clock64();
  for (int i = 0; i < N; i++) { __nanosleep(20); }
  long long t7 = clock64();
  long long c0 = clock64();
  for (int i = 0; i < N; i++) { c0 += clock64() & 1; }
  long long t8 = clock64();
  unsigned b = 0;
  for (int i = 0; i < N; i++) { b += __ballot_sync(0xffffffffu, (u + i + b) & 1); }
  long long t9 = clock64();
  for (int i = 0; i < N; i++) { __threadfence(); *flag = i; }
  long long t10 = clock64();
  double sh = f;
  for (int i = 0; i < N; i++) { sh = __shfl_sync(0xffffffffu, sh, (lane + 1) & 31) + 1.0; }
  long long t11 = clock64();
  if (lane == 0) {
    t[0] = (t1 - t0) / N; t[1] = (t2 - t1) / N; t[2] = (t3 - t2) / N; t[3] = (t4 - t3) / N;
    t[4] = (t5 - t4) / N; t[5] = (t6 - t5) / N; t[6] = (t7 - t6) / N; t[7] = (t8 - t7) / N;
    t[8] = (t9 - t8) / N; t[9] = (t10 - t9) / N; t[10] = (t11 - t10) / N;
  }
  out[lane] = x + z + f + u + acc + c0 + b + sh;
}
int main() {
  double* o; long long* t; int* fl;
  cudaMalloc(&o, 32 * 8); cudaMalloc(&t, 16 * 8); cudaMalloc(&fl, 4);
  for (int r = 0; r < 2; r++) k<<<1, 32>>>(o, t, fl);
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, t, 16 * 8, cudaMemcpyDeviceToHost);
  const char* nm[] = {"ddiv+dadd chain", "dmul+dadd chain", "fit_fitness chain", "redux.max chain", "smem ld->cvt->use chain",
                      "fence_block+st.shared", "nanosleep(20)", "clock64", "ballot chain", "threadfence+st.global", "shfl f64 + dadd"};
  for (int i = 0; i < 11; i++) printf("%-28s %lld cycles\n", nm[i], h[i]);
  return 0;
}
