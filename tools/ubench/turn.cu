// Latency of the matcher's serial "turn" (one warp): read the newest log entry from shared
// memory, evaluate the job against it, warp-argmax with precomputed per-lane items, append
// the winner to the log.  Each iteration depends on the previous one through shared memory.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LOGN = 1024;
struct Sh {
  int l_vm[LOGN]; double l_ac[LOGN], l_am[LOGN], l_lc[LOGN], l_lm[LOGN], l_rc[LOGN], l_rm[LOGN];
  int l_an[LOGN], l_pu[LOGN], l_k[LOGN], l_pu0[LOGN];
  int latest[8192];
  volatile int ncommit, gdone;
};
__device__ __forceinline__ bool better(double f, int v, double g, int w) { return f > g || (f == g && v < w); }
__device__ __forceinline__ double fit_branchy(double jc, double jm, double ac, double am, double lc, double lm, double rc, double rm) {
  if (ac + jc > lc) return 0.0;
  if (am + jm > lm) return 0.0;
  double cpu_fit = ((jc + ac) + rc) / (lc + rc);
  double mem_fit = ((jm + am) + rm) / (lm + rm);
  return (cpu_fit + mem_fit) / 2.0;
}
__device__ __forceinline__ double fit_select(double jc, double jm, double ac, double am, double lc, double lm, double rc, double rm) {
  const bool no = (ac + jc > lc) | (am + jm > lm);
  double cpu_fit = ((jc + ac) + rc) / (lc + rc);
  double mem_fit = ((jm + am) + rm) / (lm + rm);
  double f = (cpu_fit + mem_fit) * 0.5;
  return no ? 0.0 : f;
}
__device__ __forceinline__ double argmax_fast(double f, int v, int& wv, int& wl) {
  const unsigned hi = (unsigned)__double2hiint(f);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  unsigned cand = __ballot_sync(0xffffffffu, hi == mh);
  if (mh == 0u) { wv = 0x7fffffff; wl = 0; return 0.0; }
  if (__popc(cand) > 1) {
    const unsigned lo = hi == mh ? (unsigned)__double2loint(f) : 0u;
    const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
    cand = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    if (__popc(cand) > 1) {
      const unsigned key = ((cand >> (threadIdx.x & 31)) & 1u) ? (unsigned)v : 0xffffffffu;
      const unsigned mk = __reduce_min_sync(0xffffffffu, key);
      cand = __ballot_sync(0xffffffffu, key == mk);
    }
  }
  wl = __ffs(cand) - 1;
  wv = __shfl_sync(0xffffffffu, v, wl);
  return __shfl_sync(0xffffffffu, f, wl);
}
// V: 0 = per-lane items + argmax (current design); 1 = same with branch-free fit;
//    2 = uniform top-2 scheme (no collectives on the chain)
template <int V>
__global__ void k(long long* out, double* sink, int iters) {
  extern __shared__ unsigned char raw[];
  Sh& S = *reinterpret_cast<Sh*>(raw);
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < LOGN; i += blockDim.x) {
    S.l_vm[i] = i & 4095; S.l_ac[i] = 1.0; S.l_am[i] = 1024.0; S.l_lc[i] = 64.0; S.l_lm[i] = 262144.0;
    S.l_rc[i] = 8.0; S.l_rm[i] = 4096.0; S.l_an[i] = 0; S.l_pu[i] = 0;
  }
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) S.latest[i] = -1;
  if (threadIdx.x == 0) { S.ncommit = 1; S.gdone = 0; }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // per-lane "old" items (as if evaluated before the turn)
  double yf = 0.3 + lane * 0.001; int yv = 5000 + lane;
  double b1f = 0.5, b2f = 0.4; int b1v = 6000, b2v = 6001;
  double acc = 0;
  long long t0 = clock64();
  int c = 1;
  for (int it = 0; it < iters; it++) {
    const double jc = 0.5, jm = 512.0;
    // wait for the turn (already ours) and the newest entry
    while (S.gdone != it) {}
    const int cn = S.ncommit;
    const int idx = (cn - 1) & (LOGN - 1);
    const int vm = S.l_vm[idx];
    const double ac = S.l_ac[idx], am = S.l_am[idx], lc = S.l_lc[idx], lm = S.l_lm[idx], rc = S.l_rc[idx], rm = S.l_rm[idx];
    const int an = S.l_an[idx], pu = S.l_pu[idx];
    int wv; double wf; double w_ac, w_am, w_lc, w_lm, w_rc, w_rm; int w_an, w_pu; bool writer;
    if (V == 0 || V == 1) {
      double xf = 0.0;
      if (lane == (it & 31)) xf = V == 0 ? fit_branchy(jc, jm, ac, am, lc, lm, rc, rm) : fit_select(jc, jm, ac, am, lc, lm, rc, rm);
      const bool xin = xf > 0.0 && S.latest[vm] < cn;
      const bool ok = S.latest[yv] < 0;
      const bool use_y = ok && (!xin || better(yf, yv, xf, vm));
      const double lf = use_y ? yf : (xin ? xf : 0.0);
      const int lv = use_y ? yv : (xin ? vm : 0x7fffffff);
      int wl;
      wf = argmax_fast(lf, lv, wv, wl);
      const bool exact = wf > 0.0 && better(wf, wv, 0.1, 77);
      writer = exact && lane == wl;
      w_ac = ac; w_am = am; w_lc = lc; w_lm = lm; w_rc = rc; w_rm = rm; w_an = an; w_pu = pu;
    } else {
      const double xf = fit_select(jc, jm, ac, am, lc, lm, rc, rm);  // uniform
      const bool first = b1v != vm;
      const double pf = first ? b1f : b2f; const int pv = first ? b1v : b2v;
      const bool takex = better(xf, vm, pf, pv);
      wf = takex ? xf : pf; wv = takex ? vm : pv;
      writer = lane == 0 && wf > 0.0;
      w_ac = ac; w_am = am; w_lc = lc; w_lm = lm; w_rc = rc; w_rm = rm; w_an = an; w_pu = pu;
    }
    if (writer) {
      const int o = c & (LOGN - 1);
      S.l_vm[o] = wv & 4095; S.l_ac[o] = w_ac + 0.0; S.l_am[o] = w_am + 0.0; S.l_lc[o] = w_lc; S.l_lm[o] = w_lm;
      S.l_rc[o] = w_rc; S.l_rm[o] = w_rm; S.l_an[o] = w_an + 1; S.l_pu[o] = w_pu; S.l_k[o] = it; S.l_pu0[o] = w_pu;
      S.latest[wv & 4095] = c;
      asm volatile("fence.acq_rel.cta;" ::: "memory");
      S.ncommit = c + 1;
      S.gdone = it + 1;
    }
    c++;
    acc += wf;
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) { out[0] = (t1 - t0) / iters; sink[0] = acc; }
}

struct __align__(16) Ent { int vm, an, pu, k; double ac, am, lc, lm, rc, rm, yc, ym; };  // 80 B
struct Sh2 { Ent e[LOGN]; volatile unsigned long long flag; };  // flag = (gdone << 32) | ncommit
__device__ __forceinline__ double div_rcp(double a, double b, double y) {  // y = RN(1/b)
  const double q0 = a * y;
  const double r = fma(-b, q0, a);
  return fma(r, y, q0);
}
template <int V>
__global__ void k2(long long* out, double* sink, int iters) {
  extern __shared__ unsigned char raw[];
  Sh2& S = *reinterpret_cast<Sh2*>(raw);
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < LOGN; i += blockDim.x) {
    Ent e; e.vm = i & 4095; e.an = 0; e.pu = 0; e.k = 0; e.ac = 1.0; e.am = 1024.0; e.lc = 64.0; e.lm = 262144.0;
    e.rc = 8.0; e.rm = 4096.0; e.yc = 1.0 / (e.lc + e.rc); e.ym = 1.0 / (e.lm + e.rm);
    S.e[i] = e;
  }
  if (threadIdx.x == 0) S.flag = 1ull;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  double b1f = 0.5, b2f = 0.4; int b1v = 6000, b2v = 6001;
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double jc = 0.5, jm = 512.0;
    unsigned long long fl;
    do { fl = S.flag; } while ((int)(fl >> 32) != it);
    const int cn = (int)(unsigned)fl;
    const Ent* ep = &S.e[(cn - 1) & (LOGN - 1)];
    const int4 h = *reinterpret_cast<const int4*>(ep);
    const double2 a0 = *reinterpret_cast<const double2*>(&ep->ac);
    const double2 a1 = *reinterpret_cast<const double2*>(&ep->lc);
    const double2 a2 = *reinterpret_cast<const double2*>(&ep->rc);
    const double2 a3 = *reinterpret_cast<const double2*>(&ep->yc);
    const double ac = a0.x, am = a0.y, lc = a1.x, lm = a1.y, rc = a2.x, rm = a2.y;
    const bool no = (ac + jc > lc) | (am + jm > lm);
    double cf, mf;
    if (V == 3) { cf = ((jc + ac) + rc) / (lc + rc); mf = ((jm + am) + rm) / (lm + rm); }
    else { cf = div_rcp((jc + ac) + rc, lc + rc, a3.x); mf = div_rcp((jm + am) + rm, lm + rm, a3.y); }
    const double xf = no ? 0.0 : (cf + mf) * 0.5;
    const bool first = b1v != h.x;
    const double pf = first ? b1f : b2f; const int pv = first ? b1v : b2v;
    const bool takex = better(xf, h.x, pf, pv);
    const double wf = takex ? xf : pf; const int wv = takex ? h.x : pv;
    if (lane == 0 && wf > 0.0) {
      Ent* o = &S.e[cn & (LOGN - 1)];
      *reinterpret_cast<int4*>(o) = make_int4(wv & 4095, h.y + 1, h.z, it);
      *reinterpret_cast<double2*>(&o->ac) = make_double2(ac + 0.0, am + 0.0);
      *reinterpret_cast<double2*>(&o->lc) = a1;
      *reinterpret_cast<double2*>(&o->rc) = a2;
      *reinterpret_cast<double2*>(&o->yc) = a3;
      S.flag = ((unsigned long long)(it + 1) << 32) | (unsigned)(cn + 1);
    }
    acc += wf;
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) { out[0] = (t1 - t0) / iters; sink[0] = acc; }
}
int main() {
  long long* o; double* sk;
  cudaMalloc(&o, 8); cudaMalloc(&sk, 8);
  const int smem = sizeof(Sh);
  const char* nm[3] = {"lanes+argmax, branchy fit", "lanes+argmax, select fit", "uniform top-2 (no collectives)"};
  for (int v = 0; v < 3; v++) {
    auto kf = v == 0 ? k<0> : (v == 1 ? k<1> : k<2>);
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int r = 0; r < 2; r++) kf<<<1, 128, smem>>>(o, sk, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %lld cycles/turn (%s)\n", nm[v], h, cudaGetErrorString(e));
  }
  for (int v = 3; v <= 4; v++) {
    auto kf = v == 3 ? k2<3> : k2<4>;
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sh2));
    for (int r = 0; r < 2; r++) kf<<<1, 128, sizeof(Sh2)>>>(o, sk, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %lld cycles/turn (%s)\n", v == 3 ? "uniform, AoS 128-bit, 1 flag word" : "  + reciprocal division", h, cudaGetErrorString(e));
  }
  return 0;
}
