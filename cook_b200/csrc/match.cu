// match.cu — considerable-job filter (M0) and the exact greedy best-fit matcher
// (M3/M4) on the GPU.  Replaces pending-jobs->considerable-jobs
// (scheduler/scheduler.clj:729-762, tools.clj:903-973) and Fenzo's
// TaskScheduler.scheduleOnce as Cook calls it (scheduler.clj:665-671) with
// good-enough-fitness >= 1.0 (every VM evaluated for every task).
//
// Exactness on a parallel machine.  Fenzo places requests one at a
// time; each placement mutates one VM and may change every later argmax.  The
// kernel keeps that order but splits the work:
//
//   evaluators (all CTAs but #0): for a block of B jobs, one warp per job scans
//     ALL offers against a SNAPSHOT of the dynamic VM state and keeps, per lane
//     (= chunk of offers v == lane mod 32), the best two (fitness, v) pairs.
//   resolver (warp 0 of CTA 0): walks the jobs in rank order.  VMs touched since
//     the snapshot ("dirty", <= 2B of them, state in shared memory) are
//     re-evaluated exactly; every other VM is unchanged, so the row's best
//     clean candidate per chunk is still exact.  If a chunk's two candidates
//     are both dirty, its remaining VMs are bounded above by the second
//     candidate's fitness; the chunk is re-scanned only when that bound could
//     beat the winner.  Jobs whose constraints depend on same-cycle placements
//     of other jobs (groups) take a full re-scan against current state.
//
//   The two roles are software-pipelined: while the resolver places block t the
//   evaluators score block t+1 against the state published after block t-1
//   (double-buffered), one grid barrier per block.
//
// Tie-break: equal fitness => lowest hostname (offers are index-sorted by
// name_rank on upload, so "lowest v").  All f64 ops are IEEE (div.rn.f64,
// -fmad=false): identical to oracle/cook_oracle.cpp eval_pair bit for bit.
#include <cooperative_groups.h>

#include <algorithm>
#include <numeric>

#include "common.cuh"
#include "sort.cuh"

namespace {

constexpr int RES_THREADS = 256;   // threads per CTA of the match kernel
constexpr int MAXB = 256;          // max jobs per block
constexpr int MAXD = 2 * MAXB;     // dirty list capacity (two blocks)
constexpr int TOPK = 8;            // candidates kept per (job, chunk)
constexpr int ROW_V_OFF = TOPK * 32 * 8;             // byte offset of the v part of a row
constexpr int ROW_BYTES = TOPK * 32 * (8 + 4);       // 3072 B per job

// Per-VM state is AoS, 32 B per record, so one record is two 128-bit loads and a
// clean candidate's state can be staged with 16-byte async copies.
struct __align__(32) VmStatic { double lc, lm, rc, rm; };
struct __align__(32) VmDyn { double ac, am; int an, pu; int pad0, pad1; };

struct JobDev {   // columns in ORIGINAL job index space (may be null)
  const int32_t* user;
  const double* cpus;
  const double* mem;
  const double* gpus;
  const int32_t* ports;
  const uint8_t* allowed;
  const uint8_t* plugin;
  const int32_t* novel_off; const int32_t* novel_host;
  const int32_t* gpu_model;
  const double* disk_request; const int32_t* disk_type;
  const int32_t* attr_off; const int32_t* attr_col; const int32_t* attr_val;
  const int64_t* est_end_ms;
  const int32_t* ckpt_location;
  const int32_t* reserved_host;
  const int32_t* group_off; const int32_t* group_idx;
};

struct OfferDev {
  int O;
  // hot columns, gathered into rank-sorted index space v
  const VmStatic* vs;   // {lease cpus, lease mem, running cpus, running mem} per VM
  const int32_t* perm;  // v -> original offer index
  // constraint columns, ORIGINAL index space (may be null)
  const int32_t* hostname_id;
  const int32_t* run_count;
  const int32_t* ports_total;  // computed on device; null when no ports
  const int32_t* port_off; const int32_t* port_begin; const int32_t* port_end;
  const uint8_t* is_k8s;
  const int32_t* location;
  const int32_t* gpu_off; const int32_t* gpu_model; const double* gpu_count;
  const int32_t* disk_off; const int32_t* disk_type; const double* disk_space;
  const int32_t* max_tasks; const int32_t* num_tasks;
  const int64_t* host_start;
  int n_attr_cols; const int32_t* attr;
  const uint8_t* reserved;
};

struct GroupDev {
  int n_groups;
  const int32_t* kind; const int32_t* attr_col; const int32_t* minimum;
  const int32_t* cot_off; const int32_t* cot_host; const int32_t* cot_attr;
  const int32_t* gp_off;  // capacity offsets of per-group placed lists
  int32_t* gp_n;          // dynamic count per group
  int32_t* gp_vm;         // placed VM (v space)
};

struct DynBuf {  // dynamic per-VM state, index space v, double buffered
  VmDyn* d[2];
};

struct MatchArgs {
  JobDev jb;
  OfferDev of;
  GroupDev gr;
  DynBuf dyn;
  int n_cons;
  const int32_t* cons;     // k -> job index
  const double* kc;        // gathered cpus per k
  const double* km;        // gathered mem per k
  const uint8_t* kflags;   // bit0: has groups
  int B;                   // jobs per block
  int host_lifetime_mins;
  unsigned char* rows;     // [2][B][ROW_BYTES]: f[TOPK][32] f64 then v[TOPK][32] i32
  const double* kg;        // gathered gpus per k (constraint kernel)
  const int32_t* kports;   // gathered port counts per k
  uint8_t* feas;           // [2][B] any feasible VM at the snapshot
  unsigned* rows_ready;    // [nblk] rows scored per block
  unsigned* published;     // # blocks resolved and published
  int32_t* assign;         // [n_cons] v (rank space) or -1
  int32_t* ports_start;    // [n_cons] ports_used of the VM before assignment
  uint8_t* fail;           // [n_cons]
  unsigned long long* stats;  // [0]=fast [1]=chunk rescans [2]=full rescans [3]=matched [4]=offers used
};

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct JobRegs {  // per-job values the hot loop keeps in registers
  double c, m, g;
  int j, ports;
};

template <bool CONSTR>
__device__ __forceinline__ JobRegs load_job(const MatchArgs& a, int k) {
  JobRegs r;
  r.c = a.kc[k]; r.m = a.km[k];
  r.j = a.cons[k];
  r.g = 0.0; r.ports = 0;
  if (CONSTR) {
    r.g = a.kg[k];
    r.ports = a.kports[k];
  }
  return r;
}

__device__ __forceinline__ double csr_lookup(const int32_t* off, const int32_t* key,
                                             const double* val, int o, int k) {
  if (!off) return 0.0;
  for (int i = off[o]; i < off[o + 1]; i++)
    if (key[i] == k) return val[i];
  return 0.0;
}

// Static + count-dependent hard constraints of one (job, VM) pair, in Cook's
// evaluation order (see oracle eval_pair; constraints.clj).  Group constraints
// are handled by group_pass().  `an` = tasks assigned to the VM this cycle.
__device__ bool constraints_pass(const MatchArgs& a, const JobRegs& r, int v, int an) {
  const JobDev& jb = a.jb;
  const OfferDev& of = a.of;
  const int o = of.perm[v];
  const int j = r.j;
  if (jb.ckpt_location && jb.ckpt_location[j] >= 0) {
    int loc = of.location ? of.location[o] : -1;
    if (loc != jb.ckpt_location[j]) return false;
  }
  if (jb.est_end_ms && jb.est_end_ms[j] >= 0 && of.host_start && of.host_start[o] >= 0) {
    long long death = 1000LL * of.host_start[o] + 60000LL * a.host_lifetime_mins;
    if (!(jb.est_end_ms[j] < death)) return false;
  }
  if (jb.attr_off) {
    for (int k = jb.attr_off[j]; k < jb.attr_off[j + 1]; k++) {
      int col = jb.attr_col[k], val = jb.attr_val[k];
      if (col < 0 || col >= of.n_attr_cols) return false;
      int hv = of.attr[(size_t)col * of.O + o];
      if (val <= 0 || hv != val) return false;
    }
  }
  const bool k8s = of.is_k8s && of.is_k8s[o];
  if (jb.disk_request && jb.disk_request[j] >= 0.0 && k8s) {
    double space = csr_lookup(of.disk_off, of.disk_type, of.disk_space, o,
                              jb.disk_type ? jb.disk_type[j] : -1);
    if (!(space >= jb.disk_request[j])) return false;
  }
  if (k8s) {
    if (r.g > 0.0) {
      double have = csr_lookup(of.gpu_off, of.gpu_model, of.gpu_count, o,
                               jb.gpu_model ? jb.gpu_model[j] : -1);
      int on_vm = (of.run_count ? of.run_count[o] : 0) + an;
      if (!(have == r.g && on_vm == 0)) return false;
    } else {
      int nmodels = of.gpu_off ? of.gpu_off[o + 1] - of.gpu_off[o] : 0;
      if (nmodels != 0) return false;
    }
  } else if (!(r.g == 0.0)) {
    return false;
  }
  if (jb.novel_off) {
    int h = of.hostname_id[o];
    for (int k = jb.novel_off[j]; k < jb.novel_off[j + 1]; k++)
      if (jb.novel_host[k] == h) return false;
  }
  if (of.max_tasks && of.max_tasks[o] >= 0) {
    int total = (of.num_tasks ? of.num_tasks[o] : 0) + an;
    if (!(total < of.max_tasks[o])) return false;
  }
  if (of.reserved && of.reserved[o]) {
    int mine = jb.reserved_host ? jb.reserved_host[j] : -1;
    if (mine != of.hostname_id[o]) return false;
  }
  return true;
}

__device__ __forceinline__ int vm_attr(const OfferDev& of, int col, int v) {
  return (col >= 0 && col < of.n_attr_cols) ? of.attr[(size_t)col * of.O + of.perm[v]] : 0;
}

// Group constraints (constraints.clj:586-678) against the CURRENT group state
// (running cotasks known to Fenzo + cotasks placed earlier in this cycle).
__device__ bool group_pass(const MatchArgs& a, const JobRegs& r, int v) {
  const JobDev& jb = a.jb;
  const GroupDev& gr = a.gr;
  const OfferDev& of = a.of;
  for (int k = jb.group_off[r.j]; k < jb.group_off[r.j + 1]; k++) {
    const int g = jb.group_idx[k];
    const int kind = gr.kind[g];
    const int c0 = gr.cot_off[g], c1 = gr.cot_off[g + 1];
    const int p0 = gr.gp_off[g], pn = __ldcg(gr.gp_n + g);
    if (kind == COOK_GROUP_UNIQUE) {
      const int h = of.hostname_id[of.perm[v]];
      for (int c = c0; c < c1; c++)
        if (gr.cot_host[c] == h) return false;
      for (int p = 0; p < pn; p++)
        if (__ldcg(gr.gp_vm + p0 + p) == v) return false;
    } else {
      const int col = gr.attr_col[g];
      const int target = vm_attr(of, col, v);
      const int n = (c1 - c0) + pn;
      if (n == 0) continue;
      auto val_at = [&](int i) { return i < c1 - c0 ? gr.cot_attr[c0 + i] : vm_attr(of, col, __ldcg(gr.gp_vm + p0 + i - (c1 - c0))); };
      int tf = 0;
      for (int i = 0; i < n; i++) tf += (val_at(i) == target);
      if (kind == COOK_GROUP_ATTR_EQUALS) {
        if (tf == 0) return false;
      } else {  // balanced
        if (tf == 0) continue;  // (nil? target-freq) => passes
        int mn = 0x7fffffff, mx = 0, distinct = 0;
        for (int i = 0; i < n; i++) {
          int vi = val_at(i), f = 0;
          bool first = true;
          for (int q = 0; q < n; q++) {
            int vq = val_at(q);
            if (vq == vi) { f++; if (q < i) first = false; }
          }
          if (first) { distinct++; mn = min(mn, f); mx = max(mx, f); }
        }
        if (gr.minimum[g] > distinct) mn = 0;
        if (!(mn == mx || tf < mx)) return false;
      }
    }
  }
  return true;
}

// FENZO 3a + 4 (see oracle): resource fit then cpuMemBinPacker fitness.
__device__ __forceinline__ double fit_fitness(double jc, double jm, double ac, double am,
                                              double lc, double lm, double rc, double rm) {
  if (ac + jc > lc) return 0.0;
  if (am + jm > lm) return 0.0;
  double cpu_fit = ((jc + ac) + rc) / (lc + rc);
  double mem_fit = ((jm + am) + rm) / (lm + rm);
  return (cpu_fit + mem_fit) / 2.0;
}

// Full evaluation of (job, VM v) with explicit dynamic state.
template <bool CONSTR>
__device__ __forceinline__ double eval_vm(const MatchArgs& a, const JobRegs& r, int v, double ac,
                                          double am, int an, int pu, double lc, double lm,
                                          double rc, double rm, bool with_groups) {
  if (CONSTR) {
    if (ac + r.c > lc) return 0.0;
    if (am + r.m > lm) return 0.0;
    if (r.ports > 0) {
      int tot = a.of.ports_total ? a.of.ports_total[a.of.perm[v]] : 0;
      if (r.ports > tot - pu) return 0.0;
    }
    if (!constraints_pass(a, r, v, an)) return 0.0;
    if (with_groups && !group_pass(a, r, v)) return 0.0;
  }
  return fit_fitness(r.c, r.m, ac, am, lc, lm, rc, rm);
}

// ------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// TMA (bulk async copy engine): global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------- evaluators
// One CTA scores one job against ALL offers (snapshot state).  Thread t scans
// v = t, t+256, ... (coalesced 128-bit loads); v mod 32 == lane, so every
// thread's VMs belong to chunk `lane`.  Each thread keeps its best TOPK
// (fitness desc, v asc); the 8 warps' lists are tree-merged per lane through
// shared memory into the row.
struct EvalShared {
  double f[RES_THREADS / 32][TOPK][32];
  int32_t v[RES_THREADS / 32][TOPK][32];
};

__device__ __forceinline__ bool better(double f, int v, double g, int w) {
  return f > g || (f == g && v < w);
}

// merge the sorted list E[w] into the sorted register list (f, vv)
__device__ __forceinline__ void merge_list(double* f, int* vv, const EvalShared& E, int w, int lane) {
#pragma unroll 1
  for (int i = 0; i < TOPK; i++) {
    double x = E.f[w][i][lane];
    int xv = E.v[w][i][lane];
    if (!(x > 0.0) || !better(x, xv, f[TOPK - 1], vv[TOPK - 1])) break;  // lists are sorted
    f[TOPK - 1] = x; vv[TOPK - 1] = xv;
#pragma unroll
    for (int q = TOPK - 1; q > 0; q--) {
      if (better(f[q], vv[q], f[q - 1], vv[q - 1])) {
        double tf = f[q]; f[q] = f[q - 1]; f[q - 1] = tf;
        int tv = vv[q]; vv[q] = vv[q - 1]; vv[q - 1] = tv;
      }
    }
  }
}

template <bool CONSTR, bool PROF>
__device__ void evaluate_row(const MatchArgs& a, int k, int blk, EvalShared& E, unsigned long long* ep) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long e0 = PROF ? clock64() : 0;
  const double2* st2 = reinterpret_cast<const double2*>(a.of.vs);
  const double2* dy2 = reinterpret_cast<const double2*>(a.dyn.d[blk & 1]);  // S_{b-2} = buffer b&1
  double f[TOPK];
  int vv[TOPK];
#pragma unroll
  for (int i = 0; i < TOPK; i++) { f[i] = 0.0; vv[i] = -1; }
  const bool grp = CONSTR && (a.kflags[k] & 1);
  if (!grp) {
    const JobRegs r = load_job<CONSTR>(a, k);
#pragma unroll 4
    for (int v = tid; v < a.of.O; v += RES_THREADS) {
      const double2 s0 = __ldg(st2 + 2 * v), s1 = __ldg(st2 + 2 * v + 1);
      const double2 d0 = __ldcg(dy2 + 2 * v);
      int2 d1 = make_int2(0, 0);
      if (CONSTR) d1 = __ldcg(reinterpret_cast<const int2*>(dy2 + 2 * v + 1));
      double x = eval_vm<CONSTR>(a, r, v, d0.x, d0.y, d1.x, d1.y, s0.x, s0.y, s1.x, s1.y, false);
      if (x > f[TOPK - 1]) {  // v ascends within a thread: strict > keeps the lower v on ties
        f[TOPK - 1] = x; vv[TOPK - 1] = v;
#pragma unroll
        for (int i = TOPK - 1; i > 0; i--) {
          if (f[i] > f[i - 1]) {
            double tf = f[i]; f[i] = f[i - 1]; f[i - 1] = tf;
            int tv = vv[i]; vv[i] = vv[i - 1]; vv[i - 1] = tv;
          }
        }
      }
    }
  }
  long long e1 = PROF ? clock64() : 0;
  const int any = __syncthreads_or(f[0] > 0.0 ? 1 : 0);
  if (any) {
    // tree merge across the 8 warps: 3 levels
#pragma unroll
    for (int step = 1; step < RES_THREADS / 32; step <<= 1) {
      if ((warp & (2 * step - 1)) == step) {
#pragma unroll
        for (int i = 0; i < TOPK; i++) { E.f[warp][i][lane] = f[i]; E.v[warp][i][lane] = vv[i]; }
      }
      __syncthreads();
      if ((warp & (2 * step - 1)) == 0) merge_list(f, vv, E, warp + step, lane);
      __syncthreads();
    }
  }
  long long e2 = PROF ? clock64() : 0;
  if (warp == 0) {
    unsigned char* row = a.rows + ((size_t)(blk & 1) * a.B + (k - blk * a.B)) * ROW_BYTES;
    if (any || grp) {
      double* rf = reinterpret_cast<double*>(row);
      int32_t* rv = reinterpret_cast<int32_t*>(row + ROW_V_OFF);
#pragma unroll
      for (int i = 0; i < TOPK; i++) {
        __stcg(rf + i * 32 + lane, f[i]);
        __stcg(rv + i * 32 + lane, vv[i]);
      }
    }
    if (lane == 0) {
      __stcg(a.feas + (size_t)(blk & 1) * a.B + (k - blk * a.B), (uint8_t)((any || grp) ? 1 : 0));
      __threadfence();
      atomicAdd(a.rows_ready + blk, 1u);
    }
  }
  if (PROF) {
    long long e3 = clock64();
    ep[0] += (unsigned long long)(e1 - e0); ep[1] += (unsigned long long)(e2 - e1);
    ep[2] += (unsigned long long)(e3 - e2);
  }
}

// --------------------------------------------------------------- resolver
// CTA 0.  Jobs are resolved in ROUNDS of up to SPEC_W consecutive feasible jobs:
//   P  (parallel, one warp per job): exact best and runner-up VM of the job
//      against the ROUND-START state = clean chunk candidates from its row +
//      exact re-evaluation of all dirty VMs (+ rare chunk re-scans).
//   C  (chain): job i consumes the commits of jobs 0..i-1 of the round in order;
//      a commit changes ONE VM, so job i re-evaluates just that VM and updates
//      its (best, runner-up) pair; when commit i-1 is in, its best is exact and
//      it publishes its own commit.  If a job loses its best and does not know
//      its exact runner-up any more it aborts: the round is truncated there and
//      the job restarts the next round (where it has no predecessors).
//   A  (warp 0): commits are applied to the dirty list, outputs written.
constexpr int SPEC_W = RES_THREADS / 32;
constexpr int ROW_RING = 2 * SPEC_W;

struct VmRec {  // state of one VM as seen by the resolver
  double ac, am, lc, lm, rc, rm;
  int an, pu, slot, pad;  // slot: index in the dirty list, -1 if clean at round start
};

struct CommitRec {
  VmRec st;      // state AFTER this job's placement
  int vm;        // -1: job not placed
  int aborted;
  int pu_before;
  int pad;
};

struct ResolverShared {
  // dirty list (VMs touched since the snapshot), SoA so lane d reads entry d
  double d_ac[MAXD], d_am[MAXD];
  double d_lc[MAXD], d_lm[MAXD], d_rc[MAXD], d_rm[MAXD];
  int32_t d_vm[MAXD], d_an[MAXD], d_pu[MAXD], d_touch[MAXD];
  // per-block job requests
  double jc[MAXB], jm[MAXB], jg[MAXB];
  int32_t jports[MAXB], jj[MAXB];
  uint8_t jgrp[MAXB];
  int16_t feas_list[MAXB];
  // TMA-staged rows of the next feasible jobs
  __align__(128) unsigned char rows[ROW_RING][ROW_BYTES];
  unsigned long long bar[ROW_RING];
  CommitRec commits[SPEC_W];
  volatile int commit_flag[SPEC_W];
  // round header (written by warp 0 between the two CTA barriers)
  volatile int r_mode;   // 0 = speculative round, 1 = exit
  volatile int r_n;      // jobs in the round
  volatile int r_q0;     // index of the round's first job in feas_list
  volatile int r_seq;    // round sequence number (commit_flag target)
  volatile int r_blk, r_k0, r_nD;
  volatile unsigned r_gq0;  // global feasible-job counter of the round's first job (row ring position)
  unsigned long long n_rescan;
};

__device__ __forceinline__ double warp_max_f64(double f) {  // f >= 0
  unsigned hi = (unsigned)__double2hiint(f);
  unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  unsigned lo = hi == mh ? (unsigned)__double2loint(f) : 0u;
  unsigned ml = __reduce_max_sync(0xffffffffu, lo);
  return __hiloint2double((int)mh, (int)ml);
}

// argmax over lanes of (f desc, v asc).  Returns winner fitness; wv/wl = VM and lane.
__device__ __forceinline__ double warp_argmax(double f, int v, int& wv, int& wl) {
  const double wf = warp_max_f64(f);
  unsigned key = (f == wf && wf > 0.0) ? (unsigned)v : 0xffffffffu;
  unsigned mv = __reduce_min_sync(0xffffffffu, key);
  wv = (int)mv;
  wl = __ffs(__ballot_sync(0xffffffffu, key == mv)) - 1;
  return wf;
}

struct Top2 {  // a lane's two best candidates (distinct VMs)
  double f1, f2;
  int v1, v2, s1, s2;
  __device__ __forceinline__ void init() { f1 = f2 = 0.0; v1 = v2 = 0x7fffffff; s1 = s2 = -1; }
  __device__ __forceinline__ void ins(double f, int v, int s) {
    if (!(f > 0.0)) return;
    if (better(f, v, f1, v1)) { f2 = f1; v2 = v1; s2 = s1; f1 = f; v1 = v; s1 = s; }
    else if (better(f, v, f2, v2)) { f2 = f; v2 = v; s2 = s; }
  }
};

__device__ __forceinline__ VmRec load_rec(const MatchArgs& a, const ResolverShared& S, const VmDyn* snapd,
                                          int v, int slot, bool constr) {
  VmRec r;
  r.slot = slot; r.pad = 0;
  if (slot >= 0) {
    r.ac = S.d_ac[slot]; r.am = S.d_am[slot]; r.lc = S.d_lc[slot]; r.lm = S.d_lm[slot];
    r.rc = S.d_rc[slot]; r.rm = S.d_rm[slot]; r.an = S.d_an[slot]; r.pu = S.d_pu[slot];
  } else {
    const double2* st2 = reinterpret_cast<const double2*>(a.of.vs + v);
    const double2 s0 = __ldg(st2), s1 = __ldg(st2 + 1);
    const double2 d0 = __ldcg(reinterpret_cast<const double2*>(snapd + v));
    int2 d1 = make_int2(0, 0);
    if (constr) d1 = __ldcg(reinterpret_cast<const int2*>(snapd + v) + 2);
    r.ac = d0.x; r.am = d0.y; r.an = d1.x; r.pu = d1.y;
    r.lc = s0.x; r.lm = s0.y; r.rc = s1.x; r.rm = s1.y;
  }
  return r;
}

struct SpecResult {
  double bf, sf;
  int bv, sv;          // -1: none
  VmRec brec, srec;
};

// P phase for one job: exact best + runner-up against the round-start state.
template <bool CONSTR>
__device__ void spec_phase(const MatchArgs& a, int blk, int ib, const unsigned char* rowp,
                           ResolverShared& S, const unsigned* bitmap, int nD, SpecResult& out,
                           JobRegs& r) {
  const int lane = threadIdx.x & 31;
  const VmDyn* snapd = a.dyn.d[blk & 1];
  r.c = S.jc[ib]; r.m = S.jm[ib]; r.g = CONSTR ? S.jg[ib] : 0.0;
  r.ports = CONSTR ? S.jports[ib] : 0; r.j = S.jj[ib];
  Top2 t;
  t.init();
  double bound = 0.0;
  {
    const double* rf = reinterpret_cast<const double*>(rowp);
    const int32_t* rv = reinterpret_cast<const int32_t*>(rowp + ROW_V_OFF);
    double f[TOPK];
    int v[TOPK];
#pragma unroll
    for (int i = 0; i < TOPK; i++) { f[i] = rf[i * 32 + lane]; v[i] = rv[i * 32 + lane]; }
    unsigned dirty_bits = 0u, live_bits = 0u;
#pragma unroll
    for (int i = 0; i < TOPK; i++) {
      const bool live = f[i] > 0.0;
      const int vi = live ? v[i] : 0;
      const unsigned w = bitmap[vi >> 5];
      live_bits |= (live ? 1u : 0u) << i;
      dirty_bits |= (((w >> (vi & 31)) & 1u) & (live ? 1u : 0u)) << i;
    }
    unsigned clean = live_bits & ~dirty_bits;  // sorted list: live_bits is a prefix mask
    int nclean = 0;
#pragma unroll
    for (int pick = 0; pick < 2; pick++) {
      if (clean) {
        const int i = __ffs(clean) - 1;
        clean &= clean - 1;
#pragma unroll
        for (int q = 0; q < TOPK; q++)
          if (q == i) t.ins(f[q], v[q], -1);
        nclean++;
      }
    }
    // fewer than two clean entries out of a FULL list: the rest of the chunk is
    // only known to be <= the last entry
    if (nclean < 2 && live_bits == ((1u << TOPK) - 1u)) bound = f[TOPK - 1];
  }
  // exact re-evaluation of dirty VMs against their round-start state
  for (int d = lane; d < nD; d += 32) {
    double f = eval_vm<CONSTR>(a, r, S.d_vm[d], S.d_ac[d], S.d_am[d], S.d_an[d], S.d_pu[d],
                               S.d_lc[d], S.d_lm[d], S.d_rc[d], S.d_rm[d], false);
    t.ins(f, S.d_vm[d], d);
  }
  int bv, bl, sv, sl;
  double bf = warp_argmax(t.f1, t.v1, bv, bl);
  // runner-up: the winner lane offers its second candidate instead
  double sf = warp_argmax(lane == bl ? t.f2 : t.f1, lane == bl ? t.v2 : t.v1, sv, sl);
  const double mb = warp_max_f64(bound);
  if (mb > 0.0 && mb >= sf) {
    // rare: a chunk whose listed candidates are (almost) all dirty could still
    // hold the best or the runner-up => exact re-scan of those chunks
    unsigned need = __ballot_sync(0xffffffffu, bound > 0.0 && bound >= sf);
    const double2* st2 = reinterpret_cast<const double2*>(a.of.vs);
    const double2* dy2 = reinterpret_cast<const double2*>(snapd);
    while (need) {
      const int c = __ffs(need) - 1;
      need &= need - 1;
      for (int v = c + 32 * lane; v < a.of.O; v += 32 * 32) {
        if ((bitmap[v >> 5] >> (v & 31)) & 1u) continue;
        if (v == t.v1 || v == t.v2) continue;  // already a candidate of this lane
        const double2 s0 = st2[2 * v], s1 = st2[2 * v + 1];
        const double2 d0 = __ldcg(dy2 + 2 * v);
        int2 d1 = make_int2(0, 0);
        if (CONSTR) d1 = __ldcg(reinterpret_cast<const int2*>(dy2 + 2 * v + 1));
        t.ins(eval_vm<CONSTR>(a, r, v, d0.x, d0.y, d1.x, d1.y, s0.x, s0.y, s1.x, s1.y, false), v, -1);
      }
      if (lane == 0) S.n_rescan++;
    }
    // a re-scanned VM may also sit in another lane's list (lane c's own clean
    // entries): duplicates are harmless for the max, but the runner-up must be a
    // DIFFERENT VM than the winner, so mask the winner VM explicitly.
    bf = warp_argmax(t.f1, t.v1, bv, bl);
    const bool a1 = t.v1 != bv;
    sf = warp_argmax(a1 ? t.f1 : t.f2, a1 ? t.v1 : t.v2, sv, sl);
  }
  out.bf = bf; out.sf = sf;
  out.bv = bf > 0.0 ? bv : -1;
  out.sv = sf > 0.0 ? sv : -1;
  if (out.bv >= 0) {
    const int slot = __shfl_sync(0xffffffffu, t.s1, bl);
    out.brec = load_rec(a, S, snapd, out.bv, slot, CONSTR);
  }
  if (out.sv >= 0) {
    int slot = __shfl_sync(0xffffffffu, (t.v1 == out.sv) ? t.s1 : t.s2, sl);
    out.srec = load_rec(a, S, snapd, out.sv, slot, CONSTR);
  }
}

// C phase: consume predecessors' commits in order, then publish own commit.
template <bool CONSTR>
__device__ void chain_phase(const MatchArgs& a, ResolverShared& S, int wi, int seq, const JobRegs& r,
                            SpecResult& sp) {
  const int lane = threadIdx.x & 31;
  bool second_known = true;
  bool aborted = false;
  for (int j = 0; j < wi && !aborted; j++) {
    while (S.commit_flag[j] != seq) __nanosleep(20);
    __threadfence_block();
    const CommitRec& c = S.commits[j];
    if (c.aborted) { aborted = true; break; }
    const int x = c.vm;
    if (x < 0) continue;
    const VmRec st = c.st;
    const double f = eval_vm<CONSTR>(a, r, x, st.ac, st.am, st.an, st.pu, st.lc, st.lm, st.rc, st.rm, false);
    if (x == sp.bv) {
      if (f > 0.0) { sp.bf = f; sp.brec = st; }  // fuller => fitness grew, still the best
      else if (second_known) {                    // my best no longer fits
        sp.bf = sp.sf; sp.bv = sp.sv; sp.brec = sp.srec;
        sp.sf = 0.0; sp.sv = -1;
        second_known = sp.bv < 0;  // nothing left at all => trivially known
      } else aborted = true;
    } else if (x == sp.sv) {
      if (f > 0.0) {
        sp.sf = f; sp.srec = st;
        if (better(sp.sf, sp.sv, sp.bf, sp.bv)) {
          double tf = sp.bf; sp.bf = sp.sf; sp.sf = tf;
          int tv = sp.bv; sp.bv = sp.sv; sp.sv = tv;
          VmRec tr = sp.brec; sp.brec = sp.srec; sp.srec = tr;
        }
      } else { sp.sf = 0.0; sp.sv = -1; second_known = false; }
    } else if (f > 0.0) {
      if (sp.bv < 0 || better(f, x, sp.bf, sp.bv)) {
        if (sp.bv >= 0) { sp.sf = sp.bf; sp.sv = sp.bv; sp.srec = sp.brec; second_known = true; }
        sp.bf = f; sp.bv = x; sp.brec = st;
      } else if (second_known && (sp.sv < 0 || better(f, x, sp.sf, sp.sv))) {
        sp.sf = f; sp.sv = x; sp.srec = st;
      }
    }
  }
  if (lane == 0) {
    CommitRec& m = S.commits[wi];
    m.aborted = aborted ? 1 : 0;
    m.vm = -1;
    m.pu_before = 0;
    if (!aborted && sp.bv >= 0 && sp.bf > 0.0) {
      VmRec st = sp.brec;
      m.pu_before = st.pu;
      st.ac = st.ac + r.c;
      st.am = st.am + r.m;
      st.an += 1;
      st.pu += r.ports;
      m.st = st;
      m.vm = sp.bv;
    }
    __threadfence_block();
    S.commit_flag[wi] = seq;
  }
  __syncwarp();
}

// Single-warp exact resolution of one job against live state (group jobs).
template <bool CONSTR>
__device__ void resolve_group_job(const MatchArgs& a, int blk, int k, int ib, ResolverShared& S,
                                  unsigned* bitmap, int& nD, unsigned long long* lstats) {
  const int lane = threadIdx.x & 31;
  const VmDyn* snapd = a.dyn.d[blk & 1];
  JobRegs r;
  r.c = S.jc[ib]; r.m = S.jm[ib]; r.g = CONSTR ? S.jg[ib] : 0.0;
  r.ports = CONSTR ? S.jports[ib] : 0; r.j = S.jj[ib];
  double cf = 0.0;
  int cv = 0x7fffffff, cslot = -1;
  for (int d = lane; d < nD; d += 32) {
    double f = eval_vm<CONSTR>(a, r, S.d_vm[d], S.d_ac[d], S.d_am[d], S.d_an[d], S.d_pu[d],
                               S.d_lc[d], S.d_lm[d], S.d_rc[d], S.d_rm[d], true);
    int v = S.d_vm[d];
    if (f > cf || (f == cf && f > 0.0 && v < cv)) { cf = f; cv = v; cslot = d; }
  }
  const double2* st2 = reinterpret_cast<const double2*>(a.of.vs);
  const double2* dy2 = reinterpret_cast<const double2*>(snapd);
  for (int v = lane; v < a.of.O; v += 32) {
    if ((bitmap[v >> 5] >> (v & 31)) & 1u) continue;
    const double2 s0 = st2[2 * v], s1 = st2[2 * v + 1];
    const double2 d0 = __ldcg(dy2 + 2 * v);
    int2 d1 = make_int2(0, 0);
    if (CONSTR) d1 = __ldcg(reinterpret_cast<const int2*>(dy2 + 2 * v + 1));
    double f = eval_vm<CONSTR>(a, r, v, d0.x, d0.y, d1.x, d1.y, s0.x, s0.y, s1.x, s1.y, true);
    if (f > cf || (f == cf && f > 0.0 && v < cv)) { cf = f; cv = v; cslot = -1; }
  }
  lstats[2]++;
  int wv, wl;
  const double wf = warp_argmax(cf, cv, wv, wl);
  if (wf > 0.0) {
    int slot = __shfl_sync(0xffffffffu, cslot, wl);
    if (slot < 0) {
      slot = nD;
      if (lane == 0) {
        const VmRec rec = load_rec(a, S, snapd, wv, -1, CONSTR);
        S.d_vm[slot] = wv;
        S.d_ac[slot] = rec.ac; S.d_am[slot] = rec.am; S.d_an[slot] = rec.an; S.d_pu[slot] = rec.pu;
        S.d_lc[slot] = rec.lc; S.d_lm[slot] = rec.lm; S.d_rc[slot] = rec.rc; S.d_rm[slot] = rec.rm;
        bitmap[wv >> 5] |= 1u << (wv & 31);
      }
      nD++;
    }
    if (lane == 0) {
      a.ports_start[k] = S.d_pu[slot];
      S.d_ac[slot] = S.d_ac[slot] + r.c;
      S.d_am[slot] = S.d_am[slot] + r.m;
      S.d_an[slot] += 1;
      S.d_pu[slot] += r.ports;
      S.d_touch[slot] = blk;
      a.assign[k] = wv;
      a.fail[k] = COOK_FAIL_NONE;
      if (CONSTR) {
        for (int q = a.jb.group_off[r.j]; q < a.jb.group_off[r.j + 1]; q++) {
          int g = a.jb.group_idx[q];
          int n = __ldcg(a.gr.gp_n + g);
          a.gr.gp_vm[a.gr.gp_off[g] + n] = wv;
          __threadfence_block();
          a.gr.gp_n[g] = n + 1;
        }
      }
    }
    lstats[3]++;
  } else if (lane == 0) {
    a.assign[k] = -1; a.fail[k] = COOK_FAIL_CONSTRAINT;
  }
  __syncwarp();
}

// A phase (warp 0): apply the round's commits to the dirty list; returns the
// number of jobs of the round that are final.
__device__ int apply_round(const MatchArgs& a, ResolverShared& S, unsigned* bitmap, int& nD, int blk,
                           int k0, int q0, int n, unsigned long long* lstats) {
  const int lane = threadIdx.x & 31;
  const bool in = lane < n;
  const int ab = in ? S.commits[lane].aborted : 0;
  const unsigned abm = __ballot_sync(0xffffffffu, in && ab);
  const int n_done = abm ? (__ffs(abm) - 1) : n;
  const bool act = lane < n_done;
  const int vm = act ? S.commits[lane].vm : -1;
  int slot = (act && vm >= 0) ? S.commits[lane].st.slot : -1;
  // first / last commit of each distinct VM inside the round
  bool first = act && vm >= 0, last = act && vm >= 0;
  int first_lane = lane;
  for (int j = 0; j < SPEC_W; j++) {
    const int vj = __shfl_sync(0xffffffffu, vm, j);
    if (vm >= 0 && vj == vm) {
      if (j < lane) { first = false; if (j < first_lane) first_lane = j; }
      if (j > lane) last = false;
    }
  }
  // new dirty slots for VMs that were clean at round start
  const unsigned newm = __ballot_sync(0xffffffffu, first && slot < 0);
  if (first && slot < 0) slot = nD + __popc(newm & ((1u << lane) - 1u));
  // a later commit of the same VM carries the slot of the first one
  const int fslot = __shfl_sync(0xffffffffu, slot, first_lane);
  if (act && vm >= 0 && !first) slot = fslot;
  if (last) {
    const VmRec st = S.commits[lane].st;
    S.d_vm[slot] = vm; S.d_ac[slot] = st.ac; S.d_am[slot] = st.am; S.d_an[slot] = st.an;
    S.d_pu[slot] = st.pu; S.d_lc[slot] = st.lc; S.d_lm[slot] = st.lm; S.d_rc[slot] = st.rc;
    S.d_rm[slot] = st.rm; S.d_touch[slot] = blk;
  }
  if (first && S.commits[lane].st.slot < 0) atomicOr(&bitmap[vm >> 5], 1u << (vm & 31));
  if (act) {
    const int k = k0 + S.feas_list[q0 + lane];
    a.assign[k] = vm;
    a.fail[k] = vm >= 0 ? COOK_FAIL_NONE : COOK_FAIL_CONSTRAINT;
    if (vm >= 0) a.ports_start[k] = S.commits[lane].pu_before;
  }
  nD += __popc(newm);
  lstats[3] += __popc(__ballot_sync(0xffffffffu, act && vm >= 0));
  lstats[0] += n_done;
  __syncwarp();
  return n_done;
}

// Warp 0 between rounds: block transitions (publish, wait for rows, compaction,
// feasibility list), TMA row issue, next round header.
template <bool CONSTR, bool PROF>
struct Driver {
  int blk, nblk, k0, nj, nfeas, qdone, qissued, nD;
  unsigned gq;  // global feasible counter at qdone
};

template <bool CONSTR>
__device__ void begin_block(const MatchArgs& a, ResolverShared& S, unsigned* bitmap, int blk, int& nD,
                            int& nfeas, unsigned long long* lstats) {
  const int lane = threadIdx.x & 31;
  const int k0 = blk * a.B;
  const int k1 = min(k0 + a.B, a.n_cons);
  const int nj = k1 - k0;
  for (int i = lane; i < nj; i += 32) {
    S.jc[i] = a.kc[k0 + i]; S.jm[i] = a.km[k0 + i]; S.jj[i] = a.cons[k0 + i];
    if (CONSTR) { S.jg[i] = a.kg[k0 + i]; S.jports[i] = a.kports[k0 + i]; S.jgrp[i] = a.kflags[k0 + i] & 1; }
  }
  // drop dirty entries not touched in the previous block: they are part of the
  // snapshot this block's rows were scored against.
  {
    int keep_n = 0;
    for (int base = 0; base < nD; base += 32) {
      int d = base + lane;
      bool keep = d < nD && S.d_touch[d] >= blk - 1;
      unsigned kb = __ballot_sync(0xffffffffu, keep);
      int vm = 0, an = 0, pu = 0, tc = 0; double ac = 0, am = 0, lc = 0, lm = 0, rc = 0, rm = 0;
      if (d < nD) {
        vm = S.d_vm[d]; an = S.d_an[d]; pu = S.d_pu[d]; tc = S.d_touch[d];
        ac = S.d_ac[d]; am = S.d_am[d]; lc = S.d_lc[d]; lm = S.d_lm[d]; rc = S.d_rc[d]; rm = S.d_rm[d];
        if (!keep) atomicAnd(&bitmap[vm >> 5], ~(1u << (vm & 31)));
      }
      __syncwarp();
      if (keep) {
        int t = keep_n + __popc(kb & ((1u << lane) - 1u));
        S.d_vm[t] = vm; S.d_an[t] = an; S.d_pu[t] = pu; S.d_touch[t] = tc;
        S.d_ac[t] = ac; S.d_am[t] = am; S.d_lc[t] = lc; S.d_lm[t] = lm; S.d_rc[t] = rc; S.d_rm[t] = rm;
      }
      keep_n += __popc(kb);
      __syncwarp();
    }
    nD = keep_n;
  }
  // rows of this block ready?
  if (lane == 0) {
    while (ld_acquire_u32(a.rows_ready + blk) < (unsigned)nj) __nanosleep(20);
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async-proxy (TMA) reads
  }
  __syncwarp();
  // jobs with no feasible VM at the snapshot are unplaceable now too (resources
  // and count constraints only tighten within a cycle): skip them wholesale.
  const uint8_t* feas = a.feas + (size_t)(blk & 1) * a.B;
  nfeas = 0;
  for (int base = 0; base < nj; base += 32) {
    const int i = base + lane;
    const bool valid = i < nj;
    const bool fz = valid && __ldcg(feas + i) != 0;
    if (valid && !fz) { a.assign[k0 + i] = -1; a.fail[k0 + i] = COOK_FAIL_RESOURCES; }
    const unsigned mask = __ballot_sync(0xffffffffu, fz);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    lstats[0] += __popc(vmask & ~mask);
    if (fz) S.feas_list[nfeas + __popc(mask & ((1u << lane) - 1u))] = (int16_t)i;
    nfeas += __popc(mask);
  }
  __syncwarp();
}

__device__ void end_block(const MatchArgs& a, ResolverShared& S, int blk, int nD) {
  // publish every dirty entry (touched in this or the previous block) into the
  // buffer the evaluators read for block blk+2.
  const int lane = threadIdx.x & 31;
  VmDyn* pub = a.dyn.d[blk & 1];
  for (int d = lane; d < nD; d += 32) {
    int v = S.d_vm[d];
    __stcg(reinterpret_cast<double2*>(pub + v), make_double2(S.d_ac[d], S.d_am[d]));
    __stcg(reinterpret_cast<int2*>(pub + v) + 2, make_int2(S.d_an[d], S.d_pu[d]));
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.published), "r"((unsigned)(blk + 1)) : "memory");
  }
  __syncwarp();
}

// Pipeline.  The resolver CTA places block t while the evaluator CTAs score
// block t+1 against the state published after block t-1 (buffer (t+1)&1).
// Synchronisation is by two monotone counters only:
//   rows_ready[b]  evaluators -> resolver (one arrival per scored row)
//   published      resolver -> evaluators (# blocks resolved and published)
template <bool CONSTR, bool PROF>
__global__ void __launch_bounds__(RES_THREADS, 1) match_kernel(MatchArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nblk = (a.n_cons + a.B - 1) / a.B;
  if (blockIdx.x == 0) {
    ResolverShared& S = *reinterpret_cast<ResolverShared*>(smem_raw);
    unsigned* bitmap = reinterpret_cast<unsigned*>(smem_raw + sizeof(ResolverShared));
    const int words = (a.of.O + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < words; i += RES_THREADS) bitmap[i] = 0u;
    if (threadIdx.x == 0) {
      for (int i = 0; i < ROW_RING; i++) mbar_init(&S.bar[i], 1);
      for (int i = 0; i < SPEC_W; i++) S.commit_flag[i] = 0;
      S.n_rescan = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- driver state (warp 0 only)
    int blk = 0, k0 = 0, nfeas = 0, qdone = 0, qissued = 0, nD = 0, seq = 0;
    unsigned gq = 0;  // global feasible-job counter at qdone (row ring position)
    bool open = false;  // a block is open
    unsigned long long lstats[4] = {0, 0, 0, 0};
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_start = clock64();
    while (true) {
      if (warp == 0) {
        long long td0 = PROF ? clock64() : 0;
        // advance to a state where a round can be issued (or exit)
        int mode = 0, n = 0;
        while (true) {
          if (!open) {
            if (blk >= nblk) { mode = 1; break; }
            k0 = blk * a.B;
            begin_block<CONSTR>(a, S, bitmap, blk, nD, nfeas, lstats);
            qdone = 0; qissued = 0;
            open = true;
          }
          // keep the row ring full
          if (lane == 0) {
            const unsigned char* gbase = a.rows + (size_t)(blk & 1) * a.B * ROW_BYTES;
            while (qissued < nfeas && qissued < qdone + ROW_RING) {
              const unsigned s = (gq + (unsigned)(qissued - qdone)) % ROW_RING;
              mbar_expect_tx(&S.bar[s], ROW_BYTES);
              tma_load_1d(S.rows[s], gbase + (size_t)S.feas_list[qissued] * ROW_BYTES, ROW_BYTES, &S.bar[s]);
              qissued++;
            }
          }
          qissued = __shfl_sync(0xffffffffu, qissued, 0);
          if (qdone >= nfeas) {  // block finished
            end_block(a, S, blk, nD);
            open = false;
            blk++;
            continue;
          }
          // group-constrained jobs are resolved one at a time against live state
          if (CONSTR && S.jgrp[S.feas_list[qdone]]) {
            const int ib = S.feas_list[qdone];
            // its (unused) row slot still has to be consumed to keep the ring in phase
            mbar_wait(&S.bar[gq % ROW_RING], (gq / ROW_RING) & 1u);
            resolve_group_job<CONSTR>(a, blk, k0 + ib, ib, S, bitmap, nD, lstats);
            qdone++; gq++;
            continue;
          }
          n = 0;
          while (n < SPEC_W && qdone + n < nfeas && !(CONSTR && S.jgrp[S.feas_list[qdone + n]])) n++;
          break;
        }
        seq++;
        if (lane == 0) {
          S.r_mode = mode; S.r_n = n; S.r_q0 = qdone; S.r_seq = seq; S.r_blk = blk; S.r_k0 = k0;
          S.r_nD = nD; S.r_gq0 = gq;
        }
        if (PROF) prof[0] += (unsigned long long)(clock64() - td0);
      }
      __syncthreads();  // (A) round header visible
      if (S.r_mode == 1) break;
      const int rn = S.r_n;
      long long tr0 = PROF ? clock64() : 0;
      if (warp < rn) {
        const int rq0 = S.r_q0, rblk = S.r_blk, rseq = S.r_seq, rnD = S.r_nD;
        const unsigned g = S.r_gq0 + (unsigned)warp;
        const int ib = S.feas_list[rq0 + warp];
        mbar_wait(&S.bar[g % ROW_RING], (g / ROW_RING) & 1u);
        SpecResult sp;
        JobRegs r;
        spec_phase<CONSTR>(a, rblk, ib, S.rows[g % ROW_RING], S, bitmap, rnD, sp, r);
        if (PROF && warp == 0) prof[1] += (unsigned long long)(clock64() - tr0);
        chain_phase<CONSTR>(a, S, warp, rseq, r, sp);
      }
      long long tr1 = PROF ? clock64() : 0;
      __syncthreads();  // (B) all commits of the round published
      if (warp == 0) {
        if (PROF) { prof[2] += (unsigned long long)(clock64() - tr1); prof[4]++; }
        long long ta0 = PROF ? clock64() : 0;
        const int n_done = apply_round(a, S, bitmap, nD, blk, k0, qdone, rn, lstats);
        qdone += n_done;
        gq += (unsigned)n_done;
        if (PROF) { prof[3] += (unsigned long long)(clock64() - ta0); prof[5] += (unsigned long long)(rn - n_done); }
      }
    }
    if (threadIdx.x == 0) {
      a.stats[0] = lstats[0]; a.stats[1] = S.n_rescan; a.stats[2] = lstats[2]; a.stats[3] = lstats[3];
      for (int i = 0; i < 6; i++) a.stats[4 + i] = prof[i];
      a.stats[12] = (unsigned long long)(clock64() - t_start);
    }
  } else {
    EvalShared& E = *reinterpret_cast<EvalShared*>(smem_raw);
    const int n_eval = gridDim.x - 1;
    unsigned long long work = 0, wait = 0;
    unsigned long long ep[3] = {0, 0, 0};
    for (int b = 0; b < nblk; b++) {
      const int k0 = b * a.B, k1 = min(k0 + a.B, a.n_cons);
      long long w0 = clock64();
      // rows of block b need S_{b-2}: published >= b-1
      if (b >= 2) {
        if (threadIdx.x == 0)
          while ((int)ld_acquire_u32(a.published) < b - 1) __nanosleep(32);
        __syncthreads();
      }
      long long w1 = clock64();
      for (int k = k0 + (int)blockIdx.x - 1; k < k1; k += n_eval) evaluate_row<CONSTR, PROF>(a, k, b, E, ep);
      wait += (unsigned long long)(w1 - w0);
      work += (unsigned long long)(clock64() - w1);
    }
    if (blockIdx.x == 1 && threadIdx.x == 0) {
      a.stats[14] = work; a.stats[15] = wait;
      if (PROF) { a.stats[16] = ep[0]; a.stats[17] = ep[1]; a.stats[18] = ep[2]; }
    }
  }
}

// ------------------------------------------------------------ considerable
struct ConsArgs {
  const int32_t* ranked; int n_ranked;
  JobDev jb;
  int n_users;
  const double *q_count, *q_cpus, *q_mem, *q_gpus;
  const double *u_count, *u_cpus, *u_mem, *u_gpus;
  const int32_t* tokens;
  int enforce_rate_limit;
  cook_pool_quota pool_q;
  int num_considerable;
};

struct LessUserPos {
  const int32_t* ranked;
  const int32_t* user;
  __device__ bool operator()(int32_t a, int32_t b) const {
    int ua = user[ranked[a]], ub = user[ranked[b]];
    if (ua != ub) return ua < ub;
    return a < b;
  }
};

__global__ void iota_k(int32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void cons_seg_kernel(const int32_t* pos_by_user, const int32_t* ranked,
                                const int32_t* user, int n, int32_t* seg_start, int32_t* seg_end) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int u = user[ranked[pos_by_user[p]]];
  if (p == 0 || user[ranked[pos_by_user[p - 1]]] != u) seg_start[u] = p;
  if (p == n - 1 || user[ranked[pos_by_user[p + 1]]] != u) seg_end[u] = p + 1;
}

// tools.clj:903-915 + :940-959: warp per user, lane-serial left fold over the
// user's queued jobs in queue order, starting from the user's running usage.
__global__ void __launch_bounds__(128) cons_user_kernel(ConsArgs a, const int32_t* pos_by_user,
                                                        const int32_t* seg_start,
                                                        const int32_t* seg_end, uint8_t* keep) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (u >= a.n_users) return;
  const int s = seg_start[u], e = seg_end[u];
  if (e <= s) return;
  double an = a.u_count ? a.u_count[u] : 0.0, ac = a.u_cpus ? a.u_cpus[u] : 0.0;
  double am = a.u_mem ? a.u_mem[u] : 0.0, ag = a.u_gpus ? a.u_gpus[u] : 0.0;
  const double qn = a.q_count[u], qc = a.q_cpus[u], qm = a.q_mem[u], qg = a.q_gpus[u];
  const int tokens = a.tokens ? a.tokens[u] : 0x7fffffff;
  int seen = 0;
  for (int base = s; base < e; base += 32) {
    int p = base + lane;
    double xc = 0, xm = 0, xg = 0;
    int pos = -1;
    if (p < e) {
      pos = pos_by_user[p];
      int j = a.ranked[pos];
      xc = a.jb.cpus[j]; xm = a.jb.mem[j]; xg = a.jb.gpus ? a.jb.gpus[j] : 0.0;
    }
    double mc = 0, mm = 0, mg = 0, mn = 0;
    int cntn = min(32, e - base);
    for (int l = 0; l < cntn; l++) {
      an = an + 1.0;
      ac = ac + __shfl_sync(0xffffffffu, xc, l);
      am = am + __shfl_sync(0xffffffffu, xm, l);
      ag = ag + __shfl_sync(0xffffffffu, xg, l);
      if (lane == l) { mn = an; mc = ac; mm = am; mg = ag; }
    }
    bool ok = (p < e) && (mn <= qn && mc <= qc && mm <= qm && mg <= qg);
    unsigned ob = __ballot_sync(0xffffffffu, ok);
    int kth = seen + __popc(ob & (0xffffffffu >> (31 - lane)));  // k-th surviving job of the user
    bool limited = kth > tokens;
    if (ok && limited && a.enforce_rate_limit) ok = false;
    if (p < e) keep[pos] = ok ? 1 : 0;
    seen += __popc(ob);
  }
}

// Queue-order pass (single warp): pool quota over survivors (tools.clj:917-933),
// allowed + launch-plugin masks (scheduler.clj:749-750), take N (:751);
// gathers the per-k hot columns.
__global__ void cons_queue_kernel(ConsArgs a, const uint8_t* keep, int32_t* cons, double* kc,
                                  double* km, double* kg, int32_t* kports, uint8_t* kflags,
                                  int32_t* out_n) {
  const int lane = threadIdx.x;
  double pn = 0, pc = 0, pm = 0, pg = 0;
  if (a.pool_q.enabled) {  // (reduce (partial merge-with +) (vals user->usage)), tools.clj:969
    for (int base = 0; base < a.n_users; base += 32) {
      int u = base + lane;
      double xn = (u < a.n_users && a.u_count) ? a.u_count[u] : 0.0;
      double xc = (u < a.n_users && a.u_cpus) ? a.u_cpus[u] : 0.0;
      double xm = (u < a.n_users && a.u_mem) ? a.u_mem[u] : 0.0;
      double xg = (u < a.n_users && a.u_gpus) ? a.u_gpus[u] : 0.0;
      int cntn = min(32, a.n_users - base);
      for (int l = 0; l < cntn; l++) {
        pn = pn + __shfl_sync(0xffffffffu, xn, l);
        pc = pc + __shfl_sync(0xffffffffu, xc, l);
        pm = pm + __shfl_sync(0xffffffffu, xm, l);
        pg = pg + __shfl_sync(0xffffffffu, xg, l);
      }
    }
  }
  int n_out = 0;
  for (int base = 0; base < a.n_ranked && n_out < a.num_considerable; base += 32) {
    int i = base + lane;
    bool k = i < a.n_ranked && keep[i];
    int j = i < a.n_ranked ? a.ranked[i] : 0;
    double xc = 0, xm = 0, xg = 0;
    if (k) { xc = a.jb.cpus[j]; xm = a.jb.mem[j]; xg = a.jb.gpus ? a.jb.gpus[j] : 0.0; }
    if (a.pool_q.enabled) {
      unsigned mask = __ballot_sync(0xffffffffu, k);
      double mc = 0, mm = 0, mg = 0, mn = 0;
      while (mask) {
        int l = __ffs(mask) - 1;
        mask &= mask - 1;
        pn = pn + 1.0;
        pc = pc + __shfl_sync(0xffffffffu, xc, l);
        pm = pm + __shfl_sync(0xffffffffu, xm, l);
        pg = pg + __shfl_sync(0xffffffffu, xg, l);
        if (lane == l) { mn = pn; mc = pc; mm = pm; mg = pg; }
      }
      if (k) k = mn <= a.pool_q.count && mc <= a.pool_q.cpus && mm <= a.pool_q.mem && mg <= a.pool_q.gpus;
    }
    if (k && a.jb.allowed && !a.jb.allowed[j]) k = false;
    if (k && a.jb.plugin && !a.jb.plugin[j]) k = false;
    unsigned kb = __ballot_sync(0xffffffffu, k);
    int slot = n_out + __popc(kb & ((1u << lane) - 1u));
    if (k && slot < a.num_considerable) {
      cons[slot] = j;
      kc[slot] = a.jb.cpus[j];
      km[slot] = a.jb.mem[j];
      kg[slot] = a.jb.gpus ? a.jb.gpus[j] : 0.0;
      kports[slot] = a.jb.ports ? a.jb.ports[j] : 0;
      uint8_t fl = 0;
      if (a.jb.group_off && a.jb.group_off[j + 1] > a.jb.group_off[j]) fl |= 1;
      kflags[slot] = fl;
    }
    n_out += __popc(kb);
  }
  if (lane == 0) *out_n = min(n_out, a.num_considerable);
}

// Parallel form of the queue-order pass for the common case of NO global pool
// quota (tools.clj:923 `(if (nil? quota) queue ...)`): the filters are then
// element-wise and "take N" is a stable compaction => three-kernel scan.
constexpr int SCAN_TB = 256;
constexpr int SCAN_ITEMS = 4;  // elements per thread

__device__ __forceinline__ bool cons_flag(const ConsArgs& a, const uint8_t* keep, int i) {
  if (i >= a.n_ranked || !keep[i]) return false;
  const int j = a.ranked[i];
  if (a.jb.allowed && !a.jb.allowed[j]) return false;
  if (a.jb.plugin && !a.jb.plugin[j]) return false;
  return true;
}

__global__ void __launch_bounds__(SCAN_TB) cons_count_kernel(ConsArgs a, const uint8_t* keep,
                                                             int32_t* block_sums) {
  __shared__ int warp_sums[SCAN_TB / 32];
  const int base = (blockIdx.x * SCAN_TB + threadIdx.x) * SCAN_ITEMS;
  int c = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) c += cons_flag(a, keep, base + q) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SCAN_TB / 32; w++) t += warp_sums[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void cons_scan_blocks_kernel(int32_t* block_sums, int nblocks, int32_t* out_n, int cap) {
  // single warp, sequential over chunks of 32 block sums (nblocks <= ~10k)
  const int lane = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < nblocks; base += 32) {
    int i = base + lane;
    int v = i < nblocks ? block_sums[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (i < nblocks) block_sums[i] = carry + incl - v;  // exclusive
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) *out_n = min(carry, cap);
}

__global__ void __launch_bounds__(SCAN_TB) cons_scatter_kernel(ConsArgs a, const uint8_t* keep,
                                                               const int32_t* block_off, int32_t* cons,
                                                               double* kc, double* km, double* kg,
                                                               int32_t* kports, uint8_t* kflags) {
  __shared__ int warp_off[SCAN_TB / 32];
  const int base = (blockIdx.x * SCAN_TB + threadIdx.x) * SCAN_ITEMS;
  bool f[SCAN_ITEMS];
  int c = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) { f[q] = cons_flag(a, keep, base + q); c += f[q] ? 1 : 0; }
  int incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_off[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SCAN_TB / 32; w++) { int x = warp_off[w]; warp_off[w] = t; t += x; }
  }
  __syncthreads();
  int slot = block_off[blockIdx.x] + warp_off[warp] + incl - c;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) {
    if (!f[q]) continue;
    if (slot < a.num_considerable) {
      const int j = a.ranked[base + q];
      cons[slot] = j;
      kc[slot] = a.jb.cpus[j];
      km[slot] = a.jb.mem[j];
      kg[slot] = a.jb.gpus ? a.jb.gpus[j] : 0.0;
      kports[slot] = a.jb.ports ? a.jb.ports[j] : 0;
      uint8_t fl = 0;
      if (a.jb.group_off && a.jb.group_off[j + 1] > a.jb.group_off[j]) fl |= 1;
      kflags[slot] = fl;
    }
    slot++;
  }
}

// ------------------------------------------------------------------ setup
__global__ void gather_offers_kernel(const int32_t* perm, int O, const double* c, const double* m,
                                     const double* rc, const double* rm, VmStatic* vs) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= O) return;
  int o = perm[v];
  VmStatic x;
  x.lc = c[o]; x.lm = m[o];
  x.rc = rc ? rc[o] : 0.0; x.rm = rm ? rm[o] : 0.0;
  vs[v] = x;
}

__global__ void ports_total_kernel(const int32_t* off, const int32_t* b, const int32_t* e, int O,
                                   int32_t* total) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  int t = 0;
  for (int k = off[o]; k < off[o + 1]; k++) t += e[k] - b[k] + 1;
  total[o] = t;
}

// assign (rank space) -> original offer index; assigned port numbers
// (FENZO F6: first n free ports scanning ranges in lease order).
__global__ void finalize_kernel(MatchArgs a, int32_t* out_assign, int32_t* out_ports, int max_ports,
                                int32_t* used_flag) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_cons) return;
  int v = a.assign[k];
  int o = v >= 0 ? a.of.perm[v] : -1;
  out_assign[k] = o;
  if (v >= 0) used_flag[v] = 1;
  if (out_ports && max_ports > 0) {
    for (int p = 0; p < max_ports; p++) out_ports[(size_t)k * max_ports + p] = -1;
    int want = (o >= 0 && a.jb.ports) ? a.jb.ports[a.cons[k]] : 0;
    if (want > 0 && a.of.port_off) {
      int skip = a.ports_start[k], got = 0;
      for (int r = a.of.port_off[o]; r < a.of.port_off[o + 1] && got < want; r++) {
        int len = a.of.port_end[r] - a.of.port_begin[r] + 1;
        if (skip >= len) { skip -= len; continue; }
        for (int p = a.of.port_begin[r] + skip; p <= a.of.port_end[r] && got < want; p++) {
          if (got < max_ports) out_ports[(size_t)k * max_ports + got] = p;
          got++;
        }
        skip = 0;
      }
    }
  }
}

__global__ void count_flags_kernel(const int32_t* flags, int n, int32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int v = (i < n && flags[i]) ? 1 : 0;
  unsigned b = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

}  // namespace

// ------------------------------------------------------------------ host side
// A MatchPlan is the device-resident image of one cook_match call's inputs plus
// all scratch.  With params->reuse_resident the upload stage is skipped (the
// inputs of the previous call on this handle are still in HBM) and only the
// kernels + result download run — the "inputs already resident" measurement.
struct MatchPlan {
  Arena arena;
  bool valid = false;
  int J = 0, O = 0, U = 0, n_ranked = 0, NC = 0, max_ports = 0, G = 0, B = 0;
  bool constr = false;
  size_t n_memb = 0;
  int64_t h2d_bytes = 0;
  ConsArgs ca;
  MatchArgs ma;
  // scratch
  int32_t *d_perm = nullptr, *d_pos = nullptr, *d_tmp = nullptr, *d_seg_s = nullptr, *d_seg_e = nullptr;
  uint8_t* d_keep = nullptr;
  double *d_oc = nullptr, *d_om = nullptr, *d_orc = nullptr, *d_orm = nullptr;
  VmStatic* d_vs = nullptr;
  double* d_kg = nullptr;
  int32_t* d_kports = nullptr;
  int32_t* d_ports_total = nullptr;
  int32_t *d_cons = nullptr, *d_out_assign = nullptr, *d_out_ports = nullptr, *d_used = nullptr;
  double *d_kc = nullptr, *d_km = nullptr;
  uint8_t* d_kflags = nullptr;
  unsigned long long* d_stats = nullptr;
  int32_t* d_counters = nullptr;
};

static void plan_free(void* p) {
  MatchPlan* mp = static_cast<MatchPlan*>(p);
  if (mp) { mp->arena.release(); delete mp; }
}

#define UP(dst, src, n)                                                   \
  do {                                                                    \
    CK(pool, upload(ar, st, (src), (size_t)(n), &(dst)));                 \
    if (src) mp->h2d_bytes += (int64_t)sizeof(*(src)) * (int64_t)(n);     \
  } while (0)

static int32_t build_plan(cook_pool* pool, MatchPlan* mp, const int32_t* ranked_idx,
                          int32_t n_ranked, const cook_jobs_soa* jobs,
                          const cook_offers_soa* offers, const cook_groups* groups,
                          const cook_user_table* users, const cook_pool_quota* pool_quota,
                          const cook_match_params* params, int32_t max_ports) {
  cudaStream_t st = pool->stream;
  Arena& ar = mp->arena;
  const int J = jobs->n, O = offers->n, U = users->n_users, NC = params->num_considerable;
  mp->valid = false;
  mp->h2d_bytes = 0;
  // host-side prep: offers sorted by hostname rank (tie-break order)
  std::vector<int32_t> perm(O);
  std::iota(perm.begin(), perm.end(), 0);
  std::sort(perm.begin(), perm.end(),
            [&](int32_t x, int32_t y) { return offers->name_rank[x] < offers->name_rank[y]; });
  // per-group capacity of the placed list = #member jobs
  const int G = groups ? groups->n_groups : 0;
  std::vector<int32_t> gp_off(G + 1, 0);
  size_t n_memb = 0;
  if (G && jobs->group_off) {
    n_memb = jobs->group_off[J];
    for (size_t i = 0; i < n_memb; i++) {
      int g = jobs->group_idx[i];
      if (g < 0 || g >= G) return set_err(pool, COOK_E_BADARG, "cook_match: bad group index");
      gp_off[g + 1]++;
    }
    for (int g = 0; g < G; g++) gp_off[g + 1] += gp_off[g];
  }
  // The cpu+mem-only kernel is used when no constraint column can affect a
  // placement (all-zero ports/gpus columns count as absent).
  bool constr_eff = jobs->novel_off || jobs->gpu_model || jobs->disk_request || jobs->attr_off ||
                    jobs->est_end_ms || jobs->ckpt_location || jobs->reserved_host ||
                    (jobs->group_off && G) || offers->is_k8s || offers->max_tasks ||
                    offers->reserved || offers->gpu_off;
  if (!constr_eff && jobs->gpus)
    for (int j = 0; j < J && !constr_eff; j++) constr_eff = jobs->gpus[j] != 0.0;
  if (!constr_eff && jobs->ports)
    for (int j = 0; j < J && !constr_eff; j++) constr_eff = jobs->ports[j] != 0;

  int B = 128;
  if (const char* eb = getenv("COOK_MATCH_B")) { int v = atoi(eb); if (v >= 8 && v <= MAXB) B = v; }
  Sizer sz;
  sz.add<int32_t>(n_ranked);
  for (int k = 0; k < 3; k++) sz.add<double>(J + 1);
  sz.add<int32_t>(J + 1); sz.add<int32_t>(J + 1);
  sz.add<uint8_t>(J + 1); sz.add<uint8_t>(J + 1);
  size_t csr_j = (jobs->novel_off ? jobs->novel_off[J] : 0) + 2 * (size_t)(jobs->attr_off ? jobs->attr_off[J] : 0) + n_memb;
  sz.add<int32_t>(4 * (size_t)(J + 2) + csr_j + 64);
  sz.add<double>(J + 1); sz.add<int32_t>(3 * (size_t)(J + 1)); sz.add<int64_t>(J + 1);
  sz.add<int32_t>(O + 1);
  for (int k = 0; k < 8; k++) sz.add<double>(O + 1);
  sz.add<int32_t>(12 * (size_t)(O + 2));
  size_t csr_o = (offers->port_off ? 2 * (size_t)offers->port_off[O] : 0) +
                 (offers->gpu_off ? (size_t)offers->gpu_off[O] : 0) +
                 (offers->disk_off ? (size_t)offers->disk_off[O] : 0);
  sz.add<int32_t>(csr_o + 64); sz.add<double>(csr_o + 64);
  sz.add<int64_t>(O + 1); sz.add<uint8_t>(2 * (size_t)(O + 1));
  sz.add<int32_t>((size_t)offers->n_attr_cols * O + 1);
  if (G) {
    sz.add<int32_t>(6 * (size_t)(G + 2));
    sz.add<int32_t>(2 * (size_t)(groups->cot_off ? groups->cot_off[G] : 0) + 64);
    sz.add<int32_t>(n_memb + 64);
  }
  for (int k = 0; k < 9; k++) sz.add<double>(U);
  sz.add<int32_t>(U);
  for (int k = 0; k < 4; k++) sz.add<double>(O + 1);
  for (int k = 0; k < 4; k++) sz.add<int32_t>(O + 1);
  for (int k = 0; k < 6; k++) sz.add<int32_t>(n_ranked + 1);
  sz.add<uint8_t>(n_ranked + 1);
  sz.add<int32_t>(NC + 1); sz.add<double>(NC + 1); sz.add<double>(NC + 1); sz.add<uint8_t>(NC + 1);
  sz.add<unsigned char>((size_t)2 * B * ROW_BYTES); sz.add<double>(NC + B + 1); sz.add<int32_t>(NC + B + 1);
  sz.add<uint8_t>(2 * B + 16); sz.add<unsigned>((size_t)NC / B + 16);
  sz.add<int32_t>(NC + 1); sz.add<int32_t>(NC + 1); sz.add<uint8_t>(NC + 1);
  sz.add<int32_t>(NC + 1); sz.add<int32_t>((size_t)NC * std::max(max_ports, 1) + 1);
  sz.add<int32_t>(O + 1);
  sz.add<unsigned long long>(32); sz.add<int32_t>(16);
  sz.add<VmStatic>(O + 1); sz.add<VmDyn>(O + 1); sz.add<VmDyn>(O + 1);
  CK(pool, ar.reserve(sz.off + (1 << 18)));
  ar.reset();

  ConsArgs& ca = mp->ca;
  memset(&ca, 0, sizeof(ca));
  int32_t* d_ranked; UP(d_ranked, ranked_idx, n_ranked);
  JobDev jb;
  memset(&jb, 0, sizeof(jb));
  { int32_t* p; UP(p, jobs->user, J); jb.user = p; }
  { double* p; UP(p, jobs->cpus, J); jb.cpus = p; UP(p, jobs->mem, J); jb.mem = p;
    UP(p, jobs->gpus, J); jb.gpus = p; }
  { int32_t* p; UP(p, jobs->ports, J); jb.ports = p; }
  { uint8_t* p; UP(p, jobs->allowed, J); jb.allowed = p; UP(p, jobs->plugin_accept, J); jb.plugin = p; }
  if (jobs->novel_off) { int32_t* p; UP(p, jobs->novel_off, J + 1); jb.novel_off = p;
    UP(p, jobs->novel_host, std::max(1, jobs->novel_off[J])); jb.novel_host = p; }
  { int32_t* p; UP(p, jobs->gpu_model, J); jb.gpu_model = p; }
  { double* p; UP(p, jobs->disk_request, J); jb.disk_request = p; }
  { int32_t* p; UP(p, jobs->disk_type, J); jb.disk_type = p; }
  if (jobs->attr_off) { int32_t* p; UP(p, jobs->attr_off, J + 1); jb.attr_off = p;
    int na = std::max(1, jobs->attr_off[J]);
    UP(p, jobs->attr_col, na); jb.attr_col = p; UP(p, jobs->attr_val, na); jb.attr_val = p; }
  { int64_t* p; UP(p, jobs->est_end_ms, J); jb.est_end_ms = p; }
  { int32_t* p; UP(p, jobs->ckpt_location, J); jb.ckpt_location = p;
    UP(p, jobs->reserved_host, J); jb.reserved_host = p; }
  if (jobs->group_off && G) { int32_t* p; UP(p, jobs->group_off, J + 1); jb.group_off = p;
    UP(p, jobs->group_idx, std::max<size_t>(1, n_memb)); jb.group_idx = p; }

  OfferDev of;
  memset(&of, 0, sizeof(of));
  of.O = O;
  { const int32_t* hp = perm.data(); UP(mp->d_perm, hp, O); of.perm = mp->d_perm; }
  UP(mp->d_oc, offers->cpus, O); UP(mp->d_om, offers->mem, O);
  UP(mp->d_orc, offers->run_cpus, O); UP(mp->d_orm, offers->run_mem, O);
  mp->d_vs = ar.take<VmStatic>(O + 1);
  of.vs = mp->d_vs;
  { int32_t* p; UP(p, offers->hostname_id, O); of.hostname_id = p;
    UP(p, offers->run_count, O); of.run_count = p; }
  mp->d_ports_total = nullptr;
  if (offers->port_off) { int32_t* p; UP(p, offers->port_off, O + 1); of.port_off = p;
    int np = std::max(1, offers->port_off[O]);
    UP(p, offers->port_begin, np); of.port_begin = p; UP(p, offers->port_end, np); of.port_end = p;
    mp->d_ports_total = ar.take<int32_t>(O + 1); of.ports_total = mp->d_ports_total; }
  { uint8_t* p; UP(p, offers->is_k8s, O); of.is_k8s = p; UP(p, offers->reserved, O); of.reserved = p; }
  { int32_t* p; UP(p, offers->location, O); of.location = p; }
  if (offers->gpu_off) { int32_t* p; UP(p, offers->gpu_off, O + 1); of.gpu_off = p;
    int ng = std::max(1, offers->gpu_off[O]); UP(p, offers->gpu_model, ng); of.gpu_model = p;
    double* q; UP(q, offers->gpu_count, ng); of.gpu_count = q; }
  if (offers->disk_off) { int32_t* p; UP(p, offers->disk_off, O + 1); of.disk_off = p;
    int nd = std::max(1, offers->disk_off[O]); UP(p, offers->disk_type, nd); of.disk_type = p;
    double* q; UP(q, offers->disk_space, nd); of.disk_space = q; }
  { int32_t* p; UP(p, offers->max_tasks, O); of.max_tasks = p; UP(p, offers->num_tasks, O); of.num_tasks = p; }
  { int64_t* p; UP(p, offers->host_start_time, O); of.host_start = p; }
  of.n_attr_cols = offers->attr ? offers->n_attr_cols : 0;
  if (of.n_attr_cols > 0) { int32_t* p; UP(p, offers->attr, (size_t)of.n_attr_cols * O); of.attr = p; }
  if ((jb.novel_off || of.reserved || (G && jb.group_off)) && !of.hostname_id)
    return set_err(pool, COOK_E_BADARG, "cook_match: hostname_id column required by constraints");

  GroupDev gr;
  memset(&gr, 0, sizeof(gr));
  if (G && jb.group_off) {
    gr.n_groups = G;
    int32_t* p;
    UP(p, groups->kind, G); gr.kind = p;
    UP(p, groups->attr_col, G); gr.attr_col = p;
    UP(p, groups->minimum, G); gr.minimum = p;
    UP(p, groups->cot_off, G + 1); gr.cot_off = p;
    int nc = std::max(1, groups->cot_off ? groups->cot_off[G] : 0);
    UP(p, groups->cot_hostname_id, nc); gr.cot_host = p;
    UP(p, groups->cot_attr_val, nc); gr.cot_attr = p;
    { const int32_t* hp = gp_off.data(); UP(p, hp, G + 1); gr.gp_off = p; }
    gr.gp_n = ar.take<int32_t>(G + 1);
    gr.gp_vm = ar.take<int32_t>(n_memb + 1);
    if (!gr.cot_off || !gr.kind) return set_err(pool, COOK_E_BADARG, "cook_match: incomplete cook_groups");
  }
  ca.ranked = d_ranked; ca.n_ranked = n_ranked; ca.jb = jb; ca.n_users = U;
  { double* p;
    UP(p, users->quota_count, U); ca.q_count = p; UP(p, users->quota_cpus, U); ca.q_cpus = p;
    UP(p, users->quota_mem, U); ca.q_mem = p; UP(p, users->quota_gpus, U); ca.q_gpus = p;
    UP(p, users->usage_count, U); ca.u_count = p; UP(p, users->usage_cpus, U); ca.u_cpus = p;
    UP(p, users->usage_mem, U); ca.u_mem = p; UP(p, users->usage_gpus, U); ca.u_gpus = p; }
  { int32_t* p; UP(p, users->tokens, U); ca.tokens = p; }
  ca.enforce_rate_limit = params->enforce_rate_limit;
  cook_pool_quota qoff{0, 0, 0, 0, 0};
  ca.pool_q = pool_quota ? *pool_quota : qoff;
  ca.num_considerable = NC;

  MatchArgs& ma = mp->ma;
  memset(&ma, 0, sizeof(ma));
  for (int b = 0; b < 2; b++) {
    ma.dyn.d[b] = ar.take<VmDyn>(O + 1);
  }
  mp->d_pos = ar.take<int32_t>(n_ranked + 1);
  mp->d_tmp = ar.take<int32_t>(n_ranked + 1);
  mp->d_seg_s = ar.take<int32_t>(U + 1);
  mp->d_seg_e = ar.take<int32_t>(U + 1);
  mp->d_keep = ar.take<uint8_t>(n_ranked + 1);
  mp->d_cons = ar.take<int32_t>(NC + 1);
  mp->d_kc = ar.take<double>(NC + 1);
  mp->d_km = ar.take<double>(NC + 1);
  mp->d_kflags = ar.take<uint8_t>(NC + 1);
  ma.rows = ar.take<unsigned char>((size_t)2 * B * ROW_BYTES);
  mp->d_kg = ar.take<double>(NC + B + 1);
  mp->d_kports = ar.take<int32_t>(NC + B + 1);
  ma.kg = mp->d_kg; ma.kports = mp->d_kports;
  ma.feas = ar.take<uint8_t>(2 * B + 16);
  ma.rows_ready = ar.take<unsigned>((size_t)NC / B + 16);
  ma.assign = ar.take<int32_t>(NC + 1);
  ma.ports_start = ar.take<int32_t>(NC + 1);
  ma.fail = ar.take<uint8_t>(NC + 1);
  mp->d_out_assign = ar.take<int32_t>(NC + 1);
  mp->d_out_ports = ar.take<int32_t>((size_t)NC * std::max(max_ports, 1) + 1);
  mp->d_used = ar.take<int32_t>(O + 1);
  mp->d_stats = ar.take<unsigned long long>(32);
  mp->d_counters = ar.take<int32_t>(16);
  if (!mp->d_counters) return set_err(pool, COOK_E_OOM, "cook_match: arena exhausted");
  ma.jb = jb; ma.of = of; ma.gr = gr;
  ma.cons = mp->d_cons; ma.kc = mp->d_kc; ma.km = mp->d_km; ma.kflags = mp->d_kflags;
  ma.B = B; ma.host_lifetime_mins = params->host_lifetime_mins;
  ma.published = reinterpret_cast<unsigned*>(mp->d_counters + 8); ma.stats = mp->d_stats;
  mp->J = J; mp->O = O; mp->U = U; mp->n_ranked = n_ranked; mp->NC = NC; mp->max_ports = max_ports;
  mp->G = G; mp->B = B; mp->constr = constr_eff; mp->n_memb = n_memb;
  mp->valid = true;
  return COOK_OK;
}

static int32_t run_plan(cook_pool* pool, MatchPlan* mp, int32_t* out_considerable,
                        int32_t* out_assign, int32_t* out_ports, uint8_t* out_fail_reason,
                        cook_match_stats* out_stats, bool uploaded) {
  cudaStream_t st = pool->stream;
  const int O = mp->O, U = mp->U, n_ranked = mp->n_ranked, max_ports = mp->max_ports;
  MatchArgs& ma = mp->ma;
  ConsArgs& ca = mp->ca;
  int launches = 0;
  // ---- reset of per-cycle dynamic state
  for (int b = 0; b < 2; b++) {
    CK(pool, cudaMemsetAsync(ma.dyn.d[b], 0, sizeof(VmDyn) * (O + 1), st));
  }
  if (ma.gr.gp_n) CK(pool, cudaMemsetAsync(ma.gr.gp_n, 0, sizeof(int32_t) * (mp->G + 1), st));
  CK(pool, cudaMemsetAsync(mp->d_seg_s, 0, sizeof(int32_t) * (U + 1), st));
  CK(pool, cudaMemsetAsync(mp->d_seg_e, 0, sizeof(int32_t) * (U + 1), st));
  CK(pool, cudaMemsetAsync(mp->d_used, 0, sizeof(int32_t) * (O + 1), st));
  CK(pool, cudaMemsetAsync(mp->d_stats, 0, sizeof(unsigned long long) * 32, st));
  CK(pool, cudaMemsetAsync(mp->d_counters, 0, sizeof(int32_t) * 16, st));
  CK(pool, cudaMemsetAsync(ma.rows_ready, 0, sizeof(unsigned) * ((size_t)mp->NC / mp->B + 16), st));
  CK(pool, cudaEventRecord(pool->ev[1], st));

  // ---- M0 considerable
  const int TB = 256;
  if (O > 0) {
    gather_offers_kernel<<<(O + TB - 1) / TB, TB, 0, st>>>(mp->d_perm, O, mp->d_oc, mp->d_om, mp->d_orc,
                                                           mp->d_orm, mp->d_vs);
    launches++;
    if (mp->d_ports_total) {
      ports_total_kernel<<<(O + TB - 1) / TB, TB, 0, st>>>(ma.of.port_off, ma.of.port_begin,
                                                           ma.of.port_end, O, mp->d_ports_total);
      launches++;
    }
  }
  iota_k<<<(n_ranked + TB - 1) / TB, TB, 0, st>>>(mp->d_pos, n_ranked);
  CK(pool, csort::sort_indices(mp->d_pos, mp->d_tmp, n_ranked, LessUserPos{ca.ranked, ca.jb.user}, st));
  launches += 2;
  for (long long w = csort::TILE; w < n_ranked; w <<= 1) launches++;
  cons_seg_kernel<<<(n_ranked + TB - 1) / TB, TB, 0, st>>>(mp->d_pos, ca.ranked, ca.jb.user, n_ranked,
                                                           mp->d_seg_s, mp->d_seg_e);
  cons_user_kernel<<<(U + 3) / 4, 128, 0, st>>>(ca, mp->d_pos, mp->d_seg_s, mp->d_seg_e, mp->d_keep);
  if (ca.pool_q.enabled) {
    // global pool quota: an order-dependent f64 left fold over the survivors
    // (filter-sequential) => exact single-warp pass
    cons_queue_kernel<<<1, 32, 0, st>>>(ca, mp->d_keep, mp->d_cons, mp->d_kc, mp->d_km, mp->d_kg,
                                        mp->d_kports, mp->d_kflags, mp->d_counters);
    launches += 3;
  } else {
    const int per_block = SCAN_TB * SCAN_ITEMS;
    const int nsb = (n_ranked + per_block - 1) / per_block;
    cons_count_kernel<<<nsb, SCAN_TB, 0, st>>>(ca, mp->d_keep, mp->d_tmp);
    cons_scan_blocks_kernel<<<1, 32, 0, st>>>(mp->d_tmp, nsb, mp->d_counters, ca.num_considerable);
    cons_scatter_kernel<<<nsb, SCAN_TB, 0, st>>>(ca, mp->d_keep, mp->d_tmp, mp->d_cons, mp->d_kc, mp->d_km,
                                                 mp->d_kg, mp->d_kports, mp->d_kflags);
    launches += 5;
  }
  CK(pool, cudaGetLastError());
  int32_t n_cons = 0;
  CK(pool, cudaMemcpyAsync(&n_cons, mp->d_counters, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(pool, cudaEventRecord(pool->ev[2], st));
  CK(pool, cudaStreamSynchronize(st));

  // ---- M3 matcher
  ma.n_cons = n_cons;
  int n_used = 0;
  unsigned long long hstats[32] = {0};
  const bool prof_on = getenv("COOK_PROF") != nullptr;
  CK(pool, cudaEventRecord(pool->ev[5], st));
  if (n_cons > 0) {
    if (O == 0) {
      CK(pool, cudaMemsetAsync(ma.assign, 0xff, sizeof(int32_t) * n_cons, st));
      CK(pool, cudaMemsetAsync(ma.fail, COOK_FAIL_NO_OFFERS, n_cons, st));
    } else {
      size_t smem = std::max(sizeof(ResolverShared) + sizeof(unsigned) * ((O + 31) / 32) + 16,
                             sizeof(EvalShared));
      if (smem > 220 * 1024)
        return set_err(pool, COOK_E_BADARG, "cook_match: too many offers for the resolver bitmap (%d)", O);
      void* kfn = mp->constr ? (prof_on ? (void*)match_kernel<true, true> : (void*)match_kernel<true, false>)
                             : (prof_on ? (void*)match_kernel<false, true> : (void*)match_kernel<false, false>);
      CK(pool, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int grid = pool->sm_count;
      int occ = 0;
      CK(pool, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, RES_THREADS, smem));
      if (occ < 1) return set_err(pool, COOK_E_CUDA, "cook_match: kernel does not fit on an SM");
      if (grid < 2) grid = 2;
      void* kargs[] = {&ma};
      CK(pool, cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(RES_THREADS), kargs, smem, st));
      launches++;
    }
  }
  CK(pool, cudaEventRecord(pool->ev[6], st));
  if (n_cons > 0) {
    finalize_kernel<<<(n_cons + TB - 1) / TB, TB, 0, st>>>(ma, mp->d_out_assign,
                                                           out_ports ? mp->d_out_ports : nullptr,
                                                           max_ports, mp->d_used);
    count_flags_kernel<<<(O + TB) / TB, TB, 0, st>>>(mp->d_used, O, mp->d_counters + 1);
    launches += 2;
    CK(pool, cudaGetLastError());
  }
  CK(pool, cudaEventRecord(pool->ev[3], st));
  if (n_cons > 0) {
    CK(pool, cudaMemcpyAsync(out_considerable, mp->d_cons, sizeof(int32_t) * n_cons, cudaMemcpyDeviceToHost, st));
    CK(pool, cudaMemcpyAsync(out_assign, mp->d_out_assign, sizeof(int32_t) * n_cons, cudaMemcpyDeviceToHost, st));
    if (out_ports && max_ports > 0)
      CK(pool, cudaMemcpyAsync(out_ports, mp->d_out_ports, sizeof(int32_t) * (size_t)n_cons * max_ports,
                               cudaMemcpyDeviceToHost, st));
    if (out_fail_reason)
      CK(pool, cudaMemcpyAsync(out_fail_reason, ma.fail, n_cons, cudaMemcpyDeviceToHost, st));
    CK(pool, cudaMemcpyAsync(hstats, mp->d_stats, sizeof(hstats), cudaMemcpyDeviceToHost, st));
    CK(pool, cudaMemcpyAsync(&n_used, mp->d_counters + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  CK(pool, cudaEventRecord(pool->ev[4], st));
  CK(pool, cudaStreamSynchronize(st));
  if (prof_on) {
    const char* nm[15] = {"driver", "spec_phase_w0", "chain_wait_w0", "apply", "rounds", "aborted_jobs",
                          "-", "-", "res_total", "-", "eval_work", "eval_wait",
                          "eval_loop", "eval_sync", "eval_merge"};
    for (int i = 0; i < 15; i++) fprintf(stderr, "[cook_prof] %-14s %llu\n", nm[i], hstats[4 + i]);
  }
  if (out_stats) {
    out_stats->n_considerable = n_cons;
    out_stats->n_matched = (int)hstats[3];
    out_stats->head_matched = (n_cons > 0 && out_assign[0] >= 0) ? 1 : 0;
    out_stats->n_offers_used = n_used;
    out_stats->evals = (int64_t)n_cons * O;
    out_stats->n_fast = (int64_t)hstats[0];
    out_stats->n_chunk_rescan = (int64_t)hstats[1];
    out_stats->n_full_rescan = (int64_t)hstats[2];
    out_stats->ms_h2d = uploaded ? ev_ms(pool->ev[0], pool->ev[1]) : 0.0;
    out_stats->ms_considerable = ev_ms(pool->ev[1], pool->ev[2]);
    out_stats->ms_match = ev_ms(pool->ev[2], pool->ev[3]);
    out_stats->ms_match_kernel = ev_ms(pool->ev[5], pool->ev[6]);
    out_stats->ms_d2h = ev_ms(pool->ev[3], pool->ev[4]);
    out_stats->n_launches = launches;
    out_stats->h2d_bytes = uploaded ? mp->h2d_bytes : 0;
    out_stats->d2h_bytes = (int64_t)n_cons * (8 + (out_fail_reason ? 1 : 0) +
                                               (out_ports ? 4 * (int64_t)max_ports : 0)) + 44;
  }
  return COOK_OK;
}

extern "C" int32_t cook_match(cook_pool* pool, const int32_t* ranked_idx, int32_t n_ranked,
                              const cook_jobs_soa* jobs, const cook_offers_soa* offers,
                              const cook_groups* groups, const cook_user_table* users,
                              const cook_pool_quota* pool_quota, const cook_match_params* params,
                              int32_t* out_considerable, int32_t* out_assign, int32_t* out_ports,
                              int32_t max_ports, uint8_t* out_fail_reason,
                              cook_match_stats* out_stats) {
  if (!pool) return COOK_E_BADARG;
  if (!ranked_idx || !jobs || !offers || !users || !params || !out_considerable || !out_assign)
    return set_err(pool, COOK_E_BADARG, "cook_match: null argument");
  if (params->good_enough_fitness < 1.0)
    return set_err(pool, COOK_E_BADARG,
                   "cook_match: good_enough_fitness < 1.0 is Fenzo's racy early-exit mode; "
                   "only the deterministic mode (>= 1.0) is supported");
  if (params->fitness_kind != 0)
    return set_err(pool, COOK_E_UNSUPPORTED_CONSTRAINT, "cook_match: only cpuMemBinPacker");
  const int J = jobs->n, O = offers->n, U = users->n_users;
  const int NC = params->num_considerable;
  if (J < 0 || O < 0 || U <= 0 || n_ranked < 0 || NC < 0 || max_ports < 0)
    return set_err(pool, COOK_E_BADARG, "cook_match: bad sizes");
  if (out_stats) memset(out_stats, 0, sizeof(*out_stats));
  if (n_ranked == 0 || NC == 0) return COOK_OK;
  CK(pool, cudaSetDevice(pool->device));
  if (!pool->match_plan) {
    pool->match_plan = new MatchPlan();
    pool->match_plan_free = plan_free;
  }
  MatchPlan* mp = static_cast<MatchPlan*>(pool->match_plan);
  const bool reuse = params->reuse_resident && mp->valid && mp->J == J && mp->O == O && mp->U == U &&
                     mp->n_ranked == n_ranked && mp->NC == NC && mp->max_ports == max_ports;
  if (params->reuse_resident && !reuse)
    return set_err(pool, COOK_E_BADARG,
                   "cook_match: reuse_resident set but no matching resident inputs on this handle");
  CK(pool, cudaEventRecord(pool->ev[0], pool->stream));
  if (!reuse) {
    int32_t rc = build_plan(pool, mp, ranked_idx, n_ranked, jobs, offers, groups, users, pool_quota,
                            params, max_ports);
    if (rc != COOK_OK) { mp->valid = false; return rc; }
  }
  return run_plan(pool, mp, out_considerable, out_assign, out_ports, out_fail_reason, out_stats, !reuse);
}
