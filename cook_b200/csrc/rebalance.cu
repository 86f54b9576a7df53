// rebalance.cu — rebalancer preemption-victim search on the GPU.
//
// Replaces init-state (rebalancer.clj:222-266), compute-pending-default-job-dru
// (:182-208), compute-preemption-decision (:320-407), next-state (:270-309) and
// the rebalance loop (:434-467).  Decisions are inherently sequential (H6); the
// work INSIDE a decision is data-parallel over the running tasks and hosts, and
// the state is kept incrementally between decisions:
//
//   once:
//     S1 comparator sort of the tasks by (user name, -priority, start, task id,
//        job id)                                  -> per-user order (tools.clj:614-641)
//     S2 warp-per-user left fold                  -> cumulative sums + DRU (dru.clj:50-66)
//     S3 tasks grouped by host (a task never changes host)
//   per pending job:
//     P1 warp fold over the job's user            -> job-below-quota, nearest dru
//     P2 host kernel                              -> constraints per host (:358-377)
//     P3 warp per host: [spare ; eligible victims by desc dru] prefix sums by
//        repeated warp selection of the next victim, best sufficient prefix (:380-403)
//     P4 argmax over hosts (max dru, ties -> greatest hostname = `max-key` last wins)
//     P5 apply: next-state -- victims die in place (a dead task adds 0.0 to every
//        fold), the job's task is inserted into its user's order, and only the
//        users that changed are re-folded, from the first position that changed
//        (dru.clj:128-144 next-task->scored-task).
//
// GPU DRU mode is rejected: the reference itself throws there (see oracle).
#include "common.cuh"
#include "sort.cuh"

namespace {

struct RTasks {  // capacity R + max_preemption; synthetic tasks appended
  int32_t* user; int32_t* prio; int64_t* start; int64_t* tid; int64_t* jid;
  double* cpus; double* mem; double* gpus; int32_t* host; uint8_t* alive; double* dru;
  int32_t* pos;        // position inside the user's sorted list (dead tasks keep their slot)
  double* cm; double* cc;  // cumulative mem / cpus of the user up to and including this slot
};

struct LessUser {  // keys only: dead tasks keep their place
  RTasks t;
  const int32_t* name_rank;
  __device__ bool operator()(int32_t a, int32_t b) const {
    int ua = name_rank[t.user[a]], ub = name_rank[t.user[b]];
    if (ua != ub) return ua < ub;
    int pa = -t.prio[a], pb = -t.prio[b];
    if (pa != pb) return pa < pb;
    if (t.start[a] != t.start[b]) return t.start[a] < t.start[b];
    if (t.tid[a] != t.tid[b]) return t.tid[a] < t.tid[b];
    if (t.jid[a] != t.jid[b]) return t.jid[a] < t.jid[b];
    return a < b;
  }
};

struct LessHost {
  RTasks t;
  __device__ bool operator()(int32_t a, int32_t b) const {
    int ha = t.host[a], hb = t.host[b];
    if (ha != hb) return ha < hb;
    return a < b;
  }
};

__global__ void iota_r(int32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void user_seg_kernel(const int32_t* ord, RTasks t, int n, int32_t* seg_start, int32_t* seg_end) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int u = t.user[ord[p]];
  if (p == 0 || t.user[ord[p - 1]] != u) seg_start[u] = p;
  if (p == n - 1 || t.user[ord[p + 1]] != u) seg_end[u] = p + 1;
}

// Users to re-fold after a decision, written by apply_kernel.
struct Refold {
  int32_t n;
  int32_t q_ins;       // where the new task went into the user order
  int32_t pu;          // its user
  int32_t user[64];    // victims of one decision sit on one host; more than 63 distinct
  int32_t from[64];    //   users fall back to `all` (from = segment start)
  int32_t all;
};

// dru.clj:50-66: lane-serial left fold (exact association) of the users in `rf`
// (or of every user when rf == nullptr), restarted at the first slot that changed.
__global__ void __launch_bounds__(128) user_dru_kernel(const int32_t* ord, RTasks t, const double* div_mem,
                                                       const double* div_cpus, const int32_t* seg_start,
                                                       const int32_t* seg_end, int n_users, const Refold* rf) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const bool all = rf == nullptr || rf->all;
  if (w >= (all ? n_users : rf->n)) return;
  const int u = all ? w : rf->user[w];
  const int s = seg_start[u], e = seg_end[u];
  if (e <= s) return;
  int f = s;
  if (!all) {
    f = rf->from[w];
    if (f == 0x7fffffff) f = rf->q_ins;           // only the insertion touches this user
    else if (f >= rf->q_ins) f++;                 // slots at and after the insertion moved by one
    if (u == rf->pu) f = min(f, rf->q_ins);
    f = min(max(f, s), e);
  }
  const double md = div_mem[u], cd = div_cpus[u];
  double am = 0.0, ac = 0.0;
  if (f > s) { int j = ord[f - 1]; am = t.cm[j]; ac = t.cc[j]; }
  for (int base = f; base < e; base += 32) {
    int p = base + lane;
    int i = p < e ? ord[p] : -1;
    const bool live = i >= 0 && t.alive[i];
    double xm = live ? t.mem[i] : 0.0, xc = live ? t.cpus[i] : 0.0;
    double mym = 0.0, myc = 0.0;
    int cntn = min(32, e - base);
    for (int l = 0; l < cntn; l++) {
      am = am + __shfl_sync(0xffffffffu, xm, l);
      ac = ac + __shfl_sync(0xffffffffu, xc, l);
      if (lane == l) { mym = am; myc = ac; }
    }
    if (i >= 0) {
      t.cm[i] = mym; t.cc[i] = myc;
      double a = mym / md, b = myc / cd;
      t.dru[i] = a > b ? a : b;
      t.pos[i] = p - s;
    }
  }
}

struct HostCols {
  int H;
  const int32_t* hostname_id; const int32_t* name_rank;
  uint8_t* has_spare; double* spare_cpus; double* spare_mem; double* spare_gpus;
  const uint8_t* is_k8s; const int32_t* location;
  const int32_t* gpu_off; const int32_t* gpu_model; const double* gpu_count;
  const int32_t* disk_off; const int32_t* disk_type; const double* disk_space;
  const int64_t* host_start; int n_attr_cols; const int32_t* attr;
};

struct PendCols {
  const int32_t* user; const double* cpus; const double* mem; const double* gpus;
  const int64_t* jid; const int32_t* prio;
  const int32_t* novel_off; const int32_t* novel_host; const int32_t* gpu_model;
  const double* disk_request; const int32_t* disk_type;
  const int32_t* attr_off; const int32_t* attr_col; const int32_t* attr_val;
  const int64_t* est_end_ms; const int32_t* ckpt_location;
  const int32_t* group_off; const int32_t* group_idx;
};

struct GroupCols {
  int n; const int32_t* kind; const int32_t* attr_col; const int32_t* minimum;
  const int32_t* cot_off; const int32_t* cot_host; const int32_t* cot_attr;
};

struct PendScalars {  // per pending job, device resident
  int below_quota;
  double pending_dru;
};

// P1: job-below-quota (:210-220) and pending dru (:182-208) for pending job p.
__global__ void pending_kernel(const int32_t* ord, RTasks t, PendCols pc, int p, const int32_t* seg_start,
                               const int32_t* seg_end, const double* q_count, const double* q_cpus,
                               const double* q_mem, const double* q_gpus, const double* div_mem,
                               const double* div_cpus, PendScalars* out) {
  const int lane = threadIdx.x;
  const int u = pc.user[p];
  const int s = seg_start[u], e = seg_end[u];
  const double pm = pc.mem[p], pcpu = pc.cpus[p], pg = pc.gpus ? pc.gpus[p] : 0.0;
  double an = 1.0, ac = pcpu, am = pm, ag = pg;  // (conj running-jobs p): p first
  double nearest = 0.0;
  const int pprio = -pc.prio[p];
  const long long pj = pc.jid[p];
  for (int base = s; base < e; base += 32) {
    int q = base + lane;
    int i = q < e ? ord[q] : -1;
    if (i >= 0 && !t.alive[i]) i = -1;   // preempted earlier in this cycle: adds 0.0, is no neighbour
    double xc = i >= 0 ? t.cpus[i] : 0.0, xm = i >= 0 ? t.mem[i] : 0.0, xg = i >= 0 ? t.gpus[i] : 0.0;
    double xn = i >= 0 ? 1.0 : 0.0;
    int cntn = min(32, e - base);
    for (int l = 0; l < cntn; l++) {
      an = an + __shfl_sync(0xffffffffu, xn, l);
      ac = ac + __shfl_sync(0xffffffffu, xc, l);
      am = am + __shfl_sync(0xffffffffu, xm, l);
      ag = ag + __shfl_sync(0xffffffffu, xg, l);
    }
    // task <= synthetic pending task [-prio, Long/MAX, nil(-1), job id] ?
    bool le = false;
    if (i >= 0) {
      int tp = -t.prio[i];
      if (tp != pprio) le = tp < pprio;
      else if (t.start[i] != 0x7fffffffffffffffLL) le = true;
      else if (t.tid[i] != -1) le = false;  // nil < any id
      else le = t.jid[i] <= pj;
    }
    unsigned m = __ballot_sync(0xffffffffu, le);
    if (m) {
      int last = 31 - __clz(m);
      nearest = __shfl_sync(0xffffffffu, i >= 0 ? t.dru[i] : 0.0, last);
    }
  }
  if (lane == 0) {
    out->below_quota = (an <= q_count[u] && ac <= q_cpus[u] && am <= q_mem[u] && ag <= q_gpus[u]) ? 1 : 0;
    double a = nearest + pm / div_mem[u], b = nearest + pcpu / div_cpus[u];
    out->pending_dru = a > b ? a : b;
  }
}

__device__ __forceinline__ double csr_get(const int32_t* off, const int32_t* key, const double* val, int o, int k) {
  if (!off) return 0.0;
  for (int i = off[o]; i < off[o + 1]; i++)
    if (key[i] == k) return val[i];
  return 0.0;
}

// host has at least one live task? (preemptable-host->slave-id, :371-377)
__global__ void host_has_task_kernel(RTasks t, int n, uint8_t* has_task) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && t.alive[i]) has_task[t.host[i]] = 1;
}

// P2: constraints of pending job p on every host (constraints.clj:504-515, :680-697)
__global__ void host_ok_kernel(PendCols pc, int p, HostCols hc, GroupCols gc, const uint8_t* has_task,
                               const int32_t* preempted_hosts, const int32_t* n_preempted,
                               int host_lifetime_mins, uint8_t* ok) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= hc.H) return;
  const bool have = has_task[h] != 0;
  bool pass = true;
  if (have && pc.novel_off)
    for (int k = pc.novel_off[p]; k < pc.novel_off[p + 1]; k++)
      if (pc.novel_host[k] == hc.hostname_id[h]) pass = false;
  const bool k8s = have && hc.is_k8s && hc.is_k8s[h];
  const double g = pc.gpus ? pc.gpus[p] : 0.0;
  if (k8s) {
    if (g > 0.0) {
      double hv = csr_get(hc.gpu_off, hc.gpu_model, hc.gpu_count, h, pc.gpu_model ? pc.gpu_model[p] : -1);
      if (!(hv == g)) pass = false;
    } else {
      int nm = hc.gpu_off ? hc.gpu_off[h + 1] - hc.gpu_off[h] : 0;
      if (nm != 0) pass = false;
    }
  } else if (!(g == 0.0)) {
    pass = false;
  }
  if (pc.disk_request && pc.disk_request[p] >= 0.0 && k8s) {
    double space = csr_get(hc.disk_off, hc.disk_type, hc.disk_space, h, pc.disk_type ? pc.disk_type[p] : -1);
    if (!(space >= pc.disk_request[p])) pass = false;
  }
  if (pc.attr_off)
    for (int k = pc.attr_off[p]; k < pc.attr_off[p + 1]; k++) {
      int col = pc.attr_col[k], val = pc.attr_val[k];
      if (!have || col < 0 || col >= hc.n_attr_cols) { pass = false; continue; }
      int hv = hc.attr[(size_t)col * hc.H + h];
      if (val <= 0 || hv != val) pass = false;
    }
  if (pc.est_end_ms && pc.est_end_ms[p] >= 0 && have && hc.host_start && hc.host_start[h] >= 0) {
    long long death = 1000LL * hc.host_start[h] + 60000LL * host_lifetime_mins;
    if (!(pc.est_end_ms[p] < death)) pass = false;
  }
  if (pc.ckpt_location && pc.ckpt_location[p] >= 0) {
    int loc = (have && hc.location) ? hc.location[h] : -1;
    if (loc != pc.ckpt_location[p]) pass = false;
  }
  if (pass && pc.group_off && gc.n > 0) {
    const int np = *n_preempted;
    for (int k = pc.group_off[p]; k < pc.group_off[p + 1] && pass; k++) {
      const int gi = pc.group_idx[k];
      const int kind = gc.kind[gi];
      const int col = gc.attr_col ? gc.attr_col[gi] : -1;
      const int c0 = gc.cot_off[gi], c1 = gc.cot_off[gi + 1];
      auto hattr = [&](int hh) { return (col >= 0 && col < hc.n_attr_cols) ? hc.attr[(size_t)col * hc.H + hh] : 0; };
      if (kind == COOK_GROUP_UNIQUE) {
        if (!have) { pass = false; break; }
        const int hn = hc.hostname_id[h];
        for (int q = 0; q < np; q++) if (hc.hostname_id[preempted_hosts[q]] == hn) pass = false;
        for (int c = c0; c < c1; c++) if (gc.cot_host[c] == hn) pass = false;
      } else {
        const int n = np + (c1 - c0);
        if (n == 0) continue;
        auto val_at = [&](int i) { return i < np ? hattr(preempted_hosts[i]) : gc.cot_attr[c0 + i - np]; };
        const int target = have ? hattr(h) : 0;
        int tf = 0;
        for (int i = 0; i < n; i++) tf += (val_at(i) == target);
        if (kind == COOK_GROUP_ATTR_EQUALS) { if (tf == 0) pass = false; }
        else if (tf != 0) {
          int mn = 0x7fffffff, mx = 0, distinct = 0;
          for (int i = 0; i < n; i++) {
            int vi = val_at(i), f = 0; bool first = true;
            for (int q = 0; q < n; q++) { int vq = val_at(q); if (vq == vi) { f++; if (q < i) first = false; } }
            if (first) { distinct++; mn = min(mn, f); mx = max(mx, f); }
          }
          if (gc.minimum[gi] > distinct) mn = 0;
          if (!(mn == mx || tf < mx)) pass = false;
        }
      }
    }
  }
  ok[h] = pass ? 1 : 0;
}

// host segments of the tasks grouped by host
__global__ void host_seg_kernel(const int32_t* hord, RTasks t, int n, int32_t* hs, int32_t* he) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int h = t.host[hord[p]];
  if (p == 0 || t.host[hord[p - 1]] != h) hs[h] = p;
  if (p == n - 1 || t.host[hord[p + 1]] != h) he[h] = p + 1;
}

struct HostBest {  // best sufficient prefix of one host
  double dru, mem, cpus, gpus;
  int32_t n_victims;  // -1: no candidate
};

// Priority-map order inside a host (:252-256, :349): (-dru, user name); equal
// (dru, user): later same-user position first (ours), then the lower index.
struct VKey {
  double dru; int32_t urank, pos, idx;   // idx < 0: none
};
__device__ __forceinline__ bool vkey_before(const VKey& a, const VKey& b) {
  if (b.idx < 0) return a.idx >= 0;
  if (a.idx < 0) return false;
  if (a.dru != b.dru) return a.dru > b.dru;
  if (a.urank != b.urank) return a.urank < b.urank;
  if (a.pos != b.pos) return a.pos > b.pos;
  return a.idx < b.idx;
}
__device__ __forceinline__ VKey vkey_shfl_xor(const VKey& k, int o) {
  VKey r;
  r.dru = __shfl_xor_sync(0xffffffffu, k.dru, o);
  r.urank = __shfl_xor_sync(0xffffffffu, k.urank, o);
  r.pos = __shfl_xor_sync(0xffffffffu, k.pos, o);
  r.idx = __shfl_xor_sync(0xffffffffu, k.idx, o);
  return r;
}

struct SelArgs {
  const int32_t* hord; const int32_t* hs; const int32_t* he;
  RTasks t; int R; int n_tasks;             // synthetic tasks are R .. n_tasks-1
  HostCols hc; PendCols pc; const int32_t* user_rank;
  const PendScalars* ps; double min_diff, safe;
};

// P3 for one host, one warp: [spare ; victims by desc dru] prefix sums in the
// reference's left-fold order.  The next victim is chosen by a warp-wide argmax
// over the eligible tasks that come after the previous one; the first sufficient
// prefix has the highest dru of the host, longer prefixes with the SAME dru win
// the max-key tie (last wins).  With `emit` the first n_emit victims are written
// in ascending dru order (:397 conj onto a list).
__device__ HostBest host_select(const SelArgs& a, int p, int h, int lane, int32_t* emit, int n_emit) {
  HostBest b;
  b.dru = 0.0; b.mem = b.cpus = b.gpus = 0.0; b.n_victims = -1;
  const double pm = a.pc.mem[p], pcpu = a.pc.cpus[p], pg = a.pc.gpus ? a.pc.gpus[p] : 0.0;
  const int pu = a.pc.user[p];
  const bool below = a.ps->below_quota != 0;
  const double pend = a.ps->pending_dru;
  const int s0 = a.hs[h], seg = a.he[h] - s0, n_items = seg + (a.n_tasks - a.R);
  auto item = [&](int k) -> int {
    int i = k < seg ? a.hord[s0 + k] : a.R + (k - seg);
    if (k >= seg && a.t.host[i] != h) return -1;
    if (!a.t.alive[i]) return -1;
    double d = a.t.dru[i];
    if (!(below || a.t.user[i] == pu)) return -1;
    if (d < a.safe) return -1;
    if (!((d - pend) > a.min_diff)) return -1;
    return i;
  };
  double sm = 0.0, sc = 0.0, sg = 0.0;
  int nv = 0;
  bool have = false;
  double cur = 0.0;
  auto consider = [&](double dru) {
    if (sm >= pm && sc >= pcpu && (pg > 0.0 ? sg >= pg : true)) {
      if (!have || dru >= cur) { have = true; cur = dru; b.dru = dru; b.mem = sm; b.cpus = sc; b.gpus = sg; b.n_victims = nv; }
    }
  };
  if (a.hc.has_spare[h]) {
    sg = sg + a.hc.spare_gpus[h]; sm = sm + a.hc.spare_mem[h]; sc = sc + a.hc.spare_cpus[h];
    consider(1.7976931348623157e308);
  }
  if (!emit) {  // cannot reach the request with everything eligible (any summation order, wide margin)?
    double tm = 0.0, tc = 0.0, tg = 0.0;
    for (int k = lane; k < n_items; k += 32) {
      int i = item(k);
      if (i >= 0) { tm += a.t.mem[i]; tc += a.t.cpus[i]; tg += a.t.gpus[i]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      tm += __shfl_xor_sync(0xffffffffu, tm, o); tc += __shfl_xor_sync(0xffffffffu, tc, o);
      tg += __shfl_xor_sync(0xffffffffu, tg, o);
    }
    const double slack = 1.0 + 1e-6;
    if (!have && ((sm + tm) * slack < pm || (sc + tc) * slack < pcpu || (pg > 0.0 && (sg + tg) * slack < pg))) return b;
  }
  VKey last;
  last.idx = -1; last.dru = 0.0; last.urank = 0; last.pos = 0;
  bool first = true;
  while (true) {
    if (emit && nv >= n_emit) break;
    VKey best;
    best.idx = -1; best.dru = 0.0; best.urank = 0; best.pos = 0;
    for (int k = lane; k < n_items; k += 32) {
      int i = item(k);
      if (i < 0) continue;
      VKey c;
      c.dru = a.t.dru[i]; c.urank = a.user_rank[a.t.user[i]]; c.pos = a.t.pos[i]; c.idx = i;
      if (!first && !vkey_before(last, c)) continue;   // already taken
      if (vkey_before(c, best)) best = c;
    }
    for (int o = 16; o > 0; o >>= 1) {
      VKey other = vkey_shfl_xor(best, o);
      if (vkey_before(other, best)) best = other;
    }
    if (best.idx < 0) break;
    if (have && best.dru < cur) break;   // later prefixes only have smaller dru
    const int i = best.idx;
    sg = sg + a.t.gpus[i]; sm = sm + a.t.mem[i]; sc = sc + a.t.cpus[i];
    if (emit && lane == 0) emit[n_emit - 1 - nv] = i;
    nv++;
    consider(best.dru);
    last = best;
    first = false;
  }
  return b;
}

__global__ void __launch_bounds__(256) host_best_kernel(SelArgs a, int p, const uint8_t* ok, HostBest* out) {
  const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (h >= a.hc.H) return;
  HostBest b;
  b.dru = 0.0; b.mem = b.cpus = b.gpus = 0.0; b.n_victims = -1;
  if (ok[h]) b = host_select(a, p, h, lane, nullptr, 0);
  if (lane == 0) out[h] = b;
}

struct ApplyArgs {
  SelArgs sel;
  const HostBest* best;
  const int32_t* uord; const int32_t* us;   // user order and segment starts (before the insertion)
  int32_t* n_tasks;        // live + dead task count (grows by one per decision)
  int32_t* n_dec; int32_t* n_vict; int32_t* preempted_hosts; int32_t* n_preempted;
  cook_decision* dec; int32_t* victims;
  uint8_t* has_task;
  int32_t* changed;        // out: 1 if a decision was made
  Refold* rf;
};

// P4 + P5: argmax over hosts (max dru; ties -> greatest hostname) and next-state.
__global__ void __launch_bounds__(256) apply_kernel(ApplyArgs a, int p) {
  __shared__ double s_dru[256];
  __shared__ int s_rank[256];
  __shared__ int s_host[256];
  const int tid = threadIdx.x;
  const HostCols& hc = a.sel.hc;
  const PendCols& pc = a.sel.pc;
  RTasks t = a.sel.t;
  double bd = -1.0; int br = -1, bh = -1;
  for (int h = tid; h < hc.H; h += blockDim.x) {
    if (a.best[h].n_victims < 0) continue;
    double d = a.best[h].dru; int r = hc.name_rank[h];
    if (d > bd || (d == bd && r > br)) { bd = d; br = r; bh = h; }
  }
  s_dru[tid] = bd; s_rank[tid] = br; s_host[tid] = bh;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (tid < s) {
      double d = s_dru[tid + s]; int r = s_rank[tid + s];
      if (d > s_dru[tid] || (d == s_dru[tid] && r > s_rank[tid])) { s_dru[tid] = d; s_rank[tid] = r; s_host[tid] = s_host[tid + s]; }
    }
    __syncthreads();
  }
  if (tid >= 32) return;
  const int h = s_host[0];
  if (h < 0) {
    if (tid == 0) *a.changed = 0;
    return;
  }
  const HostBest b = a.best[h];
  const int vb = *a.n_vict;
  // the victims again (same selection), stored in ascending dru order
  if (b.n_victims > 0) host_select(a.sel, p, h, tid, a.victims + vb, b.n_victims);
  __syncwarp();
  const int n = *a.n_tasks, ni = n;
  const int pu = pc.user[p];
  Refold& rf = *a.rf;
  if (tid == 0) {
    const int di = *a.n_dec;
    cook_decision d;
    d.pending_idx = p; d.host = h; d.dru = b.dru; d.mem = b.mem; d.cpus = b.cpus; d.gpus = b.gpus;
    d.victim_begin = vb; d.victim_count = b.n_victims;
    rf.n = 0; rf.all = 0; rf.pu = pu;
    bool pu_listed = false;
    for (int k = b.n_victims - 1; k >= 0; k--) {     // selection order
      const int i = a.victims[vb + k];
      t.alive[i] = 0;
      a.preempted_hosts[(*a.n_preempted)++] = t.host[i];
      const int u = t.user[i], from = a.us[u] + t.pos[i];
      int e = 0;
      while (e < rf.n && rf.user[e] != u) e++;
      if (e < rf.n) rf.from[e] = min(rf.from[e], from);
      else if (rf.n < 63) { rf.user[rf.n] = u; rf.from[rf.n] = from; rf.n++; }
      else rf.all = 1;
      if (u == pu) pu_listed = true;
    }
    if (!pu_listed) { rf.user[rf.n] = pu; rf.from[rf.n] = 0x7fffffff; rf.n++; }
    *a.n_vict = vb + b.n_victims;
    a.dec[di] = d;
    *a.n_dec = di + 1;
    // synthetic running task of the pending job on host h (create-task-ent :hostname)
    *a.n_tasks = n + 1;
    t.user[ni] = pu; t.prio[ni] = pc.prio[p]; t.start[ni] = 0x7fffffffffffffffLL;
    t.tid[ni] = -1; t.jid[ni] = pc.jid[p];
    t.cpus[ni] = pc.cpus[p]; t.mem[ni] = pc.mem[p]; t.gpus[ni] = pc.gpus ? pc.gpus[p] : 0.0;
    t.host[ni] = h; t.alive[ni] = 1; t.dru[ni] = 0.0; t.pos[ni] = 0; t.cm[ni] = 0.0; t.cc[ni] = 0.0;
    a.has_task[h] = 1;
    hc.has_spare[h] = 1;
    hc.spare_mem[h] = b.mem - pc.mem[p];
    hc.spare_gpus[h] = b.gpus - (pc.gpus ? pc.gpus[p] : 0.0);
    hc.spare_cpus[h] = b.cpus - pc.cpus[p];
  }
  __syncwarp();
  // where the new task goes in the user order: after every task that is not greater
  // (32-ary search: the predicate "new task < uord[q]" is monotone in q)
  LessUser less{t, a.sel.user_rank};
  int lo = 0, hi = n;
  while (lo < hi) {
    const int step = (hi - lo + 31) / 32;
    const int q = lo + tid * step;
    const bool pred = q < hi ? less(ni, a.uord[q]) : true;
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    const int L = m ? __ffs(m) - 1 : 32;
    const int nlo = L > 0 ? lo + (L - 1) * step + 1 : lo;
    const int nhi = L < 32 ? min(hi, lo + L * step) : hi;
    lo = min(nlo, nhi); hi = nhi;
  }
  if (tid == 0) {
    rf.q_ins = lo;
    *a.changed = 1;
  }
}

// uord with the new task inserted at rf->q_ins (out of place)
__global__ void insert_kernel(const int32_t* src, int32_t* dst, int n, int ni, const Refold* rf) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  const int q = rf->q_ins;
  dst[j] = j < q ? src[j] : (j == q ? ni : src[j - 1]);
}

}  // namespace

#define RUP(dst, src, n) CK(pool, upload(ar, st, (src), (size_t)(n), &(dst)))

extern "C" int32_t cook_rebalance(cook_pool* pool, const cook_running_soa* running,
                                  const cook_jobs_soa* pending, const int64_t* pending_job_id,
                                  const int32_t* pending_priority, const cook_host_table* hosts,
                                  const cook_groups* groups, const cook_user_table* users,
                                  const cook_rebalance_params* prm, cook_decision* out_decisions,
                                  int32_t* out_victims, int32_t* out_n) {
  if (!pool) return COOK_E_BADARG;
  if (!running || !pending || !pending_job_id || !pending_priority || !hosts || !users || !prm ||
      !out_decisions || !out_victims || !out_n)
    return set_err(pool, COOK_E_BADARG, "cook_rebalance: null argument");
  if (pool->dru_mode != 0)
    return set_err(pool, COOK_E_UNSUPPORTED_CONSTRAINT,
                   "cook_rebalance: GPU DRU mode has no reference behaviour (rebalancer.clj:339-349 throws)");
  const int R = running->t.n, P = pending->n, H = hosts->n, U = users->n_users;
  const int MP = prm->max_preemption;
  *out_n = 0;
  if (P <= 0 || MP <= 0 || H <= 0) return COOK_OK;
  CK(pool, cudaSetDevice(pool->device));
  cudaStream_t st = pool->stream;
  Arena& ar = pool->arena;
  const int CAP = R + MP + 1;
  const int G = groups ? groups->n_groups : 0;
  Sizer sz;
  for (int k = 0; k < 6; k++) sz.add<int64_t>(CAP);  // generous: covers int32/int64/double columns
  for (int k = 0; k < 10; k++) sz.add<double>(CAP);
  for (int k = 0; k < 8; k++) sz.add<int32_t>(CAP);
  sz.add<Refold>(2);
  for (int k = 0; k < 12; k++) sz.add<double>(std::max(U, H) + 1);
  for (int k = 0; k < 16; k++) sz.add<int32_t>(std::max(U, H) + 2);
  size_t csr_h = (hosts->gpu_off ? hosts->gpu_off[H] : 0) + (hosts->disk_off ? hosts->disk_off[H] : 0);
  sz.add<double>(csr_h + 64); sz.add<int32_t>(csr_h + 64);
  sz.add<int32_t>((size_t)hosts->n_attr_cols * H + 1);
  sz.add<int64_t>(H + 1);
  for (int k = 0; k < 8; k++) sz.add<double>(P + 1);
  for (int k = 0; k < 12; k++) sz.add<int32_t>(P + 2);
  size_t csr_p = (pending->novel_off ? pending->novel_off[P] : 0) + 2 * (size_t)(pending->attr_off ? pending->attr_off[P] : 0) +
                 (pending->group_off ? pending->group_off[P] : 0);
  sz.add<int32_t>(csr_p + 64); sz.add<int64_t>(P + 1);
  if (G) { sz.add<int32_t>(6 * (size_t)(G + 2)); sz.add<int32_t>(2 * (size_t)(groups->cot_off ? groups->cot_off[G] : 0) + 64); }
  sz.add<HostBest>(H + 1); sz.add<cook_decision>(MP + 1); sz.add<int32_t>(CAP + MP);
  sz.add<PendScalars>(4); sz.add<int32_t>(64);
  CK(pool, ar.reserve(sz.off + (1 << 16)));
  ar.reset();

  CK(pool, cudaEventRecord(pool->ev[12], st));
  RTasks t;
  t.user = ar.take<int32_t>(CAP); t.prio = ar.take<int32_t>(CAP); t.start = ar.take<int64_t>(CAP);
  t.tid = ar.take<int64_t>(CAP); t.jid = ar.take<int64_t>(CAP); t.cpus = ar.take<double>(CAP);
  t.mem = ar.take<double>(CAP); t.gpus = ar.take<double>(CAP); t.host = ar.take<int32_t>(CAP);
  t.alive = ar.take<uint8_t>(CAP); t.dru = ar.take<double>(CAP); t.pos = ar.take<int32_t>(CAP);
  t.cm = ar.take<double>(CAP); t.cc = ar.take<double>(CAP);
  if (!t.cc) return set_err(pool, COOK_E_OOM, "cook_rebalance: arena exhausted");
  const cook_tasks_soa& rt = running->t;
#define CPY(dst, src, T) if (R) CK(pool, cudaMemcpyAsync(dst, src, sizeof(T) * R, cudaMemcpyHostToDevice, st))
  CPY(t.user, rt.user, int32_t); CPY(t.prio, rt.priority, int32_t); CPY(t.start, rt.start_time, int64_t);
  CPY(t.tid, rt.task_id, int64_t); CPY(t.jid, rt.job_id, int64_t); CPY(t.cpus, rt.cpus, double);
  CPY(t.mem, rt.mem, double); CPY(t.host, running->host, int32_t);
  if (rt.gpus) { CPY(t.gpus, rt.gpus, double); } else CK(pool, cudaMemsetAsync(t.gpus, 0, sizeof(double) * CAP, st));
#undef CPY
  CK(pool, cudaMemsetAsync(t.alive, 0, CAP, st));
  if (R) CK(pool, cudaMemsetAsync(t.alive, 1, R, st));
  CK(pool, cudaMemsetAsync(t.dru, 0, sizeof(double) * CAP, st));

  int32_t* d_urank; double *d_divm, *d_divc, *d_qn, *d_qc, *d_qm, *d_qg;
  RUP(d_urank, users->name_rank, U); RUP(d_divm, users->div_mem, U); RUP(d_divc, users->div_cpus, U);
  RUP(d_qn, users->quota_count, U); RUP(d_qc, users->quota_cpus, U); RUP(d_qm, users->quota_mem, U);
  RUP(d_qg, users->quota_gpus, U);

  HostCols hc;
  memset(&hc, 0, sizeof(hc));
  hc.H = H;
  { int32_t* p; RUP(p, hosts->hostname_id, H); hc.hostname_id = p; RUP(p, hosts->name_rank, H); hc.name_rank = p; }
  hc.has_spare = ar.take<uint8_t>(H + 1); hc.spare_cpus = ar.take<double>(H + 1);
  hc.spare_mem = ar.take<double>(H + 1); hc.spare_gpus = ar.take<double>(H + 1);
  CK(pool, cudaMemsetAsync(hc.has_spare, 0, H + 1, st));
  CK(pool, cudaMemsetAsync(hc.spare_cpus, 0, sizeof(double) * (H + 1), st));
  CK(pool, cudaMemsetAsync(hc.spare_mem, 0, sizeof(double) * (H + 1), st));
  CK(pool, cudaMemsetAsync(hc.spare_gpus, 0, sizeof(double) * (H + 1), st));
  if (hosts->has_spare) CK(pool, cudaMemcpyAsync(hc.has_spare, hosts->has_spare, H, cudaMemcpyHostToDevice, st));
  if (hosts->spare_cpus) CK(pool, cudaMemcpyAsync(hc.spare_cpus, hosts->spare_cpus, sizeof(double) * H, cudaMemcpyHostToDevice, st));
  if (hosts->spare_mem) CK(pool, cudaMemcpyAsync(hc.spare_mem, hosts->spare_mem, sizeof(double) * H, cudaMemcpyHostToDevice, st));
  if (hosts->spare_gpus) CK(pool, cudaMemcpyAsync(hc.spare_gpus, hosts->spare_gpus, sizeof(double) * H, cudaMemcpyHostToDevice, st));
  { uint8_t* p; RUP(p, hosts->is_k8s, H); hc.is_k8s = p; }
  { int32_t* p; RUP(p, hosts->location, H); hc.location = p; }
  if (hosts->gpu_off) { int32_t* p; RUP(p, hosts->gpu_off, H + 1); hc.gpu_off = p;
    int n = std::max(1, hosts->gpu_off[H]); RUP(p, hosts->gpu_model, n); hc.gpu_model = p;
    double* q; RUP(q, hosts->gpu_count, n); hc.gpu_count = q; }
  if (hosts->disk_off) { int32_t* p; RUP(p, hosts->disk_off, H + 1); hc.disk_off = p;
    int n = std::max(1, hosts->disk_off[H]); RUP(p, hosts->disk_type, n); hc.disk_type = p;
    double* q; RUP(q, hosts->disk_space, n); hc.disk_space = q; }
  { int64_t* p; RUP(p, hosts->host_start_time, H); hc.host_start = p; }
  hc.n_attr_cols = hosts->attr ? hosts->n_attr_cols : 0;
  if (hc.n_attr_cols > 0) { int32_t* p; RUP(p, hosts->attr, (size_t)hc.n_attr_cols * H); hc.attr = p; }

  PendCols pc;
  memset(&pc, 0, sizeof(pc));
  { int32_t* p; RUP(p, pending->user, P); pc.user = p; RUP(p, pending_priority, P); pc.prio = p; }
  { double* p; RUP(p, pending->cpus, P); pc.cpus = p; RUP(p, pending->mem, P); pc.mem = p; RUP(p, pending->gpus, P); pc.gpus = p; }
  { int64_t* p; RUP(p, pending_job_id, P); pc.jid = p; RUP(p, pending->est_end_ms, P); pc.est_end_ms = p; }
  if (pending->novel_off) { int32_t* p; RUP(p, pending->novel_off, P + 1); pc.novel_off = p;
    RUP(p, pending->novel_host, std::max(1, pending->novel_off[P])); pc.novel_host = p; }
  { int32_t* p; RUP(p, pending->gpu_model, P); pc.gpu_model = p; RUP(p, pending->disk_type, P); pc.disk_type = p;
    RUP(p, pending->ckpt_location, P); pc.ckpt_location = p; }
  { double* p; RUP(p, pending->disk_request, P); pc.disk_request = p; }
  if (pending->attr_off) { int32_t* p; RUP(p, pending->attr_off, P + 1); pc.attr_off = p;
    int n = std::max(1, pending->attr_off[P]); RUP(p, pending->attr_col, n); pc.attr_col = p;
    RUP(p, pending->attr_val, n); pc.attr_val = p; }
  GroupCols gc;
  memset(&gc, 0, sizeof(gc));
  if (G && pending->group_off) {
    int32_t* p; RUP(p, pending->group_off, P + 1); pc.group_off = p;
    RUP(p, pending->group_idx, std::max(1, pending->group_off[P])); pc.group_idx = p;
    gc.n = G;
    RUP(p, groups->kind, G); gc.kind = p; RUP(p, groups->attr_col, G); gc.attr_col = p;
    RUP(p, groups->minimum, G); gc.minimum = p; RUP(p, groups->cot_off, G + 1); gc.cot_off = p;
    int n = std::max(1, groups->cot_off ? groups->cot_off[G] : 0);
    RUP(p, groups->cot_hostname_id, n); gc.cot_host = p; RUP(p, groups->cot_attr_val, n); gc.cot_attr = p;
    if (!gc.cot_off || !gc.kind) return set_err(pool, COOK_E_BADARG, "cook_rebalance: incomplete cook_groups");
  }

  int32_t* d_ord = ar.take<int32_t>(CAP); int32_t* d_hord = ar.take<int32_t>(CAP);
  int32_t* d_tmp = ar.take<int32_t>(CAP);
  int32_t* d_us = ar.take<int32_t>(U + 1); int32_t* d_ue = ar.take<int32_t>(U + 1);
  int32_t* d_hs = ar.take<int32_t>(H + 1); int32_t* d_he = ar.take<int32_t>(H + 1);
  uint8_t* d_has_task = ar.take<uint8_t>(H + 1); uint8_t* d_ok = ar.take<uint8_t>(H + 1);
  HostBest* d_best = ar.take<HostBest>(H + 1);
  cook_decision* d_dec = ar.take<cook_decision>(MP + 1);
  int32_t* d_vict = ar.take<int32_t>(CAP + MP);
  int32_t* d_pre = ar.take<int32_t>(CAP + MP);
  PendScalars* d_ps = ar.take<PendScalars>(4);
  Refold* d_rf = ar.take<Refold>(1);
  int32_t* d_cnt = ar.take<int32_t>(64);  // [0] n_tasks [1] n_dec [2] n_vict [3] n_preempted [4] changed
  if (!d_cnt) return set_err(pool, COOK_E_OOM, "cook_rebalance: arena exhausted");
  int32_t h_cnt[8] = {R, 0, 0, 0, 0, 0, 0, 0};
  CK(pool, cudaMemcpyAsync(d_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, st));

  CK(pool, cudaEventRecord(pool->ev[13], st));
  const int TB = 256;
  // ---- init-state: user order + DRU of every user, tasks grouped by host
  CK(pool, cudaMemsetAsync(d_us, 0, sizeof(int32_t) * (U + 1), st));
  CK(pool, cudaMemsetAsync(d_ue, 0, sizeof(int32_t) * (U + 1), st));
  CK(pool, cudaMemsetAsync(d_hs, 0, sizeof(int32_t) * (H + 1), st));
  CK(pool, cudaMemsetAsync(d_he, 0, sizeof(int32_t) * (H + 1), st));
  CK(pool, cudaMemsetAsync(d_has_task, 0, H + 1, st));
  if (R > 0) {
    iota_r<<<(R + TB - 1) / TB, TB, 0, st>>>(d_ord, R);
    CK(pool, csort::sort_indices(d_ord, d_tmp, R, LessUser{t, d_urank}, st));
    user_seg_kernel<<<(R + TB - 1) / TB, TB, 0, st>>>(d_ord, t, R, d_us, d_ue);
    user_dru_kernel<<<(U + 3) / 4, 128, 0, st>>>(d_ord, t, d_divm, d_divc, d_us, d_ue, U, nullptr);
    iota_r<<<(R + TB - 1) / TB, TB, 0, st>>>(d_hord, R);
    CK(pool, csort::sort_indices(d_hord, d_tmp, R, LessHost{t}, st));
    host_seg_kernel<<<(R + TB - 1) / TB, TB, 0, st>>>(d_hord, t, R, d_hs, d_he);
    host_has_task_kernel<<<(R + TB - 1) / TB, TB, 0, st>>>(t, R, d_has_task);
  }
  SelArgs sa;
  sa.hord = d_hord; sa.hs = d_hs; sa.he = d_he; sa.t = t; sa.R = R; sa.n_tasks = R;
  sa.hc = hc; sa.pc = pc; sa.user_rank = d_urank; sa.ps = d_ps;
  sa.min_diff = prm->min_dru_diff; sa.safe = prm->safe_dru_threshold;
  int n_tasks = R, n_dec = 0;
  for (int p = 0; p < P && n_dec < MP; p++) {
    pending_kernel<<<1, 32, 0, st>>>(d_ord, t, pc, p, d_us, d_ue, d_qn, d_qc, d_qm, d_qg, d_divm, d_divc, d_ps);
    host_ok_kernel<<<(H + TB - 1) / TB, TB, 0, st>>>(pc, p, hc, gc, d_has_task, d_pre, d_cnt + 3,
                                                     prm->host_lifetime_mins, d_ok);
    sa.n_tasks = n_tasks;
    host_best_kernel<<<(H + 7) / 8, 256, 0, st>>>(sa, p, d_ok, d_best);
    ApplyArgs aa;
    aa.sel = sa; aa.best = d_best; aa.uord = d_ord; aa.us = d_us;
    aa.n_tasks = d_cnt; aa.n_dec = d_cnt + 1; aa.n_vict = d_cnt + 2; aa.preempted_hosts = d_pre;
    aa.n_preempted = d_cnt + 3; aa.dec = d_dec; aa.victims = d_vict; aa.has_task = d_has_task;
    aa.changed = d_cnt + 4; aa.rf = d_rf;
    apply_kernel<<<1, 256, 0, st>>>(aa, p);
    CK(pool, cudaGetLastError());
    CK(pool, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int32_t) * 5, cudaMemcpyDeviceToHost, st));
    CK(pool, cudaStreamSynchronize(st));
    n_dec = h_cnt[1];
    if (h_cnt[4] != 0) {  // next-state: the new task enters its user's order, changed users are re-folded
      const int n = n_tasks, ni = n_tasks;
      insert_kernel<<<(n + 1 + TB - 1) / TB, TB, 0, st>>>(d_ord, d_tmp, n, ni, d_rf);
      CK(pool, cudaMemcpyAsync(d_ord, d_tmp, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToDevice, st));
      CK(pool, cudaMemsetAsync(d_us, 0, sizeof(int32_t) * (U + 1), st));
      CK(pool, cudaMemsetAsync(d_ue, 0, sizeof(int32_t) * (U + 1), st));
      user_seg_kernel<<<(n + 1 + TB - 1) / TB, TB, 0, st>>>(d_ord, t, n + 1, d_us, d_ue);
      user_dru_kernel<<<(U + 3) / 4, 128, 0, st>>>(d_ord, t, d_divm, d_divc, d_us, d_ue, U, d_rf);
    }
    n_tasks = h_cnt[0];
  }
  CK(pool, cudaEventRecord(pool->ev[14], st));
  if (n_dec > 0) {
    CK(pool, cudaMemcpyAsync(out_decisions, d_dec, sizeof(cook_decision) * n_dec, cudaMemcpyDeviceToHost, st));
    if (h_cnt[2] > 0)
      CK(pool, cudaMemcpyAsync(out_victims, d_vict, sizeof(int32_t) * h_cnt[2], cudaMemcpyDeviceToHost, st));
    CK(pool, cudaStreamSynchronize(st));
  }
  CK(pool, cudaEventRecord(pool->ev[15], st));
  CK(pool, cudaStreamSynchronize(st));
  {
    cook_phase_stats& ps = pool->phase[COOK_PHASE_REBALANCE];
    ps.ms_h2d = ev_ms(pool->ev[12], pool->ev[13]);
    ps.ms_device = ev_ms(pool->ev[13], pool->ev[14]);
    ps.ms_d2h = ev_ms(pool->ev[14], pool->ev[15]);
    ps.h2d_bytes = (int64_t)R * 60 + (int64_t)P * 48 + (int64_t)H * 48 + (int64_t)U * 60;
    ps.d2h_bytes = (int64_t)n_dec * (int64_t)sizeof(cook_decision) + (int64_t)h_cnt[2] * 4;
    ps.n_launches = 0;
  }
  *out_n = n_dec;
  return COOK_OK;
}
