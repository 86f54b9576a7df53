// rank.cu — DRU fair-share ranking on the GPU.
//
// Replaces sort-jobs-by-dru-helper (scheduler/scheduler.clj:2073-2091),
// limit-over-quota-jobs (:2057-2071), dru/compute-task-scored-task-pairs
// (dru.clj:50-80), dru/sorted-merge (dru.clj:82-104), filter-based-on-quota
// (scheduler.clj:2134-2157) and filter-offensive-jobs (:2198-2229).
//
// Pipeline (all on the pool's stream, no host round trips):
//   K1 iota + comparator sort by (user name rank, -priority, start, task id,
//      job id)                      -> per-user segments in tools.clj:614-641 order
//   K2 segment bounds
//   K3 warp-per-user fold           -> cumulative usage in the reference's
//      left-fold order (lane-serial shuffle chain keeps f64 association),
//      over-quota truncation, DRU = max(mem/div, cpus/div) (IEEE div.rn.f64)
//   K4 comparator sort of positions by (dru, k-way-merge tie rule)
//   K5 single-warp queue filter     -> pending only, pool quota, group quota,
//      offensive filter, stable compaction.
#include "common.cuh"
#include "sort.cuh"

namespace {

struct TaskCols {
  const int32_t* user;
  const int32_t* prio;
  const int64_t* start;
  const int64_t* tid;
  const int64_t* jid;
  const double* cpus;
  const double* mem;
  const double* gpus;
};

// tools.clj:614-641 compare of feature vectors, prefixed by the user's name
// rank so that one global sort yields all per-user lists, users in name order.
struct LessUserTask {
  TaskCols t;
  const int32_t* name_rank;
  __device__ bool operator()(int32_t a, int32_t b) const {
    int ua = name_rank[t.user[a]], ub = name_rank[t.user[b]];
    if (ua != ub) return ua < ub;
    int pa = -t.prio[a], pb = -t.prio[b];
    if (pa != pb) return pa < pb;
    long long sa = t.start[a], sb = t.start[b];
    if (sa != sb) return sa < sb;
    long long ta = t.tid[a], tb = t.tid[b];
    if (ta != tb) return ta < tb;
    long long ja = t.jid[a], jb = t.jid[b];
    if (ja != jb) return ja < jb;
    return a < b;
  }
};

// Global emission order of dru/sorted-merge (dru.clj:82-104).  X, Y are
// positions in the user-sorted array.  The merge emits, at every step, the head
// with the smallest (dru, -arrival, name) where arrival = step at which the
// user's previous task was emitted (0 for a user's first task): the popped
// user's remainder is consed to the FRONT before a STABLE re-sort (:93-94).
// Hence for equal dru:  X before Y  <=>  prev(X) was emitted AFTER prev(Y),
// which recurses on the previous tasks with the roles swapped.
struct LessMerge {
  const double* dru;        // by sorted position; NaN = cut by limit-over-quota
  const int32_t* user_at;   // user of sorted position
  const int32_t* seg_start; // per user
  const int32_t* name_rank;
  __device__ bool operator()(int32_t x, int32_t y) const {
    double dx = dru[x], dy = dru[y];
    bool kx = dx == dx, ky = dy == dy;
    if (kx != ky) return kx;  // truncated tasks sort last
    if (!kx) return x < y;
    while (true) {
      if (dx != dy) return dx < dy;
      int ux = user_at[x], uy = user_at[y];
      if (ux == uy) return x < y;  // same user: queue order
      bool hx = x > seg_start[ux], hy = y > seg_start[uy];
      if (!hx && !hy) return name_rank[ux] < name_rank[uy];
      if (!hx) return false;  // X has arrival 0 => Y first
      if (!hy) return true;
      // before(X,Y) = before(prev(Y), prev(X)): swap roles and step back
      int nx = y - 1, ny = x - 1;
      x = nx; y = ny;
      dx = dru[x]; dy = dru[y];
    }
  }
};

__global__ void scatter_dru_kernel(const int32_t* __restrict__ idx, const double* __restrict__ dru_at,
                                   int n, double* __restrict__ dru_task) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dru_task[idx[p]] = dru_at[p];
}

__global__ void iota_kernel(int32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void seg_bounds_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ user,
                                  int n, int32_t* seg_start, int32_t* seg_end,
                                  int32_t* user_at) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int u = user[idx[p]];
  user_at[p] = u;
  if (p == 0 || user[idx[p - 1]] != u) seg_start[u] = p;
  if (p == n - 1 || user[idx[p + 1]] != u) seg_end[u] = p + 1;
}

struct UserCols {
  const double *div_mem, *div_cpus, *div_gpus;
  const double *q_count, *q_cpus, *q_mem, *q_gpus;
};

// K3: one warp per user.  Lane-serial fold keeps the reference's left-fold
// association: acc = ((acc + x0) + x1) + ... exactly as `reductions` does.
__global__ void __launch_bounds__(128) user_fold_kernel(
    const int32_t* __restrict__ idx, TaskCols t, UserCols uc, const int32_t* __restrict__ seg_start,
    const int32_t* __restrict__ seg_end, int n_users, int dru_mode, int max_over_quota,
    double* __restrict__ dru_at, int32_t* n_kept_total) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_users) return;
  const int u = warp;
  const int s = seg_start[u], e = seg_end[u];
  if (e <= s) return;
  const double md = uc.div_mem[u], cd = uc.div_cpus[u], gd = uc.div_gpus[u];
  const double qn = uc.q_count[u], qc = uc.q_cpus[u], qm = uc.q_mem[u], qg = uc.q_gpus[u];
  double am = 0.0, ac = 0.0, ag = 0.0;  // carried cumulative sums (identical in all lanes)
  int over = 0;
  bool cut = false;
  int kept = 0;
  const double NaN = __longlong_as_double(0x7ff8000000000000LL);
  for (int base = s; base < e; base += 32) {
    int p = base + lane;
    double xm = 0.0, xc = 0.0, xg = 0.0;
    if (p < e) {
      int ti = idx[p];
      xm = t.mem[ti]; xc = t.cpus[ti]; xg = t.gpus[ti];
    }
    double mym = 0.0, myc = 0.0, myg = 0.0;
    int cntn = min(32, e - base);
#pragma unroll 4
    for (int l = 0; l < cntn; l++) {
      am = am + __shfl_sync(0xffffffffu, xm, l);
      ac = ac + __shfl_sync(0xffffffffu, xc, l);
      ag = ag + __shfl_sync(0xffffffffu, xg, l);
      if (lane == l) { mym = am; myc = ac; myg = ag; }
    }
    // scheduler.clj:2057-2071: keep while #violating prefixes <= limit
    bool viol = false;
    if (p < e) {
      double cnt = (double)(p - s + 1);
      viol = !(cnt <= qn && myc <= qc && mym <= qm && myg <= qg);
    }
    unsigned vb = __ballot_sync(0xffffffffu, viol);
    int over_incl = over + __popc(vb & (0xffffffffu >> (31 - lane)));
    bool keep = (p < e) && !cut && (over_incl <= max_over_quota);
    double d = NaN;
    if (keep) {
      if (dru_mode == 0) {
        double a = mym / md, b = myc / cd;
        d = a > b ? a : b;
      } else {
        d = myg / gd;
      }
    }
    if (p < e) dru_at[p] = d;
    unsigned kb = __ballot_sync(0xffffffffu, keep);
    kept += __popc(kb);
    over += __popc(vb);
    if (over > max_over_quota) cut = true;  // later batches are all beyond the cut
  }
  if (lane == 0 && kept) atomicAdd(n_kept_total, kept);
}

struct QueueFilterArgs {
  const int32_t* pos_sorted;  // positions (into user-sorted array) in merge order
  const int32_t* idx;         // sorted position -> combined task index
  int n_kept;
  int R;                      // running count (combined index >= R => pending)
  const double* cpus;         // combined columns
  const double* mem;
  const double* gpus;
  cook_pool_quota pool_q, group_q;
  double pool_usage[4];       // filled by pool_usage_kernel
  double group_usage[4];
  int filter_offensive;
  double off_mem, off_cpus;
  int32_t* out_order;         // combined task index per emitted task
  int32_t* out_ranked;        // pending indices surviving all filters
  int32_t* out_n;
};

// Σ running usage of the pool (scheduler.clj:2118-2123) in input order.
__global__ void pool_usage_kernel(const double* cpus, const double* mem, const double* gpus, int R,
                                  double* out4) {
  // single warp, lane-serial fold => left-fold association
  const int lane = threadIdx.x;
  double ac = 0.0, am = 0.0, ag = 0.0;
  for (int base = 0; base < R; base += 32) {
    int i = base + lane;
    double xc = i < R ? cpus[i] : 0.0, xm = i < R ? mem[i] : 0.0, xg = i < R ? gpus[i] : 0.0;
    int cntn = min(32, R - base);
    for (int l = 0; l < cntn; l++) {
      ac = ac + __shfl_sync(0xffffffffu, xc, l);
      am = am + __shfl_sync(0xffffffffu, xm, l);
      ag = ag + __shfl_sync(0xffffffffu, xg, l);
    }
  }
  if (lane == 0) { out4[0] = (double)R; out4[1] = ac; out4[2] = am; out4[3] = ag; }
}

// K5: single warp; stable, order-preserving sequence of filters
// (tools.clj:654-668 filter-sequential: state advances for rejected jobs too).
__global__ void queue_filter_kernel(QueueFilterArgs a, const double* pool_usage_dev) {
  const int lane = threadIdx.x;
  double pc = pool_usage_dev[1], pm = pool_usage_dev[2], pg = pool_usage_dev[3], pn = pool_usage_dev[0];
  double gc = a.group_usage[1], gm = a.group_usage[2], gg = a.group_usage[3], gn = a.group_usage[0];
  int n_out = 0;
  for (int base = 0; base < a.n_kept; base += 32) {
    int i = base + lane;
    bool valid = i < a.n_kept;
    int ti = -1;
    if (valid) {
      ti = a.idx[a.pos_sorted[i]];
      if (a.out_order) a.out_order[i] = ti;
    }
    bool pend = valid && ti >= a.R;
    double xc = 0.0, xm = 0.0, xg = 0.0;
    if (pend) { xc = a.cpus[ti]; xm = a.mem[ti]; xg = a.gpus[ti]; }
    bool keep = pend;
    if (a.pool_q.enabled) {
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      double mc = 0, mm = 0, mg = 0, mn = 0;
      while (mask) {
        int l = __ffs(mask) - 1;
        mask &= mask - 1;
        pn = pn + 1.0;
        pc = pc + __shfl_sync(0xffffffffu, xc, l);
        pm = pm + __shfl_sync(0xffffffffu, xm, l);
        pg = pg + __shfl_sync(0xffffffffu, xg, l);
        if (lane == l) { mc = pc; mm = pm; mg = pg; mn = pn; }
      }
      if (keep)
        keep = mn <= a.pool_q.count && mc <= a.pool_q.cpus && mm <= a.pool_q.mem && mg <= a.pool_q.gpus;
    }
    if (a.group_q.enabled) {
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      double mc = 0, mm = 0, mg = 0, mn = 0;
      while (mask) {
        int l = __ffs(mask) - 1;
        mask &= mask - 1;
        gn = gn + 1.0;
        gc = gc + __shfl_sync(0xffffffffu, xc, l);
        gm = gm + __shfl_sync(0xffffffffu, xm, l);
        gg = gg + __shfl_sync(0xffffffffu, xg, l);
        if (lane == l) { mc = gc; mm = gm; mg = gg; mn = gn; }
      }
      if (keep)
        keep = mn <= a.group_q.count && mc <= a.group_q.cpus && mm <= a.group_q.mem && mg <= a.group_q.gpus;
    }
    if (keep && a.filter_offensive && (xm > a.off_mem || xc > a.off_cpus)) keep = false;
    unsigned kb = __ballot_sync(0xffffffffu, keep);
    if (keep) a.out_ranked[n_out + __popc(kb & ((1u << lane) - 1u))] = ti - a.R;
    n_out += __popc(kb);
  }
  if (lane == 0) *a.out_n = n_out;
}

template <class T>
cudaError_t upload2(Arena& ar, cudaStream_t st, const T* a, int na, const T* b, int nb, T** out) {
  T* d = ar.take<T>((size_t)na + nb + 1);
  if (!d) return cudaErrorMemoryAllocation;
  *out = d;
  cudaError_t e = cudaSuccess;
  if (na) e = cudaMemcpyAsync(d, a, sizeof(T) * na, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  if (nb) e = cudaMemcpyAsync(d + na, b, sizeof(T) * nb, cudaMemcpyHostToDevice, st);
  return e;
}

}  // namespace

extern "C" int32_t cook_rank(cook_pool* pool, const cook_tasks_soa* running,
                             const cook_tasks_soa* pending, const cook_user_table* users,
                             const cook_pool_quota* pool_quota, const cook_pool_quota* group_quota,
                             const double* group_usage, const cook_rank_params* params,
                             int32_t* out_ranked_idx, int32_t* out_n, double* out_dru,
                             int32_t* out_order, int32_t* out_order_n) {
  if (!pool) return COOK_E_BADARG;
  if (!running || !pending || !users || !params || !out_ranked_idx || !out_n)
    return set_err(pool, COOK_E_BADARG, "cook_rank: null argument");
  const int R = running->n, J = pending->n, N = R + J, U = users->n_users;
  if (R < 0 || J < 0 || U <= 0) return set_err(pool, COOK_E_BADARG, "cook_rank: bad sizes");
  *out_n = 0;
  if (out_order_n) *out_order_n = 0;
  if (N == 0) return COOK_OK;
  CK(pool, cudaSetDevice(pool->device));
  cudaStream_t st = pool->stream;
  Arena& ar = pool->arena;
  Sizer sz;
  for (int k = 0; k < 2; k++) sz.add<int32_t>(N + 1);
  for (int k = 0; k < 3; k++) sz.add<int64_t>(N + 1);
  for (int k = 0; k < 3; k++) sz.add<double>(N + 1);
  sz.add<int32_t>(U);
  for (int k = 0; k < 7; k++) sz.add<double>(U);
  for (int k = 0; k < 6; k++) sz.add<int32_t>(N + 1);  // idx,tmp,user_at,pos,out_order,out_ranked
  sz.add<double>(N + 1);                                 // dru_at
  for (int k = 0; k < 2; k++) sz.add<int32_t>(U);        // seg bounds
  sz.add<double>(8);
  sz.add<int32_t>(8);
  sz.add<double>(N + 1);                                 // dru by task (output)
  CK(pool, ar.reserve(sz.off + 4096));
  ar.reset();

  TaskCols t;
  int32_t *d_user, *d_prio; int64_t *d_start, *d_tid, *d_jid; double *d_cpus, *d_mem, *d_gpus;
  CK(pool, upload2(ar, st, running->user, R, pending->user, J, &d_user));
  CK(pool, upload2(ar, st, running->priority, R, pending->priority, J, &d_prio));
  CK(pool, upload2(ar, st, running->start_time, R, pending->start_time, J, &d_start));
  CK(pool, upload2(ar, st, running->task_id, R, pending->task_id, J, &d_tid));
  CK(pool, upload2(ar, st, running->job_id, R, pending->job_id, J, &d_jid));
  CK(pool, upload2(ar, st, running->cpus, R, pending->cpus, J, &d_cpus));
  CK(pool, upload2(ar, st, running->mem, R, pending->mem, J, &d_mem));
  CK(pool, upload2(ar, st, running->gpus, R, pending->gpus, J, &d_gpus));
  t = TaskCols{d_user, d_prio, d_start, d_tid, d_jid, d_cpus, d_mem, d_gpus};
  int32_t* d_name_rank;
  UserCols uc;
  double *dm, *dc, *dg, *qn, *qc, *qm, *qg;
  CK(pool, upload(ar, st, users->name_rank, U, &d_name_rank));
  CK(pool, upload(ar, st, users->div_mem, U, &dm));
  CK(pool, upload(ar, st, users->div_cpus, U, &dc));
  CK(pool, upload(ar, st, users->div_gpus, U, &dg));
  CK(pool, upload(ar, st, users->quota_count, U, &qn));
  CK(pool, upload(ar, st, users->quota_cpus, U, &qc));
  CK(pool, upload(ar, st, users->quota_mem, U, &qm));
  CK(pool, upload(ar, st, users->quota_gpus, U, &qg));
  uc = UserCols{dm, dc, dg, qn, qc, qm, qg};

  int32_t* d_idx = ar.take<int32_t>(N + 1);
  int32_t* d_tmp = ar.take<int32_t>(N + 1);
  int32_t* d_user_at = ar.take<int32_t>(N + 1);
  int32_t* d_pos = ar.take<int32_t>(N + 1);
  int32_t* d_out_order = ar.take<int32_t>(N + 1);
  int32_t* d_out_ranked = ar.take<int32_t>(N + 1);
  double* d_dru_at = ar.take<double>(N + 1);
  int32_t* d_seg_start = ar.take<int32_t>(U);
  int32_t* d_seg_end = ar.take<int32_t>(U);
  double* d_pool_usage = ar.take<double>(8);
  int32_t* d_counters = ar.take<int32_t>(8);  // [0]=n_kept [1]=n_out
  double* d_dru_task = ar.take<double>(N + 1);
  if (!d_dru_task) return set_err(pool, COOK_E_OOM, "cook_rank: arena exhausted");

  CK(pool, cudaMemsetAsync(d_seg_start, 0, sizeof(int32_t) * U, st));
  CK(pool, cudaMemsetAsync(d_seg_end, 0, sizeof(int32_t) * U, st));
  CK(pool, cudaMemsetAsync(d_counters, 0, sizeof(int32_t) * 8, st));

  const int TB = 256, nb = (N + TB - 1) / TB;
  iota_kernel<<<nb, TB, 0, st>>>(d_idx, N);
  CK(pool, csort::sort_indices(d_idx, d_tmp, N, LessUserTask{t, d_name_rank}, st));
  seg_bounds_kernel<<<nb, TB, 0, st>>>(d_idx, d_user, N, d_seg_start, d_seg_end, d_user_at);
  {
    int warps_per_block = 4;
    int blocks = (U + warps_per_block - 1) / warps_per_block;
    user_fold_kernel<<<blocks, warps_per_block * 32, 0, st>>>(
        d_idx, t, uc, d_seg_start, d_seg_end, U, pool->dru_mode, params->max_over_quota_jobs,
        d_dru_at, d_counters);
  }
  iota_kernel<<<nb, TB, 0, st>>>(d_pos, N);
  CK(pool, csort::sort_indices(d_pos, d_tmp, N,
                               LessMerge{d_dru_at, d_user_at, d_seg_start, d_name_rank}, st));
  pool_usage_kernel<<<1, 32, 0, st>>>(d_cpus, d_mem, d_gpus, R, d_pool_usage);
  int32_t n_kept = 0;
  CK(pool, cudaMemcpyAsync(&n_kept, d_counters, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(pool, cudaStreamSynchronize(st));
  QueueFilterArgs qa;
  qa.pos_sorted = d_pos; qa.idx = d_idx; qa.n_kept = n_kept; qa.R = R;
  qa.cpus = d_cpus; qa.mem = d_mem; qa.gpus = d_gpus;
  cook_pool_quota off{0, 0, 0, 0, 0};
  qa.pool_q = pool_quota ? *pool_quota : off;
  qa.group_q = (group_quota && group_usage) ? *group_quota : off;
  for (int k = 0; k < 4; k++) qa.group_usage[k] = group_usage ? group_usage[k] : 0.0;
  qa.filter_offensive = params->filter_offensive;
  qa.off_mem = params->offensive_max_mem_mb; qa.off_cpus = params->offensive_max_cpus;
  qa.out_order = d_out_order; qa.out_ranked = d_out_ranked; qa.out_n = d_counters + 1;
  queue_filter_kernel<<<1, 32, 0, st>>>(qa, d_pool_usage);
  CK(pool, cudaGetLastError());
  int32_t n_out = 0;
  CK(pool, cudaMemcpyAsync(&n_out, d_counters + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(pool, cudaStreamSynchronize(st));
  if (n_out > 0)
    CK(pool, cudaMemcpyAsync(out_ranked_idx, d_out_ranked, sizeof(int32_t) * n_out,
                             cudaMemcpyDeviceToHost, st));
  if (out_order && n_kept > 0)
    CK(pool, cudaMemcpyAsync(out_order, d_out_order, sizeof(int32_t) * n_kept,
                             cudaMemcpyDeviceToHost, st));
  if (out_dru) {
    scatter_dru_kernel<<<nb, TB, 0, st>>>(d_idx, d_dru_at, N, d_dru_task);
    CK(pool, cudaMemcpyAsync(out_dru, d_dru_task, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
  }
  CK(pool, cudaStreamSynchronize(st));
  *out_n = n_out;
  if (out_order_n) *out_order_n = n_kept;
  return COOK_OK;
}
