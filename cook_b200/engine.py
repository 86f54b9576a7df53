"""Thin call layer over the C ABI (include/cook_gpu.h).

`GpuEngine` drives libcookgpu.so (the product).  The same call shapes are
reused by oracle/pyoracle.py for the CPU oracle so that tests can diff the two;
this module itself never touches the oracle and has no CPU fallback: if the CUDA
library is missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcookgpu.so")


class CookError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"cook_gpu error {code}: {msg}")
        self.code = code


def _empty_tasks():
    z32 = np.zeros(1, np.int32)
    z64 = np.zeros(1, np.int64)
    zf = np.zeros(1, np.float64)
    t = abi.TasksSoA(n=0, user=z32, priority=z32, start_time=z64, task_id=z64, job_id=z64,
                     cpus=zf, mem=zf, gpus=zf)
    return t


class _CallShapes:
    """Argument packing shared by the GPU engine and the oracle wrapper."""

    @staticmethod
    def rank_buffers(running, pending):
        n = running.n + pending.n
        return (np.zeros(max(pending.n, 1), np.int32), C.c_int32(0),
                np.full(max(n, 1), np.nan, np.float64), np.zeros(max(n, 1), np.int32), C.c_int32(0))

    @staticmethod
    def match_buffers(params, max_ports):
        k = max(params.num_considerable, 1)
        return (np.full(k, -1, np.int32), np.full(k, -1, np.int32),
                np.full(k * max(max_ports, 1), -1, np.int32), np.zeros(k, np.uint8),
                abi.MatchStats())


def load_library(path=None):
    # COOK_GPU_LIB: A/B builds of the same library (tools/ only); never a CPU fallback
    path = path or os.environ.get("COOK_GPU_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.cook_gpu_version.restype = C.c_char_p
    for name in ("cook_gpu_init", "cook_gpu_shutdown", "cook_pool_open", "cook_pool_close",
                 "cook_last_error", "cook_rank", "cook_match", "cook_rebalance",
                 "cook_allgather_usage", "cook_last_stats", "cook_comm_unique_id", "cook_comm_init",
                 "cook_comm_destroy", "cook_exchange_usage", "cook_exchange_usage_batch", "cook_rebalance_trace",
                 "cook_match_failures"):
        getattr(lib, name).restype = C.c_int32
    return lib


class GpuEngine:
    """One Cook pool bound to one GPU.  Not re-entrant (same contract as
    `(locking fenzo ...)`, scheduler/scheduler.clj:665)."""

    def __init__(self, pool_name="no-pool", dru_mode=0, device=0, lib=None):
        self.lib = lib or load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.cook_gpu_init(None, C.byref(self.ctx))
        if rc != 0:
            raise CookError(rc, "cook_gpu_init failed (no CUDA device? there is no CPU fallback)")
        self.pool = C.c_void_p()
        rc = self.lib.cook_pool_open(self.ctx, pool_name.encode(), int(dru_mode), int(device),
                                     C.byref(self.pool))
        if rc != 0:
            raise CookError(rc, "cook_pool_open failed")
        self.dru_mode = dru_mode

    def close(self):
        if getattr(self, "pool", None):
            self.lib.cook_pool_close(self.pool)
            self.pool = None
        if getattr(self, "ctx", None):
            self.lib.cook_gpu_shutdown(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        buf = C.create_string_buffer(512)
        self.lib.cook_last_error(self.pool, buf, 512)
        raise CookError(rc, buf.value.decode(errors="replace"))

    # -- R1-R7 -------------------------------------------------------------
    def rank(self, running, pending, users, pool_quota=None, group_quota=None,
             group_usage=None, params=None):
        running = running or _empty_tasks()
        pool_quota = pool_quota or abi.make_pool_quota(None)
        group_quota = group_quota or abi.make_pool_quota(None)
        params = params or abi.RankParams(100, 0, 0.0, 0.0)
        gu = np.ascontiguousarray(group_usage if group_usage is not None else np.zeros(4), np.float64)
        ranked, n, dru, order, on = _CallShapes.rank_buffers(running, pending)
        rc = self.lib.cook_rank(self.pool, C.byref(running), C.byref(pending), C.byref(users),
                                C.byref(pool_quota), C.byref(group_quota), abi.ptr(gu, abi.P_F64),
                                C.byref(params), abi.ptr(ranked, abi.P_I32), C.byref(n),
                                abi.ptr(dru, abi.P_F64), abi.ptr(order, abi.P_I32), C.byref(on))
        if rc != 0:
            self._err(rc)
        return {"ranked": ranked[:n.value].copy(), "dru": dru[:running.n + pending.n],
                "order": order[:on.value].copy()}

    # -- M0-M6 -------------------------------------------------------------
    def match(self, ranked_idx, jobs, offers, users, params, groups=None, pool_quota=None,
              max_ports=0):
        ranked_idx = np.ascontiguousarray(ranked_idx, np.int32)
        pool_quota = pool_quota or abi.make_pool_quota(None)
        cons, assign, ports, fail, stats = _CallShapes.match_buffers(params, max_ports)
        rc = self.lib.cook_match(self.pool, abi.ptr(ranked_idx, abi.P_I32), len(ranked_idx),
                                 C.byref(jobs), C.byref(offers),
                                 C.byref(groups) if groups is not None else None,
                                 C.byref(users), C.byref(pool_quota), C.byref(params),
                                 abi.ptr(cons, abi.P_I32), abi.ptr(assign, abi.P_I32),
                                 abi.ptr(ports, abi.P_I32) if max_ports > 0 else None,
                                 int(max_ports), abi.ptr(fail, abi.P_U8), C.byref(stats))
        if rc != 0:
            self._err(rc)
        k = stats.n_considerable
        return {"considerable": cons[:k].copy(), "assign": assign[:k].copy(),
                "ports": ports[:k * max(max_ports, 1)].reshape(k, max(max_ports, 1)).copy(),
                "fail": fail[:k].copy(), "stats": stats.as_dict()}

    def match_failures(self, k_idx):
        """cook_match_failures: per requested considerable job of the LAST match on this handle, the
        per-reason host counts at the job's turn (see cook_b200.cycle.summarize_failures)."""
        k = np.ascontiguousarray(k_idx, np.int32)
        out = (abi.FailureCounts * max(len(k), 1))()
        rc = self.lib.cook_match_failures(self.pool, abi.ptr(k, abi.P_I32), len(k), out)
        if rc != 0:
            self._err(rc)
        return [{"n_vms": o.n_vms, "n_passed": o.n_passed, "n_ports": o.n_ports, "counts": list(o.counts)} for o in out[:len(k)]]

    # -- B1-B6 -------------------------------------------------------------
    def rebalance(self, running, pending, pending_job_id, pending_priority, hosts, users, params,
                  groups=None):
        pj = np.ascontiguousarray(pending_job_id, np.int64)
        pp = np.ascontiguousarray(pending_priority, np.int32)
        dec = (abi.Decision * max(params.max_preemption, 1))()
        vict = np.full(running.t.n + max(params.max_preemption, 1), -1, np.int32)
        n = C.c_int32(0)
        rc = self.lib.cook_rebalance(self.pool, C.byref(running), C.byref(pending),
                                     abi.ptr(pj, abi.P_I64), abi.ptr(pp, abi.P_I32),
                                     C.byref(hosts), C.byref(groups) if groups is not None else None,
                                     C.byref(users), C.byref(params), dec,
                                     abi.ptr(vict, abi.P_I32), C.byref(n))
        if rc != 0:
            self._err(rc)
        return decisions_to_list(dec, vict, n.value)


    def rebalance_trace(self, running, pending, pending_job_id, pending_priority, hosts, users, params,
                        forced=None, forced_only=True, groups=None):
        """cook_rebalance_trace: the same walk plus the rebalancer state the reference's own tests read
        (K18 pending DRU, K21 next-state, job-below-quota).  forced: [(pending_idx, host, [victims],
        mem, cpus, gpus)] applied with next-state instead of being searched."""
        return rebalance_trace_call(
            lambda *a: self.lib.cook_rebalance_trace(self.pool, *a), self._err,
            running, pending, pending_job_id, pending_priority, hosts, users, params, forced, forced_only, groups)

    # -- phase timing / §8e exchange ----------------------------------------
    def last_stats(self, phase):
        ps = abi.PhaseStats()
        rc = self.lib.cook_last_stats(self.pool, int(phase), C.byref(ps))
        if rc != 0:
            self._err(rc)
        return ps.as_dict()

    def exchange_usage(self, n_users_pad, comm=None, world=1):
        """Per-user usage delta of the last match round on this handle, computed on the device and
        all-gathered over `comm` (a cook_comm_init handle).  Returns [world, n_users_pad, 4]."""
        n_pad = 4 * int(n_users_pad)
        out = np.zeros(world * n_pad, np.float64)
        rc = self.lib.cook_exchange_usage(self.pool, comm, int(world), n_pad, abi.ptr(out, abi.P_F64))
        if rc != 0:
            self._err(rc)
        return out.reshape(world, n_users_pad, 4)


def exchange_usage_batch(engines, n_users_pad, n_slots=None, comm=None, world=1):
    """The exchange step for ALL the pools a rank ran in the cycle: their usage deltas (slot i =
    engines[i], the slots beyond stay zero) travel in ONE all-gather (cook_exchange_usage_batch).
    Returns [world, n_slots, n_users_pad, 4]."""
    if not engines:
        raise ValueError("exchange_usage_batch: at least one handle")
    n_slots = len(engines) if n_slots is None else int(n_slots)
    n_pad = 4 * int(n_users_pad)
    out = np.zeros(world * n_slots * n_pad, np.float64)
    handles = (C.c_void_p * len(engines))(*[e.pool for e in engines])
    rc = engines[0].lib.cook_exchange_usage_batch(handles, len(engines), comm, int(world), n_pad, n_slots,
                                                  abi.ptr(out, abi.P_F64))
    if rc != 0:
        engines[0]._err(rc)
    return out.reshape(world, n_slots, n_users_pad, 4)


def comm_unique_id(lib):
    buf = (C.c_uint8 * 128)()
    rc = lib.cook_comm_unique_id(buf)
    if rc != 0:
        raise CookError(rc, "cook_comm_unique_id (is libnccl loadable?)")
    return bytes(buf)


def comm_init(lib, uid, rank, world, device):
    comm = C.c_void_p()
    buf = (C.c_uint8 * 128).from_buffer_copy(uid)
    rc = lib.cook_comm_init(buf, int(rank), int(world), int(device), C.byref(comm))
    if rc != 0:
        raise CookError(rc, "cook_comm_init")
    return comm


def rebalance_trace_call(fn, on_error, running, pending, pending_job_id, pending_priority, hosts, users, params,
                         forced, forced_only, groups):
    """Argument packing of cook_rebalance_trace (`fn` binds the handle; the test suite's CPU checker
    exports the same signature and reuses this packing)."""
    pj = np.ascontiguousarray(pending_job_id, np.int64)
    pp = np.ascontiguousarray(pending_priority, np.int32)
    mp = max(params.max_preemption, 1)
    R, P, H = running.t.n, pending.n, hosts.n
    dec = (abi.Decision * mp)()
    vict = np.full(R + mp, -1, np.int32)
    n = C.c_int32(0)
    forced = forced or []
    fdec = (abi.Decision * max(len(forced), 1))()
    fv = []
    for i, (pi, h, vs, m, c, g) in enumerate(forced):
        fdec[i] = abi.Decision(pi, h, len(fv), len(vs), 0.0, m, c, g)
        fv += list(vs)
    fv = np.array(fv + [0], np.int32)
    pdru, tdru = np.full(max(P, 1), np.nan), np.zeros(R + mp)
    alive, order, n_order = np.zeros(R + mp, np.uint8), np.zeros(R + mp, np.int32), C.c_int32(0)
    hs, sm, sc, sg = np.zeros(H, np.uint8), np.zeros(H), np.zeros(H), np.zeros(H)
    below = np.full(max(P, 1), 255, np.uint8)
    tr = abi.RebTrace(len(forced), fdec, abi.ptr(fv, abi.P_I32), abi.ptr(pdru, abi.P_F64),
                      abi.ptr(tdru, abi.P_F64), abi.ptr(alive, abi.P_U8), abi.ptr(order, abi.P_I32),
                      C.pointer(n_order), abi.ptr(hs, abi.P_U8), abi.ptr(sm, abi.P_F64),
                      abi.ptr(sc, abi.P_F64), abi.ptr(sg, abi.P_F64), 1 if forced_only else 0,
                      abi.ptr(below, abi.P_U8))
    rc = fn(C.byref(running), C.byref(pending), abi.ptr(pj, abi.P_I64), abi.ptr(pp, abi.P_I32),
            C.byref(hosts), C.byref(groups) if groups is not None else None, C.byref(users),
            C.byref(params), dec, abi.ptr(vict, abi.P_I32), C.byref(n), C.byref(tr))
    if rc != 0:
        on_error(rc)
    k = n_order.value
    return {"decisions": decisions_to_list(dec, vict, n.value), "pending_dru": pdru[:P],
            "below_quota": [bool(b) if b != 255 else None for b in below[:P]],
            "order": [int(x) for x in order[:k]], "order_dru": [float(tdru[t]) for t in order[:k]],
            "spare": {h: (float(sm[h]), float(sc[h]), float(sg[h])) for h in range(H) if hs[h]}}


def decisions_to_list(dec, vict, n):
    out = []
    for i in range(n):
        d = dec[i]
        out.append({"pending_idx": d.pending_idx, "host": d.host,
                    "victims": [int(x) for x in vict[d.victim_begin:d.victim_begin + d.victim_count]],
                    "dru": d.dru, "mem": d.mem, "cpus": d.cpus, "gpus": d.gpus})
    return out
