"""ctypes wrapper around oracle/libcookoracle.so — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Mirrors cook_b200.engine.GpuEngine's
call shapes so tests can diff product vs oracle on identical inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cook_b200 import abi
from cook_b200.engine import _CallShapes, _empty_tasks, decisions_to_list, rebalance_trace_call

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcookoracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_version.restype = C.c_char_p
    for n in ("oracle_rank", "oracle_match", "oracle_considerable", "oracle_rebalance",
              "oracle_match_mt", "oracle_rebalance_trace"):
        if hasattr(lib, n):
            getattr(lib, n).restype = C.c_int32
    return lib


class OracleEngine:
    def __init__(self, dru_mode=0, lib=None):
        self.lib = lib or load()
        self.dru_mode = dru_mode

    def set_naive_merge(self, on):
        self.lib.oracle_set_naive_merge(int(on))

    def rank(self, running, pending, users, pool_quota=None, group_quota=None, group_usage=None,
             params=None):
        running = running or _empty_tasks()
        pool_quota = pool_quota or abi.make_pool_quota(None)
        group_quota = group_quota or abi.make_pool_quota(None)
        params = params or abi.RankParams(100, 0, 0.0, 0.0)
        gu = np.ascontiguousarray(group_usage if group_usage is not None else np.zeros(4), np.float64)
        ranked, n, dru, order, on = _CallShapes.rank_buffers(running, pending)
        rc = self.lib.oracle_rank(int(self.dru_mode), C.byref(running), C.byref(pending),
                                  C.byref(users), C.byref(pool_quota), C.byref(group_quota),
                                  abi.ptr(gu, abi.P_F64), C.byref(params),
                                  abi.ptr(ranked, abi.P_I32), C.byref(n), abi.ptr(dru, abi.P_F64),
                                  abi.ptr(order, abi.P_I32), C.byref(on))
        if rc != 0:
            raise RuntimeError(f"oracle_rank rc={rc}")
        return {"ranked": ranked[:n.value].copy(), "dru": dru[:running.n + pending.n],
                "order": order[:on.value].copy()}

    def match(self, ranked_idx, jobs, offers, users, params, groups=None, pool_quota=None,
              max_ports=0, threads=1):
        ranked_idx = np.ascontiguousarray(ranked_idx, np.int32)
        pool_quota = pool_quota or abi.make_pool_quota(None)
        cons, assign, ports, fail, stats = _CallShapes.match_buffers(params, max_ports)
        fn = self.lib.oracle_match
        args = [abi.ptr(ranked_idx, abi.P_I32), len(ranked_idx), C.byref(jobs), C.byref(offers),
                C.byref(groups) if groups is not None else None, C.byref(users),
                C.byref(pool_quota), C.byref(params), abi.ptr(cons, abi.P_I32),
                abi.ptr(assign, abi.P_I32), abi.ptr(ports, abi.P_I32) if max_ports > 0 else None,
                int(max_ports), abi.ptr(fail, abi.P_U8), C.byref(stats)]
        if threads > 1:
            fn = self.lib.oracle_match_mt
            args.append(int(threads))
        rc = fn(*args)
        if rc != 0:
            raise RuntimeError(f"oracle_match rc={rc}")
        k = stats.n_considerable
        return {"considerable": cons[:k].copy(), "assign": assign[:k].copy(),
                "ports": ports[:k * max(max_ports, 1)].reshape(k, max(max_ports, 1)).copy(),
                "fail": fail[:k].copy(), "stats": stats.as_dict()}

    def rebalance(self, running, pending, pending_job_id, pending_priority, hosts, users, params,
                  groups=None):
        pj = np.ascontiguousarray(pending_job_id, np.int64)
        pp = np.ascontiguousarray(pending_priority, np.int32)
        dec = (abi.Decision * max(params.max_preemption, 1))()
        vict = np.full(running.t.n + max(params.max_preemption, 1), -1, np.int32)
        n = C.c_int32(0)
        rc = self.lib.oracle_rebalance(int(self.dru_mode), C.byref(running), C.byref(pending),
                                       abi.ptr(pj, abi.P_I64), abi.ptr(pp, abi.P_I32),
                                       C.byref(hosts),
                                       C.byref(groups) if groups is not None else None,
                                       C.byref(users), C.byref(params), dec,
                                       abi.ptr(vict, abi.P_I32), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"oracle_rebalance rc={rc}")
        return decisions_to_list(dec, vict, n.value)

    def match_failures(self, ranked_idx, jobs, offers, users, params, k_idx, groups=None, pool_quota=None):
        """Twin of GpuEngine.match + match_failures: the same match, counters at the requested turns."""
        ranked_idx = np.ascontiguousarray(ranked_idx, np.int32)
        pool_quota = pool_quota or abi.make_pool_quota(None)
        k = np.ascontiguousarray(k_idx, np.int32)
        out = (abi.FailureCounts * max(len(k), 1))()
        self.lib.oracle_match_failures.restype = C.c_int32
        rc = self.lib.oracle_match_failures(abi.ptr(ranked_idx, abi.P_I32), len(ranked_idx), C.byref(jobs), C.byref(offers),
                                            C.byref(groups) if groups is not None else None, C.byref(users),
                                            C.byref(pool_quota), C.byref(params), abi.ptr(k, abi.P_I32), len(k), out)
        if rc != 0:
            raise RuntimeError(f"oracle_match_failures rc={rc}")
        return [{"n_vms": o.n_vms, "n_passed": o.n_passed, "n_ports": o.n_ports, "counts": list(o.counts)} for o in out[:len(k)]]

    def rebalance_trace(self, running, pending, pending_job_id, pending_priority, hosts, users, params,
                        forced=None, forced_only=True, groups=None):
        """Rebalancer state as the reference's own tests read it (K18 pending DRU, K21 next-state).
        forced: [(pending_idx, host, [victims], mem, cpus, gpus)] applied with next-state instead
        of searching."""
        def err(rc):
            raise RuntimeError(f"oracle_rebalance_trace rc={rc}")
        return rebalance_trace_call(
            lambda *a: self.lib.oracle_rebalance_trace(int(self.dru_mode), *a), err,
            running, pending, pending_job_id, pending_priority, hosts, users, params, forced, forced_only, groups)
