"""(Lives under tests/ because it uses the oracle as the checker beside every measurement.)
Times the other rows of the path (rank R1-R7, rebalancer B1-B6, and the matcher on a
full config-#3 pool) through the C ABI with HOST buffers, next to the CPU restatement on the
same inputs, and checks that the results are identical.  Writes one JSON object to
gpurun_out/path_measurements.json.  `PROF_NO_GPU=1` runs the generators and the CPU side only
(dry run in a container without a GPU); `PROF_SCALE=0.1` shrinks every case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cook_b200 import traces  # noqa: E402
from oracle.pyoracle import OracleEngine, build  # noqa: E402

NO_GPU = os.environ.get("PROF_NO_GPU") == "1"
SCALE = float(os.environ.get("PROF_SCALE", "1"))
build()
orc = OracleEngine()
gpu = None
if not NO_GPU:
    from cook_b200.engine import GpuEngine
    gpu = GpuEngine()


def timed(fn, reps):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best, out


def n_(x):
    return max(8, int(x * SCALE))


res = {}


def rank_case(name, seed, nj, nu, nr, quotas=False):
    from cook_b200 import abi
    t = traces.gen_pool(seed, nj, 16, nu, nr)
    args = (t["running"], t["pending"], t["users"])
    kw = {}
    r = {"pending": nj, "running": nr, "users": nu, "pool_and_group_quota": quotas}
    if quotas:   # both sequential filters bind about two thirds of the way down the queue
        tot_c = float(np.sum(t["pending"].col("cpus"))) + float(np.sum(t["running"].col("cpus")))
        q = dict(count=1e12, cpus=round(0.66 * tot_c), mem=1e15, gpus=1e12)
        kw = dict(pool_quota=abi.make_pool_quota(q), group_quota=abi.make_pool_quota(dict(q, cpus=round(0.7 * tot_c))),
                  group_usage=np.array([10.0, 100.0, 1000.0, 0.0]))
    ms_o, ro = timed(lambda: orc.rank(*args, **kw), 1)
    r["cpu_ms"] = round(ms_o, 2)
    r["ranked"] = int(len(ro["ranked"]))
    if gpu:
        gpu.rank(*args, **kw)
        ms_g, rg = timed(lambda: gpu.rank(*args, **kw), 3)
        r["gpu_ms_e2e"] = round(ms_g, 3)
        r["identical"] = bool(np.array_equal(rg["ranked"], ro["ranked"]) and np.array_equal(rg["order"], ro["order"])
                              and np.array_equal(rg["dru"], ro["dru"]))
        r["tasks_per_s_gpu"] = (nj + nr) / (ms_g / 1e3)
    r["tasks_per_s_cpu"] = (nj + nr) / (ms_o / 1e3)
    res[name] = r
    print(name, r, flush=True)


def rebalance_case(name, seed, nr, npend, nh, nu, mp):
    t = traces.gen_rebalance(seed, nr, npend, nh, nu, max_preemption=mp)
    args = (t["running"], t["pending"], t["pending_job_id"], t["pending_priority"], t["hosts"], t["users"], t["params"])
    r = {"running": nr, "pending": npend, "hosts": nh, "users": nu, "max_preemption": mp}
    ms_o, do = timed(lambda: orc.rebalance(*args, groups=t["groups"]), 1)
    r["cpu_ms"] = round(ms_o, 2)
    r["decisions"] = len(do)
    r["victims"] = sum(len(d["victims"]) for d in do)
    if gpu:
        gpu.rebalance(*args, groups=t["groups"])
        ms_g, dg = timed(lambda: gpu.rebalance(*args, groups=t["groups"]), 3)
        r["gpu_ms_e2e"] = round(ms_g, 3)
        r["identical"] = dg == do
        # the search examines every (pending job, running task) pair once per pending job
        r["pairs_per_s_gpu"] = npend * nr / (ms_g / 1e3)
    r["pairs_per_s_cpu"] = npend * nr / (ms_o / 1e3)
    res[name] = r
    print(name, r, flush=True)


def c3_case(name, seed, nj, no, nu, nr):
    t = traces.gen_c3_pool(seed, nj, no, nu, nr)
    ranked = orc.rank(t["running"], t["pending"], t["users"])["ranked"]
    prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"])
    r = {"jobs": nj, "offers": no}
    threads = min(16, os.cpu_count() or 1)   # the restatement's per-job barrier does not scale further
    ms_o, mo = timed(lambda: orc.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"],
                                       max_ports=2, threads=threads), 1)
    r["cpu_ms"] = round(ms_o, 1)
    r["cpu_threads"] = threads
    r["evals"] = int(mo["stats"]["evals"])
    r["n_matched"] = int(mo["stats"]["n_matched"])
    if gpu:
        gpu.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
        ms_g, mg = timed(lambda: gpu.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"],
                                           max_ports=2), 3)
        r["gpu_ms_e2e"] = round(ms_g, 2)
        r["gpu_kernel_ms"] = round(mg["stats"]["ms_match_kernel"], 2)
        r["identical"] = bool(np.array_equal(mg["considerable"], mo["considerable"]) and
                              np.array_equal(mg["assign"], mo["assign"]) and np.array_equal(mg["ports"], mo["ports"]))
        r["evals_per_s_gpu_e2e"] = r["evals"] / (ms_g / 1e3)
        r["rescans"] = {k: int(mg["stats"][k]) for k in ("n_fast", "n_chunk_rescan", "n_full_rescan")}
    r["evals_per_s_cpu"] = r["evals"] / (ms_o / 1e3)
    res[name] = r
    print(name, r, flush=True)


CASES = os.environ.get("PROF_CASES", "rank,rank_quota,rebalance,c3").split(",")
if "rank" in CASES:
    rank_case("rank_c2", 2, n_(100_000), n_(1_000), n_(20_000))
    rank_case("rank_c5_pool", 5, n_(625_000), n_(5_000), n_(100_000))
if "rank_quota" in CASES:
    rank_case("rank_c5_pool_quotas", 5, n_(625_000), n_(5_000), n_(100_000), quotas=True)
if "rebalance" in CASES:
    rebalance_case("rebalance_c4_pool", 4, n_(100_000), n_(400), n_(5_000), n_(2_000), 64)
if "c3" in CASES:
    c3_case("match_c3_pool", 300, n_(250_000), n_(5_000), n_(2_000), n_(50_000))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "path_measurements.json"), "w") as f:
    json.dump(res, f, indent=1)
print("OK", flush=True)
