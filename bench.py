#!/usr/bin/env python
"""bench.py — job x offer fit evaluations / second per scheduling cycle.

A "step" is one scheduling cycle of the hot path over every pool this process owns:
  cook_rank  (DRU ranking, scheduler/scheduler.clj:2073-2091)
  cook_match (considerable-job filter + exact greedy best-fit matcher = Cook's
              pending-jobs->considerable-jobs + Fenzo scheduleOnce, :729-762, :665-671)
  cook_rebalance (configs c4 / c5: preemption-victim search, rebalancer.clj:434-467)
  cook_exchange_usage (per-user usage delta computed on the device + ONE ncclAllGather; the
              gathered totals are folded into the next cycle's quota-group usage, :2125-2157)

--config (BASELINE.json `configs`):
  c2 (default)  100k pending jobs x 5k offers, cpu+mem fit only, 1 pool PER GPU (weak scaling)
  c3            1M jobs x 20k nodes, 4 pools, all constraint kinds      (strong scaling, LPT pools)
  c4            c3 + rebalancer sweep (400k running tasks, max-preemption 128)
  c5            10M jobs x 100k nodes, 16 pools, full rank + match + rebalance cycle

  value  : evals/s with inputs resident in HBM: evals / device time of the whole step (CUDA
           events on the launching stream inside the library), max over ranks.  The step
           includes ranking, rebalancing and the exchange - not only the matcher.
  e2e    : the same step through the C ABI with HOST buffers, H2D + D2H inside, wall clock.
  --impl reference : the reference algorithm's CPU restatement on the host cores, same step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": "C2: 100k pending jobs x 5k offers, cpu+mem fit only, 1 pool per GPU, all jobs considerable; "
          "cycle = rank + match + usage exchange",
    "c3": "C3: 1M jobs x 20k nodes in 4 pools (40/30/20/10 %), host-placement + gpu/ports/attribute constraints; "
          "cycle = rank + match + usage exchange per pool, pools placed on GPUs by LPT",
    "c4": "C4: C3's pools with 400k running tasks + DRU rebalancer sweep (max-preemption 128); "
          "cycle = rank + match + rebalance + usage exchange per pool, pools placed on GPUs by LPT",
    "c5": "C5: 10M jobs x 100k nodes in 16 pools (Zipf), full DRU rank + match + rebalance cycle + usage "
          "allgather; pools placed on GPUs by LPT",
}
B_EVAL = {"c2": 32, "c3": 96, "c4": 96, "c5": 96}   # algorithmic bytes per fit evaluation
B_RANK = 68                                           # bytes per ranked task
B_REBAL = 40                                          # bytes per running task per evaluated pending job


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _pin_struct(struct):
    """Re-home every column of an ABI struct into pinned host memory."""
    import torch
    from cook_b200 import abi
    kw = {}
    for name, ctype in struct._fields_:
        arr = struct.col(name)
        if arr is None:
            if ctype not in abi._NP and not isinstance(getattr(struct, name), abi._SoA):
                kw[name] = getattr(struct, name)
            continue
        if isinstance(arr, abi._SoA):
            kw[name] = _pin_struct(arr)
            continue
        t = torch.from_numpy(arr.copy()).pin_memory()
        kw[name] = t.numpy()
        kw.setdefault("_pins", []).append(t)
    pins = kw.pop("_pins", [])
    new = type(struct)(**kw)
    new._pins = pins
    return new


# ------------------------------------------------------------------ the workload
def pool_plan(config, world):
    """[(pool index, owning rank)] - Cook's shard axis is the pool (scheduler.clj:2488-2517)."""
    from cook_b200 import sharding, traces
    if config == "c2":
        return [(r, r) for r in range(world)]          # weak scaling: one C2 pool per GPU
    sizes = traces.pool_sizes(config)
    owner = sharding.assign_pools_lpt([sharding.pool_cycle_cost(j, o) for j, o, _, _ in sizes], world)
    return [(p, owner[p]) for p in range(len(sizes))]


def gen_pool_inputs(config, p, scale=1.0):
    from cook_b200 import traces
    if config == "c2":
        t = traces.gen_c2(seed=2 + p)
        t["groups"] = None
        t["host_lifetime_mins"] = 0
        return t
    return traces.gen_config_pool(config, p, scale=scale)


def config_dict(config, world, plan):
    """Identical in both arms (ours / reference): the driver compares them."""
    from cook_b200 import traces
    if config == "c2":
        pools = [{"jobs": 100_000, "offers": 5_000, "users": 1_000, "running": 20_000}] * world
    else:
        pools = [dict(zip(("jobs", "offers", "users", "running"), s)) for s in traces.pool_sizes(config)]
    return {"workload": WORKLOADS[config], "config": config, "n_pools": len(plan),
            "jobs": int(sum(p["jobs"] for p in pools)), "offers": int(sum(p["offers"] for p in pools)),
            "users": int(sum(p["users"] for p in pools)), "running": int(sum(p["running"] for p in pools)),
            "pools_per_gpu": [sum(1 for _, r in plan if r == g) for g in range(world)],
            "cycle": "rank + match" + (" + rebalance" if config in ("c4", "c5") else "") + " + usage exchange",
            "l2": "flushed between timed iterations (512 MiB memset)",
            "parallelism": (f"pool-sharded x{world} (one pool per GPU, weak)" if config == "c2" else
                            f"{len(plan)} pools placed on {world} GPU(s) by LPT (strong)") +
                           ", device-side usage delta + ONE ncclAllGather per cycle for all of a rank's pools"}


def GROUP_QUOTA():
    from cook_b200 import abi
    return abi.make_pool_quota({"count": 1e15, "cpus": 1e15, "mem": 1e18, "gpus": 1e15})


class PoolRun:
    """One pool bound to this process: inputs (pinned host columns), engine, per-cycle call sequence."""

    def __init__(self, config, p, eng_cls, device, pin):
        from cook_b200 import abi, traces
        self.config, self.p = config, p
        t = gen_pool_inputs(config, p)
        self.t = t
        f = _pin_struct if pin else (lambda x: x)
        self.running, self.pending, self.users = f(t["running"]), f(t["pending"]), f(t["users"])
        self.jobs, self.offers = f(t["jobs"]), f(t["offers"])
        self.groups = f(t["groups"]) if t.get("groups") is not None else None
        self.nj, self.no, self.nu = t["jobs"].n, t["offers"].n, t["users"].n_users
        self.max_ports = 2 if config != "c2" else 0
        self.hl = t.get("host_lifetime_mins", 0)
        self.reb = t.get("rebalance")
        if self.reb is not None and pin:
            r = self.reb
            self.reb = dict(r, running=f(r["running"]), pending=f(r["pending"]), hosts=f(r["hosts"]), users=f(r["users"]),
                            groups=f(r["groups"]) if r["groups"] is not None else None)
        self.eng = eng_cls(pool_name=f"{config}-pool-{p}", device=device) if device is not None else eng_cls()
        self.ranked = None
        self.group_usage = np.zeros(4)

    def rank(self):
        # configs with quota groups (c3-c5): the quota-group filter runs against the usage the
        # exchange step gathered (limits far above the totals: the pass executes, the queue is the
        # same every cycle, so steps stay comparable)
        gq = None if self.config == "c2" else GROUP_QUOTA()
        out = self.eng.rank(self.running, self.pending, self.users, group_quota=gq, group_usage=self.group_usage)
        self.ranked = out["ranked"]
        return out

    def match(self, resident, **kw):
        from cook_b200 import traces
        prm = traces.match_params(self.nj, host_lifetime_mins=self.hl, reuse_resident=1 if resident else 0)
        return self.eng.match(self.ranked, self.jobs, self.offers, self.users, prm, groups=self.groups,
                              max_ports=self.max_ports, **kw)

    def rebalance(self, **kw):
        r = self.reb
        return self.eng.rebalance(r["running"], r["pending"], r["pending_job_id"], r["pending_priority"], r["hosts"],
                                  r["users"], r["params"], groups=r["groups"], **kw)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cook_b200 import abi
    from cook_b200.engine import GpuEngine, comm_init, comm_unique_id, exchange_usage_batch, load_library

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
    torch.cuda.set_device(local)
    comm = None
    lib = load_library()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the library's own communicator: 128 bytes of unique id travel over the host's control
        # plane (here torch.distributed), everything else is libcookgpu + NCCL
        box = [comm_unique_id(lib) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = comm_init(lib, box[0], rank, world, local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cfg = args.config
    plan = pool_plan(cfg, world)
    mine = [p for p, r in plan if r == rank]
    slots = max(sum(1 for _, r in plan if r == g) for g in range(world))   # exchanges per cycle (same on all ranks)
    pools = [PoolRun(cfg, p, GpuEngine, local, pin=True) for p in mine]
    nu_pad = int(max_over_ranks(max([pr.nu for pr in pools] + [1])))
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    has_reb = cfg in ("c4", "c5")

    def cycle(resident):
        """One scheduling cycle over this rank's pools.  Returns per-phase device ms, evals, placements."""
        ph = {"rank": 0.0, "match": 0.0, "match_kernel": 0.0, "rebalance": 0.0, "exchange": 0.0}
        ev = pl = ln = h2d = d2h = dec = 0
        last = None
        for i in range(slots):
            pr = pools[i] if i < len(pools) else None
            if pr is not None:
                pr.rank()
                s = pr.eng.last_stats(abi.PHASE_RANK)
                ph["rank"] += s["ms_device"]; ln += s["n_launches"]; h2d += s["h2d_bytes"]; d2h += s["d2h_bytes"]
                m = pr.match(resident)
                s = m["stats"]
                ph["match"] += s["ms_considerable"] + s["ms_match"]; ph["match_kernel"] += s["ms_match_kernel"]
                ev += s["evals"]; pl += s["n_matched"]; ln += s["n_launches"]; h2d += s["h2d_bytes"]; d2h += s["d2h_bytes"]
                last = s
                if has_reb:
                    d = pr.rebalance()
                    dec += len(d)
                    s = pr.eng.last_stats(abi.PHASE_REBALANCE)
                    ph["rebalance"] += s["ms_device"]; h2d += s["h2d_bytes"]; d2h += s["d2h_bytes"]
        # ONE exchange per cycle for all of this rank's pools (cook_exchange_usage_batch): LPT balances the
        # SUM of a rank's pools; a collective per pool slot would make every slot as long as its slowest rank
        g = exchange_usage_batch([q.eng for q in pools], nu_pad, n_slots=slots, comm=comm, world=world)   # [world, slots, nu_pad, 4]
        s = pools[0].eng.last_stats(abi.PHASE_EXCHANGE)
        ph["exchange"] += s["ms_device"]; ln += s["n_launches"]; d2h += s["d2h_bytes"]
        # the collective's result is consumed: every pool's quota-group usage for the NEXT rank
        # cycle is the sum of all pools' deltas (aggregate-quota-groups, scheduler.clj:2125-2132)
        tot = g.sum(axis=(0, 1, 2))
        for q in pools:
            q.group_usage = tot
        return ph, ev, pl, ln, h2d, d2h, dec, last

    # ---------------- resident-input arm (value): upload once, then reuse_resident
    cycle(False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        torch.cuda.synchronize()
        cycle(True)
    barrier()
    tot = {"rank": 0.0, "match": 0.0, "match_kernel": 0.0, "rebalance": 0.0, "exchange": 0.0}
    evals = places = launches = decisions = 0
    last = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the device-timed region)
        torch.cuda.synchronize()
        ph, ev, pl, ln, _, _, dec, last = cycle(True)
        for k in tot:
            tot[k] += ph[k]
        evals += ev; places += pl; launches += ln; decisions += dec
    barrier()
    wall_res = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    step_dev_ms = tot["rank"] + tot["match"] + tot["rebalance"] + tot["exchange"]
    dev_ms_max = max_over_ranks(step_dev_ms)
    phases_max = {k: max_over_ranks(v) / args.steps for k, v in tot.items()}
    total_evals = sum_over_ranks(evals)
    total_places = sum_over_ranks(places)
    total_launches = sum_over_ranks(launches)

    # ---------------- end-to-end arm (host buffers, H2D + D2H inside, rank included)
    for _ in range(min(args.warmup, 2)):
        cycle(False)
    barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, _, _, h2d, d2h, _, _ = cycle(False)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d_all, d2h_all = sum_over_ranks(h2d), sum_over_ranks(d2h)

    # ---------------- C2 only, N == 1: the non-saturating variant (every job placeable): the
    # dependency-chain bound without the cluster filling up after a quarter of the queue
    nonsat = None
    if cfg == "c2" and world == 1 and not args.no_nonsat:
        from cook_b200 import traces
        t = traces.gen_c2(seed=2, offer_scale=6)
        e2 = GpuEngine(pool_name="c2-nonsaturating", device=local)
        rk = e2.rank(t["running"], t["pending"], t["users"])["ranked"]
        e2.match(rk, t["jobs"], t["offers"], t["users"], traces.match_params(t["jobs"].n))
        ms = []
        for _ in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            m2 = e2.match(rk, t["jobs"], t["offers"], t["users"], traces.match_params(t["jobs"].n, reuse_resident=1))
            ms.append(m2["stats"]["ms_match_kernel"])
        k_ms = float(np.median(ms))
        nonsat = {"workload": "C2 with 6x the offer capacity: all 100k jobs placeable", "placements": int(m2["stats"]["n_matched"]),
                  "kernel_ms": k_ms, "us_per_placement": 1e3 * k_ms / max(1, m2["stats"]["n_matched"]),
                  "evals_per_s": m2["stats"]["evals"] / (k_ms / 1e3)}
        e2.close()

    # ---------------- CPU baseline (rank 0, N == 1 only): the oracle, 1 thread, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import OracleEngine
        ora = OracleEngine()
        pr = pools[0]
        tc = time.perf_counter()
        ro = ora.rank(pr.t["running"], pr.t["pending"], pr.t["users"])
        from cook_b200 import traces
        prm = traces.match_params(pr.nj, host_lifetime_mins=pr.hl)
        mo = ora.match(ro["ranked"], pr.t["jobs"], pr.t["offers"], pr.t["users"], prm, groups=pr.t.get("groups"),
                       max_ports=pr.max_ports)
        dt = time.perf_counter() - tc
        pr.group_usage = np.zeros(4)
        pr.rank()
        mg = pr.match(False)
        same = bool(np.array_equal(mo["assign"], mg["assign"]) and np.array_equal(mo["considerable"], mg["considerable"]))
        cpu = {"value": mo["stats"]["evals"] / dt, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": f"pool 0 of the workload ({pr.nj} jobs x {pr.no} offers = {mo['stats']['evals']:.3g} evals), one "
                         "rank + match pass, single thread, C++ restatement of the reference algorithm (not the JVM)",
               "assignments_identical_to_gpu": same}

    if rank == 0:
        peak, how = _peaks()
        value = total_evals / (dev_ms_max / 1e3)
        kern_s = phases_max["match_kernel"] / 1e3
        evals_per_step_rank0 = evals / args.steps
        achieved = evals_per_step_rank0 * B_EVAL[cfg] / max(kern_s, 1e-12) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "match_kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("config", "c2") == cfg:
                traffic = tj.get("dram_bytes_per_launch")
        n_tasks = sum(pr.t["running"].n + pr.t["pending"].n for pr in pools)
        phase_block = {
            "rank": {"ms": phases_max["rank"], "roofline": {
                "bound": "hbm", "achieved": n_tasks * B_RANK / max(phases_max["rank"], 1e-9) / 1e6, "peak": peak,
                "unit": "GB/s", "frac": n_tasks * B_RANK / max(phases_max["rank"], 1e-9) / 1e6 / peak,
                "algorithmic_bytes": n_tasks * B_RANK}},
            "match": {"ms": phases_max["match"], "kernel_ms": phases_max["match_kernel"], "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "algorithmic_bytes": evals_per_step_rank0 * B_EVAL[cfg]}},
            "exchange": {"ms": phases_max["exchange"]},
        }
        if has_reb:
            rb = sum(pr.reb["running"].t.n * B_REBAL * pr.reb["pending"].n for pr in pools)
            phase_block["rebalance"] = {"ms": phases_max["rebalance"], "decisions_per_cycle": decisions / args.steps,
                                        "roofline": {"bound": "hbm", "achieved": rb / max(phases_max["rebalance"], 1e-9) / 1e6,
                                                     "peak": peak, "unit": "GB/s",
                                                     "frac": rb / max(phases_max["rebalance"], 1e-9) / 1e6 / peak,
                                                     "algorithmic_bytes": rb,
                                                     "note": "upper bound: R x 40 B per pending job walked (SURVEY §8d B_rebal)"}}
        line = {
            "metric": "job x offer fit evals/sec per scheduling cycle",
            "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak" if cfg == "c2" else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (numpy PCG64, fixed seeds; cook_b200/traces.py)",
            "config": config_dict(cfg, world, plan),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                         "kernel": "match_kernel<%s>" % ("false" if cfg == "c2" else "true"), "kernel_ms": kern_s * 1e3,
                         "algorithmic_bytes_per_launch": evals_per_step_rank0 * B_EVAL[cfg],
                         "traffic_source": "ncu --set full capture of the same kernel build (profiles/), per launch",
                         "note": "streaming-equivalent: B_eval x evals; the offer table is L2/shared-memory resident so "
                                 "DRAM traffic is far below this by design; the kernel is bound by the serial "
                                 "dependency chain of the exact greedy (see us_per_placement)"},
            "e2e": {"value": total_evals / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_s / args.steps * 1e3,
                    "includes": "cook_rank + cook_match" + (" + cook_rebalance" if has_reb else "") +
                                " + cook_exchange_usage, pinned host buffers, H2D and D2H inside"},
            "gpu_launches": int(total_launches),
            "clocks": clocks,
            "phases": phase_block,
            "placements_per_step": total_places / args.steps,
            "us_per_placement": 1e3 * phases_max["match_kernel"] / max(1.0, places / args.steps),
            "cycle": {k: last[k] for k in ("n_considerable", "n_matched", "n_offers_used",
                                           "n_fast", "n_chunk_rescan", "n_full_rescan")} if last else None,
            "wall_resident_s": wall_res,
        }
        if nonsat:
            line["nonsaturating"] = nonsat
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    for pr in pools:
        pr.eng.close()
    if comm is not None:
        lib.cook_comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU algorithm (oracle restatement; the JVM + Fenzo cannot run here) on
    all host cores: the per-task VM loop is split over threads, which is exactly what Fenzo
    parallelises.  Same step as ours (rank + match [+ rebalance]); at N > 1 rank 0 alone works and
    each step handles ONE of the workload's pools, rotating (a bounded sample of the N-GPU step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cook_b200 import traces
    from oracle.pyoracle import OracleEngine
    cfg = args.config
    world = args.gpus
    plan = pool_plan(cfg, world)
    avail = os.cpu_count() or 1
    ora = OracleEngine()
    has_reb = cfg in ("c4", "c5")
    pool_ids = [p for p, _ in plan]
    cache = {}

    def get(p):
        if p not in cache:
            if len(cache) >= 2:
                cache.pop(next(iter(cache)))
            cache[p] = gen_pool_inputs(cfg, p)
        return cache[p]

    def one(p, threads):
        t = get(p)
        r = ora.rank(t["running"], t["pending"], t["users"], group_quota=None if cfg == "c2" else GROUP_QUOTA(),
                     group_usage=np.zeros(4))
        prm = traces.match_params(t["jobs"].n, host_lifetime_mins=t.get("host_lifetime_mins", 0))
        m = ora.match(r["ranked"], t["jobs"], t["offers"], t["users"], prm, groups=t.get("groups"),
                      max_ports=2 if cfg != "c2" else 0, threads=threads)
        if has_reb:
            rb = t["rebalance"]
            ora.rebalance(rb["running"], rb["pending"], rb["pending_job_id"], rb["pending_priority"], rb["hosts"],
                          rb["users"], rb["params"], groups=rb["groups"])
        return m["stats"]["evals"]

    # "all the host threads it can use": the per-task VM loop stops scaling well before 128
    # threads; time one pass of the first pool for a few counts and keep the fastest
    cands = sorted({th for th in (1, 8, 16, 32, 64, avail) if th <= avail})
    best = (0.0, 1)
    for th in cands:
        tp = time.perf_counter()
        ev = one(pool_ids[0], th)
        rate = ev / (time.perf_counter() - tp)
        if rate > best[0]:
            best = (rate, th)
    cores = best[1]
    for i in range(max(0, args.warmup - len(cands))):
        one(pool_ids[i % len(pool_ids)], cores)
    t0 = time.perf_counter()
    evals = 0
    for i in range(args.steps):
        evals += one(pool_ids[i % len(pool_ids)], cores)
    dt = time.perf_counter() - t0
    v = evals / dt
    sample = (f"one pool of the workload per step, rotating over its {len(pool_ids)} pool(s) "
              f"({evals / args.steps:.3g} evals per step): rank + match" + (" + rebalance" if has_reb else "") +
              f", {cores} threads, C++ restatement (not the JVM)")
    emit({
        "impl": "reference", "metric": "job x offer fit evals/sec per scheduling cycle",
        "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if cfg == "c2" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (numpy PCG64, fixed seeds; cook_b200/traces.py)",
        "config": config_dict(cfg, world, plan),
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_JSON_FD = None


def emit(line):
    """The run's ONE line of stdout.  Libraries write to fd 1 behind Python's back (NCCL announces its
    version there at the first communicator, torch's or ours), so main() points fd 1 at stderr for the
    whole run and the JSON line goes to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nonsat", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
