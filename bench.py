#!/usr/bin/env python
"""bench.py — job x offer fit evaluations / second per scheduling cycle.

A "step" is one match cycle of the hot path (considerable-job filter + exact greedy
best-fit matcher = Cook's pending-jobs->considerable-jobs + Fenzo scheduleOnce,
scheduler/scheduler.clj:729-762, :665-671) over one synthetic pool of BASELINE
config #2: 100k pending jobs x 5k offers, cpu+mem fit, 1 pool, all 100k considered.
At N > 1 every rank owns one independent pool of that shape (pools are Cook's
natural shard axis) => weak scaling; after every cycle the per-user
usage deltas are all-gathered (the one exchange step).

  value  : evals/s with inputs resident in HBM (device time, CUDA events on the
           launching stream inside the library, max over ranks)
  e2e    : the same through the C ABI with HOST buffers (pinned), H2D + D2H inside
  --impl reference : the oracle's multi-threaded CPU path on the host cores.
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

WORKLOAD = "C2: 100k pending jobs x 5k offers, cpu+mem fit only, 1 pool per GPU, all jobs considerable"
B_EVAL = 32  # algorithmic bytes per fit evaluation (rem/tot cpus+mem, f64)


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
        t = torch.from_numpy(arr.copy()).pin_memory()
        kw[name] = t.numpy()
        kw.setdefault("_pins", []).append(t)
    pins = kw.pop("_pins", [])
    new = type(struct)(**kw)
    new._pins = pins
    return new


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cook_b200 import traces
    from cook_b200.engine import GpuEngine

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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

    t = traces.gen_c2(seed=2 + rank)
    eng = GpuEngine(pool_name=f"pool-{rank}", device=local)
    ranked = eng.rank(t["running"], t["pending"], t["users"])["ranked"]
    nj = t["jobs"].n
    n_users = t["users"].n_users
    jobs_p, offers_p, users_p = _pin_struct(t["jobs"]), _pin_struct(t["offers"]), _pin_struct(t["users"])
    ranked_p = torch.from_numpy(ranked.copy()).pin_memory()
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    usage_delta = torch.zeros(n_users * 4, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(world * n_users * 4, dtype=torch.float64, device="cuda")
    cols = t["cols"]["pending"]
    owners = t["jobs"].col("user")

    def exchange(m):
        """§8e: per-user usage delta of this rank's pool -> all ranks (one allgather)."""
        placed = m["considerable"][m["assign"] >= 0]
        d = np.zeros((n_users, 4))
        np.add.at(d[:, 0], owners[placed], 1.0)
        np.add.at(d[:, 1], owners[placed], cols["cpus"][placed])
        np.add.at(d[:, 2], owners[placed], cols["mem"][placed])
        usage_delta.copy_(torch.from_numpy(d.reshape(-1)))
        if world > 1:
            dist.all_gather_into_tensor(gathered, usage_delta)

    prm_up = traces.match_params(nj)
    prm_res = traces.match_params(nj, reuse_resident=1)

    # ---------------- resident-input arm (value)
    m = eng.match(ranked_p.numpy(), jobs_p, offers_p, users_p, prm_up)  # upload once
    # clocks are sampled from the warm-up on (same kernels, same load): the timed region of a
    # default run is ~0.1 s, shorter than two nvidia-smi queries
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        torch.cuda.synchronize()
        m = eng.match(ranked_p.numpy(), jobs_p, offers_p, users_p, prm_res)
        exchange(m)
    barrier()
    dev_ms = 0.0
    kern_ms = 0.0
    launches = 0
    evals = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the device-timed region)
        torch.cuda.synchronize()
        m = eng.match(ranked_p.numpy(), jobs_p, offers_p, users_p, prm_res)
        exchange(m)
        s = m["stats"]
        dev_ms += s["ms_considerable"] + s["ms_match"]
        kern_ms += s["ms_match_kernel"]
        launches += s["n_launches"]
        evals += s["evals"]
    barrier()
    wall_res = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    dev_ms_max = max_over_ranks(dev_ms)
    kern_ms_max = max_over_ranks(kern_ms)
    stats_last = m["stats"]

    # ---------------- end-to-end arm (host buffers, H2D + D2H inside)
    for _ in range(min(args.warmup, 2)):
        m = eng.match(ranked_p.numpy(), jobs_p, offers_p, users_p, prm_up)
    barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m = eng.match(ranked_p.numpy(), jobs_p, offers_p, users_p, prm_up)
        exchange(m)
        h2d = m["stats"]["h2d_bytes"]
        d2h = m["stats"]["d2h_bytes"]
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    # ---------------- CPU baseline (rank 0, N == 1 only): the oracle, 1 thread
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import OracleEngine
        ora = OracleEngine()
        tc = time.perf_counter()
        mo = ora.match(ranked, t["jobs"], t["offers"], t["users"], prm_up)
        dt = time.perf_counter() - tc
        same = bool(np.array_equal(mo["assign"], m["assign"]) and np.array_equal(mo["considerable"], m["considerable"]))
        cpu = {"value": mo["stats"]["evals"] / dt, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": "the full workload (100k x 5k = 5e8 evals), one pass, single thread, "
                         "C++ restatement of the reference algorithm (not the JVM)",
               "assignments_identical_to_gpu": same}

    if rank == 0:
        peak, how = _peaks()
        evals_per_step = evals / args.steps
        total_evals = evals * world
        value = total_evals / (dev_ms_max / 1e3)
        kern_s = kern_ms_max / 1e3 / args.steps
        achieved = evals_per_step * B_EVAL / kern_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "match_kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": "job x offer fit evals/sec per scheduling cycle",
            "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (numpy PCG64 seeds 2+rank; cook_b200/traces.py gen_c2)",
            "config": {"workload": WORKLOAD, "jobs": nj, "offers": t["offers"].n,
                       "users": n_users, "pools_per_gpu": 1,
                       "l2": "flushed between timed iterations (512 MiB memset)",
                       "parallelism": f"pool-sharded x{world}, usage allgather per cycle"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                         "kernel": "match_kernel<false>", "kernel_ms": kern_s * 1e3,
                         "algorithmic_bytes_per_launch": evals_per_step * B_EVAL,
                         "note": "streaming-equivalent: 32 B/eval x evals; the offer table is "
                                 "L2/L1-resident so DRAM traffic is far below this by design"},
            "e2e": {"value": total_evals / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cycle": {k: stats_last[k] for k in ("n_considerable", "n_matched", "n_offers_used",
                                                  "n_fast", "n_chunk_rescan", "n_full_rescan")},
            "wall_resident_s": wall_res,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU algorithm (oracle restatement; the JVM + Fenzo
    cannot run here) on all host cores: the per-task VM loop is split over
    threads, which is exactly what Fenzo parallelises."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cook_b200 import traces
    from oracle.pyoracle import OracleEngine
    avail = os.cpu_count() or 1
    t = traces.gen_c2(seed=2)
    ora = OracleEngine()
    ranked = ora.rank(t["running"], t["pending"], t["users"])["ranked"]
    # "all the host threads it can use": the per-task VM loop (5k offers) stops
    # scaling well before 128 threads; pick the fastest count on a short probe.
    probe = traces.match_params(20000)  # long enough that per-job thread hand-off costs show
    rates = []
    for th in sorted({1, 2, 4, 8, 16, 32, 64, min(avail, 64)}):
        if th > avail:
            continue
        tp = time.perf_counter()
        mp_ = ora.match(ranked, t["jobs"], t["offers"], t["users"], probe, threads=th)
        rates.append((mp_["stats"]["evals"] / (time.perf_counter() - tp), th))
    sample_jobs = t["jobs"].n  # the whole workload per step (a few seconds)
    prm = traces.match_params(sample_jobs)
    # the tail of the queue (jobs that fit nowhere) is cheap per job and favours fewer threads
    # than the probe: time one full pass for the two best probe counts and for one thread,
    # keep the fastest (these passes are the warm-up)
    cands = sorted({th for _, th in sorted(rates, reverse=True)[:2]} | {1})
    best = (0.0, 1)
    for th in cands:
        tp = time.perf_counter()
        mp_ = ora.match(ranked, t["jobs"], t["offers"], t["users"], prm, threads=th)
        rate = mp_["stats"]["evals"] / (time.perf_counter() - tp)
        if rate > best[0]:
            best = (rate, th)
    cores = best[1]
    for _ in range(max(0, args.warmup - len(cands))):
        ora.match(ranked, t["jobs"], t["offers"], t["users"], prm, threads=cores)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        m = ora.match(ranked, t["jobs"], t["offers"], t["users"], prm, threads=cores)
        evals += m["stats"]["evals"]
    dt = time.perf_counter() - t0
    v = evals / dt
    sample = (f"all {sample_jobs} considerable jobs x 5k offers per step "
              f"({sample_jobs * 5000:.3g} evals), {cores} threads, C++ restatement (not the JVM)")
    print(json.dumps({
        "impl": "reference", "metric": "job x offer fit evals/sec per scheduling cycle",
        "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (gen_c2 seed 2)",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
