"""Deterministic synthetic job+offer traces for BASELINE.json's configs.

SURVEY.md §8(d): all generators use numpy.random.Generator(PCG64(seed)) with
fixed seeds and value grids chosen so that every partial sum is exact in f64
(cpus multiples of 2^-1, mem integer MiB), which makes DRU sums and fitness
arithmetic independent of evaluation order — the same property the exact
JVM/GPU parity argument in DESIGN.md relies on.

Shapes follow the reference's own simulator inputs where it has them
(scheduler/test/cook/test/zz_simulator.clj:720-744 for C1).
"""
import numpy as np

from . import abi

INT64_MAX = abi.INT64_MAX


def _zipf_owner(rng, n, n_users, a=1.1):
    w = 1.0 / np.arange(1, n_users + 1) ** a
    w /= w.sum()
    return rng.choice(n_users, size=n, p=w).astype(np.int32)


def _tasks(rng, n, users, running, id_base, cpus_choices, mem_fn, prio_p=(0.1, 0.8, 0.1),
           gpus=None):
    prio = rng.choice(np.array([10, 50, 90], np.int32), size=n, p=prio_p).astype(np.int32)
    cpus = rng.choice(np.asarray(cpus_choices, np.float64), size=n)
    mem = mem_fn(rng, n)
    job_id = (id_base + np.arange(n)).astype(np.int64) * 2
    if running:
        start = (1_600_000_000_000 + rng.integers(0, 86_400_000, size=n)).astype(np.int64)
        task_id = job_id + 1
    else:
        start = np.full(n, INT64_MAX, np.int64)
        task_id = np.full(n, -1, np.int64)
    return dict(user=users, priority=prio, start_time=start, task_id=task_id, job_id=job_id,
                cpus=cpus, mem=mem, gpus=np.zeros(n) if gpus is None else gpus)


def gen_pool(seed, n_jobs, n_offers, n_users, n_running, *, cpus_choices=(0.5, 1, 2, 4, 8),
             mem_fn=None, offer_types=((16, 65536), (32, 131072), (64, 262144), (96, 393216)),
             offer_p=(0.4, 0.3, 0.2, 0.1), default_share=(100.0, 400000.0), big_share_frac=0.05,
             used_fraction=True, zipf=1.1, constraints=False, n_attr_cols=0, prio_p=(0.1, 0.8, 0.1)):
    """One pool's rank + match inputs.  Returns a dict of numpy columns plus the
    packed ABI structs (keys 'running', 'pending', 'users', 'jobs', 'offers')."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if mem_fn is None:
        mem_fn = lambda r, n: (512.0 * r.integers(1, 65, size=n)).astype(np.float64)
    owners_p = _zipf_owner(rng, n_jobs, n_users, zipf) if zipf else rng.integers(0, n_users, n_jobs).astype(np.int32)
    owners_r = _zipf_owner(rng, n_running, n_users, zipf) if zipf else rng.integers(0, n_users, n_running).astype(np.int32)
    pend = _tasks(rng, n_jobs, owners_p, False, 1_000_000, cpus_choices, mem_fn, prio_p)
    run = _tasks(rng, n_running, owners_r, True, 1, cpus_choices, mem_fn, prio_p)
    name_rank = rng.permutation(n_users).astype(np.int32)
    div_cpus = np.full(n_users, default_share[0])
    div_mem = np.full(n_users, default_share[1])
    big = rng.random(n_users) < big_share_frac
    div_cpus[big] *= 10.0
    div_mem[big] *= 10.0
    # offers
    ti = rng.choice(len(offer_types), size=n_offers, p=offer_p)
    tot_cpus = np.array([offer_types[i][0] for i in ti], np.float64)
    tot_mem = np.array([offer_types[i][1] for i in ti], np.float64)
    if used_fraction:
        frac = rng.integers(0, 8, size=n_offers) / 8.0  # grid-preserving
        run_cpus = np.floor(tot_cpus * frac * 2.0) / 2.0
        run_mem = np.floor(tot_mem * frac / 512.0) * 512.0
    else:
        run_cpus = np.zeros(n_offers)
        run_mem = np.zeros(n_offers)
    lease_cpus = tot_cpus - run_cpus
    lease_mem = tot_mem - run_mem
    run_count = np.where(run_cpus > 0, rng.integers(1, 9, size=n_offers), 0).astype(np.int32)
    host_rank = rng.permutation(n_offers).astype(np.int32)
    cols = dict(pending=pend, running=run, name_rank=name_rank, div_cpus=div_cpus, div_mem=div_mem,
                lease_cpus=lease_cpus, lease_mem=lease_mem, run_cpus=run_cpus, run_mem=run_mem,
                run_count=run_count, host_rank=host_rank)
    running = abi.make_tasks(**run)
    pending = abi.make_tasks(**pend)
    # running usage per user (generate-user-usage-map scheduler.clj:711-727)
    usage = {k: np.zeros(n_users) for k in ("count", "cpus", "mem", "gpus")}
    np.add.at(usage["count"], owners_r, 1.0)
    np.add.at(usage["cpus"], owners_r, run["cpus"])
    np.add.at(usage["mem"], owners_r, run["mem"])
    users = abi.make_users(n_users, name_rank=name_rank, div_mem=div_mem, div_cpus=div_cpus,
                           div_gpus=np.full(n_users, 1.0), usage=usage)
    jobs = abi.JobsSoA(n=n_jobs, user=owners_p, cpus=pend["cpus"], mem=pend["mem"],
                       gpus=pend["gpus"], ports=np.zeros(n_jobs, np.int32),
                       allowed=np.ones(n_jobs, np.uint8), plugin_accept=np.ones(n_jobs, np.uint8))
    offers = abi.OffersSoA(n=n_offers, hostname_id=np.arange(n_offers, dtype=np.int32),
                           name_rank=host_rank, cpus=lease_cpus, mem=lease_mem,
                           run_cpus=run_cpus, run_mem=run_mem, run_count=run_count,
                           n_attr_cols=0)
    return dict(cols=cols, running=running, pending=pending, users=users, jobs=jobs,
                offers=offers, n_users=n_users)


def gen_c1(seed=1):
    """BASELINE config #1: simulator-style trace, 1k pending x 100 offers, 1 pool
    (zz_simulator.clj:720-744: cpus in {1,2,3}, mem in [2000,2999], 4 users,
    hosts 20 cpus / 20000 MB, share cpus 2 mem 2000)."""
    return gen_pool(seed, 1000, 100, 4, 0, cpus_choices=(1, 2, 3),
                    mem_fn=lambda r, n: r.integers(2000, 3000, size=n).astype(np.float64),
                    offer_types=((20, 20000),), offer_p=(1.0,), default_share=(2.0, 2000.0),
                    big_share_frac=0.0, used_fraction=False, zipf=None, prio_p=(0.0, 1.0, 0.0))


def gen_c2(seed=2, n_jobs=100_000, n_offers=5_000, n_users=1_000, n_running=20_000):
    """BASELINE config #2: 100k jobs x 5k offers, cpu+mem fit only, 1 pool."""
    return gen_pool(seed, n_jobs, n_offers, n_users, n_running)


def match_params(num_considerable, enforce_rate_limit=0, host_lifetime_mins=0, reuse_resident=0):
    return abi.MatchParams(int(num_considerable), int(enforce_rate_limit), int(host_lifetime_mins),
                           0, 1.0, int(reuse_resident), 0)
