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


def gen_c2(seed=2, n_jobs=100_000, n_offers=5_000, n_users=1_000, n_running=20_000, offer_scale=1):
    """BASELINE config #2: 100k jobs x 5k offers, cpu+mem fit only, 1 pool.  offer_scale > 1 is the
    NON-SATURATING variant bench.py reports beside it: the same jobs against offers with that many
    times the capacity (6 => every job is placeable), which exposes the matcher's dependency-chain
    bound instead of the cluster filling up after a quarter of the queue."""
    types = tuple((c * offer_scale, m * offer_scale) for c, m in ((16, 65536), (32, 131072), (64, 262144), (96, 393216)))
    return gen_pool(seed, n_jobs, n_offers, n_users, n_running, offer_types=types)


def match_params(num_considerable, enforce_rate_limit=0, host_lifetime_mins=0, reuse_resident=0, max_ctas=0):
    return abi.MatchParams(int(num_considerable), int(enforce_rate_limit), int(host_lifetime_mins),
                           0, 1.0, int(reuse_resident), int(max_ctas))


def add_constraints(t, seed, *, n_attr_cols=8, attr_card=(3, 6, 24, 2, 2, 4, 5, 7), frac_attr=0.30,
                    frac_gpu_nodes=0.05, frac_gpu_jobs=0.02, frac_port_jobs=0.03,
                    frac_port_nodes=0.25, frac_group_jobs=0.05, group_size=(4, 24),
                    frac_novel=0.10, max_tasks=110, frac_k8s=0.75, frac_reserved=0.01,
                    frac_est=0.05, frac_ckpt=0.02, frac_disk=0.05, n_models=2, n_locations=3,
                    host_lifetime_mins=1440, n_running_cotasks=3):
    """Adds SURVEY §8d config-#3 style constraint columns to a gen_pool() trace, in
    place: 8 attribute columns, GPU nodes/jobs (models x counts), ports on the
    Mesos-type nodes, unique/balanced/attribute-equals groups, novel-host lists,
    max-tasks-per-host, reservations, estimated completion, checkpoint locality,
    disk.  Returns the cook_groups struct (or None)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    jobs, offers = t["jobs"], t["offers"]
    J, O = jobs.n, offers.n
    # ---- offers
    is_k8s = (rng.random(O) < frac_k8s).astype(np.uint8)
    attr = np.zeros((n_attr_cols, O), np.int32)
    for c in range(n_attr_cols):
        attr[c] = rng.integers(1, attr_card[c % len(attr_card)] + 1, O)
        attr[c][rng.random(O) < 0.05] = 0  # attribute absent on some hosts
    gpu_node = (rng.random(O) < frac_gpu_nodes) & (is_k8s == 1)
    gpu_lists_m = [[int(rng.integers(0, n_models))] if g else [] for g in gpu_node]
    gpu_lists_c = [[float(rng.choice([1, 2, 4, 8]))] if g else [] for g in gpu_node]
    gpu_off, gpu_model = abi.csr(gpu_lists_m)
    _, gpu_count = abi.csr(gpu_lists_c, np.float64)
    port_node = (rng.random(O) < frac_port_nodes) & (is_k8s == 0)
    pb = [[31000, 31500] if p else [] for p in port_node]
    pe = [[31009, 31504] if p else [] for p in port_node]
    port_off, port_begin = abi.csr(pb)
    _, port_end = abi.csr(pe)
    disk_t = [[0, 1] if k else [] for k in is_k8s]
    disk_s = [[float(rng.integers(10, 200) * 1024), float(rng.integers(0, 50) * 1024)] if k else [] for k in is_k8s]
    disk_off, disk_type = abi.csr(disk_t)
    _, disk_space = abi.csr(disk_s, np.float64)
    num_tasks = rng.integers(0, max_tasks + 5, O).astype(np.int32)
    max_t = np.where(rng.random(O) < 0.9, max_tasks, -1).astype(np.int32)
    location = rng.integers(0, n_locations, O).astype(np.int32)
    host_start = np.where(rng.random(O) < 0.7, 1_600_000_000 + rng.integers(0, 86400, O), -1).astype(np.int64)
    reserved = (rng.random(O) < frac_reserved).astype(np.uint8)
    hostname_id = offers.col("hostname_id")
    cols = {n: offers.col(n) for n in ("hostname_id", "name_rank", "cpus", "mem", "run_cpus", "run_mem", "run_count")}
    t["offers"] = abi.OffersSoA(n=O, **cols, port_off=port_off, port_begin=port_begin, port_end=port_end,
                                is_k8s=is_k8s, location=location, gpu_off=gpu_off, gpu_model=gpu_model,
                                gpu_count=gpu_count, disk_off=disk_off, disk_type=disk_type,
                                disk_space=disk_space, max_tasks=max_t, num_tasks=num_tasks,
                                host_start_time=host_start, n_attr_cols=n_attr_cols,
                                attr=attr.reshape(-1), reserved=reserved)
    # ---- jobs
    gpus = np.where(rng.random(J) < frac_gpu_jobs, rng.choice([1.0, 2.0, 4.0, 8.0], J), 0.0)
    gmodel = np.where(gpus > 0, rng.integers(0, n_models, J), -1).astype(np.int32)
    ports = np.where(rng.random(J) < frac_port_jobs, rng.integers(1, 3, J), 0).astype(np.int32)
    attr_l_c, attr_l_v = [], []
    for j in range(J):
        if rng.random() < frac_attr:
            k = int(rng.integers(1, 3))
            cs = rng.choice(n_attr_cols, size=k, replace=False)
            attr_l_c.append([int(c) for c in cs])
            attr_l_v.append([int(rng.integers(1, attr_card[int(c) % len(attr_card)] + 1)) if rng.random() < 0.97 else -1 for c in cs])
        else:
            attr_l_c.append([])
            attr_l_v.append([])
    attr_off, attr_col = abi.csr(attr_l_c)
    _, attr_val = abi.csr(attr_l_v)
    novel = [[int(x) for x in rng.choice(hostname_id, size=int(rng.integers(1, 4)), replace=False)]
             if rng.random() < frac_novel else [] for _ in range(J)]
    novel_off, novel_host = abi.csr(novel)
    now_ms = 1_600_050_000_000
    est = np.where(rng.random(J) < frac_est, now_ms + rng.integers(1, 48 * 3600_000, J), -1).astype(np.int64)
    ckpt = np.where(rng.random(J) < frac_ckpt, rng.integers(0, n_locations, J), -1).astype(np.int32)
    res_hosts = hostname_id[reserved == 1]
    reserved_host = np.full(J, -1, np.int32)
    if len(res_hosts):
        pick = rng.random(J) < 0.01
        reserved_host[pick] = rng.choice(res_hosts, size=int(pick.sum()))
    disk_req = np.where(rng.random(J) < frac_disk, rng.integers(1, 64, J) * 1024.0, -1.0)
    disk_typ = rng.integers(0, 2, J).astype(np.int32)
    # groups
    n_gj = int(J * frac_group_jobs)
    members = rng.permutation(J)[:n_gj]
    grp_lists = [[] for _ in range(J)]
    kinds, acols, mins, cot_h, cot_a = [], [], [], [], []
    g = 0
    i = 0
    while i < n_gj:
        sz = int(rng.integers(group_size[0], group_size[1] + 1))
        r = rng.random()
        kind = abi.GROUP_UNIQUE if r < 0.6 else (abi.GROUP_BALANCED if r < 0.85 else abi.GROUP_ATTR_EQUALS)
        col = int(rng.integers(0, min(3, n_attr_cols)))
        for j in members[i:i + sz]:
            grp_lists[j].append(g)
        kinds.append(kind); acols.append(col); mins.append(int(rng.integers(1, 4)))
        nc = int(rng.integers(0, n_running_cotasks + 1))
        hs = [int(x) for x in rng.choice(hostname_id, size=nc, replace=False)] if nc else []
        cot_h.append(hs)
        cot_a.append([int(attr[col][np.where(hostname_id == h)[0][0]]) for h in hs])
        g += 1
        i += sz
    group_off, group_idx = abi.csr(grp_lists)
    base = {n: jobs.col(n) for n in ("user", "cpus", "mem", "allowed", "plugin_accept")}
    t["jobs"] = abi.JobsSoA(n=J, **base, gpus=gpus, ports=ports, novel_off=novel_off, novel_host=novel_host,
                            gpu_model=gmodel, disk_request=disk_req, disk_type=disk_typ,
                            attr_off=attr_off, attr_col=attr_col, attr_val=attr_val, est_end_ms=est,
                            ckpt_location=ckpt, reserved_host=reserved_host,
                            group_off=group_off, group_idx=group_idx)
    t["host_lifetime_mins"] = host_lifetime_mins
    if g == 0:
        t["groups"] = None
        return None
    cot_off, cot_host = abi.csr(cot_h)
    _, cot_attr = abi.csr(cot_a)
    t["groups"] = abi.Groups(n_groups=g, kind=np.array(kinds, np.int32), attr_col=np.array(acols, np.int32),
                             minimum=np.array(mins, np.int32), cot_off=cot_off,
                             cot_hostname_id=cot_host, cot_attr_val=cot_attr)
    return t["groups"]


def gen_c3_pool(seed=3, n_jobs=20_000, n_offers=1_000, n_users=200, n_running=4_000, **kw):
    """One pool of BASELINE config #3's shape (scaled by the caller): host-placement
    + gpu/ports constraints, groups, novel hosts, max-tasks-per-host."""
    t = gen_pool(seed, n_jobs, n_offers, n_users, n_running)
    add_constraints(t, seed + 1000, **kw)
    return t


def gen_rebalance(seed, n_running, n_pending, n_hosts, n_users, *, max_preemption=64,
                  min_dru_diff=0.5, safe_dru_threshold=1.0, constraints=True):
    """BASELINE config #4 style rebalancer input: running tasks skewed so that some
    users sit far above their share, a few hosts with spare resources, pending jobs
    of under-served users; optional host attributes / novel-host / group constraints."""
    rng = np.random.Generator(np.random.PCG64(seed))
    owners = _zipf_owner(rng, n_running, n_users, 1.3)
    run = _tasks(rng, n_running, owners, True, 1, (0.5, 1, 2, 4), lambda r, n: (512.0 * r.integers(1, 33, size=n)))
    host = rng.integers(0, n_hosts, n_running).astype(np.int32)
    running = abi.RunningSoA(t=abi.make_tasks(**run), host=host)
    pu = rng.integers(n_users // 2, n_users, n_pending).astype(np.int32)  # tail users are under-served
    pend = _tasks(rng, n_pending, pu, False, 10_000_000, (1, 2, 4, 8), lambda r, n: (1024.0 * r.integers(1, 33, size=n)))
    name_rank = rng.permutation(n_users).astype(np.int32)
    tot_c = np.zeros(n_users); tot_m = np.zeros(n_users)
    np.add.at(tot_c, owners, run["cpus"]); np.add.at(tot_m, owners, run["mem"])
    share_c = np.maximum(8.0, np.round(np.median(tot_c[tot_c > 0]) if (tot_c > 0).any() else 8.0))
    share_m = np.maximum(4096.0, np.round(np.median(tot_m[tot_m > 0]) if (tot_m > 0).any() else 4096.0))
    quota = {"count": np.where(rng.random(n_users) < 0.2, rng.integers(5, 200, n_users), 1e12).astype(float)}
    users = abi.make_users(n_users, name_rank=name_rank, div_mem=np.full(n_users, share_m),
                           div_cpus=np.full(n_users, share_c), div_gpus=np.full(n_users, 1.0), quota=quota)
    has_spare = (rng.random(n_hosts) < 0.05).astype(np.uint8)
    hostname_id = np.arange(n_hosts, dtype=np.int32)
    hkw = dict(n=n_hosts, hostname_id=hostname_id, name_rank=rng.permutation(n_hosts).astype(np.int32),
               has_spare=has_spare, spare_cpus=np.where(has_spare, rng.integers(0, 9, n_hosts), 0).astype(float),
               spare_mem=np.where(has_spare, 1024.0 * rng.integers(0, 17, n_hosts), 0.0),
               spare_gpus=np.zeros(n_hosts), n_attr_cols=0)
    jkw = dict(n=n_pending, user=pu, cpus=pend["cpus"], mem=pend["mem"], gpus=np.zeros(n_pending))
    groups = None
    if constraints:
        ncol = 3
        attr = rng.integers(0, 4, (ncol, n_hosts)).astype(np.int32)
        hkw.update(n_attr_cols=ncol, attr=attr.reshape(-1), is_k8s=(rng.random(n_hosts) < 0.5).astype(np.uint8),
                   location=rng.integers(0, 2, n_hosts).astype(np.int32))
        al_c = [[int(rng.integers(0, ncol))] if rng.random() < 0.3 else [] for _ in range(n_pending)]
        al_v = [[int(rng.integers(1, 4))] if c else [] for c in al_c]
        attr_off, attr_col = abi.csr(al_c)
        _, attr_val = abi.csr(al_v)
        novel = [[int(x) for x in rng.choice(hostname_id, size=2, replace=False)] if rng.random() < 0.3 else []
                 for _ in range(n_pending)]
        novel_off, novel_host = abi.csr(novel)
        ng = max(1, n_pending // 8)
        gl = [[int(rng.integers(0, ng))] if rng.random() < 0.3 else [] for _ in range(n_pending)]
        group_off, group_idx = abi.csr(gl)
        kinds = rng.integers(0, 3, ng).astype(np.int32)
        acol = rng.integers(0, ncol, ng).astype(np.int32)
        cot_h = [[int(x) for x in rng.choice(hostname_id, size=int(rng.integers(0, 3)), replace=False)] for _ in range(ng)]
        cot_a = [[int(attr[acol[g]][h]) for h in cot_h[g]] for g in range(ng)]
        cot_off, cot_host = abi.csr(cot_h)
        _, cot_attr = abi.csr(cot_a)
        groups = abi.Groups(n_groups=ng, kind=kinds, attr_col=acol, minimum=rng.integers(1, 4, ng).astype(np.int32),
                            cot_off=cot_off, cot_hostname_id=cot_host, cot_attr_val=cot_attr)
        jkw.update(attr_off=attr_off, attr_col=attr_col, attr_val=attr_val, novel_off=novel_off,
                   novel_host=novel_host, group_off=group_off, group_idx=group_idx,
                   ckpt_location=np.where(rng.random(n_pending) < 0.1, rng.integers(0, 2, n_pending), -1).astype(np.int32))
    return dict(running=running, pending=abi.JobsSoA(**jkw), pending_job_id=pend["job_id"],
                pending_priority=pend["priority"], hosts=abi.HostTable(**hkw), users=users, groups=groups,
                params=abi.RebalanceParams(max_preemption, min_dru_diff, safe_dru_threshold, 0))


# ---------------------------------------------------------------------------------------------
# BASELINE configs #3-#5 at their stated sizes.  A config is a list of
# independent pools (Cook's shard axis, scheduler.clj:2488-2517); every pool carries its own
# rank + match (+ rebalance) inputs.  Pool sizes follow the 40/30/20/10 % split (C3/C4) and a
# Zipf split over 16 pools (C5); nodes split in the same proportion.
C3_SPLIT = (0.4, 0.3, 0.2, 0.1)


def pool_sizes(config):
    """[(jobs, offers, users, running)] per pool for 'c3' / 'c4' / 'c5'."""
    if config in ("c3", "c4"):
        run = 200_000 if config == "c3" else 400_000
        return [(int(1_000_000 * f), int(20_000 * f), int(5_000 * f), int(run * f)) for f in C3_SPLIT]
    if config == "c5":
        w = 1.0 / np.arange(1, 17) ** 0.5
        w /= w.sum()
        jobs = np.floor(10_000_000 * w / 1000).astype(int) * 1000
        jobs[0] += 10_000_000 - jobs.sum()
        return [(int(j), max(500, int(round(100_000 * j / 1e7))), max(100, int(round(20_000 * j / 1e7))),
                 int(round(2_000_000 * j / 1e7))) for j in jobs]
    raise ValueError(config)


def gen_config_pool(config, p, scale=1.0):
    """Pool p of BASELINE config 'c3' | 'c4' | 'c5' (scale < 1 shrinks every dimension: the
    CPU-runnable miniature of the same shape).  Returns the gen_c3_pool() dict plus 'quota'
    tables that bind for ~10 % of the users and, for c4/c5, a 'rebalance'
    entry (gen_rebalance dict sized to the pool)."""
    nj, no, nu, nr = pool_sizes(config)[p]
    nj, no, nu, nr = (max(8, int(nj * scale)), max(4, int(no * scale)), max(2, int(nu * scale)),
                      max(1, int(nr * scale)))
    seed = {"c3": 3000, "c4": 4000, "c5": 5000}[config] + 17 * p
    t = gen_pool(seed, nj, no, nu, nr)
    add_constraints_fast(t, seed + 1)
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    # 10 % of the users carry count / cpu quotas that bind
    bind = rng.random(nu) < 0.10
    quota = {"count": np.where(bind, rng.integers(20, 400, nu), 1e12).astype(float),
             "cpus": np.where(bind, rng.integers(50, 2000, nu), 1e12).astype(float)}
    usage = {k: t["users"].col("usage_" + k) for k in ("count", "cpus", "mem", "gpus")}
    t["users"] = abi.make_users(nu, name_rank=t["users"].col("name_rank"), div_mem=t["users"].col("div_mem"),
                                div_cpus=t["users"].col("div_cpus"), div_gpus=np.full(nu, 1.0),
                                quota=quota, usage=usage)
    if config in ("c4", "c5"):
        t["rebalance"] = gen_rebalance(seed + 3, nr, max(8, int(400 * (nj / 250_000))), no, nu,
                                       max_preemption=128, min_dru_diff=0.5, safe_dru_threshold=1.0)
    t["config"] = config
    t["pool"] = p
    return t


def add_constraints_fast(t, seed, *, n_attr_cols=8, attr_card=(3, 6, 24, 2, 2, 4, 5, 7), frac_attr=0.30,
                         frac_gpu_nodes=0.05, frac_gpu_jobs=0.02, frac_port_jobs=0.03,
                         frac_port_nodes=0.25, frac_group_jobs=0.05, group_size=(4, 24),
                         frac_novel=0.10, max_tasks=110, frac_k8s=0.75, frac_reserved=0.01,
                         frac_est=0.05, frac_ckpt=0.02, frac_disk=0.05, n_models=2, n_locations=3,
                         host_lifetime_mins=1440, n_running_cotasks=3):
    """Vectorised twin of add_constraints() (same column semantics and fractions, different random
    stream) for the million-job pools: no per-job Python loop."""
    rng = np.random.Generator(np.random.PCG64(seed))
    jobs, offers = t["jobs"], t["offers"]
    J, O = jobs.n, offers.n
    card = np.array([attr_card[c % len(attr_card)] for c in range(n_attr_cols)])
    # ---- offers
    is_k8s = (rng.random(O) < frac_k8s).astype(np.uint8)
    attr = (rng.integers(0, 1 << 30, (n_attr_cols, O)) % card[:, None] + 1).astype(np.int32)
    attr[rng.random((n_attr_cols, O)) < 0.05] = 0

    def csr_one(mask, vals, dtype=np.int32):   # <= 1 entry per row
        off = np.zeros(len(mask) + 1, np.int32)
        off[1:] = np.cumsum(mask)
        return off, np.ascontiguousarray(vals[mask], dtype)

    gpu_node = (rng.random(O) < frac_gpu_nodes) & (is_k8s == 1)
    gpu_off, gpu_model = csr_one(gpu_node, rng.integers(0, n_models, O))
    _, gpu_count = csr_one(gpu_node, rng.choice(np.array([1.0, 2.0, 4.0, 8.0]), O), np.float64)
    port_node = (rng.random(O) < frac_port_nodes) & (is_k8s == 0)
    port_off = np.zeros(O + 1, np.int32)
    port_off[1:] = np.cumsum(2 * port_node)
    npn = int(port_node.sum())
    port_begin = np.tile(np.array([31000, 31500], np.int32), npn)
    port_end = np.tile(np.array([31009, 31504], np.int32), npn)
    k8 = is_k8s == 1
    disk_off = np.zeros(O + 1, np.int32)
    disk_off[1:] = np.cumsum(2 * k8)
    nk = int(k8.sum())
    disk_type = np.tile(np.array([0, 1], np.int32), nk)
    disk_space = np.stack([rng.integers(10, 200, nk) * 1024.0, rng.integers(0, 50, nk) * 1024.0], 1).reshape(-1)
    num_tasks = rng.integers(0, max_tasks + 5, O).astype(np.int32)
    max_t = np.where(rng.random(O) < 0.9, max_tasks, -1).astype(np.int32)
    location = rng.integers(0, n_locations, O).astype(np.int32)
    host_start = np.where(rng.random(O) < 0.7, 1_600_000_000 + rng.integers(0, 86400, O), -1).astype(np.int64)
    reserved = (rng.random(O) < frac_reserved).astype(np.uint8)
    hostname_id = offers.col("hostname_id")
    cols = {n: offers.col(n) for n in ("hostname_id", "name_rank", "cpus", "mem", "run_cpus", "run_mem", "run_count")}
    t["offers"] = abi.OffersSoA(n=O, **cols, port_off=port_off, port_begin=port_begin, port_end=port_end,
                                is_k8s=is_k8s, location=location, gpu_off=gpu_off, gpu_model=gpu_model,
                                gpu_count=gpu_count, disk_off=disk_off, disk_type=disk_type,
                                disk_space=disk_space, max_tasks=max_t, num_tasks=num_tasks,
                                host_start_time=host_start, n_attr_cols=n_attr_cols,
                                attr=attr.reshape(-1), reserved=reserved)
    # ---- jobs
    gpus = np.where(rng.random(J) < frac_gpu_jobs, rng.choice(np.array([1.0, 2.0, 4.0, 8.0]), J), 0.0)
    gmodel = np.where(gpus > 0, rng.integers(0, n_models, J), -1).astype(np.int32)
    ports = np.where(rng.random(J) < frac_port_jobs, rng.integers(1, 3, J), 0).astype(np.int32)
    # user-defined EQUALS: 0, 1 or 2 (distinct columns) per job
    na = np.where(rng.random(J) < frac_attr, rng.integers(1, 3, J), 0).astype(np.int32)
    attr_off = np.zeros(J + 1, np.int32)
    attr_off[1:] = np.cumsum(na)
    owner = np.repeat(np.arange(J), na)
    first = np.concatenate([[True], owner[1:] != owner[:-1]]) if len(owner) else np.zeros(0, bool)
    c1 = rng.integers(0, n_attr_cols, len(owner))
    c2 = (c1 + rng.integers(1, n_attr_cols, len(owner))) % n_attr_cols     # differs from the first column
    prev = np.roll(c1, 1)
    attr_col = np.where(first, c1, np.where(c2 == prev, (c2 + 1) % n_attr_cols, c2)).astype(np.int32)
    attr_col = np.where(~first & (attr_col == prev), (attr_col + 1) % n_attr_cols, attr_col).astype(np.int32)
    attr_val = (rng.integers(0, 1 << 30, len(owner)) % card[attr_col] + 1).astype(np.int32)
    attr_val[rng.random(len(owner)) >= 0.97] = -1
    # novel-host: 1..3 distinct previous hosts for 10 % of the jobs
    nn = np.where(rng.random(J) < frac_novel, rng.integers(1, 4, J), 0).astype(np.int32)
    nn = np.minimum(nn, O)
    novel_off = np.zeros(J + 1, np.int32)
    novel_off[1:] = np.cumsum(nn)
    ownn = np.repeat(np.arange(J), nn)
    pos = np.arange(len(ownn)) - novel_off[ownn]               # 0..2 within the job
    base = rng.integers(0, O, J)
    step = rng.integers(1, max(2, O // 3), J)
    novel_host = hostname_id[(base[ownn] + pos * step[ownn]) % O].astype(np.int32) if len(ownn) else np.zeros(0, np.int32)
    if O < 8:   # tiny tables: fall back to exact de-duplication
        keep = np.ones(len(ownn), bool)
        for j in np.unique(ownn):
            idx = np.where(ownn == j)[0]
            _, f = np.unique(novel_host[idx], return_index=True)
            keep[idx] = False
            keep[idx[f]] = True
        novel_host = novel_host[keep]
        nn = np.bincount(ownn[keep], minlength=J).astype(np.int32)
        novel_off[1:] = np.cumsum(nn)
    now_ms = 1_600_050_000_000
    est = np.where(rng.random(J) < frac_est, now_ms + rng.integers(1, 48 * 3600_000, J), -1).astype(np.int64)
    ckpt = np.where(rng.random(J) < frac_ckpt, rng.integers(0, n_locations, J), -1).astype(np.int32)
    res_hosts = hostname_id[reserved == 1]
    reserved_host = np.full(J, -1, np.int32)
    if len(res_hosts):
        pick = rng.random(J) < 0.01
        reserved_host[pick] = rng.choice(res_hosts, size=int(pick.sum()))
    disk_req = np.where(rng.random(J) < frac_disk, rng.integers(1, 64, J) * 1024.0, -1.0)
    disk_typ = rng.integers(0, 2, J).astype(np.int32)
    # groups: consecutive runs of a random permutation, sizes in group_size
    n_gj = int(J * frac_group_jobs)
    members = rng.permutation(J)[:n_gj]
    sizes = []
    left = n_gj
    while left > 0:
        sz = min(left, int(rng.integers(group_size[0], group_size[1] + 1)))
        sizes.append(sz)
        left -= sz
    G = len(sizes)
    group_of = np.full(J, -1, np.int32)
    if G:
        group_of[members] = np.repeat(np.arange(G, dtype=np.int32), sizes)
    has_g = group_of >= 0
    group_off = np.zeros(J + 1, np.int32)
    group_off[1:] = np.cumsum(has_g)
    group_idx = group_of[has_g]
    base_j = {n: jobs.col(n) for n in ("user", "cpus", "mem", "allowed", "plugin_accept")}
    t["jobs"] = abi.JobsSoA(n=J, **base_j, gpus=gpus, ports=ports, novel_off=novel_off, novel_host=novel_host,
                            gpu_model=gmodel, disk_request=disk_req, disk_type=disk_typ,
                            attr_off=attr_off, attr_col=attr_col, attr_val=attr_val, est_end_ms=est,
                            ckpt_location=ckpt, reserved_host=reserved_host,
                            group_off=group_off, group_idx=group_idx)
    t["host_lifetime_mins"] = host_lifetime_mins
    if G == 0:
        t["groups"] = None
        return None
    r = rng.random(G)
    kinds = np.where(r < 0.6, abi.GROUP_UNIQUE, np.where(r < 0.85, abi.GROUP_BALANCED, abi.GROUP_ATTR_EQUALS)).astype(np.int32)
    acols = rng.integers(0, min(3, n_attr_cols), G).astype(np.int32)
    mins = rng.integers(1, 4, G).astype(np.int32)
    ncot = np.minimum(rng.integers(0, n_running_cotasks + 1, G), O).astype(np.int32)
    cot_off = np.zeros(G + 1, np.int32)
    cot_off[1:] = np.cumsum(ncot)
    owg = np.repeat(np.arange(G), ncot)
    posg = np.arange(len(owg)) - cot_off[owg]
    bg = rng.integers(0, O, G)
    sg = rng.integers(1, max(2, O // 4), G)
    cot_idx = (bg[owg] + posg * sg[owg]) % O if len(owg) else np.zeros(0, np.int64)
    cot_host = hostname_id[cot_idx].astype(np.int32) if len(owg) else np.zeros(0, np.int32)
    cot_attr = attr[acols[owg], cot_idx].astype(np.int32) if len(owg) else np.zeros(0, np.int32)
    t["groups"] = abi.Groups(n_groups=G, kind=kinds, attr_col=acols, minimum=mins, cot_off=cot_off,
                             cot_hostname_id=cot_host, cot_attr_val=cot_attr)
    return t["groups"]


def dump_for_jvm(config, out_dir):
    """Writes the C2 (or one pool of c3-c5) trace as little-endian column files + manifest.edn for
    bench/jvm/run_reference.clj (the JVM harness that times the real Clojure + Fenzo path)."""
    import os
    os.makedirs(out_dir, exist_ok=True)
    t = gen_c2() if config == "c2" else gen_config_pool(config, 0)
    cols = {"pending_user": t["pending"].col("user"), "pending_cpus": t["pending"].col("cpus"),
            "pending_mem": t["pending"].col("mem"), "pending_priority": t["pending"].col("priority"),
            "running_user": t["running"].col("user"), "running_cpus": t["running"].col("cpus"),
            "running_mem": t["running"].col("mem"), "offer_cpus": t["offers"].col("cpus"),
            "offer_mem": t["offers"].col("mem"), "offer_name_rank": t["offers"].col("name_rank")}
    for n, a in cols.items():
        np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tofile(os.path.join(out_dir, n + ".bin"))
    with open(os.path.join(out_dir, "manifest.edn"), "w") as f:
        f.write("{:config \"%s\" :jobs %d :offers %d :users %d :running %d :num-considerable %d}\n" %
                (config, t["jobs"].n, t["offers"].n, t["users"].n_users, t["running"].n, t["jobs"].n))


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "--dump":
        dump_for_jvm(sys.argv[2], sys.argv[3])
    else:
        print("usage: python -m cook_b200.traces --dump c2|c3|c4|c5 DIR")
