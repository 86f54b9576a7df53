"""-m gpu parity tests: CUDA path (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from cook_b200 import abi, traces

pytestmark = pytest.mark.gpu


def _same_dru(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(b, nan=-1.0))


@pytest.mark.parametrize("seed,nj,no,nu,nr", [
    (11, 50, 7, 3, 10), (12, 1000, 100, 4, 0), (13, 5000, 333, 50, 2000),
    (14, 20000, 1000, 200, 5000), (15, 4096, 2048, 7, 4096), (16, 1, 1, 1, 0),
    (17, 3000, 33, 1000, 100),
])
def test_rank_and_match_parity_random(gpu, oracle, seed, nj, no, nu, nr):
    t = traces.gen_pool(seed, nj, no, nu, nr)
    rg = gpu.rank(t["running"], t["pending"], t["users"])
    ro = oracle.rank(t["running"], t["pending"], t["users"])
    assert np.array_equal(rg["order"], ro["order"])
    assert np.array_equal(rg["ranked"], ro["ranked"])
    assert _same_dru(rg["dru"], ro["dru"])
    prm = traces.match_params(nj)
    mg = gpu.match(rg["ranked"], t["jobs"], t["offers"], t["users"], prm)
    mo = oracle.match(ro["ranked"], t["jobs"], t["offers"], t["users"], prm)
    assert np.array_equal(mg["considerable"], mo["considerable"])
    assert np.array_equal(mg["assign"], mo["assign"])
    for k in ("n_considerable", "n_matched", "head_matched", "n_offers_used", "evals"):
        assert mg["stats"][k] == mo["stats"][k], k


def test_c1_simulator_trace(gpu, oracle):
    t = traces.gen_c1()
    rg = gpu.rank(t["running"], t["pending"], t["users"])
    ro = oracle.rank(t["running"], t["pending"], t["users"])
    assert np.array_equal(rg["ranked"], ro["ranked"])
    prm = traces.match_params(1000)
    mg = gpu.match(rg["ranked"], t["jobs"], t["offers"], t["users"], prm)
    mo = oracle.match(ro["ranked"], t["jobs"], t["offers"], t["users"], prm)
    assert np.array_equal(mg["assign"], mo["assign"])


def test_c2_full_size_bit_identical(gpu, oracle):
    """BASELINE config #2: 100k jobs x 5k offers — assignments bit-identical."""
    t = traces.gen_c2()
    rg = gpu.rank(t["running"], t["pending"], t["users"])
    ro = oracle.rank(t["running"], t["pending"], t["users"])
    assert np.array_equal(rg["ranked"], ro["ranked"])
    assert _same_dru(rg["dru"], ro["dru"])
    prm = traces.match_params(100_000)
    mg = gpu.match(rg["ranked"], t["jobs"], t["offers"], t["users"], prm)
    mo = oracle.match(ro["ranked"], t["jobs"], t["offers"], t["users"], prm)
    assert np.array_equal(mg["considerable"], mo["considerable"])
    assert np.array_equal(mg["assign"], mo["assign"])
    assert mg["stats"]["evals"] == 500_000_000


def test_quota_and_rate_limit_considerable(gpu, oracle):
    rng = np.random.default_rng(5)
    t = traces.gen_pool(21, 4000, 100, 30, 1000)
    nu = 30
    quota = {"count": rng.integers(5, 60, nu).astype(float), "cpus": rng.integers(10, 200, nu).astype(float),
             "mem": rng.integers(20000, 900000, nu).astype(float), "gpus": np.full(nu, 1e9)}
    usage = {k: t["users"].col("usage_" + k) for k in ("count", "cpus", "mem", "gpus")}
    users = abi.make_users(nu, name_rank=t["users"].col("name_rank"), div_mem=t["users"].col("div_mem"),
                           div_cpus=t["users"].col("div_cpus"), quota=quota, usage=usage,
                           tokens=rng.integers(0, 40, nu).astype(np.int32))
    pq = abi.make_pool_quota({"count": 1500, "cpus": 4000, "mem": 2.0e7, "gpus": 1e9})
    rank_prm = abi.RankParams(5, 1, 30000.0, 6.0)
    rg = gpu.rank(t["running"], t["pending"], users, pool_quota=pq, params=rank_prm)
    ro = oracle.rank(t["running"], t["pending"], users, pool_quota=pq, params=rank_prm)
    assert np.array_equal(rg["ranked"], ro["ranked"])
    assert _same_dru(rg["dru"], ro["dru"])
    for enforce in (0, 1):
        prm = traces.match_params(700, enforce_rate_limit=enforce)
        mg = gpu.match(ro["ranked"], t["jobs"], t["offers"], users, prm, pool_quota=pq)
        mo = oracle.match(ro["ranked"], t["jobs"], t["offers"], users, prm, pool_quota=pq)
        assert np.array_equal(mg["considerable"], mo["considerable"])
        assert np.array_equal(mg["assign"], mo["assign"])


def test_resident_inputs_same_result(gpu, oracle):
    """reuse_resident skips the upload stage; results must not change, and asking
    for it without resident inputs must fail loudly."""
    from cook_b200.engine import CookError
    t = traces.gen_pool(31, 6000, 400, 40, 1500)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(6000))
    m1 = gpu.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(6000))
    assert m1["stats"]["h2d_bytes"] > 0
    for _ in range(2):
        m2 = gpu.match(ranked, t["jobs"], t["offers"], t["users"],
                       traces.match_params(6000, reuse_resident=1))
        assert m2["stats"]["h2d_bytes"] == 0
        assert np.array_equal(m2["assign"], mo["assign"])
    assert np.array_equal(m1["assign"], mo["assign"])
    with pytest.raises(CookError):
        gpu.match(ranked[:100], t["jobs"], t["offers"], t["users"],
                  traces.match_params(6000, reuse_resident=1))


@pytest.mark.parametrize("seed,nj,no,nu,nr,kw", [
    (41, 5000, 300, 50, 1000, {}),
    (42, 8000, 1000, 100, 2000, {"frac_group_jobs": 0.0}),
    (43, 3000, 97, 20, 500, {"frac_group_jobs": 0.3, "group_size": (2, 9), "frac_gpu_jobs": 0.2,
                             "frac_gpu_nodes": 0.4, "frac_port_jobs": 0.3}),
    (44, 2000, 64, 10, 100, {"frac_k8s": 0.0, "frac_port_nodes": 0.9, "frac_port_jobs": 0.5,
                             "max_tasks": 3}),
])
def test_constraint_kernel_parity(gpu, oracle, seed, nj, no, nu, nr, kw):
    """Config-#3 style constraints (attributes, gpu, ports, groups, novel hosts,
    max-tasks, reservation, est-completion, checkpoint, disk): assignments, ports
    and considerable sets bit-identical to the oracle."""
    t = traces.gen_c3_pool(seed, nj, no, nu, nr, **kw)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"])
    mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    mg = gpu.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    assert np.array_equal(mg["considerable"], mo["considerable"])
    assert np.array_equal(mg["assign"], mo["assign"])
    assert np.array_equal(mg["ports"], mo["ports"])
    assert np.array_equal(mg["fail"] == 0, mo["fail"] == 0)
    assert mg["stats"]["n_matched"] == mo["stats"]["n_matched"] > 0


def test_rank_golden_on_gpu():
    """The reference's own ranking known answers (K1-K8: DRU values, merged order, quota
    filters, GPU-mode shares) through the CUDA path."""
    import json
    import os
    from cook_b200.engine import GpuEngine
    from golden_util import check_rank_case
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "rank_golden.json")) as f:
        cases = json.load(f)
    engines = {}

    def factory(mode):
        if mode not in engines:
            engines[mode] = GpuEngine(pool_name="golden-%d" % mode, dru_mode=mode)
        return engines[mode]
    try:
        for case in cases:
            check_rank_case(case, factory)
    finally:
        for e in engines.values():
            e.close()


def test_rebalance_golden_on_gpu(gpu):
    """The reference's own rebalancer known answers (K19, K22) through the CUDA path."""
    import json
    import os
    from golden_util import check_rebalance_case
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "rebalance_golden.json")) as f:
        cases = json.load(f)
    for case in cases:
        check_rebalance_case(case, gpu)


@pytest.mark.parametrize("seed,nr,np_,nh,nu", [(51, 2000, 64, 100, 20), (52, 20000, 128, 500, 100),
                                               (53, 500, 40, 10, 5)])
def test_rebalance_parity_random(gpu, oracle, seed, nr, np_, nh, nu):
    t = traces.gen_rebalance(seed, nr, np_, nh, nu)
    args = (t["running"], t["pending"], t["pending_job_id"], t["pending_priority"], t["hosts"],
            t["users"], t["params"])
    do = oracle.rebalance(*args, groups=t["groups"])
    dg = gpu.rebalance(*args, groups=t["groups"])
    assert len(do) > 0
    assert dg == do


def test_c3_shape_four_pools_parity(gpu, oracle):
    """BASELINE config #3 at 1:10 scale: 100k jobs over 4 pools (40/30/20/10 %), 2k nodes,
    host-placement + gpu/ports constraints, groups.  Every pool bit-identical to the oracle;
    plus the size-independent invariants checked on the GPU result itself."""
    from cook_b200.engine import GpuEngine
    shapes = [(40_000, 800), (30_000, 600), (20_000, 400), (10_000, 200)]
    for p, (nj, no) in enumerate(shapes):
        t = traces.gen_c3_pool(300 + p, nj, no, 500, nj // 5)
        ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
        prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"])
        eng = GpuEngine(pool_name=f"pool-{p}")
        try:
            mg = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
        finally:
            eng.close()
        mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
        assert np.array_equal(mg["considerable"], mo["considerable"])
        assert np.array_equal(mg["assign"], mo["assign"]), p
        assert np.array_equal(mg["ports"], mo["ports"])
        # invariants: nothing over-committed, every placed job respects its resources
        placed = mg["assign"] >= 0
        cpus = np.zeros(no); mem = np.zeros(no)
        jc, jm = t["jobs"].col("cpus"), t["jobs"].col("mem")
        np.add.at(cpus, mg["assign"][placed], jc[mg["considerable"][placed]])
        np.add.at(mem, mg["assign"][placed], jm[mg["considerable"][placed]])
        assert (cpus <= t["offers"].col("cpus")).all() and (mem <= t["offers"].col("mem")).all()
        assert mg["stats"]["n_matched"] == int(placed.sum()) > 0


def test_prefix_property_at_c2_size(gpu):
    """Size-independent property of the exact greedy: the placements of the first N jobs do
    not depend on the jobs behind them (checked GPU against GPU on the 100k x 5k trace)."""
    t = traces.gen_c2()
    ranked = gpu.rank(t["running"], t["pending"], t["users"])["ranked"]
    full = gpu.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(100_000))
    for n in (1, 777, 20_000):
        part = gpu.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(n))
        assert np.array_equal(part["assign"], full["assign"][:n])
        assert np.array_equal(part["considerable"], full["considerable"][:n])


@pytest.mark.parametrize("seed,nj,no", [(71, 3000, 7000), (72, 1500, 45000)])
def test_wide_offer_tables(gpu, oracle, seed, nj, no):
    """Offer tables beyond the shared-memory fast paths: 7k offers (static VM table read from
    L2 instead of shared memory), 45k offers (newest-log-entry table in global memory, no
    live-VM mask): same assignments as the oracle."""
    t = traces.gen_pool(seed, nj, no, 50, 500)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    prm = traces.match_params(nj)
    mg = gpu.match(ranked, t["jobs"], t["offers"], t["users"], prm)
    mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm)
    assert np.array_equal(mg["considerable"], mo["considerable"])
    assert np.array_equal(mg["assign"], mo["assign"])


def test_pools_side_by_side_on_one_gpu(oracle):
    """Four pools run their cycles concurrently on one GPU (37 thread blocks each, one host
    thread per pool as in Cook's per-pool match loops): same results as the oracle."""
    import threading
    from cook_b200.engine import GpuEngine
    shapes = [(12_000, 600), (9_000, 500), (6_000, 300), (3_000, 150)]
    pools = []
    for p, (nj, no) in enumerate(shapes):
        t = traces.gen_c3_pool(400 + p, nj, no, 200, nj // 5)
        ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
        prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"], max_ctas=37)
        pools.append((t, ranked, prm, GpuEngine(pool_name=f"pool-{p}")))
    out = [None] * len(pools)

    def cycle(i):
        t, ranked, prm, eng = pools[i]
        out[i] = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)

    for rep in range(2):
        th = [threading.Thread(target=cycle, args=(i,)) for i in range(len(pools))]
        [x.start() for x in th]
        [x.join() for x in th]
        for i, (t, ranked, prm, eng) in enumerate(pools):
            mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
            assert np.array_equal(out[i]["assign"], mo["assign"]), (rep, i)
            assert np.array_equal(out[i]["ports"], mo["ports"])
    for _, _, _, eng in pools:
        eng.close()


@pytest.mark.parametrize("seed,nj,no", [(81, 2500, 7000), (82, 1200, 45000), (83, 3000, 5121)])
def test_wide_offer_tables_with_constraints(gpu, oracle, seed, nj, no):
    """The constraint kernel beyond the shared-memory fast paths: verdict-bit rows of 220 / 1407 /
    161 words, newest-log-entry table in global memory (45k offers), a last tile of one VM (5121)."""
    t = traces.gen_c3_pool(seed, nj, no, 60, nj // 5)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"])
    mo = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    mg = gpu.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    assert np.array_equal(mg["considerable"], mo["considerable"])
    assert np.array_equal(mg["assign"], mo["assign"])
    assert np.array_equal(mg["ports"], mo["ports"])
    assert mg["stats"]["n_matched"] == mo["stats"]["n_matched"] > 0


@pytest.mark.parametrize("frac_pool,frac_group", [(0.5, 0.8), (0.9, 0.3), (0.0, 0.5), (2.0, 2.0)])
def test_rank_queue_filters_parity(gpu, oracle, frac_pool, frac_group):
    """Pool quota, quota group and offensive-job filters binding at different depths of a queue
    that spans several 1024-entry chunks of the sequential filter passes (tools.clj:654-668)."""
    t = traces.gen_pool(91, 5000, 16, 80, 1500)
    tot = float(np.sum(t["pending"].col("cpus"))) + float(np.sum(t["running"].col("cpus")))
    q = lambda f: abi.make_pool_quota(dict(count=1e12, cpus=round(f * tot), mem=1e15, gpus=1e12))
    kw = dict(pool_quota=q(frac_pool), group_quota=q(frac_group), group_usage=np.array([7.0, 33.0, 4096.0, 0.0]),
              params=abi.RankParams(100, 1, 20000.0, 6.0))
    rg = gpu.rank(t["running"], t["pending"], t["users"], **kw)
    ro = oracle.rank(t["running"], t["pending"], t["users"], **kw)
    assert np.array_equal(rg["ranked"], ro["ranked"])
    assert np.array_equal(rg["order"], ro["order"])
    assert _same_dru(rg["dru"], ro["dru"])


def test_rank_and_rebalance_degenerate_inputs(gpu, oracle):
    """No pending jobs / no running tasks: empty queue, rebalancer decisions on spare resources only."""
    t = traces.gen_pool(92, 300, 16, 10, 200)
    from cook_b200.engine import _empty_tasks
    rg = gpu.rank(t["running"], _empty_tasks(), t["users"])
    ro = oracle.rank(t["running"], _empty_tasks(), t["users"])
    assert len(rg["ranked"]) == len(ro["ranked"]) == 0 and np.array_equal(rg["order"], ro["order"])
    rg = gpu.rank(None, t["pending"], t["users"])
    ro = oracle.rank(None, t["pending"], t["users"])
    assert np.array_equal(rg["ranked"], ro["ranked"])
    r = traces.gen_rebalance(93, 0, 12, 20, 8, constraints=False)
    args = (r["running"], r["pending"], r["pending_job_id"], r["pending_priority"], r["hosts"], r["users"], r["params"])
    assert gpu.rebalance(*args) == oracle.rebalance(*args)
    r = traces.gen_rebalance(94, 3000, 0, 50, 20, constraints=False)
    args = (r["running"], r["pending"], r["pending_job_id"], r["pending_priority"], r["hosts"], r["users"], r["params"])
    assert gpu.rebalance(*args) == oracle.rebalance(*args) == []
