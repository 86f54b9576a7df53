"""Considerable-job known answers transcribed BY HAND from the reference
(test/cook/test/scheduler/scheduler.clj:1566-1705,
test-pending-jobs->considerable-jobs).  One user; usage {count 1, cpus 2, mem 1024, gpus 0}."""
import numpy as np

from cook_b200 import abi, traces

NON_GPU = [(3, 2048, 0), (13, 1024, 0), (7, 4096, 0), (11, 1024, 0)]   # job-1..job-4 (cpus, mem, gpus)
GPU = [(5, 2048, 2), (19, 1024, 4)]                                   # job-5, job-6
USAGE = {"count": 1, "cpus": 2, "mem": 1024, "gpus": 0}


def _considerable(eng, jobs, quota, num_considerable, plugin_accept=1, tokens=None):
    J = len(jobs)
    jb = abi.JobsSoA(n=J, user=np.zeros(J, np.int32), cpus=np.array([j[0] for j in jobs], float),
                     mem=np.array([j[1] for j in jobs], float), gpus=np.array([j[2] for j in jobs], float),
                     allowed=np.ones(J, np.uint8), plugin_accept=np.full(J, plugin_accept, np.uint8))
    of = abi.OffersSoA(n=1, hostname_id=np.zeros(1, np.int32), name_rank=np.zeros(1, np.int32),
                       cpus=np.array([1.0]), mem=np.array([1.0]), run_cpus=np.zeros(1), run_mem=np.zeros(1),
                       run_count=np.zeros(1, np.int32), n_attr_cols=0)
    users = abi.make_users(1, quota={k: np.array([float(v)]) for k, v in quota.items()},
                           usage={k: np.array([float(v)]) for k, v in USAGE.items()},
                           tokens=None if tokens is None else np.array([tokens], np.int32))
    prm = traces.match_params(num_considerable, enforce_rate_limit=0 if tokens is None else 1)
    m = eng.match(np.arange(J, dtype=np.int32), jb, of, users, prm)
    return [int(x) for x in m["considerable"]]


def _k9(eng, jobs, users_of, quota, usage, pool_quota, nc=10):
    """jobs: [(cpus, mem)], users_of: user index per job; quota/usage: per-user dicts of lists."""
    J = len(jobs)
    nu = max(users_of) + 1
    jb = abi.JobsSoA(n=J, user=np.array(users_of, np.int32), cpus=np.array([j[0] for j in jobs], float),
                     mem=np.array([j[1] for j in jobs], float), gpus=np.zeros(J),
                     allowed=np.ones(J, np.uint8), plugin_accept=np.ones(J, np.uint8))
    of = abi.OffersSoA(n=1, hostname_id=np.zeros(1, np.int32), name_rank=np.zeros(1, np.int32),
                       cpus=np.array([1.0]), mem=np.array([1.0]), run_cpus=np.zeros(1), run_mem=np.zeros(1),
                       run_count=np.zeros(1, np.int32), n_attr_cols=0)
    users = abi.make_users(nu, quota={k: np.array(v, float) for k, v in quota.items()} if quota else None,
                           usage={k: np.array(v, float) for k, v in usage.items()})
    pq = abi.make_pool_quota(pool_quota) if pool_quota else None
    m = eng.match(np.arange(J, dtype=np.int32), jb, of, users, traces.match_params(nc), pool_quota=pq)
    return [int(x) for x in m["considerable"]]


def check_k9(eng):
    """K9: test/cook/test/tools.clj:763-816 (pool quota, user quota, user quota before pool quota)."""
    q4 = [(2, 2048), (1, 1024), (3, 4096), (1, 1024)]
    use = {"count": [1], "cpus": [2], "mem": [1024]}
    big = 1e18
    # filter-based-on-pool-quota (:763-778); the running usage counts against the pool quota
    for pq, expect in [({"count": 1, "cpus": 2, "mem": 1024, "gpus": big}, []),
                       ({"count": 10, "cpus": 20, "mem": 32768, "gpus": big}, [0, 1, 2, 3]),
                       ({"count": 4, "cpus": 20, "mem": 6144, "gpus": big}, [0, 1])]:
        assert _k9(eng, q4, [0, 0, 0, 0], None, use, pq) == expect, ("K9 pool", pq)
    # filter-based-on-user-quota (:780-797)
    for uq, expect in [({"count": [1], "cpus": [2], "mem": [1024]}, []),
                       ({"count": [10], "cpus": [20], "mem": [32768]}, [0, 1, 2, 3]),
                       ({"count": [4], "cpus": [20], "mem": [6144]}, [0, 1])]:
        assert _k9(eng, q4, [0, 0, 0, 0], uq, use, None) == expect, ("K9 user", uq)
    # filter-pending-jobs-for-quota (:799-816): user quota filters first => [job-1 job-4]
    got = _k9(eng, [(1, 1)] * 4, [0, 0, 0, 1], {"count": [2, 2], "cpus": [100, 100], "mem": [100, 100]},
              {"count": [1, 1], "cpus": [1, 1], "mem": [1, 1]}, {"count": 4, "cpus": 100, "mem": 100, "gpus": big})
    assert got == [0, 3], ("K9 order", got)


def check_all(eng):
    check_k9(eng)
    big = {"count": 10, "cpus": 50, "mem": 32768, "gpus": 10}
    # every job deferred by the launch plugin => nothing considerable
    assert _considerable(eng, NON_GPU, big, 5, plugin_accept=0) == []
    # jobs inside usage quota
    assert _considerable(eng, NON_GPU, big, 5) == [0, 1, 2, 3]
    assert _considerable(eng, GPU, big, 5) == [0, 1]
    # inside quota but beyond the launch-rate limit (one token saved): only the first job
    assert _considerable(eng, NON_GPU, big, 5, tokens=1) == [0]
    assert _considerable(eng, GPU, big, 5, tokens=1) == [0]
    # limited by num-considerable
    assert _considerable(eng, NON_GPU, big, 3) == [0, 1, 2]
    assert _considerable(eng, GPU, big, 3) == [0, 1]
    assert _considerable(eng, NON_GPU, big, 2) == [0, 1]
    assert _considerable(eng, GPU, big, 2) == [0, 1]
    assert _considerable(eng, NON_GPU, big, 1) == [0]
    assert _considerable(eng, GPU, big, 1) == [0]
    # some jobs inside usage quota
    some = {"count": 5, "cpus": 10, "mem": 4096, "gpus": 10}
    assert _considerable(eng, NON_GPU, some, 5) == [0]
    assert _considerable(eng, GPU, some, 5) == [0]
    # quota gpus not ignored
    nogpu = {"count": 5, "cpus": 10, "mem": 4096, "gpus": 0}
    assert _considerable(eng, NON_GPU, nogpu, 5) == [0]
    assert _considerable(eng, GPU, nogpu, 5) == []
    # all jobs exceed quota
    tiny = {"count": 5, "cpus": 3, "mem": 4096, "gpus": 10}
    assert _considerable(eng, NON_GPU, tiny, 5) == []
    assert _considerable(eng, GPU, tiny, 5) == []
