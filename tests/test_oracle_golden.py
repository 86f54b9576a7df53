"""CPU tests: the oracle against the reference's own known-answer tests
(SURVEY §8c K1-K8), transcribed in tests/golden/."""
import json
import os

import numpy as np
import pytest

from cook_b200 import abi
from golden_util import check_rank_case

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "rank_golden.json")) as f:
    RANK_CASES = json.load(f)


@pytest.mark.parametrize("case", RANK_CASES, ids=[c["name"] for c in RANK_CASES])
@pytest.mark.parametrize("naive", [0, 1])
def test_rank_golden(case, naive):
    from oracle.pyoracle import OracleEngine

    def factory(mode):
        e = OracleEngine(dru_mode=mode)
        e.set_naive_merge(naive)
        return e
    try:
        check_rank_case(case, factory)
    finally:
        OracleEngine().set_naive_merge(0)


with open(os.path.join(HERE, "golden", "rebalance_golden.json")) as f:
    REB_CASES = json.load(f)


@pytest.mark.parametrize("case", REB_CASES, ids=[c["name"] for c in REB_CASES])
def test_rebalance_golden(case):
    from golden_util import check_rebalance_case
    from oracle.pyoracle import OracleEngine
    check_rebalance_case(case, OracleEngine())


def _reb_case(running, pending, spare, max_preemption=8, min_dru_diff=0.0):
    return dict(name="state", running=running, pending=pending, spare=spare, share=dict(mem=25.0, cpus=25.0),
                shares={}, params=dict(max_preemption=max_preemption, min_dru_diff=min_dru_diff,
                                       safe_dru_threshold=1.0))


def _rj(user, mem, cpus, host):
    return dict(user=user, mem=float(mem), cpus=float(cpus), host=host)


def _pj(user, mem, cpus):
    return dict(user=user, mem=float(mem), cpus=float(cpus))


def check_pending_job_dru(eng):
    """K18: test/cook/test/rebalancer.clj:115-157 compute-pending-default-job-dru = 1.92 / 0.8 / 2.6.
    (`:ucpus` in job4 and job11 is not a create-dummy-job key, so those jobs get the default 1.0
    cpus, testutil.clj:234-253.)  The GPU-mode half (:159-203) has no rebalancer behaviour to
    restate: compute-preemption-decision throws in that mode (oracle header)."""
    from golden_util import rebalance_inputs
    from oracle.pyoracle import OracleEngine
    run = [_rj("ljin", 10, 10, "h"), _rj("ljin", 5, 5, "h"), _rj("ljin", 15, 25, "h"), _rj("ljin", 25, 1, "h"),
           _rj("wzhao", 8, 8, "h"), _rj("wzhao", 10, 10, "h"), _rj("wzhao", 10, 10, "h"), _rj("wzhao", 10, 10, "h")]
    pend = [_pj("wzhao", 10, 10), _pj("sunil", 20, 20), _pj("ljin", 10, 1)]
    inp = rebalance_inputs(_reb_case(run, pend, {}, min_dru_diff=1e9))   # nothing is preemptable: walk all jobs
    out = eng.rebalance_trace(inp["running"], inp["pending"], inp["pending_job_id"],
                              inp["pending_priority"], inp["hosts"], inp["users"], inp["params"])
    assert out["decisions"] == []
    assert list(out["pending_dru"]) == [1.92, 0.8, 2.6]      # `(is (= 1.92 ...))`: exact


def test_pending_job_dru_golden(oracle):
    check_pending_job_dru(oracle)


@pytest.mark.gpu
def test_pending_job_dru_golden_gpu(gpu):
    check_pending_job_dru(gpu)


K21_RUN = [_rj("ljin", 10, 10, "hostA"), _rj("ljin", 5, 5, "hostA"), _rj("ljin", 15, 25, "hostB"),
           _rj("ljin", 25, 15, "hostB"), _rj("wzhao", 8, 8, "hostA"), _rj("wzhao", 10, 10, "hostB"),
           _rj("wzhao", 10, 10, "hostA"), _rj("wzhao", 10, 10, "hostB")]
K21_PEND = [_pj("wzhao", 15, 15), _pj("sunil", 15, 15), _pj("ljin", 15, 15), _pj("sunil", 40, 40),
            _pj("sunil", 45, 45), _pj("sunil", 80, 80)]
NEW = 8   # index of the task next-state creates for the pending job
K21 = [   # (cite, pending idx, host, victims, mem, cpus, expected key order, expected drus, expected spare)
    ("813-912", 0, "hostB", [5, 7], 20.0, 20.0, [3, 2, NEW, 6, 1, 0, 4], [2.2, 1.6, 1.32, 0.72, 0.6, 0.4, 0.32],
     {"hostA": (50.0, 50.0, 0.0), "hostB": (5.0, 5.0, 0.0)}),
    ("914-941", 1, "hostA", [1, 6], 65.0, 65.0, [3, 2, 7, 5, NEW, 0, 4], [2.0, 1.4, 1.12, 0.72, 0.6, 0.4, 0.32],
     {"hostA": (50.0, 50.0, 0.0)}),
    ("943-988", 3, "hostA", [], 50.0, 50.0, [3, 2, NEW, 7, 6, 5, 1, 0, 4],
     [2.2, 1.6, 1.6, 1.52, 1.12, 0.72, 0.6, 0.4, 0.32], {"hostA": (10.0, 10.0, 0.0)}),
]


def check_next_state(eng, k):
    """K21: test/cook/test/rebalancer.clj:813-988 next-state: task->scored-task keys and scores and
    host->spare-resources after applying a given decision."""
    from golden_util import rebalance_inputs
    from oracle.pyoracle import OracleEngine
    cite, pidx, host, victims, mem, cpus, order, drus, spare = k
    inp = rebalance_inputs(_reb_case(K21_RUN, K21_PEND, {"hostA": (50.0, 50.0)}))
    hid = {h: i for i, h in enumerate(inp["hostnames"])}
    out = eng.rebalance_trace(inp["running"], inp["pending"], inp["pending_job_id"],
                              inp["pending_priority"], inp["hosts"], inp["users"], inp["params"],
                              forced=[(pidx, hid[host], victims, mem, cpus, 0.0)])
    assert out["order"] == order, (cite, out["order"])
    assert np.allclose(out["order_dru"], drus, rtol=1e-12, atol=0), (cite, out["order_dru"])
    assert out["spare"] == {hid[h]: v for h, v in spare.items()}, (cite, out["spare"])


@pytest.mark.parametrize("k", K21, ids=[k[0] for k in K21])
def test_next_state_golden(oracle, k):
    check_next_state(oracle, k)


@pytest.mark.gpu
@pytest.mark.parametrize("k", K21, ids=[k[0] for k in K21])
def test_next_state_golden_gpu(gpu, k):
    check_next_state(gpu, k)


def check_job_below_quota(eng):
    """test/cook/test/rebalancer.clj:1368-1397: testA has a count quota of 1 and one running task, so
    its waiting job is not below quota; testB (no quota) is."""
    from golden_util import rebalance_inputs
    from oracle.pyoracle import OracleEngine
    run = [_rj("testA", 10, 1, "hostA"), _rj("testB", 10, 1, "hostA")]
    pend = [_pj("testA", 10, 1), _pj("testB", 10, 1)]
    inp = rebalance_inputs(_reb_case(run, pend, {}, min_dru_diff=1e9))
    big = np.finfo(np.float64).max
    users = abi.make_users(2, name_rank=np.arange(2, dtype=np.int32), div_mem=np.full(2, 25.0), div_cpus=np.full(2, 25.0),
                           div_gpus=np.ones(2), quota=dict(count=np.array([1.0, big]), cpus=np.full(2, big),
                                                           mem=np.full(2, big), gpus=np.full(2, big)))
    out = eng.rebalance_trace(inp["running"], inp["pending"], inp["pending_job_id"],
                              inp["pending_priority"], inp["hosts"], users, inp["params"])
    assert out["below_quota"] == [False, True]


def test_job_below_quota_golden(oracle):
    check_job_below_quota(oracle)


@pytest.mark.gpu
def test_job_below_quota_golden_gpu(gpu):
    check_job_below_quota(gpu)


def check_filter_offensive_jobs(eng):
    """R7, test/cook/test/scheduler/scheduler.clj:857-888: constraints {memory-gb 10, cpus 5}; a 12 GB
    job and a 6-cpu job are offensive, the 8 GB / 4 cpu job stays in the queue."""
    from cook_b200.engine import _empty_tasks
    from oracle.pyoracle import OracleEngine
    mem = np.array([1024.0 * 12.0, 1024.0 * 8.0, 1024.0 * 8.0])
    cpus = np.array([4.0, 6.0, 4.0])
    pending = abi.make_tasks(user=np.zeros(3, np.int32), priority=np.full(3, 50, np.int32),
                             start_time=np.full(3, abi.INT64_MAX, np.int64), task_id=np.full(3, -1, np.int64),
                             job_id=np.arange(1, 4, dtype=np.int64), cpus=cpus, mem=mem)
    users = abi.make_users(1)
    out = eng.rank(_empty_tasks(), pending, users, params=abi.RankParams(100, 1, 1024.0 * 10.0, 5.0))
    assert list(out["ranked"]) == [2]
    out = eng.rank(_empty_tasks(), pending, users, params=abi.RankParams(100, 0, 0.0, 0.0))
    assert sorted(out["ranked"]) == [0, 1, 2]


def test_filter_offensive_jobs_golden(oracle):
    check_filter_offensive_jobs(oracle)


@pytest.mark.gpu
def test_filter_offensive_jobs_golden_gpu(gpu):
    check_filter_offensive_jobs(gpu)
