"""Fenzo restatement pinned at set level by the reference's own matcher tests
(K10, K12, K13, K15).  CPU: oracle.  GPU (-m gpu): the CUDA path on the same cases."""
import pytest

from match_golden_cases import check_all


def test_match_golden_oracle(oracle):
    check_all(oracle)


@pytest.mark.gpu
def test_match_golden_gpu(gpu):
    check_all(gpu)


def test_constraint_truth_tables_oracle(oracle):
    from constraint_golden_cases import check_all as check_constraints
    check_constraints(oracle)


@pytest.mark.gpu
def test_constraint_truth_tables_gpu(gpu):
    from constraint_golden_cases import check_all as check_constraints
    check_constraints(gpu)


def test_considerable_golden_oracle(oracle):
    from considerable_golden_cases import check_all as check_considerable
    check_considerable(oracle)


@pytest.mark.gpu
def test_considerable_golden_gpu(gpu):
    from considerable_golden_cases import check_all as check_considerable
    check_considerable(gpu)


def test_rebalance_constraint_golden_oracle(oracle):
    from rebalance_constraint_golden import check_all as check_reb
    check_reb(oracle)


@pytest.mark.gpu
def test_rebalance_constraint_golden_gpu(gpu):
    from rebalance_constraint_golden import check_all as check_reb
    check_reb(gpu)


def _k17(eng):
    """K17: test/cook/test/rebalancer.clj:55-113 (init-state): DRU of 8 running tasks, share
    mem 25 / cpus 25; the rebalancer's priority map lists them by descending DRU."""
    import numpy as np
    from cook_b200 import abi
    from cook_b200.engine import _empty_tasks
    run = [("ljin", 10, 10), ("ljin", 5, 5), ("ljin", 15, 25), ("ljin", 25, 15),
           ("wzhao", 8, 8), ("wzhao", 10, 10), ("wzhao", 10, 10), ("wzhao", 10, 10)]   # (user, mem, cpus)
    uid = {"ljin": 0, "wzhao": 1}
    R = len(run)
    t = abi.make_tasks(user=np.array([uid[r[0]] for r in run], np.int32), priority=np.full(R, 50, np.int32),
                       start_time=np.full(R, 1_600_000_000_000, np.int64),
                       task_id=np.arange(1000, 1000 + R, dtype=np.int64), job_id=np.arange(1, R + 1, dtype=np.int64),
                       cpus=np.array([float(r[2]) for r in run]), mem=np.array([float(r[1]) for r in run]))
    users = abi.make_users(2, div_mem=np.full(2, 25.0), div_cpus=np.full(2, 25.0), div_gpus=np.full(2, 1.0))
    out = eng.rank(t, _empty_tasks(), users)
    assert np.allclose(out["dru"][:R], [0.4, 0.6, 1.6, 2.2, 0.32, 0.72, 1.12, 1.52], rtol=1e-12, atol=0)
    # ascending merge order reversed = [t4 t3 t8 t7 t6 t2 t1 t5]
    assert [int(x) for x in out["order"][::-1]] == [3, 2, 7, 6, 5, 1, 0, 4]


def test_init_state_dru_golden_oracle(oracle):
    _k17(oracle)


@pytest.mark.gpu
def test_init_state_dru_golden_gpu(gpu):
    _k17(gpu)


def test_handle_offers_golden_oracle(oracle):
    """K15 (test/cook/test/scheduler/scheduler.clj:1947-2239): launched job sets and offer counts of
    handle-resource-offers! against the real Fenzo, for the Mesos and the Kubernetes offer tables."""
    import handle_offers_golden_cases
    assert handle_offers_golden_cases.check_all(oracle) == 2 * (13 + 1) + 2 * (13 + 6)


@pytest.mark.gpu
def test_handle_offers_golden_gpu(gpu):
    """K15 in full through the CUDA path (the same 66 assertions)."""
    import handle_offers_golden_cases
    assert handle_offers_golden_cases.check_all(gpu) == 2 * (13 + 1) + 2 * (13 + 6)


def test_rebalance_balanced_and_quota_golden_oracle(oracle):
    """K20, second half (test/cook/test/rebalancer.clj:673-811): balanced host-placement groups (one
    with a host already preempted in the cycle) and the over-quota rule (dru 100.0, own task)."""
    import rebalance_constraint_golden
    assert rebalance_constraint_golden.check_balanced_and_quota(oracle) == 3


@pytest.mark.gpu
def test_rebalance_balanced_and_quota_golden_gpu(gpu):
    """K20, second half, through the CUDA path (cook_rebalance_trace applies the given first decision)."""
    import rebalance_constraint_golden
    assert rebalance_constraint_golden.check_balanced_and_quota(gpu) == 3


def test_cycle_driver_golden_oracle(oracle):
    """M5/M6: handle-resource-offers! as a whole (return value, launch-rate filter, reservations,
    queue removal, scale-back) around the oracle."""
    import cycle_golden_cases
    assert cycle_golden_cases.check_state_machine()
    assert cycle_golden_cases.check_all(oracle) == 9
    assert cycle_golden_cases.check_autoscaling(oracle)


@pytest.mark.gpu
def test_cycle_driver_golden_gpu(gpu):
    import cycle_golden_cases
    assert cycle_golden_cases.check_all(gpu) == 9
    assert cycle_golden_cases.check_autoscaling(gpu)
