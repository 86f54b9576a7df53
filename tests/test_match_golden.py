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
