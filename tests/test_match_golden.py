"""Fenzo restatement pinned at set level by the reference's own matcher tests
(K10, K12, K13, K15).  CPU: oracle.  GPU (-m gpu): the CUDA path on the same cases."""
import pytest

from match_golden_cases import check_all


def test_match_golden_oracle(oracle):
    check_all(oracle)


@pytest.mark.gpu
def test_match_golden_gpu(gpu):
    check_all(gpu)
