"""CPU tests: the oracle against the reference's own known-answer tests
(SURVEY §8c K1-K8), transcribed in tests/golden/."""
import json
import os

import numpy as np
import pytest

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
