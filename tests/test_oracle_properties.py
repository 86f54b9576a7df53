"""CPU tests: internal consistency of the oracle (differential and property checks that do not
need the reference): the two restatements of dru/sorted-merge agree under heavy DRU ties, the
threaded matcher used by `bench.py --impl reference` equals the single-threaded one, and the
greedy matcher has the properties the GPU parity tests rely on at full size."""
import numpy as np
import pytest

from cook_b200 import abi, traces


def _rank(oracle, t, naive, **kw):
    oracle.set_naive_merge(naive)
    try:
        return oracle.rank(t["running"], t["pending"], t["users"], **kw)
    finally:
        oracle.set_naive_merge(0)


@pytest.mark.parametrize("seed,nu", [(11, 3), (12, 17), (13, 60)])
def test_sorted_merge_forms_agree_under_ties(oracle, seed, nu):
    """dru.clj:86-103 sorted-merge restated literally (re-sort the heads after every emission) and
    as a heap with the equivalent total order: same queue even when most DRUs tie (equal shares,
    requests on a coarse grid), which is where the 'most recently emitted user first' rule bites."""
    t = traces.gen_pool(seed, 1500, 8, nu, 400, cpus_choices=(1, 2))
    users = abi.make_users(nu, name_rank=t["users"].col("name_rank"), div_mem=np.full(nu, 1e12),
                           div_cpus=np.full(nu, 4.0))
    t = dict(t, users=users)
    a, b = _rank(oracle, t, 1), _rank(oracle, t, 0)
    assert np.array_equal(a["order"], b["order"])
    assert np.array_equal(a["ranked"], b["ranked"])
    dru = a["dru"][a["order"]]
    assert np.all(np.diff(dru) >= 0)                       # ascending DRU
    assert len(np.unique(dru)) < 0.7 * len(dru)            # the case really has many ties


@pytest.mark.parametrize("threads", [2, 5])
def test_threaded_restatement_equals_single_thread(oracle, threads):
    t = traces.gen_c3_pool(21, 1500, 120, 30, 300)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    prm = traces.match_params(1500, host_lifetime_mins=t["host_lifetime_mins"])
    one = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    many = oracle.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2,
                        threads=threads)
    for k in ("considerable", "assign", "ports", "fail"):
        assert np.array_equal(one[k], many[k]), k
    assert one["stats"]["n_matched"] == many["stats"]["n_matched"] > 0


def test_greedy_matcher_properties(oracle):
    """Prefix property (the first n jobs are placed as in the full run), capacity is never
    exceeded, a job that stays unplaced fits on no VM at the end of the cycle either (cpu+mem
    pools: resources only shrink), and the run is deterministic."""
    t = traces.gen_pool(31, 4000, 150, 40, 800)
    ranked = oracle.rank(t["running"], t["pending"], t["users"])["ranked"]
    full = oracle.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(4000))
    again = oracle.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(4000))
    assert np.array_equal(full["assign"], again["assign"])
    for n in (1, 37, 1000, 2500):
        part = oracle.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(n))
        assert np.array_equal(part["assign"], full["assign"][:n])
    jc, jm = t["jobs"].col("cpus"), t["jobs"].col("mem")
    oc, om = t["offers"].col("cpus"), t["offers"].col("mem")
    used_c, used_m = np.zeros(len(oc)), np.zeros(len(om))
    placed = full["assign"] >= 0
    np.add.at(used_c, full["assign"][placed], jc[full["considerable"][placed]])
    np.add.at(used_m, full["assign"][placed], jm[full["considerable"][placed]])
    assert np.all(used_c <= oc) and np.all(used_m <= om)
    left_c, left_m = oc - used_c, om - used_m
    for k in np.flatnonzero(~placed)[:200]:
        j = full["considerable"][k]
        assert not np.any((jc[j] <= left_c) & (jm[j] <= left_m))
    assert 0 < placed.sum() < len(placed)
