"""M5/M6 (handle-resource-offers! as a whole) known answers: the K15 cases whose subject is the cycle
driver rather than the matcher - the function's RETURN VALUE ("matched the head or matched nothing"),
the compute-cluster launch-rate filter, the rebalancer reservations before / after, the queue
without the matched jobs - through cook_b200.cycle.PoolCycle around either engine.
Transcribed from test/cook/test/scheduler/scheduler.clj:1947-2133 (test-handle-resource-helpers)."""
import numpy as np

from cook_b200 import abi
from cook_b200.cycle import OfferCache, PoolCycle, next_considerable, offers_from_cache
from handle_offers_golden_cases import JOBS, MESOS, QUOTA, USAGE, P100


def _jobs():
    J = len(JOBS)
    return abi.JobsSoA(n=J, user=np.zeros(J, np.int32), cpus=np.array([j[0] for j in JOBS], float),
                       mem=np.array([j[1] for j in JOBS], float), gpus=np.array([j[2] for j in JOBS], float),
                       allowed=np.ones(J, np.uint8), plugin_accept=np.ones(J, np.uint8),
                       gpu_model=np.full(J, P100, np.int32))


def _users(quota=None, usage=None):
    q, u = quota or QUOTA, usage or USAGE
    return abi.make_users(1, quota={k: np.array([float(v)]) for k, v in q.items()},
                          usage={k: np.array([float(v)]) for k, v in u.items()})


def _offers(ids):
    cache = OfferCache()
    cache.add([{"hostname": f"host-{i}", "cpus": MESOS[i][0], "mem": MESOS[i][1], "id": i} for i in ids], now_ms=0)
    names, of = offers_from_cache(cache, {f"host-{i}": i for i in MESOS})
    return cache, names, of


def check_all(eng):
    n = 0
    queue = np.arange(len(JOBS), dtype=np.int32)
    # :1957-1964 enough offers, nc 6: returns true, 4 jobs on 3 offers, the matched jobs leave the queue
    pc = PoolCycle(eng, max_considerable=6)
    _, _, of = _offers([1, 2, 3])
    r = pc.handle_resource_offers(queue, _jobs(), of, _users())
    assert r["matched_head_or_no_matches"] is True and r["launched_jobs"] == {0, 1, 2, 3} and len(r["launched_offers"]) == 3
    assert list(r["queue"]) == [4, 5, 6, 7] and r["next_considerable"] == 6
    n += 1
    # :1966-1982 limited by num-considerable 1 / 2
    for nc, want in ((1, {0}), (2, {0, 1})):
        r = PoolCycle(eng, max_considerable=nc).handle_resource_offers(queue, _jobs(), _offers([1, 2, 3])[2], _users())
        assert r["matched_head_or_no_matches"] and r["launched_jobs"] == want and len(r["launched_offers"]) == len(want)
        n += 1
    # :2015-2028 the compute cluster's launch-rate limiter is enforcing and in debt: every match of
    # the cycle is dropped, nothing launches - and the function still returns true (no matches)
    pc = PoolCycle(eng, max_considerable=10)
    r = pc.handle_resource_offers(queue, _jobs(), _offers([1, 2, 3])[2], _users(),
                                  cluster_of_offer=["cc", "cc", "cc"], cluster_tokens={"cc": -1}, cluster_enforce={"cc": True})
    assert r["matched_head_or_no_matches"] is True and not r["launched_jobs"] and not r["launched_offers"]
    assert list(r["queue"]) == list(queue) and pc.num_considerable == 10
    n += 1
    # :2002-2013 limiter with tokens left: nothing is dropped
    r = PoolCycle(eng, max_considerable=6).handle_resource_offers(
        queue, _jobs(), _offers([1, 2, 3])[2], _users(), cluster_of_offer=["cc"] * 3, cluster_tokens={"cc": 1},
        cluster_enforce={"cc": True})
    assert r["launched_jobs"] == {0, 1, 2, 3}
    n += 1
    # :2086-2092 offer fits no job: true (nothing matched), queue untouched
    r = PoolCycle(eng, max_considerable=10).handle_resource_offers(queue, _jobs(), _offers([5])[2], _users())
    assert r["matched_head_or_no_matches"] is True and not r["launched_jobs"] and len(r["queue"]) == len(JOBS)
    n += 1
    # :2112-2119 a host reserved for somebody else's job: nothing launches, reservations unchanged
    pc = PoolCycle(eng, max_considerable=10)
    pc.job_reserved_host = {99: 1}           # some other job holds host-1
    r = pc.handle_resource_offers(queue, _jobs(), _offers([1])[2], _users())
    assert not r["launched_jobs"] and pc.job_reserved_host == {99: 1}
    n += 1
    # :2121-2133 only the jobs the host is reserved for launch there; their reservations are released
    # and they are recorded as launched
    pc = PoolCycle(eng, max_considerable=10)
    pc.job_reserved_host = {0: 9, 1: 9}
    r = pc.handle_resource_offers(queue, _jobs(), _offers([9])[2], _users())
    assert r["launched_jobs"] == {0, 1} and pc.job_reserved_host == {} and pc.launched_jobs == {0, 1}
    n += 1
    # the head of the queue cannot be placed but a later job can: the function returns FALSE and the
    # pool shows Fenzo fewer jobs next cycle (handle-fenzo-pool :1613-1651)
    pc = PoolCycle(eng, max_considerable=1000)
    r = pc.handle_resource_offers(np.array([1, 0, 2, 3], np.int32), _jobs(), _offers([1])[2], _users())   # job-2 (13 cpus) leads; offer-1 has 10
    assert r["matched_head_or_no_matches"] is False and r["launched_jobs"] == {0} and r["next_considerable"] == 950
    n += 1
    return n


def check_state_machine():
    """handle-fenzo-pool's num-considerable arithmetic and the lease cache, no engine needed."""
    # max 1000, scaleback 0.95: 1000 -> 950 -> 902 -> ... reaches 1 after 88 failed cycles ("this will take 88 seconds")
    nc, at_floor, steps = 1000, 0, 0
    while nc > 1:
        nc, at_floor = next_considerable(nc, False, 1000, 0.95, at_floor, 10 ** 9)
        steps += 1
    assert 86 <= steps <= 89, steps      # the reference comment (:1621) says "88 seconds" at 1 iteration / s
    # at the floor: counts iterations, resets to max after floor_iterations_before_reset
    nc, at_floor = 1, 0
    for i in range(4):
        nc, at_floor = next_considerable(nc, False, 1000, 0.95, at_floor, 5)
        assert (nc, at_floor) == (1, i + 1)
    nc, at_floor = next_considerable(nc, False, 1000, 0.95, at_floor, 5)
    assert nc == 1000
    # a matched head restores max-considerable at once
    assert next_considerable(37, True, 1000, 0.95, 0, 5) == (1000, 0)
    # lease cache: two offers of one host merge, unused Mesos leases expire after the incubation time,
    # single-shot (Kubernetes) leases never survive a match attempt
    c = OfferCache(incubate_ms=15_000)
    c.add([{"hostname": "a", "cpus": 4.0, "mem": 100.0, "ports": [(31000, 31009)]},
           {"hostname": "b", "cpus": 1.0, "mem": 10.0, "single_shot": True}], now_ms=0)
    c.add([{"hostname": "a", "cpus": 2.0, "mem": 50.0, "ports": [(31500, 31504)]}], now_ms=10_000)
    names, of = offers_from_cache(c, {"a": 7, "b": 8})
    assert names == ["a", "b"] and list(of.col("cpus")) == [6.0, 1.0] and list(of.col("mem")) == [150.0, 10.0]
    assert list(of.col("port_off")) == [0, 2, 2] and list(of.col("port_begin")) == [31000, 31500]
    consumed, dropped = c.after_match([])
    assert not consumed and [l["hostname"] for l in dropped] == ["b"]
    assert [l["hostname"] for l in c.expire(16_000)] == ["a"] and len(c.leases) == 1     # the first lease of a
    assert [l["hostname"] for l in c.expire(26_000)] == ["a"] and not c.leases
    return True


def check_autoscaling(eng):
    """SURVEY §8f-4 (scheduler.clj:1283-1335): the scaled maximum and the quota-filtered prefix."""
    from cook_b200.cycle import autoscalable_jobs, max_jobs_for_autoscaling_scaled
    # 20 % unmatched, scale factor 2.5, max 1000 -> 500; never fewer than the unmatched jobs themselves
    assert max_jobs_for_autoscaling_scaled(100, 20, 1000, 2.5) == 500
    assert max_jobs_for_autoscaling_scaled(100, 90, 1000, 2.5) == 1000
    assert max_jobs_for_autoscaling_scaled(100, 90, 50, 1.0) == 90
    assert max_jobs_for_autoscaling_scaled(0, 0, 1000, 2.5) == 0
    # the eight K15 jobs of one user, quota {count 10 cpus 70 mem 32768}, usage {1, 2, 1024}: cpus run out after
    # 3 + 13 + 7 + 11 + 5 + 19 = 58 (+2 used) <= 70, job-7 (1 cpu) and job-8 (2 cpus) still fit -> all eight;
    # a quota of 40 cpus: job-5 would reach 41 > 40 - and filter-sequential advances the usage for rejected jobs too
    # (tools.clj:654-668), so nothing after it passes either
    q = np.arange(len(JOBS), dtype=np.int32)
    got = autoscalable_jobs(eng, q, _jobs(), _users(), number_considerable=8, number_unmatched=8, max_jobs_for_autoscaling=100)
    assert got == [0, 1, 2, 3, 4, 5, 6, 7]
    got = autoscalable_jobs(eng, q, _jobs(), _users(quota=dict(count=10, cpus=40, mem=32768, gpus=10)),
                            number_considerable=8, number_unmatched=8, max_jobs_for_autoscaling=100)
    assert got == [0, 1, 2, 3]
    # the queue after a match (jobs 0-3 launched), half unmatched, max 4 -> take 4 of the rest; recent synthetic pods dropped
    got = autoscalable_jobs(eng, np.array([4, 5, 6, 7], np.int32), _jobs(), _users(), number_considerable=8, number_unmatched=4,
                            max_jobs_for_autoscaling=4, scale_factor=1.0, recent_synthetic_pod_jobs=(5,))
    assert got == [4, 6, 7]
    return True
