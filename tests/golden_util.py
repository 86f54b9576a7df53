"""Turns the hand-transcribed golden cases (tests/golden/*.json) into ABI structs."""
import numpy as np

from cook_b200 import abi

BIG = np.finfo(np.float64).max
T0 = 1_600_000_000_000


def rank_inputs(case):
    jobs = case["jobs"]
    names = sorted({j["user"] for j in jobs})
    uid = {n: i for i, n in enumerate(names)}
    nu = len(names)
    ds = case.get("default_share") or {}
    div = {k: np.full(nu, ds.get(k, BIG)) for k in ("mem", "cpus", "gpus")}
    for u, sh in case["shares"].items():
        for k, v in sh.items():
            div[k][uid[u]] = v
    quota = {k: np.full(nu, BIG) for k in ("count", "cpus", "mem", "gpus")}
    for u, q in case["quotas"].items():
        for k, v in q.items():
            quota[k][uid[u]] = v
    users = abi.make_users(nu, name_rank=np.arange(nu, dtype=np.int32), div_mem=div["mem"],
                           div_cpus=div["cpus"], div_gpus=div["gpus"], quota=quota)
    run_idx = [i for i, j in enumerate(jobs) if j["running"]]
    order = case.get("instance_order") or run_idx
    task_id = {ji: 1000 + k for k, ji in enumerate(order)}
    pend_idx = [i for i, j in enumerate(jobs) if not j["running"]]

    def soa(idx, running):
        return abi.make_tasks(
            user=np.array([uid[jobs[i]["user"]] for i in idx], np.int32),
            priority=np.array([jobs[i]["prio"] for i in idx], np.int32),
            start_time=np.array([T0 if running else abi.INT64_MAX for _ in idx], np.int64),
            task_id=np.array([task_id[i] if running else -1 for i in idx], np.int64),
            job_id=np.array([i + 1 for i in idx], np.int64),
            cpus=np.array([jobs[i]["cpus"] for i in idx], np.float64),
            mem=np.array([jobs[i]["mem"] for i in idx], np.float64),
            gpus=np.array([jobs[i]["gpus"] for i in idx], np.float64))

    running = soa(run_idx, True)
    pending = soa(pend_idx, False)
    pq = abi.make_pool_quota(case.get("pool_quota"))
    gq = abi.make_pool_quota(case.get("group_quota"))
    gu = case.get("group_usage")
    gu = np.array([gu["count"], gu["cpus"], gu["mem"], gu["gpus"]], np.float64) if gu else None
    params = abi.RankParams(case["max_over_quota"], 0, 0.0, 0.0)
    return dict(running=running, pending=pending, users=users, pool_quota=pq, group_quota=gq,
                group_usage=gu, params=params, run_idx=run_idx, pend_idx=pend_idx)


def check_rank_case(case, engine_factory):
    inp = rank_inputs(case)
    eng = engine_factory(case["dru_mode"])
    out = eng.rank(inp["running"], inp["pending"], inp["users"], pool_quota=inp["pool_quota"],
                   group_quota=inp["group_quota"], group_usage=inp["group_usage"],
                   params=inp["params"])
    R = len(inp["run_idx"])
    comb = inp["run_idx"] + inp["pend_idx"]
    if case["expect_kind"] == "pending_jobs":
        got = [inp["pend_idx"][j] for j in out["ranked"]]
        assert got == case["expect"], (case["name"], got, case["expect"])
    else:
        got_order = [comb[t] for t in out["order"]]
        if case["expect"] is not None:
            assert got_order == case["expect"], (case["name"], got_order)
        drus = [out["dru"][t] for t in out["order"]]
        assert np.allclose(drus, case["expect_drus"], rtol=1e-12, atol=0), (case["name"], drus)
    return out


def rebalance_inputs(case):
    """spare: host -> (mem, cpus).  Pending job ids continue after the running
    jobs' ids (creation order); pending priority 50."""
    run, pend = case["running"], case["pending"]
    names = sorted({j["user"] for j in run} | {j["user"] for j in pend})
    uid = {n: i for i, n in enumerate(names)}
    nu = len(names)
    div = {k: np.full(nu, case["share"][k]) for k in ("mem", "cpus")}
    for u, sh in case["shares"].items():
        for k, v in sh.items():
            div[k][uid[u]] = v
    users = abi.make_users(nu, name_rank=np.arange(nu, dtype=np.int32), div_mem=div["mem"],
                           div_cpus=div["cpus"], div_gpus=np.full(nu, 1.0))
    hostnames = sorted({j["host"] for j in run} | set(case["spare"].keys()))
    hid = {h: i for i, h in enumerate(hostnames)}
    nh = len(hostnames)
    R = len(run)
    t = abi.make_tasks(user=np.array([uid[j["user"]] for j in run], np.int32),
                       priority=np.full(R, 50, np.int32), start_time=np.full(R, T0, np.int64),
                       task_id=np.arange(1000, 1000 + R, dtype=np.int64),
                       job_id=np.arange(1, R + 1, dtype=np.int64),
                       cpus=np.array([j["cpus"] for j in run]), mem=np.array([j["mem"] for j in run]))
    running = abi.RunningSoA(t=t, host=np.array([hid[j["host"]] for j in run], np.int32))
    P = len(pend)
    jobs = abi.JobsSoA(n=P, user=np.array([uid[j["user"]] for j in pend], np.int32),
                       cpus=np.array([j["cpus"] for j in pend]), mem=np.array([j["mem"] for j in pend]),
                       gpus=np.zeros(P))
    has_spare = np.zeros(nh, np.uint8)
    sc, sm = np.zeros(nh), np.zeros(nh)
    for h, (m, c) in case["spare"].items():
        has_spare[hid[h]] = 1
        sm[hid[h]] = m
        sc[hid[h]] = c
    hosts = abi.HostTable(n=nh, hostname_id=np.arange(nh, dtype=np.int32),
                          name_rank=np.arange(nh, dtype=np.int32), has_spare=has_spare,
                          spare_cpus=sc, spare_mem=sm, spare_gpus=np.zeros(nh), n_attr_cols=0)
    p = case["params"]
    params = abi.RebalanceParams(p["max_preemption"], p["min_dru_diff"], p["safe_dru_threshold"], 0)
    return dict(running=running, pending=jobs, pending_job_id=np.arange(R + 1, R + 1 + P, dtype=np.int64),
                pending_priority=np.full(P, 50, np.int32), hosts=hosts, users=users, params=params,
                hostnames=hostnames)


def check_rebalance_case(case, eng):
    inp = rebalance_inputs(case)
    out = eng.rebalance(inp["running"], inp["pending"], inp["pending_job_id"], inp["pending_priority"],
                        inp["hosts"], inp["users"], inp["params"])
    if "expect" in case:
        assert len(out) == len(case["expect"]), (case["name"], out)
        for d, e in zip(out, case["expect"]):
            assert inp["hostnames"][d["host"]] == e["host"], (case["name"], d)
            assert d["dru"] == e["dru"] or abs(d["dru"] - e["dru"]) <= 1e-12 * abs(e["dru"]), (case["name"], d)
            assert d["victims"] == e["victims"] and d["mem"] == e["mem"] and d["cpus"] == e["cpus"], (case["name"], d)
    else:
        assert [d["pending_idx"] for d in out] == case["expect_run"], (case["name"], out)
        assert [v for d in out for v in d["victims"]] == case["expect_preempt"], (case["name"], out)
    return out
