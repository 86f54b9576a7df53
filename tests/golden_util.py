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
