"""Rebalancer decisions under host constraints, transcribed BY HAND from the reference
(test/cook/test/rebalancer.clj:441-597).  Four hosts each filled by one
preemptable 100-cpu task of a different user ("pig1".."pig4"), default share mem 10 / cpus 10;
the pending job of user "diego" asks 1 cpu / 10 MB (create-dummy-job defaults)."""
import numpy as np

from cook_b200 import abi

HOSTS = ["bricks", "rebar", "sticks", "straw"]          # sorted: index = hostname rank
PIG_HOST = {"pig1": "straw", "pig2": "sticks", "pig3": "bricks", "pig4": "rebar"}
USERS = ["diego", "pig1", "pig2", "pig3", "pig4"]
T0 = 1_600_000_000_000


AZ = {"east": 1, "west": 2, "south": 3}


def _decide_attr_equals(eng, steel_az):
    """attribute-equals group on "az" (:598-672): straw/sticks/bricks are east, rebar west; the group's
    running member sits on a fifth host "steel" whose az is `steel_az`."""
    hosts5 = ["bricks", "rebar", "steel", "sticks", "straw"]
    az = {"bricks": "east", "rebar": "west", "steel": steel_az, "sticks": "east", "straw": "east"}
    hid = {h: i for i, h in enumerate(hosts5)}
    uid = {u: i for i, u in enumerate(USERS)}
    run = [(u, 100.0, 10.0, h) for u, h in PIG_HOST.items()] + [("diego", 1.0, 10.0, "steel")]
    R = len(run)
    t = abi.make_tasks(user=np.array([uid[r[0]] for r in run], np.int32), priority=np.full(R, 50, np.int32),
                       start_time=np.full(R, T0, np.int64), task_id=np.arange(1000, 1000 + R, dtype=np.int64),
                       job_id=np.arange(1, R + 1, dtype=np.int64), cpus=np.array([r[1] for r in run]),
                       mem=np.array([r[2] for r in run]))
    running = abi.RunningSoA(t=t, host=np.array([hid[r[3]] for r in run], np.int32))
    off, gi = abi.csr([[0]])
    jobs = abi.JobsSoA(n=1, user=np.array([uid["diego"]], np.int32), cpus=np.array([1.0]), mem=np.array([10.0]),
                       gpus=np.zeros(1), group_off=off, group_idx=gi)
    coff, chost = abi.csr([[hid["steel"]]])
    _, cattr = abi.csr([[AZ[steel_az]]])
    groups = abi.Groups(n_groups=1, kind=np.array([abi.GROUP_ATTR_EQUALS], np.int32), attr_col=np.array([0], np.int32),
                        minimum=np.zeros(1, np.int32), cot_off=coff, cot_hostname_id=chost, cot_attr_val=cattr)
    nh = len(hosts5)
    hosts = abi.HostTable(n=nh, hostname_id=np.arange(nh, dtype=np.int32), name_rank=np.arange(nh, dtype=np.int32),
                          has_spare=np.zeros(nh, np.uint8), spare_cpus=np.zeros(nh), spare_mem=np.zeros(nh),
                          spare_gpus=np.zeros(nh), n_attr_cols=1,
                          attr=np.array([AZ[az[h]] for h in hosts5], np.int32))
    users = abi.make_users(len(USERS), div_mem=np.full(len(USERS), 10.0), div_cpus=np.full(len(USERS), 10.0),
                           div_gpus=np.full(len(USERS), 1.0))
    out = eng.rebalance(running, jobs, np.array([R + 1], np.int64), np.array([50], np.int32), hosts, users,
                        abi.RebalanceParams(1, 0.05, 1.0, 0), groups=groups)
    return hosts5[out[0]["host"]] if out else None


def _decide(eng, diego_running_hosts=(), novel=None, group_kind=None):
    hid = {h: i for i, h in enumerate(HOSTS)}
    uid = {u: i for i, u in enumerate(USERS)}
    run = [(u, 100.0, 10.0, h) for u, h in PIG_HOST.items()]
    run += [("diego", 1.0, 10.0, h) for h in diego_running_hosts]   # group members already running
    R = len(run)
    t = abi.make_tasks(user=np.array([uid[r[0]] for r in run], np.int32), priority=np.full(R, 50, np.int32),
                       start_time=np.full(R, T0, np.int64), task_id=np.arange(1000, 1000 + R, dtype=np.int64),
                       job_id=np.arange(1, R + 1, dtype=np.int64), cpus=np.array([r[1] for r in run]),
                       mem=np.array([r[2] for r in run]))
    running = abi.RunningSoA(t=t, host=np.array([hid[r[3]] for r in run], np.int32))
    jkw = dict(n=1, user=np.array([uid["diego"]], np.int32), cpus=np.array([1.0]), mem=np.array([10.0]),
               gpus=np.zeros(1))
    groups = None
    if novel is not None:
        off, hosts = abi.csr([[hid[h] for h in novel]])
        jkw.update(novel_off=off, novel_host=hosts)
    if group_kind is not None:
        off, gi = abi.csr([[0]])
        jkw.update(group_off=off, group_idx=gi)
        coff, chost = abi.csr([[hid[h] for h in diego_running_hosts]])
        _, cattr = abi.csr([[0 for _ in diego_running_hosts]])
        groups = abi.Groups(n_groups=1, kind=np.array([group_kind], np.int32), attr_col=np.array([-1], np.int32),
                            minimum=np.zeros(1, np.int32), cot_off=coff, cot_hostname_id=chost, cot_attr_val=cattr)
    nh = len(HOSTS)
    hosts = abi.HostTable(n=nh, hostname_id=np.arange(nh, dtype=np.int32), name_rank=np.arange(nh, dtype=np.int32),
                          has_spare=np.zeros(nh, np.uint8), spare_cpus=np.zeros(nh), spare_mem=np.zeros(nh),
                          spare_gpus=np.zeros(nh), n_attr_cols=0)
    users = abi.make_users(len(USERS), div_mem=np.full(len(USERS), 10.0), div_cpus=np.full(len(USERS), 10.0),
                           div_gpus=np.full(len(USERS), 1.0))
    out = eng.rebalance(running, abi.JobsSoA(**jkw), np.array([R + 1], np.int64), np.array([50], np.int32),
                        hosts, users, abi.RebalanceParams(1, 0.05, 1.0, 0), groups=groups)
    return HOSTS[out[0]["host"]] if out else None


def check_all(eng):
    # novel-host: the job already failed on straw, sticks, bricks (:461-484, :485-512)
    assert _decide(eng, novel=["straw", "sticks", "bricks"]) == "rebar"
    # unconstrained job gets some host (:513-531); with equal DRUs `max-key` keeps the last = greatest hostname
    assert _decide(eng) == "straw"
    # unique host-placement group: members run on all hosts but rebar (:552-574) / on all hosts (:575-597)
    assert _decide(eng, diego_running_hosts=["straw", "sticks", "bricks"], group_kind=abi.GROUP_UNIQUE) == "rebar"
    assert _decide(eng, diego_running_hosts=["straw", "sticks", "bricks", "rebar"],
                   group_kind=abi.GROUP_UNIQUE) is None
    # attribute-equals group: the running member is in az west => only rebar qualifies (:619-645);
    # in az south => no host qualifies (:646-672)
    assert _decide_attr_equals(eng, "west") == "rebar"
    assert _decide_attr_equals(eng, "south") is None


# ---- K20, second half (test/cook/test/rebalancer.clj:673-811): balanced groups and the quota rule
B_HOSTS = ["bricks", "concrete", "gold", "rebar", "steel", "sticks", "straw", "titanium"]   # sorted
B_AZ = {"straw": "east", "sticks": "west", "bricks": "south", "rebar": "north", "concrete": "east",
        "steel": "west", "gold": "south", "titanium": "north"}
B_PIGS = ["straw", "sticks", "bricks", "rebar", "concrete", "steel", "gold", "titanium"]    # pig1..pig8
AZ4 = {"east": 1, "west": 2, "south": 3, "north": 4}


def _balanced_inputs(group_hosts, extra_pending_user=None):
    """Eight hosts each filled by one 200-cpu task of pig1..pig8, share mem 20 / cpus 20; diego's
    group (balanced on "az", minimum 4) has one 1-cpu member running per entry of `group_hosts`;
    diego's pending job (1 cpu / 10 MB) is in the group.  `extra_pending_user`: an ungrouped
    job of that user walks first (used to preempt a host earlier in the same cycle)."""
    users = ["diego"] + [f"pig{i}" for i in range(1, 9)]
    uid = {u: i for i, u in enumerate(users)}
    hid = {h: i for i, h in enumerate(B_HOSTS)}
    run = [(f"pig{i + 1}", 200.0, 10.0, h) for i, h in enumerate(B_PIGS)] + [("diego", 1.0, 10.0, h) for h in group_hosts]
    R = len(run)
    t = abi.make_tasks(user=np.array([uid[r[0]] for r in run], np.int32), priority=np.full(R, 50, np.int32),
                       start_time=np.full(R, T0, np.int64), task_id=np.arange(1000, 1000 + R, dtype=np.int64),
                       job_id=np.arange(1, R + 1, dtype=np.int64), cpus=np.array([r[1] for r in run]),
                       mem=np.array([r[2] for r in run]))
    running = abi.RunningSoA(t=t, host=np.array([hid[r[3]] for r in run], np.int32))
    pend_users = ([extra_pending_user] if extra_pending_user else []) + ["diego"]
    P = len(pend_users)
    off, gi = abi.csr([[] for _ in pend_users[:-1]] + [[0]])
    jobs = abi.JobsSoA(n=P, user=np.array([uid[u] for u in pend_users], np.int32), cpus=np.ones(P),
                       mem=np.full(P, 10.0), gpus=np.zeros(P), group_off=off, group_idx=gi)
    coff, chost = abi.csr([[hid[h] for h in group_hosts]])
    _, cattr = abi.csr([[AZ4[B_AZ[h]] for h in group_hosts]])
    groups = abi.Groups(n_groups=1, kind=np.array([abi.GROUP_BALANCED], np.int32), attr_col=np.array([0], np.int32),
                        minimum=np.array([4], np.int32), cot_off=coff, cot_hostname_id=chost, cot_attr_val=cattr)
    nh = len(B_HOSTS)
    hosts = abi.HostTable(n=nh, hostname_id=np.arange(nh, dtype=np.int32), name_rank=np.arange(nh, dtype=np.int32),
                          has_spare=np.zeros(nh, np.uint8), spare_cpus=np.zeros(nh), spare_mem=np.zeros(nh),
                          spare_gpus=np.zeros(nh), n_attr_cols=1,
                          attr=np.array([AZ4[B_AZ[h]] for h in B_HOSTS], np.int32))
    nu = len(users)
    ut = abi.make_users(nu, div_mem=np.full(nu, 20.0), div_cpus=np.full(nu, 20.0), div_gpus=np.full(nu, 1.0))
    return dict(running=running, jobs=jobs, job_ids=np.arange(R + 1, R + 1 + P, dtype=np.int64),
                prio=np.full(P, 50, np.int32), hosts=hosts, users=ut, groups=groups, hid=hid, R=R,
                params=abi.RebalanceParams(4, 0.05, 1.0, 0))


def check_balanced_and_quota(oracle):
    """Oracle only (the second case applies a given first decision through the oracle's test-only
    state view, as the reference test hands compute-preemption-decision a preempted task)."""
    full = ["straw", "sticks", "bricks", "rebar", "concrete", "steel", "gold", "titanium"]
    # :695-733 every az has 4 group members except north (3): only rebar / titanium qualify
    c = _balanced_inputs(full + full[:7])
    out = oracle.rebalance(c["running"], c["jobs"], c["job_ids"], c["prio"], c["hosts"], c["users"], c["params"],
                           groups=c["groups"])
    assert len(out) == 1 and B_HOSTS[out[0]["host"]] in ("rebar", "titanium"), out
    # :734-774 north and south have 3 members, but titanium (north) was preempted earlier in this
    # cycle, which counts towards north: only bricks / gold remain
    c = _balanced_inputs(full + full[:6], extra_pending_user="pig1")
    pig8 = B_PIGS.index("titanium")          # running task index of the pig on titanium
    tr = oracle.rebalance_trace(c["running"], c["jobs"], c["job_ids"], c["prio"], c["hosts"], c["users"], c["params"],
                                forced=[(0, c["hid"]["titanium"], [pig8], 10.0, 200.0, 0.0)], forced_only=False,
                                groups=c["groups"])
    d = tr["decisions"]
    assert len(d) == 2 and d[0]["victims"] == [pig8] and B_HOSTS[d[1]["host"]] in ("bricks", "gold"), d
    # :776-811 user over its count quota: only its own tasks may be preempted for it -> the 100-dru task of
    # testA, not the 200-dru task of testB
    uid = {"testA": 0, "testB": 1}
    t = abi.make_tasks(user=np.array([0, 1], np.int32), priority=np.array([1, 50], np.int32),
                       start_time=np.full(2, T0, np.int64), task_id=np.array([1000, 1001], np.int64),
                       job_id=np.array([1, 2], np.int64), cpus=np.array([100.0, 200.0]), mem=np.array([100.0, 200.0]))
    running = abi.RunningSoA(t=t, host=np.zeros(2, np.int32))
    jobs = abi.JobsSoA(n=1, user=np.array([uid["testA"]], np.int32), cpus=np.array([1.0]), mem=np.array([1.0]),
                       gpus=np.zeros(1))
    hosts = abi.HostTable(n=1, hostname_id=np.zeros(1, np.int32), name_rank=np.zeros(1, np.int32),
                          has_spare=np.zeros(1, np.uint8), spare_cpus=np.zeros(1), spare_mem=np.zeros(1),
                          spare_gpus=np.zeros(1), n_attr_cols=0)
    big = np.finfo(np.float64).max
    users = abi.make_users(2, div_mem=np.ones(2), div_cpus=np.ones(2), div_gpus=np.ones(2),
                           quota=dict(count=np.array([1.0, big]), cpus=np.full(2, big), mem=np.full(2, big),
                                      gpus=np.full(2, big)))
    out = oracle.rebalance(running, jobs, np.array([3], np.int64), np.array([50], np.int32), hosts, users,
                           abi.RebalanceParams(1, 0.5, 1.0, 0))
    assert len(out) == 1 and out[0]["host"] == 0 and out[0]["dru"] == 100.0 and out[0]["victims"] == [0], out
    return 3
