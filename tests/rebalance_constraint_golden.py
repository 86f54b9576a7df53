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
