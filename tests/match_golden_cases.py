"""Matcher known answers transcribed BY HAND from the reference's tests
(test/cook/test/scheduler/scheduler.clj): K10 :547-584, K12 :660-706,
K13 :982-1155.  These pin the Fenzo restatement at SET level (which jobs are
placed / how many per host) — the reference has no host-level golden data for
multi-host matches."""
import numpy as np

from cook_b200 import abi, traces


def build(jobs, offers, groups=None, n_attr_cols=0):
    """jobs: dicts(cpus, mem, groups=[...]); offers: dicts(host, cpus, mem, attrs={col: val}, run=...)"""
    J, O = len(jobs), len(offers)
    hostnames = sorted({o["host"] for o in offers} | {h for g in (groups or []) for h, _ in g.get("cotasks", [])})
    hid = {h: i for i, h in enumerate(hostnames)}
    users = abi.make_users(1)
    goff, gidx = abi.csr([j.get("groups", []) for j in jobs])
    jb = abi.JobsSoA(n=J, user=np.zeros(J, np.int32), cpus=np.array([j["cpus"] for j in jobs], float),
                     mem=np.array([j["mem"] for j in jobs], float), gpus=np.zeros(J),
                     allowed=np.ones(J, np.uint8), plugin_accept=np.ones(J, np.uint8),
                     group_off=goff if groups else None, group_idx=gidx if groups else None)
    attr = np.zeros((max(n_attr_cols, 1), O), np.int32)
    for i, o in enumerate(offers):
        for c, v in o.get("attrs", {}).items():
            attr[c, i] = v
    order = sorted(range(O), key=lambda i: offers[i]["host"])
    rank = np.zeros(O, np.int32)
    for r, i in enumerate(order):
        rank[i] = r
    of = abi.OffersSoA(n=O, hostname_id=np.array([hid[o["host"]] for o in offers], np.int32), name_rank=rank,
                       cpus=np.array([o["cpus"] for o in offers], float),
                       mem=np.array([o["mem"] for o in offers], float),
                       run_cpus=np.array([o.get("run_cpus", 0.0) for o in offers], float),
                       run_mem=np.array([o.get("run_mem", 0.0) for o in offers], float),
                       run_count=np.array([o.get("run_count", 0) for o in offers], np.int32),
                       n_attr_cols=n_attr_cols, attr=attr.reshape(-1) if n_attr_cols else None)
    gr = None
    if groups:
        coff, chost = abi.csr([[hid[h] for h, _ in g.get("cotasks", [])] for g in groups])
        _, cattr = abi.csr([[a for _, a in g.get("cotasks", [])] for g in groups])
        gr = abi.Groups(n_groups=len(groups), kind=np.array([g["kind"] for g in groups], np.int32),
                        attr_col=np.array([g.get("col", -1) for g in groups], np.int32),
                        minimum=np.array([g.get("minimum", 0) for g in groups], np.int32),
                        cot_off=coff, cot_hostname_id=chost, cot_attr_val=cattr)
    return dict(jobs=jb, offers=of, users=users, groups=gr, hid=hid, hostnames=hostnames)


def run(eng, c):
    J = c["jobs"].n
    return eng.match(np.arange(J, dtype=np.int32), c["jobs"], c["offers"], c["users"],
                     traces.match_params(J), groups=c["groups"])


def check_all(eng):
    # ---- K10 scheduler.clj:547-584: four 1-cpu/1000-MB jobs against ONE offer
    four = [dict(cpus=1.0, mem=1000.0) for _ in range(4)]
    for (cpus, mem), n in [((0, 0), 0), ((0.5, 100), 0), ((0.5, 1000), 0), ((1, 500), 0),
                           ((1, 1000), 1), ((1.5, 1500), 1), ((4, 4000), 4), ((5, 5000), 4)]:
        m = run(eng, build(four, [dict(host="h", cpus=float(cpus), mem=float(mem))]))
        assert int((m["assign"] >= 0).sum()) == n, ("K10", cpus, mem, m["assign"])
    assert run(eng, build([], [dict(host="h", cpus=2.0, mem=2000.0)]))["stats"]["n_matched"] == 0
    # ---- K12 :660-706: list order is respected (high priority first, 1 cpu offer)
    nine = [dict(cpus=1.0, mem=1000.0) for _ in range(9)]
    m = run(eng, build(nine, [dict(host="empty_host", cpus=1.0, mem=200000.0)]))
    assert list(m["assign"]) == [0] + [-1] * 8, ("K12", m["assign"])
    # ---- K13 unique, same cycle :1020-1046: two jobs of a unique group, one host -> 1 placed
    g = [dict(kind=abi.GROUP_UNIQUE)]
    two = [dict(cpus=1.0, mem=10.0, groups=[0]) for _ in range(2)]
    m = run(eng, build(two, [dict(host="test-host", cpus=100.0, mem=100000.0)], groups=g))
    assert list(m["assign"]) == [0, -1], ("K13-unique-same-cycle", m["assign"])
    # unique, different cycles :988-1018: the cotask already runs on the host -> 0 placed
    g = [dict(kind=abi.GROUP_UNIQUE, cotasks=[("test-host", 0)])]
    one = [dict(cpus=1.0, mem=10.0, groups=[0])]
    m = run(eng, build(one, [dict(host="test-host", cpus=100.0, mem=100000.0, run_cpus=1.0, run_mem=10.0,
                                  run_count=1)], groups=g))
    assert list(m["assign"]) == [-1], ("K13-unique-cross-cycle", m["assign"])
    # no group: both fit :1047-1054
    m = run(eng, build([dict(cpus=1.0, mem=10.0)], [dict(host="test-host", cpus=100.0, mem=100000.0)]))
    assert list(m["assign"]) == [0]
    # ---- K13 balanced :1056-1070: 9 jobs, HOSTNAME balanced minimum 3, 3 hosts -> 3/3/3
    hosts = ["straw", "sticks", "bricks"]
    offers = [dict(host=h, cpus=100.0, mem=100000.0, attrs={0: i + 1}) for i, h in enumerate(hosts)]
    g = [dict(kind=abi.GROUP_BALANCED, col=0, minimum=3)]
    jobs9 = [dict(cpus=1.0, mem=10.0, groups=[0]) for _ in range(9)]
    m = run(eng, build(jobs9, offers, groups=g, n_attr_cols=1))
    assert sorted(np.bincount(m["assign"], minlength=3)) == [3, 3, 3], ("K13-balanced", m["assign"])
    # without the constraint the assignment is NOT balanced :1071-1081 (bin packing piles up)
    m = run(eng, build([dict(cpus=1.0, mem=10.0) for _ in range(9)], offers, n_attr_cols=1))
    assert sorted(np.bincount(m["assign"], minlength=3)) != [3, 3, 3]
    # ---- K13 attribute-equals :1083-1145: cotask runs on an "east" host; 20 jobs vs
    # 20 one-cpu "west" offers + one 5-cpu "east" offer -> 5 placed on east, 15 fail
    EAST, WEST = 1, 2
    g = [dict(kind=abi.GROUP_ATTR_EQUALS, col=0, cotasks=[("first-east-host", EAST)])]
    offers = [dict(host=f"west-{i:02d}", cpus=1.0, mem=100000.0, attrs={0: WEST}) for i in range(20)]
    offers.append(dict(host="east-big", cpus=5.0, mem=100000.0, attrs={0: EAST}))
    jobs20 = [dict(cpus=1.0, mem=10.0, groups=[0]) for _ in range(20)]
    m = run(eng, build(jobs20, offers, groups=g, n_attr_cols=1))
    placed = m["assign"][m["assign"] >= 0]
    assert len(placed) == 5 and set(placed) == {20}, ("K13-attr-equals", m["assign"])
    assert int((m["assign"] < 0).sum()) == 15
    # no constraint: all 20 are placed using the other offers too :1137-1145
    offers15 = offers[:15] + [offers[20]]
    m = run(eng, build([dict(cpus=1.0, mem=10.0) for _ in range(20)], offers15, n_attr_cols=1))
    assert int((m["assign"] >= 0).sum()) == 20
    # ---- K15 :1957-1964: 4 jobs (3/2048, 13/1024, 7/4096, 11/1024) on 3 offers
    # (10/2048, 20/16384, 30/8192): all 4 launched using 3 offers
    jobs4 = [dict(cpus=3.0, mem=2048.0), dict(cpus=13.0, mem=1024.0), dict(cpus=7.0, mem=4096.0),
             dict(cpus=11.0, mem=1024.0)]
    offers3 = [dict(host="h1", cpus=10.0, mem=2048.0), dict(host="h2", cpus=20.0, mem=16384.0),
               dict(host="h3", cpus=30.0, mem=8192.0)]
    m = run(eng, build(jobs4, offers3))
    assert int((m["assign"] >= 0).sum()) == 4 and m["stats"]["n_offers_used"] == 3, ("K15", m["assign"])
