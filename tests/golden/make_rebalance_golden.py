"""Rebalancer known-answer vectors transcribed BY HAND from
/root/reference/scheduler/test/cook/test/rebalancer.clj (K19: :208-440
compute-preemption-decision; K22: :1028-1140 rebalance).  Job/task ids follow
creation order; every running task is created with start-time "now" (equal), so
per-user order = task id order.  Rewrites tests/golden/rebalance_golden.json."""
import json
import os

DMAX = 1.7976931348623157e308


def rj(user, mem, cpus, host):
    return dict(user=user, mem=float(mem), cpus=float(cpus), host=host)


def pj(user, mem, cpus):
    return dict(user=user, mem=float(mem), cpus=float(cpus))


CASES = []

# ---- K19 --------------------------------------------------------------------
K19_RUN = [rj("ljin", 10, 10, "hostA"), rj("ljin", 5, 5, "hostA"), rj("ljin", 15, 25, "hostB"),
           rj("ljin", 25, 15, "hostB"), rj("wzhao", 10, 10, "hostA"), rj("wzhao", 10, 10, "hostB")]
T3, T4, T7, T8 = 2, 3, 4, 5  # indices into K19_RUN
J9, J10, J11 = pj("wzhao", 15, 15), pj("sunil", 15, 15), pj("ljin", 15, 15)
J12, J13, J14 = pj("sunil", 40, 40), pj("sunil", 45, 45), pj("sunil", 80, 80)


def k19(n, line, job, spare, min_diff, expect):
    CASES.append(dict(name=f"K19-{n}", cite=f"test/cook/test/rebalancer.clj:{line}", running=K19_RUN,
                      pending=[job], spare=spare, share=dict(mem=25.0, cpus=25.0), shares={},
                      params=dict(max_preemption=1, min_dru_diff=min_diff, safe_dru_threshold=1.0),
                      expect=[expect] if expect else []))


def dec(host, dru, tasks, mem, cpus):
    return dict(pending=0, host=host, dru=dru, victims=tasks, mem=float(mem), cpus=float(cpus), gpus=0.0)


k19(1, "257-268", J9, {}, 0.05, dec("hostB", 2.2, [T4], 25, 15))
k19(2, "270-281", J9, {"hostB": (15, 15)}, 0.5, dec("hostB", DMAX, [], 15, 15))
k19(3, "283-295", J9, {"hostA": (20, 20), "hostB": (10, 10)}, 0.5, dec("hostA", DMAX, [], 20, 20))
k19(4, "297-309", J9, {"hostA": (10, 10), "hostB": (10, 10)}, 0.0, dec("hostB", 2.2, [T4], 35, 25))
k19(5, "311-322", J10, {}, 0.5, dec("hostB", 2.2, [T4], 25, 15))
k19(6, "324-335", J11, {}, 0.5, None)
k19(7, "336-347", J12, {}, 0.0, None)
k19(8, "348-359", J12, {"hostA": (40, 40)}, 0.5, dec("hostA", DMAX, [], 40, 40))
k19(9, "360-371", J12, {"hostA": (35, 35)}, 0.0, None)
k19(10, "372-384", J12, {"hostA": (35, 35), "hostB": (30, 30)}, 0.5, dec("hostB", 2.2, [T4], 55, 45))
k19(11, "385-396", J13, {}, 0.5, None)
k19(12, "397-408", J13, {}, 2.0, None)
k19(13, "409-420", J14, {}, 0.5, None)

# ---- K22 --------------------------------------------------------------------
K22_RUN = [rj("ljin", 10, 10, "hostA"), rj("ljin", 5, 5, "hostA"), rj("ljin", 15, 25, "hostB"),
           rj("ljin", 25, 15, "hostB"), rj("wzhao", 8, 8, "hostA"), rj("wzhao", 10, 10, "hostB"),
           rj("wzhao", 10, 10, "hostA"), rj("wzhao", 10, 10, "hostB")]
W = [pj("wzhao", 5, 5) for _ in range(10)]
S = [pj("sunil", 5, 5) for _ in range(10)]


def k22(name, jobs, spare, run, preempt, shares=None):
    CASES.append(dict(name=f"K22-{name}", cite="test/cook/test/rebalancer.clj:1078-1140", running=K22_RUN,
                      pending=jobs, spare=spare, share=dict(mem=25.0, cpus=25.0), shares=shares or {},
                      params=dict(max_preemption=128, min_dru_diff=0.0, safe_dru_threshold=1.0),
                      expect_run=run, expect_preempt=preempt))


k22("simple", W, {}, [0, 1, 2], [3])
k22("simple-with-available", W, {"hostB": (0.0, 10.0)}, [0, 1, 2, 3, 4], [3])   # spare {:mem 0 :cpus 10}
k22("simple-2", S, {}, list(range(8)), [3, 2])
k22("simple-2-with-available", S, {"hostB": (25.0, 25.0)}, list(range(8)), [3])
k22("share-change", S, {}, list(range(10)), [3, 2, 7], shares={"sunil": dict(mem=50.0, cpus=50.0)})

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rebalance_golden.json")
    with open(out, "w") as f:
        json.dump(CASES, f, indent=1)
    print("wrote", out, len(CASES))
