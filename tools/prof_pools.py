"""Four config-#3-shaped pools on ONE GPU: one after the other with the whole GPU each, or side by
side (one host thread per pool, max_ctas = 37 thread blocks each).  Prints wall time per round."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cook_b200 import traces  # noqa: E402
from cook_b200.engine import GpuEngine  # noqa: E402

nj, no = int(sys.argv[1]) if len(sys.argv) > 1 else 40000, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
pools = []
for p in range(4):
    t = traces.gen_c3_pool(500 + p, nj, no, 500, nj // 5)
    eng = GpuEngine(pool_name=f"pool-{p}")
    ranked = eng.rank(t["running"], t["pending"], t["users"])["ranked"]
    pools.append((t, ranked, eng))


def cycle(i, max_ctas, out):
    t, ranked, eng = pools[i]
    prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"], max_ctas=max_ctas)
    out[i] = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)


for mode, max_ctas in (("sequential, 148 blocks per pool", 0), ("side by side, 37 blocks per pool", 37)):
    ref = None
    for rep in range(3):
        out = [None] * 4
        t0 = time.perf_counter()
        if max_ctas == 0:
            for i in range(4):
                cycle(i, 0, out)
        else:
            th = [threading.Thread(target=cycle, args=(i, max_ctas, out)) for i in range(4)]
            [x.start() for x in th]
            [x.join() for x in th]
        dt = time.perf_counter() - t0
        evals = sum(o["stats"]["evals"] for o in out)
        kern = [round(o["stats"]["ms_match_kernel"], 1) for o in out]
        print(f"{mode}: wall {dt * 1e3:7.1f} ms for 4 pools ({evals / dt:.3g} evals/s), kernel ms per pool {kern}", flush=True)
