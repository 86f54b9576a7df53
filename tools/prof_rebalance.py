"""Phase timing of cook_rebalance on a config-#4-shaped pool (run on the GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cook_b200 import abi, traces  # noqa: E402
from cook_b200.engine import GpuEngine  # noqa: E402

eng = GpuEngine()
for nr, npend, nh, nu, mp in ((100_000, 400, 5_000, 2_000, 64), (160_000, 640, 8_000, 2_000, 128)):
    t = traces.gen_rebalance(4, nr, npend, nh, nu, max_preemption=mp)
    args = (t["running"], t["pending"], t["pending_job_id"], t["pending_priority"], t["hosts"], t["users"], t["params"])
    eng.rebalance(*args, groups=t["groups"])
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        d = eng.rebalance(*args, groups=t["groups"])
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    s = eng.last_stats(abi.PHASE_REBALANCE)
    print(dict(running=nr, pending=npend, hosts=nh, max_preemption=mp, decisions=len(d),
               last_pending_idx=d[-1]["pending_idx"] if d else None, e2e_ms=round(best, 3),
               **{k: round(v, 3) if isinstance(v, float) else v for k, v in s.items()}), flush=True)
