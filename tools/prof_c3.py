"""Kernel time of one config-#3-shaped pool (constraints, groups, ports, gpus) on the GPU box.
COOK_GPU_LIB selects an A/B build of the library."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cook_b200 import traces  # noqa: E402
from cook_b200.engine import GpuEngine  # noqa: E402

nj, no = int(sys.argv[1]) if len(sys.argv) > 1 else 40000, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
t = traces.gen_c3_pool(300, nj, no, 500, nj // 5)
eng = GpuEngine()
ranked = eng.rank(t["running"], t["pending"], t["users"])["ranked"]
prm = traces.match_params(nj, host_lifetime_mins=t["host_lifetime_mins"])
ms = []
for _ in range(3):
    m = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
    ms.append(round(m["stats"]["ms_match_kernel"], 3))
print("c3 pool", nj, "x", no, "kernel_ms", ms, {k: m["stats"][k] for k in ("n_matched", "n_fast", "n_chunk_rescan", "n_full_rescan")},
      "evals/s %.3g" % (m["stats"]["evals"] / (ms[-1] / 1e3)), flush=True)
