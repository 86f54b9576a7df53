"""Phase profile of the matcher on one pool of a BASELINE config (run on the GPU box):
python tools/prof_match_cfg.py c5 5   -> kernel ms, placements, then the COOK_PROF counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cook_b200 import traces  # noqa: E402
from cook_b200.engine import GpuEngine  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
p = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t = traces.gen_config_pool(cfg, p)
eng = GpuEngine()
ranked = eng.rank(t["running"], t["pending"], t["users"])["ranked"]
prm = traces.match_params(t["jobs"].n, host_lifetime_mins=t["host_lifetime_mins"])
os.environ.pop("COOK_PROF", None)
ref = None
for env in sys.argv[3:] or [""]:          # e.g. COOK_MATCH_TARGET=64,COOK_MATCH_BMIN=128
    for kv in env.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            os.environ[k] = v
    ms = []
    for _ in range(3):
        m = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
        s = m["stats"]
        ms.append(round(s["ms_match_kernel"], 3))
    if ref is None:
        ref = m["assign"].copy()
    print(cfg, p, "env", env, "jobs", t["jobs"].n, "offers", t["offers"].n, "kernel_ms", ms,
          "same" if (m["assign"] == ref).all() else "DIFFERENT",
          {k: s[k] for k in ("n_considerable", "n_matched", "n_fast", "n_chunk_rescan", "n_full_rescan", "n_offers_used")}, flush=True)
if not os.environ.get("PROF_DETAIL"):
    sys.exit(0)
os.environ["COOK_PROF"] = "1"
m = eng.match(ranked, t["jobs"], t["offers"], t["users"], prm, groups=t["groups"], max_ports=2)
print("prof_kernel_ms", round(m["stats"]["ms_match_kernel"], 3), flush=True)
