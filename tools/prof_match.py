"""Phase profile of the matcher on C2 (run on the GPU box): prints kernel ms without
instrumentation, then the clock64 phase breakdown (COOK_PROF=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cook_b200 import traces  # noqa: E402
from cook_b200.engine import GpuEngine  # noqa: E402

t = traces.gen_c2()
eng = GpuEngine()
ranked = eng.rank(t['running'], t['pending'], t['users'])['ranked']
res = traces.match_params(100000, reuse_resident=1)
ref = None
for env in sys.argv[1:] or [""]:
    for kv in env.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            os.environ[k] = v
    os.environ.pop('COOK_PROF', None)
    m = eng.match(ranked, t['jobs'], t['offers'], t['users'], traces.match_params(100000))
    if ref is None:
        ref = m['assign'].copy()
    ms = []
    for _ in range(3):
        m = eng.match(ranked, t['jobs'], t['offers'], t['users'], res)
        ms.append(round(m['stats']['ms_match_kernel'], 3))
    print("==", env, "kernel_ms", ms, "considerable_ms", round(m['stats']['ms_considerable'], 3),
          "same" if (m['assign'] == ref).all() else "DIFFERENT",
          {k: m['stats'][k] for k in ('n_matched', 'n_fast', 'n_chunk_rescan')}, flush=True)
    if os.environ.get("PROF_DETAIL"):
        os.environ['COOK_PROF'] = '1'
        m = eng.match(ranked, t['jobs'], t['offers'], t['users'], res)
        print("prof_kernel_ms", round(m['stats']['ms_match_kernel'], 3), flush=True)
