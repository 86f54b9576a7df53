"""N-GPU check of the exchange step (run under torchrun on the GPU box, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py

Every rank owns one pool of ONE quota group.  After its match round it calls cook_exchange_usage
(device-side delta + ncclAllGather through the library's own communicator); the gathered table must
equal what every rank computes on the host for its own pool (cross-checked through a second,
independent all-gather of the host results), and the next cook_rank - fed with the gathered group
usage - must equal the oracle fed with the same totals.  Prints MULTIGPU_OK on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cook_b200 import abi, sharding, traces  # noqa: E402
from cook_b200.engine import GpuEngine, comm_init, comm_unique_id, load_library  # noqa: E402
from oracle.pyoracle import OracleEngine  # noqa: E402

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_library()
box = [comm_unique_id(lib) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
comm = comm_init(lib, box[0], rank, world, local)

NU = 80
t = traces.gen_pool(500 + rank, 12_000 + 1000 * rank, 500, NU, 2_000)
eng = GpuEngine(pool_name=f"grp-{rank}", device=local)
ora = OracleEngine()
ranked = eng.rank(t["running"], t["pending"], t["users"])["ranked"]
m = eng.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(t["jobs"].n))
mo = ora.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(t["jobs"].n))
assert np.array_equal(m["assign"], mo["assign"]), "assignments differ from the oracle"
g = eng.exchange_usage(NU + 16, comm=comm, world=world)            # [world, NU + 16, 4], padded
j = t["jobs"]
mine = sharding.usage_delta(m["considerable"], m["assign"], j.col("user"), j.col("cpus"), j.col("mem"), j.col("gpus"), NU)
allh = torch.zeros(world, NU, 4, dtype=torch.float64, device="cuda")
dist.all_gather_into_tensor(allh.view(-1), torch.from_numpy(mine).cuda().reshape(-1))
allh = allh.cpu().numpy()
assert np.array_equal(g[:, :NU], allh), "library all-gather differs from the host deltas"
assert not g[:, NU:].any()
assert np.array_equal(g[rank, :NU], mine) and mine[:, 0].sum() == m["stats"]["n_matched"] > 0
tot = g.sum(axis=(0, 1))
gq = abi.make_pool_quota({"count": tot[0] + 800, "cpus": tot[1] + 2500, "mem": tot[2] + 1.0e7, "gpus": 1e9})
rg = eng.rank(t["running"], t["pending"], t["users"], group_quota=gq, group_usage=tot)
ro = ora.rank(t["running"], t["pending"], t["users"], group_quota=gq, group_usage=allh.sum(axis=(0, 1)))
r0 = ora.rank(t["running"], t["pending"], t["users"], group_quota=gq, group_usage=np.zeros(4))
assert np.array_equal(rg["ranked"], ro["ranked"]) and 0 < len(rg["ranked"]) < len(r0["ranked"])
# the batched form (one all-gather per cycle for all of a rank's pools): rank r owns 1 + (r % 2) pools,
# two slots everywhere; slot 0 = the pool above, slot 1 = a second small pool on the odd ranks, zeros elsewhere
from cook_b200.engine import exchange_usage_batch
engs = [eng]
want2 = np.zeros((NU, 4))
if rank % 2 == 1:
    t2 = traces.gen_pool(700 + rank, 4_000, 200, NU, 500)
    e2 = GpuEngine(pool_name=f"grp-{rank}-b", device=local)
    r2 = e2.rank(t2["running"], t2["pending"], t2["users"])["ranked"]
    m2 = e2.match(r2, t2["jobs"], t2["offers"], t2["users"], traces.match_params(t2["jobs"].n))
    j2 = t2["jobs"]
    want2 = sharding.usage_delta(m2["considerable"], m2["assign"], j2.col("user"), j2.col("cpus"), j2.col("mem"), j2.col("gpus"), NU)
    engs.append(e2)
eng.match(ranked, t["jobs"], t["offers"], t["users"], traces.match_params(t["jobs"].n))   # the round whose delta travels
gb = exchange_usage_batch(engs, NU + 16, n_slots=2, comm=comm, world=world)               # [world, 2, NU + 16, 4]
all2 = torch.zeros(world, NU, 4, dtype=torch.float64, device="cuda")
dist.all_gather_into_tensor(all2.view(-1), torch.from_numpy(want2).cuda().reshape(-1))
assert np.array_equal(gb[:, 0, :NU], allh), "batched exchange: slot 0 differs"
assert np.array_equal(gb[:, 1, :NU], all2.cpu().numpy()), "batched exchange: slot 1 differs"
assert not gb[:, :, NU:].any()
for e in engs[1:]:
    e.close()
s = eng.last_stats(abi.PHASE_EXCHANGE)
ok = torch.ones(1, device="cuda")
dist.all_reduce(ok)
if rank == 0:
    print(f"MULTIGPU_OK world={world} batched=ok exchange_ms={s['ms_device']:.3f} launches={s['n_launches']} "
          f"placed={[int(x) for x in g[:, :, 0].sum(axis=1)]}", flush=True)
eng.close()
lib.cook_comm_destroy(comm)
dist.destroy_process_group()
