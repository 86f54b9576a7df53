"""K15: `handle-resource-offers!` known answers transcribed BY HAND from the reference's tests
(test/cook/test/scheduler/scheduler.clj:1806-2256: `test-handle-resource-helpers`,
`test-handle-resource-offers-mesos`, `test-handle-resource-offers-k8s`).  These run against the
REAL Fenzo in the reference, so they pin the Fenzo restatement (considerable filter + resource fit +
gpu / disk / reservation constraints + bin packing) at SET level: which jobs are launched and on how
many offers.  The reference's offer hostnames are random UUIDs, so every answer must hold for any
hostname order — each case is checked with the offers ranked in listed and in reverse order.

Fixture: eight pending jobs of one user in queue order (:1860-1925)
  job-1 3 cpus/2048 MB, job-2 13/1024, job-3 7/4096, job-4 11/1024, job-5 5/2048 + 2 gpus (p100),
  job-6 19/1024 + 4 gpus (default model p100), job-7 1/2048 + disk 250000 pd-ssd, job-8 2/2048 +
  disk 10000 pd-ssd; jobs without a disk request get the pool default 10000 MiB "standard" ->
  "pd-standard" (disk-config :1937-1945; only enforced on Kubernetes hosts);
  user usage {count 1, cpus 2, mem 1024}, quota {count 10, cpus 70, mem 32768, gpus 10} (:1929-1930).
Dictionary ids (the host shim's job): gpu models p100 = 0; disk types pd-standard = 0, pd-ssd = 1.
"""
import numpy as np

from cook_b200 import abi, traces

P100 = 0
PD_STANDARD, PD_SSD = 0, 1
JOBS = [  # (cpus, mem, gpus, disk_request, disk_type)
    (3, 2048, 0, 10000.0, PD_STANDARD), (13, 1024, 0, 10000.0, PD_STANDARD), (7, 4096, 0, 10000.0, PD_STANDARD),
    (11, 1024, 0, 10000.0, PD_STANDARD), (5, 2048, 2, 10000.0, PD_STANDARD), (19, 1024, 4, 10000.0, PD_STANDARD),
    (1, 2048, 0, 250000.0, PD_SSD), (2, 2048, 0, 10000.0, PD_SSD)]
USAGE = dict(count=1.0, cpus=2.0, mem=1024.0, gpus=0.0)
QUOTA = dict(count=10.0, cpus=70.0, mem=32768.0, gpus=10.0)
# offers: (cpus, mem, {gpu model: count}, {disk type: MiB})
STD = {PD_STANDARD: 512000.0}
MESOS = {i + 1: (c, m, {}, {}) for i, (c, m) in enumerate(
    [(10, 2048), (20, 16384), (30, 8192), (4, 2048), (4, 1024), (10, 4096), (20, 4096), (30, 16384), (100, 200000)])}
K8S = {1: (10, 2048, {}, STD), 2: (20, 16384, {}, STD), 3: (30, 8192, {}, STD), 4: (4, 2048, {}, STD),
       5: (4, 1024, {}, STD), 6: (10, 4096, {P100: 2.0}, STD), 7: (20, 4096, {P100: 4.0}, STD),
       8: (30, 16384, {P100: 1.0}, STD), 9: (100, 200000, {}, STD),
       10: (30, 2048, {}, {PD_SSD: 500000.0}), 11: (30, 2048, {}, {PD_SSD: 200000.0})}


def _match(eng, table, k8s, offer_ids, num_considerable, reverse, quota=None, usage=None, tokens=None,
           reserved_offers=(), reserved_jobs=()):
    J, O = len(JOBS), len(offer_ids)
    jb = abi.JobsSoA(n=J, user=np.zeros(J, np.int32), cpus=np.array([j[0] for j in JOBS], float),
                     mem=np.array([j[1] for j in JOBS], float), gpus=np.array([j[2] for j in JOBS], float),
                     allowed=np.ones(J, np.uint8), plugin_accept=np.ones(J, np.uint8),
                     gpu_model=np.full(J, P100, np.int32),
                     disk_request=np.array([j[3] for j in JOBS], float),
                     disk_type=np.array([j[4] for j in JOBS], np.int32),
                     # reserved jobs name their host; every other job names none (-1)
                     reserved_host=np.array([0 if j in reserved_jobs else -1 for j in range(J)], np.int32))
    offs = [table[i] for i in offer_ids]
    rank = np.arange(O, dtype=np.int32)[::-1].copy() if reverse else np.arange(O, dtype=np.int32)
    goff, gmod = abi.csr([list(o[2].keys()) for o in offs])
    _, gcnt = abi.csr([list(o[2].values()) for o in offs], np.float64)
    doff, dtyp = abi.csr([list(o[3].keys()) for o in offs])
    _, dsp = abi.csr([list(o[3].values()) for o in offs], np.float64)
    of = abi.OffersSoA(n=O, hostname_id=np.arange(O, dtype=np.int32), name_rank=rank,
                       cpus=np.array([o[0] for o in offs], float), mem=np.array([o[1] for o in offs], float),
                       run_cpus=np.zeros(O), run_mem=np.zeros(O), run_count=np.zeros(O, np.int32),
                       is_k8s=np.full(O, 1 if k8s else 0, np.uint8),
                       gpu_off=goff, gpu_model=gmod, gpu_count=gcnt,
                       disk_off=doff, disk_type=dtyp, disk_space=dsp,
                       reserved=np.array([1 if i in reserved_offers else 0 for i in offer_ids], np.uint8),
                       n_attr_cols=0)
    q, u = quota or QUOTA, usage or USAGE
    users = abi.make_users(1, quota={k: np.array([float(v)]) for k, v in q.items()},
                           usage={k: np.array([float(v)]) for k, v in u.items()},
                           tokens=np.array([tokens if tokens is not None else 1 << 20], np.int32))
    prm = traces.match_params(num_considerable, enforce_rate_limit=1 if tokens is not None else 0)
    m = eng.match(np.arange(J, dtype=np.int32), jb, of, users, prm)
    launched = {int(j) + 1 for j, a in zip(m["considerable"], m["assign"]) if a >= 0}
    return launched, int(m["stats"]["n_offers_used"])


def check_all(eng):
    n = 0
    for k8s, table in ((False, MESOS), (True, K8S)):
        for reverse in (False, True):
            def run(offer_ids, nc, **kw):
                return _match(eng, table, k8s, offer_ids, nc, reverse, **kw)
            # test-handle-resource-helpers (:1947-2133), run once per back end
            cases = [
                ("enough offers, nc 6 :1957", run([1, 2, 3], 6), ({1, 2, 3, 4}, 3)),
                ("nc 1 :1966", run([1, 2, 3], 1), ({1}, 1)),
                ("nc 2 :1975", run([1, 2, 3], 2), ({1, 2}, 2)),
                ("nc 2, one launch token :1984", run([1, 2, 3], 2, tokens=1), ({1}, 1)),
                ("nc 1, quota {5 45 16384} :2047", run([1, 2, 3], 1, quota=dict(count=5, cpus=45, mem=16384, gpus=0)),
                 ({1}, 1)),
                ("nc 1, usage {5 5 16384} :2058", run([1, 2, 3], 1, usage=dict(count=5, cpus=5, mem=16384, gpus=0)),
                 ({1}, 1)),
                ("offer for single job :2068", run([4], 10), ({1}, 1)),
                ("offer for first three jobs :2077", run([3], 10), ({1, 2, 3}, 1)),
                ("offer not fit for any job :2086", run([5], 10), (set(), 0)),
                ("too little quota :2094", run([1, 2, 3], 10, quota=dict(count=5, cpus=4, mem=4096, gpus=0)),
                 (set(), 0)),
                ("user at capacity :2103", run([1, 2, 3], 10, usage=dict(count=10, cpus=50, mem=32768, gpus=10)),
                 (set(), 0)),
                ("reserved host, nobody's reservation :2112", run([1], 10, reserved_offers=(1,)), (set(), 0)),
                ("only reserved jobs on the reserved host :2121",
                 run([9], 10, reserved_offers=(9,), reserved_jobs=(0, 1)), ({1, 2}, 1)),
            ]
            if not k8s:   # :2150-2157 in mesos, jobs requesting gpus do not get matched
                cases.append(("mesos: all offers :2151", run(list(range(1, 10)), 10)[0], {1, 2, 3, 4, 7, 8}))
            else:         # :2186-2239
                cases += [
                    ("k8s: all offers :2187", run(list(range(1, 12)), 10)[0], {1, 2, 3, 4, 5, 6, 7, 8}),
                    ("k8s: gpu offers for all gpu jobs :2195", run([6, 7], 10)[0], {5, 6}),
                    ("k8s: gpu offer for single gpu job :2203", run([6], 10)[0], {5}),
                    ("k8s: gpu offer matching no gpu job :2211", run([8], 10)[0], set()),
                    ("k8s: disk offer, same disk type :2220", run([10], 10)[0], {7}),
                    ("k8s: disk offer matching no job :2228", run([11], 7)[0], set()),
                ]
            for name, got, want in cases:
                assert got == want, ("K15", "k8s" if k8s else "mesos", "reverse" if reverse else "listed", name, got)
                n += 1
    return n
