"""Constraint truth tables transcribed BY HAND from the reference's own tests
(SURVEY §8c K11 + K16).  Every case is one job against one roomy offer; the
expected value is whether the job may be placed on that host.

  K16a  test/cook/test/scheduler/constraints.clj:43-57    user-defined EQUALS (6 cases)
  K16b  constraints.clj:59-221                           gpu-host (11 cases)
  K16c  constraints.clj:223-301                          disk-host (4 cases + "no constraint")
  K16d  constraints.clj:303-350                          rebalancer reservation (2 cases)
  K16e  constraints.clj:433-439                          estimated completion (3 cases)
  K11   test/cook/test/scheduler/scheduler.clj:586-658   checkpoint locality (6 cases)

The dictionary encoding (value ids, model ids, location ids) is the host shim's job
(INTEGRATION.md); here it is done by hand next to each case.
"""
import numpy as np

from cook_b200 import abi, traces


def _run(eng, job, offer, host_lifetime_mins=0, groups=None):
    """job / offer: dicts of column overrides for a 1 x 1 match; returns placed?"""
    jkw = dict(n=1, user=np.zeros(1, np.int32), cpus=np.array([job.get("cpus", 5.0)]),
               mem=np.array([job.get("mem", 5.0)]), gpus=np.array([job.get("gpus", 0.0)]),
               allowed=np.ones(1, np.uint8), plugin_accept=np.ones(1, np.uint8))
    if "gpu_model" in job:
        jkw["gpu_model"] = np.array([job["gpu_model"]], np.int32)
    if "attrs" in job:  # list of (col, value id)
        off, col = abi.csr([[c for c, _ in job["attrs"]]])
        _, val = abi.csr([[v for _, v in job["attrs"]]])
        jkw.update(attr_off=off, attr_col=col, attr_val=val)
    if "disk_request" in job:
        jkw["disk_request"] = np.array([job["disk_request"]], np.float64)
        jkw["disk_type"] = np.array([job.get("disk_type", 0)], np.int32)
    if "est_end_ms" in job:
        jkw["est_end_ms"] = np.array([job["est_end_ms"]], np.int64)
    if "ckpt_location" in job:
        jkw["ckpt_location"] = np.array([job["ckpt_location"]], np.int32)
    if "reserved_host" in job:
        jkw["reserved_host"] = np.array([job["reserved_host"]], np.int32)
    jobs = abi.JobsSoA(**jkw)
    ncol = offer.get("n_attr_cols", 0)
    okw = dict(n=1, hostname_id=np.array([offer.get("hostname_id", 0)], np.int32),
               name_rank=np.zeros(1, np.int32), cpus=np.array([offer.get("cpus", 40.0)]),
               mem=np.array([offer.get("mem", 5000.0)]), run_cpus=np.zeros(1), run_mem=np.zeros(1),
               run_count=np.array([offer.get("run_count", 0)], np.int32), n_attr_cols=ncol)
    if ncol:
        okw["attr"] = np.array(offer["attr"], np.int32).reshape(-1)
    if "is_k8s" in offer:
        okw["is_k8s"] = np.array([offer["is_k8s"]], np.uint8)
    if "gpus" in offer:  # {model id: count}
        off, mod = abi.csr([list(offer["gpus"].keys())])
        _, cnt = abi.csr([list(offer["gpus"].values())], np.float64)
        okw.update(gpu_off=off, gpu_model=mod, gpu_count=cnt)
    if "disk" in offer:  # {type id: MiB}
        off, typ = abi.csr([list(offer["disk"].keys())])
        _, sp = abi.csr([list(offer["disk"].values())], np.float64)
        okw.update(disk_off=off, disk_type=typ, disk_space=sp)
    if "host_start_time" in offer:
        okw["host_start_time"] = np.array([offer["host_start_time"]], np.int64)
    if "location" in offer:
        okw["location"] = np.array([offer["location"]], np.int32)
    if "reserved" in offer:
        okw["reserved"] = np.array([offer["reserved"]], np.uint8)
    offers = abi.OffersSoA(**okw)
    users = abi.make_users(1)
    prm = traces.match_params(1, host_lifetime_mins=host_lifetime_mins)
    m = eng.match(np.zeros(1, np.int32), jobs, offers, users, prm, groups=groups)
    return bool(m["assign"][0] >= 0)


def check_all(eng):
    # ---- K16a user-defined constraint: is_spot EQUALS true AND instance_type EQUALS mem.large
    # columns: 0 = is_spot {true: 1, false: 2}, 1 = instance_type {mem.large: 1, cpu.large: 2}; 0 = absent
    want = [(0, 1), (1, 1)]
    for attr, expect in [([1, 1], True), ([1, 2], False), ([2, 1], False), ([1, 0], False),
                         ([0, 1], False), ([0, 0], False)]:
        got = _run(eng, dict(attrs=want), dict(n_attr_cols=2, attr=attr))
        assert got == expect, ("K16a", attr, got)

    # ---- K16b gpu-host constraint; models: p100 = 0, k80 = 1
    k8s_gpu = dict(is_k8s=1, gpus={0: 4.0})
    k8s_plain = dict(is_k8s=1, gpus={})
    mesos = dict(is_k8s=0)
    cases = [
        (dict(gpus=1.0, gpu_model=0), k8s_gpu, False),   # too many GPUs on the host (count must be equal)
        (dict(gpus=8.0, gpu_model=0), k8s_gpu, False),   # too few
        (dict(gpus=4.0, gpu_model=1), k8s_gpu, False),   # wrong model
        (dict(gpus=4.0, gpu_model=0), k8s_gpu, True),    # exact
        (dict(gpus=4.0, gpu_model=0), dict(k8s_gpu, run_count=1), False),  # a task already on the VM
        (dict(gpus=0.0), k8s_gpu, False),                # non-GPU job on a GPU host
        (dict(gpus=1.0, gpu_model=0), k8s_plain, False), # GPU job on a non-GPU host
        (dict(gpus=0.0), k8s_plain, True),
        (dict(gpus=1.0, gpu_model=-1), mesos, False),    # GPU job on a Mesos host
        (dict(gpus=0.0), mesos, True),
        (dict(gpus=0.0), k8s_plain, True),
    ]
    for i, (job, offer, expect) in enumerate(cases):
        got = _run(eng, job, offer)
        assert got == expect, ("K16b", i, got)

    # ---- K16c disk-host constraint; host disk {"pd-standard": 50}; types: pd-standard = 0, pd-ssd = 1
    # (config type-map standard -> pd-standard, default type standard: resolved by the host shim)
    disk_host = dict(is_k8s=1, gpus={}, disk={0: 50.0})
    for i, (job, expect) in enumerate([
            (dict(disk_request=10.0, disk_type=0), True),
            (dict(disk_request=50.0, disk_type=0), True),    # no type => default type
            (dict(disk_request=100.0, disk_type=0), False),
            (dict(disk_request=10.0, disk_type=1), False),   # host has no pd-ssd
            (dict(disk_request=-1.0), True)]):               # pool without the constraint
        got = _run(eng, job, disk_host)
        assert got == expect, ("K16c", i, got)

    # ---- K16d rebalancer reservation: hostB (id 1) is reserved for ANOTHER job
    assert _run(eng, dict(), dict(hostname_id=1, reserved=1)) is False
    assert _run(eng, dict(), dict(hostname_id=0, reserved=0)) is True
    # (and the job the host is reserved for may use it: scheduler.clj:645-653)
    assert _run(eng, dict(reserved_host=1), dict(hostname_id=1, reserved=1)) is True

    # ---- K16e estimated completion: end 100000 ms, host lifetime 1 min
    assert _run(eng, dict(est_end_ms=100000), dict(), host_lifetime_mins=1) is True            # no start time
    assert _run(eng, dict(est_end_ms=100000), dict(host_start_time=0), host_lifetime_mins=1) is False
    assert _run(eng, dict(est_end_ms=100000), dict(host_start_time=51), host_lifetime_mins=1) is True

    # ---- K11 checkpoint locality: locations a = 0, b = 1; cluster-1 is at a, cluster-2 at b;
    # the job's last instance ran on (cluster) => ckpt_location = that cluster's location when
    # :job/checkpoint is set, else no constraint.  Offer: 1 cpu / 1000 MB, job 1 cpu / 1000 MB.
    off = lambda loc: dict(cpus=1.0, mem=1000.0, location=loc)
    job = lambda loc: dict(cpus=1.0, mem=1000.0, ckpt_location=loc)
    for i, (offer_loc, ckpt, expect) in enumerate([(0, 0, True), (0, -1, True), (1, 0, False),
                                                    (1, -1, True), (0, 1, False), (1, 1, True)]):
        got = _run(eng, job(ckpt), off(offer_loc))
        assert got == expect, ("K11", i, got)
