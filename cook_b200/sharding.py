"""Multi-GPU plumbing for the path.

Pools are independent in the reference (separate Fenzo instance, queue, offers,
shares: scheduler/scheduler.clj:2488-2490, :2167-2170, :1578), so the cycle shards
by pool with no data-path collective.  The one exchange step is the per-pool
running-usage totals that feed other pools' rank filtering through quota groups
(aggregate-quota-groups, scheduler.clj:2125-2132): after a match round every rank
contributes its pools' usage deltas and one all-gather makes them consistent.
`torch.distributed` is plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def assign_pools_lpt(costs, n_gpus):
    """Longest-processing-time bin packing of pools onto GPUs.  costs[p]: pool_cycle_cost().
    Returns gpu index per pool.  Deterministic (ties -> lower pool index / gpu index)."""
    order = sorted(range(len(costs)), key=lambda p: (-costs[p], p))
    load = [0] * n_gpus
    out = [0] * len(costs)
    for p in order:
        g = min(range(n_gpus), key=lambda i: (load[i], i))
        out[p] = g
        load[g] += costs[p]
    return out


def pool_cycle_cost(n_jobs, n_offers):
    """Relative cost of one scheduling cycle of a pool, for the placement.  Measured on B200 (DESIGN
    section 7): a cycle follows the pool's NODE count, not jobs x nodes - the matcher is bound by its
    sequential placements (about as many as the nodes can hold) and unplaceable jobs are dismissed
    by the row pre-test; placements on clusters above ~6k nodes cost more each (the per-VM tables
    leave shared memory): 58.7 ms at 6.1k nodes, 195 ms at 15k."""
    return n_offers * max(1.0, (n_offers / 6000.0) ** 0.3)


def usage_delta(considerable, assign, user, cpus, mem, gpus, n_users):
    """{count,cpus,mem,gpus} per user of the jobs placed this round (the usage the
    next rank cycle must see), shape [n_users, 4], f64."""
    placed = np.asarray(considerable)[np.asarray(assign) >= 0]
    d = np.zeros((n_users, 4), np.float64)
    np.add.at(d[:, 0], user[placed], 1.0)
    np.add.at(d[:, 1], user[placed], cpus[placed])
    np.add.at(d[:, 2], user[placed], mem[placed])
    np.add.at(d[:, 3], user[placed], gpus[placed])
    return d


def aggregate_quota_groups(pool_usage, quota_groups):
    """scheduler.clj:2125-2132: pool -> usage[4] summed per quota group."""
    out = {}
    for pool, u in pool_usage.items():
        g = quota_groups.get(pool)
        if g is None:
            continue
        out[g] = out.get(g, np.zeros(4)) + np.asarray(u, np.float64)
    return out


def stack_slots(deltas, n_slots, n_users_pad):
    """Host twin of cook_exchange_usage_batch's layout: the deltas of a rank's pools ([n_users_p, 4] each)
    as ONE table [n_slots, n_users_pad, 4], slot i = pool i, spare slots and spare users zero."""
    if len(deltas) > n_slots:
        raise ValueError("more pools than slots")
    out = np.zeros((n_slots, n_users_pad, 4), np.float64)
    for i, d in enumerate(deltas):
        d = np.asarray(d, np.float64)
        out[i, :d.shape[0]] = d
    return out


def exchange_usage(local, device=None):
    """All-gather of this rank's usage table (any shape, f64).  Returns
    [world, *local.shape].  Single process: returns local[None]."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(local, np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.numpy()[None]
    if device is not None:
        t = t.to(device)
    out = torch.empty((dist.get_world_size(),) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.reshape(-1).contiguous())
    return out.cpu().numpy()
