"""Multi-GPU plumbing of the component shard (one process per GPU, torch.distributed).

Sibling components share no unassigned variable and no factor (src/Component.cpp:508-549), so a
wave of sibling solves partitions across ranks with no data-path collective.  The only exchange
steps are (a) the all-reduce of the per-rank partial objective and (b), when the caller needs the
assembled state, an all-gather of the disjoint solution slices.  Works with the nccl backend on
device tensors and with gloo on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_problems(ps, rank, world):
    """Indices of the components rank `rank` owns: greedy longest-processing-time packing by
    |factors|*|vars| (deterministic; every rank computes the same assignment)."""
    nv = np.diff(ps.var_off)
    nf = np.diff(ps.fac_off)
    cost = (nf * np.maximum(nv, 1)).astype(np.int64)
    order = np.argsort(-cost, kind="stable")
    owner = np.empty(ps.n, dtype=np.int64)
    if ps.n > 4 * world and cost[order[0]] * ps.n < 50 * max(int(cost.sum()), 1):
        # many comparable components: dealing the cost-sorted list round-robin is within one
        # component of the LPT bound and O(n)
        owner[order] = np.arange(ps.n) % world
    else:
        load = np.zeros(world, dtype=np.int64)
        for i in order:
            r = int(np.argmin(load))
            owner[i] = r
            load[r] += cost[i]
    return np.nonzero(owner == rank)[0]


def allreduce_objective(partial):
    """Sum of the per-rank partial objectives (fp64 tensor of one element, in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial


def gather_solution(x_full, vids, x_slice):
    """Assemble the state after a sharded wave: every rank contributes (vids, x_slice) for the
    variables of its own components; returns x_full with all ranks' slices written (same on every
    rank).  Slices are disjoint by construction, so the write order does not matter."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        x_full[vids] = x_slice
        return x_full
    world = dist.get_world_size()
    n = torch.tensor([vids.numel()], dtype=torch.int64, device=vids.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    nmax = int(max(int(s.item()) for s in sizes))
    vpad = torch.zeros(nmax, dtype=torch.int64, device=vids.device)
    xpad = torch.zeros(nmax, dtype=torch.float64, device=x_slice.device)
    vpad[: vids.numel()] = vids
    xpad[: x_slice.numel()] = x_slice
    vall = [torch.empty_like(vpad) for _ in range(world)]
    xall = [torch.empty_like(xpad) for _ in range(world)]
    dist.all_gather(vall, vpad)
    dist.all_gather(xall, xpad)
    for r in range(world):
        k = int(sizes[r].item())
        x_full[vall[r][:k]] = xall[r][:k]
    return x_full
