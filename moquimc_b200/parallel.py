"""Multi-GPU host logic for one-process-per-GPU deployments (torch.distributed, NCCL over NVLink on
the GPU box, gloo in the CPU tests).  The reference is single-GPU (one cudaSetDevice per process,
moqui/base/environments/mqi_phantom_env.hpp:45); histories are independent, so the path shards
without a data-path collective:

  * Dose / EnergyDeposition / LETd: rank r transports histories [n r / N, n (r+1) / N) of the same
    seeded source (history h draws the same counter-based stream on any rank) and the dense grids are
    summed with ONE reduce per beam;
  * Dij: contiguous blocks of SPOTS per rank -- rows of the CSR are disjoint, no reduction;
  * robust scenarios: independent jobs, round-robin;
  * statistical stopping: per pass, all-reduce the pass's sum / sum-of-squares grids, add them to the
    running totals every rank keeps, evaluate calculate_stat (mqi_tps_env.hpp:1339-1426) on the totals.

Nothing here computes physics; the transport itself is the CUDA library (capi.Engine).
"""
import numpy as np


def history_shard(n, rank, world):
    """(first, count) of rank's share of n histories; the shares tile [0, n) exactly."""
    a, b = n * rank // world, n * (rank + 1) // world
    return a, b - a


def spot_shard(histories_per_spot, rank, world):
    """Whole spots per rank: (first_spot, n_spots, first_history, n_histories)."""
    h = np.asarray(histories_per_spot, dtype=np.uint64)
    s0, s1 = len(h) * rank // world, len(h) * (rank + 1) // world
    cum = np.concatenate(([0], np.cumsum(h, dtype=np.uint64)))
    return s0, s1 - s0, int(cum[s0]), int(cum[s1] - cum[s0])


def scenario_shard(n_scenarios, rank, world):
    """Robust-evaluation scenarios of this rank (round-robin, replicas only: no collective)."""
    return list(range(rank, n_scenarios, world))


def robust_scenarios(shift_mm=3.0, density=0.035):
    """The 21 scenarios of config C5: {nominal, +-shift on one axis at a time} x {1, 1 -+ density}."""
    shifts = [(0.0, 0.0, 0.0)]
    for ax in range(3):
        for sgn in (+1.0, -1.0):
            s = [0.0, 0.0, 0.0]
            s[ax] = sgn * shift_mm
            shifts.append(tuple(s))
    return [{"XShift": s[0], "YShift": s[1], "ZShift": s[2], "DensityScaling": d}
            for s in shifts for d in (1.0, 1.0 - density, 1.0 + density)]


def reduce_dense(grid, dst=0, group=None):
    """One sum-reduce of a dense scorer grid (torch tensor, fp64) to rank dst."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(grid, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return grid


def criterion_from_partials(sum_ratio, count):
    """calculate_stat's final step: mean of sigma/mu over the selected voxels, in percent."""
    return 100.0 * sum_ratio / count if count > 0 else 0.0


class StoppingLoop:
    """run_by_beam_stat (mqi_tps_env.hpp:1242-1336) across ranks.

    transport_pass(k, pass_sum, pass_sq) must ADD this rank's share of pass k into the two zeroed
    grids and return the number of histories it transported; evaluate(total_sum, total_sq, n) returns
    (sum of sigma/mu over selected voxels, number of selected voxels) -- on a GPU box this is
    capi.Engine.stat_partial_buffers, the fused CUDA kernel."""

    def __init__(self, criteria_percent, transport_pass, evaluate, max_passes=1000, group=None):
        self.criteria = criteria_percent
        self.transport_pass = transport_pass
        self.evaluate = evaluate
        self.max_passes = max_passes
        self.group = group
        self.history = []

    def run(self, pass_sum, pass_sq, total_sum, total_sq):
        import torch
        import torch.distributed as dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        tracked, current, k = 0, 100.0, 0
        while current > self.criteria and k < self.max_passes:
            pass_sum.zero_()
            pass_sq.zero_()
            n = torch.tensor([self.transport_pass(k, pass_sum, pass_sq)], dtype=torch.int64, device=pass_sum.device)
            if multi:
                dist.all_reduce(pass_sum, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(pass_sq, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
            total_sum += pass_sum
            total_sq += pass_sq
            tracked += int(n.item())
            k += 1
            s, c = self.evaluate(total_sum, total_sq, tracked)
            current = criterion_from_partials(s, c)
            self.history.append(current)
        return tracked, current, k
