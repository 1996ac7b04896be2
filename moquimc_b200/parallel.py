"""Multi-GPU host logic for one-process-per-GPU deployments (torch.distributed, NCCL over NVLink on
the GPU box, gloo in the CPU tests).  The reference is single-GPU (one cudaSetDevice per process,
moqui/base/environments/mqi_phantom_env.hpp:45); histories are independent, so the path shards
without a data-path collective:

  * Dose / EnergyDeposition / LETd: rank r transports histories [n r / N, n (r+1) / N) of the same
    seeded source (history h draws the same counter-based stream on any rank) and the dense grids are
    summed with ONE reduce per beam;
  * Dij: contiguous blocks of SPOTS per rank -- rows of the CSR are disjoint, no reduction;
  * robust scenarios: independent jobs, round-robin;
  * statistical stopping: every rank keeps the running sum / sum-of-squares grids of its own histories; per pass the
    two grids are reduce-scattered, every rank evaluates calculate_stat (mqi_tps_env.hpp:1339-1426) on its slice of
    the summed grids, three doubles are all-reduced.  The dose itself is reduced once, after the last pass.

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


def spot_shard_blocks(histories_per_spot, rank, world, blocks_per_rank=4):
    """Whole spots per rank in `blocks_per_rank` contiguous blocks dealt round-robin: [(first_history, n_histories), ...].
    A plan lists its spots energy layer by energy layer and the cost of a spot grows with its range, so one contiguous
    block per rank (spot_shard) leaves the rank with the highest layers working about twice as long as the one with the
    lowest; the blocks of all ranks still tile the spot list exactly, and rows stay on one rank (no reduction)."""
    h = np.asarray(histories_per_spot, dtype=np.uint64)
    cum = np.concatenate(([0], np.cumsum(h, dtype=np.uint64)))
    nb = world * blocks_per_rank if world > 1 else 1
    out = []
    for b in range(rank, nb, world):
        s0, s1 = len(h) * b // nb, len(h) * (b + 1) // nb
        if cum[s1] > cum[s0]:
            out.append((int(cum[s0]), int(cum[s1] - cum[s0])))
    return out


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


def slice_len(n, world):
    """Elements per rank when n values are dealt to `world` ranks in equal slices (the last ones padded)."""
    return (n + world - 1) // world


def reduce_scatter_sum(out_slice, full, group=None):
    """out_slice <- this rank's slice of the element-wise sum of `full` over the ranks.  `full` has
    world * len(out_slice) elements.  NCCL: one ncclReduceScatter (every link carries 1/world of the grid);
    gloo has no reduce-scatter, so the CPU tests all-reduce a copy and cut the slice out of it."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = out_slice.numel()
    assert full.numel() == world * n
    if world == 1:
        out_slice.copy_(full)
    elif dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out_slice, full, op=dist.ReduceOp.SUM, group=group)
    else:
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out_slice.copy_(tmp[rank * n:(rank + 1) * n])
    return out_slice


class StoppingLoop:
    """run_by_beam_stat (mqi_tps_env.hpp:1242-1336) across ranks, without moving whole grids per pass.

    Every rank keeps its OWN running sums (sum d, sum d^2 of the histories it transported: the scorers simply keep
    accumulating).  After a pass the two grids are reduce-scattered, each rank evaluates calculate_stat
    (mqi_tps_env.hpp:1339-1426) on its slice of the summed grids, and three numbers are all-reduced: the largest mean
    dose (max), then the sum of sigma/mu and the number of voxels above the threshold (sum).

    transport_pass(k) must ADD this rank's share of pass k into total_sum / total_sq and return the number of
    histories it transported.  evaluate(sum_slice, sq_slice, n_histories, max_mean) returns (sum of sigma/mu over the
    selected voxels, number of selected voxels, largest mean dose of the slice); with max_mean < 0 only the third
    value is used.  On a GPU box evaluate is capi.Engine.stat_partial_buffers, the fused CUDA kernel.
    total_sum / total_sq hold world * slice_len(nvox, world) elements (zero padding behind the grid).  Only the
    range of the grids that some rank has scored into is exchanged (scored_range)."""

    def __init__(self, criteria_percent, transport_pass, evaluate, max_passes=1000, group=None):
        self.criteria = criteria_percent
        self.transport_pass = transport_pass
        self.evaluate = evaluate
        self.max_passes = max_passes
        self.group = group
        self.history = []
        self.stat_seconds = 0.0
        self.phase_seconds = {"range": 0.0, "exchange": 0.0, "evaluate": 0.0}   # where stat_seconds goes

    CHUNK = 4096

    @staticmethod
    def _tick(t):
        import time
        import torch
        if t.is_cuda:
            torch.cuda.synchronize()
        return time.perf_counter()

    def scored_range(self, total_sum, world):
        """[lo, lo + span) with span a multiple of `world`: the part of the grid some rank has scored into, found in
        chunks of CHUNK values (the stat grids are zero outside the beam, so only this range is exchanged)."""
        import torch
        import torch.distributed as dist
        n = total_sum.numel()
        m = n // self.CHUNK
        lo, hi = n, 0
        if m:
            nz = torch.nonzero(total_sum[:m * self.CHUNK].view(m, self.CHUNK).amax(dim=1) > 0)
            if nz.numel():
                lo, hi = int(nz[0].item()) * self.CHUNK, (int(nz[-1].item()) + 1) * self.CHUNK
        if m * self.CHUNK < n and bool((total_sum[m * self.CHUNK:].amax() > 0).item()):
            lo, hi = min(lo, m * self.CHUNK), n
        if world > 1:
            r = torch.tensor([-lo, hi], dtype=torch.int64, device=total_sum.device)
            dist.all_reduce(r, op=dist.ReduceOp.MAX, group=self.group)
            lo, hi = -int(r[0].item()), int(r[1].item())
        if hi <= lo:
            lo, hi = 0, n
        span = ((hi - lo + world - 1) // world) * world
        lo = max(0, min(lo, n - span))
        return lo, min(span, n - lo)

    def run(self, total_sum, total_sq):
        import time
        import torch
        import torch.distributed as dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        world = dist.get_world_size(self.group) if multi else 1
        n_slice = total_sum.numel() // world
        assert n_slice * world == total_sum.numel() == total_sq.numel()
        s_sum = torch.zeros(n_slice, dtype=total_sum.dtype, device=total_sum.device)
        s_sq = torch.zeros_like(s_sum)
        tracked, current, k = 0, 100.0, 0
        while current > self.criteria and k < self.max_passes:
            n = torch.tensor([self.transport_pass(k)], dtype=torch.int64, device=total_sum.device)
            if total_sum.is_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            if multi:
                dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
            tracked += int(n.item())
            lo, span = self.scored_range(total_sum, world)
            t1 = self._tick(total_sum)
            per = span // world
            reduce_scatter_sum(s_sum[:per], total_sum[lo:lo + span], self.group)
            reduce_scatter_sum(s_sq[:per], total_sq[lo:lo + span], self.group)
            t2 = self._tick(total_sum)
            mx = torch.tensor([self.evaluate(s_sum[:per], s_sq[:per], tracked, -1.0)[2]], dtype=torch.float64, device=total_sum.device)
            if multi:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            s, c, _ = self.evaluate(s_sum[:per], s_sq[:per], tracked, float(mx.item()))
            part = torch.tensor([s, c], dtype=torch.float64, device=total_sum.device)
            if multi:
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
            t3 = self._tick(total_sum)
            self.phase_seconds["range"] += t1 - t0
            self.phase_seconds["exchange"] += t2 - t1
            self.phase_seconds["evaluate"] += t3 - t2
            k += 1
            current = criterion_from_partials(float(part[0].item()), float(part[1].item()))
            self.history.append(current)
            self.exchanged_values = span
            if total_sum.is_cuda:
                torch.cuda.synchronize()
            self.stat_seconds += time.perf_counter() - t0
        return tracked, current, k
