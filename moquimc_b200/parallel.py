"""Multi-GPU host logic for one-process-per-GPU deployments (torch.distributed, NCCL over NVLink on
the GPU box, gloo in the CPU tests).  The reference is single-GPU (one cudaSetDevice per process,
moqui/base/environments/mqi_phantom_env.hpp:45); histories are independent, so the path shards
without a data-path collective:

  * Dose / EnergyDeposition / LETd: rank r transports histories [n r / N, n (r+1) / N) of the same
    seeded source (history h draws the same counter-based stream on any rank) and the dense grids are
    summed with ONE reduce per beam;
  * Dij: contiguous blocks of SPOTS per rank -- rows of the CSR are disjoint, no reduction;
  * robust scenarios: independent jobs, round-robin;
  * statistical stopping: every rank keeps the running sum / sum-of-squares grids of its own histories; per pass only the
    chunks of the grids that can hold a voxel above the criterion's dose threshold are packed and reduce-scattered,
    every rank evaluates calculate_stat (mqi_tps_env.hpp:1339-1426) on its slice, three doubles are all-reduced
    (StoppingLoop).  The dose itself is reduced once, after the last pass.

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
    accumulating).  calculate_stat (mqi_tps_env.hpp:1339-1426) averages sigma/mu over the voxels whose mean dose exceeds
    threshold x the largest mean dose, so after a pass only the voxels that CAN exceed it are exchanged:

      * a voxel that qualifies sums to more than threshold * M (M = the largest summed value), so at least one rank holds
        more than threshold * M / world of it; every rank's own largest value m_r is a lower bound of M, so a voxel whose
        local value is at most threshold * m_r / world on EVERY rank r cannot qualify.  The grids are cut into chunks of
        CHUNK voxels and a chunk is kept if any rank holds a larger value in it (one max-all-reduce of the chunk flags,
        12 800 flags for a 512 x 512 x 200 grid; no exchange is needed for the bound itself);
      * the kept chunks of sum d and sum d^2 are packed and reduce-scattered (one collective per grid: every link
        carries 1 / world of the packed values, all links at once) and every rank evaluates its slice: the largest mean
        dose (max-all-reduce of one double), then sum of sigma/mu and the voxel count (sum-all-reduce of two doubles).
        Five collectives per pass; the phase times are taken with CUDA events, not host synchronisations.

    The selected voxels and the criterion are exactly those of an evaluation on whole summed grids; at config C3 the
    packed exchange is a few per cent of the grid.  The dose itself is reduced once, after the last pass.

    transport_pass(k) must ADD this rank's share of pass k into total_sum / total_sq and return the number of histories
    it transported.  evaluate(sum_slice, sq_slice, n_histories, max_mean) returns (sum of sigma/mu over the selected
    voxels, number of selected voxels, largest mean dose of the slice); with max_mean < 0 only the third value is
    used.  On a GPU box evaluate is capi.Engine.stat_partial_buffers, the fused CUDA kernel.  total_sum / total_sq hold
    padded_len(nvox) elements (zero padding behind the grid)."""

    CHUNK = 4096

    def __init__(self, criteria_percent, transport_pass, evaluate, threshold=0.5, max_passes=1000, group=None,
                 histories_per_pass=None):
        """histories_per_pass: histories ALL ranks transport per pass, if the caller knows it (a fixed plan): saves the
        sum-all-reduce of the per-rank counts."""
        self.criteria = criteria_percent
        self.transport_pass = transport_pass
        self.evaluate = evaluate
        self.threshold = threshold
        self.max_passes = max_passes
        self.group = group
        self.histories_per_pass = histories_per_pass
        self.history = []
        self.stat_seconds = 0.0
        self.phase_seconds = {"wait_and_select": 0.0, "exchange": 0.0, "evaluate": 0.0}   # where stat_seconds goes
        self.exchanged_values = 0
        self.collectives_per_pass = 0

    @classmethod
    def padded_len(cls, n):
        """length of the stat buffers of a grid of n voxels: a whole number of chunks"""
        return (n + cls.CHUNK - 1) // cls.CHUNK * cls.CHUNK

    @staticmethod
    def _mark(t):
        """a time stamp that does not stop the host: a CUDA event on the current stream (GPU), the clock (CPU)"""
        import time
        import torch
        if t.is_cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return time.perf_counter()

    @staticmethod
    def _between(a, b):
        return a.elapsed_time(b) * 1e-3 if hasattr(a, "elapsed_time") else b - a

    def kept_chunks(self, total_sum, world):
        """indices of the chunks that can hold a voxel above the threshold (the same list on every rank).  A voxel that
        qualifies sums to more than threshold x the largest summed value M, so at least one rank holds more than
        threshold x M / world of it; every rank tests its chunks against its OWN largest value m_r <= M (no exchange
        needed for the bound) and the flags are OR-ed: a superset of the chunks that matter, the evaluation is exact."""
        import torch
        import torch.distributed as dist
        view = total_sum.view(-1, self.CHUNK)
        cmax = view.amax(dim=1)
        keep = (cmax > (self.threshold / world) * cmax.max()).to(torch.int32)
        if world > 1:
            dist.all_reduce(keep, op=dist.ReduceOp.MAX, group=self.group)
        return torch.nonzero(keep).reshape(-1), view

    def run(self, total_sum, total_sq):
        import torch
        import torch.distributed as dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        world = dist.get_world_size(self.group) if multi else 1
        assert total_sum.numel() == total_sq.numel() and total_sum.numel() % self.CHUNK == 0 and self.CHUNK % world == 0
        tracked, current, k = 0, 100.0, 0
        marks = []
        while current > self.criteria and k < self.max_passes:
            n_local = self.transport_pass(k)
            t0 = self._mark(total_sum)
            ncoll = 0
            if self.histories_per_pass is not None:
                tracked += int(self.histories_per_pass)
            elif multi:
                n = torch.tensor([n_local], dtype=torch.int64, device=total_sum.device)
                dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
                tracked += int(n.item())
                ncoll += 1
            else:
                tracked += int(n_local)
            if world == 1:      # nothing to exchange: the criterion is evaluated on the grids where they are
                t1 = t2 = t0
                s_sum, s_sq, per = total_sum, total_sq, total_sum.numel()
                n_packed = 0
            else:
                idx, view = self.kept_chunks(total_sum, world)
                t1 = self._mark(total_sum)
                # the kept chunks of each grid packed; one reduce-scatter per grid leaves every rank with the summed values of
                # 1 / world of the packed voxels
                n_packed = int(idx.numel()) * self.CHUNK
                per = n_packed // world
                s_sum = torch.empty(per, dtype=total_sum.dtype, device=total_sum.device)
                s_sq = torch.empty_like(s_sum)
                reduce_scatter_sum(s_sum, view.index_select(0, idx).reshape(-1), self.group)
                reduce_scatter_sum(s_sq, total_sq.view(-1, self.CHUNK).index_select(0, idx).reshape(-1), self.group)
                t2 = self._mark(total_sum)
                ncoll += 3
            mx = torch.tensor([self.evaluate(s_sum, s_sq, tracked, -1.0)[2] if per else 0.0], dtype=torch.float64, device=total_sum.device)
            if multi:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            s, c, _ = self.evaluate(s_sum, s_sq, tracked, float(mx.item())) if per else (0.0, 0, 0.0)
            part = torch.tensor([s, c], dtype=torch.float64, device=total_sum.device)
            if multi:
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
                ncoll += 2
            part = part.tolist()
            t3 = self._mark(total_sum)
            marks.append((t0, t1, t2, t3))
            k += 1
            current = criterion_from_partials(float(part[0]), float(part[1]))
            self.history.append(current)
            self.exchanged_values = n_packed
            self.collectives_per_pass = ncoll
        if marks and hasattr(marks[-1][3], "synchronize"):
            marks[-1][3].synchronize()
        for t0, t1, t2, t3 in marks:
            self.phase_seconds["wait_and_select"] += self._between(t0, t1)
            self.phase_seconds["exchange"] += self._between(t1, t2)
            self.phase_seconds["evaluate"] += self._between(t2, t3)
            self.stat_seconds += self._between(t0, t3)
        return tracked, current, k
