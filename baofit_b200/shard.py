"""Sharding of the independent units of the path (scan grid points, toy realisations) over
ranks, and the single collective that collects the chi-square surface (north_star: "a single
NCCL all-gather over NVLink collecting the chi-square surface"). One process per GPU; the
data path itself has no collective: every rank holds a full replica of tables, D, L and d.

Works with any torch.distributed backend: "nccl" on the GPU box, "gloo" in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block [first, first+count) of `total` units owned by `rank`; blocks differ in
    size by at most one and concatenate, in rank order, to the reference's iteration order."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_blocks(local, total, device=None):
    """All-gathers per-rank blocks (first axis) of unequal length into the full array, in rank
    order. `local` is a numpy array [count, ...]; returns numpy [total, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local.shape[0] == total
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    pad = max(counts)
    buf = torch.zeros(pad * width, dtype=torch.float64, device=device)
    buf[: counts[rank] * width] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).ravel()).to(buf.device)
    out = torch.empty(world * pad * width, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)  # the one collective of the path
    out = out.cpu().numpy().reshape(world, pad * width)
    parts = [out[r, : counts[r] * width].reshape((counts[r],) + local.shape[1:]) for r in range(world)]
    return np.concatenate(parts, axis=0)


_cores_shared = False


def _share_cores_once():
    """Every rank takes cores / ranks-on-this-host OpenMP threads for its per-fit host work (see host_api.share_host_cores)."""
    global _cores_shared
    if not _cores_shared and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        from baofit_b200 import host_api
        host_api.share_host_cores()
    _cores_shared = True


def scan_sharded(session, device=None):
    """CorrelationAnalyzer::parameterScan with the grid split over the ranks of the default
    process group; every rank returns the full surface (chi2[total], points[total, npar])."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    _share_cores_once()
    total = session.scan_size()
    first, count = shard_range(total, rank, world)
    part = session.parameter_scan(first, count)
    packed = np.concatenate([part["chi2"][:, None], part["points"], part["status"][:, None].astype(np.float64)], axis=1)
    full = gather_blocks(packed, total, device)
    return dict(chi2=full[:, 0], points=full[:, 1:-1], status=full[:, -1].astype(np.int32),
                best_values=part["best_values"], best_chi2=part["best_chi2"])


def toymc_sharded(session, ntoys, seed=1966, device=None):
    """Toy-MC study with realisations split over ranks. Noise is drawn from a counter-based
    generator keyed by the global toy index, so results do not depend on the number of ranks."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    _share_cores_once()
    first, count = shard_range(ntoys, rank, world)
    part = session.toymc(first, count, seed)
    packed = np.concatenate([part["chi2"][:, None], part["fitted"], part["status"][:, None].astype(np.float64)], axis=1)
    full = gather_blocks(packed, ntoys, device)
    return dict(chi2=full[:, 0], fitted=full[:, 1:-1], status=full[:, -1].astype(np.int32))
