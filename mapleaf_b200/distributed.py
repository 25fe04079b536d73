"""
Multi-GPU layer: one process per GPU (torchrun), samples sharded as contiguous ranges, one collective at the end.

The path shards naturally: lanes are independent once their inputs are drawn, so there is no
data-path collective. The only exchange is the end-of-run gather of the 7 result columns + status of every sample
(56 + 4 bytes per lane) and the combination of the per-GPU moments (count, mean, M2 per column) into the dispersion
statistics the Monte Carlo summary prints (MAPLEAF/SimulationRunners/MonteCarlo.py:148-166). The reference's
counterpart is the ray object-store hand-back of each flight (MonteCarlo.py:93-105).

`torch.distributed` is plumbing here: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from typing import List, Sequence, Tuple

import numpy as np

__all__ = ["shardRange", "combineMoments", "momentsOf", "gatherResults", "gatherMoments"]


def shardRange(nTotal: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range of `rank`: concatenating the ranks' results in rank order restores sample order."""
    return nTotal * rank // world, nTotal * (rank + 1) // world


def momentsOf(results: np.ndarray, status: np.ndarray = None):
    """(count, mean[ncol], M2[ncol]) of the rows that ended normally — host twin of mlb_moments."""
    r = np.asarray(results, dtype=np.float64)
    if status is not None:
        r = r[np.asarray(status) == 0]
    n = r.shape[0]
    if n == 0:
        return 0.0, np.zeros(r.shape[1]), np.zeros(r.shape[1])
    mean = r.mean(axis=0)
    return float(n), mean, ((r - mean) ** 2).sum(axis=0)


def combineMoments(parts: Sequence[Tuple[float, np.ndarray, np.ndarray]]):
    """Chan et al. pairwise update of (n, mean, M2), in rank order. Returns (n, mean, sample standard deviation)."""
    n, mean, M2 = 0.0, None, None
    for nb, mb, Sb in parts:
        mb, Sb = np.asarray(mb, dtype=np.float64), np.asarray(Sb, dtype=np.float64)
        if nb == 0:
            continue
        if n == 0:
            n, mean, M2 = float(nb), mb.copy(), Sb.copy()
            continue
        nt = n + nb
        d = mb - mean
        mean = mean + d * (nb / nt)
        M2 = M2 + Sb + d * d * (n * nb / nt)
        n = nt
    if n == 0:
        return 0.0, None, None
    std = np.sqrt(M2 / (n - 1)) if n > 1 else np.full_like(mean, np.nan)
    return n, mean, std


def gatherResults(localResults, localStatus, group=None):
    """All-gather of per-rank result blocks ([n_r, 7] float64, [n_r] int32 torch tensors on the collective's device) in rank order.
    Ranks may hold different numbers of samples (shardRange with nTotal % world != 0): the blocks are padded to the largest for the
    collective (all_gather_into_tensor needs equal sizes) and trimmed afterwards. Returns ([sum n_r, 7], [sum n_r]) in sample order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = int(localResults.shape[0])
    counts = torch.zeros(world, dtype=torch.int64, device=localResults.device)
    dist.all_gather_into_tensor(counts, torch.tensor([n], dtype=torch.int64, device=localResults.device), group=group)
    counts = [int(c) for c in counts.tolist()]
    nMax, ncol = max(counts), localResults.shape[-1]
    sendR = torch.zeros((nMax, ncol), dtype=localResults.dtype, device=localResults.device)
    sendS = torch.full((nMax,), -1, dtype=localStatus.dtype, device=localStatus.device)
    sendR[:n], sendS[:n] = localResults, localStatus
    res = torch.empty((world, nMax, ncol), dtype=localResults.dtype, device=localResults.device)
    st = torch.empty((world, nMax), dtype=localStatus.dtype, device=localStatus.device)
    dist.all_gather_into_tensor(res.view(-1), sendR.view(-1), group=group)
    dist.all_gather_into_tensor(st.view(-1), sendS.view(-1), group=group)
    if all(c == nMax for c in counts):
        return res.view(-1, ncol), st.view(-1)
    return torch.cat([res[r, :c] for r, c in enumerate(counts)]), torch.cat([st[r, :c] for r, c in enumerate(counts)])


def gatherMoments(count: float, mean: np.ndarray, M2: np.ndarray, device=None, group=None):
    """All-gather of the per-rank (count, mean, M2) and their combination. Returns (n, mean, std) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    k = len(mean)
    mine = torch.from_numpy(np.concatenate([[count], np.asarray(mean, dtype=np.float64), np.asarray(M2, dtype=np.float64)]))
    if device is not None:
        mine = mine.to(device)
    allm = torch.empty((world, 1 + 2 * k), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(allm.view(-1), mine, group=group)
    a = allm.cpu().numpy()
    return combineMoments([(row[0], row[1:1 + k], row[1 + k:]) for row in a])
