"""Multi-GPU host logic: independent reference keyframes sharded across ranks.

The depth filter has no cross-pixel and no cross-keyframe dependency
(src/seed_check.cu, src/epipolar_match.cu, src/seed_update.cu never read a
neighbour's state), so the unit of parallelism is a keyframe: rank r owns
keyframes r, r + world, r + 2*world, ... and runs them with no data-path
collective.  The only exchange is the final gather of the depth and
convergence maps to rank 0 (NCCL on GPUs; gloo in the CPU tests).

One process per GPU, launched by torchrun; the reference has no multi-GPU path
at all.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

KEYFRAME_SEED_BASE = 0x5EED0002
KEYFRAME_SEED_STRIDE = 16


def shard_keyframes(n_keyframes: int, rank: int, world: int) -> List[int]:
    """Keyframe indices owned by `rank` (round-robin, deterministic)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_keyframes, world))


def keyframe_seed(keyframe: int) -> int:
    """Seed of the synthetic sequence of a keyframe (bench.py / tests)."""
    return KEYFRAME_SEED_BASE + KEYFRAME_SEED_STRIDE * keyframe


def gather_maps(depth: torch.Tensor, convergence: torch.Tensor, dst: int = 0
                ) -> Optional[Tuple[List[torch.Tensor], List[torch.Tensor]]]:
    """Final gather of one keyframe's (depth f32, convergence i32) maps to
    `dst`.  Returns (depths, convergences) ordered by rank on `dst`, None
    elsewhere.  Works without an initialised process group (world = 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [depth], [convergence]
    world, rank = dist.get_world_size(), dist.get_rank()
    if depth.dtype != torch.float32 or convergence.dtype != torch.int32:
        raise TypeError("gather_maps: depth must be float32 and convergence int32")
    depths = [torch.empty_like(depth) for _ in range(world)] if rank == dst else None
    convs = [torch.empty_like(convergence) for _ in range(world)] if rank == dst else None
    dist.gather(depth, depths, dst=dst)
    dist.gather(convergence, convs, dst=dst)
    return (depths, convs) if rank == dst else None


class MapGatherer:
    """The same final gather as gather_maps() for a loop: depth (f32) and convergence (i32)
    live in ONE preallocated buffer per rank, so a round is a single collective with no
    allocation -- at VGA the two-collective version cost ~1 ms per keyframe on 2 B200s,
    an eighth of the keyframe itself.

        g = MapGatherer(height, width, device)
        ... write the final maps into g.depth / g.convergence (views of g.packed) ...
        out = g.gather()      # on dst: ([depth of rank 0, ...], [convergence of rank 0, ...]); None elsewhere
    """

    def __init__(self, height: int, width: int, device=None, dst: int = 0):
        self.packed = torch.empty((2, height, width), dtype=torch.int32, device=device)
        self.depth = self.packed[0].view(torch.float32)
        self.convergence = self.packed[1]
        self.dst = dst
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.recv = None
        if self.active and dist.get_rank() == dst:
            self.recv = [torch.empty_like(self.packed) for _ in range(dist.get_world_size())]

    def gather(self) -> Optional[Tuple[List[torch.Tensor], List[torch.Tensor]]]:
        if not self.active:
            return [self.depth], [self.convergence]
        dist.gather(self.packed, self.recv, dst=self.dst)
        if self.recv is None:
            return None
        return [r[0].view(torch.float32) for r in self.recv], [r[1] for r in self.recv]


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Timing convention: a multi-GPU duration is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def assemble(keyframes_per_rank: Sequence[Sequence[int]], gathered: Sequence[torch.Tensor]) -> dict:
    """Map keyframe index -> gathered tensor, given one tensor per rank and
    the shard lists (one keyframe per rank per gather round)."""
    out = {}
    for rank, kfs in enumerate(keyframes_per_rank):
        if kfs:
            out[kfs[0]] = gathered[rank]
    return out
