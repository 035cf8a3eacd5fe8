"""Frame-wise sharding of a sequence over ranks / GPUs: frames and GOFs are independent
(reference src/lib.rs:114-120, src/decoder.rs:186), so rank r simply reconstructs a contiguous slice of the frames.
There is no exchange step and therefore no data-path collective; ``torch.distributed`` is used only to agree on the
timing (max over ranks) and to sum the point counts."""
from __future__ import annotations

from typing import Tuple


def frames_for_rank(total_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the frames rank ``rank`` owns; sizes differ by at most one, order is preserved."""
    lo = total_frames * rank // world_size
    hi = total_frames * (rank + 1) // world_size
    return lo, hi


def reduce_metrics(elapsed_ms: float, points: int, frames: int, device=None):
    """max(elapsed) / sum(points) / sum(frames) over all ranks (no-op without an initialised process group)."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return elapsed_ms, points, frames
    if not (dist.is_available() and dist.is_initialized()):
        return elapsed_ms, points, frames
    dev = device if device is not None else "cpu"
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    c = torch.tensor([points, frames], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), int(c[0].item()), int(c[1].item())
