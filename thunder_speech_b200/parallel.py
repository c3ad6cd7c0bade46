"""Batch sharding for multi-GPU inference: one process per GPU, every utterance is independent
in eval mode (per-channel BatchNorm affine, per-utterance SqueezeExcite and feature normalisation), so the
path shards with NO data-path collective.  The only exchange is gathering the transcripts (host strings) --
``torch.distributed.all_gather_object`` over whatever backend the job runs (NCCL on the GPU box, gloo in the CPU
tests)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n`` utterances: the first ``n % world`` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_predict(predict: Callable[[torch.Tensor], List[str]], audio: torch.Tensor,
                    group=None) -> List[str]:
    """Every rank holds the same ``audio[B, N]`` (host or device); each runs ``predict`` on its contiguous slice
    and all ranks return the full list of ``B`` transcripts in the original order."""
    if not (dist.is_available() and dist.is_initialized()):
        return predict(audio)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(audio.shape[0], world, rank)
    local = predict(audio[lo:hi]) if hi > lo else []
    parts: List[Sequence[str]] = [None] * world  # type: ignore[list-item]
    dist.all_gather_object(parts, list(local), group=group)
    out: List[str] = []
    for p in parts:
        out.extend(p)
    return out


def sharded_predict_stream(module, batches, group=None, depth: int = 3, presharded: bool = False) -> List[List[str]]:
    """Strong-scaling serving loop (BASELINE configs 3 / 4: "batch-sharded at 1/2/4/8"): every rank runs
    ``module.predict_stream`` over ITS contiguous slice of each host batch ``[B, N]`` (``presharded=True``: the iterable
    already yields this rank's slice) -- H2D, graph replay, D2H and detokenisation of its own utterances, no data-path
    collective -- and the transcripts of all batches are gathered ONCE at the end of the stream (one
    ``all_gather_object``: the only exchange of the inference path).  Returns, on every rank, one list of ``B``
    transcripts per batch in the original utterance order."""
    on = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)

    def local(it):
        for xb in it:
            if presharded or world == 1:
                yield xb
            else:
                lo, hi = shard_bounds(xb.shape[0], world, rank)
                yield xb[lo:hi]

    mine = [texts for texts in module.predict_stream(local(batches), depth=depth)]
    if world == 1:
        return mine
    parts: List[Sequence[Sequence[str]]] = [None] * world  # type: ignore[list-item]
    dist.all_gather_object(parts, mine, group=group)
    out: List[List[str]] = []
    for i in range(len(mine)):
        row: List[str] = []
        for p in parts:
            row.extend(p[i])
        out.append(row)
    return out


def flat_grad_views(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    """Allocates ONE flat fp32 buffer holding every parameter's gradient and points each ``param.grad`` at its slice, so
    that data-parallel training (BASELINE config 5) averages gradients with a single in-place collective and no packing."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device, dtype=torch.float32)
    off = 0
    for p in params:
        if p.dtype != torch.float32:
            raise TypeError("flat_grad_views: fp32 master weights expected")
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def ensure_grad_views(params: Sequence[torch.nn.Parameter], flat: torch.Tensor) -> bool:
    """Re-points every ``param.grad`` at its slice of `flat` if user code replaced it (``zero_grad(set_to_none=True)``,
    ``param.grad = ...``).  Returns True when something had to be repaired."""
    off, repaired = 0, False
    for p in params:
        n = p.numel()
        g = p.grad
        if g is None or g.data_ptr() != flat.data_ptr() + off * flat.element_size() or g.shape != p.shape:
            p.grad = flat[off:off + n].view_as(p)
            repaired = True
        off += n
    return repaired


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean of `flat` over all ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests); no-op when
    torch.distributed is not initialised or world_size == 1.  This is what DDP does for the reference's training_step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    return flat
