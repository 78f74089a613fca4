"""Multi-GPU plumbing: weight replicas, batch sharding, ONE collective.

The reference's only multi-GPU inference strategy is a full pipeline replica per GPU fed from a work queue
(scripts/run_eval.py:143-198,221-247); samples never interact inside the transformer.  Here: one process per GPU
(torchrun), sample b runs on rank b mod G, and the step-invariant conditioning (T5 prompt embeddings, CLIP pooled
embedding, guidance, sigma schedule) is sent from rank 0 in a single NCCL broadcast over NVLink before the loop.
No per-step collective.  Everything below is backend-agnostic torch.distributed (gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(global_batch: int, rank: int, world: int) -> List[int]:
    """Samples of the global batch this rank denoises (round-robin, like draining run_eval.py's queue in order)."""
    return list(range(rank, global_batch, world))


def _pack(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.contiguous().view(torch.uint8).reshape(-1) for t in tensors])


def broadcast_conditioning(prompt_embeds: torch.Tensor, pooled: torch.Tensor, sigmas: torch.Tensor,
                           guidance_scale: float, src: int = 0, group=None) -> Tuple[torch.Tensor, torch.Tensor,
                                                                                     torch.Tensor, float]:
    """One broadcast of everything the denoising loop needs besides the per-sample latents.

    Every rank passes tensors of the right shape/dtype/device (contents only matter on `src`); returns the
    src rank's values.  The payload is one flat byte buffer, so it is exactly one collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return prompt_embeds, pooled, sigmas, float(guidance_scale)
    g = torch.tensor([guidance_scale], dtype=torch.float32, device=prompt_embeds.device)
    parts = [prompt_embeds, pooled, sigmas.to(prompt_embeds.device), g]
    flat = _pack(parts)
    dist.broadcast(flat, src=src, group=group)
    out, off = [], 0
    for t in parts:
        n = t.numel() * t.element_size()
        out.append(flat[off:off + n].view(t.dtype).reshape(t.shape).clone())
        off += n
    return out[0], out[1], out[2], float(out[3].item())


def gather_latents(latents: torch.Tensor, global_batch: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Collect the per-rank results [b_local, S, C] back into global batch order on `dst` (after the loop)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return latents
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (global_batch + world - 1) // world
    pad = torch.zeros((per,) + tuple(latents.shape[1:]), dtype=latents.dtype, device=latents.device)
    pad[: latents.shape[0]] = latents
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty((global_batch,) + tuple(latents.shape[1:]), dtype=latents.dtype, device=latents.device)
    for r in range(world):
        idx = shard_indices(global_batch, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out


def max_over_ranks(value: float, device) -> float:
    """Device-timed numbers are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
