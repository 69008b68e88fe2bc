"""Prompt / clip sharding over the GPUs of one box (BASELINE.json north_star item 4).

The units of the hot path are independent: prompts never interact inside `GenerateLoopV2.run`
(mimikit/loops/generate.py:207-219 — the batch dimension is independent in every op) and feature clips are
independent.  So the multi-GPU form is: one process per GPU (`torchrun`), every rank holds a weight replica and runs
its own persistent kernel over a CONTIGUOUS block of prompts, and the only collective of the whole path is ONE
gather of the generated index blocks at the end (uint8 on the wire when q_levels <= 256).  No collective sits inside
the step loop.  The functions below are the host-side logic; they are backend-agnostic (`nccl` on the GPUs, `gloo`
in the CPU tests) and never touch the kernels themselves.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_of", "gather_sequences", "generate_sharded", "extract_sharded"]


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous block when n_units are dealt to `world` ranks; the first n_units % world ranks
    hold one extra unit, ranks beyond n_units hold nothing."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_units), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_of(x, world: int, rank: int, n_units: Optional[int] = None):
    """Rank's block of a per-unit tensor (dim 0 = units).  Scalars / None / 1-element tensors are broadcast
    parameters (e.g. a scalar temperature, modules/targets.py:27-34) and pass through."""
    if x is None or not isinstance(x, torch.Tensor) or x.dim() == 0:
        return x
    if n_units is not None and x.shape[0] != n_units:
        if x.numel() == 1:
            return x
        raise ValueError(f"per-unit tensor has {x.shape[0]} rows, expected {n_units}")
    lo, hi = shard_bounds(x.shape[0], world, rank)
    return x[lo:hi]


def gather_sequences(local: torch.Tensor, n_units: int, q_levels: Optional[int] = 256, group=None) -> torch.Tensor:
    """THE collective of the path: all-gather of every rank's (b_r, T) index block -> (n_units, T) int64 in prompt
    order on every rank.  Blocks travel as uint8 when the alphabet allows (8x less NVLink traffic than int64);
    ragged blocks are padded to the largest one."""
    rank, world = _world(group)
    if world == 1:
        if local.shape[0] != n_units:
            raise ValueError("single process must hold every unit")
        return local
    if q_levels is None:
        wire = torch.int64                      # alphabet unknown: do not narrow
    else:
        wire = torch.uint8 if q_levels <= 256 else (torch.int32 if q_levels <= 2 ** 31 else torch.int64)  # gloo and NCCL have no int16
    if wire != torch.int64 and local.numel():
        lo_v, hi_v = int(local.min()), int(local.max())
        if lo_v < 0 or hi_v >= int(q_levels):
            raise ValueError(f"sequence values [{lo_v}, {hi_v}] do not fit the declared alphabet of {q_levels} levels")
    b_max = shard_bounds(n_units, world, 0)[1]
    T = local.shape[1]
    send = torch.zeros((b_max, T), dtype=wire, device=local.device)
    send[:local.shape[0]] = local.to(wire)
    recv = torch.empty((world * b_max, T), dtype=wire, device=local.device)   # concatenation form (dim 0)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, b_max, T)
    out = torch.empty((n_units, T), dtype=torch.int64, device=local.device)
    for r in range(world):
        lo, hi = shard_bounds(n_units, world, r)
        out[lo:hi] = recv[r, :hi - lo]
    return out


def generate_sharded(network, prompts: torch.Tensor, n_steps: int, temperature=None, noise=None, group=None,
                     gather: bool = True, **kw):
    """`network.generate` over this rank's block of a GLOBAL prompt batch (B, P), then the single gather.

    `temperature` may be None | float | (1,) | (B,) and `noise` None | (B, n_steps): per-prompt tensors are sharded
    with the prompts so that the result equals the single-GPU run on the whole batch, prompt for prompt.
    Returns the (B, P + n_steps) int64 sequences on every rank (or this rank's block when gather=False)."""
    rank, world = _world(group)
    B = prompts.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    if isinstance(temperature, (tuple, list)) and len(temperature) == B and B > 1:
        temperature = torch.as_tensor(temperature, dtype=torch.float32)
    T = shard_of(temperature, world, rank, B) if isinstance(temperature, torch.Tensor) else temperature
    U = noise[lo:hi] if noise is not None else None
    if hi > lo:
        local = network.generate(prompts[lo:hi], n_steps, temperature=T, noise=U, **kw)
        if isinstance(local, tuple):
            local = local[0]
    else:
        local = torch.empty((0, prompts.shape[1] + n_steps), dtype=torch.int64, device=prompts.device)
    if not gather:
        return local
    q = getattr(network, "q_levels", None)
    if q is None:
        try:
            q = int(network.config.io_spec.targets[0].out_dim)
        except AttributeError:
            q = None                             # unknown alphabet: the gather keeps int64 on the wire
    return gather_sequences(local, B, q, group)


def extract_sharded(functional, clips: torch.Tensor, group=None, gather: bool = False):
    """A feature Functional (mu-law, MagSpec, mel ...) over this rank's block of a GLOBAL (n_clips, L) batch.
    Features normally stay sharded (the consumer is sharded the same way); gather=True all-gathers equal blocks."""
    rank, world = _world(group)
    n = clips.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = functional(clips[lo:hi])
    if not gather or world == 1:
        return local
    b_max = shard_bounds(n, world, 0)[1]
    send = torch.zeros((b_max, *local.shape[1:]), dtype=local.dtype, device=local.device)
    send[:local.shape[0]] = local
    recv = torch.empty((world * b_max, *send.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, *send.shape)
    parts = []
    for r in range(world):
        l, h = shard_bounds(n, world, r)
        parts.append(recv[r, :h - l])
    return torch.cat(parts, dim=0)
