"""Data-parallel generation over the GPUs of one box.

Every request (image + prompt) is independent through the whole path, so the global batch is split into
contiguous slices, one per rank; weights are replicated; each rank owns the paged KV cache of its slice.  The only
exchange is ONE all-gather of the sampled token ids per decode step (NCCL over NVLink/NVSwitch, enqueued on the
compute stream right after the lm_head+argmax kernels), so every rank -- and the caller on rank 0 -- holds all
tokens.  With the `gloo` backend (CPU tests) the same code path runs with host tensors.
"""
from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (first `rem` ranks get one extra)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_step_tokens(local_tok: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gather the per-rank token vectors of one decode step into the global (B,) vector."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_tok
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    send = local_tok
    if send.shape[0] != width:      # uneven split: pad to the widest shard so one fixed-size all-gather suffices
        send = torch.zeros(width, dtype=local_tok.dtype, device=local_tok.device)
        send[: local_tok.shape[0]] = local_tok
    recv = torch.empty(world * width, dtype=local_tok.dtype, device=local_tok.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(hi - lo == width for lo, hi in sizes):
        return recv
    return torch.cat([recv[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def _native_dp_ready(eng, group) -> bool:
    """True when the per-step exchange can run inside the engine (NCCL communicator owned by the native context, all-gather
    captured in the decode CUDA graphs): a real Engine on a CUDA device under an NCCL process group."""
    if not hasattr(eng, "dp_init") or not dist.is_initialized():
        return False
    if eng.device.type != "cuda" or dist.get_backend(group) != "nccl":
        return False
    return True


@torch.no_grad()
def generate_dp(model, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_new_tokens: int, group=None,
                step_hook: Optional[Callable] = None, phase_hook: Optional[Callable] = None) -> torch.Tensor:
    """Greedy DP generation.  `input_ids` (B,T) / `pixel_values` (B,3,I,I) hold the GLOBAL batch on every rank;
    each rank computes only its slice.  Returns the (B, max_new_tokens) int64 tokens of the whole batch on every rank.
    `model` is a visualcla.VisualCLAModel (or any object with `._engine` and `._image_layout`).

    world == 1, or NCCL on CUDA: the loop is pure CUDA-graph replays (16 steps per graph); at world > 1 each captured step
    contains the NCCL all-gather of the chosen tokens on a forked branch (vcla_nccl_init), so there is no Python and no torch
    collective per token.  Otherwise (gloo on CPU, or a step_hook): one Python iteration + one torch all-gather per step."""
    from . import _native as N
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = input_ids.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    eng = model._engine
    dev = eng.device
    nloc = hi - lo
    # slice first, then move: a rank only ever copies its own requests host -> device
    ids = input_ids[lo:hi].to(dev, non_blocking=True)
    px = None if pixel_values is None else pixel_values[lo:hi].to(dev, non_blocking=True)
    if phase_hook is not None:
        phase_hook("start")
    native = world > 1 and step_hook is None and _native_dp_ready(eng, group)
    if native:
        eng.dp_init(group)
        if nloc > eng.dp_width:
            raise ValueError(f"shard of {nloc} requests exceeds the {eng.dp_width} exchange slots of this rank's engine (max_batch)")
        eng.dp_set_active(True)
    try:
        return _generate_dp_body(model, eng, N, input_ids, ids, px, max_new_tokens, group, step_hook, phase_hook, native, world, B, lo, hi)
    finally:
        if native:
            eng.dp_set_active(False)


def _generate_dp_body(model, eng, N, input_ids, ids, px, max_new_tokens, group, step_hook, phase_hook, native, world, B, lo, hi):
    dev = eng.device
    nloc = hi - lo
    nq = getattr(eng, "nq", 0)
    max_seq = getattr(eng, "max_seq", None)
    out = torch.empty(B, max_new_tokens, dtype=torch.int64, device=dev)
    # persistent per-batch token buffer: its address keys the captured decode graphs (a fresh tensor per call would capture
    # a new graph whenever the caching allocator returns another address)
    tok = eng.token_buffer(max(nloc, 1)) if hasattr(eng, "token_buffer") else torch.zeros(max(nloc, 1), dtype=torch.int32, device=dev)
    if nloc > 0:
        mode, rows = model._image_layout(ids, px)
        S = ids.shape[1] + (nq if mode == N.IMAGE_AT_HEAD else 0)
        if max_seq is not None and S + max_new_tokens > max_seq:
            raise ValueError(f"prompt ({S}) + max_new_tokens ({max_new_tokens}) exceeds the context capacity max_seq={max_seq}")
        if mode != N.TEXT_ONLY:
            eng.vision_encode(px)
        _, first, _ = eng.prefill(ids, mode, rows, all_logits=False, last_logits=False)
        tok[:nloc].copy_(first)
    elif native:
        eng.reset()                     # rewinds this rank's exchange history, as a prefill would
        eng.dp_idle_exchange()          # a rank without requests still takes part in the prefill-step exchange
    if phase_hook is not None:
        phase_hook("prefill_done")
    if world == 1 and step_hook is None:
        # single GPU: the graph appends every chosen token to a device-side history -> the loop is pure graph replays
        eng.decode_many(tok[:nloc], max_new_tokens - 1)
        out.copy_(eng.read_history(nloc, max_new_tokens).t())
    elif native:
        if nloc > 0:
            eng.decode_many(tok[:nloc], max_new_tokens - 1)
        else:
            for _ in range(max_new_tokens - 1):
                eng.dp_idle_exchange()
        hist = eng.read_history_dp(max_new_tokens)                     # (n_new, world * width)
        w = eng.dp_width
        for r in range(world):
            rlo, rhi = shard_bounds(B, world, r)
            out[rlo:rhi] = hist[:, r * w: r * w + (rhi - rlo)].t()
    else:
        for step in range(max_new_tokens):
            if step > 0 and nloc > 0:
                eng.decode_step(tok[:nloc], tok[:nloc], None)
            out[:, step] = gather_step_tokens(tok[:nloc], B, group)
            if step_hook is not None:
                step_hook(step)
    if phase_hook is not None:
        phase_hook("done")
    return out
