"""Batch-sharded data-parallel inference: one process per GPU, weights replicated, no collective inside the loop
(every op of the path is per-utterance -- SURVEY.md 8e); the only exchange is one all-gather of the finished mels.

The reference has no distributed code (README.md:32 lists multi-GPU as an unchecked to-do); this is the B200 addition.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_total, rank, world):
    """Contiguous shard [lo, hi) of `n_total` utterances for `rank`; the first n_total % world ranks get one extra."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_length(y_lengths_max_global):
    """Every shard must pad to the GLOBAL fix_len_compatibility(max y_length): unmasked norms / attention make padded
    results depend on the padded length (DEX-TTS/model/utils.py:13-17)."""
    t = int(y_lengths_max_global)
    return (t + 3) // 4 * 4


def global_max_length(y_lengths):
    """max over all ranks of the local maximum mel length (one scalar all-reduce before planning)."""
    m = y_lengths.max().to(torch.int64).reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(m.item())


def gather_mels(y_local, n_total=None):
    """All-gather (B_local, 80, T) mels of equal-size shards into (B_total, 80, T) on every rank.
    With unequal shards (n_total given) the shards are padded to the largest one and trimmed after the gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return y_local
    world = dist.get_world_size()
    if n_total is None:
        out = torch.empty((world * y_local.shape[0],) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(out, y_local.contiguous())
        return out
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((bmax,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    out = torch.empty((world * bmax,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
