"""Multi-GPU use of the path: one process per GPU, lines sharded by `scheduling.shard_lines`, no
collective on the data path; the decoded ids are gathered once at the end.
`torch.distributed` is plumbing only (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

from .scheduling import shard_lines

TOKENS_LD = 257


def recognize_sharded(images, recognize_fn, max_seq_len: int = 4096, group=None, device=None):
    """images: the SAME list of grey uint8 arrays on every rank.  recognize_fn(list_of_images) ->
    (tokens int32 [n, 257], lengths int32 [n]) for the rank's shard.
    Returns (tokens [len(images), 257], lengths [len(images)]) in input order on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shards = shard_lines([im.shape for im in images], world, max_seq_len)
    mine = shards[rank]
    tok, ln = recognize_fn([images[i] for i in mine]) if mine else (np.zeros((0, TOKENS_LD), np.int32), np.zeros(0, np.int32))
    cap = max(len(s) for s in shards) if shards else 0
    pad_tok = np.zeros((cap, TOKENS_LD + 1), np.int32)          # last column carries the length
    pad_tok[:len(mine), :TOKENS_LD] = tok
    pad_tok[:len(mine), TOKENS_LD] = ln
    t = torch.from_numpy(pad_tok)
    if device is not None:
        t = t.to(device)
    if world == 1:
        parts = [t]
    else:
        parts = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0, group=group)
    if rank != 0:
        return None
    tokens = np.zeros((len(images), TOKENS_LD), np.int32)
    lengths = np.zeros(len(images), np.int32)
    for r, idxs in enumerate(shards):
        a = parts[r].cpu().numpy()
        for j, i in enumerate(idxs):
            tokens[i] = a[j, :TOKENS_LD]
            lengths[i] = a[j, TOKENS_LD]
    return tokens, lengths
