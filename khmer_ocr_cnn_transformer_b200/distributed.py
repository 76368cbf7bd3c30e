"""Multi-GPU use of the path: one process per GPU, lines sharded by `scheduling.shard_lines`, no
collective on the data path; the decoded ids are gathered once per request.
`torch.distributed` is plumbing only (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

from .scheduling import shard_lines

TOKENS_LD = 257


_SIDE_STREAMS: dict = {}


def _side_stream(device):
    import torch
    key = str(torch.device(device))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def gather_ids(tok_local, len_local, shards, rank: int, world: int, group=None, device=None):
    """The one collective of a request: every rank contributes the ids of ITS shard (tok_local int32 [len(shards[rank]), 257],
    len_local int32 [...]), padded to the largest shard, in one `gather` to rank 0, which puts them back into input order.
    Returns (tokens [n_lines, 257], lengths [n_lines]) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    n_local = len(shards[rank])
    cap = max(len(s) for s in shards) if shards else 0
    send = np.zeros((cap, TOKENS_LD + 1), np.int32)             # last column carries the length
    send[:n_local, :TOKENS_LD] = tok_local
    send[:n_local, TOKENS_LD] = len_local
    if world == 1:
        arrs = [send]
    else:
        t = torch.from_numpy(send)
        if device is not None and torch.device(device).type == "cuda":
            # on a side stream: the default (legacy) stream must stay out of it while other host threads of the process
            # may be capturing CUDA graphs (libkocr's decode loop)
            side = _side_stream(device)
            with torch.cuda.stream(side):
                t = t.to(device, non_blocking=True)
                parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
                dist.gather(t, parts, dst=0, group=group)
                arrs = [p.cpu().numpy() for p in parts] if rank == 0 else None
                side.synchronize()
        else:
            if device is not None:
                t = t.to(device)
            parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, parts, dst=0, group=group)
            arrs = [p.cpu().numpy() for p in parts] if rank == 0 else None
    if rank != 0:
        return None
    n = sum(len(s) for s in shards)
    tokens = np.zeros((n, TOKENS_LD), np.int32)
    lengths = np.zeros(n, np.int32)
    for r, idxs in enumerate(shards):
        a = arrs[r][:len(idxs)]
        tokens[idxs] = a[:, :TOKENS_LD]
        lengths[idxs] = a[:, TOKENS_LD]
    return tokens, lengths


def recognize_sharded(images, recognize_fn, max_seq_len: int = 4096, group=None, device=None):
    """images: the SAME list of grey uint8 arrays on every rank.  recognize_fn(list_of_images) ->
    (tokens int32 [n, 257], lengths int32 [n]) for the rank's shard (e.g. `LinePipeline.recognize`).
    Returns (tokens [len(images), 257], lengths [len(images)]) in input order on rank 0, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shards = shard_lines([im.shape for im in images], world, max_seq_len)
    mine = shards[rank]
    tok, ln = recognize_fn([images[i] for i in mine]) if mine else (np.zeros((0, TOKENS_LD), np.int32), np.zeros(0, np.int32))
    return gather_ids(tok, ln, shards, rank, world, group=group, device=device)
