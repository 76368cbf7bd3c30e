"""Host-side planning: how many chunks a line needs, how lines are packed into device batches and
how a list of lines is sharded across GPUs (lines are independent, so the shard is by
line, balanced by chunk count, with no collective on the data path)."""
from __future__ import annotations


def resized_width(h: int, w: int) -> int:
    """preprocessor.py:45-47 (Python float division, truncation, minimum chunk_width // 2)."""
    return max(50, int(48 * (w / h)))


def chunks_for(h: int, w: int, max_seq_len: int = 4096) -> int:
    """preprocessor.py:21-31 (`while start < W`, stride 84), capped where the reference truncates
    the merged sequence (predictor.py:181-183)."""
    n = (resized_width(h, w) + 83) // 84
    return min(n, (max_seq_len + 31) // 32)


def plan_batches(shapes, max_lines: int, max_chunks: int, max_seq_len: int = 4096):
    """Greedy packing of line indices, in input order, into batches within the handle's capacity."""
    batches, cur, cur_chunks = [], [], 0
    for i, (h, w) in enumerate(shapes):
        n = chunks_for(h, w, max_seq_len)
        if n > max_chunks:
            raise ValueError(f"line {i} needs {n} chunks, more than the handle's capacity {max_chunks}")
        if cur and (len(cur) >= max_lines or cur_chunks + n > max_chunks):
            batches.append(cur)
            cur, cur_chunks = [], 0
        cur.append(i)
        cur_chunks += n
    if cur:
        batches.append(cur)
    return batches


def shard_lines(shapes, world_size: int, max_seq_len: int = 4096):
    """Partition line indices over `world_size` ranks balancing the chunk count: sort by chunk count
    (descending) and deal each line to the currently lightest rank.  Returns list[list[int]]; every
    index appears exactly once; within a rank indices are kept in ascending order."""
    costs = [chunks_for(h, w, max_seq_len) for (h, w) in shapes]
    order = sorted(range(len(shapes)), key=lambda i: (-costs[i], i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]
