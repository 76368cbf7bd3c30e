"""world_size-2 gloo test (CPU) of the N>1 plumbing: shard by line, recognise the shard, gather the ids
to rank 0, restore input order.  The recogniser is replaced by a deterministic fake - only the host
logic is under test here; the CUDA path is covered by the -m gpu tests."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _fake_recognize(imgs):
    tok = np.zeros((len(imgs), 257), np.int32)
    ln = np.zeros(len(imgs), np.int32)
    for i, im in enumerate(imgs):
        n = 1 + int(im.shape[1]) % 40
        tok[i, 0] = 2
        tok[i, 1:n] = (np.arange(1, n) * int(im[0, 0]) + im.shape[0]) % 120 + 4
        ln[i] = n
    return tok, ln


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(REPO))
    import torch.distributed as dist
    from khmer_ocr_cnn_transformer_b200.distributed import recognize_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, (int(rng.integers(20, 60)), int(rng.integers(60, 1500))), dtype=np.uint8)
            for _ in range(37)]
    res = recognize_sharded(imgs, _fake_recognize)
    if rank == 0:
        tok, ln = res
        want_tok, want_ln = _fake_recognize(imgs)
        ok = np.array_equal(tok, want_tok) and np.array_equal(ln, want_ln)
        Path(out_path).write_text("ok" if ok else "mismatch")
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    out = tmp_path / "result.txt"
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
