"""The N > 1 product path on real hardware: two ranks (one per GPU, NCCL), `distributed.recognize_sharded` over
`LinePipeline.recognize` on mixed-width lines - the gathered, order-restored ids on rank 0 equal a single-GPU run of the same
lines (lines are independent: predictor.py:150-193).  Skipped on a one-GPU box (the gloo test covers the host logic there)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
    import torch
    import torch.distributed as dist
    from helpers import GOLDEN
    from khmer_ocr_cnn_transformer_b200 import weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.distributed import recognize_sharded
    from khmer_ocr_cnn_transformer_b200.pipeline import LinePipeline
    from workloads import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    imgs, _ = synth.make_lines(96, 200, 1600, seed=77)
    pipe = LinePipeline(weights.pack_blob(load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")), device=rank, in_flight=3,
                        max_lines=16, max_chunks=256)
    try:
        res = recognize_sharded(imgs, pipe.recognize, max_seq_len=pipe.max_seq_len, device=torch.device("cuda", rank))
        if rank == 0:
            tok, ln = res
            want_tok, want_ln = pipe.recognize(imgs)
            same = sum(np.array_equal(tok[i, :ln[i]], want_tok[i, :want_ln[i]]) for i in range(len(imgs)))
            Path(out_path).write_text(f"{same}/{len(imgs)} passes={pipe.stats['passes']}")
        else:
            assert res is None
        dist.barrier()
    finally:
        pipe.close()
        dist.destroy_process_group()


def test_two_rank_recognize_sharded_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, str(out)), nprocs=2, join=True)
    assert out.read_text().startswith("96/96"), out.read_text()
