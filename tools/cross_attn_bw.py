"""Achieved HBM bandwidth of the decode attention kernels as a function of the lines per pass (one pass in flight,
first 16 positions = every line active).  Diagnosis tool, GPU box:  python tools/cross_attn_bw.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint

ROOT = Path(__file__).resolve().parent.parent
blob = weights.pack_blob(load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz"))
for n_lines in (256, 1024, 2048):
    imgs = []
    for b in range(n_lines // 256):
        imgs += synth.make_lines(256, 400, 800, seed=b)[0]
    rec = _native.Recognizer(blob, max_lines=n_lines, max_chunks=n_lines * 11)
    rec.set_option("use_pdl", 0)
    rec.set_option("dec_wide", 0)
    batch = _native.LineBatch(imgs)
    rec.recognize_lines(batch, max_steps=16)
    counts = rec.gather_chunks(batch)
    tokens = int(counts.sum()) * 32
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rec.recognize_lines(batch, max_steps=16)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or 0
        if t <= 0:
            continue
        if "dec_cross_attn" in e.key or "dec_self_attn" in e.key or "layernorm" in e.key or "gemm_tc_kernel<128, true>" in e.key \
                or "gemm_tc_kernel<(int)128, (bool)1>" in e.key:
            avg = t / e.count
            extra = ""
            if "dec_cross_attn" in e.key:
                mb = tokens * 1536 / 1e6
                extra = f"  {mb:.0f} MB per launch -> {mb / avg * 1e-3:.2f} TB/s"
            print(f"lines {n_lines:5d}  {e.key[:48]:48s} launches {e.count:4d}  avg {avg:8.1f} us{extra}")
    rec.close()
