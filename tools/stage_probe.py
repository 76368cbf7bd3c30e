"""Dev probe: per-kernel timings of stages 2-5a back-to-back vs interleaved with the decode loop."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
sd = load_checkpoint(Path(__file__).resolve().parent.parent / "tests/golden/fixture_se_ckpt.npz")
rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=2816)
imgs, _ = synth.make_lines(256, 400, 800, seed=0)
batch = _native.LineBatch(imgs)
keys = ["conv1_pool1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "enc_qkv", "enc_ffn1", "enc_ffn2", "patch_proj"]
def show(tag, kt, n):
    print(tag, {k: round(kt[k]["ms"] / n, 3) for k in keys}, flush=True)
rec.gather_chunks(batch)
for _ in range(3): rec.sevgg_encoder_forward()
torch.cuda.synchronize()
rec.set_option("kernel_timing", 1)
for _ in range(10): rec.sevgg_encoder_forward()
show("back-to-back stage only      ", rec.kernel_timing(), 10)
for steps in (8, 64, 256):
    rec.set_option("kernel_timing", 1)
    for _ in range(6): rec.recognize_lines(batch, max_steps=steps)
    show(f"with decode max_steps={steps:3d}   ", rec.kernel_timing(), 6)
import time
rec.set_option("kernel_timing", 1)
for _ in range(6):
    time.sleep(0.05); rec.gather_chunks(batch); rec.sevgg_encoder_forward(); torch.cuda.synchronize()
show("stage only after 50 ms idle  ", rec.kernel_timing(), 6)
