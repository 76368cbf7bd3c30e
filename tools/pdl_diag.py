import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
sd = load_checkpoint(Path(__file__).resolve().parent.parent / "tests/golden/fixture_se_ckpt.npz")
rec = _native.Recognizer(weights.pack_blob(sd), max_lines=64, max_chunks=640)
imgs, _ = synth.make_lines(8, 400, 800, seed=0)
batch = _native.LineBatch(imgs)
rec.set_option("use_graphs", 0)
names = ["dx", "dqkv", "daof", "dy", "dq", "dh", "logits"]
order = ["embed", "qkv", "selfattn", "out", "ln1", "q", "cross", "out2", "ln2", "ffn1", "ffn2", "ln3"]
for stop in range(1, 14):
    snaps = {}
    for pdl in (0, 1):
        rec.set_option("use_pdl", pdl); rec.set_option("debug_stop", stop)
        rec.recognize_lines(batch, max_steps=1)
        snaps[pdl] = {n: rec.debug_read(n).copy() for n in names}
    diffs = {n: float(np.abs(snaps[0][n] - snaps[1][n]).max()) for n in names}
    print("first", stop, "kernels (last =", order[stop - 1] if stop <= len(order) else "...", ") max|pdl0-pdl1|:", {k: round(v, 5) for k, v in diffs.items()})
