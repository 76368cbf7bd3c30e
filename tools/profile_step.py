"""Profiling driver: two short warm-up passes (8 decode positions) then ONE full pass of the hot path over the
c2 batch.  Used under ncu (see profiles/README.md for the exact command lines)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict, load_checkpoint

root = Path(__file__).resolve().parent.parent
ck = root / "tests/golden/fixture_se_ckpt.npz"
sd = load_checkpoint(ck) if ck.exists() else seeded_state_dict("se", 0, max_global_len=1024)
n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rec = _native.Recognizer(weights.pack_blob(sd), max_lines=n_lines, max_chunks=n_lines * 11)
rec.set_option("use_graphs", 0)          # plain launches so that every kernel shows up by name
imgs, _ = synth.make_lines(n_lines, 400, 800, seed=0)
batch = _native.LineBatch(imgs)
for _ in range(2):
    rec.recognize_lines(batch, max_steps=8)
torch.cuda.synchronize()
tok, ln = rec.recognize_lines(batch, max_steps=steps)
torch.cuda.synchronize()
print("profiled pass done: chunks", int(rec.gather_chunks(batch).sum()), "mean len", float(ln.mean()))
rec.close()
