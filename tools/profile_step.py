"""Profiling driver: two short warm-up passes (8 decode positions) then ONE pass of the hot path over the c2 batch
(256 lines, 1885 chunks) inside a cudaProfilerStart/Stop range.  Used under ncu with `--profile-from-start off`
(tools/profile_r02.sh holds the exact command lines).
    python tools/profile_step.py [n_lines=256] [decode positions=24] [option=value ...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
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
rec.set_option("dec_wide", 0)            # the decode GEMM shapes bench.py runs (several passes in flight)
for a in sys.argv[3:]:
    k, v = a.split("=")
    rec.set_option(k, int(v))
imgs, _ = synth.make_lines(n_lines, 400, 800, seed=0)
batch = _native.LineBatch(imgs)
for _ in range(2):
    rec.recognize_lines(batch, max_steps=8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tok, ln = rec.recognize_lines(batch, max_steps=steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled pass done: chunks", int(rec.gather_chunks(batch).sum()), "mean len", float(ln.mean()))
rec.close()
