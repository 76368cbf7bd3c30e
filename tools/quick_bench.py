"""Dev tool: per-stage CUDA-event timings of one batch (not the contract bench; see bench.py)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict, load_checkpoint

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ck = Path(__file__).resolve().parent.parent / ("tmp_ckpt.npz" if "--tmp" in sys.argv else "tests/golden/fixture_se_ckpt.npz")
sd = load_checkpoint(ck) if ck.exists() and "--seeded" not in sys.argv else seeded_state_dict("se", 0, max_global_len=1024)
rec = _native.Recognizer(weights.pack_blob(sd), max_lines=n_lines, max_chunks=n_lines * 12)
imgs, _ = synth.make_lines(n_lines, 400, 800, seed=0)
batch = _native.LineBatch(imgs)
pix = torch.from_numpy(batch.pixels).cuda()
res = {}
def timed(name, fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[name] = {"ms_min": min(ts), "ms_med": sorted(ts)[len(ts)//2]}
counts = rec.gather_chunks(batch, pixels_dev_ptr=pix.data_ptr())
nchunks = int(counts.sum())
timed("gather", lambda: rec.gather_chunks(batch, pixels_dev_ptr=pix.data_ptr()))
timed("cnn_enc", lambda: rec.sevgg_encoder_forward())
timed("bilstm_kv", lambda: rec.merge_bilstm_forward())
t0 = time.time(); tok, ln = rec.decode_greedy(n_lines); torch.cuda.synchronize(); t1 = time.time()
res["decode_wall_ms"] = (t1 - t0) * 1e3
res["decode_steps"] = int(rec.debug_read("last_steps"))
res["mean_len"] = float(ln.mean()); res["max_len"] = int(ln.max())
t0 = time.time(); rec.recognize_lines(batch); t1 = time.time()
res["e2e_wall_ms"] = (t1 - t0) * 1e3
res["n_lines"] = n_lines; res["n_chunks"] = nchunks
res["chunks_per_s_cnn_enc"] = nchunks / (res["cnn_enc"]["ms_min"] * 1e-3)
res["pct_bf16_peak"] = res["chunks_per_s_cnn_enc"] * 2.337e9 / 1618.1e12
res["lines_per_s_e2e"] = n_lines / (res["e2e_wall_ms"] * 1e-3)
print(json.dumps(res, indent=1))
