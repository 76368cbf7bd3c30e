"""GPU-box check of the reference arm: the unmodified reference (baseline/_ref) on CPU and on cuda, 8 lines of c2."""
import sys, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from baseline import run_reference as R
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
from workloads import synth
print("available:", R.available(), "cores", os.cpu_count())
sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
imgs, _ = synth.make_lines(16, 400, 800, seed=0)
out = {}
for dev in ("cpu", "cuda"):
    pred = R.load_predictor(sd, dev)
    R.time_predict_batch(pred, imgs[:2])
    v, dt, texts = R.time_predict_batch(pred, imgs)
    out[dev] = {"lines_per_s": v, "s": dt}
    out[dev + "_texts"] = texts
print(json.dumps({k: v for k, v in out.items() if not k.endswith("_texts")}))
print("cpu == cuda texts:", out["cpu_texts"] == out["cuda_texts"])
