"""Dev probe: se_excite variants (option "se_variant": 0 = fragment-order FC weights, pooled block prefetched into registers,
2 CTAs per SM; 3 / 4 = block streamed after the gate, 3 / 4 CTAs per SM; 100 + v = the same with row-major FC weights) - per-kernel CUDA-event times of stages 2-4 and bit-equality of the outputs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from khmer_ocr_cnn_transformer_b200 import _native, weights
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
from workloads import synth
sd = load_checkpoint(Path(__file__).resolve().parent.parent / "tests/golden/fixture_se_ckpt.npz")
rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=2816)
imgs, _ = synth.make_lines(256, 200, 1600, seed=3)
imgs.sort(key=lambda im: -im.shape[1])
batch = _native.LineBatch(imgs[:150])
n = int(rec.gather_chunks(batch).sum())
print("chunks", n, flush=True)
ref = None
for variant in (100, 0, 4, 104, 100, 0, 4):
    rec.set_option("se_variant", variant)
    rec.gather_chunks(batch)
    for _ in range(3): rec.sevgg_encoder_forward()
    torch.cuda.synchronize()
    out = rec.debug_read("patch_in").copy()
    if ref is None: ref = out
    rec.set_option("kernel_timing", 1)
    for _ in range(10): rec.sevgg_encoder_forward()
    kt = rec.kernel_timing()
    rec.set_option("kernel_timing", 0)
    se = {k: round(kt[k]["ms"] / 10, 4) for k in ("se3_excite", "se4_excite", "se5_excite_finalpool")}
    total = sum(v["ms"] for k, v in kt.items() if not k.startswith(("lstm", "bilstm", "cross"))) / 10
    print("se_variant", variant, se, "stage sum ms", round(total, 3), "bit-identical to the first variant:", bool(np.array_equal(out, ref)), flush=True)
rec.set_option("se_variant", 4)
