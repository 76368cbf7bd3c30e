"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer (GPU box):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from khmer_ocr_cnn_transformer_b200 import _native, weights, textline_crops as T
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint, seeded_state_dict

ROOT = Path(__file__).resolve().parent.parent
imgs = synth.make_lines(5, 90, 700, seed=4)[0] + [np.full((20, 30), 255, np.uint8)]
for variant in ("se", "vgg", "resnet"):
    sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz") if variant == "se" else seeded_state_dict(variant, 3, max_global_len=1024)
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=8, max_chunks=96)
    for opts in ({}, {"chunk_attn_impl": 0, "dec_cross_impl": 0, "kv_split": 0, "dec_wide": 0, "use_graphs": 0}):
        for k, v in opts.items():
            rec.set_option(k, v)
        tok, ln = rec.recognize_lines(_native.LineBatch(imgs), max_steps=20)
        for k in opts:
            rec.set_option(k, 1)
    print(variant, "greedy ok", ln.tolist(), flush=True)
    if variant == "se":
        rec.gather_chunks(_native.LineBatch(imgs)); rec.sevgg_encoder_forward(); rec.merge_bilstm_forward()
        pref = np.full((6, 1), 2, np.int32)
        lg = rec.beam_step_batch([0, 0, 1, 2, 2, 2], pref, None, 0)
        pref = np.concatenate([pref, lg.argmax(1)[:, None].astype(np.int32)], 1)
        rec.beam_step_batch([0, 0, 1, 2, 2, 2], pref, [1, 0, 2, 5, 3, 4], 1)
        rec.gather_chunks(_native.LineBatch(imgs)); rec.sevgg_encoder_forward()
        tgt = np.zeros((len(imgs), 9), np.int32); tgt[:, 0] = 2; tgt[:, 1:5] = 50
        out = rec.forward_teacher_forced(tgt)
        page = np.random.default_rng(0).integers(0, 256, (80, 200, 3), dtype=np.uint8)
        crops = T.crop_lines_device(rec, page, [(0, 0, 200, 30), (5, 40, 120, 80)], 10)
        b = crops.batch
        rec.recognize_lines(b, max_steps=8, pixels_dev_ptr=crops.dev_ptr)
        print("beam / teacher-forced / crops ok", out.shape, flush=True)
    rec.close()
print("sanitize smoke done")
