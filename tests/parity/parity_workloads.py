"""Token-level parity of the CUDA path against the CPU oracle on whole workloads, with a margin analysis of every
mismatch and the CER of both sides against the ground-truth labels.

  workloads: c2 = the 256-line bench batch (widths 400-800, seed 0); c3 = 1024 lines of the mixed-width config
             (widths 200-1600, seed 3); c3full = all 8192 lines of that config (the oracle takes about an hour on 8 cores)
  GPU box :  python tests/parity/parity_workloads.py dump [c2|c3]    -> gpurun_out/<w>_tokens.npz  (tokens + lengths, CUDA path)
  anywhere:  python tests/parity/parity_workloads.py oracle [c2|c3]  -> tests/golden/oracle_tokens_<w>.npz (numpy oracle, host cores)
  anywhere:  python tests/parity/parity_workloads.py compare [c2|c3] -> profiles/r02/parity_<w>.json
"""
import json, sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

WL = sys.argv[2] if len(sys.argv) > 2 else "c2"
SPEC = {"c2": (256, 400, 800, 0), "c3": (1024, 200, 1600, 3), "c3full": (8192, 200, 1600, 3)}[WL]
ORACLE_NPZ = ROOT / "tests" / "golden" / f"oracle_tokens_{WL}.npz"
GPU_NPZ = ROOT / "gpurun_out" / f"{WL}_tokens.npz"


def _oracle_line(img):
    from threadpoolctl import threadpool_limits
    from oracle import recognizer_np as O
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    with threadpool_limits(limits=1):
        if not hasattr(_oracle_line, "sd"):
            _oracle_line.sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
        sd = _oracle_line.sd
        ch = O.preprocess_gray(img)[1]
        enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, ch, "se")))
        mem = O.memory_for_line(sd, enc, "se")
        toks, logits = O.greedy_decode(sd, mem, return_logits=True)
    top = np.sort(logits, axis=1)[:, ::-1]
    return toks, (top[:, 0] - top[:, 1]).astype(np.float32)


def lines(with_labels=False):
    from workloads import synth
    imgs, labels = synth.make_lines(SPEC[0], SPEC[1], SPEC[2], seed=SPEC[3])
    return (imgs, labels) if with_labels else imgs


def dump():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=5120)
    for opt in sys.argv[3:]:                      # e.g. lstm_split=1 kv_split=0: numerics experiments against the stored oracle tokens
        k, v = opt.split("=")
        rec.set_option(k, int(v))
    imgs = lines()
    toks, lns = [], []
    for i in range(0, len(imgs), 256):
        t, l = rec.recognize_lines(_native.LineBatch(imgs[i:i + 256]))
        toks.append(t.copy()); lns.append(l.copy())
    tok, ln = np.concatenate(toks), np.concatenate(lns)
    rec.close()
    GPU_NPZ.parent.mkdir(exist_ok=True)
    np.savez_compressed(GPU_NPZ, tokens=tok, lengths=ln)
    print("wrote", GPU_NPZ, "mean len", float(ln.mean()))


def oracle():
    from multiprocessing import Pool
    t0 = time.time()
    with Pool(int(os.environ.get("ORACLE_PROCS", "8"))) as pool:
        res = pool.map(_oracle_line, lines(), chunksize=1)
    n = SPEC[0]
    tokens = np.zeros((n, 257), np.int32)
    lengths = np.zeros(n, np.int32)
    gaps = np.zeros((n, 256), np.float32)
    for i, (t, g) in enumerate(res):
        tokens[i, :len(t)] = t
        lengths[i] = len(t)
        gaps[i, :len(g)] = g
    np.savez_compressed(ORACLE_NPZ, tokens=tokens, lengths=lengths, gaps=gaps)
    print("wrote", ORACLE_NPZ, f"{time.time()-t0:.0f} s")


def compare():
    from oracle import recognizer_np as O
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import build_vocab
    g, o = np.load(GPU_NPZ), np.load(ORACLE_NPZ)
    idx2char = {v: k for k, v in build_vocab().items()}
    same, mism, cer_sum = 0, [], 0.0
    n = SPEC[0]
    labels = lines(with_labels=True)[1]
    cer_gpu = cer_orc = 0.0
    for i in range(n):
        a = [int(t) for t in g["tokens"][i, :g["lengths"][i]]]
        b = [int(t) for t in o["tokens"][i, :o["lengths"][i]]]
        truth = O.tokens_to_text([int(t) for t in labels[i]], idx2char)
        cer_gpu += O.cer(O.tokens_to_text(a, idx2char), truth)
        cer_orc += O.cer(O.tokens_to_text(b, idx2char), truth)
        if a == b:
            same += 1
            continue
        k = next((j for j, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        c = O.cer(O.tokens_to_text(a, idx2char), O.tokens_to_text(b, idx2char))
        cer_sum += c
        mism.append({"line": i, "first_diff_pos": k, "len_gpu": len(a), "len_oracle": len(b),
                     "oracle_top1_top2_gap_at_diff": float(o["gaps"][i, k - 1]) if k >= 1 else None, "cer_between": c})
    gaps = o["gaps"][o["gaps"] > 0]
    out = {"workload": f"{WL}: {n} synthetic lines, resized width {SPEC[1]}-{SPEC[2]}, fixture checkpoint", "identical": same, "of": n,
           "identity_rate": same / n, "mean_cer_gpu_vs_oracle_all_lines": cer_sum / n,
           "mean_cer_vs_labels": {"cuda": cer_gpu / n, "oracle": cer_orc / n},
           "mean_decoded_len": {"cuda": float(g["lengths"].mean()), "oracle": float(o["lengths"].mean())}, "mismatches": mism,
           "oracle_margin_percentiles": {p: float(np.percentile(gaps, p)) for p in (0.1, 1, 5, 50)}}
    (ROOT / "profiles" / "r02" / f"parity_{WL}.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"dump": dump, "oracle": oracle, "compare": compare}[sys.argv[1]]()
