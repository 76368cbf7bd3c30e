"""Token-level parity of the CUDA path against the CPU oracle on the WHOLE c2 batch (256 lines), with a margin
analysis of every mismatch.

  GPU box :  python tools/parity_c2.py dump      -> gpurun_out/c2_tokens.npz  (tokens + lengths from the CUDA path)
  anywhere:  python tools/parity_c2.py oracle    -> profiles/r01/c2_oracle_tokens.npz (numpy oracle, host cores)
  anywhere:  python tools/parity_c2.py compare   -> profiles/r01/parity_c2.json
"""
import json, sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

ORACLE_NPZ = ROOT / "profiles" / "r01" / "c2_oracle_tokens.npz"
GPU_NPZ = ROOT / "gpurun_out" / "c2_tokens.npz"


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


def lines():
    from khmer_ocr_cnn_transformer_b200 import synth
    return synth.make_lines(256, 400, 800, seed=0)[0]


def dump():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=2816)
    tok, ln = rec.recognize_lines(_native.LineBatch(lines()))
    rec.close()
    GPU_NPZ.parent.mkdir(exist_ok=True)
    np.savez_compressed(GPU_NPZ, tokens=tok, lengths=ln)
    print("wrote", GPU_NPZ, "mean len", float(ln.mean()))


def oracle():
    from multiprocessing import Pool
    t0 = time.time()
    with Pool(int(os.environ.get("ORACLE_PROCS", "8"))) as pool:
        res = pool.map(_oracle_line, lines(), chunksize=1)
    tokens = np.zeros((256, 257), np.int32)
    lengths = np.zeros(256, np.int32)
    gaps = np.zeros((256, 256), np.float32)
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
    for i in range(256):
        a = [int(t) for t in g["tokens"][i, :g["lengths"][i]]]
        b = [int(t) for t in o["tokens"][i, :o["lengths"][i]]]
        if a == b:
            same += 1
            continue
        k = next((j for j, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        c = O.cer(O.tokens_to_text(a, idx2char), O.tokens_to_text(b, idx2char))
        cer_sum += c
        mism.append({"line": i, "first_diff_pos": k, "len_gpu": len(a), "len_oracle": len(b),
                     "oracle_top1_top2_gap_at_diff": float(o["gaps"][i, k - 1]) if k >= 1 else None, "cer_between": c})
    gaps = o["gaps"][o["gaps"] > 0]
    out = {"workload": "c2 batch: 256 synthetic lines, fixture checkpoint", "identical": same, "of": 256,
           "identity_rate": same / 256, "mean_cer_gpu_vs_oracle_all_lines": cer_sum / 256, "mismatches": mism,
           "oracle_margin_percentiles": {p: float(np.percentile(gaps, p)) for p in (0.1, 1, 5, 50)}}
    (ROOT / "profiles" / "r01" / "parity_c2.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"dump": dump, "oracle": oracle, "compare": compare}[sys.argv[1]]()
