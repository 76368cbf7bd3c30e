"""Token-level parity of the CUDA path against the CPU oracle on the WHOLE c2 batch (256 lines), with a margin
analysis of every mismatch.  Runs on the GPU box (oracle on the host cores, multi-process).  Writes
gpurun_out/parity_c2.json; the summary is committed under profiles/."""
import json, sys, time, os
from pathlib import Path
from multiprocessing import Pool
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

def _oracle_line(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import recognizer_np as O
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    sd = _oracle_line.sd if hasattr(_oracle_line, "sd") else load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
    _oracle_line.sd = sd
    img = args
    ch = O.preprocess_gray(img)[1]
    enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, ch, "se")))
    mem = O.memory_for_line(sd, enc, "se")
    toks, logits = O.greedy_decode(sd, mem, return_logits=True)
    return toks, logits

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    from khmer_ocr_cnn_transformer_b200 import _native, weights, synth
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import build_vocab
    from oracle import recognizer_np as O
    sd = load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz")
    imgs, labels = synth.make_lines(256, 400, 800, seed=0)
    imgs = imgs[:n]
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=2816)
    tok, ln = rec.recognize_lines(_native.LineBatch(imgs))
    rec.close()
    t0 = time.time()
    with Pool(min(os.cpu_count(), 32)) as pool:
        ora = pool.map(_oracle_line, imgs, chunksize=2)
    dt = time.time() - t0
    idx2char = {v: k for k, v in build_vocab().items()}
    same, mism, cer_sum = 0, [], 0.0
    for i, (otoks, ologits) in enumerate(ora):
        g = [int(t) for t in tok[i, :ln[i]]]
        if g == otoks:
            same += 1
            continue
        k = next((j for j, (a, b) in enumerate(zip(g, otoks)) if a != b), min(len(g), len(otoks)))
        lg = ologits[k - 1] if 0 < k <= len(ologits) else None
        gap = None
        if lg is not None:
            top = np.sort(lg)[::-1]
            gap = float(top[0] - top[1])
        c = O.cer(O.tokens_to_text(g, idx2char), O.tokens_to_text(otoks, idx2char))
        cer_sum += c
        mism.append({"line": i, "first_diff_pos": k, "len_gpu": len(g), "len_oracle": len(otoks),
                     "oracle_top1_top2_gap_at_diff": gap, "cer_between": c})
    out = {"lines": n, "identical": same, "identity_rate": same / n, "mean_cer_gpu_vs_oracle": cer_sum / n,
           "mismatches": mism, "oracle_seconds": dt, "oracle_procs": min(os.cpu_count(), 32)}
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "parity_c2.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
