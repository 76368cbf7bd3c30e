"""BASELINE.json configs as parity / property tests (full sizes through size-independent properties):
  c1  one 48x400 line (5 chunks), batch 1                      -> tokens == reference golden (line 9 of golden_se)
  c3  mixed-width lines 200-1600 px, sharded                   -> batch-composition invariance + oracle on a sample
  c4  48x2400 line (29 chunks)                                 -> test_gpu_stages.test_long_line_c4
  c5  VGG baseline (no SE, no BiLSTM), batch 1024              -> batch-composition invariance + oracle on a sample
Lines are independent in the reference (predictor.py:150-193), so the decoded ids of a line must not depend on
which batch it travels in, on its neighbours, or on the order: that is checkable at any size without the oracle."""
import numpy as np
import pytest

from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _recognize_all(rec, imgs, max_lines, max_steps):
    from khmer_ocr_cnn_transformer_b200 import _native
    from khmer_ocr_cnn_transformer_b200.scheduling import plan_batches
    out = [None] * len(imgs)
    for idxs in plan_batches([im.shape for im in imgs], max_lines, rec.max_chunks, rec.max_seq_len):
        tok, ln = rec.recognize_lines(_native.LineBatch([imgs[i] for i in idxs]), max_steps=max_steps)
        for j, i in enumerate(idxs):
            out[i] = tok[j, :ln[j]].copy()
    return out


def test_c1_single_line_batch1_matches_reference():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    if not (GOLDEN / "golden_se.npz").exists():
        pytest.skip("golden_se.npz not generated")
    z = np.load(GOLDEN / "golden_se.npz")
    rec = _native.Recognizer(weights.pack_blob(load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")), max_lines=1, max_chunks=8)
    try:
        img = z["img9"]                                   # the line resized to exactly 48x400
        assert img.shape == (48, 400)
        counts = rec.gather_chunks(_native.LineBatch([img]))
        assert int(counts[0]) == 5                        # stride 84, `while start < W` -> 5 chunks, not 4
        tok, ln = rec.recognize_lines(_native.LineBatch([img]))
        assert np.array_equal(tok[0, :ln[0]], z["tokens9"])
    finally:
        rec.close()


def test_c3_mixed_width_batch_composition_invariance():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from workloads import synth
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.scheduling import shard_lines
    from oracle import recognizer_np as O
    sd = load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=256, max_chunks=3072)
    try:
        imgs, _ = synth.make_lines(768, 200, 1600, seed=33)          # 3-20 chunks per line
        base = _recognize_all(rec, imgs, 256, 48)
        # (a) other batch boundaries, (b) sharded like an 8-GPU run and processed shard by shard in shuffled order
        small = _recognize_all(rec, imgs, 61, 48)
        assert all(np.array_equal(a, b) for a, b in zip(base, small))
        rng = np.random.default_rng(0)
        for shard in shard_lines([im.shape for im in imgs], 8):
            order = rng.permutation(len(shard))
            got = _recognize_all(rec, [imgs[shard[j]] for j in order], 256, 48)
            for k, j in enumerate(order):
                assert np.array_equal(got[k], base[shard[j]])
        # (c) a sample against the oracle (teacher-free: memory only, decoding is covered elsewhere)
        counts = rec.gather_chunks(_native.LineBatch(imgs[:6]))
        rec.sevgg_encoder_forward(); rec.merge_bilstm_forward()
        mem = rec.debug_read("memory").reshape(-1, 384)
        cur = 0
        for i in range(6):
            ch = O.preprocess_gray(imgs[i])[1]
            want = O.memory_for_line(sd, O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, ch, "se"))), "se")
            assert rel_err(mem[cur * 32:cur * 32 + want.shape[0]], want) < 3e-2
            cur += int(counts[i])
    finally:
        rec.close()


def test_c5_vgg_baseline_batch_1024():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from workloads import synth
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    from oracle import recognizer_np as O
    sd = seeded_state_dict("vgg", seed=11, max_global_len=1024)
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=1024, max_chunks=4096)
    try:
        assert rec.variant == 1
        imgs, _ = synth.make_lines(1024, 100, 320, seed=55)          # short scene-text-like lines, 2-4 chunks
        base = _recognize_all(rec, imgs, 1024, 16)
        part = _recognize_all(rec, imgs, 100, 16)
        assert all(np.array_equal(a, b) for a, b in zip(base, part))
        counts = rec.gather_chunks(_native.LineBatch(imgs[:4]))
        rec.sevgg_encoder_forward(); rec.merge_bilstm_forward()
        mem = rec.debug_read("memory").reshape(-1, 384)
        cur = 0
        for i in range(4):
            ch = O.preprocess_gray(imgs[i])[1]
            want = O.memory_for_line(sd, O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, ch, "vgg"))), "vgg")
            assert rel_err(mem[cur * 32:cur * 32 + want.shape[0]], want) < 3e-2
            cur += int(counts[i])
    finally:
        rec.close()


def test_c5_vgg_trained_fixture_tokens_batch_1024():
    """BASELINE config c5 with a TRAINED VGG fixture (tools/train_fixture_gpu.py --variant vgg; the reference's VGG class reads
    the synthetic lines with CER 0): batch of 1024 short lines - tokens identical to the reference's own greedy outputs on
    the golden lines, identical to the oracle on a sample, CER against the labels reported."""
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from workloads import synth
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import build_vocab
    from oracle import recognizer_np as O
    ck, gp = GOLDEN / "fixture_vgg_ckpt.npz", GOLDEN / "golden_vgg_trained.npz"
    if not (ck.exists() and gp.exists()):
        pytest.skip("trained VGG fixture not generated")
    z = np.load(gp)
    sd = load_checkpoint(ck)
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=1024, max_chunks=4096)
    try:
        assert rec.variant == 1
        imgs, labels = synth.make_lines(1024, 100, 320, seed=55)
        tok, ln = rec.recognize_lines(_native.LineBatch(imgs))
        for i in range(8):                                           # golden lines 0..7 are the first lines of this batch
            assert np.array_equal(z[f"img{i}"], imgs[i])
            assert np.array_equal(tok[i, :ln[i]], z[f"tokens{i}"]), i
        sample = list(range(8, 1024, 64))
        want = O.recognise_lines(sd, [imgs[i] for i in sample], "vgg")
        same = [np.array_equal(tok[i, :ln[i]], np.asarray(w)) for i, w in zip(sample, want)]
        idx2char = {v: k for k, v in build_vocab().items()}
        cer = float(np.mean([O.cer(O.tokens_to_text([int(t) for t in tok[i, :ln[i]]], idx2char),
                                   O.tokens_to_text([int(t) for t in labels[i]], idx2char)) for i in range(256)]))
        from test_gpu_stages import _report
        _report("c5_vgg_trained", {"oracle_sample_same": int(sum(same)), "of": len(same), "mean_cer_vs_labels_256": cer})
        assert all(same), same
        assert cer < 0.02
        long_imgs = [z["img8"], z["img9"]]                          # two longer lines (500-900 px) of the golden
        t2, l2 = rec.recognize_lines(_native.LineBatch(long_imgs))
        for j, i in enumerate((8, 9)):
            assert np.array_equal(t2[j, :l2[j]], z[f"tokens{i}"])
    finally:
        rec.close()


def test_handles_on_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in) are per device: a handle created on cuda:1 AFTER one on cuda:0 in the
    same process must work and decode the same ids, also through the model-level pipeline object.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.pipeline import LinePipeline
    z = np.load(GOLDEN / "golden_se.npz")
    blob = weights.pack_blob(load_checkpoint(GOLDEN / "fixture_se_ckpt.npz"))
    imgs = [z[f"img{i}"] for i in range(int(z["n_lines"]))]
    recs = [_native.Recognizer(blob, device=d, max_lines=16, max_chunks=256) for d in (0, 1)]
    try:
        outs = [r.recognize_lines(_native.LineBatch(imgs)) for r in recs]
        outs.append(recs[0].recognize_lines(_native.LineBatch(imgs)))        # back on device 0 after device 1 was current
        for tok, ln in outs:
            for i in range(len(imgs)):
                assert np.array_equal(tok[i, :ln[i]], z[f"tokens{i}"]), i
    finally:
        for r in recs:
            r.close()
    pipe = LinePipeline(blob, device=1, in_flight=2, max_lines=4, max_chunks=64)
    try:
        tok, ln = pipe.recognize(imgs)
        for i in range(len(imgs)):
            assert np.array_equal(tok[i, :ln[i]], z[f"tokens{i}"]), i
    finally:
        pipe.close()
