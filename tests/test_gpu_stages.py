"""Parity of the CUDA path (through the C ABI) against the numpy oracle, stage by stage.

Tolerances: stage 1 is bit-exact.  Stages 2-5 compute GEMMs with bf16 operands and fp32
accumulation and keep inter-layer activations in bf16, the oracle is fp32: the bar is a relative
L2 error <= 3e-2 per intermediate (measured values are written to gpurun_out/parity_report.json)."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, bf16_u16_to_f32, nwhc_to_nchw, rel_err, max_err

pytestmark = pytest.mark.gpu
REPORT = {}


def _report(key, value):
    REPORT[key] = value
    out = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parent.parent)) / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / "parity_report.json").write_text(json.dumps(REPORT, indent=1))
    except OSError:
        pass


@pytest.fixture(scope="module")
def rec_seeded_se():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    sd = seeded_state_dict("se", 0, max_global_len=1024)
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=64, max_chunks=512)
    yield rec, sd
    rec.close()


@pytest.fixture(scope="module")
def rec_seeded_vgg():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    sd = seeded_state_dict("vgg", 11, max_global_len=1024)
    rec = _native.Recognizer(weights.pack_blob(sd), max_lines=64, max_chunks=512)
    yield rec, sd
    rec.close()


def _lines(n, lo, hi, seed):
    from workloads import synth
    return synth.make_lines(n, lo, hi, seed=seed)[0]


# ------------------------------------------------------------------------------------------------
def test_stage1_bit_exact_random_and_ragged(rec_seeded_se):
    from khmer_ocr_cnn_transformer_b200 import _native
    from oracle import recognizer_np as O
    rec, _ = rec_seeded_se
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (h, w) in
            [(30, 375), (48, 400), (61, 333), (20, 40), (48, 100), (96, 1000), (17, 911), (50, 52),
             (48, 84), (48, 85), (7, 9), (200, 3000), (48, 50), (31, 31)]]
    imgs += _lines(8, 100, 1600, seed=3)
    counts = rec.gather_chunks(_native.LineBatch(imgs))
    got = rec.debug_read("chunks").reshape(-1, 1, 48, 100)
    want = [O.preprocess_gray(im)[1] for im in imgs]
    assert [int(c) for c in counts] == [w.shape[0] for w in want]
    want = np.concatenate(want, 0)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "stage 1 must be bit-exact"
    _report("stage1_bit_exact_chunks", int(want.shape[0]))


def test_stage1_against_reference_golden(rec_seeded_se):
    from khmer_ocr_cnn_transformer_b200 import _native
    p = GOLDEN / "golden_preprocess.npz"
    if not p.exists():
        pytest.skip("golden_preprocess.npz not generated")
    rec, _ = rec_seeded_se
    z = np.load(p)
    n = len([k for k in z.files if k.startswith("img")])
    imgs = [z[f"img{i}"] for i in range(n)]
    rec.gather_chunks(_native.LineBatch(imgs))
    got = rec.debug_read("chunks").reshape(-1, 1, 48, 100)
    want = np.concatenate([z[f"chunks{i}"] for i in range(n)], 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_stage1_empty_batch_and_errors(rec_seeded_se):
    from khmer_ocr_cnn_transformer_b200 import _native
    rec, _ = rec_seeded_se
    assert len(rec.gather_chunks(_native.LineBatch([]))) == 0
    tok, ln = rec.recognize_lines(_native.LineBatch([]))
    assert tok.shape[0] == 0
    with pytest.raises(_native.KocrError, match="down-scale"):
        rec.gather_chunks(_native.LineBatch([np.full((48 * 40, 60), 255, np.uint8)]))
    with pytest.raises(_native.KocrError, match="exceed capacity"):
        rec.gather_chunks(_native.LineBatch([np.full((48, 100), 255, np.uint8)] * 65))


# ------------------------------------------------------------------------------------------------
# conv4 / conv6 / conv7 never store their un-pooled output (column-fused GEMM epilogue): they are checked through the SE column
# means they emit (= mean over H of the oracle's conv4 / conv6 taps), the gated + pooled pool3 / pool4 and the final pool
GEOM = {"pool1": (24, 50, 64), "pool2": (12, 25, 128), "conv3": (12, 25, 256),
        "pool3": (6, 25, 256), "conv5": (6, 25, 512), "pool4": (3, 25, 512)}


def _check_backbone(rec, sd, variant, imgs, tag):
    from khmer_ocr_cnn_transformer_b200 import _native
    from oracle import recognizer_np as O
    counts = rec.gather_chunks(_native.LineBatch(imgs))
    rec.sevgg_encoder_forward()
    rec.merge_bilstm_forward()
    n = int(counts.sum())
    chunks = np.concatenate([O.preprocess_gray(im)[1] for im in imgs], 0)
    taps = {}
    f = O.cnn_forward(sd, chunks, variant, taps)
    worst = {}
    for name, (H, W, C) in GEOM.items():
        worst[name] = rel_err(nwhc_to_nchw(rec.debug_read(name), n, H, W, C), taps[name])
    if variant == "se":
        for name, tap, C in (("se_mean3", "conv4", 256), ("se_mean4", "conv6", 512)):
            got = bf16_u16_to_f32(rec.debug_read(name)).reshape(n, 25, C)      # [n][w][c]
            worst[name] = rel_err(got, taps[tap].mean(axis=2).transpose(0, 2, 1))
    else:           # VGG baseline: conv7 is a bare conv; its epilogue writes the two adaptive-pool row bins as sums
        y = taps["conv7"]                                                      # (n, 512, 3, 25)
        want = np.stack([y[:, :, 0] + y[:, :, 1], y[:, :, 1] + y[:, :, 2]], axis=2)   # (n, 512, 2, 25)
        got = bf16_u16_to_f32(rec.debug_read("bins7")).reshape(n, 25, 2, 512).transpose(0, 3, 2, 1)
        worst["conv7_row_bins"] = rel_err(got, want)
    # conv7 tap in the oracle is post-SE; compare through the final pool / patch operand instead
    pin = bf16_u16_to_f32(rec.debug_read("patch_in")).reshape(n, 32, 2, 512)     # [n][k][kh][c]
    worst["final_pool"] = rel_err(pin.transpose(0, 3, 2, 1), f)
    enc_o = O.encoder_forward(sd, O.patch_forward(sd, f))
    enc_g = rec.debug_read("enc").reshape(n, 32, 384)
    mem_g = rec.debug_read("memory").reshape(n * 32, 384)
    cur, e_err, m_err = 0, [], []
    for c in counts:
        c = int(c)
        merged = O.merge_line(sd, enc_o[cur:cur + c])
        T = merged.shape[0]
        e_err.append(rel_err(enc_g[cur:cur + c].reshape(-1, 384)[:T], merged))
        m_err.append(rel_err(mem_g[cur * 32:cur * 32 + T], O.memory_for_line(sd, enc_o[cur:cur + c], variant)))
        cur += c
    worst["enc+global_pos"] = max(e_err)
    worst["memory"] = max(m_err)
    _report(f"backbone_rel_err_{tag}", worst)
    for k, v in worst.items():
        assert v < 3e-2, f"{tag}/{k}: relative error {v:.4f} exceeds 3e-2 ({worst})"


def test_backbone_and_encoder_parity_se(rec_seeded_se):
    rec, sd = rec_seeded_se
    _check_backbone(rec, sd, "se", _lines(5, 100, 900, seed=21), "se_seeded")


def test_backbone_and_encoder_parity_vgg(rec_seeded_vgg):
    rec, sd = rec_seeded_vgg
    _check_backbone(rec, sd, "vgg", _lines(4, 100, 700, seed=22), "vgg_seeded")


@pytest.mark.parametrize("lstm_impl", [1, 0])
def test_long_line_c4(rec_seeded_se, lstm_impl):
    """48x2400 line = 29 chunks, T = 928 (BASELINE config 4): BiLSTM over a long merged sequence, with both
    recurrence kernels (1 = tensor-core fragments in registers, 0 = CUDA-core with W_hh in shared memory)."""
    from khmer_ocr_cnn_transformer_b200 import _native
    from oracle import recognizer_np as O
    rec, sd = rec_seeded_se
    rec.set_option("lstm_impl", lstm_impl)
    img = _lines(1, 2400, 2400, seed=9)[0]
    from PIL import Image
    img = np.asarray(Image.fromarray(img).resize((2400, 48), Image.Resampling.BILINEAR))
    counts = rec.gather_chunks(_native.LineBatch([img, _lines(1, 300, 300, seed=10)[0]]))
    assert int(counts[0]) == 29
    rec.sevgg_encoder_forward()
    rec.merge_bilstm_forward()
    chunks = O.preprocess_gray(img)[1]
    enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "se")))
    mem = O.memory_for_line(sd, enc, "se")
    got = rec.debug_read("memory").reshape(-1, 384)[:928]
    rec.set_option("lstm_impl", 1)
    e = rel_err(got, mem)
    _report(f"c4_memory_rel_err_lstm_impl{lstm_impl}", e)
    assert e < 3e-2


# ------------------------------------------------------------------------------------------------
def _teacher_forced_logits(rec, imgs, token_rows, steps):
    from khmer_ocr_cnn_transformer_b200 import _native
    rec.set_option("trace_logits", 1)
    rec.set_option("force_tokens", 1)
    forced = np.zeros((len(imgs), 257), np.int32)
    for i, t in enumerate(token_rows):
        forced[i, :len(t)] = t
    rec.set_forced_tokens(forced)
    try:
        rec.recognize_lines(_native.LineBatch(imgs), max_steps=steps)
        return rec.debug_read("logits_trace").reshape(len(imgs), 256, 128)[:, :steps, :124]
    finally:
        rec.set_option("trace_logits", 0)
        rec.set_option("force_tokens", 0)


def test_decoder_teacher_forced_logits_seeded(rec_seeded_se):
    """Same prefix on both sides (incl. a <pad> token, which becomes a masked key: se_model.py:190)."""
    from oracle import recognizer_np as O
    rec, sd = rec_seeded_se
    imgs = _lines(3, 200, 700, seed=31)
    rng = np.random.default_rng(0)
    rows = []
    for i in range(3):
        t = [2] + [int(x) for x in rng.integers(4, 124, 11)]
        t[4 + i] = 0
        rows.append(t)
    got = _teacher_forced_logits(rec, imgs, rows, steps=12)
    errs = []
    for i, im in enumerate(imgs):
        chunks = O.preprocess_gray(im)[1]
        enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "se")))
        mem = O.memory_for_line(sd, enc, "se")
        want = O.decoder_forward(sd, rows[i], mem)           # (12, 124)
        errs.append(rel_err(got[i], want))
    _report("teacher_forced_logits_rel_err_seeded", errs)
    assert max(errs) < 3e-2, errs


def test_greedy_decode_semantics_seeded(rec_seeded_se):
    """Free-running greedy decode: <sos> first, no <eos> stored, length <= 257, and the tokens are the
    argmax of the library's own traced logits (ties -> lowest index)."""
    from khmer_ocr_cnn_transformer_b200 import _native
    rec, sd = rec_seeded_se
    imgs = _lines(6, 100, 600, seed=41)
    rec.set_option("trace_logits", 1)
    try:
        tok, ln = rec.recognize_lines(_native.LineBatch(imgs), max_steps=40)
        trace = rec.debug_read("logits_trace").reshape(len(imgs), 256, 128)
    finally:
        rec.set_option("trace_logits", 0)
    for i in range(len(imgs)):
        assert tok[i, 0] == 2 and 1 <= ln[i] <= 41
        ids = tok[i, :ln[i]]
        assert 3 not in ids[1:]
        for t in range(ln[i] - 1):
            assert ids[t + 1] == int(np.argmax(trace[i, t, :124]))
        if ln[i] < 41:
            assert int(np.argmax(trace[i, ln[i] - 1, :124])) == 3      # stopped because of <eos>


def test_chunk_attention_mma_agrees_with_cuda_core_kernel(rec_seeded_se):
    """Per-chunk encoder attention on mma.sync (bf16 probabilities, fp32 accumulation) against the fp32 CUDA-core
    kernel: encoder outputs agree far inside the parity tolerance."""
    from khmer_ocr_cnn_transformer_b200 import _native
    rec, _ = rec_seeded_se
    imgs = _lines(6, 100, 1200, seed=22)
    got = {}
    try:
        for mode in (0, 1):
            rec.set_option("chunk_attn_impl", mode)
            rec.gather_chunks(_native.LineBatch(imgs))
            rec.sevgg_encoder_forward()
            got[mode] = rec.debug_read("enc").copy()
    finally:
        rec.set_option("chunk_attn_impl", 1)
    err = rel_err(got[1], got[0])
    _report("chunk_attention_mma_vs_fp32_rel_err", err)
    assert err < 5e-3, err
