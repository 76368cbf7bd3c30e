"""CPU tests of the host-side logic and of the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

from khmer_ocr_cnn_transformer_b200 import _native, weights, scheduling
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict, state_dict_spec, validate_state_dict
from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import Tokenizer, build_vocab
from khmer_ocr_cnn_transformer_b200.recognition.config import OCRConfig

REPO = Path(__file__).resolve().parent.parent
REC = REPO / "khmer_ocr_cnn_transformer_b200" / "recognition"


def test_vocab_layout():
    v = build_vocab()
    assert len(v) == 124
    assert [v["<pad>"], v["<unk>"], v["<sos>"], v["<eos>"], v[" "]] == [0, 1, 2, 3, 4]
    assert v["ក"] == 42 and v["‹"] == 122 and v["›"] == 123
    assert json.load(open(REC / "char2idx.json", encoding="utf-8")) == v
    ref = Path("/root/reference/netra_ocr/recognition/char2idx.json")
    if ref.exists():
        assert json.load(open(ref, encoding="utf-8")) == v


def test_tokenizer_decode():
    tok = Tokenizer(REC / "char2idx.json")
    assert (tok.sos_idx, tok.eos_idx, tok.pad_idx, len(tok)) == (2, 3, 0, 124)
    assert tok.decode([2, 42, 0, 4, 43, 3, 44]) == "ក ខ"      # skip sos/pad, stop at eos
    with pytest.raises(FileNotFoundError):
        Tokenizer(REC / "nope.json")


def test_config_defaults():
    c = OCRConfig()
    assert (c.img_height, c.chunk_width, c.chunk_overlap, c.emb_dim, c.max_seq_len, c.decode_max_len) == \
        (48, 100, 16, 384, 4096, 256)


def test_state_dict_spec_counts():
    se = state_dict_spec("se")
    vgg = state_dict_spec("vgg")
    # 137 entries in the reference = 116 parameters + 14 running stats + 7 num_batches_tracked
    assert len(se) == 137 - 7 and len(vgg) == 112 - 6
    n_se = sum(int(np.prod(s)) for k, s in se.items() if "running" not in k)
    assert n_se == 17_579_980                       # SURVEY.md §2.1 parameter count


def test_bf16_rounding_matches_torch():
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.0000001, 65504.0]
    mine = weights.f32_to_bf16_bits(x)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(mine, ref)


def test_a16_conversion_is_fp16_rne_saturating():
    """The library's 16-bit operand format (csrc/common.cuh): fp16, round-to-nearest-even, saturating."""
    import torch
    assert weights.A16_FORMAT == 1
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    x[:6] = [0.0, -0.0, 1.0000001, 65504.0, 1e6, -1e6]
    mine = weights.f32_to_a16_bits(x)
    ref = torch.from_numpy(np.clip(x, -65504, 65504)).to(torch.float16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(mine, ref)
    back = weights.a16_bits_to_f32(mine)
    assert back[4] == 65504.0 and back[5] == -65504.0 and np.all(np.isfinite(back))
    assert np.abs(back[6:] - x[6:]).max() <= np.abs(x[6:]).max() * 2.0 ** -11


def test_pack_blob_layout_and_folding():
    sd = seeded_state_dict("se", 3)
    assert validate_state_dict(sd)[0] == "se"
    t = weights.pack_tensors(sd)
    # BN folding identity on one output channel
    w, b = weights.fold_bn(sd["cnn.conv2.0.weight"], sd["cnn.conv2.0.bias"], sd["cnn.conv2.1.weight"],
                           sd["cnn.conv2.1.bias"], sd["cnn.conv2.1.running_mean"], sd["cnn.conv2.1.running_var"])
    x = np.random.default_rng(1).standard_normal((64, 3, 3)).astype(np.float32)
    pre = (sd["cnn.conv2.0.weight"][5] * x).sum() + sd["cnn.conv2.0.bias"][5]
    bn = (pre - sd["cnn.conv2.1.running_mean"][5]) / np.sqrt(sd["cnn.conv2.1.running_var"][5] + 1e-5) * \
        sd["cnn.conv2.1.weight"][5] + sd["cnn.conv2.1.bias"][5]
    assert abs(((w[5] * x).sum() + b[5]) - bn) < 1e-4
    # K-major conv layout: k = (r*3+s)*Cin + c
    km = weights.conv_to_kmajor(w)
    assert km.shape == (128, 576) and km[7, (1 * 3 + 2) * 64 + 9] == w[7, 9, 1, 2]
    # patch layout: k = kh*512 + c
    pw = weights.a16_bits_to_f32(t["patch.w"][1]).reshape(384, 1024)
    ref = weights.a16_bits_to_f32(weights.f32_to_a16_bits(sd["patch.proj.weight"][3, 17, 1, 0].reshape(1)))[0]
    assert pw[3, 512 + 17] == ref
    # LSTM recurrent packing: [dir][rank][kp][row][pair]
    whh = weights.a16_bits_to_f32(t["lstm.w_hh"][1]).reshape(2, 2, 96, 384, 2)
    src = sd["context_bilstm.weight_hh_l0_reverse"]
    gate, rank, jj, kp, pair = 2, 1, 40, 33, 1
    want = weights.a16_bits_to_f32(weights.f32_to_a16_bits(src[gate * 192 + rank * 96 + jj, 2 * kp + pair].reshape(1)))[0]
    assert whh[1, rank, kp, gate * 96 + jj, pair] == want
    # tensor-core LSTM fragments: [dir][rank][warp][mtile][kstep][lane][reg][2]
    fr = weights.a16_bits_to_f32(t["lstm.w_hh_mma"][1]).reshape(2, 2, 12, 2, 12, 32, 4, 2)
    rank, warp, mt, ks, lane, e = 1, 7, 1, 5, 22, 1
    gg, tig = lane // 4, lane % 4
    unit = rank * 96 + warp * 8 + gg
    bf = lambda v: weights.a16_bits_to_f32(weights.f32_to_a16_bits(np.asarray([v], np.float32)))[0]
    assert fr[1, rank, warp, mt, ks, lane, 0, e] == bf(src[2 * 192 + unit, ks * 16 + 2 * tig + e])        # gate g, row g
    assert fr[1, rank, warp, mt, ks, lane, 1, e] == bf(src[3 * 192 + unit, ks * 16 + 2 * tig + e])        # gate o, row g+8
    assert fr[1, rank, warp, mt, ks, lane, 3, e] == bf(src[3 * 192 + unit, ks * 16 + 2 * tig + 8 + e])
    blob = weights.pack_blob(sd)
    assert blob[:8] == b"KOCRW001" and len(blob) % 256 == 0
    vt = weights.pack_tensors(seeded_state_dict("vgg", 4))
    assert "lstm.w_ih" not in vt and "se3.w0p" not in vt and vt["meta"][1][0] == 1


def test_scheduling_chunk_counts_match_reference_rule():
    # W=400 -> 5, 800 -> 10, 1600 -> 20, 2400 -> 29, 100 -> 2
    for w, n in [(400, 5), (800, 10), (1600, 20), (2400, 29), (100, 2), (84, 1), (85, 2), (50, 1)]:
        assert scheduling.chunks_for(48, w) == n
    assert scheduling.resized_width(30, 375) == 600 and scheduling.resized_width(20, 10) == 50
    assert scheduling.chunks_for(48, 48 * 300, max_seq_len=4096) == 128       # truncation at 4096 tokens


def test_plan_batches_and_sharding():
    rng = np.random.default_rng(0)
    shapes = [(int(rng.integers(20, 70)), int(rng.integers(100, 2000))) for _ in range(500)]
    batches = scheduling.plan_batches(shapes, max_lines=64, max_chunks=300)
    assert sorted(i for b in batches for i in b) == list(range(500))
    for b in batches:
        assert len(b) <= 64 and sum(scheduling.chunks_for(*shapes[i]) for i in b) <= 300
    for ws in (1, 2, 4, 8):
        shards = scheduling.shard_lines(shapes, ws)
        assert sorted(i for s in shards for i in s) == list(range(500))
        loads = [sum(scheduling.chunks_for(*shapes[i]) for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(scheduling.chunks_for(*s) for s in shapes)
    assert scheduling.plan_batches([], 8, 8) == []


def test_synth_lines_are_deterministic():
    a, la = synth.make_lines(4, 400, 800, seed=0)
    b, lb = synth.make_lines(4, 400, 800, seed=0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and all(np.array_equal(x, y) for x, y in zip(la, lb))
    for im in a:
        assert im.dtype == np.uint8 and im.ndim == 2 and (im == 255).mean() > 0.5


def test_c_abi_exports_every_declared_symbol():
    header = (REPO / "include" / "kocr.h").read_text()
    declared = set(re.findall(r"\b(kocr_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kocr_abi_version() == 2


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    blob = weights.pack_blob(seeded_state_dict("vgg", 0, max_global_len=64))
    with pytest.raises(_native.KocrError, match="no CUDA device"):
        _native.Recognizer(blob, max_lines=2, max_chunks=8)
    from khmer_ocr_cnn_transformer_b200.recognition.predictor import OCRPredictor
    from khmer_ocr_cnn_transformer_b200.recognition.model.se_model import KhmerOCR
    with pytest.raises(RuntimeError, match="no CPU path"):
        OCRPredictor("x.pth", Tokenizer(REC / "char2idx.json"), OCRConfig(device="cpu"), KhmerOCR)


def test_resnet_variant_spec_and_packing():
    """Third checkpoint family (model/resnet_model.py): key layout, variant detection, blob entries, BN folding of the
    bias-free convs."""
    from khmer_ocr_cnn_transformer_b200.checkpoint import detect_variant, RESNET_BLOCKS
    sd = seeded_state_dict("resnet", 13)
    assert detect_variant(sd) == "resnet" and validate_state_dict(sd)[0] == "resnet"
    assert "cnn.layer2.1.shortcut.0.weight" not in sd and sd["cnn.layer2.0.shortcut.0.weight"].shape == (256, 128, 1, 1)
    assert "context_bilstm.weight_ih_l0" not in sd and "cnn.conv7.weight" not in sd
    t = weights.pack_tensors(sd)
    assert t["meta"][1][0] == 2 and "lstm.w_ih" not in t and "conv2.w" not in t
    for bi, (name, cin, cout) in enumerate(RESNET_BLOCKS):
        assert t[f"res{bi}.c1.w"][1].shape == (cout, 9 * cin) and t[f"res{bi}.c2.w"][1].shape == (cout, 9 * cout)
        assert (f"res{bi}.sc.w" in t) == (cin != cout)
    # folded bias of a bias-free conv = beta - mean * gamma / sqrt(var + eps)
    p = "cnn.layer3.0.bn2"
    want = sd[p + ".bias"] - sd[p + ".running_mean"] * sd[p + ".weight"] / np.sqrt(sd[p + ".running_var"] + 1e-5)
    assert np.allclose(t["res3.c2.b"][1], want, atol=1e-6)


def test_textline_box_arithmetic_matches_reference_rules():
    """textline_detection.py:17-34 / ocr_engine.py:72-76 in Python ints (host side of kocr_crop_lines)."""
    from khmer_ocr_cnn_transformer_b200 import textline_crops as T

    class Box:
        def __init__(self, poly):
            self.polygon = poly

    class Pred:
        bboxes = [Box([[10.9, 20.2], [99.9, 20.2], [99.9, 40.7], [10.9, 40.7]]), Box([[-3.0, 5.0], [4.0, 5.0], [4.0, 9.0], [-3.0, 9.0]]),
                  Box([[300.0, 10.0], [310.0, 10.0], [310.0, 20.0], [300.0, 20.0]])]
    assert T.textline_boxes((200, 100), Pred(), 5) == [(5, 15, 104, 45), (0, 0, 9, 14)]       # third box is outside: skipped
    assert T.textline_boxes((200, 100), [p.polygon for p in Pred.bboxes], 0) == [(10, 20, 99, 40), (0, 5, 4, 9)]
    assert T.element_boxes((200, 100), [((50, 60, 150, 90), 3), ((0, 0, 20, 10), 3)], 8) == [(0, 0, 28, 18), (42, 52, 158, 98)]


def test_batched_beam_bookkeeping_equals_reference_loop():
    """recognition/beam.BatchedBeam against a line-by-line transcription of the reference loop (predictor.py:104-136) on a
    deterministic fake decoder (logits = function of the prefix), incl. score ties, early <eos>, lines that finish at
    different positions and lines that never emit <eos>."""
    import torch
    import torch.nn.functional as F
    from khmer_ocr_cnn_transformer_b200.recognition.beam import BatchedBeam
    SOS, EOS, V, MAXLEN = 2, 3, 124, 24

    def fake_logits(line, prefix):
        rng = np.random.default_rng(abs(hash((line, tuple(int(p) for p in prefix)))) % (2 ** 32))
        lg = np.round(rng.standard_normal(V) * 2.0, 1).astype(np.float32)        # coarse values -> frequent exact ties
        if line % 4 != 3:                                                        # every 4th line never ends
            lg[EOS] += 0.35 * len(prefix) - 2.0
        else:
            lg[EOS] = -30.0
        return lg

    def reference_loop(line, bw):
        beams, completed = [(0.0, [SOS])], []
        for _ in range(MAXLEN):
            lp = F.log_softmax(torch.from_numpy(np.stack([fake_logits(line, s) for _, s in beams])), dim=-1)
            cands = []
            for i, (score, seq) in enumerate(beams):
                tp, ti = lp[i].topk(bw)
                for k in range(bw):
                    cands.append((score + tp[k].item(), seq + [ti[k].item()]))
            cands.sort(key=lambda x: x[0], reverse=True)
            nxt = []
            for sc, seq in cands:
                if seq[-1] == EOS:
                    completed.append((sc / len(seq), seq))
                elif len(nxt) < bw:
                    nxt.append((sc, seq))
            beams = nxt
            if not beams:
                break
        return sorted(completed, key=lambda x: x[0], reverse=True)[0][1] if completed else beams[0][1]

    for bw in (1, 2, 3, 5):
        n = 13
        beam = BatchedBeam(n, bw, SOS, EOS, MAXLEN)
        for t in range(MAXLEN):
            if beam.live_lines().size == 0:
                break
            row_line, prefixes, parents = beam.rows()
            assert prefixes.shape == (len(row_line), t + 1)
            if t == 0:
                assert list(row_line) == list(range(n)) and np.all(prefixes[:, 0] == SOS)
            lg = np.stack([fake_logits(int(l), p) for l, p in zip(row_line, prefixes)])
            tv, ti = F.log_softmax(torch.from_numpy(lg), dim=-1).topk(bw, dim=-1)
            beam.update(tv.numpy(), ti.numpy())
        got = beam.results()
        for line in range(n):
            assert got[line] == reference_loop(line, bw), (bw, line)


def test_top_level_recognition_import_like_the_reference():
    """netra_ocr/ocr_engine.py:6-10 puts ITS package directory on sys.path and imports `recognition.recognize_text` as a
    top-level package; INTEGRATION.md option A points that path at this repository's package directory."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from recognition.recognize_text import recognize_batch, recognize\n"
        "from recognition.predictor import OCRPredictor\n"
        "from recognition.tokenizer import Tokenizer\n"
        "from recognition.config import OCRConfig\n"
        "from recognition.utils import autodetect_config\n"
        "assert recognize_batch([]) == []\n"
        "assert OCRPredictor.__module__ == 'recognition.predictor'\n"
        "print('ok')\n" % str(REPO / "khmer_ocr_cnn_transformer_b200"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


@pytest.mark.parametrize("wrapped", [False, True])
def test_pth_checkpoint_round_trip(tmp_path, wrapped):
    """A reference-style `.pth` (torch.save of a state_dict with `num_batches_tracked` entries, bare or wrapped in
    {'model_state_dict': ...}: predictor.py:38-46) loads to the same arrays, the same config and the same packed blob."""
    import torch
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    from khmer_ocr_cnn_transformer_b200.recognition.utils import autodetect_config
    sd = seeded_state_dict("se", seed=3, max_global_len=512)
    state = {}
    for k, v in sd.items():
        state[k] = torch.from_numpy(v.copy())
        if k.endswith("running_var"):
            state[k.replace("running_var", "num_batches_tracked")] = torch.tensor(7, dtype=torch.int64)
    path = tmp_path / "khmerocr_se_transformer.pth"
    torch.save({"model_state_dict": state, "epoch": 3} if wrapped else state, path)
    got = load_checkpoint(path)
    assert set(got) == set(sd)
    assert all(np.array_equal(got[k], sd[k]) for k in sd)
    assert autodetect_config(path) == {"max_seq_len": 512, "emb_dim": 384, "decode_max_len": 256}
    assert weights.pack_blob(got) == weights.pack_blob(sd)


def test_build_skips_missing_dependencies(monkeypatch, tmp_path):
    from khmer_ocr_cnn_transformer_b200 import build as B
    if not B.LIB.exists():
        pytest.skip("library not built")
    monkeypatch.setattr(B, "PKG", B.PKG)        # needs_build() must not raise when ../include/kocr.h is absent
    real = B.PKG.parent / "include" / "kocr.h"
    assert real.exists()
    assert isinstance(B.needs_build(), bool)


def test_se_fragments_layout():
    """weights.se_fragments: B operand of mma.sync.m16n8k16 in fragment order (lane (g, t) of n-tile nt, k-step ks holds
    w[nt*8+g][ks*16 + 2t, +1] and [ks*16 + 8 + 2t, +1]; k-steps interleaved in pairs per lane)."""
    from khmer_ocr_cnn_transformer_b200.weights import se_fragments
    rng = np.random.default_rng(1)
    for N, K in ((32, 512), (16, 256), (512, 32), (256, 16)):
        w = rng.standard_normal((N, K)).astype(np.float32)
        f = se_fragments(w)
        assert f.shape == (N * K,) and np.array_equal(np.sort(f), np.sort(w.reshape(-1)))      # a permutation
        pair = K // 16 >= 2
        f = f.reshape(N // 8, K // 32, 32, 2, 4) if pair else f.reshape(N // 8, 1, 32, 1, 4)
        for nt in (0, N // 8 - 1):
            for lane in (0, 7, 18, 31):
                g, t = lane >> 2, lane & 3
                for ks in range(K // 16):
                    got = f[nt, ks // 2, lane, ks % 2] if pair else f[nt, 0, lane, 0]
                    k0 = ks * 16
                    want = [w[nt * 8 + g, k0 + 2 * t], w[nt * 8 + g, k0 + 2 * t + 1], w[nt * 8 + g, k0 + 8 + 2 * t],
                            w[nt * 8 + g, k0 + 9 + 2 * t]]
                    assert np.array_equal(got, want)
