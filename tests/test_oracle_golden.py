"""The oracle (oracle/recognizer_np.py) against the committed outputs of the REFERENCE ITSELF
(tests/golden/*.npz, written by tests/golden/make_fixtures.py from /root/reference), and - when the
reference is importable (build container only) - against the live reference."""
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, rel_err, max_err, load_fixture_ckpt
from oracle import recognizer_np as O

REF = Path("/root/reference")


def _need(name):
    p = GOLDEN / name
    if not p.exists():
        pytest.skip(f"{name} not generated yet")
    return np.load(p)


def test_preprocess_bit_exact_against_reference_outputs():
    z = _need("golden_preprocess.npz")
    n = len([k for k in z.files if k.startswith("img")])
    assert n >= 10
    for i in range(n):
        got = O.preprocess_gray(z[f"img{i}"])[1]
        want = z[f"chunks{i}"]
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"case {i}"
    assert np.array_equal(O.rgb_to_l(z["rgb"]), z["rgb_l"])


def test_chunk_count_and_white_padding_rules():
    # SURVEY.md §0 table + notebook cell 13 (last column of the last chunk is white)
    for w, n in [(400, 5), (800, 10), (1600, 20), (2400, 29), (100, 2)]:
        assert O.n_chunks_for_width(w) == n
    line = np.zeros((48, 130), np.uint8)
    ch = O.chunk_resized_line(line)
    assert ch.shape == (2, 1, 48, 100)
    assert np.all(ch[1, 0, :, 46:] == 1.0) and np.all(ch[1, 0, :, :46] == -1.0)
    assert np.all(ch[0] == -1.0)


def test_se_model_stages_against_reference_outputs():
    z = _need("golden_se.npz")
    sd = load_fixture_ckpt()
    n = int(z["n_lines"])
    for i in range(n):
        chunks = O.preprocess_gray(z[f"img{i}"])[1]
        f = O.cnn_forward(sd, chunks, "se")
        assert rel_err(f[:2], z[f"cnn{i}"]) < 1e-4
        enc = O.encoder_forward(sd, O.patch_forward(sd, f))
        assert rel_err(enc, z[f"enc{i}"]) < 1e-4
        mem = O.memory_for_line(sd, enc, "se")
        assert rel_err(mem, z[f"mem{i}"]) < 1e-4
        if i < 3:       # the full-prefix greedy loop is slow in numpy: three lines are enough
            toks, logits = O.greedy_decode(sd, mem, return_logits=True)
            assert toks == [int(t) for t in z[f"tokens{i}"]]
            assert max_err(logits, z[f"step_logits{i}"]) < 1e-3


def test_beam_search_against_reference_outputs():
    """oracle.beam_search == OCRPredictor._beam_search (predictor.py:101-136) on the reference's own memory."""
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import build_vocab
    z = _need("golden_se.npz")
    if "beam3_texts" not in z.files:
        pytest.skip("golden_se.npz predates the beam-search goldens")
    idx2char = {v: k for k, v in build_vocab().items()}
    sd = load_fixture_ckpt()
    for i in (0, 1, 6):
        assert O.tokens_to_text(O.beam_search(sd, z[f"mem{i}"], 3), idx2char) == str(z["beam3_texts"][i])
    assert O.tokens_to_text(O.beam_search(sd, z["mem2"], 2), idx2char) == str(z["beam2_texts"][2])
    # width 1 degenerates to the greedy loop
    assert O.beam_search(sd, z["mem3"], 1)[:-1] == [int(t) for t in z["tokens3"]] or \
        O.beam_search(sd, z["mem3"], 1) == [int(t) for t in z["tokens3"]]


def test_vgg_baseline_against_reference_outputs():
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    z = _need("golden_vgg.npz")
    sd = seeded_state_dict("vgg", seed=11, max_global_len=1024)
    for i in range(int(z["n_lines"])):
        chunks = O.preprocess_gray(z[f"img{i}"])[1]
        enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "vgg")))
        assert rel_err(enc, z[f"enc{i}"]) < 1e-4
        mem = O.memory_for_line(sd, enc, "vgg")
        assert rel_err(mem, z[f"mem{i}"]) < 1e-4
        toks = O.greedy_decode(sd, mem, max_len=24)
        want = [int(t) for t in z[f"tokens{i}"]]
        assert toks[:len(want)] == want[:len(toks)]


def test_tokens_to_text_and_cer():
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import build_vocab
    idx2char = {v: k for k, v in build_vocab().items()}
    assert O.tokens_to_text([2, 42, 0, 43, 3, 44], idx2char) == "កខ"
    assert O.cer("abc", "abc") == 0.0 and abs(O.cer("abd", "abc") - 1 / 3) < 1e-9 and O.cer("", "") == 0.0


@pytest.mark.skipif(not REF.exists(), reason="reference checkout only exists in the build container")
def test_oracle_against_live_reference_random_weights():
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, str(REF))
    import torch
    from PIL import Image
    from netra_ocr.recognition.model.se_model import KhmerOCR
    from netra_ocr.recognition.preprocessor import ImagePreprocessor
    from netra_ocr.recognition.config import OCRConfig
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    rng = np.random.default_rng(3)
    pre = ImagePreprocessor(OCRConfig(device="cpu"))
    for (h, w) in [(30, 375), (61, 333), (20, 40), (96, 1000)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = pre.process(Image.fromarray(img)).numpy()
        assert np.array_equal(ref.view(np.uint32), O.preprocess_gray(img)[1].view(np.uint32))
    sd = seeded_state_dict("se", 5, max_global_len=256)
    m = KhmerOCR(vocab_size=124, pad_idx=0, emb_dim=384, max_global_len=256)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m.eval()
    chunks = O.preprocess_gray(rng.integers(0, 256, (40, 250), dtype=np.uint8))[1]
    with torch.no_grad():
        f = m.cnn(torch.from_numpy(chunks))
        p, _ = m.patch(f)
        e = m.enc(p.transpose(0, 1).contiguous()).transpose(0, 1)
        mem, _ = m.context_bilstm(e.reshape(1, -1, 384) + m.global_pos[: e.shape[0] * 32].unsqueeze(0))
        toks = [2, 5, 9, 0, 44, 17]
        lg = m.dec(torch.tensor([toks]), mem, torch.zeros(1, mem.shape[1], dtype=torch.bool))
    enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "se")))
    assert rel_err(enc, e.numpy()) < 1e-5
    mem2 = O.memory_for_line(sd, enc, "se")
    assert rel_err(mem2, mem[0].numpy()) < 1e-5
    assert max_err(O.decoder_forward(sd, toks, mem2), lg[0].numpy()) < 1e-4


def test_textline_crops_against_pillow_outputs():
    """oracle.textline_boxes / crop_line_gray == the Pillow calls of extract_textline_crops (textline_detection.py:17-47)
    + convert('L') (preprocessor.py:41), bit for bit, incl. clipping at the page edges and skipped boxes."""
    z = _need("golden_crops.npz")
    page, polys = z["page"], z["polys"].tolist()
    for tag in "abc":
        expansion, padding = [int(v) for v in z[f"{tag}_params"]]
        boxes = O.textline_boxes((page.shape[1], page.shape[0]), polys, expansion)
        assert len(boxes) == int(z[f"{tag}_n"]) == 9                   # 10 polygons, one entirely outside the page
        for i, b in enumerate(boxes):
            assert list(b) == [int(v) for v in z[f"{tag}_box{i}"]]
            assert np.array_equal(O.crop_line_gray(page, b, padding), z[f"{tag}_crop{i}"]), (tag, i)
    assert np.array_equal(O.rgb_to_l(page), z["page_l"])
    assert np.array_equal(O.crop_line_gray(z["page_l"], (10, 5, 300, 60), 0), z["l_crop"])


def test_teacher_forced_forward_against_reference_outputs():
    """oracle.forward_teacher_forced == the reference's KhmerOCR.forward (se_model.py:240-289): padded memory, BiLSTM
    over the pads, memory_key_padding_mask, <pad> target keys masked."""
    z = _need("golden_forward.npz")
    sd = load_fixture_ckpt()
    n = int(z["n"])
    enc = []
    for i in range(n):
        chunks = O.preprocess_gray(z[f"img{i}"])[1]
        enc.append(O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "se"))))
    got = O.forward_teacher_forced(sd, enc, z["tgt"], "se")
    assert got.shape == z["logits"].shape
    assert rel_err(got, z["logits"]) < 1e-4 and max_err(got, z["logits"]) < 2e-3


def test_resnet_baseline_against_reference_outputs():
    """oracle 'resnet' variant == the reference's ResNet-Transformer (model/resnet_model.py) with the seeded init."""
    from khmer_ocr_cnn_transformer_b200.checkpoint import seeded_state_dict
    z = _need("golden_resnet.npz")
    sd = seeded_state_dict("resnet", seed=13, max_global_len=1024)
    for i in range(int(z["n_lines"])):
        chunks = O.preprocess_gray(z[f"img{i}"])[1]
        f = O.cnn_forward(sd, chunks, "resnet")
        assert rel_err(f, z[f"cnn{i}"]) < 1e-4
        enc = O.encoder_forward(sd, O.patch_forward(sd, f))
        assert rel_err(enc, z[f"enc{i}"]) < 1e-4
        mem = O.memory_for_line(sd, enc, "resnet")
        assert rel_err(mem, z[f"mem{i}"]) < 1e-4
        toks = O.greedy_decode(sd, mem, max_len=24)
        want = [int(t) for t in z[f"tokens{i}"]]
        assert toks[:len(want)] == want[:len(toks)]


def test_vgg_trained_fixture_against_reference_outputs():
    """oracle 'vgg' variant on the TRAINED VGG fixture == the reference's VGG class + OCRPredictor (memory, greedy tokens)."""
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    z = _need("golden_vgg_trained.npz")
    sd = load_checkpoint(GOLDEN / "fixture_vgg_ckpt.npz")
    for i in (0, 3, 8):
        chunks = O.preprocess_gray(z[f"img{i}"])[1]
        enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "vgg")))
        mem = O.memory_for_line(sd, enc, "vgg")
        assert rel_err(mem, z[f"mem{i}"]) < 1e-4
        assert O.greedy_decode(sd, mem) == [int(t) for t in z[f"tokens{i}"]]


def test_stored_oracle_tokens_match_live_oracle():
    """tests/golden/oracle_tokens_*.npz (what tests/test_gpu_token_identity.py compares the CUDA path with on the full c2 / c3
    workloads) are outputs of THIS oracle: re-derive a few lines, including one beyond the first 1024 of the 8192-line set."""
    from workloads import synth
    sd = load_fixture_ckpt()
    cases = {"c2": ((256, 400, 800, 0), (0, 131)), "c3": ((1024, 200, 1600, 3), (7,)), "c3full": ((8192, 200, 1600, 3), (4099,))}
    for name, (spec, picks) in cases.items():
        z = _need(f"oracle_tokens_{name}.npz")
        n, lo, hi, seed = spec
        imgs, _ = synth.make_lines(max(picks) + 1, lo, hi, seed=seed)      # the generator is sequential: a prefix is enough
        for i in picks:
            want = [int(t) for t in z["tokens"][i, :z["lengths"][i]]]
            assert O.recognise_lines(sd, [imgs[i]], "se")[0] == want, (name, i)
