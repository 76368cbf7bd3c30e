"""The reference's MODEL protocol on the handle (`cnn(chunks)`, `patch(f) -> (x, N)`, `enc(p)` seq-first, `global_pos`,
optional `context_bilstm`, `dec(tgt, memory, mask)`): every stage against the oracle, and the UNMODIFIED reference
`OCRPredictor` (baseline/_ref) driving THIS model object end to end."""
import sys

import numpy as np
import pytest

from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pred():
    from khmer_ocr_cnn_transformer_b200.recognition.config import OCRConfig
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import Tokenizer
    from khmer_ocr_cnn_transformer_b200.recognition.predictor import OCRPredictor
    from khmer_ocr_cnn_transformer_b200.recognition.utils import autodetect_config
    from khmer_ocr_cnn_transformer_b200.recognition.model.se_model import KhmerOCR
    from khmer_ocr_cnn_transformer_b200.recognition import recognize_text
    ckpt = GOLDEN / "fixture_se_ckpt.npz"
    p = OCRPredictor(ckpt, Tokenizer(recognize_text.DEFAULT_VOCAB_PATH), OCRConfig(**autodetect_config(ckpt)), KhmerOCR,
                     max_lines=64, max_chunks=512, in_flight=1)
    yield p
    p.close()


def test_model_protocol_stages_match_oracle(pred):
    import torch
    from helpers import load_fixture_ckpt
    from oracle import recognizer_np as O
    sd = load_fixture_ckpt()
    z = np.load(GOLDEN / "golden_se.npz")
    chunks = O.preprocess_gray(z["img3"])[1]                       # (n, 1, 48, 100)
    m = pred.model
    f = m.cnn(torch.from_numpy(chunks))
    want_f = O.cnn_forward(sd, chunks, "se")
    assert tuple(f.shape) == want_f.shape == (chunks.shape[0], 512, 2, 32)
    x, N = m.patch(torch.from_numpy(want_f))
    want_x = O.patch_forward(sd, want_f)
    assert N == 32 and tuple(x.shape) == want_x.shape
    enc = m.enc(torch.from_numpy(want_x).transpose(0, 1).contiguous()).transpose(0, 1)
    want_enc = O.encoder_forward(sd, want_x)
    merged = O.merge_line(sd, want_enc)                            # (T, 384) incl. global_pos
    mem, _ = m.context_bilstm(torch.from_numpy(merged)[None])
    want_mem = O.bilstm(sd, merged)
    toks = [int(t) for t in z["tokens3"]][:24]
    logits = m.dec(torch.tensor([toks]), torch.from_numpy(want_mem)[None], torch.zeros((1, want_mem.shape[0]), dtype=torch.bool))
    want_logits = O.decoder_forward(sd, toks, want_mem)
    errs = {"cnn": rel_err(f.numpy(), want_f), "patch": rel_err(x.numpy(), want_x), "enc": rel_err(enc.numpy(), want_enc),
            "context_bilstm": rel_err(mem[0].numpy(), want_mem), "dec": rel_err(logits[0].numpy(), want_logits)}
    from test_gpu_stages import _report
    _report("model_protocol_rel_err_vs_oracle", errs)
    assert max(errs.values()) < 3e-2, errs
    assert m.global_pos.shape == (sd["global_pos"].shape[0], 384) and hasattr(m, "context_bilstm")
    # memory_key_padding_mask: padded memory rows must not matter
    pad = np.concatenate([want_mem, np.full((7, 384), 9.0, np.float32)])[None]
    mask = torch.zeros((1, pad.shape[1]), dtype=torch.bool)
    mask[0, want_mem.shape[0]:] = True
    again = m.dec(torch.tensor([toks]), torch.from_numpy(pad), mask)
    assert np.array_equal(again.numpy(), logits.numpy())


def test_unmodified_reference_predictor_runs_on_this_model(pred):
    """predictor.py:48-99,138-199 of the REFERENCE (baseline/_ref), with `self.model` replaced by this repository's handle
    object: greedy texts of `predict` and `predict_batch` equal the reference's own outputs on the golden lines."""
    from pathlib import Path
    ref_dir = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref_dir / "netra_ocr" / "recognition" / "predictor.py").exists():
        pytest.skip("baseline/_ref not installed (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    from PIL import Image
    from baseline import run_reference as R
    from helpers import load_fixture_ckpt
    z = np.load(GOLDEN / "golden_se.npz")
    ref_pred = R.load_predictor(load_fixture_ckpt(), "cpu")        # the reference's own predictor object (torch model inside) ...
    ref_pred.model = pred.model                                     # ... now driving the CUDA path through the model protocol
    imgs = [Image.fromarray(z[f"img{i}"]) for i in range(4)]
    want = [str(t) for t in z["texts"][:4]]
    assert [ref_pred.predict(im, beam_width=1) for im in imgs[:2]] == want[:2]
    import contextlib, io
    with contextlib.redirect_stderr(io.StringIO()):
        assert ref_pred.predict_batch(imgs, beam_width=1, batch_size=3) == want
