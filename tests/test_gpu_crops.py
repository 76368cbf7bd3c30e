"""Input side: page -> text-line crops on the GPU (kocr_crop_lines), bit-exact with the Pillow pipeline
of the reference (textline_detection.py:7-53, ocr_engine.py:72-76, preprocessor.py:39-41)."""
import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rec():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    r = _native.Recognizer(weights.pack_blob(load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")), max_lines=32, max_chunks=512)
    yield r
    r.close()


def test_crops_bit_exact_against_pillow_goldens(rec):
    from khmer_ocr_cnn_transformer_b200 import textline_crops as T
    z = np.load(GOLDEN / "golden_crops.npz")
    page, polys = z["page"], z["polys"].tolist()
    for tag in "abc":
        expansion, padding = [int(v) for v in z[f"{tag}_params"]]
        boxes = T.textline_boxes((page.shape[1], page.shape[0]), polys, expansion)
        assert [list(b) for b in boxes] == [[int(v) for v in z[f"{tag}_box{i}"]] for i in range(int(z[f"{tag}_n"]))]
        got = T.crop_lines_device(rec, page, boxes, padding).to_host()
        for i, g in enumerate(got):
            assert np.array_equal(g, z[f"{tag}_crop{i}"]), (tag, i)
    # grey page in, no canvas (the custom-detector branch, ocr_engine.py:72-76)
    got = T.crop_lines_device(rec, z["page_l"], [(10, 5, 300, 60)], 0).to_host()
    assert np.array_equal(got[0], z["l_crop"])
    assert T.element_boxes((900, 420), [((5, 200, 100, 230), 1), ((880, 10, 899, 40), 1)], 4) == \
        [(876, 6, 900, 44), (1, 196, 104, 234)]


def test_crop_errors_and_empty(rec):
    from khmer_ocr_cnn_transformer_b200 import _native, textline_crops as T
    page = np.full((50, 60, 3), 255, np.uint8)
    assert T.crop_lines_device(rec, page, [], 10).to_host() == []
    with pytest.raises(_native.KocrError, match="outside"):
        T.crop_lines_device(rec, page, [(0, 0, 61, 10)], 0)
    with pytest.raises(_native.KocrError, match="empty"):
        T.crop_lines_device(rec, page, [(5, 5, 5, 10)], 0)


def test_predict_page_equals_recognition_of_the_pillow_crops():
    """Page + polygons -> texts: crops cut on the GPU and recognised from device memory give exactly the texts of
    recognising the Pillow-made crops one by one (the reference's extract_textline_crops -> recognize_batch sequence)."""
    from PIL import Image
    from khmer_ocr_cnn_transformer_b200.recognition.config import OCRConfig
    from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import Tokenizer
    from khmer_ocr_cnn_transformer_b200.recognition.predictor import OCRPredictor
    from khmer_ocr_cnn_transformer_b200.recognition.utils import autodetect_config
    from khmer_ocr_cnn_transformer_b200.recognition.model.se_model import KhmerOCR
    from khmer_ocr_cnn_transformer_b200.recognition import recognize_text
    z = np.load(GOLDEN / "golden_crops.npz")
    ckpt = GOLDEN / "fixture_se_ckpt.npz"
    pred = OCRPredictor(ckpt, Tokenizer(recognize_text.DEFAULT_VOCAB_PATH), OCRConfig(**autodetect_config(ckpt)), KhmerOCR,
                        max_lines=16, max_chunks=256)
    try:
        got = pred.predict_page(Image.fromarray(z["page"]), z["polys"].tolist(), expansion_px=5, padding_px=10)
        want = pred.predict_batch([Image.fromarray(z[f"a_crop{i}"]) for i in range(int(z["a_n"]))], beam_width=1)
        assert got == want and len(got) == 9
        assert sum(len(t) > 3 for t in got) >= 7          # the seven synthetic lines are actually read
    finally:
        pred.close()
