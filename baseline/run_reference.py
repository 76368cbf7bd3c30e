"""The UNMODIFIED reference recogniser as a comparison arm (bench.py `--impl reference`, `cpu_baseline`).

`baseline/_ref/` holds the reference installed with
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
(git-ignored, NOT gpurun-ignored: it travels to the GPU box; `__graft_entry__.build()` re-creates it when /root/reference
is present).  Nothing here is on the product path, and nothing of this repository's kernels / engine is on the
reference's path: the reference's own `OCRPredictor.predict_batch` (netra_ocr/recognition/predictor.py:138-199) runs its
own torch modules, on the device named in its own OCRConfig.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"


def available() -> tuple[bool, str]:
    if not (REF_DIR / "netra_ocr" / "recognition" / "predictor.py").exists():
        return False, "baseline/_ref/netra_ocr/recognition is absent (run __graft_entry__.build() where /root/reference exists)"
    try:
        import torchvision  # noqa: F401  (the reference's preprocessor needs it)
    except Exception as e:      # pragma: no cover
        return False, f"torchvision missing: {e}"
    return True, ""


def save_pth(state_dict_np: dict, path: Path) -> Path:
    """Reference-style checkpoint: torch.save of a bare state_dict incl. the BatchNorm `num_batches_tracked` entries."""
    import torch
    state = {}
    for k, v in state_dict_np.items():
        state[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
        if k.endswith("running_var"):
            state[k[:-len("running_var")] + "num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    torch.save(state, path)
    return path


def load_predictor(state_dict_np: dict, device: str = "cpu", variant: str = "se"):
    """The reference's own OCRPredictor, built the way its recognize_text._get_predictor does (recognize_text.py:29-59)."""
    import torch
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    from netra_ocr.recognition.config import OCRConfig
    from netra_ocr.recognition.tokenizer import Tokenizer
    from netra_ocr.recognition.predictor import OCRPredictor
    from netra_ocr.recognition.utils import autodetect_config
    if variant == "vgg":
        from netra_ocr.recognition.model.vgg_model import KhmerOCR
    elif variant == "resnet":
        from netra_ocr.recognition.model.resnet_model import KhmerOCR
    else:
        from netra_ocr.recognition.model.se_model import KhmerOCR
    tmp = Path(tempfile.mkdtemp(prefix="kocr_ref_")) / f"khmerocr_{variant}_transformer.pth"
    save_pth(state_dict_np, tmp)
    cfg = OCRConfig(**autodetect_config(tmp))
    cfg.device = device
    tok = Tokenizer(REF_DIR / "netra_ocr" / "recognition" / "char2idx.json")
    pred = OCRPredictor(model_path=tmp, tokenizer=tok, config=cfg, model_class=KhmerOCR)
    try:
        os.remove(tmp)
    except OSError:
        pass
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: give the arm every host core
    return pred


def to_pil(images):
    from PIL import Image
    return [Image.fromarray(im) for im in images]


def time_predict_batch(pred, images, batch_size: int = 8):
    """lines/s of `OCRPredictor.predict_batch(images, beam_width=1, batch_size=8)`; returns (lines_per_s, seconds, texts)."""
    import contextlib
    import io
    pil = to_pil(images)
    t0 = time.perf_counter()
    with contextlib.redirect_stderr(io.StringIO()):        # the reference draws a tqdm bar on stderr
        texts = pred.predict_batch(pil, beam_width=1, batch_size=batch_size)
    if pred.device.type == "cuda":
        import torch
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return len(images) / dt, dt, texts
