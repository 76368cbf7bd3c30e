"""`OCRConfig` - same seven fields and defaults as the reference dataclass
(reference: netra_ocr/recognition/config.py:4-13), including `device`: "cuda" when a CUDA device is visible, else
"cpu".  This package has no CPU path, so a "cpu" config fails loudly in `OCRPredictor` instead of running slowly."""
from dataclasses import dataclass


def _default_device() -> str:
    try:
        import torch
        return "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        return "cpu"


@dataclass
class OCRConfig:
    """Configuration for OCR Inference Pipeline."""
    img_height: int = 48
    chunk_width: int = 100
    chunk_overlap: int = 16
    emb_dim: int = 384
    max_seq_len: int = 4096
    decode_max_len: int = 256
    device: str = _default_device()
