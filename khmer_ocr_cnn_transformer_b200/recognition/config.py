"""`OCRConfig` - same seven fields and defaults as the reference dataclass
(reference: netra_ocr/recognition/config.py:4-13).  `device` defaults to CUDA because this
package has no CPU path; asking for anything else fails loudly in `OCRPredictor`."""
from dataclasses import dataclass


@dataclass
class OCRConfig:
    """Configuration for OCR Inference Pipeline."""
    img_height: int = 48
    chunk_width: int = 100
    chunk_overlap: int = 16
    emb_dim: int = 384
    max_seq_len: int = 4096
    decode_max_len: int = 256
    device: str = "cuda"
