"""ImagePreprocessor - host side of stage 1 (reference: netra_ocr/recognition/preprocessor.py:8-58).

The host only opens the image and converts it to 8-bit grey (`convert('L')`, exactly what the
reference does at :39-41).  Resize, chunking, white padding and normalisation run on the GPU
(csrc/preprocess.cu) and are bit-exact with the reference's Pillow/torchvision pipeline."""
from pathlib import Path

import numpy as np

from .config import OCRConfig
from . import _core
LineBatch = _core.native.LineBatch


class ImagePreprocessor:
    def __init__(self, config: OCRConfig, recognizer=None):
        self.cfg = config
        self._rec = recognizer
        if (config.img_height, config.chunk_width, config.chunk_overlap) != (48, 100, 16):
            raise ValueError("the CUDA path is specialised for img_height=48, chunk_width=100, chunk_overlap=16")

    @staticmethod
    def to_gray(image_source) -> np.ndarray:
        """path / PIL image / 2-D uint8 array -> (h, w) uint8 grey; same error behaviour as the reference."""
        if isinstance(image_source, np.ndarray):
            if image_source.ndim == 2 and image_source.dtype == np.uint8:
                return np.ascontiguousarray(image_source)
            raise ValueError("Input must be a path or PIL Image")
        from PIL import Image
        if isinstance(image_source, (str, Path)):
            if not Path(image_source).exists():
                raise FileNotFoundError(f"Image not found: {image_source}")
            image = Image.open(image_source).convert("L")
        elif isinstance(image_source, Image.Image):
            image = image_source.convert("L")
        else:
            raise ValueError("Input must be a path or PIL Image")
        return np.asarray(image, dtype=np.uint8)

    def process(self, image_source):
        """Return the normalised chunk tensor (n, 1, 48, 100) fp32, computed on the GPU."""
        import torch
        if self._rec is None:
            raise RuntimeError("ImagePreprocessor.process needs the CUDA recogniser (no CPU path)")
        batch = LineBatch([self.to_gray(image_source)])
        n = int(self._rec.gather_chunks(batch)[0])
        chunks = self._rec.debug_read("chunks").reshape(n, 1, 48, 100)
        return torch.from_numpy(chunks.copy())
