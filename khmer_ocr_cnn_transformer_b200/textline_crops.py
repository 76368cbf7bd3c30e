"""Input side of the recogniser: text-line crops of a page image.

Reference: `extract_textline_crops` (netra_ocr/textline_detection.py:7-53) expands every detected polygon's bounding
box by `expansion_px`, clips it to the page, crops, and pastes the crop on a white canvas with `padding_px` on every
side; the custom-detector branch of `OCREngine.process_image` (netra_ocr/ocr_engine.py:72-76) crops the clipped,
padded box without a canvas.  `recognize_batch` then converts every crop to 8-bit grey (preprocessor.py:39-41).

Here the box arithmetic stays in Python ints exactly as the reference writes it, and the pixels are cut, padded and
converted on the GPU (`kocr_crop_lines`, csrc/preprocess.cu) straight into the buffer that stage 1 reads - the page is
uploaded once, the crops never exist on the host.  Bit-exact with the Pillow pipeline (tests/test_gpu_crops.py)."""
from __future__ import annotations

import numpy as np

from ._native import line_batch_from_shapes


def _polygon_of(obj):
    return obj.polygon if hasattr(obj, "polygon") else obj


def textline_boxes(image_size, textline_pred, expansion_px: int = 5):
    """Boxes (x0, y0, x1, y1) of `extract_textline_crops` steps 1-2 (textline_detection.py:17-34): int() truncation of
    the polygon extremes, expansion, clipping to the page, empty boxes skipped.  `textline_pred` is a Surya-style
    prediction (`.bboxes[i].polygon`) or a plain list of polygons [[x, y], ...]."""
    img_w, img_h = image_size
    polys = textline_pred.bboxes if hasattr(textline_pred, "bboxes") else textline_pred
    boxes = []
    for obj in polys:
        poly = _polygon_of(obj)
        xs = [p[0] for p in poly]
        ys = [p[1] for p in poly]
        x0, y0 = int(min(xs)), int(min(ys))
        x1, y1 = int(max(xs)), int(max(ys))
        x0 = max(0, x0 - expansion_px)
        y0 = max(0, y0 - expansion_px)
        x1 = min(img_w, x1 + expansion_px)
        y1 = min(img_h, y1 + expansion_px)
        if x1 - x0 <= 0 or y1 - y0 <= 0:
            continue
        boxes.append((x0, y0, x1, y1))
    return boxes


def element_boxes(image_size, elements, padding: int):
    """Boxes of the custom-detector branch (ocr_engine.py:66-76): elements = [((x1, y1, x2, y2), class_id), ...] already
    filtered to text classes; sorted by y, padded by `padding`, clipped; cropped WITHOUT a white canvas."""
    img_w, img_h = image_size
    out = []
    for el in sorted(elements, key=lambda e: e[0][1]):
        x1, y1, x2, y2 = el[0]
        out.append((max(0, x1 - padding), max(0, y1 - padding), min(img_w, x2 + padding), min(img_h, y2 + padding)))
    return out


class DeviceCrops:
    """Grey line crops in device memory + the LineBatch table that kocr_gather_chunks / kocr_recognize_lines take."""

    def __init__(self, buffer, batch, boxes):
        self.buffer, self.batch, self.boxes = buffer, batch, boxes

    @property
    def dev_ptr(self) -> int:
        return self.buffer.data_ptr()

    def to_host(self):
        """list of (h, w) uint8 arrays (tests / debugging)."""
        flat = self.buffer.cpu().numpy()
        return [flat[o:o + h * w].reshape(h, w).copy()
                for o, h, w in zip(self.batch.offsets, self.batch.heights, self.batch.widths)]


def crop_lines_device(recognizer, image, boxes, padding_px: int = 10) -> DeviceCrops:
    """`image`: PIL image (RGB or L; other modes are converted to RGB first, like Pillow's paste onto an RGB canvas) or a
    uint8 array (H, W) / (H, W, 3).  Returns the padded grey crops of `boxes` in device memory."""
    import torch
    if not isinstance(image, np.ndarray):
        if image.mode not in ("RGB", "L"):
            image = image.convert("RGB")
        image = np.asarray(image, dtype=np.uint8)
    page = np.ascontiguousarray(image)
    shapes = [(y1 - y0 + 2 * padding_px, x1 - x0 + 2 * padding_px) for (x0, y0, x1, y1) in boxes]
    batch = line_batch_from_shapes(shapes)
    buf = torch.empty(max(batch.pixel_bytes, 1), dtype=torch.uint8, device=f"cuda:{recognizer.device}")
    if boxes:
        recognizer.crop_lines(page, np.asarray(boxes, np.int32), padding_px, buf.data_ptr(), batch.offsets)
    return DeviceCrops(buf, batch, list(boxes))
