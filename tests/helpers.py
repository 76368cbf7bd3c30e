"""Shared test helpers (layout conversions, error metrics, fixture loading)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def bf16_u16_to_f32(a: np.ndarray) -> np.ndarray:
    """Raw 16-bit activation words of the library (fp16 by default, bf16 for the KOCR_A16_BF16 build) -> fp32."""
    from khmer_ocr_cnn_transformer_b200.weights import a16_bits_to_f32
    return a16_bits_to_f32(a)


def nwhc_to_nchw(buf_u16: np.ndarray, n: int, H: int, W: int, C: int) -> np.ndarray:
    """dense NWHC 16-bit activation of the library [(n*W + w)*H + h][C] -> fp32 (n, C, H, W)."""
    x = bf16_u16_to_f32(buf_u16).reshape(n, W, H, C)
    return np.ascontiguousarray(x.transpose(0, 3, 2, 1))


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2"""
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) /
                 (np.linalg.norm(b.astype(np.float64)) + 1e-30))


def max_err(a, b) -> float:
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def load_fixture_ckpt():
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    return load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")
