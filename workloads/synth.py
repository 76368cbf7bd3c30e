"""Synthetic Khmer text-line images for tests and benchmarks (no fonts needed at run time).

The reference renders training lines with Pillow from a corpus and the bundled fonts
(scripts/generate_document_text.py:86-127: words joined by spaces, black on white, 5 px margin).
Neither the corpus nor (on the GPU box) the fonts are available, so `tests/golden/make_fixtures.py`
renders a bank of vocab-derived pseudo-words once, offline, with those fonts and commits it as
`workloads/wordbank.npz`.  This module composes lines from that bank with a seeded RNG:
same pixel statistics (anti-aliased black glyphs on white, native heights 20-60 px so that the
height-48 resize is exercised), unlimited supply, deterministic.
"""
from __future__ import annotations

from pathlib import Path
import numpy as np

DEFAULT_BANK = Path(__file__).resolve().parent / "wordbank.npz"


class WordBank:
    """Word images grouped by (font, size); every word of a group shares the canvas height."""

    def __init__(self, path=DEFAULT_BANK):
        z = np.load(path)
        self.pixels = z["pixels"]              # uint8 flat
        self.offsets = z["offsets"]            # int64 [n_words+1]
        self.widths = z["widths"]              # int32 [n_words]
        self.heights = z["heights"]            # int32 [n_words]  (== group height)
        self.group = z["group"]                # int32 [n_words]
        self.space = z["space"]                # int32 [n_groups]  space width in px
        self.tok_flat = z["tok_flat"]          # int32 flat token ids of all words
        self.tok_off = z["tok_off"]            # int64 [n_words+1]
        self.n_groups = int(self.space.shape[0])
        self.members = [np.nonzero(self.group == g)[0] for g in range(self.n_groups)]

    def word(self, i: int) -> np.ndarray:
        h, w = int(self.heights[i]), int(self.widths[i])
        return self.pixels[self.offsets[i]:self.offsets[i + 1]].reshape(h, w)

    def tokens(self, i: int) -> np.ndarray:
        return self.tok_flat[self.tok_off[i]:self.tok_off[i + 1]]


def compose_line(bank: WordBank, rng: np.random.Generator, target_width: int,
                 space_token: int = 4, margin: int = 5):
    """Compose one line whose width AFTER the height-48 resize is close to `target_width`.
    Returns (uint8 grey image (h, w), int32 label token ids without sos/eos)."""
    g = int(rng.integers(bank.n_groups))
    ids = bank.members[g]
    h = int(bank.heights[ids[0]]) + 2 * margin
    native_target = max(20, int(round(target_width * h / 48.0)))
    parts, label, w = [], [], 2 * margin
    sp = int(bank.space[g])
    while True:
        i = int(ids[int(rng.integers(len(ids)))])
        wi = int(bank.widths[i])
        add = wi + (sp if parts else 0)
        if parts and w + add > native_target:
            break
        if parts:
            label.append(space_token)
        parts.append(i)
        label.extend(int(t) for t in bank.tokens(i))
        w += add
        if w >= native_target:
            break
    img = np.full((h, w), 255, np.uint8)
    x = margin
    for k, i in enumerate(parts):
        if k:
            x += sp
        wi = int(bank.widths[i])
        img[margin:h - margin, x:x + wi] = bank.word(i)
        x += wi
    return img, np.asarray(label, np.int32)


def make_lines(n: int, width_lo: int, width_hi: int, seed: int = 0, bank: WordBank | None = None):
    """`n` seeded lines with resized widths ~ uniform in [width_lo, width_hi]."""
    bank = bank or WordBank()
    rng = np.random.Generator(np.random.PCG64(seed))
    imgs, labels = [], []
    for _ in range(n):
        tw = int(rng.integers(width_lo, width_hi + 1))
        im, lb = compose_line(bank, rng, tw)
        imgs.append(im)
        labels.append(lb)
    return imgs, labels
