"""Tokenizer - id <-> character mapping, byte-identical vocab layout to the reference
(reference: netra_ocr/recognition/tokenizer.py:4-38 and char2idx.json: 124 contiguous ids;
0 <pad>, 1 <unk>, 2 <sos>, 3 <eos>, then the characters in code-point order)."""
import json
from pathlib import Path

# The vocabulary as code-point ranges (inclusive), in id order after the four specials.
_SPECIALS = ["<pad>", "<unk>", "<sos>", "<eos>"]
_RANGES = [
    (0x20, 0x3A), (0x3C, 0x40), (0x5B, 0x5B), (0x5D, 0x5D), (0x5F, 0x5F), (0x7C, 0x7C),
    (0xAB, 0xAB), (0xBB, 0xBB),
    (0x1780, 0x179C), (0x179F, 0x17A2), (0x17A5, 0x17A5), (0x17A7, 0x17A7), (0x17AC, 0x17AC),
    (0x17AF, 0x17AF), (0x17B1, 0x17B2), (0x17B6, 0x17CD), (0x17CF, 0x17D0), (0x17D2, 0x17D2),
    (0x17D4, 0x17D7), (0x17E0, 0x17E9), (0x2039, 0x203A),
]


def build_vocab() -> dict:
    """char -> id for the 124-symbol vocabulary."""
    chars = list(_SPECIALS)
    for lo, hi in _RANGES:
        chars.extend(chr(c) for c in range(lo, hi + 1))
    return {c: i for i, c in enumerate(chars)}


class Tokenizer:
    """Handles mapping between characters and integer IDs."""

    def __init__(self, char2idx_path):
        self.char2idx_path = Path(char2idx_path)
        self.char2idx, self.idx2char = self._load_vocab()
        self.sos_idx = self.char2idx.get("<sos>", 1)
        self.eos_idx = self.char2idx.get("<eos>", 2)
        self.pad_idx = self.char2idx.get("<pad>", 0)

    def _load_vocab(self):
        if not self.char2idx_path.exists():
            raise FileNotFoundError(f"Vocab file not found: {self.char2idx_path}")
        with open(self.char2idx_path, "r", encoding="utf-8") as f:
            char2idx = json.load(f)
        idx2char = {v: k for k, v in char2idx.items()}
        return char2idx, idx2char

    def decode(self, token_ids) -> str:
        """Converts a list of token IDs back to a string: skip sos/pad, stop at eos."""
        result = []
        for idx in token_ids:
            if idx == self.sos_idx or idx == self.pad_idx:
                continue
            if idx == self.eos_idx:
                break
            result.append(self.idx2char.get(idx, ""))
        return "".join(result)

    def __len__(self):
        return len(self.char2idx)
