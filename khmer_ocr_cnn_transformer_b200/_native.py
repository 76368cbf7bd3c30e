"""ctypes binding of libkocr_b200.so (include/kocr.h).  No CPU fallback: a missing library or a
missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libkocr_b200.so"
TOKENS_LD = 257

EXPORTS = [
    "kocr_abi_version", "kocr_last_error", "kocr_create", "kocr_destroy", "kocr_workspace_bytes",
    "kocr_model_info", "kocr_gather_chunks", "kocr_sevgg_encoder_forward", "kocr_merge_bilstm_forward",
    "kocr_decode_greedy", "kocr_recognize_lines", "kocr_set_option", "kocr_set_forced_tokens",
    "kocr_debug_read", "kocr_launch_count", "kocr_test_gemm", "kocr_read_kernel_timing", "kocr_read_unfinished", "kocr_beam_step", "kocr_beam_step_batch", "kocr_crop_lines", "kocr_forward_teacher_forced", "kocr_beam_search",
    "kocr_model_cnn", "kocr_model_patch", "kocr_model_enc", "kocr_model_bilstm", "kocr_model_dec",
]

_lib = None


class KocrError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (building it with nvcc first if it is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        if _build.needs_build():
            _build.build()
    if not LIB_PATH.exists():
        raise KocrError(f"{LIB_PATH} is missing: run `python -m khmer_ocr_cnn_transformer_b200.build` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.kocr_abi_version.restype = i32
    lib.kocr_last_error.restype = C.c_char_p
    lib.kocr_create.argtypes = [vp, sz, i32, i32, i32, C.POINTER(vp)]
    lib.kocr_destroy.argtypes = [vp]
    lib.kocr_workspace_bytes.argtypes = [vp]
    lib.kocr_workspace_bytes.restype = sz
    lib.kocr_model_info.argtypes = [vp] + [C.POINTER(i32)] * 5
    lib.kocr_gather_chunks.argtypes = [vp, vp, sz, i32, vp, vp, vp, i32, vp, vp]
    lib.kocr_sevgg_encoder_forward.argtypes = [vp, vp]
    lib.kocr_merge_bilstm_forward.argtypes = [vp, vp]
    lib.kocr_decode_greedy.argtypes = [vp, i32, vp, vp, vp]
    lib.kocr_recognize_lines.argtypes = [vp, vp, sz, i32, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.kocr_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.kocr_set_forced_tokens.argtypes = [vp, vp, i32]
    lib.kocr_debug_read.argtypes = [vp, C.c_char_p, vp, sz, C.POINTER(sz)]
    lib.kocr_beam_step.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp]
    lib.kocr_beam_step.restype = i32
    lib.kocr_beam_step_batch.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp]
    lib.kocr_beam_step_batch.restype = i32
    for f in ("kocr_model_cnn", "kocr_model_patch", "kocr_model_enc"):
        getattr(lib, f).argtypes = [vp, vp, i32, vp, vp]
        getattr(lib, f).restype = i32
    lib.kocr_model_bilstm.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.kocr_model_bilstm.restype = i32
    lib.kocr_model_dec.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp, vp]
    lib.kocr_model_dec.restype = i32
    lib.kocr_beam_search.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.kocr_beam_search.restype = i32
    lib.kocr_crop_lines.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp]
    lib.kocr_crop_lines.restype = i32
    lib.kocr_forward_teacher_forced.argtypes = [vp, vp, i32, vp, vp]
    lib.kocr_forward_teacher_forced.restype = i32
    lib.kocr_read_unfinished.argtypes = [vp, vp]
    lib.kocr_read_unfinished.restype = i32
    lib.kocr_read_kernel_timing.argtypes = [vp, C.c_char_p, sz]
    lib.kocr_read_kernel_timing.restype = i32
    lib.kocr_launch_count.restype = i64
    lib.kocr_test_gemm.argtypes = [i32, vp, i64, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp]
    for f in ("kocr_create", "kocr_destroy", "kocr_model_info", "kocr_gather_chunks",
              "kocr_sevgg_encoder_forward", "kocr_merge_bilstm_forward", "kocr_decode_greedy",
              "kocr_recognize_lines", "kocr_set_option", "kocr_set_forced_tokens", "kocr_debug_read",
              "kocr_test_gemm"):
        getattr(lib, f).restype = i32
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load_library().kocr_last_error().decode("utf-8", "replace")
        raise KocrError(f"libkocr_b200 error {rc}: {msg}")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


class LineBatch:
    """Concatenated grey uint8 line images + the offset/height/width tables the C ABI takes."""

    def __init__(self, images):
        hs, ws, offs, total = [], [], [], 0
        for im in images:
            if im.ndim != 2 or im.dtype != np.uint8:
                raise ValueError("line images must be 2-D uint8 (grey) arrays")
            hs.append(im.shape[0]); ws.append(im.shape[1]); offs.append(total)
            total += im.size
        self.n = len(images)
        self.pixels = np.empty(max(total, 1), np.uint8)
        for im, o in zip(images, offs):
            self.pixels[o:o + im.size] = im.reshape(-1)
        self.pixel_bytes = total
        self.offsets = np.asarray(offs, np.int64)
        self.heights = np.asarray(hs, np.int32)
        self.widths = np.asarray(ws, np.int32)


def line_batch_from_shapes(shapes):
    """A LineBatch that only describes lines living in DEVICE memory (kocr_crop_lines output): shapes = [(h, w), ...]."""
    b = LineBatch.__new__(LineBatch)
    hs = [int(h) for h, _ in shapes]
    ws = [int(w) for _, w in shapes]
    sizes = [h * w for h, w in zip(hs, ws)]
    b.n = len(shapes)
    b.pixel_bytes = int(sum(sizes))
    b.pixels = np.empty(1, np.uint8)
    b.offsets = np.asarray(np.cumsum([0] + sizes[:-1]) if sizes else [], np.int64)
    b.heights = np.asarray(hs, np.int32)
    b.widths = np.asarray(ws, np.int32)
    return b


class Recognizer:
    """Thin owner of a `kocr_handle` (one per GPU, one host thread)."""

    DEBUG_DTYPES = {"chunks": np.float32, "enc": np.float32, "memory": np.float32, "logits_trace": np.float32,
                    "dx": np.float32, "dy": np.float32, "daof": np.float32, "dq": np.float32, "dqkv": np.float32,
                    "dh": np.float32, "logits": np.float32}

    def __init__(self, weight_blob: bytes, device: int = 0, max_lines: int = 256, max_chunks: int = 4096):
        self.lib = load_library()
        if self.lib.kocr_abi_version() != 2:
            raise KocrError("ABI version mismatch between _native.py and libkocr_b200.so")
        self._h = C.c_void_p()
        buf = (C.c_char * len(weight_blob)).from_buffer_copy(weight_blob)
        check(self.lib.kocr_create(buf, len(weight_blob), device, max_lines, max_chunks, C.byref(self._h)))
        self.max_lines, self.max_chunks, self.device = max_lines, max_chunks, device
        v = [C.c_int() for _ in range(5)]
        check(self.lib.kocr_model_info(self._h, *[C.byref(x) for x in v]))
        self.variant, self.emb_dim, self.max_seq_len, self.decode_max_len, self.vocab_size = [x.value for x in v]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.kocr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stages ---------------------------------------------------------------------------
    def gather_chunks(self, batch: LineBatch, stream=None, pixels_dev_ptr=None):
        counts = np.zeros(max(batch.n, 1), np.int32)
        pix = _ptr(pixels_dev_ptr) if pixels_dev_ptr is not None else _ptr(batch.pixels)
        check(self.lib.kocr_gather_chunks(self._h, pix, batch.pixel_bytes, 1 if pixels_dev_ptr is not None else 0,
                                          _ptr(batch.offsets), _ptr(batch.heights), _ptr(batch.widths), batch.n,
                                          _ptr(counts), _ptr(stream)))
        return counts[:batch.n]

    def sevgg_encoder_forward(self, stream=None):
        check(self.lib.kocr_sevgg_encoder_forward(self._h, _ptr(stream)))

    def merge_bilstm_forward(self, stream=None):
        check(self.lib.kocr_merge_bilstm_forward(self._h, _ptr(stream)))

    def decode_greedy(self, n_lines: int, max_steps: int = 0, stream=None):
        tokens = np.zeros((max(n_lines, 1), TOKENS_LD), np.int32)
        lengths = np.zeros(max(n_lines, 1), np.int32)
        check(self.lib.kocr_decode_greedy(self._h, max_steps, _ptr(tokens), _ptr(lengths), _ptr(stream)))
        return tokens[:n_lines], lengths[:n_lines]

    def recognize_lines(self, batch: LineBatch, max_steps: int = 0, stream=None, pixels_dev_ptr=None,
                        tokens_out=None, lengths_out=None):
        tokens = tokens_out if tokens_out is not None else np.zeros((max(batch.n, 1), TOKENS_LD), np.int32)
        lengths = lengths_out if lengths_out is not None else np.zeros(max(batch.n, 1), np.int32)
        pix = _ptr(pixels_dev_ptr) if pixels_dev_ptr is not None else _ptr(batch.pixels)
        check(self.lib.kocr_recognize_lines(self._h, pix, batch.pixel_bytes, 1 if pixels_dev_ptr is not None else 0,
                                            _ptr(batch.offsets), _ptr(batch.heights), _ptr(batch.widths), batch.n,
                                            max_steps, _ptr(tokens), _ptr(lengths), _ptr(stream)))
        return tokens[:batch.n], lengths[:batch.n]

    # ---- test hooks -----------------------------------------------------------------------
    def set_option(self, name: str, value: int):
        check(self.lib.kocr_set_option(self._h, name.encode(), int(value)))

    def set_forced_tokens(self, tokens: np.ndarray):
        tokens = np.ascontiguousarray(tokens, np.int32)
        assert tokens.ndim == 2 and tokens.shape[1] == TOKENS_LD
        check(self.lib.kocr_set_forced_tokens(self._h, _ptr(tokens), tokens.shape[0]))

    def debug_read(self, name: str) -> np.ndarray:
        n = C.c_size_t()
        check(self.lib.kocr_debug_read(self._h, name.encode(), None, 0, C.byref(n)))
        if name in ("last_steps", "host_launch_us", "host_wait_us"):       # scalar hooks come back in the size field
            return np.asarray(n.value)
        dtype = self.DEBUG_DTYPES.get(name, np.uint16)       # bf16 buffers come back as raw uint16
        out = np.empty(n.value // np.dtype(dtype).itemsize, dtype)
        check(self.lib.kocr_debug_read(self._h, name.encode(), _ptr(out), out.nbytes, C.byref(n)))
        return out

    def beam_step(self, line: int, prefixes: np.ndarray, parents, t: int) -> np.ndarray:
        """One decoder position for the hypotheses `prefixes` [n_rows, t+1] of `line`; returns logits [n_rows, 124]."""
        prefixes = np.ascontiguousarray(prefixes, np.int32)
        n_rows = prefixes.shape[0]
        assert prefixes.shape[1] == t + 1
        par = np.ascontiguousarray(parents if parents is not None else np.zeros(n_rows), np.int32)
        logits = np.zeros((n_rows, 128), np.float32)
        check(self.lib.kocr_beam_step(self._h, line, n_rows, _ptr(par), _ptr(prefixes), t, _ptr(logits), None))
        return logits[:, :124]

    def forward_teacher_forced(self, tgt_tokens: np.ndarray) -> np.ndarray:
        """KhmerOCR.forward semantics on the batch whose stages 1-4 have just run: tgt_tokens int [B, L] -> logits
        fp32 [B, L, 124] (padded memory, un-packed BiLSTM, memory_key_padding_mask)."""
        tgt = np.ascontiguousarray(tgt_tokens, np.int32)
        B, L = tgt.shape
        out = np.zeros((B, L, 128), np.float32)
        check(self.lib.kocr_forward_teacher_forced(self._h, _ptr(tgt), L, _ptr(out), None))
        return out[:, :, :124]

    def beam_step_batch(self, row_line, prefixes: np.ndarray, parents, t: int) -> np.ndarray:
        """One decoder position for hypotheses of many lines: row r continues hypothesis `parents[r]` (row of the previous
        call) of line `row_line[r]`; prefixes [n_rows, t+1]; returns logits [n_rows, 124]."""
        prefixes = np.ascontiguousarray(prefixes, np.int32)
        n_rows = prefixes.shape[0]
        assert prefixes.shape[1] == t + 1
        rl = np.ascontiguousarray(row_line, np.int32)
        par = np.ascontiguousarray(parents if parents is not None else np.zeros(n_rows), np.int32)
        logits = np.zeros((n_rows, 128), np.float32)
        check(self.lib.kocr_beam_step_batch(self._h, n_rows, _ptr(rl), _ptr(par), _ptr(prefixes), t, _ptr(logits), None))
        return logits[:, :124]

    def beam_search(self, n_lines: int, beam_width: int, max_len: int = 0):
        """`OCRPredictor._beam_search` for every line of the batch whose stages 1-5a have just run: (tokens [n, 257], lengths)."""
        tokens = np.zeros((max(n_lines, 1), TOKENS_LD), np.int32)
        lengths = np.zeros(max(n_lines, 1), np.int32)
        check(self.lib.kocr_beam_search(self._h, int(beam_width), int(max_len), _ptr(tokens), _ptr(lengths), None))
        return tokens[:n_lines], lengths[:n_lines]

    # ---- the reference's model protocol (predictor.py:53-78,166-192), numpy in / numpy out ------------------------------
    def model_cnn(self, chunks: np.ndarray) -> np.ndarray:
        """`model.cnn`: chunks fp32 (n, 1, 48, 100) -> features fp32 (n, 512, 2, 32)."""
        x = np.ascontiguousarray(chunks, np.float32).reshape(-1, 1, 48, 100)
        out = np.empty((x.shape[0], 512, 2, 32), np.float32)
        check(self.lib.kocr_model_cnn(self._h, _ptr(x), x.shape[0], _ptr(out), None))
        return out

    def model_patch(self, f: np.ndarray) -> np.ndarray:
        """`model.patch`: features (n, 512, 2, 32) -> patch embeddings + local positions (n, 32, 384)."""
        x = np.ascontiguousarray(f, np.float32).reshape(-1, 512, 2, 32)
        out = np.empty((x.shape[0], 32, self.emb_dim), np.float32)
        check(self.lib.kocr_model_patch(self._h, _ptr(x), x.shape[0], _ptr(out), None))
        return out

    def model_enc(self, p: np.ndarray) -> np.ndarray:
        """`model.enc`: seq-first (32, n, 384) -> (32, n, 384)."""
        x = np.ascontiguousarray(p, np.float32)
        assert x.ndim == 3 and x.shape[0] == 32 and x.shape[2] == self.emb_dim
        out = np.empty_like(x)
        check(self.lib.kocr_model_enc(self._h, _ptr(x), x.shape[1], _ptr(out), None))
        return out

    def model_bilstm(self, merged: np.ndarray) -> np.ndarray:
        """`model.context_bilstm` (zero initial state): (B, T, 384) -> (B, T, 384)."""
        x = np.ascontiguousarray(merged, np.float32)
        assert x.ndim == 3 and x.shape[2] == self.emb_dim
        out = np.empty_like(x)
        check(self.lib.kocr_model_bilstm(self._h, _ptr(x), x.shape[0], x.shape[1], _ptr(out), None))
        return out

    def model_dec(self, tgt: np.ndarray, memory: np.ndarray, pad_mask=None) -> np.ndarray:
        """`model.dec(tgt, memory, memory_key_padding_mask)`: tgt int (B, t), memory (B, T, 384), mask bool (B, T) with True =
        padded (a suffix per line) or None -> logits fp32 (B, t, vocab)."""
        tg = np.ascontiguousarray(tgt, np.int32)
        mem = np.ascontiguousarray(memory, np.float32)
        assert tg.ndim == 2 and mem.ndim == 3 and mem.shape[0] == tg.shape[0] and mem.shape[2] == self.emb_dim
        mk = np.ascontiguousarray(pad_mask, np.uint8) if pad_mask is not None else None
        out = np.empty((tg.shape[0], tg.shape[1], self.vocab_size), np.float32)
        check(self.lib.kocr_model_dec(self._h, _ptr(tg), tg.shape[0], tg.shape[1], _ptr(mem), mem.shape[1], _ptr(mk), _ptr(out), None))
        return out

    def crop_lines(self, page, boxes, pad_px: int, out_dev_ptr: int, out_offsets, page_dev_ptr=None, stream=None):
        """Cut `boxes` (int32 [n, 4] = x0, y0, x1, y1, clipped) out of `page` (uint8 (H, W) or (H, W, 3); pass
        `page_dev_ptr` if it already lives on the device), add `pad_px` of white and convert to grey, into the device
        buffer at `out_dev_ptr` (line i at out_offsets[i])."""
        boxes = np.ascontiguousarray(boxes, np.int32).reshape(-1, 4)
        offs = np.ascontiguousarray(out_offsets, np.int64)
        ch = 1 if page.ndim == 2 else int(page.shape[2])
        src = _ptr(page_dev_ptr) if page_dev_ptr is not None else _ptr(np.ascontiguousarray(page))
        check(self.lib.kocr_crop_lines(self._h, src, int(page.shape[0]), int(page.shape[1]), ch,
                                       1 if page_dev_ptr is not None else 0, _ptr(boxes), boxes.shape[0], int(pad_px),
                                       _ptr(out_dev_ptr), _ptr(offs), _ptr(stream)))

    def unfinished(self, n_lines: int) -> np.ndarray:
        """Flags of the lines left incomplete by an early return (option "straggler_threshold")."""
        flags = np.zeros(max(n_lines, 1), np.int32)
        check(self.lib.kocr_read_unfinished(self._h, _ptr(flags)))
        return flags[:n_lines]

    def kernel_timing(self) -> dict:
        """{site: {"ms": total, "launches": n, "flops": algorithmic FLOPs over those launches}}"""
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.kocr_read_kernel_timing(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n, fl = line.split()
            out[name] = {"ms": float(ms), "launches": int(n), "flops": float(fl)}
        return out

    def workspace_bytes(self) -> int:
        return int(self.lib.kocr_workspace_bytes(self._h))


def launch_count() -> int:
    return int(load_library().kocr_launch_count())
