"""OCRPredictor - inference driver with the reference's constructor and predict / predict_batch
signatures (reference: netra_ocr/recognition/predictor.py:12-199), running on libkocr_b200.so.

Differences that do not change results: chunks of many lines are batched on the GPU regardless of
`batch_size` (lines are independent, predictor.py:150-193), BiLSTM/decoding are batched across lines
with a KV cache instead of one line and one full-prefix pass per token.  Beam search (beam_width > 1) runs inside the
library for all lines of a batch (kocr_beam_search: the reference's bookkeeping, KV-cached positions).  `self.model` is the
device handle and carries the reference's model protocol (cnn / patch / enc / global_pos / context_bilstm / dec).
There is no CPU path."""
import logging
from pathlib import Path

import numpy as np

from .config import OCRConfig
from .tokenizer import Tokenizer
from .preprocessor import ImagePreprocessor
from . import _core
load_checkpoint, detect_variant = _core.checkpoint.load_checkpoint, _core.checkpoint.detect_variant
pack_blob = _core.weights.pack_blob
Recognizer, LineBatch = _core.native.Recognizer, _core.native.LineBatch
plan_batches = _core.scheduling.plan_batches

logger = logging.getLogger(__name__)


def _attach_model_protocol(rec, sd, variant):
    """The attributes the reference's predictor uses on `self.model` (predictor.py:26-31,53-78,166-192): `cnn(chunks)`,
    `patch(f) -> (x, N)`, `enc(p)` seq-first, `global_pos`, `context_bilstm(merged) -> (out, state)` (SE-VGG family only: the
    reference probes it with hasattr) and `dec(tgt, memory, mask)`, on torch tensors, backed by the kocr_model_* entry points.
    With them the reference's own OCRPredictor code runs unchanged on this object (tests/test_gpu_model_protocol.py)."""
    import torch

    def f32(t):
        return t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t, np.float32)

    rec.global_pos = torch.from_numpy(np.array(sd["global_pos"], np.float32))
    rec.cnn = lambda chunks: torch.from_numpy(rec.model_cnn(f32(chunks)))
    rec.patch = lambda f: (torch.from_numpy(rec.model_patch(f32(f))), 32)
    rec.enc = lambda p: torch.from_numpy(rec.model_enc(f32(p)))
    if variant == "se":
        class _ContextBiLSTM:
            def __call__(self, merged):
                return torch.from_numpy(rec.model_bilstm(f32(merged))), None

            def flatten_parameters(self):           # (predictor.py:75 calls it on the nn.LSTM)
                pass
        rec.context_bilstm = _ContextBiLSTM()

    def dec(tgt, memory, mask=None):
        tg = tgt.detach().cpu().numpy() if hasattr(tgt, "detach") else np.asarray(tgt)
        mk = None if mask is None else (mask.detach().cpu().numpy() if hasattr(mask, "detach") else np.asarray(mask))
        return torch.from_numpy(rec.model_dec(tg.astype(np.int32), f32(memory), None if mk is None else mk.astype(np.uint8)))
    rec.dec = dec
    rec.eval = lambda: rec                      # (nn.Module no-ops the reference calls on its model)
    rec.to = lambda *a, **k: rec


class OCRPredictor:
    def __init__(self, model_path, tokenizer: Tokenizer, config: OCRConfig, model_class,
                 max_lines: int = 256, max_chunks: int = 2816, in_flight: int = 6):
        self.cfg = config
        self.tokenizer = tokenizer
        import torch
        self.device = torch.device(self.cfg.device)
        if self.device.type != "cuda":
            raise RuntimeError("khmer_ocr_cnn_transformer_b200 has no CPU path: OCRConfig.device must be 'cuda'")
        # the kernels hard-code the reference vocabulary layout (char2idx.json: <pad>=0, <sos>=2, <eos>=3, 124 symbols)
        layout = (tokenizer.pad_idx, tokenizer.sos_idx, tokenizer.eos_idx, len(tokenizer))
        if layout != (0, 2, 3, 124):
            raise ValueError(f"vocabulary layout (pad, sos, eos, size) = {layout}; the CUDA path is specialised for the "
                             "reference's char2idx.json layout (0, 2, 3, 124)")
        logger.info(f"Init Model: dim={self.cfg.emb_dim}, max_seq={self.cfg.max_seq_len}")
        self.model_spec = model_class(vocab_size=len(tokenizer), pad_idx=tokenizer.pad_idx,
                                      emb_dim=self.cfg.emb_dim, max_global_len=self.cfg.max_seq_len)
        self._max_lines, self._max_chunks, self._in_flight = max_lines, max_chunks, in_flight
        self._load_weights(model_path)
        self.preprocessor = ImagePreprocessor(config, self.model)

    def _load_weights(self, path):
        sd = load_checkpoint(Path(path))
        variant = detect_variant(sd)
        want = getattr(self.model_spec, "variant", variant)
        if variant != want:
            logger.warning(f"checkpoint looks like '{variant}' but model_class is '{want}'; using the checkpoint")
        index = self.device.index if self.device.index is not None else 0
        # `in_flight` recognisers (handle + workspace + stream each, ~2 MB of workspace per chunk of capacity) behind one
        # pipeline: greedy batches stream through all of them; `self.model` is the first one (beam search, page crops,
        # teacher-forced forward and the stage-level calls use it alone)
        self._pipe = _core.pipeline().LinePipeline(pack_blob(sd), device=index, in_flight=self._in_flight,
                                                   max_lines=self._max_lines, max_chunks=self._max_chunks)
        self.model = self._pipe.recs[0]
        _attach_model_protocol(self.model, sd, variant)

    def close(self):
        """Release every device handle of the predictor (weights, workspaces, streams)."""
        self._pipe.close()

    # ------------------------------------------------------------------------------------
    def _decode_ids(self, tokens, lengths):
        """`Tokenizer.decode` (tokenizer.py:26-35: skip <sos> / <pad>, stop at <eos>, unknown ids -> "") for every row,
        through a lookup table instead of a Python loop per token (the loop was 30 % of `predict_batch` on 8192 lines)."""
        tk = self.tokenizer
        lut = getattr(self, "_id2char_lut", None)
        if lut is None:
            lut = np.array([tk.idx2char.get(i, "") for i in range(max(len(tk), 128))], dtype=object)
            lut[tk.sos_idx] = ""
            lut[tk.pad_idx] = ""
            self._id2char_lut = lut
        out = []
        for i in range(tokens.shape[0]):
            row = tokens[i, :lengths[i]]
            stop = np.nonzero(row == tk.eos_idx)[0]
            if stop.size:
                row = row[:stop[0]]
            out.append("".join(lut[np.clip(row, 0, lut.shape[0] - 1)].tolist()))
        return out

    def _recognize_gray(self, grays):
        """Greedy recognition of grey uint8 lines through the pipeline: batches within the handles' capacity (sorted by
        length, results in input order), several passes in flight, the long tail of every pass pooled and decoded
        together (pipeline.py); every batch is detokenised by its worker thread as soon as its last line is final, while
        the other passes are still on the GPU.  Greedy decoding is deterministic and lines are independent: results do
        not depend on the batching."""
        n = len(grays)
        texts = [None] * n
        if n == 0:
            return texts
        Job = _core.pipeline().Job
        tokens = np.zeros((n, _core.native.TOKENS_LD), np.int32)
        lengths = np.zeros(n, np.int32)
        jobs = [Job(ids, images=[grays[i] for i in ids]) for ids in self._pipe.plan([g.shape for g in grays])]

        def on_done(job):
            for i, text in zip(job.ids, self._decode_ids(tokens[job.ids], lengths[job.ids])):
                texts[i] = text

        self._pipe.run_jobs(jobs, tokens, lengths, max_steps=self.cfg.decode_max_len, image_of=lambda i: grays[i], on_done=on_done)
        return texts

    # ------------------------------------------------------------------------------------
    def _beam_search_batch(self, n_lines: int, beam_width: int) -> list:
        """`OCRPredictor._beam_search` (reference predictor.py:101-136) for every line of the batch whose stages 1-5a
        have just run.  Each decoder position of ALL lines' hypotheses is one GPU pass (kocr_beam_step_batch, KV-cached:
        the reference re-runs the whole prefix); log-softmax and top-k are torch's, like the reference's; the
        candidate / pruning / completion bookkeeping is `beam.BatchedBeam` (same arithmetic and tie order, vectorised)."""
        import torch
        import torch.nn.functional as F
        from .beam import BatchedBeam
        beam = BatchedBeam(n_lines, beam_width, self.tokenizer.sos_idx, self.tokenizer.eos_idx, self.cfg.decode_max_len)
        for t in range(self.cfg.decode_max_len):
            if beam.live_lines().size == 0:
                break
            row_line, prefixes, parents = beam.rows()
            logits = self.model.beam_step_batch(row_line, prefixes, parents if t > 0 else None, t)
            log_probs = F.log_softmax(torch.from_numpy(logits.copy()), dim=-1)
            top_probs, top_idxs = log_probs.topk(beam_width, dim=-1)
            beam.update(top_probs.numpy(), top_idxs.numpy())
        return [self.tokenizer.decode(seq) for seq in beam.results()]

    def _beam_gray(self, grays, beam_width: int):
        """Beam search for a list of grey lines: stages 1-5a per batch, then `kocr_beam_search` - the whole of
        `OCRPredictor._beam_search` (reference predictor.py:101-136) for every line of the batch inside the library
        (device log-softmax + top-k, the reference's bookkeeping in C++), several batches in flight on their own handles.
        `_beam_search_batch` (the same bookkeeping in numpy over `kocr_beam_step_batch`) is kept as the cross-check."""
        if beam_width > 8:
            raise ValueError("beam_width > 8 is not supported by the CUDA path")
        results = [None] * len(grays)
        # every hypothesis is a row of the decode workspace: at most max_lines // beam_width lines per pass
        lines_per_pass = max(1, self._max_lines // max(beam_width, 1))
        import threading
        left, lock = [], threading.Lock()

        def make_run(tail_div):
            def run(rec, idxs):
                # long tail: low-probability alternatives that never emit <eos> ramble on to decode_max_len and would hold the
                # whole pass; a pass stops once <= 1/tail_div of its lines still have live hypotheses, those are searched
                # again together afterwards (deterministic: same result)
                rec.set_option("straggler_threshold", len(idxs) // tail_div if tail_div else 0)
                rec.gather_chunks(LineBatch([grays[i] for i in idxs]))
                rec.sevgg_encoder_forward()
                rec.merge_bilstm_forward()
                tokens, lengths = rec.beam_search(len(idxs), beam_width, self.cfg.decode_max_len)
                todo = rec.unfinished(len(idxs)) if tail_div else np.zeros(len(idxs), np.int32)
                texts = self._decode_ids(tokens, lengths)
                with lock:
                    for j, i in enumerate(idxs):
                        if todo[j]:
                            left.append(i)
                        else:
                            results[i] = texts[j]
            return run

        def plan(ids):
            return [[ids[j] for j in b] for b in plan_batches([grays[i].shape for i in ids], lines_per_pass, self._max_chunks,
                                                              self.cfg.max_seq_len)]

        self._pipe.map_batches(plan(list(range(len(grays)))), make_run(8))
        if left:
            again, left[:] = sorted(left), []
            self._pipe.map_batches(plan(again), make_run(0))
        for r in self._pipe.recs:
            r.set_option("straggler_threshold", 0)
        return results

    def _beam_gray_host(self, grays, beam_width: int):
        """The round-1 path: one device pass per decoder position, bookkeeping in numpy (`beam.BatchedBeam`)."""
        if beam_width > 8:
            raise ValueError("beam_width > 8 is not supported by the CUDA path")
        results = [None] * len(grays)
        lines_per_pass = max(1, self._max_lines // max(beam_width, 1))
        for idxs in plan_batches([g.shape for g in grays], lines_per_pass, self._max_chunks, self.cfg.max_seq_len):
            self.model.gather_chunks(LineBatch([grays[i] for i in idxs]))
            self.model.sevgg_encoder_forward()
            self.model.merge_bilstm_forward()
            for i, text in zip(idxs, self._beam_search_batch(len(idxs), beam_width)):
                results[i] = text
        return results

    def forward_logits(self, image_list: list, tgt_tokens) -> np.ndarray:
        """Teacher-forced batched forward with the reference's TRAINING-time semantics - `KhmerOCR.forward(chunk_lists,
        tgt_tokens)` (model/se_model.py:240-289), as used by its CER evaluation loop (notebook cell 19): memory padded to
        the longest line of the batch, BiLSTM over the pads (no packing), memory_key_padding_mask, <pad> target keys
        masked.  image_list: paths / PIL images / grey arrays (the reference builds `chunk_lists` from the same
        ImagePreprocessor); tgt_tokens: int (B, L), rows = <sos> + target right-padded with <pad>.  Returns fp32
        logits (B, L, vocab).  The whole list is one device batch (B <= max_lines, B * Tmax <= 32 * max_chunks)."""
        grays = [ImagePreprocessor.to_gray(im) for im in image_list]
        tgt = np.asarray(tgt_tokens)
        if tgt.ndim != 2 or tgt.shape[0] != len(grays):
            raise ValueError("tgt_tokens must have shape (len(image_list), L)")
        self.model.gather_chunks(LineBatch(grays))
        self.model.sevgg_encoder_forward()
        return self.model.forward_teacher_forced(tgt)

    def predict_page(self, image, textline_pred, expansion_px: int = 5, padding_px: int = 10) -> list:
        """Page image + detected text-line polygons -> one greedy-decoded string per line, in detection order: the
        `extract_textline_crops` -> `recognize_batch(crops, beam_width=1)` sequence of OCREngine.process_image
        (ocr_engine.py:54-86, textline_detection.py:7-53) with the crops cut, padded and grey-converted on the GPU."""
        tc = _core.textline_crops()
        textline_boxes, crop_lines_device = tc.textline_boxes, tc.crop_lines_device
        from PIL import Image
        if isinstance(image, (str, Path)):
            image = Image.open(image).convert("RGB")
        size = image.size if hasattr(image, "size") and not isinstance(image, np.ndarray) else (image.shape[1], image.shape[0])
        boxes = textline_boxes(size, textline_pred, expansion_px)
        if not boxes:
            return []
        crops = crop_lines_device(self.model, image, boxes, padding_px)
        shapes = list(zip(crops.batch.heights.tolist(), crops.batch.widths.tolist()))
        line_batch_from_shapes = _core.native.line_batch_from_shapes
        results = [None] * len(boxes)
        self.model.set_option("straggler_threshold", 0)
        for idxs in plan_batches(shapes, self._max_lines, self._max_chunks, self.cfg.max_seq_len):
            if list(idxs) != list(range(idxs[0], idxs[0] + len(idxs))):
                raise RuntimeError("plan_batches must keep device-resident crops contiguous")
            sub = line_batch_from_shapes([shapes[i] for i in idxs])
            tokens, lengths = self.model.recognize_lines(sub, max_steps=self.cfg.decode_max_len,
                                                         pixels_dev_ptr=crops.dev_ptr + int(crops.batch.offsets[idxs[0]]))
            for j, text in zip(idxs, self._decode_ids(tokens, lengths)):
                results[j] = text
        return results

    def predict(self, image_input, beam_width: int = 3) -> str:
        gray = ImagePreprocessor.to_gray(image_input)
        if beam_width <= 1:
            return self._recognize_gray([gray])[0]
        return self._beam_gray([gray], beam_width)[0]

    def predict_batch(self, image_list: list, beam_width: int = 1, batch_size: int = 8) -> list:
        if not image_list:
            return []
        grays = [ImagePreprocessor.to_gray(im) for im in image_list]
        if beam_width <= 1:
            return self._recognize_gray(grays)
        return self._beam_gray(grays, beam_width)
