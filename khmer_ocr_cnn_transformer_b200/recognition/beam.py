"""Beam-search bookkeeping of `OCRPredictor._beam_search` (reference: netra_ocr/recognition/predictor.py:101-136) for MANY
lines at once, vectorised with numpy - the decoder positions run on the GPU (`kocr_beam_step_batch`), this module owns the
host side and reproduces the reference's arithmetic and tie order exactly:

  * scores are Python floats there (sums of `.item()` values) = IEEE float64 adds of the fp32 log-probabilities here;
  * candidates of a line are listed hypothesis-major, top-k-minor and sorted with Python's stable `sort(reverse=True)`
    = `np.argsort(-score, kind="stable")` (equal scores keep list order);
  * EVERY candidate ending in <eos> is moved to `completed` with score / len(seq) (len counts <sos> and <eos>), wherever it
    ranks; the first `beam_width` other candidates survive;
  * the answer is the best completed hypothesis - `sorted(completed, reverse=True)[0]`, i.e. the first-appended among
    equal scores; all <eos> candidates of one step have the same length, so only the best of a step can ever win and
    a strict `>` against the best so far keeps the reference's order - else the first live hypothesis.
"""
from __future__ import annotations

import numpy as np


class BatchedBeam:
    def __init__(self, n_lines: int, beam_width: int, sos: int, eos: int, max_len: int):
        self.n, self.bw, self.sos, self.eos, self.max_len = n_lines, beam_width, sos, eos, max_len
        self.t = 0                                                    # prefixes have t + 1 tokens
        self.k = np.ones(n_lines, np.int64)                           # live hypotheses per line (0 = line finished)
        self.scores = np.zeros((n_lines, beam_width), np.float64)
        self.seqs = np.zeros((n_lines, beam_width, max_len + 2), np.int32)
        self.seqs[:, 0, 0] = sos
        self.prev_row = np.zeros((n_lines, beam_width), np.int64)     # device row of each live hypothesis in the last pass
        self.best_score = np.full(n_lines, -np.inf, np.float64)       # best completed hypothesis so far
        self.best_seq = [None] * n_lines
        self._row_start = np.zeros(n_lines, np.int64)

    # ---- device pass description ------------------------------------------------------------------------------------
    def live_lines(self) -> np.ndarray:
        return np.nonzero(self.k > 0)[0]

    def rows(self):
        """(row_line, prefixes [R, t+1], parents [R]) of the next device pass, line-major / hypothesis-minor."""
        live = self.live_lines()
        counts = self.k[live]
        row_line = np.repeat(live, counts)
        starts = np.cumsum(counts) - counts
        self._row_start[:] = -1
        self._row_start[live] = starts
        hyp = np.arange(row_line.shape[0]) - np.repeat(starts, counts)
        prefixes = self.seqs[row_line, hyp, : self.t + 1]
        parents = self.prev_row[row_line, hyp]
        self._row_hyp = hyp
        return row_line.astype(np.int32), np.ascontiguousarray(prefixes), parents.astype(np.int32)

    # ---- one position -----------------------------------------------------------------------------------------------
    def update(self, top_vals: np.ndarray, top_idx: np.ndarray) -> None:
        """top_vals / top_idx [R, beam_width]: `log_probs[row].topk(beam_width)` of the rows returned by `rows()`."""
        bw, n = self.bw, self.n
        live = self.live_lines()
        R = top_vals.shape[0]
        row_line = np.repeat(live, self.k[live])
        hyp = self._row_hyp
        # candidate table [n_lines, bw * bw]: hypothesis-major, top-k-minor; missing hypotheses = -inf
        cand = np.full((n, bw, bw), -np.inf, np.float64)
        tok = np.zeros((n, bw, bw), np.int64)
        cand[row_line, hyp] = self.scores[row_line, hyp][:, None] + top_vals.astype(np.float64)
        tok[row_line, hyp] = top_idx
        cand = cand.reshape(n, bw * bw)
        tok = tok.reshape(n, bw * bw)
        order = np.argsort(-cand, axis=1, kind="stable")              # Python's stable sort(reverse=True)
        s_sorted = np.take_along_axis(cand, order, 1)
        t_sorted = np.take_along_axis(tok, order, 1)
        valid = np.isfinite(s_sorted)
        is_eos = valid & (t_sorted == self.eos)
        # completed: the first <eos> candidate in sorted order is the best of this step (same length for all)
        seq_len = self.t + 2
        has = is_eos.any(axis=1)
        first = np.argmax(is_eos, axis=1)
        for l in np.nonzero(has)[0]:
            sc = s_sorted[l, first[l]] / seq_len
            if sc > self.best_score[l]:
                parent = order[l, first[l]] // bw
                self.best_score[l] = sc
                self.best_seq[l] = self.seqs[l, parent, : self.t + 1].tolist() + [self.eos]
        # survivors: the first `bw` non-<eos> candidates
        keep = valid & ~is_eos
        rank = np.cumsum(keep, axis=1) - 1
        keep &= rank < bw
        new_k = keep.sum(axis=1)
        new_scores = np.zeros_like(self.scores)
        new_seqs = np.zeros_like(self.seqs)
        new_prev = np.zeros_like(self.prev_row)
        ls, cs = np.nonzero(keep)
        if ls.size:
            slot = rank[ls, cs]
            parent = order[ls, cs] // bw
            new_scores[ls, slot] = s_sorted[ls, cs]
            new_seqs[ls, slot, : self.t + 1] = self.seqs[ls, parent, : self.t + 1]
            new_seqs[ls, slot, self.t + 1] = t_sorted[ls, cs]
            new_prev[ls, slot] = self._row_start[ls] + parent
        self.scores, self.seqs, self.prev_row, self.k = new_scores, new_seqs, new_prev, new_k
        self.t += 1
        del R

    def results(self):
        """Token id list per line, as the reference hands it to `Tokenizer.decode`."""
        out = []
        for l in range(self.n):
            if self.best_seq[l] is not None:
                out.append(self.best_seq[l])
            else:
                out.append(self.seqs[l, 0, : self.t + 1].tolist())
        return out
