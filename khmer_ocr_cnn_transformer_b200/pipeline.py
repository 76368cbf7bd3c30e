"""LinePipeline - keeps several device passes in flight on ONE GPU.

One decode chain (a batch of lines stepping through ~100 greedy positions) leaves most of the 148 SMs idle, while the
SE-VGG / encoder stage of another batch can use them.  The pipeline therefore owns `in_flight` recognisers - each a
`kocr_handle` with its own weights, workspace and stream - and one host thread per recogniser; the threads pull jobs
(batches of lines) from a shared queue and spend their time inside the C ABI (ctypes releases the GIL), so the device
sees `in_flight` independent streams.  Lines are independent in the reference (predictor.py:150-193), so the decoded ids
do not depend on how lines are batched.

Long tail: a pass returns as soon as at most `straggler_threshold` lines per 256 are still decoding
(kocr_read_unfinished); those stragglers go to a pool shared by the passes and are decoded to the end together in later
passes (greedy decoding is deterministic: same ids).

Used by `OCRPredictor.predict_batch` (the public API) and by bench.py (its e2e measurement)."""
from __future__ import annotations

import queue
import threading

import numpy as np

from . import _native
from .scheduling import chunks_for, plan_batches

TOKENS_LD = _native.TOKENS_LD


class Job:
    """A batch of lines for one device pass.  `ids` are the caller's line numbers (rows of the output arrays)."""
    __slots__ = ("ids", "images", "batch", "dev_ptr", "tag")

    def __init__(self, ids, images=None, batch=None, dev_ptr=None, tag=None):
        self.ids, self.images, self.batch, self.dev_ptr, self.tag = ids, images, batch, dev_ptr, tag


class LinePipeline:
    def __init__(self, weight_blob: bytes, device: int = 0, in_flight: int = 4, max_lines: int = 256,
                 max_chunks: int = 2816, straggler_per_256: int = 8, options: dict | None = None):
        self.device, self.in_flight = device, max(1, int(in_flight))
        self.max_lines, self.max_chunks = max_lines, max_chunks
        self.straggler_per_256 = straggler_per_256
        self._blob, self._options = weight_blob, dict(options or {})
        # recognisers are created on demand (a single `predict` needs one; ~2 MB of workspace per chunk of capacity each)
        self.recs = []
        self._mode = None
        self._ensure(1)
        self.max_seq_len = self.recs[0].max_seq_len
        self._lock = threading.Lock()
        self.stats = {"passes": 0, "straggler_passes": 0, "stragglers": 0}

    def _ensure(self, n: int):
        while len(self.recs) < min(n, self.in_flight):
            r = _native.Recognizer(self._blob, device=self.device, max_lines=self.max_lines, max_chunks=self.max_chunks)
            for k, v in self._options.items():
                r.set_option(k, v)
            self.recs.append(r)
            self._mode = None

    def _set_mode(self, concurrent: bool):
        """Throughput settings when several passes share the GPU, latency settings for a lone pass."""
        if self._mode == concurrent:
            return
        for r in self.recs:
            r.set_option("blocking_wait", 1 if concurrent else 0)   # many host threads per GPU must not spin inside stream waits
            r.set_option("use_pdl", 0 if concurrent else 1)         # early-resident dependents only hold SM slots when streams compete
            r.set_option("dec_wide", 0 if concurrent else 1)        # few fat CTAs per decode GEMM leave the SMs to the other streams
            for k, v in self._options.items():
                r.set_option(k, v)
        self._mode = concurrent

    def close(self):
        for r in self.recs:
            r.close()
        self.recs = []

    # ---- planning -----------------------------------------------------------------------------------------------
    def plan(self, shapes, sort_by_length: bool = True):
        """Batches of line indices within the handles' capacity.  Sorting by chunk count makes batches homogeneous: the
        lines of a pass then finish decoding together and a pass of short lines holds more of them."""
        order = list(range(len(shapes)))
        if sort_by_length:
            order.sort(key=lambda i: (-chunks_for(shapes[i][0], shapes[i][1], self.max_seq_len), i))
        groups = plan_batches([shapes[i] for i in order], self.max_lines, self.max_chunks, self.max_seq_len)
        return [[order[j] for j in g] for g in groups]

    # ---- execution ----------------------------------------------------------------------------------------------
    def run_jobs(self, jobs, tokens: np.ndarray, lengths: np.ndarray, max_steps: int = 0, image_of=None,
                 on_done=None, final_exact: bool = True):
        """Run `jobs` (list[Job]) over the in-flight recognisers; results go to rows `job.ids` of `tokens` / `lengths`.
        `image_of(i)` returns the grey image of caller line i (needed to re-submit stragglers).  `on_done(job)` is called
        (from a worker thread) when every line of the job has its final result."""
        q: queue.Queue = queue.Queue()
        for j in jobs:
            q.put(j)
        n_jobs = len(jobs)
        n_threads = max(1, min(self.in_flight, n_jobs))
        self._ensure(n_threads)
        self._set_mode(n_threads > 1)
        pool: list = []            # (line id, owning job) of stragglers waiting for a pass of their own
        pending = {id(j): 0 for j in jobs}     # stragglers of a job still in the pool / in a straggler pass
        errors: list = []
        can_pool = image_of is not None and self.straggler_per_256 > 0

        def finish(job):
            if on_done is not None:
                on_done(job)

        def straggler_pass(rec, part):
            rec.set_option("straggler_threshold", 0)
            imgs = [image_of(i) for i, _ in part]
            for sub in self.plan([im.shape for im in imgs]):          # (a pool of long lines can exceed one pass's chunk capacity)
                ids = [part[k][0] for k in sub]
                tok, ln = rec.recognize_lines(_native.LineBatch([imgs[k] for k in sub]), max_steps=max_steps)
                tokens[ids] = tok
                lengths[ids] = ln
            done = []
            with self._lock:
                self.stats["straggler_passes"] += 1
                for _, owner in part:
                    pending[id(owner)] -= 1
                    if pending[id(owner)] == 0:
                        done.append(owner)
            for owner in {id(o): o for o in done}.values():
                finish(owner)

        def worker(rec):
            try:
                while True:
                    try:
                        job = q.get_nowait()
                    except queue.Empty:
                        break
                    n = len(job.ids)
                    # the last pass of every worker decodes all of its lines in place: the run must not end with a lone,
                    # latency-bound straggler pass
                    last = final_exact and q.qsize() < n_threads
                    thr = 0 if (not can_pool or last) else max(1, self.straggler_per_256 * n // 256)
                    rec.set_option("straggler_threshold", thr)
                    batch = job.batch if job.batch is not None else _native.LineBatch(job.images)
                    tok, ln = rec.recognize_lines(batch, max_steps=max_steps, pixels_dev_ptr=job.dev_ptr)
                    tokens[job.ids] = tok
                    lengths[job.ids] = ln
                    todo = np.nonzero(rec.unfinished(n))[0] if thr > 0 else ()
                    part = None
                    with self._lock:
                        self.stats["passes"] += 1
                        self.stats["stragglers"] += len(todo)
                        pending[id(job)] += len(todo)
                        pool.extend((job.ids[k], job) for k in todo)
                        if len(pool) >= self.max_lines:
                            part = pool[:self.max_lines]
                            del pool[:self.max_lines]
                    if len(todo) == 0:
                        finish(job)
                    if part:
                        straggler_pass(rec, part)
                # no jobs left: drain the pool (every worker takes what is there until it is empty)
                while True:
                    with self._lock:
                        part = pool[:self.max_lines]
                        del pool[:self.max_lines]
                    if not part:
                        break
                    straggler_pass(rec, part)
            except Exception as e:          # surface worker failures instead of hanging
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(self.recs[k],)) for k in range(n_threads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def map_batches(self, batches, fn):
        """Run `fn(recogniser, batch)` for every element of `batches`, up to `in_flight` at a time, each call on its own
        recogniser and host thread (beam search, or any other multi-call sequence that must stay on one handle)."""
        q: queue.Queue = queue.Queue()
        for b in batches:
            q.put(b)
        n_threads = max(1, min(self.in_flight, len(batches)))
        self._ensure(n_threads)
        self._set_mode(n_threads > 1)
        errors: list = []

        def worker(rec):
            try:
                while True:
                    try:
                        b = q.get_nowait()
                    except queue.Empty:
                        return
                    fn(rec, b)
            except Exception as e:
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(self.recs[k],)) for k in range(n_threads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def recognize(self, grays, max_steps: int = 0):
        """Greedy recognition of grey uint8 line images -> (tokens int32 [n, 257], lengths int32 [n]) in input order."""
        n = len(grays)
        tokens = np.zeros((n, TOKENS_LD), np.int32)
        lengths = np.zeros(n, np.int32)
        if n == 0:
            return tokens, lengths
        jobs = [Job(ids, images=[grays[i] for i in ids]) for ids in self.plan([g.shape for g in grays])]
        self.run_jobs(jobs, tokens, lengths, max_steps=max_steps, image_of=lambda i: grays[i])
        return tokens, lengths
