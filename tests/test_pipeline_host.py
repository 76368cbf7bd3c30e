"""Host logic of LinePipeline (job queue, worker threads, straggler pool, per-job completion callbacks) with a fake
recogniser in place of the CUDA library: every line must come back exactly once, in input order, whatever the batching."""
import threading

import numpy as np
import pytest

from khmer_ocr_cnn_transformer_b200 import _native, pipeline


class FakeRecognizer:
    """`recognize_lines` "decodes" a line to [2, h, w, first pixel]; with a straggler threshold > 0 it leaves the lines whose
    first pixel is odd unfinished (up to the threshold), like an early return of kocr_decode_greedy."""
    created = 0

    def __init__(self, blob, device=0, max_lines=256, max_chunks=4096):
        FakeRecognizer.created += 1
        self.max_lines, self.max_chunks, self.max_seq_len = max_lines, max_chunks, 4096
        self.opts, self._flags, self.calls = {}, None, 0

    def set_option(self, k, v):
        self.opts[k] = v

    def recognize_lines(self, batch, max_steps=0, stream=None, pixels_dev_ptr=None, tokens_out=None, lengths_out=None):
        self.calls += 1
        assert batch.n <= self.max_lines
        from khmer_ocr_cnn_transformer_b200.scheduling import chunks_for
        assert sum(chunks_for(int(h), int(w)) for h, w in zip(batch.heights, batch.widths)) <= self.max_chunks
        tok = np.zeros((batch.n, _native.TOKENS_LD), np.int32)
        ln = np.zeros(batch.n, np.int32)
        flags = np.zeros(batch.n, np.int32)
        left = self.opts.get("straggler_threshold", 0)
        for i in range(batch.n):
            first = int(batch.pixels[batch.offsets[i]])
            if left > 0 and first % 2 == 1:
                flags[i] = 1
                left -= 1
                continue
            tok[i, :4] = [2, batch.heights[i], batch.widths[i], first]
            ln[i] = 4
        self._flags = flags
        return tok, ln

    def unfinished(self, n):
        return self._flags[:n]

    def close(self):
        pass


@pytest.fixture
def fake(monkeypatch):
    FakeRecognizer.created = 0
    monkeypatch.setattr(_native, "Recognizer", FakeRecognizer)
    return FakeRecognizer


def _lines(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        h, w = int(rng.integers(20, 60)), int(rng.integers(60, 2000))
        im = np.full((h, w), 255, np.uint8)
        im[0, 0] = i % 251
        out.append(im)
    return out


@pytest.mark.parametrize("in_flight,n", [(1, 5), (4, 700), (12, 2000)])
def test_pipeline_returns_every_line_in_order(fake, in_flight, n):
    pipe = pipeline.LinePipeline(b"blob", in_flight=in_flight, max_lines=64, max_chunks=600, straggler_per_256=16)
    imgs = _lines(n, seed=n)
    tok, ln = pipe.recognize(imgs)
    assert np.all(ln == 4)
    for i, im in enumerate(imgs):
        assert list(tok[i, :4]) == [2, im.shape[0], im.shape[1], i % 251]
    assert pipe.stats["passes"] >= (n + 63) // 64
    if n > 64:
        assert pipe.stats["stragglers"] > 0 and pipe.stats["straggler_passes"] > 0      # the pool was exercised
    assert fake.created == min(in_flight, max(1, len(pipe.plan([im.shape for im in imgs]))))   # handles are created on demand


def test_pipeline_job_completion_callbacks(fake):
    pipe = pipeline.LinePipeline(b"blob", in_flight=3, max_lines=32, max_chunks=400, straggler_per_256=32)
    imgs = _lines(300, seed=5)
    plan = pipe.plan([im.shape for im in imgs])
    assert sorted(i for g in plan for i in g) == list(range(300))
    jobs = [pipeline.Job(ids, images=[imgs[i] for i in ids], tag=k) for k, ids in enumerate(plan)]
    tok = np.zeros((300, _native.TOKENS_LD), np.int32)
    ln = np.zeros(300, np.int32)
    done, lock = [], threading.Lock()

    def on_done(job):
        with lock:
            assert np.all(ln[job.ids] == 4), "a job is reported done only when all of its lines have their result"
            done.append(job.tag)

    pipe.run_jobs(jobs, tok, ln, image_of=lambda i: imgs[i], on_done=on_done)
    assert sorted(done) == list(range(len(jobs)))


def test_pipeline_empty_and_errors(fake):
    pipe = pipeline.LinePipeline(b"blob", in_flight=2, max_lines=8, max_chunks=64)
    tok, ln = pipe.recognize([])
    assert tok.shape == (0, _native.TOKENS_LD) and ln.shape == (0,)

    def boom(*a, **k):
        raise _native.KocrError("device error")
    pipe.recs[0].recognize_lines = boom
    with pytest.raises(_native.KocrError):
        pipe.recognize(_lines(3))
