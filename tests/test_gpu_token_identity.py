"""Greedy-decoded token sequences of the CUDA path against the CPU oracle on WHOLE workloads (north-star bar: identical on
>= 99.9 % of lines).  The oracle's tokens are committed fixtures (tests/golden/oracle_tokens_{c2,c3,c3full}.npz, written by
`tests/parity/parity_workloads.py oracle <w>` - the numpy restatement of predictor.py:85-99 / se_model.py:182-208 on the
fixture checkpoint; the 8192-line set takes it 66 minutes on 8 cores), so the GPU side is a few seconds per test.

  c2      256 lines, resized width 400-800, seed 0 (the bench batch)      -> every line identical
  c3      first 1024 lines of the mixed-width config (200-1600, seed 3)   -> >= 99.9 %
  c3full  all 8192 lines of BASELINE config c3                            -> >= 99.9 % (at most 8 lines may differ)
"""
import numpy as np
import pytest

from helpers import GOLDEN
from test_gpu_stages import _report

pytestmark = pytest.mark.gpu

SPEC = {"c2": (256, 400, 800, 0), "c3": (1024, 200, 1600, 3), "c3full": (8192, 200, 1600, 3)}


@pytest.fixture(scope="module")
def rec():
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
    r = _native.Recognizer(weights.pack_blob(load_checkpoint(GOLDEN / "fixture_se_ckpt.npz")), max_lines=256, max_chunks=5120)
    yield r
    r.close()


def _identity(rec, workload):
    from khmer_ocr_cnn_transformer_b200 import _native
    from workloads import synth
    n, lo, hi, seed = SPEC[workload]
    path = GOLDEN / f"oracle_tokens_{workload}.npz"
    if not path.exists():
        pytest.skip(f"{path.name} not generated")
    o = np.load(path)
    assert o["tokens"].shape[0] == n
    imgs, _ = synth.make_lines(n, lo, hi, seed=seed)
    rec.set_option("straggler_threshold", 0)
    diff = []
    for i0 in range(0, n, 256):
        tok, ln = rec.recognize_lines(_native.LineBatch(imgs[i0:i0 + 256]))
        for j in range(tok.shape[0]):
            i = i0 + j
            if ln[j] != o["lengths"][i] or not np.array_equal(tok[j, :ln[j]], o["tokens"][i, :ln[j]]):
                a, b = tok[j, :ln[j]], o["tokens"][i, :o["lengths"][i]]
                k = next((p for p, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
                diff.append({"line": i, "first_diff_pos": int(k), "oracle_top1_top2_gap": float(o["gaps"][i, max(k - 1, 0)])})
    same = n - len(diff)
    _report(f"token_identity_{workload}", {"identical": same, "of": n, "rate": same / n, "mismatches": diff})
    print(f"\ntoken identity {workload}: {same} / {n} = {100.0 * same / n:.3f} %  mismatching lines: {[d['line'] for d in diff]}")
    return same, n


def test_c2_all_256_lines_identical_to_oracle(rec):
    same, n = _identity(rec, "c2")
    assert same == n


def test_c3_first_1024_lines_identity_vs_oracle(rec):
    same, n = _identity(rec, "c3")
    assert same / n >= 0.999


def test_c3_all_8192_lines_identity_vs_oracle(rec):
    same, n = _identity(rec, "c3full")
    assert same / n >= 0.999
