"""tcgen05 GEMM unit tests through the C ABI: tcgen05 kernel vs the CUDA-core check kernel vs numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _a16():
    import torch
    from khmer_ocr_cnn_transformer_b200 import weights
    return torch.float16 if weights.A16_FORMAT == 1 else torch.bfloat16


def _run(impl, a, w, M, N, taps, cin, tap_off, bias, relu, pl):
    import torch
    from khmer_ocr_cnn_transformer_b200 import _native
    lib = _native.load_library()
    out32 = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    out16 = torch.zeros(M, N, dtype=_a16(), device="cuda")
    to = np.asarray(tap_off, np.int32)
    b = torch.from_numpy(bias).cuda() if bias is not None else None
    _native.check(lib.kocr_test_gemm(impl, a.data_ptr(), a.shape[0], w.data_ptr(), M, N, taps, cin,
                                     to.ctypes.data, b.data_ptr() if b is not None else None, relu,
                                     pl[0], pl[1], out32.data_ptr(), out16.data_ptr(), None))
    torch.cuda.synchronize()
    return out32.cpu().numpy(), out16.float().cpu().numpy()


def _reference(a, w, M, N, taps, cin, tap_off, bias, relu, pl):
    A = a.float().cpu().numpy().astype(np.float64)
    W = w.float().cpu().numpy().astype(np.float64).reshape(N, taps, cin)
    rows = A.shape[0]
    out = np.zeros((M, N))
    for t in range(taps):
        idx = np.arange(M) + tap_off[t]
        ok = (idx >= 0) & (idx < rows)
        sh = np.zeros((M, cin))
        sh[ok] = A[idx[ok]]
        out += sh @ W[:, t, :].T
    if bias is not None:
        out += bias
    if relu:
        out = np.maximum(out, 0)
    if pl[0] > 0:
        S, P = (pl[0] + 1) * (pl[1] + 1), pl[1] + 1
        r = np.arange(M) % S
        valid = ((r // P) < pl[0]) & ((r % P) < pl[1])
        out[~valid] = 0
    return out


CASES = [
    # M, N, taps, cin, relu, pl
    (128, 128, 1, 64, 0, (0, 0)),          # one tile, one k-block
    (128, 128, 1, 384, 0, (0, 0)),         # k loop wraps the 6-stage ring once
    (300, 384, 1, 1024, 1, (0, 0)),        # partial M tile, 3 N tiles, ring wraps many times
    (1000, 256, 1, 384, 0, (0, 0)),        # BN=256 path
    (2 * 338, 256, 9, 128, 1, (12, 25)),   # conv3-like implicit GEMM with row mask
    (5 * 104, 512, 9, 512, 1, (3, 25)),    # conv7-like, K = 4608
    (37, 128, 1, 384, 0, (0, 0)),          # decode-step sized
    (700, 1152, 1, 384, 1, (0, 0)),        # N = 1152 = 9 x 128 (or 6 x 192 with the gemm_bn192 option)
    (148 * 128 * 2 + 77, 128, 1, 64, 0, (0, 0)),   # persistent loop: several tiles per CTA
]


@pytest.mark.parametrize("M,N,taps,cin,relu,pl", CASES)
def test_gemm_tcgen05_matches_check_kernel_and_numpy(M, N, taps, cin, relu, pl):
    import torch
    rng = np.random.default_rng(M * 7 + N)
    rows = M
    a = torch.from_numpy(rng.standard_normal((rows, cin)).astype(np.float32)).cuda().to(_a16())
    w = torch.from_numpy((rng.standard_normal((N, taps * cin)) / np.sqrt(taps * cin)).astype(np.float32)).cuda().to(_a16())
    bias = rng.standard_normal(N).astype(np.float32)
    if taps == 9:
        P = pl[1] + 1
        tap_off = [(r - 1) * P + (s - 1) for r in range(3) for s in range(3)]
    else:
        tap_off = [0]
    ref = _reference(a, w, M, N, taps, cin, tap_off, bias, relu, pl)
    chk32, _ = _run(1, a, w, M, N, taps, cin, tap_off, bias, relu, pl)
    assert np.abs(chk32 - ref).max() < 2e-3, "CUDA-core check kernel disagrees with numpy"
    tc32, tc16 = _run(0, a, w, M, N, taps, cin, tap_off, bias, relu, pl)
    err = np.abs(tc32 - ref).max()
    assert err < 2e-3, f"tcgen05 fp32 output max err {err}"
    # the 16-bit output is the fp32 result rounded to nearest-even
    want16 = torch.from_numpy(tc32).to(_a16()).float().numpy()
    assert np.array_equal(tc16, want16)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 1152, 384), (37, 128, 384), (256, 384, 1536), (300, 1536, 384)])
def test_gemm_tf32_operands(M, N, K):
    """kind::tf32 path used by the decoder: fp32 operands in global memory, read by the tensor core as TF32."""
    import torch
    from khmer_ocr_cnn_transformer_b200 import _native
    from khmer_ocr_cnn_transformer_b200.weights import round_to_tf32
    lib = _native.load_library()
    rng = np.random.default_rng(M + N + K)
    a_np = round_to_tf32(rng.standard_normal((M, K)).astype(np.float32))
    w_np = round_to_tf32((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    a, w, b = torch.from_numpy(a_np).cuda(), torch.from_numpy(w_np).cuda(), torch.from_numpy(bias).cuda()
    out = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    to = np.zeros(1, np.int32)
    _native.check(lib.kocr_test_gemm(2, a.data_ptr(), M, w.data_ptr(), M, N, 1, K, to.ctypes.data, b.data_ptr(), 0,
                                     0, 0, out.data_ptr(), None, None))
    torch.cuda.synchronize()
    ref = a_np.astype(np.float64) @ w_np.astype(np.float64).T + bias
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err < 1e-4, f"tf32 GEMM with tf32-representable operands: max err {err}"
    # un-rounded fp32 activations: the hardware drops the low mantissa bits (error ~2^-11 relative per element)
    a2_np = rng.standard_normal((M, K)).astype(np.float32)
    a2 = torch.from_numpy(a2_np).cuda()
    _native.check(lib.kocr_test_gemm(2, a2.data_ptr(), M, w.data_ptr(), M, N, 1, K, to.ctypes.data, b.data_ptr(), 0,
                                     0, 0, out.data_ptr(), None, None))
    torch.cuda.synchronize()
    ref2 = a2_np.astype(np.float64) @ w_np.astype(np.float64).T + bias
    rel = np.linalg.norm(out.cpu().numpy() - ref2) / np.linalg.norm(ref2)
    assert rel < 1e-3, f"tf32 GEMM with fp32 activations: relative error {rel}"
