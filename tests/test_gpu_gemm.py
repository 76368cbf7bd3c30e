"""tcgen05 GEMM unit tests through the C ABI: tcgen05 kernel vs the CUDA-core check kernel vs numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _a16():
    import torch
    from khmer_ocr_cnn_transformer_b200 import weights
    return torch.float16 if weights.A16_FORMAT == 1 else torch.bfloat16


def _run(impl, a, w, M, N, taps, cin, bias, relu, conv=(0, 0), tile_cols=0, col_mode=0):
    """One launch through kocr_test_gemm.  Returns (fp32 out, 16-bit out) or, with col_mode, (pooled 16-bit, column means)."""
    import torch
    from khmer_ocr_cnn_transformer_b200 import _native
    lib = _native.load_library()
    b = torch.from_numpy(bias).cuda() if bias is not None else None
    H, W = conv
    if col_mode:
        cols = M // H
        rows_out = {1: cols * (H // 2), 2: cols * 2, 3: (cols // 2) * (H // 2)}[col_mode]
        pool = torch.zeros(rows_out, N, dtype=_a16(), device="cuda")
        mean = torch.zeros(cols, N, dtype=_a16(), device="cuda")
        _native.check(lib.kocr_test_gemm(impl, a.data_ptr(), a.shape[0], w.data_ptr(), M, N, taps, cin, H, W, tile_cols, col_mode,
                                         b.data_ptr() if b is not None else None, relu, None, None, pool.data_ptr(),
                                         mean.data_ptr(), None))
        torch.cuda.synchronize()
        return pool.float().cpu().numpy(), mean.float().cpu().numpy()
    out32 = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    out16 = torch.zeros(M, N, dtype=_a16(), device="cuda")
    _native.check(lib.kocr_test_gemm(impl, a.data_ptr(), a.shape[0], w.data_ptr(), M, N, taps, cin, H, W, tile_cols, 0,
                                     b.data_ptr() if b is not None else None, relu, out32.data_ptr(), out16.data_ptr(),
                                     None, None, None))
    torch.cuda.synchronize()
    return out32.cpu().numpy(), out16.float().cpu().numpy()


def _reference(a, w, M, N, taps, cin, bias, relu, conv=(0, 0)):
    """float64 numpy: plain GEMM, or the 3x3 / pad-1 convolution over the dense NWHC activation [n][W][H][cin]."""
    A = a.float().cpu().numpy().astype(np.float64)
    Wt = w.float().cpu().numpy().astype(np.float64).reshape(N, taps, cin)
    if taps == 1:
        out = A[:M] @ Wt[:, 0, :].T
    else:
        H, W_ = conv
        n = M // (H * W_)
        x = A.reshape(n, W_, H, cin)
        xp = np.zeros((n, W_ + 2, H + 2, cin))
        xp[:, 1:-1, 1:-1] = x
        out = np.zeros((n, W_, H, N))
        for kh in range(3):
            for kw in range(3):
                out += xp[:, kw:kw + W_, kh:kh + H] @ Wt[:, kh * 3 + kw, :].T
        out = out.reshape(M, N)
    if bias is not None:
        out = out + bias
    if relu:
        out = np.maximum(out, 0)
    return out


CASES = [
    # M, N, taps, cin, relu, (H, W), tile_cols
    (128, 128, 1, 64, 0, (0, 0), 0),            # one tile, one k-block
    (128, 128, 1, 384, 0, (0, 0), 0),           # k loop wraps the operand ring once
    (300, 384, 1, 1024, 1, (0, 0), 0),          # partial M tile, 3 N tiles, ring wraps many times
    (1000, 256, 1, 384, 0, (0, 0), 0),          # BN=256 path
    (2 * 300, 256, 9, 128, 1, (12, 25), 0),     # conv3-like implicit GEMM (TMA im2col), 128 consecutive pixels per tile
    (3 * 1200, 128, 9, 64, 1, (24, 50), 0),     # conv2-like: tiles straddle columns and images
    (5 * 75, 512, 9, 512, 1, (3, 25), 0),       # conv7-like, K = 4608
    (5 * 150, 512, 9, 256, 1, (6, 25), 21),     # conv5-like with whole-column tiles (126 rows), standard epilogue
    (37, 128, 1, 384, 0, (0, 0), 0),            # decode-step sized
    (700, 1152, 1, 384, 1, (0, 0), 0),          # N = 1152 = 9 x 128 (or 6 x 192 with the gemm_bn192 option)
    (148 * 128 * 2 + 77, 128, 1, 64, 0, (0, 0), 0),   # persistent loop: several tiles per CTA
]


@pytest.mark.parametrize("M,N,taps,cin,relu,conv,tile_cols", CASES)
def test_gemm_tcgen05_matches_check_kernel_and_numpy(M, N, taps, cin, relu, conv, tile_cols):
    import torch
    rng = np.random.default_rng(M * 7 + N)
    a = torch.from_numpy(rng.standard_normal((M, cin)).astype(np.float32)).cuda().to(_a16())
    w = torch.from_numpy((rng.standard_normal((N, taps * cin)) / np.sqrt(taps * cin)).astype(np.float32)).cuda().to(_a16())
    bias = rng.standard_normal(N).astype(np.float32)
    ref = _reference(a, w, M, N, taps, cin, bias, relu, conv)
    chk32, _ = _run(1, a, w, M, N, taps, cin, bias, relu, conv, tile_cols)
    assert np.abs(chk32 - ref).max() < 2e-3, "CUDA-core check kernel disagrees with numpy"
    tc32, tc16 = _run(0, a, w, M, N, taps, cin, bias, relu, conv, tile_cols)
    err = np.abs(tc32 - ref).max()
    assert err < 2e-3, f"tcgen05 fp32 output max err {err}"
    # the 16-bit output is the fp32 result rounded to nearest-even
    want16 = torch.from_numpy(tc32).to(_a16()).float().numpy()
    assert np.array_equal(tc16, want16)


COL_CASES = [
    # n_img, (H, W), tile_cols, cin, N, col_mode, relu
    (5, (12, 25), 10, 256, 256, 1, 1),          # conv4: 10 columns x 12 rows per tile, (2,1) max-pool + column means
    (7, (6, 25), 21, 512, 512, 1, 1),           # conv6: 21 columns x 6 rows (126 of 128 MMA rows), tiles straddle images
    (9, (3, 25), 42, 512, 512, 2, 1),           # conv7 (SE model): 42 columns x 3 rows, adaptive-pool row-bin sums
    (4, (3, 25), 42, 512, 512, 2, 0),           # conv7 of the VGG baseline: no ReLU
    (3, (24, 50), 4, 64, 128, 3, 1),            # conv2: 4 columns x 24 rows per tile, 2x2 max-pool (N tile 128)
]


@pytest.mark.parametrize("n_img,conv,tile_cols,cin,N,col_mode,relu", COL_CASES)
def test_gemm_column_fused_epilogue(n_img, conv, tile_cols, cin, N, col_mode, relu):
    """Column-fused conv epilogue (whole-column M tiles): pooled rows / row-bin sums and SE column means straight from the
    fp32 accumulators, against numpy and against the CUDA-core check kernel."""
    import torch
    H, W = conv
    M = n_img * H * W
    rng = np.random.default_rng(M + N + col_mode)
    a = torch.from_numpy(rng.standard_normal((M, cin)).astype(np.float32)).cuda().to(_a16())
    w = torch.from_numpy((rng.standard_normal((N, 9 * cin)) / np.sqrt(9 * cin)).astype(np.float32)).cuda().to(_a16())
    bias = rng.standard_normal(N).astype(np.float32)
    y = _reference(a, w, M, N, 9, cin, bias, relu, conv).reshape(n_img * W, H, N)          # [col][h][n]
    want_mean = y.mean(axis=1)
    if col_mode == 1:
        want_pool = np.maximum(y[:, 0::2], y[:, 1::2]).reshape(-1, N)
    elif col_mode == 2:
        want_pool = np.stack([y[:, 0] + y[:, 1], y[:, 1] + y[:, 2]], axis=1).reshape(-1, N)
    else:
        want_pool = y.reshape(n_img * W // 2, 2, H // 2, 2, N).max(axis=(1, 3)).reshape(-1, N)
    ulp = 2.0 ** -10 if _a16() == torch.float16 else 2.0 ** -7
    for impl in (1, 0):
        pool, mean = _run(impl, a, w, M, N, 9, cin, bias, relu, conv, tile_cols, col_mode)
        if col_mode != 3:       # (the 2x2 pool after conv2 has no SE block behind it: no column means)
            assert np.all(np.abs(mean - want_mean) <= 2e-3 + np.abs(want_mean) * ulp), f"impl {impl}: column means"
        # pooled values are rounded to 16 bits once: compare within one 16-bit ulp of the fp64 result
        tol = 2e-3 + np.abs(want_pool) * ulp
        assert np.all(np.abs(pool - want_pool) <= tol), f"impl {impl}: pooled rows, max err {np.abs(pool - want_pool).max()}"


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
    _native.check(lib.kocr_test_gemm(2, a.data_ptr(), M, w.data_ptr(), M, N, 1, K, 0, 0, 0, 0, b.data_ptr(), 0,
                                     out.data_ptr(), None, None, None, None))
    torch.cuda.synchronize()
    ref = a_np.astype(np.float64) @ w_np.astype(np.float64).T + bias
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err < 1e-4, f"tf32 GEMM with tf32-representable operands: max err {err}"
    # un-rounded fp32 activations: the hardware drops the low mantissa bits (error ~2^-11 relative per element)
    a2_np = rng.standard_normal((M, K)).astype(np.float32)
    a2 = torch.from_numpy(a2_np).cuda()
    _native.check(lib.kocr_test_gemm(2, a2.data_ptr(), M, w.data_ptr(), M, N, 1, K, 0, 0, 0, 0, b.data_ptr(), 0,
                                     out.data_ptr(), None, None, None, None))
    torch.cuda.synchronize()
    ref2 = a2_np.astype(np.float64) @ w_np.astype(np.float64).T + bias
    rel = np.linalg.norm(out.cpu().numpy() - ref2) / np.linalg.norm(ref2)
    assert rel < 1e-3, f"tf32 GEMM with fp32 activations: relative error {rel}"
