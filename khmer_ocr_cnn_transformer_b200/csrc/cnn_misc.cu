// CUDA-core / HBM-bound pieces of the SE-VGG backbone (everything that is not a dense contraction):
//   conv1 (Cin = 1, too thin for tensor cores) + BN + ReLU + 2x2 pool      se_model.py:39-40,64
//   2x2 max-pool after conv2                                                 se_model.py:43,65
//   SequenceSE gate (mean over H -> FC -> ReLU -> FC -> sigmoid) + (2,1) pool  se_model.py:19-30,48-49,53-54
//   SequenceSE gate + AdaptiveAvgPool2d((2,32)) -> patch-projection operand  se_model.py:59-61,76-78
// All activations are bf16 in the padded-linear NHWC layout (common.cuh PLGeom); SE math is fp32.
#include "kernels.cuh"

namespace kocr {

// ------------------------------------------------------------------------------------------
// conv1 + pool1.  grid = (4 bands of 6 pooled rows, n_chunks), block = 256.
// ------------------------------------------------------------------------------------------
static constexpr int C1_BAND = 6;                 // pooled rows per CTA
static constexpr int C1_IN_ROWS = 2 * C1_BAND + 2;
static constexpr int C1_IN_COLS = CHUNK_W + 2;

__global__ void __launch_bounds__(256) conv1_pool_kernel(const float* __restrict__ chunks,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         __nv_bfloat16* __restrict__ out) {
    __shared__ float s_in[C1_IN_ROWS][C1_IN_COLS];
    __shared__ float s_w[9][64];
    __shared__ float s_b[64];
    const int band = blockIdx.x, n = blockIdx.y;
    const PLGeom g = make_pl(IMG_H / 2, CHUNK_W / 2);     // 24 x 50 output
    const float* src = chunks + (long)n * IMG_H * CHUNK_W;
    const int y0 = band * 2 * C1_BAND - 1;                // first input row (with halo)
    for (int i = threadIdx.x; i < C1_IN_ROWS * C1_IN_COLS; i += blockDim.x) {
        const int r = i / C1_IN_COLS, c = i - r * C1_IN_COLS;
        const int y = y0 + r, x = c - 1;
        s_in[r][c] = (y >= 0 && y < IMG_H && x >= 0 && x < CHUNK_W) ? src[y * CHUNK_W + x] : 0.f;
    }
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) s_w[i / 64][i % 64] = w[(i % 64) * 9 + i / 64];
    if (threadIdx.x < 64) s_b[threadIdx.x] = b[threadIdx.x];
    __syncthreads();

    const int rows_here = C1_BAND + (band == 3 ? 1 : 0);   // last band also writes the shared pad row
    const int items = rows_here * g.P * 8;
    // 256 % 8 == 0: a thread always works on the same group of 8 output channels, so its 72 folded weights and
    // 8 biases live in registers for all of its pixels (shared memory then only serves the 4x4 input patches).
    const int cg = threadIdx.x & 7;
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[t][j] = s_w[t][cg * 8 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = s_b[cg * 8 + j];
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int pos = it >> 3;
        const int pr = pos / g.P, pw = pos - pr * g.P;
        const int oh = band * C1_BAND + pr;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (oh < g.H && pw < g.W) {
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) in[r][c] = s_in[2 * pr + r][2 * pw + c];
            float res[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float wv = wr[r * 3 + c][j];
                        a00 = fmaf(in[r][c], wv, a00);
                        a01 = fmaf(in[r][c + 1], wv, a01);
                        a10 = fmaf(in[r + 1][c], wv, a10);
                        a11 = fmaf(in[r + 1][c + 1], wv, a11);
                    }
                res[j] = fmaxf(fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + br[j], 0.f);
            }
            o = make_uint4(pack_bf16(res[0], res[1]), pack_bf16(res[2], res[3]), pack_bf16(res[4], res[5]),
                           pack_bf16(res[6], res[7]));
        }
        const long q = (long)n * g.S + (long)oh * g.P + pw;
        reinterpret_cast<uint4*>(out + q * 64)[cg] = o;
    }
}

int launch_conv1_pool(const float* d_chunks, const float* w, const float* b, __nv_bfloat16* out, int n_chunks,
                      cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    conv1_pool_kernel<<<dim3(4, n_chunks), 256, 0, stream>>>(d_chunks, w, b, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// 2x2 max-pool between padded-linear layouts.  One thread per (output position, 8 channels).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max4(uint4 a, uint4 b) {
    return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y), bf16x2_max(a.z, b.z), bf16x2_max(a.w, b.w));
}

__global__ void __launch_bounds__(256) pool2x2_kernel(const __nv_bfloat16* __restrict__ in,
                                                      __nv_bfloat16* __restrict__ out, long total, int H, int W,
                                                      int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const PLGeom gi = make_pl(H, W), go = make_pl(H / 2, W / 2);
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const long q = idx / cgs;
    const int n = (int)(q / go.S);
    const int r = (int)(q - (long)n * go.S);
    const int oh = r / go.P, ow = r - oh * go.P;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (oh < go.H && ow < go.W) {
        const long q00 = (long)n * gi.S + (long)(2 * oh) * gi.P + 2 * ow;
        const uint4* p = reinterpret_cast<const uint4*>(in);
        const uint4 a = p[q00 * cgs + cg], b = p[(q00 + 1) * cgs + cg];
        const uint4 c = p[(q00 + gi.P) * cgs + cg], d = p[(q00 + gi.P + 1) * cgs + cg];
        o = max4(max4(a, b), max4(c, d));
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

int launch_pool2x2(const __nv_bfloat16* in, __nv_bfloat16* out, int n_chunks, int H, int W, int C,
                   cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const PLGeom go = make_pl(H / 2, W / 2);
    const long total = (long)n_chunks * go.S * (C / 8);
    pool2x2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// 1D-SE (SequenceSE, se_model.py:8-30).  The squeeze (mean over H) and the gated pooling are
// HBM-bound elementwise kernels; the two 1x1 Conv1d layers of the excitation are real contractions
// over channels and run on the tcgen05 GEMM (reduced width padded to 128), see kocr_api.cu.
// ------------------------------------------------------------------------------------------
// squeeze: padded-linear (H, W, C) bf16 -> column means [n*W + w][C] bf16.  Thread per (n, w, 8 channels).
__global__ void __launch_bounds__(256) se_col_mean_kernel(const __nv_bfloat16* __restrict__ in,
                                                          __nv_bfloat16* __restrict__ means, long total, int H,
                                                          int W, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const PLGeom gi = make_pl(H, W);
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const long col = idx / cgs;                 // n*W + w
    const int n = (int)(col / W), w = (int)(col - (long)n * W);
    const uint4* p = reinterpret_cast<const uint4*>(in + ((long)n * gi.S + w) * C) + cg;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < H; ++h) {
        const uint4 a = __ldg(p + (long)h * gi.P * cgs);
        acc[0] += bf16_lo(a.x); acc[1] += bf16_hi(a.x); acc[2] += bf16_lo(a.y); acc[3] += bf16_hi(a.y);
        acc[4] += bf16_lo(a.z); acc[5] += bf16_hi(a.z); acc[6] += bf16_lo(a.w); acc[7] += bf16_hi(a.w);
    }
    const float inv = 1.f / (float)H;
    reinterpret_cast<uint4*>(means)[idx] =
        make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                   pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
}

int launch_se_col_mean(const __nv_bfloat16* in, __nv_bfloat16* means, int n_chunks, int H, int W, int C,
                       cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const long total = (long)n_chunks * W * (C / 8);
    se_col_mean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, means, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

__device__ __forceinline__ void load_gate8(const float* __restrict__ gate, long off, float (&g)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gate + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(gate + off) + 1);
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
}

// gate (optional, fp32 [n*W + w][C], already sigmoid-ed) * max over row pairs -> padded-linear (H/2, W, C).
// The gate is positive, so max(x1*g, x2*g) == g*max(x1, x2) exactly (se_model.py:30,49).
__global__ void __launch_bounds__(256) se_apply_pool_kernel(const __nv_bfloat16* __restrict__ in,
                                                            const float* __restrict__ gate,
                                                            __nv_bfloat16* __restrict__ out, long total, int H,
                                                            int W, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const PLGeom gi = make_pl(H, W), go = make_pl(H / 2, W);
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const long q = idx / cgs;
    const int n = (int)(q / go.S);
    const int r = (int)(q - (long)n * go.S);
    const int oh = r / go.P, ow = r - oh * go.P;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (oh < go.H && ow < go.W) {
        const uint4* p = reinterpret_cast<const uint4*>(in + ((long)n * gi.S + (long)(2 * oh) * gi.P + ow) * C) + cg;
        o = max4(__ldg(p), __ldg(p + (long)gi.P * cgs));
        if (gate) {
            float g[8];
            load_gate8(gate, ((long)n * W + ow) * C + cg * 8, g);
            o = make_uint4(pack_bf16(bf16_lo(o.x) * g[0], bf16_hi(o.x) * g[1]),
                           pack_bf16(bf16_lo(o.y) * g[2], bf16_hi(o.y) * g[3]),
                           pack_bf16(bf16_lo(o.z) * g[4], bf16_hi(o.z) * g[5]),
                           pack_bf16(bf16_lo(o.w) * g[6], bf16_hi(o.w) * g[7]));
        }
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

int launch_se_apply_pool(const __nv_bfloat16* in, const float* gate, __nv_bfloat16* out, int n_chunks, int H, int W,
                         int C, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const PLGeom go = make_pl(H / 2, W);
    const long total = (long)n_chunks * go.S * (C / 8);
    se_apply_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, gate, out, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// gate (optional) * x, then AdaptiveAvgPool2d((2, 32)) -> patch-projection operand
// out[n*32 + k][kh*C + c] (se_model.py:61,78: bin k covers columns [floor(k*W/32), ceil((k+1)*W/32))).
__global__ void __launch_bounds__(256) se_apply_finalpool_kernel(const __nv_bfloat16* __restrict__ in,
                                                                 const float* __restrict__ gate,
                                                                 __nv_bfloat16* __restrict__ out, long total, int H,
                                                                 int W, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const PLGeom gi = make_pl(H, W);
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const int kh = (int)((idx / cgs) & 1);
    const long nk = idx / (2 * cgs);                 // n*32 + k
    const int n = (int)(nk / TOK_PER_CHUNK), k = (int)(nk - (long)n * TOK_PER_CHUNK);
    const int h0 = (kh * H) / 2, h1 = ((kh + 1) * H + 1) / 2;
    const int w0 = (k * W) / TOK_PER_CHUNK, w1 = ((k + 1) * W + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int w = w0; w < w1; ++w) {
        float g[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (gate) load_gate8(gate, ((long)n * W + w) * C + cg * 8, g);
        for (int h = h0; h < h1; ++h) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(in + ((long)n * gi.S + (long)h * gi.P + w) * C) + cg);
            acc[0] += bf16_lo(a.x) * g[0]; acc[1] += bf16_hi(a.x) * g[1];
            acc[2] += bf16_lo(a.y) * g[2]; acc[3] += bf16_hi(a.y) * g[3];
            acc[4] += bf16_lo(a.z) * g[4]; acc[5] += bf16_hi(a.z) * g[5];
            acc[6] += bf16_lo(a.w) * g[6]; acc[7] += bf16_hi(a.w) * g[7];
        }
    }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
    reinterpret_cast<uint4*>(out)[idx] =
        make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                   pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
}

int launch_se_apply_finalpool(const __nv_bfloat16* in, const float* gate, __nv_bfloat16* out, int n_chunks, int H,
                              int W, int C, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const long total = (long)n_chunks * TOK_PER_CHUNK * 2 * (C / 8);
    se_apply_finalpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, gate, out, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace kocr
