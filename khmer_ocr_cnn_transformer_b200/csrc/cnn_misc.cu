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
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int cg = it & 7;
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
                const int ch = cg * 8 + j;
                float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float wv = s_w[r * 3 + c][ch];
                        a00 = fmaf(in[r][c], wv, a00);
                        a01 = fmaf(in[r][c + 1], wv, a01);
                        a10 = fmaf(in[r + 1][c], wv, a10);
                        a11 = fmaf(in[r + 1][c + 1], wv, a11);
                    }
                res[j] = fmaxf(fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + s_b[ch], 0.f);
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
// SE gate + pooling.  One CTA (256 threads) per chunk; smem: gate/mean [W][C] f32 + z [W][R].
// FINAL == false: out = padded-linear (H/2, W, C) of gate * max over row pairs.
// FINAL == true : out = [n*32 + k][kh*C + c] = adaptive average over (rows kh..kh+1, column bin k).
// ------------------------------------------------------------------------------------------
template <bool FINAL>
__global__ void __launch_bounds__(256) se_pool_kernel(const __nv_bfloat16* __restrict__ in,
                                                      __nv_bfloat16* __restrict__ out, int H, int W, int C,
                                                      SEWeights se, int use_se) {
    extern __shared__ float s_dyn[];
    float* s_gate = s_dyn;                 // [W][C]   (column means first, then the gate)
    float* s_z = s_dyn + W * C;            // [W][R]
    const int n = blockIdx.x;
    const PLGeom gi = make_pl(H, W);
    const __nv_bfloat16* src = in + (long)n * gi.S * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (use_se) {
        // (1) squeeze: mean over H per (column, channel)            se_model.py:23
        const int cp = C / 2;
        for (int it = tid; it < W * cp; it += blockDim.x) {
            const int w = it / cp, c2 = it - w * cp;
            float s0 = 0.f, s1 = 0.f;
            for (int h = 0; h < H; ++h) {
                const uint32_t v = reinterpret_cast<const uint32_t*>(src + ((long)h * gi.P + w) * C)[c2];
                s0 += bf16_lo(v); s1 += bf16_hi(v);
            }
            s_gate[w * C + 2 * c2] = s0 / (float)H;
            s_gate[w * C + 2 * c2 + 1] = s1 / (float)H;
        }
        __syncthreads();
        // (2) excite FC1 + ReLU: warp owns reduced channel r, weights in registers, shuffle-reduce
        const int R = se.R;
        for (int r = warp; r < R; r += 8) {
            float wr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) wr[j] = (lane + 32 * j) < C ? se.w0[(long)r * C + lane + 32 * j] : 0.f;
            const float br = se.b0[r];
            for (int w = 0; w < W; ++w) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (lane + 32 * j < C) acc = fmaf(wr[j], s_gate[w * C + lane + 32 * j], acc);
                acc = warp_sum(acc);
                if (lane == 0) s_z[w * R + r] = fmaxf(acc + br, 0.f);
            }
        }
        __syncthreads();
        // (3) FC2 + sigmoid: thread owns channel c, loops over columns
        for (int c = tid; c < C; c += blockDim.x) {
            float w2[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) w2[r] = r < R ? se.w2[(long)c * R + r] : 0.f;
            const float b2 = se.b2[c];
            for (int w = 0; w < W; ++w) {
                float acc = b2;
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    if (r < R) acc = fmaf(w2[r], s_z[w * R + r], acc);
                s_gate[w * C + c] = 1.f / (1.f + __expf(-acc));
            }
        }
        __syncthreads();
    }

    const int cgs = C / 8;
    if (!FINAL) {
        // (4a) gate * max over the row pair, write padded-linear (H/2, W, C) including zero pads
        const PLGeom go = make_pl(H / 2, W);
        uint4* dst = reinterpret_cast<uint4*>(out + (long)n * go.S * C);
        for (int it = tid; it < go.S * cgs; it += blockDim.x) {
            const int cg = it % cgs, r = it / cgs;
            const int oh = r / go.P, ow = r - oh * go.P;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (oh < go.H && ow < go.W) {
                const uint4 a = reinterpret_cast<const uint4*>(src + ((long)(2 * oh) * gi.P + ow) * C)[cg];
                const uint4 b = reinterpret_cast<const uint4*>(src + ((long)(2 * oh + 1) * gi.P + ow) * C)[cg];
                o = max4(a, b);
                if (use_se) {
                    const float* gp = s_gate + ow * C + cg * 8;
                    o = make_uint4(pack_bf16(bf16_lo(o.x) * gp[0], bf16_hi(o.x) * gp[1]),
                                   pack_bf16(bf16_lo(o.y) * gp[2], bf16_hi(o.y) * gp[3]),
                                   pack_bf16(bf16_lo(o.z) * gp[4], bf16_hi(o.z) * gp[5]),
                                   pack_bf16(bf16_lo(o.w) * gp[6], bf16_hi(o.w) * gp[7]));
                }
            }
            dst[it] = o;
        }
    } else {
        // (4b) AdaptiveAvgPool2d((2, 32)) over gate * x; rows [kh*(H)/2 .. ) generalised bins
        __nv_bfloat16* dst = out + (long)n * TOK_PER_CHUNK * 2 * C;
        for (int it = tid; it < TOK_PER_CHUNK * 2 * cgs; it += blockDim.x) {
            const int cg = it % cgs;
            const int kh = (it / cgs) & 1;
            const int k = it / (2 * cgs);
            const int h0 = (kh * H) / 2, h1 = ((kh + 1) * H + 1) / 2;
            const int w0 = (k * W) / TOK_PER_CHUNK, w1 = ((k + 1) * W + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) {
                    const uint4 a = reinterpret_cast<const uint4*>(src + ((long)h * gi.P + w) * C)[cg];
                    float x[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                                  bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
                    if (use_se) {
                        const float* gp = s_gate + w * C + cg * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] *= gp[j];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += x[j];
                }
            const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
            const uint4 o = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                                       pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
            reinterpret_cast<uint4*>(dst + ((long)k * 2 + kh) * C)[cg] = o;
        }
    }
}

template <bool FINAL>
static int launch_se_generic(const __nv_bfloat16* in, __nv_bfloat16* out, int n_chunks, int H, int W, int C,
                             const SEWeights* se, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    KOCR_CHECK(C % 64 == 0 && C <= 512, "se_pool: unsupported channel count %d", C);
    KOCR_CHECK(se == nullptr || se->R <= 32, "se_pool: reduction width %d > 32", se ? se->R : 0);
    const size_t smem = (size_t)W * C * 4 + (size_t)W * 32 * 4;
    static bool attr_set = false;
    if (!attr_set) {
        KOCR_CUDA(cudaFuncSetAttribute(se_pool_kernel<FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    KOCR_CHECK(smem <= 64 * 1024, "se_pool: smem %zu too large", smem);
    SEWeights z = {nullptr, nullptr, nullptr, nullptr, 0};
    se_pool_kernel<FINAL><<<n_chunks, 256, smem, stream>>>(in, out, H, W, C, se ? *se : z, se ? 1 : 0);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

int launch_se_pool(const __nv_bfloat16* in, __nv_bfloat16* out, int n_chunks, int H, int W, int C,
                   const SEWeights* se, cudaStream_t stream) {
    return launch_se_generic<false>(in, out, n_chunks, H, W, C, se, stream);
}
int launch_se_finalpool(const __nv_bfloat16* in, __nv_bfloat16* out, int n_chunks, int H, int W, int C,
                        const SEWeights* se, cudaStream_t stream) {
    return launch_se_generic<true>(in, out, n_chunks, H, W, C, se, stream);
}

}  // namespace kocr
