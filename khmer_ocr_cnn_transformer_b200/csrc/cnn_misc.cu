// Pieces of the backbones that are not tcgen05 GEMMs:
//   conv1 (Cin = 1: K = 9) + BN + ReLU + 2x2 pool on mma.sync (CUDA-core version kept for A/B)   se_model.py:39-40,64
//   2x2 max-pool after conv2                                                                       se_model.py:43,65
//   fused SequenceSE block: mean over H -> FC -> ReLU -> FC -> sigmoid (mma.sync) -> gate * x ->
//     (2,1) max-pool or AdaptiveAvgPool2d((2,32)) (patch-projection operand)                       se_model.py:19-30,48-61,76-78
//   the four-kernel SE version (column means / apply+pool; the FCs then run on the GEMM) - VGG baseline pools and A/B tests
//   16-bit -> fp32 copy for the identity shortcuts of the ResNet baseline                           resnet_model.py:17,33
// All activations are 16-bit (act16_t) in the padded-linear NHWC layout (common.cuh PLGeom); SE math is fp32.
#include "kernels.cuh"

namespace kocr {

// ------------------------------------------------------------------------------------------
// conv1 + pool1 on the FP32 pipe (A/B reference of the tensor-core kernel below).  grid = (4 bands of 6 pooled rows, n_chunks), block = 256.
// ------------------------------------------------------------------------------------------
static constexpr int C1_BAND = 6;                 // pooled rows per CTA
static constexpr int C1_IN_ROWS = 2 * C1_BAND + 2;
static constexpr int C1_IN_COLS = CHUNK_W + 2;

__global__ void __launch_bounds__(256) conv1_pool_kernel(const float* __restrict__ chunks,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         act16_t* __restrict__ out) {
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
            o = make_uint4(pack_a16(res[0], res[1]), pack_a16(res[2], res[3]), pack_a16(res[4], res[5]),
                           pack_a16(res[6], res[7]));
        }
        const long q = (long)n * g.S + (long)oh * g.P + pw;
        reinterpret_cast<uint4*>(out + q * 64)[cg] = o;
    }
}

int launch_conv1_pool(const float* d_chunks, const float* w, const float* b, act16_t* out, int n_chunks,
                      cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    conv1_pool_kernel<<<dim3(4, n_chunks), 256, 0, stream>>>(d_chunks, w, b, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

__device__ __forceinline__ void mma_a16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." KOCR_MMA_A16 "." KOCR_MMA_A16 ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// conv1 + pool1 on the tensor cores.  The CUDA-core kernel above issues 9 FMAs per output value (10.4 GFLOP over the
// c2 batch = 0.32 ms of FP32 pipe); here one CTA per chunk keeps the chunk as a zero-bordered 16-bit tile in shared
// memory and runs the 3x3 conv as an implicit GEMM with K = 9 taps padded to 16 on mma.sync.m16n8k16
// (M = 16 pixels, N = 8 channels, fp32 accumulation; K = 9 is far too thin for a tcgen05 tile, and the kernel
// is bound by its 163 KB of output per chunk anyway).
// Row mapping of the two M-tiles of a "pool tile" (8 pooled pixels of one pooled row): tile A rows g / g+8 are the
// pixels (2py, 2px) / (2py, 2px+1), tile B the same on row 2py+1, px = 8i + g - so the four partners of a 2x2 pool
// window are the accumulators c0/c2 (c1/c3) of the two tiles of ONE thread: pooling needs no data exchange.
// bias and ReLU commute with the max.  Output: padded-linear (24, 50, 64), staged per warp in shared memory and
// written as one contiguous 1 KB run (8 pixels x 128 B) per pool tile.
// ------------------------------------------------------------------------------------------
static constexpr int C1M_LD = 104;                // halves per tile row: 102 used (x = -1 .. 100), 16-byte multiple

__global__ void __launch_bounds__(256) conv1_pool_mma_kernel(const float* __restrict__ chunks,
                                                             const act16_t* __restrict__ w16 /*[64][16], k = tap, 9..15 zero*/,
                                                             const float* __restrict__ b, act16_t* __restrict__ out) {
    __shared__ __align__(16) act16_t s_tile[(IMG_H + 2) * C1M_LD];
    __shared__ __align__(16) uint32_t s_stage[8][8][32];          // [warp][pooled pixel][64 channels as 32 words]
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const PLGeom go = make_pl(IMG_H / 2, CHUNK_W / 2);            // 24 x 50 output
    {   // zero-bordered 16-bit copy of the chunk: s_tile[(y + 1) * LD + (x + 1)]
        uint32_t* z = reinterpret_cast<uint32_t*>(s_tile);
        for (int i = tid; i < (IMG_H + 2) * C1M_LD / 2; i += 256) z[i] = 0u;
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(chunks + (long)n * IMG_H * CHUNK_W);
#pragma unroll 5
        for (int i = tid; i < IMG_H * (CHUNK_W / 4); i += 256) {
            const int y = i / (CHUNK_W / 4), x4 = i - y * (CHUNK_W / 4);
            const float4 v = __ldg(src + i);
            act16_t* d = s_tile + (y + 1) * C1M_LD + x4 * 4 + 1;
            d[0] = to_a16(v.x); d[1] = to_a16(v.y); d[2] = to_a16(v.z); d[3] = to_a16(v.w);
        }
    }
    // weights as B fragments (B[k = tap][n = channel] = w16[channel][tap]) and the bias of this thread's channels
    uint32_t bw[8][2];
    float bs[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const act16_t* wr = w16 + (nt * 8 + g) * 16 + 2 * t;
        bw[nt][0] = *reinterpret_cast<const uint32_t*>(wr);
        bw[nt][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
        bs[nt][0] = __ldg(b + nt * 8 + 2 * t);
        bs[nt][1] = __ldg(b + nt * 8 + 2 * t + 1);
    }
    __syncthreads();
    // taps of this thread's A registers: k = 2t, 2t + 1 (a0/a1) and k = 8 (a2/a3, t == 0 only); tap k = (k / 3, k % 3)
    const int k0 = 2 * t, k1 = 2 * t + 1;
    const int off0 = (k0 / 3) * C1M_LD + (k0 % 3), off1 = (k1 / 3) * C1M_LD + (k1 % 3), off8 = 2 * C1M_LD + 2;
    const unsigned short* tile16 = reinterpret_cast<const unsigned short*>(s_tile);
    constexpr int GROUPS = (CHUNK_W / 2 + 7) / 8;                 // 7 groups of 8 pooled columns per pooled row
    for (int pt = warp; pt < (IMG_H / 2) * GROUPS; pt += 8) {
        const int py = pt / GROUPS, gi = pt - py * GROUPS;
        const int px = min(gi * 8 + g, CHUNK_W / 2 - 1);          // clamped: results of columns >= 50 are never stored
        uint32_t a[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {                             // tile A: input row 2py, tile B: 2py + 1
            const unsigned short* p0 = tile16 + (2 * py + r) * C1M_LD + 2 * px;       // pixel (2py + r, 2px), tap (0, 0)
            a[r][0] = (uint32_t)p0[off0] | ((uint32_t)p0[off1] << 16);
            a[r][1] = (uint32_t)p0[off0 + 1] | ((uint32_t)p0[off1 + 1] << 16);        // pixel (2py + r, 2px + 1)
            a[r][2] = t == 0 ? (uint32_t)p0[off8] : 0u;
            a[r][3] = t == 0 ? (uint32_t)p0[off8 + 1] : 0u;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
            mma_a16_16816(ca, a[0], bw[nt][0], bw[nt][1]);
            mma_a16_16816(cb, a[1], bw[nt][0], bw[nt][1]);
            const float v0 = fmaxf(fmaxf(fmaxf(ca[0], ca[2]), fmaxf(cb[0], cb[2])) + bs[nt][0], 0.f);
            const float v1 = fmaxf(fmaxf(fmaxf(ca[1], ca[3]), fmaxf(cb[1], cb[3])) + bs[nt][1], 0.f);
            s_stage[warp][g][((nt ^ g) & 7) * 4 + t] = pack_a16(v0, v1);         // XOR swizzle: conflict-free both ways
        }
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(out + ((long)n * go.S + (long)py * go.P + gi * 8) * 64);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int idx = lane + 32 * r, x = idx >> 3, p = idx & 7;             // pooled pixel x of the tile, 16-byte piece p
            const int col = gi * 8 + x;
            const uint4 v = *reinterpret_cast<const uint4*>(&s_stage[warp][x][((p ^ x) & 7) * 4]);
            if (col < go.W) dst[idx] = v;
            else if (col == go.W) dst[idx] = make_uint4(0, 0, 0, 0);             // the shared zero pad column
        }
        __syncwarp();
    }
    if (warp == 0) {                                                             // the zero pad row below the chunk
        uint4* dst = reinterpret_cast<uint4*>(out + ((long)n * go.S + (long)go.H * go.P) * 64);
        for (int i = lane; i < go.P * 8; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
    }
}

static int g_conv1_impl = 1;           // 1: tensor-core kernel, 0: CUDA-core kernel (A/B tests)
void set_conv1_impl(int impl) { g_conv1_impl = impl; }

int launch_conv1_pool_mma(const float* d_chunks, const act16_t* w16, const float* b, act16_t* out, int n_chunks,
                          cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    conv1_pool_mma_kernel<<<n_chunks, 256, 0, stream>>>(d_chunks, w16, b, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}
int conv1_impl() { return g_conv1_impl; }

// ------------------------------------------------------------------------------------------
// 2x2 max-pool between padded-linear layouts.  One thread per (output position, 8 channels).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 max4(uint4 a, uint4 b) {
    return make_uint4(a16x2_max(a.x, b.x), a16x2_max(a.y, b.y), a16x2_max(a.z, b.z), a16x2_max(a.w, b.w));
}

__global__ void __launch_bounds__(256) pool2x2_kernel(const act16_t* __restrict__ in,
                                                      act16_t* __restrict__ out, long total, int H, int W,
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

int launch_pool2x2(const act16_t* in, act16_t* out, int n_chunks, int H, int W, int C,
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
// squeeze: padded-linear (H, W, C) a16 -> column means [n*W + w][C] a16.  Thread per (n, w, 8 channels).
__global__ void __launch_bounds__(256) se_col_mean_kernel(const act16_t* __restrict__ in,
                                                          act16_t* __restrict__ means, long total, int H,
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
        acc[0] += a16_lo(a.x); acc[1] += a16_hi(a.x); acc[2] += a16_lo(a.y); acc[3] += a16_hi(a.y);
        acc[4] += a16_lo(a.z); acc[5] += a16_hi(a.z); acc[6] += a16_lo(a.w); acc[7] += a16_hi(a.w);
    }
    const float inv = 1.f / (float)H;
    reinterpret_cast<uint4*>(means)[idx] =
        make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                   pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
}

int launch_se_col_mean(const act16_t* in, act16_t* means, int n_chunks, int H, int W, int C,
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
__global__ void __launch_bounds__(256) se_apply_pool_kernel(const act16_t* __restrict__ in,
                                                            const float* __restrict__ gate,
                                                            act16_t* __restrict__ out, long total, int H,
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
            o = make_uint4(pack_a16(a16_lo(o.x) * g[0], a16_hi(o.x) * g[1]),
                           pack_a16(a16_lo(o.y) * g[2], a16_hi(o.y) * g[3]),
                           pack_a16(a16_lo(o.z) * g[4], a16_hi(o.z) * g[5]),
                           pack_a16(a16_lo(o.w) * g[6], a16_hi(o.w) * g[7]));
        }
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

int launch_se_apply_pool(const act16_t* in, const float* gate, act16_t* out, int n_chunks, int H, int W,
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
__global__ void __launch_bounds__(256) se_apply_finalpool_kernel(const act16_t* __restrict__ in,
                                                                 const float* __restrict__ gate,
                                                                 act16_t* __restrict__ out, long total, int H,
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
            acc[0] += a16_lo(a.x) * g[0]; acc[1] += a16_hi(a.x) * g[1];
            acc[2] += a16_lo(a.y) * g[2]; acc[3] += a16_hi(a.y) * g[3];
            acc[4] += a16_lo(a.z) * g[4]; acc[5] += a16_hi(a.z) * g[5];
            acc[6] += a16_lo(a.w) * g[6]; acc[7] += a16_hi(a.w) * g[7];
        }
    }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
    reinterpret_cast<uint4*>(out)[idx] =
        make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                   pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
}

int launch_se_apply_finalpool(const act16_t* in, const float* gate, act16_t* out, int n_chunks, int H,
                              int W, int C, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const long total = (long)n_chunks * TOK_PER_CHUNK * 2 * (C / 8);
    se_apply_finalpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, gate, out, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Fused 1D-SE block (SequenceSE.forward, se_model.py:19-30, + the pooling that follows it, :69-78):
// ONE CTA per chunk does squeeze (column means) -> FC1+ReLU -> FC2+sigmoid -> gate * x -> pool.
// The chunk's activations (150 KB) are read twice by the same CTA, a few microseconds apart: the second read is an
// L2 hit, so HBM sees one read of the conv output and one write of the pooled output (225 KB per chunk instead of
// the ~580 KB of the four-kernel version with its a16 means / fp32 gate round trips).
// The two 1x1 Conv1d layers are [25 columns -> 32] x C x C/16 contractions: far below a tcgen05 tile, so they
// run on mma.sync.m16n8k16 (a16, fp32 accumulate) with the column means / hidden vector as A operands in shared
// memory and the weights read straight from L2 as B fragments.
// ------------------------------------------------------------------------------------------

// sigmoid with the approximate reciprocal (MUFU.RCP, 1 ulp): the IEEE division of `1.f / (1.f + __expf(-x))` expands to a
// branchy ~20-instruction sequence and was a third of this kernel's stall samples; the gate only scales 16-bit activations
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

static constexpr int SE_W = 25;            // columns of every SE stage (100 / 4)
static constexpr int SE_THREADS = 256;

template <int C> struct SeSmem {
    static constexpr int R = C / 16;                       // reduced width (se_model.py:9,13)
    static constexpr int LDA = C + 8;                      // a16 elements per row of the means (conflict-free frags)
    static constexpr int LDZ = R + 8;
    static constexpr size_t A_BYTES = 32 * LDA * 2;
    static constexpr size_t Z_BYTES = 32 * LDZ * 2;
    static constexpr size_t G_BYTES = (size_t)SE_W * C * 4;
    // the gate (written by FC2) re-uses the storage of the column means (dead once FC1 is done): 54 KB per CTA for
    // C = 512 instead of 87 KB -> 4 resident CTAs per SM, which is what hides the HBM latency of the two streaming phases
    static constexpr size_t BYTES = Z_BYTES + (A_BYTES > G_BYTES ? A_BYTES : G_BYTES);
};

// FINAL = false: (2,1) max-pool -> padded-linear (H/2, 25, C);  FINAL = true: AdaptiveAvgPool2d((2,32)) -> patch operand.
// STAGED = true: the chunk (H * 26 * C * 2 bytes = 80-160 KB, contiguous in HBM) is first copied to shared memory with
// cp.async - every byte of the chunk is in flight at once, no registers involved - and both passes over it (squeeze, gate +
// pool) read shared memory: one CTA per SM, HBM traffic = one read + one write.  STAGED = false: both passes read global
// memory (second pass: L2 hits), 4 CTAs per SM.
template <int C, int H, bool FINAL, bool STAGED>
__global__ void __launch_bounds__(SE_THREADS, STAGED ? 1 : 4) se_fused_kernel(const act16_t* __restrict__ in,
                                                              const act16_t* __restrict__ w0p /*[128][C]*/,
                                                              const float* __restrict__ b0p,
                                                              const act16_t* __restrict__ w2p /*[C][128]*/,
                                                              const float* __restrict__ b2,
                                                              act16_t* __restrict__ out) {
    using S = SeSmem<C>;
    constexpr int R = S::R, LDA = S::LDA, LDZ = S::LDZ, CG = C / 8, TPG = SE_THREADS / CG;
    extern __shared__ __align__(16) uint8_t se_smem[];
    act16_t* sZ = reinterpret_cast<act16_t*>(se_smem);
    act16_t* sA = reinterpret_cast<act16_t*>(se_smem + S::Z_BYTES);
    float* sG = reinterpret_cast<float*>(se_smem + S::Z_BYTES);          // aliases sA (see SeSmem)
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PLGeom gi = make_pl(H, SE_W);
    const uint4* gsrc = reinterpret_cast<const uint4*>(in + (long)n * gi.S * C);
    const uint4* src = gsrc;
    if (STAGED) {
        uint8_t* chunk = se_smem + S::BYTES;                               // after the Z / means / gate region
        const uint32_t cbase = smem_u32(chunk);
        constexpr int TOTAL = H * (SE_W + 1) * CG;                         // 16-byte pieces of the valid rows (pad row not needed)
        for (int i = tid; i < TOTAL; i += SE_THREADS) cp_async_16(cbase + i * 16, gsrc + i);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        src = reinterpret_cast<const uint4*>(chunk);
    }
    auto ld = [&](const uint4* p) -> uint4 { return STAGED ? *p : __ldg(p); };

    // ---- squeeze: mean over H of every (column, channel) -> a16 A operand [32][C] (rows 25..31 zero) ----
    {
        const int cg = tid % CG, sub = tid / CG;
        const float inv = 1.f / (float)H;
        // H is a compile-time constant: all H loads of a column (and two columns when H <= 6) are in flight at once
#pragma unroll(H <= 6 ? 2 : 1)
        for (int w = sub; w < 32; w += TPG) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (w < SE_W) {
                uint4 v[H];
#pragma unroll
                for (int h = 0; h < H; ++h) v[h] = ld(src + (long)(h * gi.P + w) * CG + cg);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const uint4 a = v[h];
                    acc[0] += a16_lo(a.x); acc[1] += a16_hi(a.x); acc[2] += a16_lo(a.y); acc[3] += a16_hi(a.y);
                    acc[4] += a16_lo(a.z); acc[5] += a16_hi(a.z); acc[6] += a16_lo(a.w); acc[7] += a16_hi(a.w);
                }
            }
            *reinterpret_cast<uint4*>(sA + w * LDA + cg * 8) =
                make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                           pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
        }
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    // ---- FC1 + ReLU: Z[32][R] = relu(A[32][C] * W0^T + b0); one (m-tile, n-tile) pair per warp ----
    {
        constexpr int NT = R / 8;                          // 2 or 4 n-tiles; 2 m-tiles
        if (warp < 2 * NT) {
            const int mt = warp / NT, nt = warp % NT;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            const act16_t* arow0 = sA + (mt * 16 + g) * LDA + 2 * t;
            const act16_t* wrow = w0p + (long)(nt * 8 + g) * C + 2 * t;
#pragma unroll 8
            for (int k = 0; k < C; k += 16) {
                uint32_t a[4];
                a[0] = *reinterpret_cast<const uint32_t*>(arow0 + k);
                a[1] = *reinterpret_cast<const uint32_t*>(arow0 + 8 * LDA + k);
                a[2] = *reinterpret_cast<const uint32_t*>(arow0 + k + 8);
                a[3] = *reinterpret_cast<const uint32_t*>(arow0 + 8 * LDA + k + 8);
                const uint32_t b0 = __ldg(reinterpret_cast<const uint32_t*>(wrow + k));
                const uint32_t b1 = __ldg(reinterpret_cast<const uint32_t*>(wrow + k + 8));
                mma_a16_16816(d, a, b0, b1);
            }
            const int col = nt * 8 + 2 * t;
            const float bb0 = __ldg(b0p + col), bb1 = __ldg(b0p + col + 1);
            *reinterpret_cast<uint32_t*>(sZ + (mt * 16 + g) * LDZ + col) = pack_a16(fmaxf(d[0] + bb0, 0.f), fmaxf(d[1] + bb1, 0.f));
            *reinterpret_cast<uint32_t*>(sZ + (mt * 16 + g + 8) * LDZ + col) = pack_a16(fmaxf(d[2] + bb0, 0.f), fmaxf(d[3] + bb1, 0.f));
        }
    }
    __syncthreads();
    // ---- FC2 + sigmoid: G[25][C] = sigmoid(Z[32][R] * W2^T + b2); each warp owns C/8 channels ----
    {
        constexpr int KS = R / 16;                         // 1 or 2 k-steps
        uint32_t a[2][KS][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const act16_t* zr = sZ + (mt * 16 + g) * LDZ + ks * 16 + 2 * t;
                a[mt][ks][0] = *reinterpret_cast<const uint32_t*>(zr);
                a[mt][ks][1] = *reinterpret_cast<const uint32_t*>(zr + 8 * LDZ);
                a[mt][ks][2] = *reinterpret_cast<const uint32_t*>(zr + 8);
                a[mt][ks][3] = *reinterpret_cast<const uint32_t*>(zr + 8 * LDZ + 8);
            }
        constexpr int NT_PER_WARP = C / 8 / 8;             // n-tiles of 8 channels per warp
#pragma unroll 4
        for (int i = 0; i < NT_PER_WARP; ++i) {
            const int c0 = (warp * NT_PER_WARP + i) * 8;
            const act16_t* wrow = w2p + (long)(c0 + g) * 128 + 2 * t;
            uint32_t b[KS][2];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                b[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ks * 16));
                b[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ks * 16 + 8));
            }
            const int col = c0 + 2 * t;
            const float bb0 = __ldg(b2 + col), bb1 = __ldg(b2 + col + 1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) mma_a16_16816(d, a[mt][ks], b[ks][0], b[ks][1]);
                const int r0 = mt * 16 + g, r1 = r0 + 8;
                if (r0 < SE_W)
                    *reinterpret_cast<float2*>(sG + r0 * C + col) =
                        make_float2(fast_sigmoid(d[0] + bb0), fast_sigmoid(d[1] + bb1));
                if (r1 < SE_W)
                    *reinterpret_cast<float2*>(sG + r1 * C + col) =
                        make_float2(fast_sigmoid(d[2] + bb0), fast_sigmoid(d[3] + bb1));
            }
        }
    }
    __syncthreads();
    // ---- excite + pool (second read of the chunk: L2 hits) ----
    if (!FINAL) {
        const PLGeom go = make_pl(H / 2, SE_W);
        uint4* dst = reinterpret_cast<uint4*>(out + (long)n * go.S * C);
#pragma unroll 4
        for (int idx = tid; idx < go.S * CG; idx += SE_THREADS) {
            const int cg = idx % CG, pos = idx / CG;
            const int oh = pos / go.P, ow = pos - oh * go.P;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (oh < go.H && ow < go.W) {
                const uint4* p = src + (long)(2 * oh * gi.P + ow) * CG + cg;
                o = max4(ld(p), ld(p + (long)gi.P * CG));
                const float4 ga = *reinterpret_cast<const float4*>(sG + ow * C + cg * 8);
                const float4 gb = *reinterpret_cast<const float4*>(sG + ow * C + cg * 8 + 4);
                o = make_uint4(pack_a16(a16_lo(o.x) * ga.x, a16_hi(o.x) * ga.y),
                               pack_a16(a16_lo(o.y) * ga.z, a16_hi(o.y) * ga.w),
                               pack_a16(a16_lo(o.z) * gb.x, a16_hi(o.z) * gb.y),
                               pack_a16(a16_lo(o.w) * gb.z, a16_hi(o.w) * gb.w));
            }
            dst[idx] = o;
        }
    } else {
        uint4* dst = reinterpret_cast<uint4*>(out + (long)n * TOK_PER_CHUNK * 2 * C);
#pragma unroll 4
        for (int idx = tid; idx < TOK_PER_CHUNK * 2 * CG; idx += SE_THREADS) {
            const int cg = idx % CG, kh = (idx / CG) & 1, k = idx / (2 * CG);
            const int h0 = (kh * H) / 2, h1 = ((kh + 1) * H + 1) / 2;
            const int w0 = (k * SE_W) / TOK_PER_CHUNK, w1 = ((k + 1) * SE_W + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int w = w0; w < w1; ++w) {
                const float4 ga = *reinterpret_cast<const float4*>(sG + w * C + cg * 8);
                const float4 gb = *reinterpret_cast<const float4*>(sG + w * C + cg * 8 + 4);
                for (int h = h0; h < h1; ++h) {
                    const uint4 a = ld(src + (long)(h * gi.P + w) * CG + cg);
                    acc[0] += a16_lo(a.x) * ga.x; acc[1] += a16_hi(a.x) * ga.y;
                    acc[2] += a16_lo(a.y) * ga.z; acc[3] += a16_hi(a.y) * ga.w;
                    acc[4] += a16_lo(a.z) * gb.x; acc[5] += a16_hi(a.z) * gb.y;
                    acc[6] += a16_lo(a.w) * gb.z; acc[7] += a16_hi(a.w) * gb.w;
                }
            }
            const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
            dst[idx] = make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                                  pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
        }
    }
}

static int g_se_staged = 0;            // 1: chunk staged in shared memory (cp.async) - measured 0.14 / 0.20 / 0.22 ms vs 0.14 / 0.23 / 0.17 ms
                                       // for the two-pass version (one CTA per SM serialises load, FC latency and store): off by default
void set_se_staged(int on) { g_se_staged = on; }

template <int C, int H, bool FINAL, bool STAGED>
static int launch_se_fused_variant(const act16_t* in, const SEWeights& w, act16_t* out, int n_chunks, cudaStream_t stream) {
    const size_t smem = SeSmem<C>::BYTES + (STAGED ? (size_t)H * (SE_W + 1) * C * 2 : 0);
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, se_fused_kernel<C, H, FINAL, STAGED>, (int)smem));
    se_fused_kernel<C, H, FINAL, STAGED><<<n_chunks, SE_THREADS, smem, stream>>>(in, w.w0p, w.b0p, w.w2p, w.b2, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

template <int C, int H, bool FINAL>
static int launch_se_fused_impl(const act16_t* in, const SEWeights& w, act16_t* out, int n_chunks,
                                cudaStream_t stream) {
    if (g_se_staged) return launch_se_fused_variant<C, H, FINAL, true>(in, w, out, n_chunks, stream);
    return launch_se_fused_variant<C, H, FINAL, false>(in, w, out, n_chunks, stream);
}

// The three SE sites of the backbone (se_model.py:47,53,59): (C, H) = (256, 12), (512, 6) and (512, 3) + final pool.
int launch_se_fused(const act16_t* in, const SEWeights& w, act16_t* out, int n_chunks, int H, int W, int C,
                    bool final_pool, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    if (W == SE_W && C == 256 && H == 12 && !final_pool) return launch_se_fused_impl<256, 12, false>(in, w, out, n_chunks, stream);
    if (W == SE_W && C == 512 && H == 6 && !final_pool) return launch_se_fused_impl<512, 6, false>(in, w, out, n_chunks, stream);
    if (W == SE_W && C == 512 && H == 3 && final_pool) return launch_se_fused_impl<512, 3, true>(in, w, out, n_chunks, stream);
    KOCR_CHECK(false, "se_fused: unsupported geometry H=%d W=%d C=%d final=%d", H, W, C, (int)final_pool);
    return 2;
}

// 16-bit -> fp32 copy (identity shortcut of a ResNet BasicBlock as the fp32 addend of the block's second conv GEMM).
__global__ void __launch_bounds__(256) a16_to_f32_kernel(const act16_t* __restrict__ in, float* __restrict__ out, long n8) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    o[0] = make_float4(a16_lo(v.x), a16_hi(v.x), a16_lo(v.y), a16_hi(v.y));
    o[1] = make_float4(a16_lo(v.z), a16_hi(v.z), a16_lo(v.w), a16_hi(v.w));
}

int launch_a16_to_f32(const act16_t* in, float* out, long n_elems, cudaStream_t stream) {
    if (n_elems == 0) return 0;
    const long n8 = n_elems / 8;
    a16_to_f32_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(in, out, n8);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace kocr
