// Pieces of the backbones that are not tcgen05 GEMMs:
//   conv1 (Cin = 1: K = 9) + BN + ReLU + 2x2 pool on mma.sync                                      se_model.py:39-40,64
//   2x2 max-pool after conv2, (2,1) max-pool of the ResNet baseline                                 se_model.py:43,65
//   SequenceSE excitation: FC -> ReLU -> FC -> sigmoid (mma.sync) on the column means the conv epilogue produced,
//     then gate * (pooled activations), in place - or gate * bins -> AdaptiveAvgPool2d((2,32)) = the patch-projection
//     operand                                                                                       se_model.py:19-30,48-61,76-78
//   AdaptiveAvgPool2d((2,32)) without a gate (VGG / ResNet baselines)                               vgg_model.py:48,59
//   16-bit -> fp32 copy for the identity shortcuts of the ResNet baseline                           resnet_model.py:17,33
// All activations are 16-bit (act16_t) in the DENSE NWHC layout: [chunk][w][h][C], pixel index (n * W + w) * H + h, so that
// the rows of one image column are adjacent (whole-column GEMM tiles, (2,1) pools and SE column means are local).
// SE math is fp32.
#include "kernels.cuh"

namespace kocr {

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
// Row mapping of the two M-tiles of a "pool tile" (8 pooled pixels of one pooled COLUMN, py = 8i + g): tile A rows g / g+8
// are the pixels (2py, 2px) / (2py, 2px+1), tile B the same on row 2py+1 - so the four partners of a 2x2 pool
// window are the accumulators c0/c2 (c1/c3) of the two tiles of ONE thread: pooling needs no data exchange.
// bias and ReLU commute with the max.  Output: dense NWHC (50 columns x 24 rows x 64), staged per warp in shared memory
// and written as one contiguous 1 KB run (8 vertically adjacent pixels x 128 B) per pool tile.
// ------------------------------------------------------------------------------------------
static constexpr int C1M_LD = 106;                // halves per tile row: 102 used (x = -1 .. 100); 53 words: the 8 rows of a pool tile hit 8 banks

__global__ void __launch_bounds__(256) conv1_pool_mma_kernel(const float* __restrict__ chunks,
                                                             const act16_t* __restrict__ w16 /*[64][16], k = tap, 9..15 zero*/,
                                                             const float* __restrict__ b, act16_t* __restrict__ out) {
    __shared__ __align__(16) act16_t s_tile[(IMG_H + 2) * C1M_LD];
    __shared__ __align__(16) uint32_t s_stage[8][8][32];          // [warp][pooled pixel][64 channels as 32 words]
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int OH = IMG_H / 2, OW = CHUNK_W / 2;               // 24 x 50 output
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
    constexpr int GROUPS = OH / 8;                                // 3 groups of 8 pooled rows per pooled column
    for (int pt = warp; pt < OW * GROUPS; pt += 8) {
        const int px = pt / GROUPS, gi = pt - px * GROUPS;
        const int py = gi * 8 + g;
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
        // pooled pixels (px, gi * 8 .. gi * 8 + 7) are 8 consecutive rows of one column: 1 KB contiguous
        uint4* dst = reinterpret_cast<uint4*>(out + (((long)n * OW + px) * OH + gi * 8) * 64);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int idx = lane + 32 * r, x = idx >> 3, p = idx & 7;             // pooled pixel x of the tile, 16-byte piece p
            dst[idx] = *reinterpret_cast<const uint4*>(&s_stage[warp][x][((p ^ x) & 7) * 4]);
        }
        __syncwarp();
    }
}

int launch_conv1_pool_mma(const float* d_chunks, const act16_t* w16, const float* b, act16_t* out, int n_chunks,
                          cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    conv1_pool_mma_kernel<<<n_chunks, 256, 0, stream>>>(d_chunks, w16, b, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Max-pools between dense NWHC layouts.  One thread per (output pixel, 8 channels).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 max4(uint4 a, uint4 b) {
    return make_uint4(a16x2_max(a.x, b.x), a16x2_max(a.y, b.y), a16x2_max(a.z, b.z), a16x2_max(a.w, b.w));
}

__global__ void __launch_bounds__(256) pool2x2_kernel(const act16_t* __restrict__ in,
                                                      act16_t* __restrict__ out, long total, int H, int W,
                                                      int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cgs = C / 8, Ho = H / 2, Wo = W / 2;
    const int cg = (int)(idx % cgs);
    const long q = idx / cgs;                       // (n * Wo + ow) * Ho + oh
    const int oh = (int)(q % Ho);
    const long nc = q / Ho;
    const int ow = (int)(nc % Wo);
    const long n = nc / Wo;
    const uint4* p = reinterpret_cast<const uint4*>(in) + (((n * W + 2 * ow) * H + 2 * oh) * cgs + cg);
    const uint4 a = __ldg(p), b = __ldg(p + cgs);                              // rows 2oh, 2oh + 1 of column 2ow
    const uint4 c = __ldg(p + (long)H * cgs), d = __ldg(p + (long)(H + 1) * cgs);     // ... of column 2ow + 1
    reinterpret_cast<uint4*>(out)[idx] = max4(max4(a, b), max4(c, d));
}

int launch_pool2x2(const act16_t* in, act16_t* out, int n_chunks, int H, int W, int C,
                   cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const long total = (long)n_chunks * (H / 2) * (W / 2) * (C / 8);
    pool2x2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, total, H, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// (2,1) max-pool: rows 2i, 2i+1 of a column are adjacent in NWHC
__global__ void __launch_bounds__(256) pool_h2_kernel(const act16_t* __restrict__ in, act16_t* __restrict__ out, long total, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const long q = idx / cgs;                       // output pixel = col * Ho + oh  ->  input pixels 2q, 2q + 1
    const uint4* p = reinterpret_cast<const uint4*>(in) + (2 * q * cgs + cg);
    reinterpret_cast<uint4*>(out)[idx] = max4(__ldg(p), __ldg(p + cgs));
}

int launch_pool_h2(const act16_t* in, act16_t* out, int n_chunks, int H, int W, int C, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    const long total = (long)n_chunks * W * (H / 2) * (C / 8);
    pool_h2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, total, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// AdaptiveAvgPool2d((2, 32)) of a (3, 25) map -> patch-projection operand out[n*32 + k][kh*C + c] without a gate
// (se_model.py:61,78 / vgg_model.py:48,59: bin k covers columns [floor(k*W/32), ceil((k+1)*W/32)), row bin kh covers rows
// {kh, kh + 1}).  rows_in == 2: `in` holds the row-bin SUMS the conv7 epilogue wrote ([col][2][C]); rows_in == 3: the raw
// rows ([col][3][C], ResNet baseline).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalpool_kernel(const act16_t* __restrict__ in, act16_t* __restrict__ out, long total,
                                                        int rows_in, int W, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cgs = C / 8;
    const int cg = (int)(idx % cgs);
    const int kh = (int)((idx / cgs) & 1);
    const long nk = idx / (2 * cgs);                 // n*32 + k
    const long n = nk / TOK_PER_CHUNK;
    const int k = (int)(nk - n * TOK_PER_CHUNK);
    const int w0 = (k * W) / TOK_PER_CHUNK, w1 = ((k + 1) * W + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int w = w0; w < w1; ++w) {
        const uint4* p = reinterpret_cast<const uint4*>(in) + (((n * W + w) * rows_in + kh) * cgs + cg);
        for (int r = 0; r < (rows_in == 2 ? 1 : 2); ++r) {
            const uint4 a = __ldg(p + (long)r * cgs);
            acc[0] += a16_lo(a.x); acc[1] += a16_hi(a.x); acc[2] += a16_lo(a.y); acc[3] += a16_hi(a.y);
            acc[4] += a16_lo(a.z); acc[5] += a16_hi(a.z); acc[6] += a16_lo(a.w); acc[7] += a16_hi(a.w);
        }
    }
    const float inv = 1.f / (float)(2 * (w1 - w0));
    reinterpret_cast<uint4*>(out)[idx] =
        make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                   pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
}

int launch_finalpool(const act16_t* in, int rows_in, act16_t* out, int n_chunks, int W, int C, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    KOCR_CHECK(rows_in == 2 || rows_in == 3, "finalpool: rows_in must be 2 (bin sums) or 3 (raw rows)");
    const long total = (long)n_chunks * TOK_PER_CHUNK * 2 * (C / 8);
    finalpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, total, rows_in, W, C);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// SequenceSE excitation + gating (SequenceSE.forward, se_model.py:19-30, and the pooling that follows it, :69-78).
// The squeeze (column means over H, from the fp32 accumulators) and the (2,1) max-pool / row-bin sums were done by the
// producing conv's epilogue (gemm_tc.cu, column-fused mode): what is left per chunk is
//     gate[25][C] = sigmoid(W2 relu(W0 mean + b0) + b2)            two [25 -> 32] x C x C/16 contractions
//     pooled[w][r][c] *= gate[w][c]                                 (gate > 0: max-pool and gate commute exactly)
// or, for the last block, gate * row-bin sums -> AdaptiveAvgPool2d((2,32)) -> patch operand.
// Persistent CTAs, one chunk per iteration; the contractions are far below a tcgen05 tile, so they run on
// mma.sync.m16n8k16 (a16, fp32 accumulate) with the means / hidden vector as A operands in shared memory and the weights
// read through L1 as B fragments (fragment order: one coalesced 16-byte load per lane per two k-steps).
// HBM traffic per chunk: 25*C*2 B of means + one read and one write of the POOLED tensor (the r01 kernel read the
// un-pooled conv output twice).
// ------------------------------------------------------------------------------------------

// sigmoid with the approximate reciprocal (MUFU.RCP, 1 ulp): the IEEE division of `1.f / (1.f + __expf(-x))` expands to a
// branchy ~20-instruction sequence; the gate only scales 16-bit activations
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

static constexpr int SE_W = 25;            // columns of every SE stage (100 / 4)
static constexpr int SE_THREADS = 256;

template <int C, bool FRAG> struct SeSmem {
    static constexpr int R = C / 16;                       // reduced width (se_model.py:9,13)
    static constexpr int LDA = C + 8;                      // a16 elements per row of the means (conflict-free frags)
    static constexpr int LDZ = R + 8;
    // gate row stride in floats: + 8 spreads the 8 accumulator rows of an FC2 store over the banks (2 wavefronts per
    // float2 store instead of 8 - ncu, round 2: 45 % of the kernel's shared-memory wavefronts were bank conflicts)
    static constexpr int LDG = FRAG ? C + 8 : C;
    static constexpr size_t A_BYTES = 32 * LDA * 2;
    static constexpr size_t Z_BYTES = 32 * LDZ * 2;
    static constexpr size_t G_BYTES = (size_t)SE_W * LDG * 4;
    // the gate (written by FC2) re-uses the storage of the column means (dead once FC1 is done); the final-pool variant
    // adds room for the chunk's 25 x 2 row-bin sums
    static constexpr size_t AG_BYTES = A_BYTES > G_BYTES ? A_BYTES : G_BYTES;
    static constexpr size_t BYTES = Z_BYTES + AG_BYTES;
    static constexpr size_t BINS_BYTES = (size_t)SE_W * 2 * C * 2;
};

// ROWS = pooled rows per column (H/2) - or 2 row-bin sums when FINAL.  FINAL = false: pooled is scaled in place;
// FINAL = true: pooled = bins [col][2][C], out = patch operand [n*32 + k][kh*C + c].
// PREFETCH = true: the pooled block is loaded into registers before the contractions (76 registers: 2 CTAs per SM);
// PREFETCH = false (not FINAL): the block is streamed load -> scale -> store after them (<= 64 registers: 4 CTAs per SM).
// FRAG = true: the FC weights are read in mma-fragment order (w0f / w2f, weights.se_fragments): one coalesced 16-byte load
// per lane covers two k-steps, where the row-major layout (w0p / w2p) costs 8 L1 wavefronts per 4-byte fragment load - the
// L1/TEX pipe was the busiest unit of this kernel (ncu: 61-71 %).
template <int C, int ROWS, bool FINAL, bool PREFETCH, bool FRAG>
__global__ void __launch_bounds__(SE_THREADS, PREFETCH ? 2 : 4) se_excite_kernel(const act16_t* __restrict__ means /*[n*25 + w][C]*/,
                                                                  const act16_t* __restrict__ w0p /*[128][C]*/,
                                                                  const float* __restrict__ b0p,
                                                                  const act16_t* __restrict__ w2p /*[C][128]*/,
                                                                  const float* __restrict__ b2,
                                                                  const act16_t* __restrict__ w0f, const act16_t* __restrict__ w2f,
                                                                  act16_t* __restrict__ pooled, act16_t* __restrict__ out,
                                                                  int n_chunks) {
    using S = SeSmem<C, FRAG>;
    constexpr int R = S::R, LDA = S::LDA, LDZ = S::LDZ, CG = C / 8, LDG = S::LDG;
    extern __shared__ __align__(16) uint8_t se_smem[];
    act16_t* sZ = reinterpret_cast<act16_t*>(se_smem);
    act16_t* sA = reinterpret_cast<act16_t*>(se_smem + S::Z_BYTES);
    float* sG = reinterpret_cast<float*>(se_smem + S::Z_BYTES);          // aliases sA (see SeSmem)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // PERSISTENT over chunks (grid = 2 CTAs per SM): the 2 x 32 KB of FC weights a CTA reads through __ldg stay in ITS L1
    // from the second chunk on - with one CTA per chunk every CTA paid their L2 latency inside two dependent MMA chains
    // (ncu, round 2: 17 % SM busy, long-scoreboard stalls).
#pragma unroll 1
    for (int n = blockIdx.x; n < n_chunks; n += gridDim.x) {
    // ---- the chunk's pooled block (ROWS * 25 * C 16-bit values, contiguous) starts streaming into REGISTERS now: its HBM
    //      latency is hidden behind the two small contractions below instead of following them ----
    constexpr int TOTAL = SE_W * ROWS * CG;                 // 16-byte pieces of the block
    constexpr int PER_THREAD = (TOTAL + SE_THREADS - 1) / SE_THREADS;
    const uint4* blk_in = reinterpret_cast<const uint4*>(pooled + (long)n * SE_W * ROWS * C);
    uint4 pre[PREFETCH ? PER_THREAD : 1];
    if (PREFETCH) {
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + i * SE_THREADS;
            pre[i] = idx < TOTAL ? __ldg(blk_in + idx) : make_uint4(0, 0, 0, 0);
        }
    }
    // ---- column means (16-bit, written by the conv epilogue from its fp32 accumulators) -> A operand [32][C] (rows 25..31 zero) ----
    {
        const uint4* msrc = reinterpret_cast<const uint4*>(means + (long)n * SE_W * C);
        for (int i = tid; i < 32 * CG; i += SE_THREADS) {
            const int w = i / CG, cg = i - w * CG;
            *reinterpret_cast<uint4*>(sA + w * LDA + cg * 8) = w < SE_W ? __ldg(msrc + i) : make_uint4(0, 0, 0, 0);
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
            auto a_frag = [&](int k, uint32_t (&a)[4]) {
                a[0] = *reinterpret_cast<const uint32_t*>(arow0 + k);
                a[1] = *reinterpret_cast<const uint32_t*>(arow0 + 8 * LDA + k);
                a[2] = *reinterpret_cast<const uint32_t*>(arow0 + k + 8);
                a[3] = *reinterpret_cast<const uint32_t*>(arow0 + 8 * LDA + k + 8);
            };
            if (FRAG) {
                const uint4* wf = reinterpret_cast<const uint4*>(w0f) + (long)nt * (C / 32) * 32 + lane;
#pragma unroll 4
                for (int kp = 0; kp < C / 32; ++kp) {
                    const uint4 b = __ldg(wf + kp * 32);
                    uint32_t a[4];
                    a_frag(kp * 32, a);
                    mma_a16_16816(d, a, b.x, b.y);
                    a_frag(kp * 32 + 16, a);
                    mma_a16_16816(d, a, b.z, b.w);
                }
            } else {
                const act16_t* wrow = w0p + (long)(nt * 8 + g) * C + 2 * t;
#pragma unroll 8
                for (int k = 0; k < C; k += 16) {
                    uint32_t a[4];
                    a_frag(k, a);
                    const uint32_t b0 = __ldg(reinterpret_cast<const uint32_t*>(wrow + k));
                    const uint32_t b1 = __ldg(reinterpret_cast<const uint32_t*>(wrow + k + 8));
                    mma_a16_16816(d, a, b0, b1);
                }
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
            uint32_t b[KS][2];
            if (FRAG) {
                const int ntile = warp * NT_PER_WARP + i;
                if (KS == 2) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(w2f) + ntile * 32 + lane);
                    b[0][0] = v.x; b[0][1] = v.y; b[KS - 1][0] = v.z; b[KS - 1][1] = v.w;
                } else {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(w2f) + ntile * 32 + lane);
                    b[0][0] = v.x; b[0][1] = v.y;
                }
            } else {
                const act16_t* wrow = w2p + (long)(c0 + g) * 128 + 2 * t;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    b[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ks * 16));
                    b[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ks * 16 + 8));
                }
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
                    *reinterpret_cast<float2*>(sG + r0 * LDG + col) =
                        make_float2(fast_sigmoid(d[0] + bb0), fast_sigmoid(d[1] + bb1));
                if (r1 < SE_W)
                    *reinterpret_cast<float2*>(sG + r1 * LDG + col) =
                        make_float2(fast_sigmoid(d[2] + bb0), fast_sigmoid(d[3] + bb1));
            }
        }
    }
    __syncthreads();
    // ---- excite ----
    if (!FINAL) {
        uint4* blk = reinterpret_cast<uint4*>(pooled + (long)n * SE_W * ROWS * C);    // scaled in place
        auto scale = [&](int idx, const uint4 o) {
            const int cg = idx % CG, w = idx / (ROWS * CG);
            const float4 ga = *reinterpret_cast<const float4*>(sG + w * LDG + cg * 8);
            const float4 gb = *reinterpret_cast<const float4*>(sG + w * LDG + cg * 8 + 4);
            blk[idx] = make_uint4(pack_a16(a16_lo(o.x) * ga.x, a16_hi(o.x) * ga.y),
                                  pack_a16(a16_lo(o.y) * ga.z, a16_hi(o.y) * ga.w),
                                  pack_a16(a16_lo(o.z) * gb.x, a16_hi(o.z) * gb.y),
                                  pack_a16(a16_lo(o.w) * gb.z, a16_hi(o.w) * gb.w));
        };
        if (PREFETCH) {
#pragma unroll
            for (int i = 0; i < PER_THREAD; ++i) {
                const int idx = tid + i * SE_THREADS;
                if (idx < TOTAL) scale(idx, pre[i]);
            }
        } else {
            // streamed: 4 loads in flight per thread, 32 warps per SM
            constexpr int FULL = TOTAL / SE_THREADS / 4 * 4;
#pragma unroll 1
            for (int i0 = 0; i0 < FULL; i0 += 4) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = __ldg(blk_in + tid + (i0 + j) * SE_THREADS);
#pragma unroll
                for (int j = 0; j < 4; ++j) scale(tid + (i0 + j) * SE_THREADS, v[j]);
            }
#pragma unroll
            for (int i = FULL; i < PER_THREAD; ++i) {
                const int idx = tid + i * SE_THREADS;
                if (idx < TOTAL) scale(idx, __ldg(blk_in + idx));
            }
        }
    } else {
        // the prefetched row-bin sums go to shared memory: the adaptive pool below reads 1-2 columns per output bin
        uint4* sB = reinterpret_cast<uint4*>(se_smem + S::BYTES);                     // [25 * 2 * CG] pieces
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + i * SE_THREADS;
            if (idx < TOTAL) sB[idx] = pre[i];
        }
        __syncthreads();
        const uint4* bins = sB;
        uint4* dst = reinterpret_cast<uint4*>(out + (long)n * TOK_PER_CHUNK * 2 * C);
#pragma unroll 4
        for (int idx = tid; idx < TOK_PER_CHUNK * 2 * CG; idx += SE_THREADS) {
            const int cg = idx % CG, kh = (idx / CG) & 1, k = idx / (2 * CG);
            const int w0 = (k * SE_W) / TOK_PER_CHUNK, w1 = ((k + 1) * SE_W + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int w = w0; w < w1; ++w) {
                const float4 ga = *reinterpret_cast<const float4*>(sG + w * LDG + cg * 8);
                const float4 gb = *reinterpret_cast<const float4*>(sG + w * LDG + cg * 8 + 4);
                const uint4 a = bins[(w * 2 + kh) * CG + cg];
                acc[0] += a16_lo(a.x) * ga.x; acc[1] += a16_hi(a.x) * ga.y;
                acc[2] += a16_lo(a.y) * ga.z; acc[3] += a16_hi(a.y) * ga.w;
                acc[4] += a16_lo(a.z) * gb.x; acc[5] += a16_hi(a.z) * gb.y;
                acc[6] += a16_lo(a.w) * gb.z; acc[7] += a16_hi(a.w) * gb.w;
            }
            const float inv = 1.f / (float)(2 * (w1 - w0));
            dst[idx] = make_uint4(pack_a16(acc[0] * inv, acc[1] * inv), pack_a16(acc[2] * inv, acc[3] * inv),
                                  pack_a16(acc[4] * inv, acc[5] * inv), pack_a16(acc[6] * inv, acc[7] * inv));
        }
    }
    __syncthreads();        // shared memory is reused by the next chunk
    }   // chunk loop
}

static int g_se_stream_variant = 4;
void set_se_excite_variant(int v) { g_se_stream_variant = v; }

template <int C, int ROWS, bool FINAL, bool PREFETCH, bool FRAG>
static int launch_se_excite_impl(const act16_t* means, const SEWeights& w, act16_t* pooled, act16_t* out, int n_chunks,
                                 cudaStream_t stream) {
    size_t smem = SeSmem<C, FRAG>::BYTES + (FINAL ? SeSmem<C, FRAG>::BINS_BYTES : 0);
    int per_sm = PREFETCH ? 2 : 4;
    if (!PREFETCH && g_se_stream_variant % 100 == 3) { per_sm = 3; if (smem < 72 * 1024) smem = 72 * 1024; }   // (probe: 3 CTAs per SM)
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, se_excite_kernel<C, ROWS, FINAL, PREFETCH, FRAG>, 72 * 1024 > (int)smem ? 72 * 1024 : (int)smem));
    int dev = 0, sms = 148;
    KOCR_CUDA(cudaGetDevice(&dev));
    KOCR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = n_chunks < per_sm * sms ? n_chunks : per_sm * sms;
    se_excite_kernel<C, ROWS, FINAL, PREFETCH, FRAG><<<grid, SE_THREADS, smem, stream>>>(means, w.w0p, w.b0p, w.w2p, w.b2, w.w0f, w.w2f, pooled, out,
                                                                                       n_chunks);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// The three SE sites of the backbone (se_model.py:47,53,59): (C, pooled rows) = (256, 6), (512, 3) and (512, 2 bins) + final pool.
int launch_se_excite(const act16_t* means, const SEWeights& w, act16_t* pooled, act16_t* out, int n_chunks, int rows, int W,
                     int C, bool final_pool, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    // g_se_stream_variant (option "se_variant", probes only): 4 (default) / 3 = fragment-order weights, pooled block streamed
    // after the gate with 4 / 3 CTAs per SM; 0 = fragment-order weights + register prefetch (2 CTAs per SM); 100 + v = the
    // same with the row-major weights (the kernel of the first half of round 2).  Measured (tools/se_probe.py, 2039 chunks,
    // se3 + se4 + se5): 100: 0.418 ms, 104: 0.400, 0: 0.334, 4: 0.301 - all bit-identical.
    const bool frag = g_se_stream_variant < 100;
    const bool streamed = g_se_stream_variant % 100 != 0;
#define KOCR_SE(CC, RR, FF)                                                                                                      \
    if (frag) return streamed && !FF ? launch_se_excite_impl<CC, RR, FF, FF, true>(means, w, pooled, out, n_chunks, stream)      \
                                     : launch_se_excite_impl<CC, RR, FF, true, true>(means, w, pooled, out, n_chunks, stream);   \
    return streamed && !FF ? launch_se_excite_impl<CC, RR, FF, FF, false>(means, w, pooled, out, n_chunks, stream)               \
                           : launch_se_excite_impl<CC, RR, FF, true, false>(means, w, pooled, out, n_chunks, stream);
    if (W == SE_W && C == 256 && rows == 6 && !final_pool) { KOCR_SE(256, 6, false) }
    if (W == SE_W && C == 512 && rows == 3 && !final_pool) { KOCR_SE(512, 3, false) }
    if (W == SE_W && C == 512 && rows == 2 && final_pool) { KOCR_SE(512, 2, true) }
#undef KOCR_SE
    KOCR_CHECK(false, "se_excite: unsupported geometry rows=%d W=%d C=%d final=%d", rows, W, C, (int)final_pool);
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
