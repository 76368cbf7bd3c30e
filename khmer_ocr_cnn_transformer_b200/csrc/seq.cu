// Sequence-side kernels: per-chunk attention (mma.sync; CUDA-core version kept for A/B), LayerNorm, the padded
// memory of the teacher-forced forward, BiLSTM recurrence (mma.sync cluster kernel + CUDA-core version), and the
// KV-cached decoder step kernels (self-attention, one-pass cross-attention, embed, argmax).  All softmax / LayerNorm /
// LSTM state math is fp32.
//
// Reference ops replaced:
//   nn.TransformerEncoderLayer self-attention over 32 tokens            se_model.py:119-126
//   nn.LayerNorm(384, eps 1e-5)                                          (inside enc / dec layers)
//   nn.LSTM(384 -> 192, bidirectional) recurrence                        se_model.py:228-234
//   TransformerDecoderWrapper + OCRPredictor._greedy_decode              se_model.py:182-208, predictor.py:85-99
#include "kernels.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace kocr {

// ------------------------------------------------------------------------------------------
// Per-chunk attention: one CTA per chunk, warp = head, lane = query token.
// K/V of the head live in shared memory (fp32), scores/softmax in registers.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_attention_kernel(const act16_t* __restrict__ qkv,
                                                              act16_t* __restrict__ out) {
    extern __shared__ float s_kv[];                 // [8 warps][2][32][48]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sk = s_kv + warp * (2 * 32 * HEAD_DIM);
    float* sv = sk + 32 * HEAD_DIM;
    const long row = (long)blockIdx.x * TOK_PER_CHUNK + lane;
    const act16_t* base = qkv + row * (3 * D_MODEL) + warp * HEAD_DIM;
    float q[HEAD_DIM];
    const float scale = rsqrtf((float)HEAD_DIM);
#pragma unroll
    for (int i = 0; i < HEAD_DIM / 8; ++i) {
        const uint4 a = reinterpret_cast<const uint4*>(base)[i];
        const uint4 b = reinterpret_cast<const uint4*>(base + D_MODEL)[i];
        const uint4 c = reinterpret_cast<const uint4*>(base + 2 * D_MODEL)[i];
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[i * 8 + 2 * j] = a16_lo(av[j]) * scale;
            q[i * 8 + 2 * j + 1] = a16_hi(av[j]) * scale;
            sk[lane * HEAD_DIM + i * 8 + 2 * j] = a16_lo(bv[j]);
            sk[lane * HEAD_DIM + i * 8 + 2 * j + 1] = a16_hi(bv[j]);
            sv[lane * HEAD_DIM + i * 8 + 2 * j] = a16_lo(cv[j]);
            sv[lane * HEAD_DIM + i * 8 + 2 * j + 1] = a16_hi(cv[j]);
        }
    }
    __syncwarp();
    float s[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < HEAD_DIM; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(sk + j * HEAD_DIM + d);
            acc = fmaf(q[d], k4.x, acc); acc = fmaf(q[d + 1], k4.y, acc);
            acc = fmaf(q[d + 2], k4.z, acc); acc = fmaf(q[d + 3], k4.w, acc);
        }
        s[j] = acc;
        mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { s[j] = __expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.f / sum;
    float o[HEAD_DIM];
#pragma unroll
    for (int d = 0; d < HEAD_DIM; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float pj = s[j] * inv;
#pragma unroll
        for (int d = 0; d < HEAD_DIM; d += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(sv + j * HEAD_DIM + d);
            o[d] = fmaf(pj, v4.x, o[d]); o[d + 1] = fmaf(pj, v4.y, o[d + 1]);
            o[d + 2] = fmaf(pj, v4.z, o[d + 2]); o[d + 3] = fmaf(pj, v4.w, o[d + 3]);
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + row * D_MODEL + warp * HEAD_DIM);
#pragma unroll
    for (int i = 0; i < HEAD_DIM / 8; ++i)
        dst[i] = make_uint4(pack_a16(o[i * 8], o[i * 8 + 1]), pack_a16(o[i * 8 + 2], o[i * 8 + 3]),
                            pack_a16(o[i * 8 + 4], o[i * 8 + 5]), pack_a16(o[i * 8 + 6], o[i * 8 + 7]));
}

// ------------------------------------------------------------------------------------------
// Tensor-core version: one CTA per chunk, warp = head.  The chunk's [32][1152] a16 QKV block is copied to shared
// memory with coalesced 16-byte loads (it is contiguous in HBM), every head then runs S = Q K^T (m16n8k16, fp32
// accumulators), a register softmax (quad shuffles), and O = P V with the probabilities re-used directly as A
// fragments (accumulator layout of S == A layout of the next MMA) and V read through ldmatrix.trans.  The output is
// staged in the head's dead Q columns and written back as whole 768-byte rows.
// 32 x 32 x 48 per head is far below a tcgen05 tile (M = 128), hence mma.sync here.
// ------------------------------------------------------------------------------------------
static constexpr int CA_LD = 3 * D_MODEL + 8;      // a16 elements per smem row: 2320 B = odd multiple of 16 B

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." KOCR_MMA_A16 "." KOCR_MMA_A16 ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3) chunk_attention_mma_kernel(const act16_t* __restrict__ qkv,
                                                                     act16_t* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t ca_smem[];
    act16_t* sm = reinterpret_cast<act16_t*>(ca_smem);          // [32][CA_LD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + (long)blockIdx.x * TOK_PER_CHUNK * 3 * D_MODEL);
    constexpr int ROW_U4 = 3 * D_MODEL / 8;        // 144 16-byte pieces per row
#pragma unroll 6
    for (int i = tid; i < TOK_PER_CHUNK * ROW_U4; i += 256) {
        const int r = i / ROW_U4, c = i - r * ROW_U4;
        *reinterpret_cast<uint4*>(sm + r * CA_LD + c * 8) = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t sbase = smem_u32(sm);
    const int qc = warp * HEAD_DIM, kc = D_MODEL + warp * HEAD_DIM, vc = 2 * D_MODEL + warp * HEAD_DIM;
    const int g = lane >> 2, t = lane & 3;
    // ldmatrix row addresses: x4 (A operand): lane -> matrix lane/8 = {rows 0-7 | rows 8-15} x {k 0-7 | k 8-15};
    // x2 (B operand): lanes 0-7 -> rows at k0, lanes 8-15 -> rows at k0 + 8 (other lanes' addresses are ignored)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
    const int b_row = lane & 7, b_col = ((lane >> 3) & 1) * 8;
    const float scale_log2 = rsqrtf((float)HEAD_DIM) * 1.4426950408889634f;    // softmax in base 2
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
        // ---- S = Q K^T for 16 queries x 32 keys ----
        float sacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HEAD_DIM / 16; ++ks) {
            uint32_t a[4];
            ldsm_x4(a, sbase + ((mt * 16 + a_row) * CA_LD + qc + ks * 16 + a_col) * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t b0, b1;
                ldsm_x2(b0, b1, sbase + ((j * 8 + b_row) * CA_LD + kc + ks * 16 + b_col) * 2);
                mma16816(sacc[j], a, b0, b1);
            }
        }
        // ---- softmax over the 32 keys of rows g and g + 8 (values spread over the 4 lanes of a quad) ----
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
            m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        uint32_t pa[2][4];                           // P as A fragments of the two 16-key k-steps
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = pack_a16(exp2f((sacc[j][0] - m0) * scale_log2), exp2f((sacc[j][1] - m0) * scale_log2));
            const uint32_t hi = pack_a16(exp2f((sacc[j][2] - m1) * scale_log2), exp2f((sacc[j][3] - m1) * scale_log2));
            s0 += a16_lo(lo) + a16_hi(lo);         // sums of the ROUNDED weights: the applied weights add up to 1
            s1 += a16_lo(hi) + a16_hi(hi);
            pa[j >> 1][(j & 1) * 2] = lo;
            pa[j >> 1][(j & 1) * 2 + 1] = hi;
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float inv0 = 1.f / s0, inv1 = 1.f / s1;
        // ---- O = P V: 16 queries x 48 dims, K = 32 keys ----
        float oacc[HEAD_DIM / 8][4];
#pragma unroll
        for (int j = 0; j < HEAD_DIM / 8; ++j) { oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int j = 0; j < HEAD_DIM / 8; ++j) {
                uint32_t b0, b1;
                // V[key][d] row-major, keys = K dimension: transposed 8x8 loads give B[k = 2t, 2t+1][n = g]
                ldsm_x2_trans(b0, b1, sbase + ((ks * 16 + b_row + b_col) * CA_LD + vc + j * 8) * 2);
                mma16816(oacc[j], pa[ks], b0, b1);
            }
        }
        // ---- stage the output in this head's Q columns (rows of this m-tile: no other warp reads them) ----
        __syncwarp();
#pragma unroll
        for (int j = 0; j < HEAD_DIM / 8; ++j) {
            *reinterpret_cast<uint32_t*>(sm + (mt * 16 + g) * CA_LD + qc + j * 8 + 2 * t) = pack_a16(oacc[j][0] * inv0, oacc[j][1] * inv0);
            *reinterpret_cast<uint32_t*>(sm + (mt * 16 + g + 8) * CA_LD + qc + j * 8 + 2 * t) = pack_a16(oacc[j][2] * inv1, oacc[j][3] * inv1);
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (long)blockIdx.x * TOK_PER_CHUNK * D_MODEL);
    constexpr int OUT_U4 = D_MODEL / 8;            // 48 pieces per output row
#pragma unroll 2
    for (int i = tid; i < TOK_PER_CHUNK * OUT_U4; i += 256) {
        const int r = i / OUT_U4, c = i - r * OUT_U4;
        dst[i] = *reinterpret_cast<const uint4*>(sm + r * CA_LD + c * 8);
    }
}

static int g_chunk_attn_impl = 1;      // 1: mma.sync kernel, 0: CUDA-core kernel (kept for A/B tests)
void set_chunk_attention_impl(int impl) { g_chunk_attn_impl = impl; }

int launch_chunk_attention(const act16_t* qkv, act16_t* out, int n_chunks, cudaStream_t stream) {
    if (n_chunks > 0 && g_chunk_attn_impl == 1) {
        const size_t smem_mma = (size_t)TOK_PER_CHUNK * CA_LD * 2;
        static PerDeviceOnce attr_once_mma;
        KOCR_CUDA(opt_in_dynamic_smem(attr_once_mma, chunk_attention_mma_kernel, (int)smem_mma));
        chunk_attention_mma_kernel<<<n_chunks, 256, smem_mma, stream>>>(qkv, out);
        KOCR_CUDA(cudaGetLastError());
        return 0;
    }
    if (n_chunks == 0) return 0;
    const size_t smem = 8 * 2 * 32 * HEAD_DIM * sizeof(float);
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, chunk_attention_kernel, (int)smem));
    chunk_attention_kernel<<<n_chunks, 256, smem, stream>>>(qkv, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// LayerNorm / positional add over rows of 384.  Warp per row, 12 elements per lane.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_row_outputs(const float (&y)[12], long row, int lane, float* out_f32,
                                                  act16_t* out_a16, act16_t* out_lo) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int col = j * 128 + lane * 4;
        if (out_f32)
            *reinterpret_cast<float4*>(out_f32 + row * D_MODEL + col) =
                make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        if (out_a16) {
            const uint32_t p0 = pack_a16(y[4 * j], y[4 * j + 1]), p1 = pack_a16(y[4 * j + 2], y[4 * j + 3]);
            *reinterpret_cast<uint2*>(out_a16 + row * D_MODEL + col) = make_uint2(p0, p1);
            if (out_lo) {
                const uint32_t l0 = pack_a16(y[4 * j] - a16_lo(p0), y[4 * j + 1] - a16_hi(p0));
                const uint32_t l1 = pack_a16(y[4 * j + 2] - a16_lo(p1), y[4 * j + 3] - a16_hi(p1));
                *reinterpret_cast<uint2*>(out_lo + row * D_MODEL + col) = make_uint2(l0, l1);
            }
        }
    }
}

// Input row = sum of `nsplit` split-K partial results (x + s*rows*384) + in_bias + residual (each optional).
template <bool DO_LN>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ b, const float* __restrict__ pos,
                                                        const int* __restrict__ row_pos, float* out_f32,
                                                        act16_t* out_a16, act16_t* out_lo, int rows,
                                                        int nsplit, const float* __restrict__ in_bias,
                                                        const float* resid, float* out_tf32) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    float v[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // ld.global.cg: under PDL this kernel may be resident before its producer finishes, so data written by the
        // producer must not be read through L1 / the non-coherent path (stale lines from an earlier step)
        float4 a = __ldcg(reinterpret_cast<const float4*>(x + row * D_MODEL + j * 128 + lane * 4));
        for (int sp = 1; sp < nsplit; ++sp) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(x + ((long)sp * rows + row) * D_MODEL + j * 128 + lane * 4));
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (in_bias) {
            const float4 p = *reinterpret_cast<const float4*>(in_bias + j * 128 + lane * 4);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (resid) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(resid + row * D_MODEL + j * 128 + lane * 4));
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
    if (DO_LN) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 12; ++j) s += v[j];
        const float mean = warp_sum(s) * (1.f / D_MODEL);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 12; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D_MODEL) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 gg = *reinterpret_cast<const float4*>(g + j * 128 + lane * 4);
            const float4 bb = *reinterpret_cast<const float4*>(b + j * 128 + lane * 4);
            v[4 * j] = (v[4 * j] - mean) * rstd * gg.x + bb.x;
            v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * gg.y + bb.y;
            v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * gg.z + bb.z;
            v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * gg.w + bb.w;
        }
    }
    if (pos) {
        const long pr = row_pos[row];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 pp = *reinterpret_cast<const float4*>(pos + pr * D_MODEL + j * 128 + lane * 4);
            v[4 * j] += pp.x; v[4 * j + 1] += pp.y; v[4 * j + 2] += pp.z; v[4 * j + 3] += pp.w;
        }
    }
    store_row_outputs(v, row, lane, out_f32, out_a16, out_lo);
    if (out_tf32) {         // copy of the row as TF32 GEMM operand, rounded to nearest (the exact row stays the residual)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            *reinterpret_cast<float4*>(out_tf32 + row * D_MODEL + j * 128 + lane * 4) =
                make_float4(rna_tf32(v[4 * j]), rna_tf32(v[4 * j + 1]), rna_tf32(v[4 * j + 2]), rna_tf32(v[4 * j + 3]));
    }
}

int launch_layernorm(const float* x, const float* g, const float* b, const float* pos, const int* row_pos,
                     float* out_f32, act16_t* out_a16, act16_t* out_a16_lo, int rows,
                     cudaStream_t stream, int nsplit, const float* in_bias, const float* resid, float* out_tf32) {
    if (rows == 0) return 0;
    KOCR_CUDA(launch_kernel(layernorm_kernel<true>, dim3((rows + 7) / 8), dim3(256), 0, stream, x, g, b, pos, row_pos,
                            out_f32, out_a16, out_a16_lo, rows, nsplit, in_bias, resid, out_tf32));
    return 0;
}
int launch_add_pos(const float* x, const float* pos, const int* row_pos, float* out_f32, act16_t* out_a16,
                   act16_t* out_a16_lo, int rows, cudaStream_t stream) {
    if (rows == 0) return 0;
    layernorm_kernel<false><<<(rows + 7) / 8, 256, 0, stream>>>(x, nullptr, nullptr, pos, row_pos, out_f32, out_a16,
                                                                out_a16_lo, rows, 1, nullptr, nullptr, nullptr);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

static constexpr int TOK_LD = DEC_MAX + 1;        // ints per row of the token table (<sos> + 256 ids)

// Row compaction of the greedy decode loop.  Once half of the lines of a batch have emitted <eos>, the still-active
// rows of the tail [A, n_rows) are swapped into the finished rows ("holes") of the head [0, A), A = active count rounded
// up to the 128-row GEMM tile, and the loop continues with A rows: the 13 GEMMs of a position lose an M tile, the
// attention / LayerNorm grids shrink.  pairs[i] = (src row in the tail, dst row in the head); the two sets are disjoint.
// grid = (n_pairs, 5): y < 4 copies the live part [0, t) of one self-attention cache plane (layer, K|V) src -> dst (the
// hole's cache is dead); y == 4 swaps the small per-row state so that the finished line keeps its result in the tail.
__global__ void __launch_bounds__(256) decode_compact_kernel(const int2* __restrict__ pairs, int t, float* __restrict__ kcache,
                                                             float* __restrict__ vcache, size_t layer_stride,
                                                             int* __restrict__ tokens, int* __restrict__ lengths,
                                                             int* __restrict__ finished, int* __restrict__ tok_off,
                                                             int* __restrict__ line_T) {
    const int2 pr = pairs[blockIdx.x];
    const int src = pr.x, dst = pr.y;
    if (blockIdx.y < 4) {
        float* base = ((blockIdx.y & 1) ? vcache : kcache) + (size_t)(blockIdx.y >> 1) * layer_stride;
        const float4* s4 = reinterpret_cast<const float4*>(base + (size_t)src * DEC_MAX * D_MODEL);
        float4* d4 = reinterpret_cast<float4*>(base + (size_t)dst * DEC_MAX * D_MODEL);
        for (int i = threadIdx.x; i < t * (D_MODEL / 4); i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < TOK_LD; i += blockDim.x) {
            const int a = tokens[(size_t)src * TOK_LD + i], b = tokens[(size_t)dst * TOK_LD + i];
            tokens[(size_t)src * TOK_LD + i] = b;
            tokens[(size_t)dst * TOK_LD + i] = a;
        }
        if (threadIdx.x == 0) {
            int a;
            a = lengths[src]; lengths[src] = lengths[dst]; lengths[dst] = a;
            a = finished[src]; finished[src] = finished[dst]; finished[dst] = a;
            a = tok_off[src]; tok_off[src] = tok_off[dst]; tok_off[dst] = a;
            a = line_T[src]; line_T[src] = line_T[dst]; line_T[dst] = a;
        }
    }
}

int launch_decode_compact(const int* pairs, int n_pairs, int t, float* kcache, float* vcache, size_t layer_stride, int* tokens,
                          int* lengths, int* finished, int* tok_off, int* line_T, cudaStream_t stream) {
    if (n_pairs == 0) return 0;
    decode_compact_kernel<<<dim3(n_pairs, 5), 256, 0, stream>>>(reinterpret_cast<const int2*>(pairs), t, kcache, vcache, layer_stride,
                                                               tokens, lengths, finished, tok_off, line_T);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// Split-precision operand of the cross-attention K/V projection: row = [hi | lo | hi] with hi = a16(m), lo = a16(m - hi).
// Against weights packed as [w_hi | w_hi | w_lo] one 16-bit GEMM with K = 3 * 384 computes hi*w_hi + lo*w_hi + hi*w_lo,
// i.e. the fp32 product to ~20 bits - the 16-bit rounding of memory and weights in this one projection is what flips
// near-tied argmaxes against the fp32 reference (tests/parity/parity_rootcause.py).
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ m, act16_t* __restrict__ out, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    constexpr int CG = D_MODEL / 8;
    const int cg = (int)(idx % CG);
    const long row = idx / CG;
    const float4 a = reinterpret_cast<const float4*>(m + row * D_MODEL)[2 * cg];
    const float4 b = reinterpret_cast<const float4*>(m + row * D_MODEL)[2 * cg + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hi[j] = pack_a16(v[2 * j], v[2 * j + 1]);
        lo[j] = pack_a16(v[2 * j] - a16_lo(hi[j]), v[2 * j + 1] - a16_hi(hi[j]));
    }
    uint4* o = reinterpret_cast<uint4*>(out + row * (3 * D_MODEL));
    const uint4 h4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    o[cg] = h4;
    o[CG + cg] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    o[2 * CG + cg] = h4;
}

int launch_split3(const float* m, act16_t* out, long rows, cudaStream_t stream) {
    const long total = rows * (D_MODEL / 8);
    if (total == 0) return 0;
    split3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(m, out, total);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// Training-time memory layout of KhmerOCR.forward (se_model.py:262-273): every line padded to Tmax rows; real rows are
// the merged encoder output + global_pos (already in `xb`), pad rows are 0 + global_pos[t].  One thread per 8 channels.
__global__ void __launch_bounds__(256) pad_memory_kernel(const act16_t* __restrict__ xb, const float* __restrict__ global_pos,
                                                         const int* __restrict__ src_off, const int* __restrict__ line_T,
                                                         int Tmax, long total, act16_t* __restrict__ out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    constexpr int CG = D_MODEL / 8;
    const int cg = (int)(idx % CG);
    const long row = idx / CG;
    const int b = (int)(row / Tmax), t = (int)(row - (long)b * Tmax);
    uint4 v;
    if (t < line_T[b]) {
        v = reinterpret_cast<const uint4*>(xb + ((long)src_off[b] + t) * D_MODEL)[cg];
    } else {
        const float4 p0 = reinterpret_cast<const float4*>(global_pos + (long)t * D_MODEL)[2 * cg];
        const float4 p1 = reinterpret_cast<const float4*>(global_pos + (long)t * D_MODEL)[2 * cg + 1];
        v = make_uint4(pack_a16(p0.x, p0.y), pack_a16(p0.z, p0.w), pack_a16(p1.x, p1.y), pack_a16(p1.z, p1.w));
    }
    reinterpret_cast<uint4*>(out)[idx] = v;
}

int launch_pad_memory(const act16_t* xb, const float* global_pos, const int* src_off, const int* line_T, int n_lines,
                      int Tmax, act16_t* out, cudaStream_t stream) {
    const long total = (long)n_lines * Tmax * (D_MODEL / 8);
    if (total == 0) return 0;
    pad_memory_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(xb, global_pos, src_off, line_T, Tmax, total, out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// BiLSTM recurrence.  Cluster of 2 CTAs per (group of <= 8 lines, direction); CTA `rank` owns hidden
// units [rank*96, rank*96+96) i.e. 384 gate rows (i,f,g,o x 96), whose recurrent weights stay in
// shared memory (a16x2, k-pair major) for all timesteps.  Each step: gates = gin + W_hh h (fp32
// FMA), cell update, and the new half of h is written to both CTAs' shared memory (DSMEM) followed
// by one cluster barrier.
// whh_packed layout (built on the host): [dir][rank][kp = 0..95][row = 0..383] u32 = a16x2
// {W[grow][2kp], W[grow][2kp+1]}, grow = gate*192 + rank*96 + jj for row = gate*96 + jj.
// ------------------------------------------------------------------------------------------
static constexpr int LSTM_ROWS = 384;      // gate rows per CTA
static constexpr int LSTM_KP = LSTM_H / 2; // 96 k-pairs
static constexpr int LSTM_LPG = 8;         // lines per group

size_t bilstm_whh_packed_elems() { return (size_t)2 * 2 * LSTM_KP * LSTM_ROWS * 2; }   // a16 elements

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LSTM_ROWS, 1)
bilstm_kernel(const float* __restrict__ gin, const uint32_t* __restrict__ whh, const int* __restrict__ line_tok_off,
              const int* __restrict__ line_T, const LstmGroup* __restrict__ groups, float* __restrict__ mem_f32,
              act16_t* __restrict__ mem_a16, act16_t* __restrict__ mem_lo) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint32_t* s_w = reinterpret_cast<uint32_t*>(s_raw);                                 // [96][384]
    float* s_h = reinterpret_cast<float*>(s_raw + LSTM_KP * LSTM_ROWS * 4);              // [2][8][192]
    float* s_g = s_h + 2 * LSTM_LPG * LSTM_H;                                            // [8][384]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.y & 1;
    const LstmGroup grp = groups[blockIdx.y >> 1];
    const int tid = threadIdx.x;

    int tok_off[LSTM_LPG], T[LSTM_LPG];
    int maxT = 0;
#pragma unroll
    for (int l = 0; l < LSTM_LPG; ++l) {
        const int li = grp.line[l];
        tok_off[l] = li >= 0 ? line_tok_off[li] : 0;
        T[l] = li >= 0 ? line_T[li] : 0;
        maxT = max(maxT, T[l]);
    }
    {   // load this CTA's weight slice
        const uint4* src = reinterpret_cast<const uint4*>(whh + ((size_t)dir * 2 + rank) * LSTM_KP * LSTM_ROWS);
        uint4* dst = reinterpret_cast<uint4*>(s_w);
        for (int i = tid; i < LSTM_KP * LSTM_ROWS / 4; i += blockDim.x) dst[i] = src[i];
        for (int i = tid; i < 2 * LSTM_LPG * LSTM_H; i += blockDim.x) s_h[i] = 0.f;
    }
    float* peer_h = cluster.map_shared_rank(s_h, rank ^ 1);
    cluster.sync();

    const int gate = tid / 96, jj = tid - gate * 96;
    const int grow = dir * 768 + gate * LSTM_H + rank * 96 + jj;      // column of gin for this thread's row
    float c_state[2] = {0.f, 0.f};
    int cur = 0;
    for (int s = 0; s < maxT; ++s) {
        // ---- phase A: gate pre-activations for this CTA's 384 rows x 8 lines
        float gv[LSTM_LPG];
#pragma unroll
        for (int l = 0; l < LSTM_LPG; ++l) {
            gv[l] = 0.f;
            if (s < T[l]) {
                const int pos = dir == 0 ? s : T[l] - 1 - s;
                gv[l] = __ldg(gin + (long)(tok_off[l] + pos) * (8 * LSTM_H) + grow);
            }
        }
        float acc[LSTM_LPG];
#pragma unroll
        for (int l = 0; l < LSTM_LPG; ++l) acc[l] = 0.f;
        const float* hc = s_h + cur * LSTM_LPG * LSTM_H;
#pragma unroll 4
        for (int kp = 0; kp < LSTM_KP; kp += 2) {
            const uint32_t w01 = s_w[kp * LSTM_ROWS + tid];
            const uint32_t w23 = s_w[(kp + 1) * LSTM_ROWS + tid];
            const float w0 = a16_lo(w01), w1 = a16_hi(w01), w2 = a16_lo(w23), w3 = a16_hi(w23);
#pragma unroll
            for (int l = 0; l < LSTM_LPG; ++l) {
                const float4 h4 = *reinterpret_cast<const float4*>(hc + l * LSTM_H + 2 * kp);
                acc[l] = fmaf(w0, h4.x, acc[l]); acc[l] = fmaf(w1, h4.y, acc[l]);
                acc[l] = fmaf(w2, h4.z, acc[l]); acc[l] = fmaf(w3, h4.w, acc[l]);
            }
        }
#pragma unroll
        for (int l = 0; l < LSTM_LPG; ++l) s_g[l * LSTM_ROWS + tid] = acc[l] + gv[l];
        __syncthreads();
        // ---- phase B: cell update for (line, hidden unit) items; 768 items over 384 threads
        float* hn = s_h + (cur ^ 1) * LSTM_LPG * LSTM_H;
        float* hn_peer = peer_h + (cur ^ 1) * LSTM_LPG * LSTM_H;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int item = tid + rep * LSTM_ROWS;
            const int l = item / 96, j = item - l * 96;
            if (s < T[l]) {
                const float* gp = s_g + l * LSTM_ROWS + j;
                const float ig = sigmoid_acc(gp[0]), fg = sigmoid_acc(gp[96]);
                const float gg = tanhf(gp[192]), og = sigmoid_acc(gp[288]);
                const float c = fg * c_state[rep] + ig * gg;
                c_state[rep] = c;
                const float h = og * tanhf(c);
                const int hidx = l * LSTM_H + rank * 96 + j;
                hn[hidx] = h;
                hn_peer[hidx] = h;
                const int pos = dir == 0 ? s : T[l] - 1 - s;
                const long o = (long)(tok_off[l] + pos) * D_MODEL + dir * LSTM_H + rank * 96 + j;
                mem_f32[o] = h;
                if (mem_a16) {
                    const act16_t hb = to_a16(h);
                    mem_a16[o] = hb;
                    if (mem_lo) mem_lo[o] = to_a16(h - from_a16(hb));
                }
            }
        }
        cluster.sync();
        cur ^= 1;
    }
}

int launch_bilstm(const float* gin, const act16_t* whh_packed, const int* line_tok_off, const int* line_T,
                  const LstmGroup* groups, int n_groups, float* mem_f32, act16_t* mem_a16,
                  act16_t* mem_a16_lo, cudaStream_t stream) {
    if (n_groups == 0) return 0;
    const size_t smem = (size_t)LSTM_KP * LSTM_ROWS * 4 + 2 * LSTM_LPG * LSTM_H * 4 + LSTM_LPG * LSTM_ROWS * 4;
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, bilstm_kernel, (int)smem));
    bilstm_kernel<<<dim3(2, n_groups * 2), LSTM_ROWS, smem, stream>>>(
        gin, reinterpret_cast<const uint32_t*>(whh_packed), line_tok_off, line_T, groups, mem_f32, mem_a16,
        mem_a16_lo);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// BiLSTM recurrence, tensor-core version.  Cluster of 2 CTAs per (group of <= 16 lines, direction), 12 warps
// per CTA.  Warp w of CTA `rank` owns hidden units [rank*96 + w*8, +8): its 32 gate rows x 192 recurrent weights
// live in REGISTERS as mma.sync m16n8k16 A-fragments for all timesteps (2 m-tiles {i|f}, {g|o} x 12 k-steps).
// Per step: D[32 gate rows x 16 lines] = W_hh[32 x 192] . h[192 x 16 lines] with h split into a16 hi + lo parts
// (both accumulated, fp32 accumulators), + the input projection; the fragment layout leaves all four gates of a
// (unit, line) cell in one thread, so the cell update needs no exchange; the new h is written to both CTAs'
// shared memory (DSMEM) and one cluster barrier closes the step.
// The step is M = 32 rows per warp and strictly sequential: latency-bound, far too small for a tcgen05 tile.
// whh_mma layout (host): [dir][rank][warp][mtile][kstep][lane] uint4 = the A fragment registers.
// ------------------------------------------------------------------------------------------
static constexpr int LM_LPG = 16;          // lines per group
static constexpr int LM_HS = 200;          // padded row stride (a16) of the h buffers: conflict-free B loads
static constexpr int LM_THREADS = 384;

size_t bilstm_whh_mma_elems() { return (size_t)2 * 2 * 12 * 2 * 12 * 32 * 8; }   // a16 elements

// Gate non-linearities of the tensor-core recurrence on the MUFU ex2 / rcp units (errors ~1e-6).  The branchy accurate
// expf / tanhf sequences were the longest part of a recurrence step; `tanh.approx.f32` (one MUFU op, but a relative error
// of 2^-11 = 5e-4 on every cell and hidden value) was as large as the whole 16-bit error budget of the memory, so tanh is
// computed as 1 - 2 / (1 + e^2x) instead: two MUFU ops, absolute error ~1e-7.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * fminf(x, 15.f));            // (e^30 is finite; tanh(15) == 1 in fp32)
    return 1.f - __fdividef(2.f, 1.f + e);
}

__device__ __forceinline__ void mma_a16_16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32." KOCR_MMA_A16 "." KOCR_MMA_A16 ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LM_THREADS, 1)
bilstm_mma_kernel(const float* __restrict__ gin, const uint4* __restrict__ whh, const int* __restrict__ line_tok_off,
                  const int* __restrict__ line_T, const LstmGroup16* __restrict__ groups, float* __restrict__ mem_f32,
                  act16_t* __restrict__ mem_a16, act16_t* __restrict__ mem_lo) {
    __shared__ __align__(16) act16_t s_h[2][2][LM_LPG][LM_HS];     // [buffer][hi/lo][line][k]
    extern __shared__ __align__(16) float s_gin[];                       // [2][16][384] input-projection prefetch
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.y & 1;
    const LstmGroup16 grp = groups[blockIdx.y >> 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;

    // recurrent weights -> registers (96 per thread)
    uint4 afrag[2][12];
    {
        const uint4* src = whh + ((size_t)((dir * 2 + rank) * 12 + warp) * 2 * 12) * 32 + lane;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) afrag[mt][ks] = __ldg(src + (mt * 12 + ks) * 32);
    }
    // this thread's four (unit, line) cells: unit = g, lines nt*8 + 2*tig + {0, 1}
    int cell_off[4], cell_T[4];
    int maxT = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int li = grp.line[(c >> 1) * 8 + 2 * tig + (c & 1)];
        cell_off[c] = li >= 0 ? line_tok_off[li] : 0;
        cell_T[c] = li >= 0 ? line_T[li] : 0;
    }
    for (int l = 0; l < LM_LPG; ++l) {
        const int li = grp.line[l];
        if (li >= 0) maxT = max(maxT, line_T[li]);
    }
    for (int i = tid; i < 2 * 2 * LM_LPG * LM_HS; i += LM_THREADS) (&s_h[0][0][0][0])[i] = to_a16(0.f);
    act16_t* peer = cluster.map_shared_rank(&s_h[0][0][0][0], rank ^ 1);
    cluster.sync();

    const int hid = rank * 96 + warp * 8 + g;                 // hidden unit of this thread's cells
    const int gcol = dir * 4 * LSTM_H + hid;                  // gin column of gate 0 (i); gates are LSTM_H apart
    float c_state[4] = {0.f, 0.f, 0.f, 0.f};
    int cur = 0;
    // The input projection of step s+1 is fetched with cp.async (global -> shared, no registers) while step s
    // computes, so its HBM latency never sits on the recurrence's critical path.
    auto prefetch_gin = [&](int step) {
        float* dst = s_gin + (step & 1) * 16 * LM_THREADS + tid;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (step < cell_T[c]) {
                const int pos = dir == 0 ? step : cell_T[c] - 1 - step;
                const float* gp = gin + (long)(cell_off[c] + pos) * (8 * LSTM_H) + gcol;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + (c * 4 + q) * LM_THREADS)),
                                 "l"(gp + q * LSTM_H) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch_gin(0);
    for (int s = 0; s < maxT; ++s) {
        prefetch_gin(s + 1);            // (an empty group once s + 1 >= every T)
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
        const act16_t* hhi = &s_h[cur][0][0][0];
        const act16_t* hlo = &s_h[cur][1][0][0];
#pragma unroll
        for (int ks = 0; ks < 12; ++ks) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int o = (nt * 8 + g) * LM_HS + ks * 16 + 2 * tig;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(hhi + o);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(hhi + o + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(hlo + o);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(hlo + o + 8);
                mma_a16_16816(acc[0][nt], afrag[0][ks], bh0, bh1);
                mma_a16_16816(acc[1][nt], afrag[1][ks], bh0, bh1);
                mma_a16_16816(acc[0][nt], afrag[0][ks], bl0, bl1);
                mma_a16_16816(acc[1][nt], afrag[1][ks], bl0, bl1);
            }
        }
        // cell update: acc[0][nt] = {i(l0), i(l1), f(l0), f(l1)}, acc[1][nt] = {g(l0), g(l1), o(l0), o(l1)}
        asm volatile("cp.async.wait_group 1;" ::: "memory");        // this step's input projection has landed
        const float* gsrc = s_gin + (s & 1) * 16 * LM_THREADS + tid;
        const int nxt = cur ^ 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (s < cell_T[c]) {
                const int nt = c >> 1, e = c & 1;
                const float ig = sigmoid_fast(acc[0][nt][e] + gsrc[(c * 4 + 0) * LM_THREADS]);
                const float fg = sigmoid_fast(acc[0][nt][2 + e] + gsrc[(c * 4 + 1) * LM_THREADS]);
                const float gg = tanh_fast(acc[1][nt][e] + gsrc[(c * 4 + 2) * LM_THREADS]);
                const float og = sigmoid_fast(acc[1][nt][2 + e] + gsrc[(c * 4 + 3) * LM_THREADS]);
                const float cc = fg * c_state[c] + ig * gg;
                c_state[c] = cc;
                const float h = og * tanh_fast(cc);
                const act16_t hb = to_a16(h);
                const act16_t lb = to_a16(h - from_a16(hb));
                const int line = nt * 8 + 2 * tig + e;
                const int ohi = ((nxt * 2 + 0) * LM_LPG + line) * LM_HS + hid;
                const int olo = ((nxt * 2 + 1) * LM_LPG + line) * LM_HS + hid;
                (&s_h[0][0][0][0])[ohi] = hb; (&s_h[0][0][0][0])[olo] = lb;
                peer[ohi] = hb; peer[olo] = lb;
                const int pos = dir == 0 ? s : cell_T[c] - 1 - s;
                const long o = (long)(cell_off[c] + pos) * D_MODEL + dir * LSTM_H + hid;
                mem_f32[o] = h;
                if (mem_a16) { mem_a16[o] = hb; if (mem_lo) mem_lo[o] = lb; }
            }
        }
        cluster.sync();
        cur = nxt;
    }
}

int launch_bilstm_mma(const float* gin, const act16_t* whh_mma, const int* line_tok_off, const int* line_T,
                      const LstmGroup16* groups, int n_groups, float* mem_f32, act16_t* mem_a16,
                      act16_t* mem_a16_lo, cudaStream_t stream) {
    if (n_groups == 0) return 0;
    const size_t smem = 2 * 16 * LM_THREADS * sizeof(float);
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, bilstm_mma_kernel, (int)smem));
    bilstm_mma_kernel<<<dim3(2, n_groups * 2), LM_THREADS, smem, stream>>>(
        gin, reinterpret_cast<const uint4*>(whh_mma), line_tok_off, line_T, groups, mem_f32, mem_a16, mem_a16_lo);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Decoder step kernels (one launch each per generated position t, batched over lines).
// tokens: int32 [n_lines, DEC_MAX + 1]; tokens[l][0] = <sos>.
// ------------------------------------------------------------------------------------------

__global__ void dec_embed_kernel(const int* __restrict__ tokens, const int* __restrict__ step_base, int step_off,
                                 const float* __restrict__ tok_emb,
                                 const float* __restrict__ pos_emb, float* __restrict__ x, float* __restrict__ x_tf32,
                                 act16_t* __restrict__ xb, act16_t* __restrict__ xb_lo, int n_lines) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // n_lines * 96 float4
    pdl_trigger();
    pdl_wait();
    if (idx >= n_lines * (D_MODEL / 4)) return;
    const int l = idx / (D_MODEL / 4), c4 = idx - l * (D_MODEL / 4);
    const int t = __ldcg(step_base) + step_off;
    const int tok = __ldcg(tokens + l * TOK_LD + t);
    const float4 e = reinterpret_cast<const float4*>(tok_emb + (long)tok * D_MODEL)[c4];
    const float4 p = reinterpret_cast<const float4*>(pos_emb + (long)t * D_MODEL)[c4];
    const float4 v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
    reinterpret_cast<float4*>(x)[idx] = v;
    if (x_tf32) reinterpret_cast<float4*>(x_tf32)[idx] = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
    if (xb == nullptr) return;
    const uint32_t p0 = pack_a16(v.x, v.y), p1 = pack_a16(v.z, v.w);
    reinterpret_cast<uint2*>(xb)[idx] = make_uint2(p0, p1);
    if (xb_lo)
        reinterpret_cast<uint2*>(xb_lo)[idx] = make_uint2(pack_a16(v.x - a16_lo(p0), v.y - a16_hi(p0)),
                                                          pack_a16(v.z - a16_lo(p1), v.w - a16_hi(p1)));
}

int launch_dec_embed(const int* tokens, const int* step_base, int step_off, const float* tok_emb,
                     const float* pos_emb, float* x, float* x_tf32, act16_t* xb, act16_t* xb_lo, int n_lines,
                     cudaStream_t stream) {
    const int total = n_lines * (D_MODEL / 4);
    KOCR_CUDA(launch_kernel(dec_embed_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, tokens, step_base, step_off,
                            tok_emb, pos_emb, x, x_tf32, xb, xb_lo, n_lines));
    return 0;
}

// attention outputs are consumed only as the A operand of the TF32 out-projection GEMM: round to nearest here
__device__ __forceinline__ void store_attn_out(float o, long idx, float* out) { out[idx] = rna_tf32(o); }

// Causal self-attention for the newest position t against the cache (keys 0..t); keys whose token is
// <pad> are masked (tgt_key_padding_mask, se_model.py:190).  CTA per line, warp per head.
// The cache is fp32: rounding the self-attention K/V to a16 flips ~2 % of the fixture lines against the fp32
// reference (near-tie argmaxes late in long sequences), while a16 cross-attention K/V is harmless.
__global__ void __launch_bounds__(256) dec_self_attn_kernel(const float* __restrict__ qkv,
                                                            float* __restrict__ kcache,
                                                            float* __restrict__ vcache,
                                                            const int* __restrict__ tokens,
                                                            const int* __restrict__ step_base, int step_off,
                                                            const int* __restrict__ finished,
                                                            float* __restrict__ out, int nsplit, int n_lines,
                                                            const float* __restrict__ bias) {
    __shared__ float s_q[D_MODEL];
    __shared__ float s_p[N_HEAD][DEC_MAX];
    const int l = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    pdl_wait();
    if (__ldcg(finished + l)) return;             // line already emitted <eos>: nothing downstream reads it
    const int t = __ldcg(step_base) + step_off;
    const float* row = qkv + (long)l * 3 * D_MODEL;
    float* kc = kcache + (long)l * DEC_MAX * D_MODEL;
    float* vc = vcache + (long)l * DEC_MAX * D_MODEL;
    for (int i = tid; i < D_MODEL; i += blockDim.x) {
        float qv = bias[i], kv_ = bias[D_MODEL + i], vv = bias[2 * D_MODEL + i];
        for (int sp = 0; sp < nsplit; ++sp) {               // sum the split-K partial results of the QKV projection
            const float* r = row + (long)sp * n_lines * 3 * D_MODEL;
            qv += __ldcg(r + i); kv_ += __ldcg(r + D_MODEL + i); vv += __ldcg(r + 2 * D_MODEL + i);
        }
        s_q[i] = qv * rsqrtf((float)HEAD_DIM);
        kc[(long)t * D_MODEL + i] = kv_;
        vc[(long)t * D_MODEL + i] = vv;
    }
    __syncthreads();
    const int nk = t + 1;
    const float* qh = s_q + warp * HEAD_DIM;
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) {
        float acc = -INFINITY;
        if (__ldcg(tokens + l * TOK_LD + j) != 0) {
            const float4* kp = reinterpret_cast<const float4*>(kc + (long)j * D_MODEL + warp * HEAD_DIM);
            acc = 0.f;
#pragma unroll
            for (int i = 0; i < HEAD_DIM / 4; ++i) {
                const float4 k4 = __ldcg(kp + i);
                acc = fmaf(qh[4 * i], k4.x, acc); acc = fmaf(qh[4 * i + 1], k4.y, acc);
                acc = fmaf(qh[4 * i + 2], k4.z, acc); acc = fmaf(qh[4 * i + 3], k4.w, acc);
            }
        }
        s_p[warp][j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
        const float e = __expf(s_p[warp][j] - mx);
        s_p[warp][j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    if (lane < HEAD_DIM / 2) {           // lanes 0..23 own one pair of the 48 head dims
        float o0 = 0.f, o1 = 0.f;
        const float2* vp = reinterpret_cast<const float2*>(vc + warp * HEAD_DIM) + lane;
        int j = 0;
        for (; j + 8 <= nk; j += 8) {    // 8 value rows in flight: the loop is a chain of L2 latencies otherwise
            float2 v2[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v2[u] = __ldcg(vp + (long)(j + u) * (D_MODEL / 2));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                o0 = fmaf(s_p[warp][j + u], v2[u].x, o0);
                o1 = fmaf(s_p[warp][j + u], v2[u].y, o1);
            }
        }
        for (; j < nk; ++j) {
            const float2 v2 = __ldcg(vp + (long)j * (D_MODEL / 2));
            o0 = fmaf(s_p[warp][j], v2.x, o0);
            o1 = fmaf(s_p[warp][j], v2.y, o1);
        }
        const long oi = (long)l * D_MODEL + warp * HEAD_DIM + 2 * lane;
        store_attn_out(o0 * inv, oi, out);
        store_attn_out(o1 * inv, oi + 1, out);
    }
}

int launch_dec_self_attn(const float* qkv, float* kcache, float* vcache, const int* tokens,
                         const int* step_base, int step_off, const int* finished, float* out, int n_lines,
                         cudaStream_t stream, int nsplit, const float* bias) {
    KOCR_CUDA(launch_kernel(dec_self_attn_kernel, dim3(n_lines), dim3(256), 0, stream, qkv, kcache, vcache, tokens,
                            step_base, step_off, finished, out, nsplit, n_lines, bias));
    return 0;
}

// Cross-attention of one query per line over the line's T memory tokens (K/V precomputed once per line,
// a16 [Mtok, 1536]: layer*768 + {0: K, 384: V}).  CTA per line.  A K (or V) row of all 8 heads is 768
// contiguous bytes: a warp reads one key per iteration, 24 B (12 dims, a quarter of one head) per lane, so
// every global read is a fully coalesced 768-byte burst; the 8 warps stride over the keys.
__global__ void __launch_bounds__(256) dec_cross_attn_kernel(const float* __restrict__ q,
                                                             const act16_t* __restrict__ kv, int layer,
                                                             const int* __restrict__ line_tok_off,
                                                             const int* __restrict__ line_T, int max_T,
                                                             const int* __restrict__ finished,
                                                             float* __restrict__ out, int nsplit, int n_lines,
                                                             const float* __restrict__ bias) {
    extern __shared__ __align__(16) float s_dyn[];       // [8 heads][max_T] scores | [8 warps][384] partial outputs
    const int l = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    pdl_wait();
    if (__ldcg(finished + l)) return;
    float* s_p = s_dyn;
    float* s_o = s_dyn + (long)N_HEAD * max_T;
    __shared__ float s_inv[N_HEAD];
    const int T = line_T[l];
    const int head = lane >> 2;                  // 4 lanes per head, 12 dims each
    const uint2* kbase = reinterpret_cast<const uint2*>(kv + (long)line_tok_off[l] * (4 * D_MODEL) + layer * 2 * D_MODEL) + lane * 3;
    const uint2* vbase = kbase + (D_MODEL * 2) / 8;           // +768 bytes
    const long row_stride = (4 * D_MODEL * 2) / 8;            // uint2 per token row
    float qv[12];
    {
        const float sc = rsqrtf((float)HEAD_DIM);
        const float4* qp = reinterpret_cast<const float4*>(q + (long)l * D_MODEL + lane * 12);
        const float4* bp = reinterpret_cast<const float4*>(bias + lane * 12);
        float4 a = bp[0], b = bp[1], c = bp[2];
        for (int sp = 0; sp < nsplit; ++sp) {               // sum the split-K partial results of the Q projection
            const float4* r = qp + (long)sp * n_lines * (D_MODEL / 4);
            const float4 x = __ldcg(r), y = __ldcg(r + 1), z = __ldcg(r + 2);
            a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
            b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
            c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
        }
        qv[0] = a.x * sc; qv[1] = a.y * sc; qv[2] = a.z * sc; qv[3] = a.w * sc;
        qv[4] = b.x * sc; qv[5] = b.y * sc; qv[6] = b.z * sc; qv[7] = b.w * sc;
        qv[8] = c.x * sc; qv[9] = c.y * sc; qv[10] = c.z * sc; qv[11] = c.w * sc;
    }
    // ---- scores
    for (int j0 = warp; j0 < T; j0 += 32) {           // 4 keys in flight per warp
        uint2 k[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
            if (j < T) {
                const uint2* kp = kbase + (long)j * row_stride;
                k[u][0] = __ldg(kp); k[u][1] = __ldg(kp + 1); k[u][2] = __ldg(kp + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
            if (j < T) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc = fmaf(qv[4 * i], a16_lo(k[u][i].x), acc); acc = fmaf(qv[4 * i + 1], a16_hi(k[u][i].x), acc);
                    acc = fmaf(qv[4 * i + 2], a16_lo(k[u][i].y), acc); acc = fmaf(qv[4 * i + 3], a16_hi(k[u][i].y), acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if ((lane & 3) == 0) s_p[head * max_T + j] = acc;
            }
        }
    }
    __syncthreads();
    // ---- softmax: warp w owns head w
    {
        float* ph = s_p + (long)warp * max_T;
        float mx = -INFINITY;
        for (int j = lane; j < T; j += 32) mx = fmaxf(mx, ph[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) {
            const float e = __expf(ph[j] - mx);
            ph[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) s_inv[warp] = 1.f / sum;
    }
    __syncthreads();
    // ---- P.V partial sums over this warp's keys
    float o[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = 0.f;
    const float* ph = s_p + (long)head * max_T;
    for (int j0 = warp; j0 < T; j0 += 32) {
        uint2 v[4][3];
        float pj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
            pj[u] = 0.f;
            if (j < T) {
                const uint2* vp = vbase + (long)j * row_stride;
                v[u][0] = __ldg(vp); v[u][1] = __ldg(vp + 1); v[u][2] = __ldg(vp + 2);
                pj[u] = ph[j];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + 8 * u < T) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    o[4 * i] = fmaf(pj[u], a16_lo(v[u][i].x), o[4 * i]); o[4 * i + 1] = fmaf(pj[u], a16_hi(v[u][i].x), o[4 * i + 1]);
                    o[4 * i + 2] = fmaf(pj[u], a16_lo(v[u][i].y), o[4 * i + 2]); o[4 * i + 3] = fmaf(pj[u], a16_hi(v[u][i].y), o[4 * i + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s_o[warp * D_MODEL + lane * 12 + i] = o[i];
    __syncthreads();
    for (int d = tid; d < D_MODEL; d += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_o[w * D_MODEL + d];
        store_attn_out(acc * s_inv[d / HEAD_DIM], (long)l * D_MODEL + d, out);
    }
}

// One-pass ("flash-decoding") version of the same cross-attention: every warp walks its keys ONCE, loading the K and
// the V row of a key together (2 x 768 coalesced bytes, 4 keys = 6 KB in flight per warp), keeps a running
// (max, sum, weighted V) per head with the online-softmax rescaling, and the 8 warps are merged through shared
// memory at the end.  No score buffer, no intermediate __syncthreads, twice the bytes in flight: this kernel is the
// HBM-bound part of a decode position (it streams the K/V of every active line at every position).
__global__ void __launch_bounds__(256, 3) dec_cross_attn_flash_kernel(const float* __restrict__ q,
                                                                   const act16_t* __restrict__ kv, int layer,
                                                                   const int* __restrict__ line_tok_off,
                                                                   const int* __restrict__ line_T,
                                                                   const int* __restrict__ finished,
                                                                   float* __restrict__ out, int nsplit, int n_lines,
                                                                   const float* __restrict__ bias) {
    __shared__ float s_m[8][N_HEAD], s_l[8][N_HEAD];
    __shared__ __align__(16) float s_o[8][D_MODEL];
    const int l = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    pdl_wait();
    if (__ldcg(finished + l)) return;
    const int T = line_T[l];
    const int head = lane >> 2;                  // 4 lanes per head, 12 dims each
    const uint2* kbase = reinterpret_cast<const uint2*>(kv + (long)line_tok_off[l] * (4 * D_MODEL) + layer * 2 * D_MODEL) + lane * 3;
    const uint2* vbase = kbase + (D_MODEL * 2) / 8;           // +768 bytes
    const long row_stride = (4 * D_MODEL * 2) / 8;            // uint2 per token row
    float qv[12];
    {
        const float sc = rsqrtf((float)HEAD_DIM) * 1.4426950408889634f;      // scores in the base-2 domain
        const float4* qp = reinterpret_cast<const float4*>(q + (long)l * D_MODEL + lane * 12);
        const float4* bp = reinterpret_cast<const float4*>(bias + lane * 12);
        float4 a = bp[0], b = bp[1], c = bp[2];
        for (int sp = 0; sp < nsplit; ++sp) {               // sum the split-K partial results of the Q projection
            const float4* r = qp + (long)sp * n_lines * (D_MODEL / 4);
            const float4 x = __ldcg(r), y = __ldcg(r + 1), z = __ldcg(r + 2);
            a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
            b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
            c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
        }
        qv[0] = a.x * sc; qv[1] = a.y * sc; qv[2] = a.z * sc; qv[3] = a.w * sc;
        qv[4] = b.x * sc; qv[5] = b.y * sc; qv[6] = b.z * sc; qv[7] = b.w * sc;
        qv[8] = c.x * sc; qv[9] = c.y * sc; qv[10] = c.z * sc; qv[11] = c.w * sc;
    }
    float m = -INFINITY, lsum = 0.f, o[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = 0.f;
    for (int j0 = warp; j0 < T; j0 += 32) {           // keys j0, j0+8, j0+16, j0+24 of this warp in flight together
        uint2 k[4][3], v[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
            if (j < T) {
                const uint2* kp = kbase + (long)j * row_stride;
                const uint2* vp = vbase + (long)j * row_stride;
                k[u][0] = __ldg(kp); k[u][1] = __ldg(kp + 1); k[u][2] = __ldg(kp + 2);
                v[u][0] = __ldg(vp); v[u][1] = __ldg(vp + 1); v[u][2] = __ldg(vp + 2);
            }
        }
        float sc4[4];
        float mx = m;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float acc = -INFINITY;
            if (j0 + 8 * u < T) {
                acc = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc = fmaf(qv[4 * i], a16_lo(k[u][i].x), acc); acc = fmaf(qv[4 * i + 1], a16_hi(k[u][i].x), acc);
                    acc = fmaf(qv[4 * i + 2], a16_lo(k[u][i].y), acc); acc = fmaf(qv[4 * i + 3], a16_hi(k[u][i].y), acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);        // j0 + 8u < T is warp-uniform
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            }
            sc4[u] = acc;
            mx = fmaxf(mx, acc);
        }
        const float resc = exp2f(m - mx);           // first iteration: exp2(-inf) = 0 (key j0 always exists, mx is finite)
        m = mx;
        lsum *= resc;
#pragma unroll
        for (int i = 0; i < 12; ++i) o[i] *= resc;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + 8 * u < T) {
                const float pj = exp2f(sc4[u] - mx);
                lsum += pj;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    o[4 * i] = fmaf(pj, a16_lo(v[u][i].x), o[4 * i]); o[4 * i + 1] = fmaf(pj, a16_hi(v[u][i].x), o[4 * i + 1]);
                    o[4 * i + 2] = fmaf(pj, a16_lo(v[u][i].y), o[4 * i + 2]); o[4 * i + 3] = fmaf(pj, a16_hi(v[u][i].y), o[4 * i + 3]);
                }
            }
        }
    }
    // ---- merge the 8 warps (a warp with no key at all has m = -inf, l = 0, o = 0 and drops out)
    if ((lane & 3) == 0) { s_m[warp][head] = m; s_l[warp][head] = lsum; }
#pragma unroll
    for (int i = 0; i < 12; i += 4)
        *reinterpret_cast<float4*>(&s_o[warp][lane * 12 + i]) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
    __syncthreads();
    for (int d = tid; d < D_MODEL; d += 256) {
        const int h = d / HEAD_DIM;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w][h]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float f = exp2f(s_m[w][h] - M);
            num = fmaf(s_o[w][d], f, num);
            den = fmaf(s_l[w][h], f, den);
        }
        store_attn_out(num / den, (long)l * D_MODEL + d, out);
    }
}

static int g_dec_cross_impl = 1;       // 1: one-pass kernel, 0: three-phase kernel (kept for A/B tests)
void set_dec_cross_attention_impl(int impl) { g_dec_cross_impl = impl; }

int launch_dec_cross_attn(const float* q, const act16_t* kv, int layer, const int* line_tok_off,
                          const int* line_T, int max_T, const int* finished, float* out, int n_lines,
                          cudaStream_t stream, int nsplit, const float* bias) {
    if (g_dec_cross_impl == 1) {
        KOCR_CUDA(launch_kernel(dec_cross_attn_flash_kernel, dim3(n_lines), dim3(256), 0, stream, q, kv, layer, line_tok_off,
                                line_T, finished, out, nsplit, n_lines, bias));
        return 0;
    }
    const size_t smem = ((size_t)N_HEAD * max_T + 8 * D_MODEL) * sizeof(float);
    KOCR_CHECK(smem <= 200 * 1024, "cross-attention: memory length %d too long for shared memory", max_T);
    static PerDeviceOnce attr_once;     // set outside any stream capture (the first decode group of a handle runs eagerly)
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, dec_cross_attn_kernel, 200 * 1024));
    KOCR_CUDA(launch_kernel(dec_cross_attn_kernel, dim3(n_lines), dim3(256), smem, stream, q, kv, layer, line_tok_off,
                            line_T, max_T, finished, out, nsplit, n_lines, bias));
    return 0;
}

// argmax over the 124 real logits (ties -> lowest index, torch.argmax) and the greedy bookkeeping of
// predictor.py:90-97 (stop BEFORE appending <eos>).  Warp per line.  Optional: copy the logits row into the
// trace [n_lines, DEC_MAX, 128]; teacher forcing (the next id comes from `forced`, lines never finish).
__global__ void __launch_bounds__(128) dec_argmax_kernel(const float* __restrict__ logits, int* __restrict__ tokens,
                                                         int* __restrict__ lengths, int* __restrict__ finished,
                                                         int* __restrict__ n_active, const int* __restrict__ step_base,
                                                         int step_off, int n_lines, const int* __restrict__ forced,
                                                         float* __restrict__ trace, int nsplit,
                                                         const float* __restrict__ bias) {
    const int l = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (l >= n_lines) return;
    if (__ldcg(finished + l)) return;
    const int t = __ldcg(step_base) + step_off;
    float4 v4 = reinterpret_cast<const float4*>(bias)[lane];
    for (int sp = 0; sp < nsplit; ++sp) {                   // sum the split-K partial results of out_proj
        const float4 p = __ldcg(reinterpret_cast<const float4*>(logits + ((long)sp * n_lines + l) * VOCAB_PAD) + lane);
        v4.x += p.x; v4.y += p.y; v4.z += p.z; v4.w += p.w;
    }
    if (trace) reinterpret_cast<float4*>(trace + ((long)l * DEC_MAX + t) * VOCAB_PAD)[lane] = v4;
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = lane * 4 + j;
        if (i < VOCAB && v[j] > best) { best = v[j]; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane != 0) return;
    if (bi == 0x7fffffff) bi = 3;        // no finite maximum (NaN logits): end the line instead of emitting an out-of-range id
    if (forced) {
        tokens[l * TOK_LD + t + 1] = forced[l * TOK_LD + t + 1];
        lengths[l] = t + 2;
        return;
    }
    if (bi == 3) {                       // <eos>
        finished[l] = 1;
    } else {
        tokens[l * TOK_LD + t + 1] = bi;
        lengths[l] = t + 2;
        if (t + 1 >= DEC_MAX) finished[l] = 1; else atomicAdd(n_active + t, 1);
    }
}

int launch_dec_argmax(const float* logits, int* tokens, int* lengths, int* finished, int* n_active,
                      const int* step_base, int step_off, int n_lines, const int* forced, float* trace,
                      cudaStream_t stream, int nsplit, const float* bias) {
    KOCR_CUDA(launch_kernel(dec_argmax_kernel, dim3((n_lines + 3) / 4), dim3(128), 0, stream, logits, tokens, lengths,
                            finished, n_active, step_base, step_off, n_lines, forced, trace, nsplit, bias));
    return 0;
}

__global__ void dec_bump_kernel(int* step_base, int n) {
    pdl_trigger();
    pdl_wait();
    *step_base = __ldcg(step_base) + n;
}
int launch_dec_bump(int* step_base, int n, cudaStream_t stream) {
    KOCR_CUDA(launch_kernel(dec_bump_kernel, dim3(1), dim3(1), 0, stream, step_base, n));
    return 0;
}

}  // namespace kocr
