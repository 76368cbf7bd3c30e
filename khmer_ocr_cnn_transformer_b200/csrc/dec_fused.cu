// Fused kernels of the greedy decode loop (TransformerDecoderWrapper.forward + OCRPredictor._greedy_decode,
// se_model.py:182-208, predictor.py:85-99).  A generated position used to be 25 launches; every launch boundary costs
// a few microseconds of dependent-chain latency per batch and, with a dozen batches in flight, a measurable share of the
// whole GPU's time (profiles/r02/decode_attribution.md).  Two fusions take the count to 18:
//
//   dec_gemm_ln_kernel        out-projection (self-attention / cross-attention) or FFN2  +  bias + residual + LayerNorm
//                             = post-norm sub-layer tail  x <- LN(x + W a + b)   (replaces: split-K GEMM, LayerNorm kernel).
//                             A cluster of 3 CTAs owns a 128-row tile, one 128-column slab of the 384 outputs each (one CTA
//                             for all 384 columns was tried first: a single SM cannot stream the weights fast enough);
//                             the row statistics are combined through distributed shared memory.
//   dec_out_argmax_kernel     output projection (384 -> 124)  +  argmax  +  greedy bookkeeping (stop BEFORE <eos>)
//                             (replaces: split-K GEMM, dec_argmax_kernel; a first version also embedded the next position,
//                             one row at a time per warp - a serial chain of L2 latencies, slower than the 3 us kernel it replaced)
//
// Both are single-tile tcgen05 GEMMs (TF32 operands, fp32 accumulation in TMEM, operands staged by TMA) whose epilogue
// threads own one accumulator row each (tcgen05.ld 32x32b: lane = row), which is exactly the access pattern a row-wise
// LayerNorm / argmax wants: no cross-thread reduction at all.
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace kocr {

static constexpr int DF_BM = 128;
static constexpr int DF_KE = 32;                       // fp32 elements per 128-byte K block
static constexpr int DF_STAGES = 3;
static constexpr int DF_THREADS = 192;                 // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue (quadrant = warp & 3)
static constexpr int TOK_LD = DEC_MAX + 1;

template <int N> struct DfCfg {
    static constexpr int A_BYTES = DF_BM * 128;                    // 16 KB
    static constexpr int B_BYTES = N * 128;                        // 48 KB (N = 384) / 16 KB (N = 128)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = N <= 128 ? 128 : 512;
    static constexpr int SMEM_BYTES = DF_STAGES * STAGE_BYTES + 1024 + 256;
};

struct DfCommon {          // producer / MMA halves shared by the two kernels
    uint8_t* smem;
    uint64_t *full_bar, *empty_bar, *tmem_full;
    uint32_t tmem_base;
};

// Sets up barriers + TMEM and runs the TMA producer (warp 0) and the MMA issuer (warp 1) for ONE 128 x N tile starting at
// row m0; returns true in the epilogue warps right away (they overlap their own prefetches with the main loop and then
// call df_wait_accumulator), false in the producer / MMA warps once those are done.  N is a multiple of 128: the B
// tile is loaded and multiplied as N / 128 slabs of 128 weight rows (TMA boxes hold at most 256 rows).
template <int N>
__device__ __forceinline__ bool df_mainloop(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, int m0, int b_row0, int num_kb,
                                            uint8_t* smem_raw, DfCommon& c) {
    using Cfg = DfCfg<N>;
    c.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(c.smem + DF_STAGES * Cfg::STAGE_BYTES);
    c.full_bar = bars; c.empty_bar = bars + DF_STAGES; c.tmem_full = bars + 2 * DF_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * DF_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < DF_STAGES; ++i) { mbar_init(&c.full_bar[i], 1); mbar_init(&c.empty_bar[i], 1); }
        mbar_init(c.tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem_base = *tmem_ptr;
    pdl_wait();               // the producers of our operands (previous kernels of the chain) have completed
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&c.empty_bar[stage], phase ^ 1);
                uint8_t* sa = c.smem + stage * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(&c.full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_2d(&tmap_a, &c.full_bar[stage], sa, kb * DF_KE, m0);
#pragma unroll
                for (int j = 0; j < N / 128; ++j)
                    tma_load_2d(&tmap_b, &c.full_bar[stage], sa + Cfg::A_BYTES + j * (128 * 128), kb * DF_KE, b_row0 + j * 128);
                if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return false;
    }
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(DF_BM, 128, 2u);       // TF32 operands, 128 x 128 per instruction
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&c.full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(c.smem + stage * Cfg::STAGE_BYTES);
                const uint64_t da = make_sw128_kmajor_desc(sa);
#pragma unroll
                for (int j = 0; j < N / 128; ++j) {
                    const uint64_t db = make_sw128_kmajor_desc(sa + Cfg::A_BYTES + j * (128 * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(c.tmem_base + j * 128, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                }
                umma_commit(&c.empty_bar[stage]);
                if (kb == num_kb - 1) umma_commit(c.tmem_full);
            }
            __syncwarp();
            if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
        }
        return false;
    }
    return true;              // epilogue warps: the caller prefetches what it needs, then df_wait_accumulator()
}

__device__ __forceinline__ void df_wait_accumulator(const DfCommon& c) {
    mbar_wait(c.tmem_full, 0);
    tc_fence_after();
}

template <int N>
__device__ __forceinline__ void df_finish(const DfCommon& c) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) tmem_dealloc(c.tmem_base, DfCfg<N>::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// x <- LayerNorm(resid + A W^T + bias) * gamma + beta over rows of 384 (eps 1e-5, two-pass variance like nn.LayerNorm);
// writes the exact fp32 row (residual stream) and its TF32-rounded copy (A operand of the next GEMM).
// A = [L][K] fp32 (already TF32-rounded by its producer), W = [384][K] fp32 (TF32-rounded on the host).
// grid = (3, m tiles), cluster (3,1,1): CTA `slab` computes output columns [slab*128, +128) of a 128-row tile.
// ------------------------------------------------------------------------------------------
static constexpr int LN_SLABS = D_MODEL / 128;     // 3
static constexpr int LN_RES_PITCH = 528;           // bytes per row of the shared-memory residual slab (512 + 16)
static constexpr int LN_SMEM_BYTES = DfCfg<128>::SMEM_BYTES + DF_BM * LN_RES_PITCH;

__global__ void __cluster_dims__(LN_SLABS, 1, 1) __launch_bounds__(DF_THREADS, 1)
dec_gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int L, int num_kb,
                   const float* __restrict__ bias, const float* resid, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* out_x, float* out_xt) {
    extern __shared__ uint8_t df_smem[];
    __shared__ float s_stat[2][DF_BM];             // per row of the tile: partial sum / partial sum of squared deviations of this slab
    cg::cluster_group cluster = cg::this_cluster();
    DfCommon c;
    const int slab = (int)cluster.block_rank();
    const int m0 = blockIdx.y * DF_BM;
    const int col0 = slab * 128;
    const bool epi = df_mainloop<128>(tmap_a, tmap_b, m0, col0, num_kb, df_smem, c);
    const int quad = (threadIdx.x >> 5) & 3, lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;                // row inside the tile (epilogue threads)
    const long row = (long)m0 + r;
    const bool valid = row < L;
    const uint32_t t_row = c.tmem_base + (uint32_t(quad * 32) << 16);
    // residual slab [128 rows][128 floats] of this CTA in shared memory, row pitch 528 B (row-per-thread float4 reads are then
    // bank-conflict free).  The epilogue warps are idle while the MMAs run: they stream it in with coalesced cp.async.
    const uint32_t s_res = smem_u32(c.smem + DF_STAGES * DfCfg<128>::STAGE_BYTES + 256);
    uint32_t v[32];
    if (epi) {
        const int te = threadIdx.x - 64;            // 0..127
        for (int idx = te; idx < DF_BM * 32; idx += 128) {
            const int rr = idx >> 5, c4 = idx & 31;
            if ((long)m0 + rr < L) cp_async_16(s_res + rr * LN_RES_PITCH + c4 * 16, resid + ((long)m0 + rr) * D_MODEL + col0 + c4 * 4);
        }
        cp_async_commit();
        df_wait_accumulator(c);
        cp_async_wait<0>();
        named_bar_sync(1, 128);                     // every epilogue thread's part of the slab has landed
        // pass 1: v = acc + bias + residual, written back to TMEM; partial row sum of this slab
        float sum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(t_row + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + ch * 32 + j));
                const float4 rr = valid ? ld_shared_v4(s_res + r * LN_RES_PITCH + (ch * 32 + j) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float f0 = __uint_as_float(v[j]) + b.x + rr.x, f1 = __uint_as_float(v[j + 1]) + b.y + rr.y;
                const float f2 = __uint_as_float(v[j + 2]) + b.z + rr.z, f3 = __uint_as_float(v[j + 3]) + b.w + rr.w;
                sum += (f0 + f1) + (f2 + f3);
                v[j] = __float_as_uint(f0); v[j + 1] = __float_as_uint(f1); v[j + 2] = __float_as_uint(f2); v[j + 3] = __float_as_uint(f3);
            }
            tmem_st32(t_row + ch * 32, v);
        }
        tmem_st_wait();
        s_stat[0][r] = sum;
    }
    cluster.sync();                                 // every slab's partial sums are visible
    float mean = 0.f;
    if (epi) {
#pragma unroll
        for (int k = 0; k < LN_SLABS; ++k) mean += cluster.map_shared_rank(&s_stat[0][0], k)[r];
        mean *= (1.f / D_MODEL);
        float q = 0.f;                              // pass 2: squared deviations of this slab around the row mean
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(t_row + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(v[j]) - mean; q = fmaf(d, d, q); }
        }
        s_stat[1][r] = q;
    }
    cluster.sync();
    if (epi) {      // pass 3: normalise, scale, shift; exact row + TF32-rounded copy
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < LN_SLABS; ++k) q += cluster.map_shared_rank(&s_stat[1][0], k)[r];
        const float rstd = rsqrtf(q * (1.f / D_MODEL) + 1e-5f);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(t_row + ch * 32, v);
            tmem_ld_wait();
            if (valid) {
                float4* ox = reinterpret_cast<float4*>(out_x + row * D_MODEL + col0 + ch * 32);
                float4* ot = reinterpret_cast<float4*>(out_xt + row * D_MODEL + col0 + ch * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col0 + ch * 32 + j));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col0 + ch * 32 + j));
                    float4 y;
                    y.x = (__uint_as_float(v[j]) - mean) * rstd * g.x + b.x;
                    y.y = (__uint_as_float(v[j + 1]) - mean) * rstd * g.y + b.y;
                    y.z = (__uint_as_float(v[j + 2]) - mean) * rstd * g.z + b.z;
                    y.w = (__uint_as_float(v[j + 3]) - mean) * rstd * g.w + b.w;
                    ox[j >> 2] = y;
                    ot[j >> 2] = make_float4(rna_tf32(y.x), rna_tf32(y.y), rna_tf32(y.z), rna_tf32(y.w));
                }
            }
        }
    }
    cluster.sync();                                 // nobody exits while a peer may still read its s_stat
    df_finish<128>(c);
}

int launch_dec_gemm_ln(const float* a, int L, int K, const float* w, const float* bias, const float* resid,
                       const float* gamma, const float* beta, float* out_x, float* out_xt, cudaStream_t stream) {
    KOCR_CHECK(K % DF_KE == 0 && L > 0, "dec_gemm_ln: bad shape L=%d K=%d", L, K);
    CUtensorMap ta, tb;
    KOCR_TRY(make_tmap_2d(&ta, a, (uint64_t)L, (uint64_t)K, DF_BM, 4));
    KOCR_TRY(make_tmap_2d(&tb, w, (uint64_t)D_MODEL, (uint64_t)K, 128, 4));
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, dec_gemm_ln_kernel, LN_SMEM_BYTES));
    KOCR_CUDA(launch_kernel(dec_gemm_ln_kernel, dim3(LN_SLABS, (L + DF_BM - 1) / DF_BM), dim3(DF_THREADS), LN_SMEM_BYTES,
                            stream, ta, tb, L, K / DF_KE, bias, resid, gamma, beta, out_x, out_xt));
    gemm_tc_count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// logits = x W_out^T + b (384 -> 124, padded to 128), argmax (ties -> lowest index, like torch.argmax) and the greedy
// bookkeeping of predictor.py:90-97 (stop BEFORE appending <eos>).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DF_THREADS, 2)
dec_out_argmax_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int L, int num_kb,
                      const float* __restrict__ bias, int* tokens, int* lengths, int* finished, int* n_active,
                      const int* __restrict__ step_base, int step_off, const int* __restrict__ forced, float* trace) {
    extern __shared__ uint8_t df_smem[];
    DfCommon c;
    const int m0 = blockIdx.x * DF_BM;
    if (df_mainloop<VOCAB_PAD>(tmap_a, tmap_b, m0, 0, num_kb, df_smem, c)) {
        df_wait_accumulator(c);
        const int quad = (threadIdx.x >> 5) & 3, lane = threadIdx.x & 31;
        const int l = m0 + quad * 32 + lane;
        const int t = __ldcg(step_base) + step_off;
        const bool live = l < L && __ldcg(finished + l) == 0;
        const uint32_t t_row = c.tmem_base + (uint32_t(quad * 32) << 16);
        float best = -INFINITY;
        int bi = -1;
        uint32_t v[32];
#pragma unroll 1
        for (int ch = 0; ch < VOCAB_PAD / 32; ++ch) {
            tmem_ld32(t_row + ch * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + ch * 32 + j));
                f[j] = __uint_as_float(v[j]) + b.x; f[j + 1] = __uint_as_float(v[j + 1]) + b.y;
                f[j + 2] = __uint_as_float(v[j + 2]) + b.z; f[j + 3] = __uint_as_float(v[j + 3]) + b.w;
            }
            if (live && trace) {
                float4* tr = reinterpret_cast<float4*>(trace + ((long)l * DEC_MAX + t) * VOCAB_PAD + ch * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) tr[j >> 2] = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (ch * 32 + j < VOCAB && f[j] > best) { best = f[j]; bi = ch * 32 + j; }      // ascending scan, strict >: lowest index wins ties
        }
        if (live) {
            if (forced) {
                tokens[l * TOK_LD + t + 1] = forced[l * TOK_LD + t + 1];
                lengths[l] = t + 2;
            } else if (bi == 3 || bi < 0) {         // <eos> (or no finite logit at all): finished, nothing appended
                finished[l] = 1;
            } else {
                tokens[l * TOK_LD + t + 1] = bi;
                lengths[l] = t + 2;
                if (t + 1 >= DEC_MAX) finished[l] = 1; else atomicAdd(n_active + t, 1);
            }
        }
    }
    df_finish<VOCAB_PAD>(c);
}

int launch_dec_out_argmax(const float* a, int L, const float* w /*[128][384]*/, const float* bias, int* tokens, int* lengths,
                          int* finished, int* n_active, const int* step_base, int step_off, const int* forced, float* trace,
                          cudaStream_t stream) {
    KOCR_CHECK(L > 0, "dec_out_argmax: empty batch");
    CUtensorMap ta, tb;
    KOCR_TRY(make_tmap_2d(&ta, a, (uint64_t)L, (uint64_t)D_MODEL, DF_BM, 4));
    KOCR_TRY(make_tmap_2d(&tb, w, (uint64_t)VOCAB_PAD, (uint64_t)D_MODEL, 128, 4));
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, dec_out_argmax_kernel, DfCfg<VOCAB_PAD>::SMEM_BYTES));
    KOCR_CUDA(launch_kernel(dec_out_argmax_kernel, dim3((L + DF_BM - 1) / DF_BM), dim3(DF_THREADS), DfCfg<VOCAB_PAD>::SMEM_BYTES,
                            stream, ta, tb, L, D_MODEL / DF_KE, bias, tokens, lengths, finished, n_active, step_base, step_off,
                            forced, trace));
    gemm_tc_count_launch();
    return 0;
}

}  // namespace kocr
