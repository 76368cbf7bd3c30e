// Persistent, warp-specialised tcgen05 GEMM (see gemm_tc.cuh for the contract).
//
// CTA: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), warps 2.. = epilogue
// (TMEM lane quadrant = warp_idx % 4).
// Tile = 128 (M) x BN (N), K step 64 a16 = one 128-byte swizzle atom per row.
// Accumulators: 2 stages x BN fp32 columns in TMEM so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Operands: NSTAGE-deep ring of {A 16 KB, B BN*128 B} in shared memory.
// A operand: plain 2-D tiled TMA (linear layers) or IM2COL-mode TMA (3x3 convs: `tile_rows` consecutive output pixels of
// the dense NWHC activation per load, halo zero-filled by the hardware - no padded rows are computed or stored).
// Epilogues: the standard one (bias / activation / fp32 addend / fp32 + 16-bit outputs through swizzled per-warp staging)
// and, for convs whose M tile holds whole image columns, the COLUMN-FUSED one (COLF): the tile's post-activation values go
// through a shared fp32 staging block and come out as (2,1)-max-pooled rows + SequenceSE column means (or the adaptive-
// average-pool row bins of conv7), so the un-pooled conv output never reaches HBM.
#include "gemm_tc.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

namespace kocr {

static constexpr int BM = 128;
static constexpr int BK = 64;
// Two instantiation families:
//  * 16-bit operands (stages 2-5a, throughput): 8 epilogue warps (two per TMEM lane quadrant, each takes half of the
//    tile's columns), operand ring as deep as shared memory allows, one CTA per SM;
//  * TF32 operands (decode loop, latency chains of tiny GEMMs from many streams): 4 epilogue warps, 3-stage ring,
//    ~113 KB and 192 threads per CTA so that two of them - or one plus the attention CTAs of other streams - fit on an SM.
template <int BN, bool TF32> struct GemmCfg {
    static constexpr int EPI_WARPS = TF32 ? 4 : 8;
    static constexpr int THREADS = 64 + EPI_WARPS * 32;
    static constexpr int MIN_CTAS = (TF32 && BN <= 128) ? 2 : 1;
    static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
    static constexpr int B_BYTES = BN * BK * 2;                 // 8 / 16 / 32 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NSTAGE = TF32 ? 3 : ((BN >= 192) ? 4 : 5);
    // per-epilogue-warp staging: one 4 KB tile; a second one where the fp32 addend is prefetched (16-bit, BN <= 128 only)
    static constexpr int STG_PER_WARP = (TF32 || BN == 256) ? 4096 : 8192;
    static constexpr int TMEM_COLS = (BN == 192) ? 512 : 2 * BN;   // 128 / 256 / 512 (tcgen05.alloc takes powers of two)
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * STG_PER_WARP;
};

struct GemmKernelParams {
    int M, N, taps, cin_blocks;      // cin_blocks = cin / 64
    int num_m_tiles, num_n_tiles;
    int split_k, kb_per_split;       // tile index = (m_tile * num_n_tiles + n_tile) * split_k + slice
    int tile_rows;                   // rows of M per tile: 128, or tile_cols * conv_H for whole-column conv tiles
    int conv_H, conv_W, tile_cols;   // taps == 9 only
    long total_cols;                 // n_img * conv_W
    GemmEpilogue ep;
};

// Column phase of the fused epilogue: the 4 warps of a half re-read the [tile rows][32 channels] fp32 staging block with
// lane = channel, warp = every 4th column, and emit the pooled rows and the column mean of each whole column.
template <int H, int MODE>
__device__ __forceinline__ void column_phase(uint32_t hs_base, int wq, int lane, int kc, long col0, long total_cols,
                                             act16_t* __restrict__ out_pool, act16_t* __restrict__ out_colmean, int N, int nc) {
    constexpr int HO = MODE == 1 ? H / 2 : 2;
    const uint32_t lane_off = (uint32_t)(lane & 3) << 2;
    const int piece = lane >> 2;
    for (int j = wq; j < kc; j += 4) {
        const long gcol = col0 + j;
        if (gcol >= total_cols) break;
        float v[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int r = j * H + h;
            v[h] = ld_shared_f32(hs_base + r * 128 + (((piece ^ (r & 7)) << 4) | lane_off));
        }
        if (out_colmean) {
            float sum = 0.f;
#pragma unroll
            for (int h = 0; h < H; ++h) sum += v[h];
            out_colmean[gcol * N + nc + lane] = to_a16(sum * (1.f / (float)H));
        }
        act16_t* o = out_pool + (gcol * HO) * (long)N + nc + lane;
#pragma unroll
        for (int i = 0; i < HO; ++i)
            o[(long)i * N] = to_a16(MODE == 1 ? fmaxf(v[2 * i], v[2 * i + 1]) : v[i] + v[i + 1]);
    }
}

// BK is 128 bytes of K per row in both precisions: 64 a16 or 32 fp32 (TF32) elements.
// 2x2 max-pool variant (conv2, H = 24): the tile holds an even number of columns; pooled pixel (column pair j2, row pair i)
// is the max over rows 2i, 2i+1 of columns 2*j2, 2*j2+1.  Output [pooled column][H/2][N].
template <int H>
__device__ __forceinline__ void column_phase_2x2(uint32_t hs_base, int wq, int lane, int kc, long col0, long total_cols,
                                                 act16_t* __restrict__ out_pool, int N, int nc) {
    constexpr int HO = H / 2;
    const uint32_t lane_off = (uint32_t)(lane & 3) << 2;
    const int piece = lane >> 2;
    auto ld = [&](int r) { return ld_shared_f32(hs_base + r * 128 + (((piece ^ (r & 7)) << 4) | lane_off)); };
    for (int pp = wq; pp < (kc / 2) * HO; pp += 4) {
        const int j2 = pp / HO, i = pp - j2 * HO;
        const long gcol = col0 + 2 * j2;
        if (gcol + 1 >= total_cols) continue;
        const int r0 = (2 * j2) * H + 2 * i, r1 = r0 + H;
        const float v = fmaxf(fmaxf(ld(r0), ld(r0 + 1)), fmaxf(ld(r1), ld(r1 + 1)));
        out_pool[((gcol >> 1) * HO + i) * (long)N + nc + lane] = to_a16(v);
    }
}

template <int BN, bool TF32, bool COLF>
__global__ void __launch_bounds__((GemmCfg<BN, TF32>::THREADS), (GemmCfg<BN, TF32>::MIN_CTAS))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmKernelParams p) {
    using Cfg = GemmCfg<BN, TF32>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                          // [NSTAGE]
    uint64_t* empty_bar = bars + Cfg::NSTAGE;           // [NSTAGE]
    uint64_t* tmem_full = bars + 2 * Cfg::NSTAGE;       // [2]
    uint64_t* tmem_empty = bars + 2 * Cfg::NSTAGE + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NSTAGE + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();            // a dependent (PDL-launched) kernel may begin its own prologue now

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < Cfg::NSTAGE; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], Cfg::EPI_WARPS * 32); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();               // PDL: the producer of our operands (previous kernel in the stream) has completed

    constexpr int BKE = TF32 ? 32 : 64;           // K elements per 128-byte row
    const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
    const int total_kb = p.taps * p.cin_blocks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const bool conv = p.taps == 9;
            const uint32_t stage_tx = (uint32_t)p.tile_rows * 128u + (uint32_t)Cfg::B_BYTES;   // im2col signals tile_rows * 128 B
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mn = tile / p.split_k, sp = tile - mn * p.split_k;
                const int m0 = (mn / p.num_n_tiles) * p.tile_rows;
                const int n0 = (mn % p.num_n_tiles) * BN;
                const int kb0 = sp * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
                int cn = 0, cw = 0, chh = 0;                   // first output pixel of the tile: (image, column, row)
                if (conv) {
                    const int per_img = p.conv_W * p.conv_H;
                    cn = m0 / per_img;
                    const int rem = m0 - cn * per_img;
                    cw = rem / p.conv_H;
                    chh = rem - cw * p.conv_H;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tap = kb / p.cin_blocks;
                    const int cb = kb - tap * p.cin_blocks;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (conv) {
                        // tap = r * 3 + s: r shifts along H (the map's first spatial dim), s along W (weights.py packs
                        // the 3x3 kernel as [Cout][kh][kw][Cin])
                        const int r = tap / 3, sft = tap - r * 3;
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        tma_load_im2col_4d(&tmap_a, &full_bar[stage], sa, cb * BKE, chh - 1, cw - 1, cn, (uint16_t)r, (uint16_t)sft);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(&tmap_a, &full_bar[stage], sa, cb * BKE, m0);
                    }
                    tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * BKE, n0);
                    if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(BM, BN, TF32 ? 2u : IDESC_FMT_A16);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            const int sp = tile % p.split_k;
            const int kb0 = sp * p.kb_per_split, num_kb = min(total_kb, kb0 + p.kb_per_split) - kb0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // one MMA consumes 32 B of K (16 a16 / 8 tf32) inside the 128 B swizzle row: +2 (16-byte units)
                        if (TF32) umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma_a16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);                 // frees the smem slot
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);   // accumulator ready
                }
                __syncwarp();
                if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        // Thread = one accumulator row (TMEM lane).  Global traffic goes through a per-warp 32 x 128-byte staging tile
        // in shared memory so that every global access instruction covers whole 64/128-byte row segments (4-8 rows
        // per instruction) instead of 32 scattered 16-byte pieces: the short-K GEMMs are paced by their epilogue.
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
        const int row_in_tile = quad * 32 + lane;
        const GemmEpilogue& ep = p.ep;
        // two 4 KB buffers per warp: the addend block of chunk c+1 streams in (cp.async) while chunk c is processed
        const int half = (warp - 2) >> 2;              // which half of the tile's 32-column chunks this warp handles
        const uint32_t stg_base = smem_u32(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES + 256 + (warp - 2) * Cfg::STG_PER_WARP);
        // staging addressing (16-byte pieces, XOR-swizzled so that both the row-wise and the piece-wise access
        // patterns are bank-conflict free): fp32 rows of 8 pieces, a16 rows of 4 pieces
        const uint32_t own32 = lane * 128, own16 = lane * 64;     // offsets inside a staging buffer
        const int sw32 = lane & 7, sw16 = (lane >> 1) & 3;
        const int r32 = lane >> 3, p32 = lane & 7;     // cooperative fp32 access: 4 rows x 8 pieces per instruction
        const int r16 = lane >> 2, p16 = lane & 3;     // cooperative a16 access: 8 rows x 4 pieces per instruction
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int mn = tile / p.split_k, sp = tile - mn * p.split_k;
            const int m_tile = mn / p.num_n_tiles;
            const int m0 = m_tile * p.tile_rows;
            const int n0 = (mn % p.num_n_tiles) * BN;
            const long row0 = (long)m0 + quad * 32;    // first row of this warp
            const long row = row0 + lane;
            // rows of this warp that exist: inside the tile's `tile_rows` and inside the matrix (may be <= 0)
            const int rows_here = (int)min((long)min(32, p.tile_rows - quad * 32), (long)p.M - row0);
            // addend rows this lane fetches for the warp (4 rows x 8 pieces per instruction); element offsets fit 32 bits
            uint32_t add_off[8];
            if (ep.addend) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t gr = (uint32_t)row0 + i * 4 + r32;
                    const uint32_t ar = ep.add_period > 0 ? gr % (uint32_t)ep.add_period : gr;
                    add_off[i] = ar * (uint32_t)ep.ld_add + n0 + p32 * 4;
                }
            }
            auto fetch_addend = [&](int c) {           // async copy of the 32 x 32 fp32 addend block of chunk c (L2 path)
                const uint32_t dstb = stg_base + (c & 1) * (Cfg::STG_PER_WARP - 4096);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + r32;
                    if (rr < rows_here) cp_async_16(dstb + rr * 128 + ((p32 ^ (rr & 7)) << 4), ep.addend + add_off[i] + c * 32);
                }
                cp_async_commit();
            };
            constexpr int CPW = BN / 32 / (Cfg::EPI_WARPS / 4);      // chunks per epilogue warp
            const int c0 = half * CPW;
            if (ep.addend) fetch_addend(c0);           // overlaps the wait for the accumulator
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;

            auto process = [&](const uint32_t (&v)[32], int c) {
                const int nc = n0 + c * 32;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + j));
                        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                    }
                }
                const uint32_t stg_u32 = stg_base + (c & 1) * (Cfg::STG_PER_WARP - 4096);
                if (ep.addend) {
                    // the block of this chunk was requested one chunk ago; request the next one before waiting
                    if (c + 1 < c0 + CPW) { fetch_addend(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 a = ld_shared_v4(stg_u32 + own32 + ((j ^ sw32) << 4));
                        f[4 * j] += a.x; f[4 * j + 1] += a.y; f[4 * j + 2] += a.z; f[4 * j + 3] += a.w;
                    }
                    __syncwarp();
                }
                // activation: the (warp-uniform) mode is tested OUTSIDE the unrolled loops so that the sigmoid's
                // MUFU work is not if-converted into every GEMM's epilogue
                if (ep.relu == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                } else if (ep.relu == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 1.f / (1.f + __expf(-f[j]));
                }
                if (COLF) {
                    // ---- column-fused epilogue: fp32 staging block [tile rows][32 channels] of this half, then the column phase
                    const uint32_t hs_base = smem_u32(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES + 256) + half * 16384;
                    named_bar_sync(1 + half, 128);                  // the previous chunk's column phase has finished reading
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(hs_base + row_in_tile * 128 + ((j ^ (row_in_tile & 7)) << 4),
                                     make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    named_bar_sync(1 + half, 128);                  // all 128 rows of the block are in shared memory
                    const long col0 = (long)m_tile * p.tile_cols;
                    if (p.conv_H == 24) column_phase_2x2<24>(hs_base, quad, lane, p.tile_cols, col0, p.total_cols, ep.out_pool, p.N, nc);
                    else if (p.conv_H == 12) column_phase<12, 1>(hs_base, quad, lane, p.tile_cols, col0, p.total_cols, ep.out_pool, ep.out_colmean, p.N, nc);
                    else if (p.conv_H == 6) column_phase<6, 1>(hs_base, quad, lane, p.tile_cols, col0, p.total_cols, ep.out_pool, ep.out_colmean, p.N, nc);
                    else column_phase<3, 2>(hs_base, quad, lane, p.tile_cols, col0, p.total_cols, ep.out_pool, ep.out_colmean, p.N, nc);
                    return;
                }
                if (ep.out_f32) {
                    if (ep.round_tf32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = rna_tf32(f[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(stg_u32 + own32 + ((j ^ sw32) << 4), make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    __syncwarp();
                    float* o = ep.out_f32 + ((long)sp * p.M + row0) * ep.ld_f32 + nc + p32 * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + r32;
                        const float4 a = ld_shared_v4(stg_u32 + rr * 128 + ((p32 ^ (rr & 7)) << 4));
                        if (rr < rows_here) *reinterpret_cast<float4*>(o + (long)rr * ep.ld_f32) = a;
                    }
                    __syncwarp();
                }
                if (ep.out_a16) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4u(stg_u32 + own16 + ((j ^ sw16) << 4),
                                      make_uint4(pack_a16(f[8 * j], f[8 * j + 1]), pack_a16(f[8 * j + 2], f[8 * j + 3]),
                                                 pack_a16(f[8 * j + 4], f[8 * j + 5]), pack_a16(f[8 * j + 6], f[8 * j + 7])));
                    __syncwarp();
                    act16_t* o = ep.out_a16 + row0 * ep.ld_a16 + nc + p16 * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + r16;
                        const uint4 a = ld_shared_v4u(stg_u32 + rr * 64 + ((p16 ^ ((rr >> 1) & 3)) << 4));
                        if (rr < rows_here) *reinterpret_cast<uint4*>(o + (long)rr * ep.ld_a16) = a;
                    }
                    __syncwarp();
                    if (ep.out_a16_lo && lane < rows_here) {
                        uint4* ol = reinterpret_cast<uint4*>(ep.out_a16_lo + row * ep.ld_a16 + nc);
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] -= from_a16(to_a16(f[j]));
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            ol[j] = make_uint4(pack_a16(f[8 * j], f[8 * j + 1]), pack_a16(f[8 * j + 2], f[8 * j + 3]),
                                               pack_a16(f[8 * j + 4], f[8 * j + 5]), pack_a16(f[8 * j + 6], f[8 * j + 7]));
                    }
                }
            };

            // TMEM reads are double-buffered in registers: the load of chunk c+1 is in flight while chunk c is
            // processed; the accumulator stage is released as soon as its last chunk is in registers.
            uint32_t va[32], vb[32];
            tmem_ld32(t_row + c0 * 32, va);
#pragma unroll 1
            for (int i = 0; i < CPW; i += 2) {
                tmem_ld_wait();
                if (i + 1 < CPW) {
                    tmem_ld32(t_row + (c0 + i + 1) * 32, vb);
                } else {
                    tc_fence_before();
                    mbar_arrive(&tmem_empty[acc]);
                }
                if (COLF || rows_here > 0) process(va, c0 + i);
                if (i + 1 < CPW) {
                    tmem_ld_wait();
                    if (i + 2 < CPW) {
                        tmem_ld32(t_row + (c0 + i + 2) * 32, va);
                    } else {
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    if (COLF || rows_here > 0) process(vb, c0 + i + 1);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// CUDA-core check kernels (tests only): fp32 accumulation, no tiling, written independently of the kernel above.
// ------------------------------------------------------------------------------------------
__device__ float check_dot(const act16_t* __restrict__ a, long rowsA, const act16_t* __restrict__ w, const GemmKernelParams& p,
                           long row, int n) {
    const int cin = p.cin_blocks * BK;
    float acc = 0.f;
    if (p.taps == 1) {
        if (row < rowsA) {
            const act16_t* ap = a + row * cin;
            const act16_t* wp = w + (long)n * cin;
            for (int c = 0; c < cin; ++c) acc = fmaf(from_a16(ap[c]), from_a16(wp[c]), acc);
        }
        return acc;
    }
    const int per_img = p.conv_W * p.conv_H;
    const int img = (int)(row / per_img), rem = (int)(row % per_img), wc = rem / p.conv_H, hr = rem % p.conv_H;
    for (int t = 0; t < 9; ++t) {
        const int hh = hr + t / 3 - 1, ww = wc + t % 3 - 1;
        if (hh < 0 || hh >= p.conv_H || ww < 0 || ww >= p.conv_W) continue;
        const act16_t* ap = a + (((long)img * p.conv_W + ww) * p.conv_H + hh) * cin;
        const act16_t* wp = w + (long)n * 9 * cin + (long)t * cin;
        for (int c = 0; c < cin; ++c) acc = fmaf(from_a16(ap[c]), from_a16(wp[c]), acc);
    }
    return acc;
}

__device__ float check_act(float acc, const GemmEpilogue& ep, long row, int n) {
    if (ep.bias) acc += ep.bias[n];
    if (ep.addend) {
        const long ar = ep.add_period > 0 ? (row % ep.add_period) : row;
        acc += ep.addend[ar * ep.ld_add + n];
    }
    if (ep.relu == 1) acc = fmaxf(acc, 0.f);
    else if (ep.relu == 2) acc = 1.f / (1.f + __expf(-acc));
    return acc;
}

__global__ void gemm_simt_check_kernel(const act16_t* __restrict__ a, long rowsA,
                                       const act16_t* __restrict__ w, GemmKernelParams p) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)p.M * p.N) return;
    const long row = idx / p.N;
    const int n = (int)(idx - row * p.N);
    const GemmEpilogue& ep = p.ep;
    const float acc = check_act(check_dot(a, rowsA, w, p, row, n), ep, row, n);
    if (ep.out_f32) ep.out_f32[row * ep.ld_f32 + n] = ep.round_tf32 ? rna_tf32(acc) : acc;
    if (ep.out_a16) {
        ep.out_a16[row * ep.ld_a16 + n] = to_a16(acc);
        if (ep.out_a16_lo)
            ep.out_a16_lo[row * ep.ld_a16 + n] =
                to_a16(acc - from_a16(to_a16(acc)));
    }
}

// column-fused outputs: one thread per (image column, output channel)
__global__ void gemm_simt_check_col_kernel(const act16_t* __restrict__ a, long rowsA,
                                           const act16_t* __restrict__ w, GemmKernelParams p) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.total_cols * p.N) return;
    const long col = idx / p.N;
    const int n = (int)(idx - col * p.N);
    const GemmEpilogue& ep = p.ep;
    const int H = p.conv_H;
    float y[24];
    float sum = 0.f;
    for (int h = 0; h < H; ++h) {
        const long row = col * H + h;
        y[h] = check_act(check_dot(a, rowsA, w, p, row, n), ep, row, n);
        sum += y[h];
    }
    if (ep.out_colmean) ep.out_colmean[col * p.N + n] = to_a16(sum / (float)H);
    if (ep.col_mode == 3) {          // 2x2 max-pool: this thread owns the pooled column col / 2 (even columns only)
        if ((col & 1) || col + 1 >= p.total_cols) return;
        for (int i = 0; i < H / 2; ++i) {
            float m = fmaxf(y[2 * i], y[2 * i + 1]);
            for (int r = 0; r < 2; ++r) {
                const long row = (col + 1) * H + 2 * i + r;
                m = fmaxf(m, check_act(check_dot(a, rowsA, w, p, row, n), ep, row, n));
            }
            ep.out_pool[((col >> 1) * (H / 2) + i) * p.N + n] = to_a16(m);
        }
    } else if (ep.col_mode == 1) {
        for (int i = 0; i < H / 2; ++i) ep.out_pool[(col * (H / 2) + i) * p.N + n] = to_a16(fmaxf(y[2 * i], y[2 * i + 1]));
    } else {
        for (int i = 0; i < 2; ++i) ep.out_pool[(col * 2 + i) * p.N + n] = to_a16(y[i] + y[i + 1]);
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

static void* driver_entry(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        return p;
    return nullptr;
}
static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] { fn = reinterpret_cast<PFN_encodeTiled>(driver_entry("cuTensorMapEncodeTiled")); });
    return fn;
}
static PFN_encodeIm2col get_encode_im2col_fn() {
    static PFN_encodeIm2col fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] { fn = reinterpret_cast<PFN_encodeIm2col>(driver_entry("cuTensorMapEncodeIm2col")); });
    return fn;
}

static constexpr CUtensorMapDataType A16_TMAP_TYPE =
    KOCR_A16_FORMAT == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;

// Tensor maps are cached per (pointer, geometry): encoding costs microseconds of host time per launch otherwise.
// key = (ptr, d0, d1, d2, d3, box / pixels, kind)   kind: 2 / 4 = element size of a tiled map, 9 = im2col
typedef std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t, int> TmapKey;
static std::map<TmapKey, CUtensorMap> g_tmap_cache;
static std::mutex g_tmap_mu;
static bool tmap_lookup(const TmapKey& key, CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it == g_tmap_cache.end()) return false;
    *out = it->second;
    return true;
}
static void tmap_store(const TmapKey& key, const CUtensorMap& m) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 4096) g_tmap_cache.clear();
    g_tmap_cache[key] = m;
}

// row-major [rows, cols] matrix of a16 (esize 2) or fp32 (esize 4), box = {128 bytes of K, box_rows},
// 128-byte swizzle, zero OOB fill.
static int make_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int esize) {
    const TmapKey key(ptr, rows, cols, 0, 0, box_rows, esize);
    if (tmap_lookup(key, out)) return 0;
    PFN_encodeTiled enc = get_encode_fn();
    KOCR_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * (uint64_t)esize};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esize), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, esize == 2 ? A16_TMAP_TYPE : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KOCR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu box_rows=%u", (int)r,
               ptr, (unsigned long long)rows, (unsigned long long)cols, box_rows);
    tmap_store(key, *out);
    return 0;
}

// Dense NWHC activation [n_img][W][H][C] (a16) for a 3x3 / pad-1 convolution: TMA dims (C, H, W, N) - the map's first
// spatial dimension is our H because rows are the fastest pixel index -, bounding box of base pixels [-1, d - 1) in both
// spatial dims (lower corner -pad, upper corner pad - (3 - 1)), 64 channels x `pixels` output pixels per load.
static int make_tmap_im2col(CUtensorMap* out, const void* ptr, int n_img, int W, int H, int C, int pixels) {
    const TmapKey key(ptr, (uint64_t)C, (uint64_t)H, (uint64_t)W, (uint64_t)n_img, (uint32_t)pixels, 9);
    if (tmap_lookup(key, out)) return 0;
    PFN_encodeIm2col enc = get_encode_im2col_fn();
    KOCR_CHECK(enc != nullptr, "cuTensorMapEncodeIm2col entry point not available");
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)n_img};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * C * 2, (cuuint64_t)W * H * C * 2};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, A16_TMAP_TYPE, 4, const_cast<void*>(ptr), gdim, gstr, lower, upper, 64, (cuuint32_t)pixels, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KOCR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) ptr=%p n=%d W=%d H=%d C=%d pixels=%d", (int)r, ptr,
               n_img, W, H, C, pixels);
    tmap_store(key, *out);
    return 0;
}

static std::atomic<long> g_gemm_launches{0};
static int g_bn192 = 0;                 // 192-wide N tile for N = 1152 / 384: measured SLOWER than 128 (qkv 0.206 vs 0.161 ms), kept as an A/B option
void set_gemm_bn192(int on) { g_bn192 = on; }
long gemm_tc_launch_count() { return g_gemm_launches.load(); }
void gemm_tc_count_launch() { ++g_gemm_launches; }
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int esize) {
    return make_tmap(out, ptr, rows, cols, box_rows, esize);
}

static int fill_params(GemmKernelParams& kp, const GemmProblem& p, int BN) {
    const int bke = p.tf32 ? 32 : 64;
    KOCR_CHECK(p.cin % bke == 0, "gemm: cin %d not a multiple of %d", p.cin, bke);
    KOCR_CHECK(p.N % BN == 0, "gemm: N %d not a multiple of the N tile %d", p.N, BN);
    KOCR_CHECK(p.taps == 1 || p.taps == 9, "gemm: taps must be 1 or 9");
    KOCR_CHECK(p.M > 0, "gemm: empty M");
    KOCR_CHECK(!((BN == 256 || p.tf32) && p.ep.addend), "gemm: the fp32 addend needs the double staging buffer of the 16-bit N tiles <= 128");
    kp.M = p.M; kp.N = p.N; kp.taps = p.taps; kp.cin_blocks = p.cin / bke;
    kp.tile_rows = BM; kp.conv_H = kp.conv_W = kp.tile_cols = 0; kp.total_cols = 0;
    if (p.taps == 9) {
        KOCR_CHECK(!p.tf32, "gemm: convolutions take 16-bit operands");
        KOCR_CHECK(p.conv_H > 0 && p.conv_W > 0 && p.n_img > 0 && (long)p.n_img * p.conv_W * p.conv_H == (long)p.M,
                   "gemm: conv geometry %d x %d x %d does not match M = %d", p.n_img, p.conv_W, p.conv_H, p.M);
        kp.conv_H = p.conv_H; kp.conv_W = p.conv_W; kp.tile_cols = p.tile_cols;
        kp.total_cols = (long)p.n_img * p.conv_W;
        if (p.tile_cols > 0) {
            kp.tile_rows = p.tile_cols * p.conv_H;
            KOCR_CHECK(kp.tile_rows <= BM, "gemm: %d columns of %d rows exceed the 128-row tile", p.tile_cols, p.conv_H);
        }
    }
    if (p.ep.col_mode) {
        KOCR_CHECK(p.taps == 9 && p.tile_cols > 0 && (BN == 256 || BN == 128) && !p.tf32 && p.ep.out_pool && !p.ep.addend,
                   "gemm: the column-fused epilogue needs a conv with whole-column tiles, a 128 / 256-wide N tile and out_pool");
        KOCR_CHECK((p.ep.col_mode == 1 && (p.conv_H == 12 || p.conv_H == 6)) || (p.ep.col_mode == 2 && p.conv_H == 3) ||
                   (p.ep.col_mode == 3 && p.conv_H == 24 && p.tile_cols % 2 == 0 && p.conv_W % 2 == 0),
                   "gemm: column-fused epilogue: unsupported (mode %d, H %d, %d columns per tile)", p.ep.col_mode, p.conv_H, p.tile_cols);
    }
    kp.num_m_tiles = (p.M + kp.tile_rows - 1) / kp.tile_rows;
    kp.num_n_tiles = p.N / BN;
    kp.split_k = p.split_k > 1 ? p.split_k : 1;
    const int total_kb = kp.taps * kp.cin_blocks;
    KOCR_CHECK(kp.split_k <= total_kb, "gemm: split_k %d exceeds the %d K blocks", kp.split_k, total_kb);
    kp.kb_per_split = (total_kb + kp.split_k - 1) / kp.split_k;
    KOCR_CHECK((kp.split_k - 1) * kp.kb_per_split < total_kb, "gemm: split_k %d leaves an empty K slice", kp.split_k);
    if (kp.split_k > 1)
        KOCR_CHECK(p.taps == 1 && p.ep.out_f32 && !p.ep.out_a16 && !p.ep.bias && !p.ep.addend && !p.ep.relu,
                   "gemm: split-K writes raw fp32 partial sums only");
    kp.ep = p.ep;
    return 0;
}

template <int BN, bool TF32, bool COLF>
static int launch_impl(const void* a, long rowsA, const void* w, const GemmProblem& p, int num_sms,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<BN, TF32>;
    GemmKernelParams kp;
    KOCR_TRY(fill_params(kp, p, BN));
    CUtensorMap ta, tb;
    const int esize = TF32 ? 4 : 2;
    if (p.taps == 9) KOCR_TRY(make_tmap_im2col(&ta, a, p.n_img, p.conv_W, p.conv_H, p.cin, kp.tile_rows));
    else KOCR_TRY(make_tmap(&ta, a, (uint64_t)rowsA, (uint64_t)p.cin, BM, esize));
    KOCR_TRY(make_tmap(&tb, w, (uint64_t)p.N, (uint64_t)p.taps * p.cin, BN, esize));
    static PerDeviceOnce attr_once;
    KOCR_CUDA(opt_in_dynamic_smem(attr_once, gemm_tc_kernel<BN, TF32, COLF>, Cfg::SMEM_BYTES));
    const int tiles = kp.num_m_tiles * kp.num_n_tiles * kp.split_k;
    const int grid = tiles < num_sms ? tiles : num_sms;
    KOCR_CUDA(launch_kernel(gemm_tc_kernel<BN, TF32, COLF>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, ta, tb, kp));
    ++g_gemm_launches;
    return 0;
}

int launch_gemm_tc(const void* a, long rowsA, const void* w, const GemmProblem& p, int num_sms,
                   cudaStream_t stream) {
    // N tile: 256 where it divides N, else 128 (option gemm_bn192: 192 for N = 1152 / 384 - fewer, fatter tiles, but slower)
    const int bn = p.bn > 0 ? p.bn : (p.N % 256 == 0 ? 256 : ((!p.tf32 && g_bn192 && p.N % 192 == 0) ? 192 : 128));
    if (p.ep.col_mode) {
        KOCR_CHECK((bn == 256 || bn == 128) && !p.tf32, "gemm: the column-fused epilogue needs a 128 / 256-wide 16-bit tile");
        if (bn == 128) return launch_impl<128, false, true>(a, rowsA, w, p, num_sms, stream);
        return launch_impl<256, false, true>(a, rowsA, w, p, num_sms, stream);
    }
    if (p.tf32) {
        if (bn == 256) return launch_impl<256, true, false>(a, rowsA, w, p, num_sms, stream);
        if (bn == 128) return launch_impl<128, true, false>(a, rowsA, w, p, num_sms, stream);
        if (bn == 64) return launch_impl<64, true, false>(a, rowsA, w, p, num_sms, stream);
    } else {
        if (bn == 256) return launch_impl<256, false, false>(a, rowsA, w, p, num_sms, stream);
        if (bn == 192) return launch_impl<192, false, false>(a, rowsA, w, p, num_sms, stream);
        if (bn == 128) return launch_impl<128, false, false>(a, rowsA, w, p, num_sms, stream);
    }
    KOCR_CHECK(false, "gemm: unsupported N tile %d", bn);
    return 2;
}

int launch_gemm_simt_check(const act16_t* a, long rowsA, const act16_t* w, const GemmProblem& p,
                           cudaStream_t stream) {
    KOCR_CHECK(!p.tf32, "gemm check kernel: a16 operands only");
    GemmKernelParams kp;
    KOCR_TRY(fill_params(kp, p, (p.ep.col_mode && p.N % 256 == 0) ? 256 : 128));
    if (p.ep.col_mode) {
        const long total = kp.total_cols * p.N;
        gemm_simt_check_col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, rowsA, w, kp);
    } else {
        const long total = (long)p.M * p.N;
        gemm_simt_check_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, rowsA, w, kp);
    }
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace kocr
