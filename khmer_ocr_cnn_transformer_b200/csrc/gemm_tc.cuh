// tcgen05 / TMEM / TMA GEMM used for every dense contraction on the path:
//   * conv2..conv7 (and the ResNet blocks) as implicit GEMMs whose A tiles are gathered by TMA in IM2COL mode straight
//     from the dense NWHC activation (hardware zero halo, no padded rows: M = chunks * W * H exactly),
//   * patch projection, encoder/decoder projections and FFNs, BiLSTM input projection,
//     cross-attention K/V precompute, output projection (plain 2-D tiled TMA).
// D[M,N] = sum_taps A[pixel shifted by tap, cin-block] * W[N, tap*cin + cin-block]^T  (+ epilogue)
//
// Reference ops replaced: nn.Conv2d / nn.Linear / in_proj / out_proj GEMMs dispatched to
// cuDNN/cuBLAS by the reference (se_model.py:39-61,92-97,121-125,167-173,228-234).
#pragma once
#include "common.cuh"

namespace kocr {

struct GemmEpilogue {
    const float* bias;        // [N] or nullptr
    int relu;                 // activation: 0 none, 1 max(x, 0), 2 sigmoid
    int round_tf32;           // 1: out_f32 is the A operand of a TF32 GEMM: round it to TF32 (nearest) here, see rna_tf32()
    // optional fp32 addend: out += addend[(period ? row % period : row) * ld_add + n]
    const float* addend;
    int ld_add;
    int add_period;
    // outputs (either may be null)
    act16_t* out_a16;
    int ld_a16;
    float* out_f32;
    int ld_f32;
    // optional second a16 copy holding the rounding residual (x - a16(x)) for split-precision
    // ("a16x3") consumers; written at out_a16_lo with the same leading dimension.
    act16_t* out_a16_lo;
    // Column-fused epilogue of the conv GEMMs (conv only, whole-column M tiles, N tile 256, 16-bit operands):
    //   col_mode 1: (2,1) max-pool over the row pairs of every column  -> out_pool [col][H/2][N]   (se_model.py:69-73 pool3/pool4)
    //   col_mode 2: H == 3: the two overlapping row bins of AdaptiveAvgPool2d((2, .)), as SUMS y0+y1, y1+y2
    //                                                                   -> out_pool [col][2][N]     (se_model.py:61,78)
    //   col_mode 3: H == 24: 2x2 max-pool over column PAIRS (tile_cols even)        -> out_pool [col/2][12][N]  (se_model.py:65 pool2)
    //   modes 1 / 2 also write the column means over H (the SequenceSE squeeze, se_model.py:20-22) -> out_colmean [col][N],
    //   16-bit, computed from the fp32 accumulators.  out_a16 / out_f32 are not written in these modes.
    int col_mode;
    act16_t* out_pool;
    act16_t* out_colmean;     // may be null (baselines without SE)
};

struct GemmProblem {
    int M, N;                 // output rows / cols (N multiple of the N tile)
    int taps;                 // 1 (linear) or 9 (3x3 conv, pad 1)
    int cin;                  // K per tap in elements (multiple of 64 for a16, of 32 for tf32)
    int tf32;                 // 0: a16 operands (kind::f16), 1: fp32 operands consumed as TF32 (kind::tf32)
    int split_k;              // <= 1: whole K per tile.  s > 1 (linear, fp32 output only): K is cut into s slices, slice i
                              // writes its raw partial sums to out_f32 + i*M*ld_f32 (no bias/addend/activation: the
                              // consumer adds the slices).  Used by the decode loop to spread tiny GEMMs over many SMs.
    int bn;                   // N tile override (64, 128, 256); 0 = 256 if N % 256 == 0 else 128
    // taps == 9: A is the dense NWHC activation [n_img][conv_W][conv_H][cin] (pixel index m = (n*W + w)*H + h, h fastest);
    // M must equal n_img * conv_W * conv_H.  tile_cols > 0: an M tile holds tile_cols WHOLE columns (tile_cols * conv_H
    // <= 128 rows; the rest of the 128-row MMA is idle) - required by the column-fused epilogue; 0: 128 consecutive pixels.
    int conv_H, conv_W, n_img, tile_cols;
    GemmEpilogue ep;
};

// Host launchers (defined in gemm_tc.cu).  a: [rowsA, cin] row-major, w: [N, taps*cin] (a16 unless p.tf32).
// With p.tf32 the operands are fp32 arrays (a: [rowsA, cin], w: [N, taps*cin]) read by the tensor core as TF32.
int launch_gemm_tc(const void* a, long rowsA, const void* w,
                   const GemmProblem& p, int num_sms, cudaStream_t stream);
// CUDA-core restatement of the same contract (incl. the column-fused outputs); used ONLY by tests to localise tcgen05 bugs.
int launch_gemm_simt_check(const act16_t* a, long rowsA, const act16_t* w,
                           const GemmProblem& p, cudaStream_t stream);
long gemm_tc_launch_count();
void gemm_tc_count_launch();          // other tcgen05 kernels (dec_fused.cu) add their launches to the same counter
// Cached 2-D tiled tensor map of a row-major [rows, cols] matrix (esize 2: a16, 4: fp32): box = {128 bytes of K, box_rows},
// 128-byte swizzle, zero fill out of bounds.
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int esize);
void set_gemm_bn192(int on);   // 1: N = 1152 / 384 16-bit GEMMs use 128 x 192 tiles (measured slower); 0 (default): 128 x 128

}  // namespace kocr
