// tcgen05 / TMEM / TMA GEMM used for every dense contraction on the path:
//   * conv2..conv7 as implicit GEMM over the padded-linear NHWC layout (9 row-shifted taps),
//   * patch projection, encoder/decoder projections and FFNs, BiLSTM input projection,
//     cross-attention K/V precompute, output projection.
// D[M,N] = sum_taps A[rows + tap_off, cin-block] * W[N, tap*cin + cin-block]^T  (+ epilogue)
//
// Reference ops replaced: nn.Conv2d / nn.Linear / in_proj / out_proj GEMMs dispatched to
// cuDNN/cuBLAS by the reference (se_model.py:39-61,92-97,121-125,167-173,228-234).
#pragma once
#include "common.cuh"

namespace kocr {

struct GemmEpilogue {
    const float* bias;        // [N] or nullptr
    int relu;                 // activation: 0 none, 1 max(x, 0), 2 sigmoid
    // row-validity mask for padded-linear outputs: rows whose (h, w) is a pad position get 0.
    int pl_S, pl_P, pl_H, pl_W;   // pl_S == 0 -> no mask
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
};

struct GemmProblem {
    int M, N;                 // output rows / cols (N multiple of the N tile)
    int taps;                 // 1 (linear) or 9 (3x3 conv)
    int cin;                  // K per tap in elements (multiple of 64 for a16, of 32 for tf32)
    int tf32;                 // 0: a16 operands (kind::f16), 1: fp32 operands consumed as TF32 (kind::tf32)
    int split_k;              // <= 1: whole K per tile.  s > 1 (linear, fp32 output only): K is cut into s slices, slice i
                              // writes its raw partial sums to out_f32 + i*M*ld_f32 (no bias/addend/activation: the
                              // consumer adds the slices).  Used by the decode loop to spread tiny GEMMs over many SMs.
    int bn;                   // N tile override (64, 128, 256); 0 = 256 if N % 256 == 0 else 128
    int tap_off[9];           // row shift per tap
    GemmEpilogue ep;
};

// Host launchers (defined in gemm_tc.cu).  a: [rowsA, cin] row-major, w: [N, taps*cin] (a16 unless p.tf32).
// With p.tf32 the operands are fp32 arrays (a: [rowsA, cin], w: [N, taps*cin]) read by the tensor core as TF32.
int launch_gemm_tc(const void* a, long rowsA, const void* w,
                   const GemmProblem& p, int num_sms, cudaStream_t stream);
// CUDA-core restatement of the same contract; used ONLY by tests to localise tcgen05 bugs.
int launch_gemm_simt_check(const act16_t* a, long rowsA, const act16_t* w,
                           const GemmProblem& p, cudaStream_t stream);
long gemm_tc_launch_count();
void set_gemm_bn192(int on);   // 1: N = 1152 / 384 16-bit GEMMs use 128 x 192 tiles (measured slower); 0 (default): 128 x 128

}  // namespace kocr
