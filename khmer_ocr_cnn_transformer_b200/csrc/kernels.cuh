// Launchers of the non-GEMM kernels (CUDA-core / HBM-bound parts of the path).
#pragma once
#include "common.cuh"

namespace kocr {

static constexpr int IMG_H = 48;          // config.py:7
static constexpr int CHUNK_W = 100;       // config.py:8
static constexpr int CHUNK_STRIDE = 84;   // chunk_width - chunk_overlap (preprocessor.py:31)
static constexpr int TOK_PER_CHUNK = 32;  // AdaptiveAvgPool2d((2,32)) -> 32 patches (se_model.py:61)
static constexpr int D_MODEL = 384;
static constexpr int N_HEAD = 8;
static constexpr int HEAD_DIM = 48;
static constexpr int LSTM_H = 192;
static constexpr int VOCAB = 124;
static constexpr int VOCAB_PAD = 128;
static constexpr int DEC_MAX = 256;       // dec.pos_emb rows (se_model.py:171)

struct LineDesc {
    long long src_off;    // byte offset of the (h, w) grey image in the pixel buffer
    long long mid_off;    // byte offset of the (h, new_w) horizontally-resized intermediate
    int h, w, new_w;
    int first_chunk, n_chunks;
    int pad_;
};

// ---- stage 1 ---------------------------------------------------------------------------
int launch_preprocess(const uint8_t* d_pixels, uint8_t* d_mid, const LineDesc* d_lines, const int* d_chunk_line,
                      int* d_vtab, float* d_chunks, int n_lines, int n_chunks, int max_new_w, cudaStream_t stream);
// text-line crops of a page (box + white padding + RGB->L), see preprocess.cu
int launch_crop_lines(const uint8_t* d_page, int page_w, int channels, const int* d_boxes, const long long* d_offsets,
                      int n_lines, int pad, uint8_t* d_out, cudaStream_t stream);
int preprocess_vtab_ints_per_line();
int preprocess_kmax();

// ---- stage 2 helpers -------------------------------------------------------------------
// Activations: 16-bit dense NWHC [chunk][w][h][C] (pixel index (n*W + w)*H + h; see cnn_misc.cu).
// conv1 (Cin=1) + folded BN + ReLU + 2x2 max-pool on mma.sync (K = 9 taps padded to 16): f32 chunks -> (50 x 24, 64);
// w16 = a16 [64][16].
int launch_conv1_pool_mma(const float* d_chunks, const act16_t* w16, const float* b, act16_t* out, int n_chunks,
                          cudaStream_t stream);
// 16-bit -> fp32 copy (n_elems multiple of 8)
int launch_a16_to_f32(const act16_t* in, float* out, long n_elems, cudaStream_t stream);
// 2x2 and (2,1) max-pools (C multiple of 8).
int launch_pool2x2(const act16_t* in, act16_t* out, int n_chunks, int H, int W, int C, cudaStream_t stream);
int launch_pool_h2(const act16_t* in, act16_t* out, int n_chunks, int H, int W, int C, cudaStream_t stream);
// AdaptiveAvgPool2d((2,32)) of a (3, W) map without a gate -> patch GEMM operand [n*32 + k][kh*C + c];
// rows_in = 2: `in` = row-bin sums [col][2][C] from the conv7 epilogue, rows_in = 3: raw rows [col][3][C].
int launch_finalpool(const act16_t* in, int rows_in, act16_t* out, int n_chunks, int W, int C, cudaStream_t stream);
// 1D-SE excitation on the 16-bit column means [n*W + w][C] written by the conv epilogue: gate = sigmoid(FC2(relu(FC1(mean))));
// pooled [col][rows][C] *= gate in place, or (final_pool) gate * row-bin sums -> AdaptiveAvgPool2d((2,32)) -> out.
struct SEWeights { const act16_t* w0p; const float* b0p; const act16_t* w2p; const float* b2;
                   const act16_t* w0f; const act16_t* w2f; };   // w0f / w2f: the same weights in mma fragment order (weights.se_fragments)
void set_se_excite_variant(int v);   // probe hook, see launch_se_excite (cnn_misc.cu)
int launch_se_excite(const act16_t* means, const SEWeights& w, act16_t* pooled, act16_t* out, int n_chunks, int rows, int W,
                     int C, bool final_pool, cudaStream_t stream);

// ---- stage 4/5 helpers -----------------------------------------------------------------
// per-chunk 32-token, 8-head attention: qkv a16 [M, 1152] -> out a16 [M, 384].
int launch_chunk_attention(const act16_t* qkv, act16_t* out, int n_chunks, cudaStream_t stream);
void set_chunk_attention_impl(int impl);   // 1 = mma.sync kernel (default), 0 = CUDA-core kernel (A/B tests)
// LayerNorm over 384: y = LN(x)*g + b (+ pos[row_pos[row]]); writes f32 and/or a16 (+ residual lo part).
// Input row = sum of nsplit split-K partials (x + s*rows*384) + in_bias + resid (optional).
int launch_layernorm(const float* x, const float* g, const float* b, const float* pos, const int* row_pos,
                     float* out_f32, act16_t* out_a16, act16_t* out_a16_lo, int rows,
                     cudaStream_t stream, int nsplit = 1, const float* in_bias = nullptr,
                     const float* resid = nullptr, float* out_tf32 = nullptr);
// f32 [rows, 384] (+ pos[row_pos[row]]) -> f32 + a16 copies (VGG merge path without LN).
int launch_add_pos(const float* x, const float* pos, const int* row_pos, float* out_f32, act16_t* out_a16,
                   act16_t* out_a16_lo, int rows, cudaStream_t stream);

// fp32 [rows, 384] -> 16-bit [rows, 1152] = [hi | lo | hi] (split-precision GEMM operand)
int launch_split3(const float* m, act16_t* out, long rows, cudaStream_t stream);
// padded memory of the teacher-forced batched forward: out[b*Tmax + t] = t < T_b ? xb[src_off_b + t] : a16(global_pos[t])
int launch_pad_memory(const act16_t* xb, const float* global_pos, const int* src_off, const int* line_T, int n_lines,
                      int Tmax, act16_t* out, cudaStream_t stream);
// BiLSTM recurrence (input projection already in gin): persistent 2-CTA cluster kernel.
struct LstmGroup { int line[8]; };   // lines handled together by one cluster (-1 = unused)
int launch_bilstm(const float* gin /*[Mtok,1536]*/, const act16_t* whh_packed, const int* line_tok_off,
                  const int* line_T, const LstmGroup* groups, int n_groups, float* mem_f32,
                  act16_t* mem_a16, act16_t* mem_a16_lo, cudaStream_t stream);
size_t bilstm_whh_packed_elems();
// tensor-core version: groups of 16 lines, recurrent weights as register-resident mma.sync fragments
struct LstmGroup16 { int line[16]; };
int launch_bilstm_mma(const float* gin, const act16_t* whh_mma, const int* line_tok_off, const int* line_T,
                      const LstmGroup16* groups, int n_groups, float* mem_f32, act16_t* mem_a16,
                      act16_t* mem_a16_lo, cudaStream_t stream);
size_t bilstm_whh_mma_elems();

// ---- decoder step kernels ---------------------------------------------------------------
// The generated position of a step is t = *step_base + step_off: step_base lives on the device so that a
// captured CUDA graph of 8 steps can be replayed for every group of 8 positions.
int launch_dec_embed(const int* tokens /*[L, DEC_MAX+1]*/, const int* step_base, int step_off, const float* tok_emb,
                     const float* pos_emb, float* x, float* x_tf32 /* TF32-rounded copy, GEMM operand */, act16_t* xb,
                     act16_t* xb_lo, int n_lines, cudaStream_t stream);
int launch_dec_self_attn(const float* qkv /*[L,1152]*/, float* kcache,
                         float* vcache /*[L, DEC_MAX, 384]*/, const int* tokens, const int* step_base,
                         int step_off, const int* finished, float* out, int n_lines, cudaStream_t stream,
                         int nsplit, const float* bias);
int launch_dec_cross_attn(const float* q /*[L,384]*/, const act16_t* kv /*[Mtok,1536]*/, int layer,
                          const int* line_tok_off, const int* line_T, int max_T, const int* finished,
                          float* out, int n_lines, cudaStream_t stream, int nsplit, const float* bias);
void set_dec_cross_attention_impl(int impl);   // 1 = one-pass online-softmax kernel (default), 0 = three-phase kernel
int launch_dec_argmax(const float* logits /*[L,128]*/, int* tokens, int* lengths, int* finished, int* n_active,
                      const int* step_base, int step_off, int n_lines, const int* forced, float* trace,
                      cudaStream_t stream, int nsplit, const float* bias);
int launch_dec_bump(int* step_base, int n, cudaStream_t stream);
// Fused decode kernels (dec_fused.cu): x <- LN(resid + a W^T + bias) as ONE single-tile TF32 GEMM whose epilogue threads own
// a row each (a: [L][K] fp32, w: [384][K]); writes the exact row (out_x, may alias resid) and its TF32-rounded copy (out_xt).
int launch_dec_gemm_ln(const float* a, int L, int K, const float* w, const float* bias, const float* resid,
                       const float* gamma, const float* beta, float* out_x, float* out_xt, cudaStream_t stream);
// Output projection + argmax + greedy bookkeeping.
int launch_dec_out_argmax(const float* a, int L, const float* w /*[128][384]*/, const float* bias, int* tokens, int* lengths,
                          int* finished, int* n_active, const int* step_base, int step_off, const int* forced, float* trace,
                          cudaStream_t stream);
// row compaction of the greedy loop: pairs = int2 (src row in the tail, dst row in the head), see seq.cu
int launch_decode_compact(const int* pairs, int n_pairs, int t, float* kcache, float* vcache, size_t layer_stride, int* tokens,
                          int* lengths, int* finished, int* tok_off, int* line_T, cudaStream_t stream);

}  // namespace kocr
