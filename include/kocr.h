/* kocr.h - C ABI of the B200-native text-line recogniser (libkocr_b200.so).
 *
 * Drop-in boundary for the recognition forward path of netra-ai-lab/Khmer-OCR-CNN-Transformer.
 * The reference has no FFI layer (pure Python on torch); these entry points are what a binding
 * for the path's stages would call.  Each one names the reference code it replaces
 * (paths relative to netra_ocr/recognition/).
 *
 * Conventions: plain pointers and sizes, no torch types.  Every function returns 0 on success and
 * a non-zero status on failure; kocr_last_error() then returns a thread-local message.  Nothing
 * throws across the boundary.  A handle owns its device weights and workspace, is bound to one
 * GPU, and must be used from one host thread at a time.  `stream` is a cudaStream_t passed as
 * void* (NULL = a non-blocking stream owned by the handle; device buffers passed in are then ordered after the work the
 * caller has enqueued on the default stream, and every call that returns data to the host synchronises before returning).  There is NO CPU fallback: without a CUDA device every
 * compute entry fails with an error.
 */
#ifndef KOCR_H_
#define KOCR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOCR_ABI_VERSION 2
#define KOCR_TOKENS_LD 257          /* 1 <sos> + up to 256 generated ids per line */

typedef struct kocr_handle kocr_handle;

/* ABI version of the loaded library (== KOCR_ABI_VERSION). */
int kocr_abi_version(void);

/* Message of the last failure on the calling thread ("" if none). */
const char* kocr_last_error(void);

/* Build a recogniser from a packed weight blob (see khmer_ocr_cnn_transformer_b200/weights.py:
 * reference state_dict -> BN-folded, K-major 16-bit (fp16) GEMM operands + fp32 vectors; a blob packed for the other
 * 16-bit format of a -DKOCR_A16_BF16 build is rejected).
 * Replaces OCRPredictor.__init__/_load_weights (predictor.py:13-46).
 * max_lines / max_chunks bound one batch; the workspace is allocated here, once. */
int kocr_create(const void* weight_blob, size_t blob_bytes, int device, int max_lines, int max_chunks,
                kocr_handle** out);
int kocr_destroy(kocr_handle* h);

/* Bytes of device memory held by the handle (weights + workspace). */
size_t kocr_workspace_bytes(const kocr_handle* h);

/* Model facts read from the blob: variant (0 = SE-VGG + BiLSTM, 1 = VGG baseline, 2 = ResNet baseline), emb_dim,
 * max_seq_len (global_pos rows), decode_max_len, vocab size.  (utils.py:14-43 autodetect_config) */
int kocr_model_info(const kocr_handle* h, int* variant, int* emb_dim, int* max_seq_len, int* decode_max_len,
                    int* vocab_size);

/* Stage 1 - ImagePreprocessor.process for a batch of grey lines (preprocessor.py:35-58) including
 * Pillow's BILINEAR resize to height 48, 48x100/overlap-16 chunking, white padding, normalisation.
 * pixels: concatenated uint8 (h_i x w_i) images, on the host (pixels_on_device = 0; copied with
 * cudaMemcpyAsync) or already on the device (1).  offsets[i] = byte offset of line i.
 * chunk_counts_out[i] (host, may be NULL) receives the chunks kept for line i
 * (= min(ceil(W'/84), ceil(max_seq_len/32))).  Leaves fp32 (n,1,48,100) chunks in the workspace. */
int kocr_gather_chunks(kocr_handle* h, const uint8_t* pixels, size_t pixel_bytes, int pixels_on_device,
                       const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_lines,
                       int32_t* chunk_counts_out, void* stream);

/* Stages 2-4 - model.cnn -> model.patch -> model.enc on all chunks of the batch
 * (predictor.py:166-170; se_model.py:63-79,104-117,119-126; vgg_model.py:50-59 / resnet_model.py:75-91 for the
 * baselines), then merge + global_pos (predictor.py:174-183).  Consumes the chunks left by kocr_gather_chunks. */
int kocr_sevgg_encoder_forward(kocr_handle* h, void* stream);

/* Stage 5a - context_bilstm over each line's merged sequence (predictor.py:185-186;
 * se_model.py:228-234) and the decoder's cross-attention K/V precompute.  For the VGG / ResNet baselines
 * (no BiLSTM) the merged sequence is the memory. */
int kocr_merge_bilstm_forward(kocr_handle* h, void* stream);

/* Stage 5b - OCRPredictor._greedy_decode for every line of the batch (predictor.py:85-99) with a
 * KV cache.  tokens_out: host int32 [n_lines, KOCR_TOKENS_LD], row = <sos> then generated ids;
 * lengths_out: host int32 [n_lines] = number of valid entries in the row (eos is not stored).
 * max_steps <= decode_max_len (0 = decode_max_len).  Synchronises the stream before returning. */
int kocr_decode_greedy(kocr_handle* h, int max_steps, int32_t* tokens_out, int32_t* lengths_out, void* stream);

/* Whole path, host in / host out: gather_chunks -> sevgg_encoder_forward -> merge_bilstm_forward ->
 * decode_greedy.  This is what OCRPredictor.predict_batch (predictor.py:138-199) calls per batch. */
int kocr_recognize_lines(kocr_handle* h, const uint8_t* pixels, size_t pixel_bytes, int pixels_on_device,
                         const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_lines,
                         int max_steps, int32_t* tokens_out, int32_t* lengths_out, void* stream);

/* Options (tests / parity tooling):
 *   "trace_logits"  1 -> decode_greedy records the last-position logits of every step
 *   "force_tokens"  1 -> decode_greedy feeds the ids set with kocr_set_forced_tokens instead of its argmax
 *   "straggler_threshold" n -> see kocr_read_unfinished;  "lstm_impl" 0/1, "use_graphs" 0/1, "big_gemm_sms" n: tuning knobs
 *   "compact_rows" 1 -> greedy loop: once a 128-row tile of lines has emitted <eos>, swap the active rows to the front and
 *                       continue on fewer rows (same tokens; a measured +0.5 %, off by default); "kv_split" 0 / "lstm_split" 0 -> plain
 *                       16-bit cross-attention K/V / BiLSTM input projection instead of the split-precision ones
 *   "blocking_wait" 1 -> host waits inside the calls sleep (blocking-sync event) instead of spinning: for processes that
 *                        keep many handles / host threads in flight (bench.py sets it when --in-flight > 1)
 *   A/B switches of kernel variants (process-wide): "chunk_attn_impl", "dec_cross_impl", "gemm_bn192", "dec_wide",
 *                        "use_pdl" - see DESIGN.md section 4
 *   "kernel_timing" 1 -> per-launch CUDA-event timing (see kocr_read_kernel_timing); setting it clears the totals */
int kocr_set_option(kocr_handle* h, const char* name, int value);
/* Beam search support - OCRPredictor._beam_search (predictor.py:101-136).  After kocr_gather_chunks /
 * kocr_sevgg_encoder_forward / kocr_merge_bilstm_forward on a batch, runs ONE decoder position for n_rows (<= 8)
 * hypotheses of line `line`: prefixes = host int32 [n_rows, t+1] (token ids of positions 0..t of each hypothesis),
 * parents[r] = row of the previous call whose cache hypothesis r continues (ignored for t = 0).
 * logits_out = host fp32 [n_rows, 128] (first 124 valid) for position t+1.  The beam bookkeeping (log-softmax,
 * top-k, stable sort, pruning, length normalisation) stays on the host, identical to the reference's Python. */
int kocr_beam_step(kocr_handle* h, int line, int n_rows, const int32_t* parents, const int32_t* prefixes, int t,
                   float* logits_out, void* stream);
/* The same for the hypotheses of MANY lines in one pass (what predict_batch(beam_width > 1) uses): row r belongs to line
 * row_line[r] of the current batch; n_rows <= max_lines of the handle; parents index the rows of the previous call. */
int kocr_beam_step_batch(kocr_handle* h, int n_rows, const int32_t* row_line, const int32_t* parents,
                         const int32_t* prefixes, int t, float* logits_out, void* stream);

/* The whole of OCRPredictor._beam_search (predictor.py:101-136) for every line of the current batch (after kocr_gather_chunks /
 * kocr_sevgg_encoder_forward / kocr_merge_bilstm_forward), in one call: KV-cached decoder positions for n_lines * beam_width
 * hypothesis rows (<= max_lines), log-softmax + top-k on the device, the reference's bookkeeping (float64 score sums, stable
 * sort, every <eos> candidate completed with score / len, first beam_width others survive) in the library.
 * tokens_out: host int32 [n_lines, KOCR_TOKENS_LD] = the winning sequence as the reference hands it to Tokenizer.decode
 * (<sos> ... and <eos> when a hypothesis completed); lengths_out: its length.  max_len <= decode_max_len (0 = decode_max_len). */
int kocr_beam_search(kocr_handle* h, int beam_width, int max_len, int32_t* tokens_out, int32_t* lengths_out, void* stream);

/* Teacher-forced batched forward - KhmerOCR.forward (recognition/model/se_model.py:240-289; vgg_model.py:214-246), the
 * training-time / evaluation-loop semantics: after kocr_gather_chunks + kocr_sevgg_encoder_forward on a batch of B
 * lines, every merged sequence is padded to Tmax = max T_i, global_pos is added to the pad rows too, the BiLSTM runs
 * over all Tmax rows of every line WITHOUT packing (the backward direction reads the pad rows first), the cross-
 * attention is masked to the real length (memory_key_padding_mask) and the self-attention is causal with <pad> keys
 * masked.  tgt_tokens = host int32 [B][L] (row = <sos> + target, right-padded with <pad> = 0), 1 <= L <= 256;
 * logits_out = host fp32 [B][L][128] (first 124 valid): row t = distribution of token t + 1 given tokens 0..t.
 * Needs B * Tmax <= 32 * max_chunks.  Leaves the handle's memory in the padded layout (kocr_decode_greedy then decodes
 * with the training-time memory); the next kocr_gather_chunks resets it. */
int kocr_forward_teacher_forced(kocr_handle* h, const int32_t* tgt_tokens, int L, float* logits_out, void* stream);

/* The reference's MODEL protocol - what its OCRPredictor calls on `self.model` (predictor.py:53-78,166-192; modules at
 * model/se_model.py:35-79 cnn, :81-117 patch, :119-126 enc, :228-234 context_bilstm, :162-208 dec) - as stage-level entry
 * points on HOST tensors in the reference's own layouts (fp32, C-contiguous).  They reuse the handle's batch state: do not
 * interleave them with a kocr_gather_chunks ... kocr_decode_greedy sequence on the same handle.
 *   cnn:    chunks (n, 1, 48, 100)        -> f (n, 512, 2, 32)
 *   patch:  f (n, 512, 2, 32)             -> x (n, 32, 384)            (+ bias + local positions; the reference also returns N = 32)
 *   enc:    p seq-first (32, n, 384)      -> (32, n, 384)
 *   bilstm: merged (B, T, 384)            -> (B, T, 384)               (zero initial state; SE-VGG family only)
 *   dec:    tgt int32 (B, t), memory (B, T, 384), pad_mask uint8 (B, T) (1 = padded, a suffix per line; may be NULL)
 *                                         -> logits (B, t, vocab = 124)  (causal self-attention, <pad> target keys masked) */
int kocr_model_cnn(kocr_handle* h, const float* chunks, int n, float* f_out, void* stream);
int kocr_model_patch(kocr_handle* h, const float* f, int n, float* x_out, void* stream);
int kocr_model_enc(kocr_handle* h, const float* p, int n, float* out, void* stream);
int kocr_model_bilstm(kocr_handle* h, const float* merged, int B, int T, float* out, void* stream);
int kocr_model_dec(kocr_handle* h, const int32_t* tgt, int B, int t, const float* memory, int T, const uint8_t* pad_mask,
                   float* logits_out, void* stream);

/* Input side - extract_textline_crops (netra_ocr/textline_detection.py:7-53) and the custom-detector crop of
 * OCREngine (netra_ocr/ocr_engine.py:72-76), followed by the `convert('L')` of ImagePreprocessor.process
 * (recognition/preprocessor.py:39-41).  page = uint8 [page_h][page_w][channels] (3 = RGB, 1 = L), host or device.
 * boxes = host int32 [n_lines][4] = (x0, y0, x1, y1), already expanded and clipped to the page by the caller (the
 * reference does that arithmetic in Python ints).  Line i is written to out_pixels_dev + out_offsets[i] (DEVICE
 * memory owned by the caller) as a grey image of (y1 - y0 + 2 pad_px) rows x (x1 - x0 + 2 pad_px) columns: the crop
 * in the middle of a white (255) canvas - ready for kocr_gather_chunks / kocr_recognize_lines with
 * pixels_on_device = 1.  Bit-exact with Pillow (crop, paste, rgb2l = (19595 R + 38470 G + 7471 B + 0x8000) >> 16). */
int kocr_crop_lines(kocr_handle* h, const uint8_t* page, int page_h, int page_w, int channels, int page_on_device,
                    const int32_t* boxes, int n_lines, int pad_px, uint8_t* out_pixels_dev, const int64_t* out_offsets,
                    void* stream);

/* Long-tail handling.  With option "straggler_threshold" = n > 0, kocr_decode_greedy / kocr_recognize_lines return as
 * soon as at most n lines are still decoding (checked every 8 positions).  flags_out[i] = 1 marks the lines whose
 * row is incomplete; the caller re-submits those lines in a later batch (greedy decoding is deterministic, so the
 * result is the same as decoding them to the end here).  Host int32 [n_lines]. */
int kocr_read_unfinished(kocr_handle* h, int32_t* flags_out);
/* With option "kernel_timing" = 1 every launch of the one-time stages (2-5a) is bracketed by CUDA events
 * on its stream.  This call synchronises and writes one line per launch site:
 * "<site> <total ms> <launches> <algorithmic FLOPs summed over those launches>\n". */
int kocr_read_kernel_timing(kocr_handle* h, char* text_out, size_t cap);
int kocr_set_forced_tokens(kocr_handle* h, const int32_t* tokens /* [n_lines, KOCR_TOKENS_LD] */, int n_lines);

/* Copy an intermediate of the last batch to the host (tests only).  Names: "chunks" f32 (n,1,48,100);
 * "pool1" "conv2" "pool2" "conv3" "pool3" "conv5" "pool4" 16-bit dense NWHC activations [n][w][h][C] (pool3 / pool4 are the
 * SE-gated, (2,1)-pooled outputs of conv4 / conv6 - the un-pooled conv4 / conv6 / conv7 outputs exist only for the ResNet
 * baseline); "bins7" 16-bit [n*25 + w][2][512] row-bin sums of conv7; "se_mean3/4/5" 16-bit [n*25 + w][C] column means;
 * "patch_in" 16-bit [n*32,1024]; "enc" f32 [n*32,384] (encoder output + global_pos); "memory" f32 [tokens,384];
 * "logits_trace" f32 [n_lines, steps, 128].  Returns the byte size through *bytes_out when dst is NULL. */
int kocr_debug_read(kocr_handle* h, const char* name, void* dst, size_t dst_bytes, size_t* bytes_out);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t kocr_launch_count(void);

/* Unit-test hook for the tcgen05 GEMM: D = A * W^T + bias (device pointers, 16-bit operands; impl 2: fp32 operands as TF32).
 * taps = 1: A = [rows_a, cin] row-major.  taps = 9: 3x3 / pad-1 convolution, A = dense NWHC activation
 * [m / (conv_h * conv_w)][conv_w][conv_h][cin], W = [n][9 * cin] with tap = kh * 3 + kw; tile_cols > 0 selects whole-column M
 * tiles; col_mode 1 / 2 selects the column-fused epilogue (out_pool: pooled rows / row-bin sums, out_colmean: column means;
 * out_f32 / out_a16 unused).  impl 0 = tcgen05 kernel, 1 = CUDA-core check kernel. */
int kocr_test_gemm(int impl, const void* a_a16, int64_t rows_a, const void* w_a16, int m, int n, int taps, int cin,
                   int conv_h, int conv_w, int tile_cols, int col_mode, const float* bias, int relu, float* out_f32,
                   void* out_a16, void* out_pool, void* out_colmean, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KOCR_H_ */
