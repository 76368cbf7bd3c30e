// C-ABI of libkocr_b200.so (include/kocr.h): handle, packed-weight lookup, workspace carving and the
// orchestration of the five stages of the recognition forward path on one CUDA stream.
#include "../../include/kocr.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <chrono>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace kocr {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static thread_local bool g_pdl = false;
bool pdl_active() { return g_pdl; }
void pdl_set_active(bool on) { g_pdl = on; }

static std::atomic<int64_t> g_launches{0};     // non-GEMM kernel launches (GEMM launches are counted in gemm_tc.cu)

// ---- packed weight blob ---------------------------------------------------------------
struct BlobHeader { char magic[8]; uint32_t n_entries; uint32_t reserved; };
struct BlobEntry { char name[48]; uint32_t dtype; uint32_t pad; uint64_t offset; uint64_t nbytes; };

struct ResBlockW {                 // BasicBlock of the ResNet baseline (model/resnet_model.py:5-37), BN folded
    const act16_t *c1w, *c2w, *scw;     // 3x3 convs [Cout][9][Cin]; 1x1 shortcut conv [Cout][Cin] or null (identity)
    const float *c1b, *c2b, *scb;
};
struct EncLayerW {
    const act16_t *in_w, *out_w, *l1_w, *l2_w;
    const float *in_b, *out_b, *l1_b, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b;
};
struct DecLayerW {
    const float *sa_in_w, *sa_out_w, *ca_q_w, *ca_out_w, *l1_w, *l2_w;       // fp32, pre-rounded to TF32
    const float *sa_in_b, *sa_out_b, *ca_q_b, *ca_out_b, *l1_b, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b, *n3_g, *n3_b;
};

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace kocr

using namespace kocr;

struct kocr_handle;
extern "C" int kocr_decode_greedy(kocr_handle* h, int max_steps, int32_t* tokens_out, int32_t* lengths_out, void* stream);

struct kocr_handle {
    int device = 0;
    int num_sms = 148;
    int variant = 0, emb_dim = 384, max_seq_len = 4096, dec_max_len = 256, vocab = 124;
    int max_lines = 0, max_chunks = 0;
    // weights
    uint8_t* d_blob = nullptr;
    size_t blob_bytes = 0;
    std::map<std::string, std::pair<const void*, size_t>> w;
    const float *conv1_w, *conv1_b;
    ResBlockW res[6];
    const act16_t* conv1_w16;      // [64][16]: 9 taps + 7 zeros, tensor-core conv1
    const act16_t* conv_w[8];
    const float* conv_b[8];
    SEWeights se[3];
    const act16_t* patch_w; const float *patch_b, *patch_pos;
    EncLayerW enc[2];
    const float* global_pos;
    const act16_t *lstm_w_ih, *lstm_w_ih3, *lstm_w_hh, *lstm_w_hh_mma; const float* lstm_b;
    int straggler_threshold = 0; // >0: decode_greedy returns early once <= this many lines are still active;
                                 // the caller re-submits those lines (kocr_read_unfinished) in a later batch
    int big_gemm_sms = 0;        // >0: persistent grid size of the stage 2-5a GEMMs (leave SMs to other streams)
    int lstm_impl = 1;           // 1 = tensor-core recurrence (mma fragments in registers), 0 = CUDA-core / SMEM weights
    const float *dec_tok_emb, *dec_pos;
    DecLayerW dec[2];
    const act16_t* dec_kv_w; const act16_t* dec_kv_w3; const float* dec_out_w; const float *dec_kv_b, *dec_out_b;
    // workspace
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    std::map<std::string, Buf> named;
    Buf pixels_dev, mid_dev, trace;
    uint8_t* staging_host = nullptr;   // pinned
    uint8_t* staging_dev = nullptr;
    size_t staging_bytes = 0;
    cudaEvent_t staging_done = nullptr;
    cudaEvent_t sync_event = nullptr;     // cudaEventBlockingSync: host waits sleep instead of spinning
    cudaEvent_t flag_event[2] = {nullptr, nullptr};   // decode loop: "the n_active copy of group g has landed" (look-ahead ring)
    int32_t* pinned_flag = nullptr;    // pinned int for early-exit polling
    int32_t* fin_host = nullptr;       // pinned copy of the per-line finished flags of the last decode
    // batch state
    int n_lines = 0, n_chunks = 0, n_tok = 0, max_T = 0, n_groups = 0, max_new_w = 0;
    std::vector<int> line_T, line_first_chunk, line_n_chunks;
    int last_steps = 0, last_max_steps = 256;   // positions run / allowed by the last decode (kocr_read_unfinished)
    double host_launch_us = 0, host_wait_us = 0;   // decode loop: host time inside graph / kernel launches and inside stream waits
    // options
    int trace_logits = 0, force_tokens = 0;
    bool have_forced = false;
    cudaStream_t own_stream = nullptr;       // used when the caller passes stream == NULL.  NON-blocking: a blocking stream that is
                                             // capturing a CUDA graph makes every legacy-stream operation of any other host
                                             // thread fail (cudaErrorStreamCaptureImplicit - seen with 8 ranks x 12 passes while
                                             // another thread used torch's default stream); device inputs produced on the default
                                             // stream are ordered explicitly instead (order_after_default_stream)
    cudaEvent_t order_event = nullptr;
    struct DecGraph { cudaGraphExec_t exec = nullptr; size_t nodes = 0; };
    std::map<std::tuple<int, int, int, int>, DecGraph> dec_graphs;   // (n_lines, max_T bucket, trace, force)
    bool decode_warmed = false;
    int use_graphs = 1;
    int use_pdl = 1;             // programmatic dependent launch inside the decode loop
    int compact_rows = 0;        // greedy loop: swap the still-active rows to the front once an M tile of lines has finished
                                 // (identical tokens; measured +0.5 % with 12 batches in flight - the loop is not bound by its GEMM tiles - so off by default)
    int dec_rows = 0;            // rows the decode kernels currently process (<= n_lines)
    std::vector<int> row_orig;   // row -> line of the caller's batch (identity until the loop compacts)
    int32_t* out_stage = nullptr;    // pinned [max_lines][257 + 2]: tokens / lengths / finished in row order before un-permuting
    Buf compact_tab;             // device int2 pairs
    int lstm_split = 1;          // BiLSTM input projection in split precision (same trick as kv_split; 8184 vs 8181 of 8192 c3 lines)
    int kv_split = 1;            // cross-attention K/V projection in split precision (hi + lo operands, K = 3 x 384)
    int blocking_wait = 0;       // 1: host waits sleep on a blocking-sync event (many handles / host threads per process)
    static const int BEAM_MAX = 8;
    Buf tf_x, tf_tab;             // kocr_forward_teacher_forced: padded memory operand, per-line tables
    Buf crop_page, crop_tab;      // kocr_crop_lines: device copy of a host page, boxes + offsets
    Buf beam_scratch;            // kocr_beam_search: parents / tokens / top-k tables
    Buf beam_cache;              // [2 ping-pong][K,V][2 layers][max_lines][DEC_MAX][384] fp32, allocated on first use
    int beam_cur = 0;
    int dec_wide = 1;            // 1: decode GEMMs as 128x64 tiles + split-K over ~50-100 CTAs (lowest latency);
                                 // 0: 128x128 tiles, no split (fewest CTAs: leaves the SMs to other in-flight batches)
    int debug_stop = 0;          // >0 (tests/diagnosis): a decode step launches only its first n kernels
    int dec_lookahead = 0;       // 1: enqueue decode group g+1 before polling the outcome of group g
    int dec_fused = 0;           // 1: GEMM+LayerNorm and out_proj+argmax kernels of dec_fused.cu (18 launches per position); 0: 25 launches
    int pool2_fused = 0;         // 1: the 2x2 max-pool after conv2 runs in conv2's epilogue (4 whole columns = 96 of 128 MMA rows per tile:
                                 // measured 5 % SLOWER than conv2 + pool2x2_kernel, profiles/r02); 0 (default): separate kernel
    int dec_skip = 0;            // diagnosis (tools/inflight_probe.py): bit mask of kernel classes a decode step does NOT launch
                                 // (1 cross-attention, 2 self-attention, 4 LayerNorm, 8 GEMMs, 16 embed + argmax); results are garbage
    // optional per-launch CUDA-event timing of the one-time stages (bench.py roofline)
    int kernel_timing = 0;
    struct Site { std::string name; double flops = 0; double ms = 0; int count = 0; };
    std::vector<Site> sites;
    struct Pending { int site; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    // device arrays inside staging_dev
    LineDesc* d_lines = nullptr; int* d_chunk_line = nullptr; int* d_row_pos = nullptr;
    int* d_line_tok_off = nullptr; int* d_line_T = nullptr; LstmGroup* d_groups = nullptr;
    LstmGroup16* d_groups16 = nullptr; int n_groups16 = 0;
};

namespace {

template <typename T> T* buf(kocr_handle* h, const char* name) { return reinterpret_cast<T*>(h->named[name].p); }

int lookup(kocr_handle* h, const char* name, const void** out, size_t min_bytes) {
    auto it = h->w.find(name);
    KOCR_CHECK(it != h->w.end(), "weight blob lacks entry '%s'", name);
    KOCR_CHECK(it->second.second >= min_bytes, "weight blob entry '%s' has %zu bytes, expected >= %zu", name,
               it->second.second, min_bytes);
    *out = it->second.first;
    return 0;
}
#define W_F32(field, name, n) KOCR_TRY(lookup(h, name, reinterpret_cast<const void**>(&(field)), (size_t)(n) * 4))
#define W_A16(field, name, n) KOCR_TRY(lookup(h, name, reinterpret_cast<const void**>(&(field)), (size_t)(n) * 2))

// channels of the six BasicBlocks in execution order: layer1.0, layer2.0, layer2.1, layer3.0, layer3.1, layer4.0
static const int RES_CIN[6] = {64, 128, 256, 256, 512, 512}, RES_COUT[6] = {128, 256, 256, 512, 512, 512};

int resolve_weights(kocr_handle* h) {
    const int D = D_MODEL;
    static const int cin[8] = {0, 1, 64, 128, 256, 256, 512, 512};
    static const int cout[8] = {0, 64, 128, 256, 256, 512, 512, 512};
    W_F32(h->conv1_w, "conv1.w", 64 * 9);
    W_F32(h->conv1_b, "conv1.b", 64);
    W_A16(h->conv1_w16, "conv1.w16", 64 * 16);
    char nm[64];
    for (int i = 2; i <= 7 && h->variant != 2; ++i) {
        snprintf(nm, sizeof nm, "conv%d.w", i); W_A16(h->conv_w[i], nm, (size_t)cout[i] * 9 * cin[i]);
        snprintf(nm, sizeof nm, "conv%d.b", i); W_F32(h->conv_b[i], nm, cout[i]);
    }
    if (h->variant == 2) {
        for (int b = 0; b < 6; ++b) {
            const int ci = RES_CIN[b], co = RES_COUT[b];
            ResBlockW& r = h->res[b];
            snprintf(nm, sizeof nm, "res%d.c1.w", b); W_A16(r.c1w, nm, (size_t)co * 9 * ci);
            snprintf(nm, sizeof nm, "res%d.c1.b", b); W_F32(r.c1b, nm, co);
            snprintf(nm, sizeof nm, "res%d.c2.w", b); W_A16(r.c2w, nm, (size_t)co * 9 * co);
            snprintf(nm, sizeof nm, "res%d.c2.b", b); W_F32(r.c2b, nm, co);
            r.scw = nullptr; r.scb = nullptr;
            if (ci != co) {
                snprintf(nm, sizeof nm, "res%d.sc.w", b); W_A16(r.scw, nm, (size_t)co * ci);
                snprintf(nm, sizeof nm, "res%d.sc.b", b); W_F32(r.scb, nm, co);
            }
        }
    }
    if (h->variant == 0) {
        static const int sc[3] = {256, 512, 512};
        for (int i = 0; i < 3; ++i) {
            const int C = sc[i];
            snprintf(nm, sizeof nm, "se%d.w0p", i + 3); W_A16(h->se[i].w0p, nm, 128 * C);
            snprintf(nm, sizeof nm, "se%d.b0p", i + 3); W_F32(h->se[i].b0p, nm, 128);
            snprintf(nm, sizeof nm, "se%d.w2p", i + 3); W_A16(h->se[i].w2p, nm, C * 128);
            snprintf(nm, sizeof nm, "se%d.b2", i + 3); W_F32(h->se[i].b2, nm, C);
            snprintf(nm, sizeof nm, "se%d.w0f", i + 3); W_A16(h->se[i].w0f, nm, C / 16 * C);
            snprintf(nm, sizeof nm, "se%d.w2f", i + 3); W_A16(h->se[i].w2f, nm, C * (C / 16));
        }
        W_A16(h->lstm_w_ih, "lstm.w_ih", 8 * LSTM_H * D);
        W_A16(h->lstm_w_ih3, "lstm.w_ih3", 8 * LSTM_H * 3 * D);
        W_F32(h->lstm_b, "lstm.b", 8 * LSTM_H);
        W_A16(h->lstm_w_hh, "lstm.w_hh", bilstm_whh_packed_elems());
        W_A16(h->lstm_w_hh_mma, "lstm.w_hh_mma", bilstm_whh_mma_elems());
    }
    W_A16(h->patch_w, "patch.w", D * 1024);
    W_F32(h->patch_b, "patch.b", D);
    W_F32(h->patch_pos, "patch.pos", 32 * D);
    for (int l = 0; l < 2; ++l) {
        EncLayerW& e = h->enc[l];
        snprintf(nm, sizeof nm, "enc%d.in_w", l); W_A16(e.in_w, nm, 3 * D * D);
        snprintf(nm, sizeof nm, "enc%d.in_b", l); W_F32(e.in_b, nm, 3 * D);
        snprintf(nm, sizeof nm, "enc%d.out_w", l); W_A16(e.out_w, nm, D * D);
        snprintf(nm, sizeof nm, "enc%d.out_b", l); W_F32(e.out_b, nm, D);
        snprintf(nm, sizeof nm, "enc%d.l1_w", l); W_A16(e.l1_w, nm, 1024 * D);
        snprintf(nm, sizeof nm, "enc%d.l1_b", l); W_F32(e.l1_b, nm, 1024);
        snprintf(nm, sizeof nm, "enc%d.l2_w", l); W_A16(e.l2_w, nm, D * 1024);
        snprintf(nm, sizeof nm, "enc%d.l2_b", l); W_F32(e.l2_b, nm, D);
        snprintf(nm, sizeof nm, "enc%d.n1_g", l); W_F32(e.n1_g, nm, D);
        snprintf(nm, sizeof nm, "enc%d.n1_b", l); W_F32(e.n1_b, nm, D);
        snprintf(nm, sizeof nm, "enc%d.n2_g", l); W_F32(e.n2_g, nm, D);
        snprintf(nm, sizeof nm, "enc%d.n2_b", l); W_F32(e.n2_b, nm, D);
    }
    W_F32(h->global_pos, "global_pos", (size_t)h->max_seq_len * D);
    W_F32(h->dec_tok_emb, "dec.tok_emb", (size_t)h->vocab * D);
    W_F32(h->dec_pos, "dec.pos", (size_t)h->dec_max_len * D);
    for (int l = 0; l < 2; ++l) {
        DecLayerW& d = h->dec[l];
        snprintf(nm, sizeof nm, "dec%d.sa_in_w", l); W_F32(d.sa_in_w, nm, 3 * D * D);
        snprintf(nm, sizeof nm, "dec%d.sa_in_b", l); W_F32(d.sa_in_b, nm, 3 * D);
        snprintf(nm, sizeof nm, "dec%d.sa_out_w", l); W_F32(d.sa_out_w, nm, D * D);
        snprintf(nm, sizeof nm, "dec%d.sa_out_b", l); W_F32(d.sa_out_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.ca_q_w", l); W_F32(d.ca_q_w, nm, D * D);
        snprintf(nm, sizeof nm, "dec%d.ca_q_b", l); W_F32(d.ca_q_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.ca_out_w", l); W_F32(d.ca_out_w, nm, D * D);
        snprintf(nm, sizeof nm, "dec%d.ca_out_b", l); W_F32(d.ca_out_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.l1_w", l); W_F32(d.l1_w, nm, 4 * D * D);
        snprintf(nm, sizeof nm, "dec%d.l1_b", l); W_F32(d.l1_b, nm, 4 * D);
        snprintf(nm, sizeof nm, "dec%d.l2_w", l); W_F32(d.l2_w, nm, 4 * D * D);
        snprintf(nm, sizeof nm, "dec%d.l2_b", l); W_F32(d.l2_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n1_g", l); W_F32(d.n1_g, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n1_b", l); W_F32(d.n1_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n2_g", l); W_F32(d.n2_g, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n2_b", l); W_F32(d.n2_b, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n3_g", l); W_F32(d.n3_g, nm, D);
        snprintf(nm, sizeof nm, "dec%d.n3_b", l); W_F32(d.n3_b, nm, D);
    }
    W_A16(h->dec_kv_w, "dec.ca_kv_w", 4 * D * D);
    W_A16(h->dec_kv_w3, "dec.ca_kv_w3", 4 * D * 3 * D);
    W_F32(h->dec_kv_b, "dec.ca_kv_b", 4 * D);
    W_F32(h->dec_out_w, "dec.out_w", VOCAB_PAD * D);
    W_F32(h->dec_out_b, "dec.out_b", VOCAB_PAD);
    return 0;
}

// activation geometry per stage of the backbone (dense NWHC: [chunk][W][H][C]); cols = whole columns per GEMM M tile
struct StageGeom { int H, W, cols; };
const StageGeom G1 = {24, 50, 4}, G2 = {12, 25, 10}, G3 = {6, 25, 21}, G4 = {3, 25, 42};
inline size_t px(const StageGeom& g) { return (size_t)g.H * g.W; }

struct WsItem { const char* name; size_t bytes; };

int ensure(Buf& b, size_t bytes);

// Wait for stream `s` WITHOUT spinning: cudaStreamSynchronize busy-waits by default, and bench.py / a serving process
// keeps a dozen host threads per GPU blocked in here (x 8 GPUs on one host).  A blocking-sync event lets them sleep.
cudaError_t wait_stream(kocr_handle* h, cudaStream_t s) {
    if (!h->blocking_wait) return cudaStreamSynchronize(s);      // one handle, latency matters: spin (a sleep costs ~0.4 ms)
    cudaError_t e = cudaEventRecord(h->sync_event, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(h->sync_event);
}

// The caller handed us DEVICE memory and no stream: whatever it enqueued on the default (legacy) stream to produce that memory
// must complete before our own (non-blocking) stream touches it.
int order_after_default_stream(kocr_handle* h, cudaStream_t s) {
    if (s != h->own_stream) return 0;           // the caller's own stream: the caller orders its work
    KOCR_CUDA(cudaEventRecord(h->order_event, cudaStreamLegacy));
    KOCR_CUDA(cudaStreamWaitEvent(s, h->order_event, 0));
    return 0;
}

int carve_workspace(kocr_handle* h) {
    const size_t NC = h->max_chunks, L = h->max_lines, M = NC * TOK_PER_CHUNK;
    const size_t D = D_MODEL;
    std::vector<WsItem> items = {
        {"chunks", NC * IMG_H * CHUNK_W * 4},
        {"pool1", NC * px(G1) * 64 * 2},   {"conv2", NC * px(G1) * 128 * 2}, {"pool2", NC * px(G2) * 128 * 2},
        {"conv3", NC * px(G2) * 256 * 2},  {"pool3", NC * px(G3) * 256 * 2},
        {"conv5", NC * px(G3) * 512 * 2},  {"pool4", NC * px(G4) * 512 * 2},
        {"bins7", NC * 25 * 2 * 512 * 2},  {"patch_in", M * 1024 * 2},
        {"se_mean3", NC * 25 * 256 * 2},   {"se_mean4", NC * 25 * 512 * 2}, {"se_mean5", NC * 25 * 512 * 2},
        {"x", M * D * 4},   {"xb", M * D * 2},  {"qkv", M * 3 * D * 2}, {"ao", M * D * 2}, {"y", M * D * 4},
        {"hff", M * 1024 * 2},
        {"gin", M * 8 * LSTM_H * 4}, {"mem", M * D * 4}, {"memb", M * D * 2}, {"kv", M * 4 * D * 2}, {"kv_a3", M * 3 * D * 2},
        {"vtab", L * (size_t)preprocess_vtab_ints_per_line() * 4},
        {"tokens", L * KOCR_TOKENS_LD * 4}, {"forced", L * KOCR_TOKENS_LD * 4}, {"lengths", L * 4},
        {"finished", L * 4}, {"n_active", (DEC_MAX + 1) * 4}, {"step_base", 64},
        {"dx", L * D * 4}, {"dxb", L * D * 2}, {"dqkv", L * 3 * D * 4}, {"dao", L * D * 2}, {"dy", L * D * 4},
        {"dq", L * D * 4}, {"dh", L * 4 * D * 4}, {"daof", L * D * 4}, {"logits", L * VOCAB_PAD * 4},
        {"dparts", 8 * L * 3 * D * 4},
        {"kcache", 2 * L * DEC_MAX * D * 4}, {"vcache", 2 * L * DEC_MAX * D * 4},
    };
    if (h->variant == 2) {      // ResNet baseline: un-pooled block outputs, second 24x50x128 activation, the fp32 shortcut (largest: layer1)
        items.push_back({"conv4", NC * px(G2) * 256 * 2});
        items.push_back({"conv6", NC * px(G3) * 512 * 2});
        items.push_back({"conv7", NC * px(G4) * 512 * 2});
        items.push_back({"res_t1", NC * px(G1) * 128 * 2});
        items.push_back({"res_add", NC * px(G1) * 128 * 4});
    }
    size_t total = 0;
    for (auto& it : items) total += (it.bytes + 1023) / 1024 * 1024;
    KOCR_CUDA(cudaMalloc(&h->ws, total));
    KOCR_CUDA(cudaMemset(h->ws, 0, total));
    h->ws_bytes = total;
    size_t off = 0;
    for (auto& it : items) {
        Buf b; b.p = h->ws + off; b.bytes = it.bytes;
        h->named[it.name] = b;
        off += (it.bytes + 1023) / 1024 * 1024;
    }
    // staging for the small per-batch integer tables
    h->staging_bytes = L * sizeof(LineDesc) + NC * 4 + M * 4 + L * 4 * 2 + (L / 8 + 2) * sizeof(LstmGroup) +
                       (L / 16 + 2) * sizeof(LstmGroup16) + 4096;
    KOCR_CUDA(cudaMallocHost(&h->staging_host, h->staging_bytes));
    KOCR_CUDA(cudaMalloc(&h->staging_dev, h->staging_bytes));
    KOCR_CUDA(cudaMallocHost(&h->pinned_flag, 64));
    KOCR_CUDA(cudaMallocHost(&h->fin_host, (size_t)h->max_lines * 4));
    KOCR_CUDA(cudaMallocHost(&h->out_stage, (size_t)h->max_lines * (KOCR_TOKENS_LD + 2) * 4));
    KOCR_CUDA(cudaEventCreateWithFlags(&h->staging_done, cudaEventDisableTiming | cudaEventBlockingSync));
    KOCR_CUDA(cudaEventCreateWithFlags(&h->sync_event, cudaEventDisableTiming | cudaEventBlockingSync));
    for (int i = 0; i < 2; ++i) KOCR_CUDA(cudaEventCreateWithFlags(&h->flag_event[i], cudaEventDisableTiming | cudaEventBlockingSync));
    KOCR_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    KOCR_CUDA(cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming));
    // input staging sized for the handle's capacity up front (4x the bytes of the height-48 chunks: source lines are
    // rarely more than 2x oversampled): a cudaMalloc in the middle of a run would synchronise every in-flight batch
    KOCR_TRY(ensure(h->pixels_dev, (size_t)h->max_chunks * CHUNK_STRIDE * IMG_H * 4));
    KOCR_TRY(ensure(h->mid_dev, (size_t)h->max_chunks * CHUNK_STRIDE * IMG_H * 2));
    return 0;
}

int ensure(Buf& b, size_t bytes) {
    if (b.bytes >= bytes) return 0;
    if (b.p) KOCR_CUDA(cudaFree(b.p));
    b.p = nullptr; b.bytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    KOCR_CUDA(cudaMalloc(&b.p, want));
    b.bytes = want;
    return 0;
}

// ---- per-launch timing ------------------------------------------------------------------------
struct SiteTimer {
    kocr_handle* h; cudaStream_t s; int idx = -1; cudaEvent_t a = nullptr, b = nullptr;
    SiteTimer(kocr_handle* h_, const char* name, double flops, cudaStream_t s_) : h(h_), s(s_) {
        if (!h->kernel_timing) return;
        for (size_t i = 0; i < h->sites.size(); ++i) if (h->sites[i].name == name) idx = (int)i;
        if (idx < 0) { kocr_handle::Site st; st.name = name; h->sites.push_back(st); idx = (int)h->sites.size() - 1; }
        h->sites[idx].flops += flops;
        auto get = [&]() { cudaEvent_t e = nullptr; if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        a = get(); b = get();
        cudaEventRecord(a, s);
    }
    ~SiteTimer() {
        if (idx < 0) return;
        cudaEventRecord(b, s);
        kocr_handle::Pending p; p.site = idx; p.a = a; p.b = b;
        h->pending.push_back(p);
    }
};
#define TIMED(name, flops, call) do { SiteTimer _t(h, name, flops, s); KOCR_TRY(call); } while (0)

// ---- GEMM convenience wrappers -------------------------------------------------------------
GemmEpilogue ep_none() {
    GemmEpilogue e;
    memset(&e, 0, sizeof e);
    return e;
}

int gemm_linear(kocr_handle* h, const void* a, long rows, const void* w, int N, int K,
                const GemmEpilogue& ep, cudaStream_t s, int tf32 = 0) {
    GemmProblem p;
    memset(&p, 0, sizeof p);
    p.M = (int)rows; p.N = N; p.taps = 1; p.cin = K; p.ep = ep; p.tf32 = tf32;
    const int sms = (h->big_gemm_sms > 0 && rows > 4096) ? std::min(h->big_gemm_sms, h->num_sms) : h->num_sms;
    return launch_gemm_tc(a, rows, w, p, sms, s);
}

// 3x3 / pad-1 convolution + folded BN (+ ReLU) as an implicit GEMM over the dense NWHC activation (TMA im2col mode).
// Standard form: out = act(conv + bias (+ addend)) [chunk][W][H][Cout].
int gemm_conv(kocr_handle* h, const act16_t* in, act16_t* out, int n_chunks, const StageGeom& g, int Cin,
              int Cout, const act16_t* w, const float* b, int relu, cudaStream_t s, const float* addend = nullptr) {
    GemmProblem p;
    memset(&p, 0, sizeof p);
    p.M = n_chunks * g.H * g.W; p.N = Cout; p.taps = 9; p.cin = Cin;
    p.conv_H = g.H; p.conv_W = g.W; p.n_img = n_chunks; p.tile_cols = 0;      // 128 consecutive pixels per tile
    p.ep = ep_none();
    p.ep.bias = b; p.ep.relu = relu;
    p.ep.out_a16 = out; p.ep.ld_a16 = Cout;
    if (addend) {               // residual connection: out = act(conv + bias + addend); the addend path needs the 128-wide N tile
        p.ep.addend = addend; p.ep.ld_add = Cout; p.ep.add_period = 0;
        p.bn = 128;
    }
    const int sms = h->big_gemm_sms > 0 ? std::min(h->big_gemm_sms, h->num_sms) : h->num_sms;
    return launch_gemm_tc(in, (long)p.M, w, p, sms, s);
}

// The same convolution with the column-fused epilogue (whole-column M tiles): writes the (2,1)-max-pooled rows
// (mode 1: [col][H/2][Cout]) or the AdaptiveAvgPool row-bin sums (mode 2, H = 3: [col][2][Cout]) and, if `colmean` is
// given, the SequenceSE column means [col][Cout] fp32 - the un-pooled conv output is never stored.
int gemm_conv_colfused(kocr_handle* h, const act16_t* in, act16_t* out_pool, act16_t* colmean, int mode, int n_chunks,
                       const StageGeom& g, int Cin, int Cout, const act16_t* w, const float* b, int relu, cudaStream_t s) {
    GemmProblem p;
    memset(&p, 0, sizeof p);
    p.M = n_chunks * g.H * g.W; p.N = Cout; p.taps = 9; p.cin = Cin;
    p.conv_H = g.H; p.conv_W = g.W; p.n_img = n_chunks; p.tile_cols = g.cols;
    p.ep = ep_none();
    p.ep.bias = b; p.ep.relu = relu;
    p.ep.col_mode = mode; p.ep.out_pool = out_pool; p.ep.out_colmean = colmean;
    p.bn = Cout % 256 == 0 ? 256 : 128;
    const int sms = h->big_gemm_sms > 0 ? std::min(h->big_gemm_sms, h->num_sms) : h->num_sms;
    return launch_gemm_tc(in, (long)p.M, w, p, sms, s);
}

// One BasicBlock (model/resnet_model.py:28-37) on the implicit-GEMM kernel: c1 = relu(bn1(conv1 x)); shortcut as an
// fp32 tensor (1x1 conv + BN as a per-position linear layer, or a widening copy of x); out = relu(bn2(conv2 c1) + shortcut)
// through the GEMM's fp32 addend.  `out` may alias `in` for an identity block (the addend was copied out before).
int resnet_block(kocr_handle* h, int bi, const act16_t* in, act16_t* tmp, act16_t* out, int NC, const StageGeom& g, cudaStream_t s) {
    const int ci = RES_CIN[bi], co = RES_COUT[bi];
    const ResBlockW& r = h->res[bi];
    const long rows = (long)NC * g.H * g.W;
    float* add = buf<float>(h, "res_add");
    char nm[32];
    snprintf(nm, sizeof nm, "res%d_conv1", bi);
    TIMED(nm, 2.0 * NC * g.H * g.W * 9.0 * ci * co, gemm_conv(h, in, tmp, NC, g, ci, co, r.c1w, r.c1b, 1, s));
    snprintf(nm, sizeof nm, "res%d_shortcut", bi);
    if (r.scw) {
        GemmEpilogue e = ep_none();
        e.bias = r.scb; e.out_f32 = add; e.ld_f32 = co;
        TIMED(nm, 2.0 * NC * g.H * g.W * ci * co, gemm_linear(h, in, rows, r.scw, co, ci, e, s));
    } else {
        TIMED(nm, 0, launch_a16_to_f32(in, add, rows * ci, s)); ++g_launches;
    }
    snprintf(nm, sizeof nm, "res%d_conv2", bi);
    TIMED(nm, 2.0 * NC * g.H * g.W * 9.0 * co * co, gemm_conv(h, tmp, out, NC, g, co, co, r.c2w, r.c2b, 1, s, add));
    return 0;
}

int stage_resnet_backbone(kocr_handle* h, cudaStream_t s) {
    const int NC = h->n_chunks;
    auto B = [&](const char* n) { return buf<act16_t>(h, n); };
    TIMED("conv1_pool1", 2.0 * NC * 48 * 100 * 9.0 * 64, launch_conv1_pool_mma(buf<float>(h, "chunks"), h->conv1_w16, h->conv1_b, B("pool1"), NC, s));
    ++g_launches;
    KOCR_TRY(resnet_block(h, 0, B("pool1"), B("res_t1"), B("conv2"), NC, G1, s));                    // layer1
    TIMED("pool2", 0, launch_pool2x2(B("conv2"), B("pool2"), NC, 24, 50, 128, s)); ++g_launches;
    KOCR_TRY(resnet_block(h, 1, B("pool2"), B("conv3"), B("conv4"), NC, G2, s));                     // layer2.0
    KOCR_TRY(resnet_block(h, 2, B("conv4"), B("conv3"), B("conv4"), NC, G2, s));                     // layer2.1 (identity)
    TIMED("pool3", 0, launch_pool_h2(B("conv4"), B("pool3"), NC, 12, 25, 256, s)); ++g_launches;
    KOCR_TRY(resnet_block(h, 3, B("pool3"), B("conv5"), B("conv6"), NC, G3, s));                     // layer3.0
    KOCR_TRY(resnet_block(h, 4, B("conv6"), B("conv5"), B("conv6"), NC, G3, s));                     // layer3.1 (identity)
    TIMED("pool4", 0, launch_pool_h2(B("conv6"), B("pool4"), NC, 6, 25, 512, s)); ++g_launches;
    KOCR_TRY(resnet_block(h, 5, B("pool4"), B("conv7"), B("pool4"), NC, G4, s));                     // layer4 (identity)
    TIMED("final_pool", 0, launch_finalpool(B("pool4"), 3, B("patch_in"), NC, 25, 512, s)); ++g_launches;
    return 0;
}

// parts: 1 = backbone (chunks -> patch_in), 2 = patch projection (patch_in -> x), 4 = encoder layers (x -> x); the whole
// path runs all three with the merge's global_pos added by the last LayerNorm (the model-protocol entry points run one part)
int stage_cnn_encoder(kocr_handle* h, cudaStream_t s, int parts = 7, bool add_global_pos = true) {
    const int NC = h->n_chunks;
    if (NC == 0) return 0;
    const bool se = h->variant == 0;
    auto B = [&](const char* n) { return buf<act16_t>(h, n); };
    const double nc = NC;
    auto cf = [&](int H, int W, int ci, int co) { return 2.0 * nc * H * W * 9.0 * ci * co; };   // algorithmic conv FLOPs
    if (!(parts & 1)) {
    } else if (h->variant == 2) {
        KOCR_TRY(stage_resnet_backbone(h, s));
    } else {
    // SE-VGG (se_model.py:63-79) / VGG baseline (vgg_model.py:50-59).  conv4 / conv6 / conv7 never store their un-pooled
    // output: their epilogue emits the (2,1)-max-pooled rows (conv7: the adaptive-pool row bins) plus the SE column means;
    // the SE excitation then scales the pooled tensor in place (gate > 0 commutes with the max).
    TIMED("conv1_pool1", cf(48, 100, 1, 64), launch_conv1_pool_mma(buf<float>(h, "chunks"), h->conv1_w16, h->conv1_b, B("pool1"), NC, s));
    ++g_launches;
    if (h->pool2_fused) {       // conv2 + 2x2 max-pool in one kernel (4 whole columns per M tile): the 24x50x128 conv output never reaches HBM
        TIMED("conv2", cf(24, 50, 64, 128), gemm_conv_colfused(h, B("pool1"), B("pool2"), nullptr, 3, NC, G1, 64, 128, h->conv_w[2], h->conv_b[2], 1, s));
    } else {
        TIMED("conv2", cf(24, 50, 64, 128), gemm_conv(h, B("pool1"), B("conv2"), NC, G1, 64, 128, h->conv_w[2], h->conv_b[2], 1, s));
        TIMED("pool2", 0, launch_pool2x2(B("conv2"), B("pool2"), NC, 24, 50, 128, s)); ++g_launches;
    }
    TIMED("conv3", cf(12, 25, 128, 256), gemm_conv(h, B("pool2"), B("conv3"), NC, G2, 128, 256, h->conv_w[3], h->conv_b[3], 1, s));
    TIMED("conv4", cf(12, 25, 256, 256), gemm_conv_colfused(h, B("conv3"), B("pool3"), se ? buf<act16_t>(h, "se_mean3") : nullptr, 1, NC, G2, 256, 256,
                                                            h->conv_w[4], h->conv_b[4], 1, s));
    if (se) { TIMED("se3_excite", 4.0 * nc * 25 * 256 * 16, launch_se_excite(buf<act16_t>(h, "se_mean3"), h->se[0], B("pool3"), nullptr, NC, 6, 25, 256, false, s)); ++g_launches; }
    TIMED("conv5", cf(6, 25, 256, 512), gemm_conv(h, B("pool3"), B("conv5"), NC, G3, 256, 512, h->conv_w[5], h->conv_b[5], 1, s));
    TIMED("conv6", cf(6, 25, 512, 512), gemm_conv_colfused(h, B("conv5"), B("pool4"), se ? buf<act16_t>(h, "se_mean4") : nullptr, 1, NC, G3, 512, 512,
                                                           h->conv_w[6], h->conv_b[6], 1, s));
    if (se) { TIMED("se4_excite", 4.0 * nc * 25 * 512 * 32, launch_se_excite(buf<act16_t>(h, "se_mean4"), h->se[1], B("pool4"), nullptr, NC, 3, 25, 512, false, s)); ++g_launches; }
    // conv7: SE model = conv + bn7 + relu7 (se_model.py:75); VGG baseline = bare conv (vgg_model.py:57)
    TIMED("conv7", cf(3, 25, 512, 512), gemm_conv_colfused(h, B("pool4"), B("bins7"), se ? buf<act16_t>(h, "se_mean5") : nullptr, 2, NC, G4, 512, 512,
                                                           h->conv_w[7], h->conv_b[7], se ? 1 : 0, s));
    if (se) { TIMED("se5_excite_finalpool", 4.0 * nc * 25 * 512 * 32, launch_se_excite(buf<act16_t>(h, "se_mean5"), h->se[2], B("bins7"), B("patch_in"), NC, 2, 25, 512, true, s)); ++g_launches; }
    else { TIMED("final_pool", 0, launch_finalpool(B("bins7"), 2, B("patch_in"), NC, 25, 512, s)); ++g_launches; }
    }   // SE / VGG backbone

    const long M = (long)NC * TOK_PER_CHUNK;
    float* x = buf<float>(h, "x"); float* y = buf<float>(h, "y");
    act16_t* xb = B("xb");
    if (parts & 2) {   // patch projection + bias + local positional encoding (se_model.py:108-115)
        GemmEpilogue e = ep_none();
        e.bias = h->patch_b; e.addend = h->patch_pos; e.ld_add = D_MODEL; e.add_period = TOK_PER_CHUNK;
        e.out_f32 = x; e.ld_f32 = D_MODEL; e.out_a16 = xb; e.ld_a16 = D_MODEL;
        TIMED("patch_proj", 2.0 * M * 1024 * D_MODEL, gemm_linear(h, B("patch_in"), M, h->patch_w, D_MODEL, 1024, e, s));
    }
    for (int l = 0; l < 2 && (parts & 4); ++l) {
        const EncLayerW& w = h->enc[l];
        GemmEpilogue e = ep_none();
        e.bias = w.in_b; e.out_a16 = B("qkv"); e.ld_a16 = 3 * D_MODEL;
        TIMED("enc_qkv", 2.0 * M * D_MODEL * 3 * D_MODEL, gemm_linear(h, xb, M, w.in_w, 3 * D_MODEL, D_MODEL, e, s));
        TIMED("enc_attention", 2.0 * 2.0 * M * 32 * D_MODEL, launch_chunk_attention(B("qkv"), B("ao"), NC, s)); ++g_launches;
        e = ep_none();
        e.bias = w.out_b; e.addend = x; e.ld_add = D_MODEL; e.out_f32 = y; e.ld_f32 = D_MODEL;
        TIMED("enc_out_proj", 2.0 * M * D_MODEL * D_MODEL, gemm_linear(h, B("ao"), M, w.out_w, D_MODEL, D_MODEL, e, s));
        TIMED("enc_layernorm", 0, launch_layernorm(y, w.n1_g, w.n1_b, nullptr, nullptr, x, xb, nullptr, (int)M, s)); ++g_launches;
        e = ep_none();
        e.bias = w.l1_b; e.relu = 1; e.out_a16 = B("hff"); e.ld_a16 = 1024;
        TIMED("enc_ffn1", 2.0 * M * D_MODEL * 1024, gemm_linear(h, xb, M, w.l1_w, 1024, D_MODEL, e, s));
        e = ep_none();
        e.bias = w.l2_b; e.addend = x; e.ld_add = D_MODEL; e.out_f32 = y; e.ld_f32 = D_MODEL;
        TIMED("enc_ffn2", 2.0 * M * D_MODEL * 1024, gemm_linear(h, B("hff"), M, w.l2_w, D_MODEL, 1024, e, s));
        // last layer: fuse the merge's "+ global_pos[t]" (predictor.py:178-183) into the LayerNorm
        const bool last = l == 1 && add_global_pos;
        TIMED("enc_layernorm", 0, launch_layernorm(y, w.n2_g, w.n2_b, last ? h->global_pos : nullptr, last ? h->d_row_pos : nullptr, x,
                                  xb, nullptr, (int)M, s)); ++g_launches;
    }
    return 0;
}

// Cross-attention K/V of both decoder layers for `rows` memory rows (once per line: depends only on the memory).
// mem_f32 / memb: the same memory in fp32 and in 16 bits.
int project_cross_kv(kocr_handle* h, long rows, const float* mem_f32, const act16_t* memb, cudaStream_t s) {
    GemmEpilogue e = ep_none();
    e.bias = h->dec_kv_b; e.out_a16 = buf<act16_t>(h, "kv"); e.ld_a16 = 4 * D_MODEL;
    if (h->kv_split) {
        act16_t* a3 = buf<act16_t>(h, "kv_a3");
        TIMED("cross_kv_split", 0, launch_split3(mem_f32, a3, rows, s)); ++g_launches;
        TIMED("cross_kv_proj", 2.0 * rows * D_MODEL * 4 * D_MODEL, gemm_linear(h, a3, rows, h->dec_kv_w3, 4 * D_MODEL, 3 * D_MODEL, e, s));
    } else {
        TIMED("cross_kv_proj", 2.0 * rows * D_MODEL * 4 * D_MODEL, gemm_linear(h, memb, rows, h->dec_kv_w, 4 * D_MODEL, D_MODEL, e, s));
    }
    return 0;
}

int stage_memory(kocr_handle* h, cudaStream_t s, int parts = 3) {
    const long M = h->n_tok;
    if (M == 0) return 0;
    const act16_t* memb = buf<act16_t>(h, "xb");
    if (h->variant == 0 && !(parts & 1)) memb = buf<act16_t>(h, "memb");
    if (h->variant == 0 && (parts & 1)) {
        GemmEpilogue e = ep_none();
        e.bias = h->lstm_b; e.out_f32 = buf<float>(h, "gin"); e.ld_f32 = 8 * LSTM_H;
        if (h->lstm_split) {        // split-precision input projection: fp32 merged sequence as [hi | lo | hi] rows (see project_cross_kv)
            act16_t* a3 = buf<act16_t>(h, "kv_a3");
            TIMED("lstm_in_split", 0, launch_split3(buf<float>(h, "x"), a3, M, s)); ++g_launches;
            TIMED("lstm_in_proj", 2.0 * M * D_MODEL * 8 * LSTM_H, gemm_linear(h, a3, M, h->lstm_w_ih3, 8 * LSTM_H, 3 * D_MODEL, e, s));
        } else
        TIMED("lstm_in_proj", 2.0 * M * D_MODEL * 8 * LSTM_H, gemm_linear(h, buf<act16_t>(h, "xb"), M, h->lstm_w_ih, 8 * LSTM_H, D_MODEL, e, s));
        if (h->lstm_impl == 1)
            TIMED("bilstm_recurrence", 2.0 * M * 8 * LSTM_H * LSTM_H, launch_bilstm_mma(buf<float>(h, "gin"), h->lstm_w_hh_mma, h->d_line_tok_off, h->d_line_T, h->d_groups16,
                               h->n_groups16, buf<float>(h, "mem"), buf<act16_t>(h, "memb"), nullptr, s));
        else
        TIMED("bilstm_recurrence", 2.0 * M * 8 * LSTM_H * LSTM_H, launch_bilstm(buf<float>(h, "gin"), h->lstm_w_hh, h->d_line_tok_off, h->d_line_T, h->d_groups,
                               h->n_groups, buf<float>(h, "mem"), buf<act16_t>(h, "memb"), nullptr, s));
        ++g_launches;
        memb = buf<act16_t>(h, "memb");
    }
    if (!(parts & 2)) return 0;
    return project_cross_kv(h, M, buf<float>(h, h->variant == 0 ? "mem" : "x"), memb, s);
}

// Decoder GEMM for a handful of rows: TF32, 128x64 tiles and split-K so that ~50-100 CTAs each stream one
// pipeline-fill of operands (the kernel is a latency chain, not a throughput problem); the slices are written
// as raw partial sums [split][L][N] and added up (with bias / residual) by the consuming kernel.
int gemm_dec(kocr_handle* h, const float* a, int L, const float* w, int N, int K, int split, float* parts,
             cudaStream_t s) {
    GemmProblem p;
    memset(&p, 0, sizeof p);
    p.M = L; p.N = N; p.taps = 1; p.cin = K; p.tf32 = 1; p.split_k = split; p.bn = h->dec_wide ? 64 : 128;
    p.ep.out_f32 = parts; p.ep.ld_f32 = N;
    return launch_gemm_tc(a, L, w, p, h->num_sms, s);
}

// One generated position for every line of the batch; the position is *step_base + off (device side).
// Decoder GEMMs run on the tensor cores in TF32 (fp32 operands): with a16 operands ~7 % of the lines of the
// fixture batch decode to a different sequence than the fp32 reference, with TF32 the flips disappear.
// Optional override: the rows are `n_rows` hypotheses (beams) that share the memory of ONE line and use a private,
// small self-attention cache.
struct DecRows {
    int n_rows;
    float *kcache, *vcache;          // [2 layers][max_lines][DEC_MAX][384]
    size_t layer_stride;
    const int *tok_off, *T;          // device arrays [n_rows]
};

int decode_step(kocr_handle* h, int off, int max_T, cudaStream_t s, const DecRows* rows = nullptr, bool embed_first = true) {
    const int L = rows ? rows->n_rows : h->dec_rows;
    const int D = D_MODEL;
    int n_launched = 0;
#define DSTEP(call) do { if (h->debug_stop <= 0 || n_launched < h->debug_stop) { KOCR_TRY(call); } ++n_launched; } while (0)
#define DSKIP(bit, stmt) do { if (!(h->dec_skip & (bit))) { stmt; } } while (0)
    int* tokens = buf<int>(h, "tokens");
    const int* sb = buf<int>(h, "step_base");
    const int* fin = buf<int>(h, "finished");
    float* dx = buf<float>(h, "dx");                 // residual stream, exact fp32
    float* dxt = buf<float>(h, "dq");                // the same rows rounded to TF32 (nearest): A operand of the next GEMM
    float* parts = buf<float>(h, "dparts");          // split-K partial results of the current projection
    float* dao = buf<float>(h, "daof");
    float* dh = buf<float>(h, "dh");
    const int S2 = h->dec_wide ? 2 : 1, S8 = h->dec_wide ? 8 : 1;   // K = 384 -> 2 slices of 6 k-blocks; K = 1536 -> 8
    const bool fused = h->dec_fused != 0;
    (void)embed_first;
    DSKIP(16, DSTEP(launch_dec_embed(tokens, sb, off, h->dec_tok_emb, h->dec_pos, dx, dxt, nullptr, nullptr, L, s)); ++g_launches);
    for (int l = 0; l < 2; ++l) {
        const DecLayerW& w = h->dec[l];
        float* kc = rows ? rows->kcache + l * rows->layer_stride : buf<float>(h, "kcache") + (size_t)l * h->max_lines * DEC_MAX * D;
        float* vc = rows ? rows->vcache + l * rows->layer_stride : buf<float>(h, "vcache") + (size_t)l * h->max_lines * DEC_MAX * D;
        DSKIP(8, DSTEP(gemm_dec(h, dxt, L, w.sa_in_w, 3 * D, D, S2, parts, s)));
        DSKIP(2, DSTEP(launch_dec_self_attn(parts, kc, vc, tokens, sb, off, fin, dao, L, s, S2, w.sa_in_b)); ++g_launches);
        if (fused) DSKIP(8, DSTEP(launch_dec_gemm_ln(dao, L, D, w.sa_out_w, w.sa_out_b, dx, w.n1_g, w.n1_b, dx, dxt, s)));
        else {
        DSKIP(8, DSTEP(gemm_dec(h, dao, L, w.sa_out_w, D, D, S2, parts, s)));
        DSKIP(4, DSTEP(launch_layernorm(parts, w.n1_g, w.n1_b, nullptr, nullptr, dx, nullptr, nullptr, L, s, S2, w.sa_out_b, dx, dxt)); ++g_launches);
        }
        DSKIP(8, DSTEP(gemm_dec(h, dxt, L, w.ca_q_w, D, D, S2, parts, s)));
        DSKIP(1, DSTEP(launch_dec_cross_attn(parts, buf<act16_t>(h, "kv"), l, rows ? rows->tok_off : h->d_line_tok_off,
                                    rows ? rows->T : h->d_line_T, max_T, fin,
                                    dao, L, s, S2, w.ca_q_b)); ++g_launches);
        if (fused) DSKIP(8, DSTEP(launch_dec_gemm_ln(dao, L, D, w.ca_out_w, w.ca_out_b, dx, w.n2_g, w.n2_b, dx, dxt, s)));
        else {
        DSKIP(8, DSTEP(gemm_dec(h, dao, L, w.ca_out_w, D, D, S2, parts, s)));
        DSKIP(4, DSTEP(launch_layernorm(parts, w.n2_g, w.n2_b, nullptr, nullptr, dx, nullptr, nullptr, L, s, S2, w.ca_out_b, dx, dxt)); ++g_launches);
        }
        {   // FFN1 keeps its ReLU epilogue (no split): N = 1536 already gives 48 CTAs
            GemmProblem p;
            memset(&p, 0, sizeof p);
            p.M = L; p.N = 4 * D; p.taps = 1; p.cin = D; p.tf32 = 1; p.bn = h->dec_wide ? 64 : 256;
            p.ep.bias = w.l1_b; p.ep.relu = 1; p.ep.out_f32 = dh; p.ep.ld_f32 = 4 * D; p.ep.round_tf32 = 1;
            DSKIP(8, DSTEP(launch_gemm_tc(dxt, L, w.l1_w, p, h->num_sms, s)));
        }
        if (fused) DSKIP(8, DSTEP(launch_dec_gemm_ln(dh, L, 4 * D, w.l2_w, w.l2_b, dx, w.n3_g, w.n3_b, dx, dxt, s)));
        else {
        DSKIP(8, DSTEP(gemm_dec(h, dh, L, w.l2_w, D, 4 * D, S8, parts, s)));
        DSKIP(4, DSTEP(launch_layernorm(parts, w.n3_g, w.n3_b, nullptr, nullptr, dx, nullptr, nullptr, L, s, S8, w.l2_b, dx, dxt)); ++g_launches);
        }
    }
    const int* forced = (h->force_tokens && h->have_forced) ? buf<int>(h, "forced") : nullptr;
    float* trace = (h->trace_logits || rows) ? reinterpret_cast<float*>(h->trace.p) : nullptr;
    if (fused) {
        DSKIP(8, DSTEP(launch_dec_out_argmax(dxt, L, h->dec_out_w, h->dec_out_b, tokens, buf<int>(h, "lengths"), buf<int>(h, "finished"),
                                             buf<int>(h, "n_active"), sb, off, forced, trace, s)));
        return 0;
    }
    DSKIP(8, DSTEP(gemm_dec(h, dxt, L, h->dec_out_w, VOCAB_PAD, D, S2, parts, s)));
    DSKIP(16, DSTEP(launch_dec_argmax(parts, tokens, buf<int>(h, "lengths"), buf<int>(h, "finished"), buf<int>(h, "n_active"), sb,
                            off, L, forced, trace, s, S2, h->dec_out_b)); ++g_launches);
#undef DSTEP
#undef DSKIP
    return 0;
}

static const int DEC_GROUP = 8;     // positions per captured graph / per early-exit poll

// `n` consecutive positions followed by the step_base bump, eagerly on stream s.
int decode_group_eager(kocr_handle* h, int n, int max_T, cudaStream_t s) {
    struct PdlScope { PdlScope(bool on) { pdl_set_active(on); } ~PdlScope() { pdl_set_active(false); } } scope(h->use_pdl != 0);
    for (int i = 0; i < n; ++i) KOCR_TRY(decode_step(h, i, max_T, s, nullptr, /*embed_first=*/false));
    KOCR_TRY(launch_dec_bump(buf<int>(h, "step_base"), n, s)); ++g_launches;
    return 0;
}

// The same 8 positions as one CUDA graph (captured once per (n_lines, max_T bucket, options)).
int decode_group_graph(kocr_handle* h, int max_T, cudaStream_t s) {
    auto key = std::make_tuple(h->dec_rows, max_T, h->dec_skip * 16 + h->dec_fused * 8 + h->trace_logits * 4 + h->use_pdl * 2 + h->dec_wide, (h->force_tokens && h->have_forced) ? 1 : 0);
    auto it = h->dec_graphs.find(key);
    if (it == h->dec_graphs.end()) {
        cudaGraph_t graph = nullptr;
        KOCR_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = decode_group_eager(h, DEC_GROUP, max_T, s);
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        KOCR_CHECK(ce == cudaSuccess && graph != nullptr, "decode graph capture failed: %s", cudaGetErrorString(ce));
        kocr_handle::DecGraph g;
        // kernel nodes of the graph = launches this thread "made" while capturing (the global counters are shared with the
        // other in-flight handles' threads, so a before/after difference would over-count): captured, not launched -
        // take them back out and count them per replay below
        size_t n_nodes = 0;
        cudaError_t ne = cudaGraphGetNodes(graph, nullptr, &n_nodes);
        ce = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        KOCR_CHECK(ne == cudaSuccess, "cudaGraphGetNodes failed: %s", cudaGetErrorString(ne));
        KOCR_CHECK(ce == cudaSuccess, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        g.nodes = n_nodes;
        g_launches -= (int64_t)g.nodes;
        if (h->dec_graphs.size() > 64) {
            for (auto& d : h->dec_graphs) cudaGraphExecDestroy(d.second.exec);
            h->dec_graphs.clear();
        }
        it = h->dec_graphs.emplace(key, g).first;
    }
    KOCR_CUDA(cudaGraphLaunch(it->second.exec, s));
    g_launches += (int64_t)it->second.nodes;
    return 0;
}


// ---- tables of a batch given as token counts (model-protocol entry points; kocr_gather_chunks derives them from images) --
// Line i owns rows [off_i, off_i + T_i) of the token-major buffers, off_i = sum of the previous lines' T rounded up to 32.
int set_token_tables(kocr_handle* h, int n_lines, const int* T, cudaStream_t s) {
    KOCR_CHECK(n_lines > 0 && n_lines <= h->max_lines, "model call: %d lines exceed the handle's capacity %d", n_lines, h->max_lines);
    KOCR_CUDA(cudaEventSynchronize(h->staging_done));
    uint8_t* sp = h->staging_host;
    LineDesc* lines = reinterpret_cast<LineDesc*>(sp); sp += (size_t)h->max_lines * sizeof(LineDesc);
    int* chunk_line = reinterpret_cast<int*>(sp); sp += (size_t)h->max_chunks * 4;
    int* row_pos = reinterpret_cast<int*>(sp); sp += (size_t)h->max_chunks * TOK_PER_CHUNK * 4;
    int* tok_off = reinterpret_cast<int*>(sp); sp += (size_t)h->max_lines * 4;
    int* lineT = reinterpret_cast<int*>(sp); sp += (size_t)h->max_lines * 4;
    sp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~uintptr_t(15));
    LstmGroup* groups = reinterpret_cast<LstmGroup*>(sp);
    auto dev_of = [&](void* hp) { return h->staging_dev + (reinterpret_cast<uint8_t*>(hp) - h->staging_host); };
    h->d_lines = reinterpret_cast<LineDesc*>(dev_of(lines));
    h->d_chunk_line = reinterpret_cast<int*>(dev_of(chunk_line));
    h->d_row_pos = reinterpret_cast<int*>(dev_of(row_pos));
    h->d_line_tok_off = reinterpret_cast<int*>(dev_of(tok_off));
    h->d_line_T = reinterpret_cast<int*>(dev_of(lineT));
    h->d_groups = reinterpret_cast<LstmGroup*>(dev_of(groups));
    h->n_lines = n_lines; h->max_T = 0; h->max_new_w = 0;
    h->line_T.assign(n_lines, 0); h->line_first_chunk.assign(n_lines, 0); h->line_n_chunks.assign(n_lines, 0);
    int nc = 0;
    for (int i = 0; i < n_lines; ++i) {
        KOCR_CHECK(T[i] > 0 && T[i] <= h->max_seq_len, "model call: line %d has %d tokens (1..%d supported)", i, T[i], h->max_seq_len);
        const int n = (T[i] + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
        KOCR_CHECK(nc + n <= h->max_chunks, "model call: batch needs more than %d x 32 token rows", h->max_chunks);
        memset(&lines[i], 0, sizeof(LineDesc));
        lines[i].first_chunk = nc; lines[i].n_chunks = n;
        for (int k = 0; k < n; ++k) chunk_line[nc + k] = i;
        for (int r = 0; r < n * TOK_PER_CHUNK; ++r) row_pos[(size_t)nc * TOK_PER_CHUNK + r] = std::min(r, h->max_seq_len - 1);
        tok_off[i] = nc * TOK_PER_CHUNK; lineT[i] = T[i];
        h->line_T[i] = T[i]; h->line_first_chunk[i] = nc; h->line_n_chunks[i] = n;
        h->max_T = std::max(h->max_T, T[i]);
        nc += n;
    }
    h->n_chunks = nc; h->n_tok = nc * TOK_PER_CHUNK;
    std::vector<int> order(n_lines);
    for (int i = 0; i < n_lines; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->line_T[a] > h->line_T[b]; });
    h->n_groups = (n_lines + 7) / 8;
    for (int g = 0; g < h->n_groups; ++g)
        for (int j = 0; j < 8; ++j) groups[g].line[j] = (g * 8 + j < n_lines) ? order[g * 8 + j] : -1;
    LstmGroup16* groups16 = reinterpret_cast<LstmGroup16*>(groups + h->n_groups);
    h->d_groups16 = reinterpret_cast<LstmGroup16*>(dev_of(groups16));
    h->n_groups16 = (n_lines + 15) / 16;
    for (int g = 0; g < h->n_groups16; ++g)
        for (int j = 0; j < 16; ++j) groups16[g].line[j] = (g * 16 + j < n_lines) ? order[g * 16 + j] : -1;
    const size_t used = reinterpret_cast<uint8_t*>(groups16 + h->n_groups16) - h->staging_host;
    KOCR_CHECK(used <= h->staging_bytes, "internal: staging overflow");
    KOCR_CUDA(cudaMemcpyAsync(h->staging_dev, h->staging_host, used, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaEventRecord(h->staging_done, s));
    return 0;
}

// host fp32 -> the library's 16-bit format (round to nearest even, saturating like pack_a16)
inline act16_t host_a16(float x) {
#ifdef KOCR_A16_BF16
    return __float2bfloat16_rn(x);
#else
    if (x > 65504.f) x = 65504.f;
    if (x < -65504.f) x = -65504.f;
    return __float2half_rn(x);
#endif
}
inline float host_f32(act16_t x) {
#ifdef KOCR_A16_BF16
    return __bfloat162float(x);
#else
    return __half2float(x);
#endif
}

// forced-token decode of the current batch (memory + cross K/V in place): logits of positions 0 .. L-1 -> host [B][L][ld_out]
int run_forced_decode(kocr_handle* h, const int32_t* tgt, int B, int L, float* logits_out, int n_valid, void* stream, cudaStream_t s) {
    std::vector<int32_t> forced((size_t)B * KOCR_TOKENS_LD, 0);
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < L; ++t) forced[(size_t)b * KOCR_TOKENS_LD + t] = tgt[(size_t)b * L + t];
    KOCR_CUDA(cudaMemcpyAsync(buf<int>(h, "forced"), forced.data(), forced.size() * 4, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(wait_stream(h, s));              // `forced` is a host vector of this call
    const int sv_force = h->force_tokens, sv_trace = h->trace_logits, sv_thr = h->straggler_threshold;
    const bool sv_have = h->have_forced;
    h->force_tokens = 1; h->trace_logits = 1; h->have_forced = true; h->straggler_threshold = 0;
    const int rc = kocr_decode_greedy(h, L, nullptr, nullptr, stream);
    h->force_tokens = sv_force; h->trace_logits = sv_trace; h->have_forced = sv_have; h->straggler_threshold = sv_thr;
    if (rc) return rc;
    if (n_valid == VOCAB_PAD) {
        KOCR_CUDA(cudaMemcpy2DAsync(logits_out, (size_t)L * VOCAB_PAD * 4, h->trace.p, (size_t)DEC_MAX * VOCAB_PAD * 4,
                                    (size_t)L * VOCAB_PAD * 4, B, cudaMemcpyDeviceToHost, s));
        KOCR_CUDA(wait_stream(h, s));
    } else {        // compact rows of `n_valid` logits
        std::vector<float> tmp((size_t)B * L * VOCAB_PAD);
        KOCR_CUDA(cudaMemcpy2DAsync(tmp.data(), (size_t)L * VOCAB_PAD * 4, h->trace.p, (size_t)DEC_MAX * VOCAB_PAD * 4,
                                    (size_t)L * VOCAB_PAD * 4, B, cudaMemcpyDeviceToHost, s));
        KOCR_CUDA(wait_stream(h, s));
        for (size_t r = 0; r < (size_t)B * L; ++r) memcpy(logits_out + r * n_valid, tmp.data() + r * VOCAB_PAD, (size_t)n_valid * 4);
    }
    return 0;
}
}  // namespace

// ===========================================================================================
// C ABI
// ===========================================================================================
extern "C" {

int kocr_abi_version(void) { return KOCR_ABI_VERSION; }
const char* kocr_last_error(void) { return get_error(); }
int64_t kocr_launch_count(void) { return g_launches.load() + gemm_tc_launch_count(); }

int kocr_create(const void* weight_blob, size_t blob_bytes, int device, int max_lines, int max_chunks,
                kocr_handle** out) {
    KOCR_CHECK(out != nullptr && weight_blob != nullptr, "kocr_create: null argument");
    *out = nullptr;
    KOCR_CHECK(max_lines > 0 && max_chunks > 0, "kocr_create: max_lines/max_chunks must be positive");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    KOCR_CHECK(ce == cudaSuccess && ndev > 0, "kocr_create: no CUDA device (%s); this library has no CPU path",
               cudaGetErrorString(ce));
    KOCR_CHECK(device >= 0 && device < ndev, "kocr_create: device %d out of range (%d devices)", device, ndev);
    KOCR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KOCR_CUDA(cudaGetDeviceProperties(&prop, device));
    KOCR_CHECK(prop.major == 10, "kocr_create: device is sm_%d%d; this library is built for sm_100a only", prop.major,
               prop.minor);
    const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(weight_blob);
    KOCR_CHECK(blob_bytes >= sizeof(BlobHeader) && memcmp(hd->magic, "KOCRW001", 8) == 0, "kocr_create: bad blob magic");
    const BlobEntry* ents = reinterpret_cast<const BlobEntry*>(hd + 1);
    KOCR_CHECK(blob_bytes >= sizeof(BlobHeader) + (size_t)hd->n_entries * sizeof(BlobEntry), "kocr_create: truncated blob");

    kocr_handle* h = new kocr_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->max_lines = max_lines;
    h->max_chunks = max_chunks;
    auto fail = [&](int rc) { kocr_destroy(h); return rc; };
    if (cudaMalloc(&h->d_blob, blob_bytes) != cudaSuccess) { set_error("kocr_create: cudaMalloc(%zu) failed", blob_bytes); return fail(1); }
    h->blob_bytes = blob_bytes;
    if (cudaMemcpy(h->d_blob, weight_blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("kocr_create: H2D weight copy failed"); return fail(1); }
    const int32_t* meta = nullptr;
    for (uint32_t i = 0; i < hd->n_entries; ++i) {
        const BlobEntry& e = ents[i];
        if (e.offset + e.nbytes > blob_bytes) { set_error("kocr_create: entry %u out of bounds", i); return fail(2); }
        char nm[49]; memcpy(nm, e.name, 48); nm[48] = 0;
        h->w[nm] = std::make_pair((const void*)(h->d_blob + e.offset), (size_t)e.nbytes);
        if (strcmp(nm, "meta") == 0) meta = reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(weight_blob) + e.offset);
    }
    if (!meta) { set_error("kocr_create: blob lacks 'meta'"); return fail(2); }
    h->variant = meta[0]; h->emb_dim = meta[1]; h->max_seq_len = meta[2]; h->dec_max_len = meta[3]; h->vocab = meta[4];
    if (meta[5] != KOCR_A16_FORMAT) {
        set_error("kocr_create: weight blob packs 16-bit operands as %s but this library was built for %s",
                  meta[5] == 1 ? "fp16" : "bf16", KOCR_A16_FORMAT == 1 ? "fp16" : "bf16");
        return fail(2);
    }
    if (h->emb_dim != D_MODEL || h->dec_max_len != DEC_MAX || h->vocab != VOCAB) {
        set_error("kocr_create: unsupported dims emb=%d dec_max=%d vocab=%d (kernels are specialised for 384/256/124)",
                  h->emb_dim, h->dec_max_len, h->vocab);
        return fail(2);
    }
    int rc = resolve_weights(h);
    if (rc) return fail(rc);
    rc = carve_workspace(h);
    if (rc) return fail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("kocr_create: device synchronisation failed"); return fail(1); }   // weight upload / workspace memset ran on the default stream; ours is non-blocking
    *out = h;
    return 0;
}

int kocr_destroy(kocr_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->d_blob) cudaFree(h->d_blob);
    if (h->ws) cudaFree(h->ws);
    if (h->pixels_dev.p) cudaFree(h->pixels_dev.p);
    if (h->mid_dev.p) cudaFree(h->mid_dev.p);
    if (h->trace.p) cudaFree(h->trace.p);
    if (h->beam_cache.p) cudaFree(h->beam_cache.p);
    if (h->beam_scratch.p) cudaFree(h->beam_scratch.p);
    if (h->crop_page.p) cudaFree(h->crop_page.p);
    if (h->tf_x.p) cudaFree(h->tf_x.p);
    if (h->tf_tab.p) cudaFree(h->tf_tab.p);
    if (h->crop_tab.p) cudaFree(h->crop_tab.p);
    if (h->staging_host) cudaFreeHost(h->staging_host);
    if (h->staging_dev) cudaFree(h->staging_dev);
    if (h->pinned_flag) cudaFreeHost(h->pinned_flag);
    if (h->fin_host) cudaFreeHost(h->fin_host);
    if (h->out_stage) cudaFreeHost(h->out_stage);
    if (h->compact_tab.p) cudaFree(h->compact_tab.p);
    if (h->staging_done) cudaEventDestroy(h->staging_done);
    if (h->sync_event) cudaEventDestroy(h->sync_event);
    for (int i = 0; i < 2; ++i) if (h->flag_event[i]) cudaEventDestroy(h->flag_event[i]);
    for (auto& g : h->dec_graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->order_event) cudaEventDestroy(h->order_event);
    for (auto& p : h->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    delete h;
    return 0;
}

size_t kocr_workspace_bytes(const kocr_handle* h) {
    return h ? h->ws_bytes + h->blob_bytes + h->pixels_dev.bytes + h->mid_dev.bytes + h->trace.bytes + h->staging_bytes : 0;
}

int kocr_model_info(const kocr_handle* h, int* variant, int* emb_dim, int* max_seq_len, int* decode_max_len,
                    int* vocab_size) {
    KOCR_CHECK(h != nullptr, "kocr_model_info: null handle");
    if (variant) *variant = h->variant;
    if (emb_dim) *emb_dim = h->emb_dim;
    if (max_seq_len) *max_seq_len = h->max_seq_len;
    if (decode_max_len) *decode_max_len = h->dec_max_len;
    if (vocab_size) *vocab_size = h->vocab;
    return 0;
}

int kocr_gather_chunks(kocr_handle* h, const uint8_t* pixels, size_t pixel_bytes, int pixels_on_device,
                       const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_lines,
                       int32_t* chunk_counts_out, void* stream) {
    KOCR_CHECK(h != nullptr, "kocr_gather_chunks: null handle");
    KOCR_CHECK(n_lines >= 0 && n_lines <= h->max_lines, "kocr_gather_chunks: %d lines exceed capacity %d", n_lines,
               h->max_lines);
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    KOCR_CUDA(cudaSetDevice(h->device));
    h->n_lines = n_lines; h->n_chunks = 0; h->n_tok = 0; h->max_T = 0; h->n_groups = 0; h->max_new_w = 0;
    h->line_T.assign(n_lines, 0); h->line_first_chunk.assign(n_lines, 0); h->line_n_chunks.assign(n_lines, 0);
    if (n_lines == 0) return 0;
    KOCR_CHECK(pixels && offsets && heights && widths, "kocr_gather_chunks: null argument");
    KOCR_CUDA(cudaEventSynchronize(h->staging_done));      // previous batch's table upload has been consumed

    // carve the staging buffer
    uint8_t* sp = h->staging_host;
    LineDesc* lines = reinterpret_cast<LineDesc*>(sp); sp += (size_t)h->max_lines * sizeof(LineDesc);
    int* chunk_line = reinterpret_cast<int*>(sp); sp += (size_t)h->max_chunks * 4;
    int* row_pos = reinterpret_cast<int*>(sp); sp += (size_t)h->max_chunks * TOK_PER_CHUNK * 4;
    int* tok_off = reinterpret_cast<int*>(sp); sp += (size_t)h->max_lines * 4;
    int* lineT = reinterpret_cast<int*>(sp); sp += (size_t)h->max_lines * 4;
    sp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~uintptr_t(15));
    LstmGroup* groups = reinterpret_cast<LstmGroup*>(sp);
    auto dev_of = [&](void* hp) { return h->staging_dev + (reinterpret_cast<uint8_t*>(hp) - h->staging_host); };
    h->d_lines = reinterpret_cast<LineDesc*>(dev_of(lines));
    h->d_chunk_line = reinterpret_cast<int*>(dev_of(chunk_line));
    h->d_row_pos = reinterpret_cast<int*>(dev_of(row_pos));
    h->d_line_tok_off = reinterpret_cast<int*>(dev_of(tok_off));
    h->d_line_T = reinterpret_cast<int*>(dev_of(lineT));
    h->d_groups = reinterpret_cast<LstmGroup*>(dev_of(groups));

    const int max_line_chunks = (h->max_seq_len + TOK_PER_CHUNK - 1) / TOK_PER_CHUNK;
    const int kmax = preprocess_kmax();
    long long mid_total = 0;
    int nc = 0;
    for (int i = 0; i < n_lines; ++i) {
        const int hh = heights[i], ww = widths[i];
        KOCR_CHECK(hh > 0 && ww > 0, "kocr_gather_chunks: line %d has empty size %dx%d", i, hh, ww);
        KOCR_CHECK(offsets[i] >= 0 && (size_t)offsets[i] + (size_t)hh * ww <= pixel_bytes,
                   "kocr_gather_chunks: line %d exceeds the pixel buffer", i);
        // preprocessor.py:45-47: aspect = w / h (float division); new_w = max(50, int(48 * aspect))
        const double aspect = (double)ww / (double)hh;
        int new_w = (int)((double)IMG_H * aspect);
        if (new_w < CHUNK_W / 2) new_w = CHUNK_W / 2;
        const double sh = std::max((double)ww / new_w, 1.0), sv = std::max((double)hh / IMG_H, 1.0);
        KOCR_CHECK((int)std::ceil(sh) * 2 + 1 <= kmax && (int)std::ceil(sv) * 2 + 1 <= kmax,
                   "kocr_gather_chunks: line %d (%dx%d) needs a down-scale beyond the supported %dx", i, hh, ww,
                   (kmax - 1) / 2);
        int n = (new_w + CHUNK_STRIDE - 1) / CHUNK_STRIDE;          // preprocessor.py:21-31
        if (n > max_line_chunks) n = max_line_chunks;               // tokens beyond max_seq_len are dropped (predictor.py:181-183)
        KOCR_CHECK(nc + n <= h->max_chunks, "kocr_gather_chunks: batch needs more than %d chunks", h->max_chunks);
        LineDesc& L = lines[i];
        L.src_off = offsets[i]; L.mid_off = mid_total; L.h = hh; L.w = ww; L.new_w = new_w;
        L.first_chunk = nc; L.n_chunks = n; L.pad_ = 0;
        mid_total += (long long)hh * new_w;
        const int T = std::min(n * TOK_PER_CHUNK, h->max_seq_len);
        for (int k = 0; k < n; ++k) chunk_line[nc + k] = i;
        for (int r = 0; r < n * TOK_PER_CHUNK; ++r) row_pos[(size_t)nc * TOK_PER_CHUNK + r] = std::min(r, h->max_seq_len - 1);
        tok_off[i] = nc * TOK_PER_CHUNK; lineT[i] = T;
        h->line_T[i] = T; h->line_first_chunk[i] = nc; h->line_n_chunks[i] = n;
        h->max_T = std::max(h->max_T, T);
        h->max_new_w = std::max(h->max_new_w, new_w);
        if (chunk_counts_out) chunk_counts_out[i] = n;
        nc += n;
    }
    h->n_chunks = nc; h->n_tok = nc * TOK_PER_CHUNK;
    // LSTM groups: lines sorted by length (descending) so that a cluster's 8 lines finish together
    std::vector<int> order(n_lines);
    for (int i = 0; i < n_lines; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->line_T[a] > h->line_T[b]; });
    h->n_groups = (n_lines + 7) / 8;
    for (int g = 0; g < h->n_groups; ++g)
        for (int j = 0; j < 8; ++j) groups[g].line[j] = (g * 8 + j < n_lines) ? order[g * 8 + j] : -1;
    LstmGroup16* groups16 = reinterpret_cast<LstmGroup16*>(groups + h->n_groups);
    h->d_groups16 = reinterpret_cast<LstmGroup16*>(dev_of(groups16));
    h->n_groups16 = (n_lines + 15) / 16;
    for (int g = 0; g < h->n_groups16; ++g)
        for (int j = 0; j < 16; ++j) groups16[g].line[j] = (g * 16 + j < n_lines) ? order[g * 16 + j] : -1;
    const size_t used = reinterpret_cast<uint8_t*>(groups16 + h->n_groups16) - h->staging_host;
    KOCR_CHECK(used <= h->staging_bytes, "internal: staging overflow");
    KOCR_CUDA(cudaMemcpyAsync(h->staging_dev, h->staging_host, used, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaEventRecord(h->staging_done, s));

    const uint8_t* d_pix = pixels;
    if (pixels_on_device) KOCR_TRY(order_after_default_stream(h, s));
    if (!pixels_on_device) {
        KOCR_TRY(ensure(h->pixels_dev, pixel_bytes));
        KOCR_CUDA(cudaMemcpyAsync(h->pixels_dev.p, pixels, pixel_bytes, cudaMemcpyHostToDevice, s));
        d_pix = reinterpret_cast<const uint8_t*>(h->pixels_dev.p);
    }
    KOCR_TRY(ensure(h->mid_dev, (size_t)mid_total));
    KOCR_TRY(launch_preprocess(d_pix, reinterpret_cast<uint8_t*>(h->mid_dev.p), h->d_lines, h->d_chunk_line,
                               buf<int>(h, "vtab"), buf<float>(h, "chunks"), n_lines, nc, h->max_new_w, s));
    g_launches += 3;
    return 0;
}

int kocr_sevgg_encoder_forward(kocr_handle* h, void* stream) {
    KOCR_CHECK(h != nullptr, "kocr_sevgg_encoder_forward: null handle");
    KOCR_CUDA(cudaSetDevice(h->device));
    return stage_cnn_encoder(h, stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream);
}

int kocr_merge_bilstm_forward(kocr_handle* h, void* stream) {
    KOCR_CHECK(h != nullptr, "kocr_merge_bilstm_forward: null handle");
    KOCR_CUDA(cudaSetDevice(h->device));
    return stage_memory(h, stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream);
}

int kocr_decode_greedy(kocr_handle* h, int max_steps, int32_t* tokens_out, int32_t* lengths_out, void* stream) {
    KOCR_CHECK(h != nullptr, "kocr_decode_greedy: null handle");
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    const int L = h->n_lines;
    if (L == 0) return 0;
    if (max_steps <= 0 || max_steps > h->dec_max_len) max_steps = h->dec_max_len;
    int* tokens = buf<int>(h, "tokens");
    KOCR_CUDA(cudaMemsetAsync(tokens, 0, (size_t)L * KOCR_TOKENS_LD * 4, s));
    KOCR_CUDA(cudaMemsetAsync(buf<int>(h, "finished"), 0, (size_t)L * 4, s));
    KOCR_CUDA(cudaMemsetAsync(buf<int>(h, "n_active"), 0, (DEC_MAX + 1) * 4, s));
    KOCR_CUDA(cudaMemsetAsync(buf<int>(h, "step_base"), 0, 64, s));
    {   // tokens[:, 0] = <sos> (2), lengths = 1
        std::vector<int32_t> init((size_t)L, 2), ones((size_t)L, 1);
        KOCR_CUDA(cudaMemcpy2DAsync(tokens, KOCR_TOKENS_LD * 4, init.data(), 4, 4, L, cudaMemcpyHostToDevice, s));
        KOCR_CUDA(cudaMemcpyAsync(buf<int>(h, "lengths"), ones.data(), (size_t)L * 4, cudaMemcpyHostToDevice, s));
        KOCR_CUDA(wait_stream(h, s));   // the two host vectors die at scope end
    }
    if (h->trace_logits) {
        KOCR_TRY(ensure(h->trace, (size_t)h->max_lines * DEC_MAX * VOCAB_PAD * 4));
        KOCR_CUDA(cudaMemsetAsync(h->trace.p, 0, (size_t)L * DEC_MAX * VOCAB_PAD * 4, s));
    }
    const bool forcing = h->force_tokens && h->have_forced;
    const int max_T = (h->max_T + 127) / 128 * 128;      // bucketed: only sizes the cross-attention scratch
    // Row compaction (see decode_compact_kernel): plain greedy decoding only - the logits trace and forced tokens are
    // indexed by the caller's line numbers
    const bool may_compact = h->compact_rows && !forcing && !h->trace_logits;
    h->dec_rows = L;
    h->row_orig.resize(L);
    for (int i = 0; i < L; ++i) h->row_orig[i] = i;
    bool compacted = false;
    int done = 0;
    // One group = DEC_GROUP positions (a CUDA-graph replay) followed by a 4-byte copy of "lines still active" to the host.
    auto enqueue_group = [&](int slot) -> int {
        const int n = std::min(DEC_GROUP, max_steps - done);
        // the first group of a handle runs eagerly (it sets the kernels' function attributes)
        const auto t0 = std::chrono::steady_clock::now();
        if (n == DEC_GROUP && h->use_graphs && h->decode_warmed) KOCR_TRY(decode_group_graph(h, max_T, s));
        else KOCR_TRY(decode_group_eager(h, n, max_T, s));
        h->host_launch_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        h->decode_warmed = true;
        done += n;
        if (!forcing && done < max_steps) {
            KOCR_CUDA(cudaMemcpyAsync(h->pinned_flag + slot, buf<int>(h, "n_active") + (done - 1), 4, cudaMemcpyDeviceToHost, s));
            KOCR_CUDA(cudaEventRecord(h->flag_event[slot], s));
        }
        return 0;
    };
    auto wait_flag = [&](int slot) -> cudaError_t {
        const auto t1 = std::chrono::steady_clock::now();
        cudaError_t e = cudaSuccess;
        if (h->blocking_wait) e = cudaEventSynchronize(h->flag_event[slot]);
        else while ((e = cudaEventQuery(h->flag_event[slot])) == cudaErrorNotReady) {}
        h->host_wait_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count();
        return e;
    };
    if (!may_compact && h->dec_lookahead) {
        // Look-ahead (option, off by default - measured slightly slower with 12-18 passes in flight: the speculative group's GEMMs
        // still run on every row, and the decode phase is bound by total work, not by the host round trip): group g + 1 is enqueued BEFORE the host learns how group g ended, so the stream never idles on the
        // host round trip (a blocking-sync wake-up is 0.1-0.4 ms, x 13 groups per batch).  A group that starts after every
        // line has finished costs little: all per-line kernels exit at once.  Tokens are unaffected (a finished line
        // stays finished; stragglers are re-decoded from scratch by the caller).
        int slot = 0;
        KOCR_TRY(enqueue_group(slot));
        while (!forcing && done < max_steps) {
            const int polled_slot = slot;               // the group whose outcome we are about to read
            slot ^= 1;
            KOCR_TRY(enqueue_group(slot));              // ... after the next one is already in the stream
            KOCR_CUDA(wait_flag(polled_slot));
            if (h->pinned_flag[polled_slot] <= h->straggler_threshold) break;   // every line (or all but a few stragglers) has emitted <eos>
        }
        while (forcing && done < max_steps) KOCR_TRY(enqueue_group(0));
    } else
    while (done < max_steps) {
        KOCR_TRY(enqueue_group(0));
        if (!forcing && done < max_steps) {
            KOCR_CUDA(wait_flag(0));
            const int n_active = h->pinned_flag[0];
            if (n_active <= h->straggler_threshold) break;     // every line (or all but a few stragglers) has emitted <eos>
            const int target = (n_active + 127) / 128 * 128;
            if (may_compact && target < h->dec_rows) {
                const int R = h->dec_rows;
                int32_t* fin = h->out_stage;                    // pinned scratch
                KOCR_CUDA(cudaMemcpyAsync(fin, buf<int>(h, "finished"), (size_t)R * 4, cudaMemcpyDeviceToHost, s));
                KOCR_CUDA(wait_stream(h, s));
                std::vector<int32_t> pairs;
                int hole = 0;
                for (int src = target; src < R; ++src) {
                    if (fin[src]) continue;                     // finished line in the tail: stays where it is
                    while (hole < target && !fin[hole]) ++hole; // next finished row of the head
                    KOCR_CHECK(hole < target, "internal: decode compaction found no hole for row %d", src);
                    pairs.push_back(src); pairs.push_back(hole);
                    std::swap(h->row_orig[src], h->row_orig[hole]);
                    ++hole;
                }
                if (!pairs.empty()) {
                    KOCR_TRY(ensure(h->compact_tab, pairs.size() * 4));
                    KOCR_CUDA(cudaMemcpyAsync(h->compact_tab.p, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, s));
                    KOCR_TRY(launch_decode_compact(reinterpret_cast<const int*>(h->compact_tab.p), (int)pairs.size() / 2, done,
                                                   buf<float>(h, "kcache"), buf<float>(h, "vcache"),
                                                   (size_t)h->max_lines * DEC_MAX * D_MODEL, tokens, buf<int>(h, "lengths"),
                                                   buf<int>(h, "finished"), h->d_line_tok_off, h->d_line_T, s));
                    ++g_launches;
                    KOCR_CUDA(wait_stream(h, s));               // `pairs` is a host vector of this scope
                    compacted = true;
                }
                h->dec_rows = target;
            }
        }
    }
    h->last_steps = done; h->last_max_steps = max_steps;
    // results come back in row order; un-permute them into the caller's line order
    int32_t* st_tok = h->out_stage;
    int32_t* st_len = st_tok + (size_t)h->max_lines * KOCR_TOKENS_LD;
    int32_t* st_fin = st_len + h->max_lines;
    if (tokens_out) KOCR_CUDA(cudaMemcpyAsync(st_tok, tokens, (size_t)L * KOCR_TOKENS_LD * 4, cudaMemcpyDeviceToHost, s));
    if (lengths_out) KOCR_CUDA(cudaMemcpyAsync(st_len, buf<int>(h, "lengths"), (size_t)L * 4, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(cudaMemcpyAsync(st_fin, buf<int>(h, "finished"), (size_t)L * 4, cudaMemcpyDeviceToHost, s));
    if (compacted) {    // put the per-line tables back for whoever uses this batch next (beam search, another decode)
        const uint8_t* host_off = h->staging_host + (reinterpret_cast<uint8_t*>(h->d_line_tok_off) - h->staging_dev);
        const uint8_t* host_T = h->staging_host + (reinterpret_cast<uint8_t*>(h->d_line_T) - h->staging_dev);
        KOCR_CUDA(cudaMemcpyAsync(h->d_line_tok_off, host_off, (size_t)L * 4, cudaMemcpyHostToDevice, s));
        KOCR_CUDA(cudaMemcpyAsync(h->d_line_T, host_T, (size_t)L * 4, cudaMemcpyHostToDevice, s));
    }
    KOCR_CUDA(wait_stream(h, s));
    for (int r = 0; r < L; ++r) {
        const int line = h->row_orig[r];
        if (tokens_out) memcpy(tokens_out + (size_t)line * KOCR_TOKENS_LD, st_tok + (size_t)r * KOCR_TOKENS_LD, KOCR_TOKENS_LD * 4);
        if (lengths_out) lengths_out[line] = st_len[r];
        h->fin_host[line] = st_fin[r];
    }
    h->dec_rows = L;
    return 0;
}

int kocr_recognize_lines(kocr_handle* h, const uint8_t* pixels, size_t pixel_bytes, int pixels_on_device,
                         const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_lines,
                         int max_steps, int32_t* tokens_out, int32_t* lengths_out, void* stream) {
    KOCR_TRY(kocr_gather_chunks(h, pixels, pixel_bytes, pixels_on_device, offsets, heights, widths, n_lines, nullptr,
                                stream));
    KOCR_TRY(kocr_sevgg_encoder_forward(h, stream));
    KOCR_TRY(kocr_merge_bilstm_forward(h, stream));
    return kocr_decode_greedy(h, max_steps, tokens_out, lengths_out, stream);
}

int kocr_set_option(kocr_handle* h, const char* name, int value) {
    KOCR_CHECK(h != nullptr && name != nullptr, "kocr_set_option: null argument");
    if (strcmp(name, "trace_logits") == 0) { h->trace_logits = value; return 0; }
    if (strcmp(name, "force_tokens") == 0) { h->force_tokens = value; return 0; }
    if (strcmp(name, "use_graphs") == 0) { h->use_graphs = value; return 0; }
    if (strcmp(name, "use_pdl") == 0) { h->use_pdl = value; return 0; }
    if (strcmp(name, "blocking_wait") == 0) { h->blocking_wait = value; return 0; }
    if (strcmp(name, "kv_split") == 0) { h->kv_split = value; return 0; }
    if (strcmp(name, "lstm_split") == 0) { h->lstm_split = value; return 0; }
    if (strcmp(name, "compact_rows") == 0) { h->compact_rows = value; return 0; }
    if (strcmp(name, "dec_cross_impl") == 0) { set_dec_cross_attention_impl(value); return 0; }          // process-wide
    if (strcmp(name, "se_variant") == 0) { set_se_excite_variant(value); return 0; }                      // process-wide
    if (strcmp(name, "gemm_bn192") == 0) { set_gemm_bn192(value); return 0; }                             // process-wide
    if (strcmp(name, "chunk_attn_impl") == 0) { set_chunk_attention_impl(value); return 0; }   // process-wide
    if (strcmp(name, "debug_stop") == 0) { h->debug_stop = value; return 0; }
    if (strcmp(name, "dec_skip") == 0) { h->dec_skip = value; return 0; }
    if (strcmp(name, "pool2_fused") == 0) { h->pool2_fused = value; return 0; }
    if (strcmp(name, "dec_fused") == 0) { h->dec_fused = value; return 0; }
    if (strcmp(name, "dec_lookahead") == 0) { h->dec_lookahead = value; return 0; }
    if (strcmp(name, "dec_wide") == 0) { h->dec_wide = value; return 0; }
    if (strcmp(name, "lstm_impl") == 0) { h->lstm_impl = value; return 0; }
    if (strcmp(name, "big_gemm_sms") == 0) { h->big_gemm_sms = value; return 0; }
    if (strcmp(name, "straggler_threshold") == 0) { h->straggler_threshold = value; return 0; }
    if (strcmp(name, "kernel_timing") == 0) {
        h->kernel_timing = value;
        if (value) { h->sites.clear(); }
        return 0;
    }
    KOCR_CHECK(false, "kocr_set_option: unknown option '%s'", name);
    return 0;
}

int kocr_set_forced_tokens(kocr_handle* h, const int32_t* tokens, int n_lines) {
    KOCR_CHECK(h != nullptr && tokens != nullptr, "kocr_set_forced_tokens: null argument");
    KOCR_CHECK(n_lines > 0 && n_lines <= h->max_lines, "kocr_set_forced_tokens: bad n_lines %d", n_lines);
    KOCR_CUDA(cudaSetDevice(h->device));
    KOCR_CUDA(cudaMemcpy(buf<int>(h, "forced"), tokens, (size_t)n_lines * KOCR_TOKENS_LD * 4, cudaMemcpyHostToDevice));
    h->have_forced = true;
    return 0;
}

// ---- beam search support (OCRPredictor._beam_search, predictor.py:101-136) ---------------------------------
// One decoder position for `n_rows` hypotheses of line `line`.  The host owns the beam bookkeeping (scores, pruning,
// tie order - identical to the reference's Python); the device keeps one self-attention cache row per hypothesis.
namespace {
__global__ void beam_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst, const int* __restrict__ parents,
                                    int n_rows, int t, size_t layer_stride, size_t kv_stride) {
    // grid = (n_rows, 2 layers, 2 {K,V}); copies positions [0, t) of the parent's cache row
    const int r = blockIdx.x, layer = blockIdx.y, kv = blockIdx.z;
    if (parents[r] < 0) return;                      // dead row (kocr_beam_search)
    const float4* s4 = reinterpret_cast<const float4*>(src + kv * kv_stride + layer * layer_stride + (size_t)parents[r] * DEC_MAX * D_MODEL);
    float4* d4 = reinterpret_cast<float4*>(dst + kv * kv_stride + layer * layer_stride + (size_t)r * DEC_MAX * D_MODEL);
    for (int i = threadIdx.x; i < t * (D_MODEL / 4); i += blockDim.x) d4[i] = s4[i];
}
}  // namespace

int kocr_beam_step_batch(kocr_handle* h, int n_rows, const int32_t* row_line, const int32_t* parents,
                         const int32_t* prefixes, int t, float* logits_out, void* stream) {
    KOCR_CHECK(h != nullptr && prefixes != nullptr && logits_out != nullptr && row_line != nullptr, "kocr_beam_step_batch: null argument");
    const int R = h->max_lines;                          // every decode buffer of the handle has max_lines rows
    KOCR_CHECK(n_rows >= 1 && n_rows <= R, "kocr_beam_step_batch: %d hypotheses (handle capacity %d rows)", n_rows, R);
    KOCR_CHECK(t >= 0 && t < h->dec_max_len, "kocr_beam_step_batch: position %d out of range", t);
    KOCR_CHECK(t == 0 || parents != nullptr, "kocr_beam_step_batch: parents required for t > 0");
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    const size_t row = (size_t)DEC_MAX * D_MODEL, layer_stride = (size_t)R * row, kv_stride = 2 * layer_stride,
                 buf_stride = 2 * kv_stride;
    KOCR_TRY(ensure(h->beam_cache, 2 * buf_stride * sizeof(float)));
    KOCR_TRY(ensure(h->trace, (size_t)h->max_lines * DEC_MAX * VOCAB_PAD * 4));
    float* base = reinterpret_cast<float*>(h->beam_cache.p);
    int* tokens = buf<int>(h, "tokens");
    int* scratch = buf<int>(h, "forced");             // [0, R): parents, [R, 2R): tok_off, [2R, 3R): T  (buffer: R x 257 ints)
    h->have_forced = false;                           // ... which overwrites any tokens set with kocr_set_forced_tokens
    std::vector<int32_t> tab((size_t)3 * R, 0);
    int max_T = 0;
    for (int r = 0; r < n_rows; ++r) {
        const int line = row_line[r];
        KOCR_CHECK(line >= 0 && line < h->n_lines, "kocr_beam_step_batch: row %d refers to line %d outside the current batch of %d", r, line, h->n_lines);
        tab[r] = t > 0 ? parents[r] : 0;
        KOCR_CHECK(tab[r] >= 0 && tab[r] < R, "kocr_beam_step_batch: bad parent index %d", tab[r]);
        tab[(size_t)R + r] = h->line_first_chunk[line] * TOK_PER_CHUNK;
        tab[(size_t)2 * R + r] = h->line_T[line];
        max_T = std::max(max_T, h->line_T[line]);
    }
    KOCR_CUDA(cudaMemcpyAsync(scratch, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
    // token prefixes of all hypotheses (positions 0..t): row r of `prefixes` is [t + 1] ints
    KOCR_CUDA(cudaMemcpy2DAsync(tokens, KOCR_TOKENS_LD * 4, prefixes, (size_t)(t + 1) * 4, (size_t)(t + 1) * 4, n_rows,
                                cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaMemsetAsync(buf<int>(h, "finished"), 0, (size_t)R * 4, s));
    const int32_t step = t;
    KOCR_CUDA(cudaMemcpyAsync(buf<int>(h, "step_base"), &step, 4, cudaMemcpyHostToDevice, s));
    if (t == 0) h->beam_cur = 0;
    if (t > 0) {
        const int nxt = h->beam_cur ^ 1;
        beam_reorder_kernel<<<dim3(n_rows, 2, 2), 256, 0, s>>>(base + h->beam_cur * buf_stride, base + nxt * buf_stride, scratch,
                                                              n_rows, t, layer_stride, kv_stride);
        KOCR_CUDA(cudaGetLastError());
        ++g_launches;
        h->beam_cur = nxt;
    }
    DecRows rows;
    rows.n_rows = n_rows;
    rows.kcache = base + h->beam_cur * buf_stride;
    rows.vcache = rows.kcache + kv_stride;
    rows.layer_stride = layer_stride;
    rows.tok_off = scratch + R;
    rows.T = scratch + 2 * R;
    const int saved_force = h->force_tokens;
    h->force_tokens = 0;
    int rc = decode_step(h, 0, (max_T + 127) / 128 * 128, s, &rows);
    h->force_tokens = saved_force;
    if (rc) return rc;
    // the argmax kernel wrote the (bias-added, slice-summed) logits of position t into the trace rows
    KOCR_CUDA(cudaMemcpy2DAsync(logits_out, VOCAB_PAD * 4, reinterpret_cast<float*>(h->trace.p) + (size_t)t * VOCAB_PAD,
                                (size_t)DEC_MAX * VOCAB_PAD * 4, VOCAB_PAD * 4, n_rows, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(wait_stream(h, s));     // tab / prefixes are host memory of this call
    return 0;
}

// ---- whole beam search in one call (OCRPredictor._beam_search, predictor.py:101-136, for every line of the batch) -------
// The per-position host round trip of kocr_beam_step_batch (Python bookkeeping + torch log-softmax / top-k on the logits)
// bounded beam search at ~0.6 k lines/s.  Here the loop lives in the library: hypotheses sit in fixed row slots (line l owns
// rows [l * bw, (l + 1) * bw)), a device kernel turns the logits into log-softmax top-`bw` (value, index) pairs per row, the
// host keeps the reference's bookkeeping in C++ (float64 score sums of the fp32 log-probabilities, hypothesis-major /
// top-k-minor candidate order, STABLE descending sort, every <eos> candidate completed with score / len(seq), the first `bw`
// others survive, best = first-appended among equal completed scores, else the first live hypothesis), and the next
// position's prefixes / self-attention caches are re-ordered on the device from the parent indices.
namespace {
__global__ void __launch_bounds__(128) beam_topk_kernel(const float* __restrict__ trace, int t, int bw, int n_rows,
                                                        float* __restrict__ vals, int* __restrict__ idxs) {
    // warp per row: log_softmax over the 124 logits of position t (fp32, max-subtracted like torch), then `bw` rounds of
    // warp arg-max (ties -> lowest index)
    const int r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const float4 v4 = reinterpret_cast<const float4*>(trace + ((long)r * DEC_MAX + t) * VOCAB_PAD)[lane];
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (lane * 4 + j < VOCAB) mx = fmaxf(mx, v[j]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (lane * 4 + j < VOCAB) sum += expf(v[j] - mx);
    sum = warp_sum(sum);
    const float lse = logf(sum);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (lane * 4 + j < VOCAB) ? (v[j] - mx) - lse : -INFINITY;
    for (int k = 0; k < bw; ++k) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (v[j] > best) { best = v[j]; bi = lane * 4 + j; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { vals[r * 8 + k] = best; idxs[r * 8 + k] = bi; }
        if ((bi >> 2) == lane) v[bi & 3] = -INFINITY;           // remove the winner
    }
}

// next position's row state from the parent indices: prefix tokens [0, t] of the parent + the chosen token at t + 1, and the
// "dead row" flag the per-line kernels skip on.  grid = n_rows.
__global__ void __launch_bounds__(64) beam_rows_kernel(const int* __restrict__ tok_old, int* __restrict__ tok_new,
                                                       const int* __restrict__ parents, const int* __restrict__ new_tok,
                                                       int* __restrict__ finished, int t) {
    const int r = blockIdx.x, p = parents[r];
    if (threadIdx.x == 0) finished[r] = p < 0 ? 1 : 0;
    if (p < 0) return;
    for (int i = threadIdx.x; i <= t; i += blockDim.x) tok_new[r * KOCR_TOKENS_LD + i] = tok_old[p * KOCR_TOKENS_LD + i];
    if (threadIdx.x == 0) tok_new[r * KOCR_TOKENS_LD + t + 1] = new_tok[r];
}
}  // namespace

int kocr_beam_search(kocr_handle* h, int beam_width, int max_len, int32_t* tokens_out, int32_t* lengths_out, void* stream) {
    KOCR_CHECK(h != nullptr && tokens_out != nullptr && lengths_out != nullptr, "kocr_beam_search: null argument");
    const int bw = beam_width, n_lines = h->n_lines, R = n_lines * bw;
    KOCR_CHECK(bw >= 1 && bw <= kocr_handle::BEAM_MAX, "kocr_beam_search: beam width %d outside [1, %d]", bw, kocr_handle::BEAM_MAX);
    KOCR_CHECK(n_lines > 0 && h->n_tok > 0, "kocr_beam_search: run kocr_gather_chunks + kocr_sevgg_encoder_forward + kocr_merge_bilstm_forward on a batch first");
    KOCR_CHECK(R <= h->max_lines, "kocr_beam_search: %d lines x beam %d exceed the handle's %d decode rows", n_lines, bw, h->max_lines);
    if (max_len <= 0 || max_len > h->dec_max_len) max_len = h->dec_max_len;
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    const int ML = h->max_lines;
    const size_t row = (size_t)DEC_MAX * D_MODEL, layer_stride = (size_t)ML * row, kv_stride = 2 * layer_stride, buf_stride = 2 * kv_stride;
    KOCR_TRY(ensure(h->beam_cache, 2 * buf_stride * sizeof(float)));
    KOCR_TRY(ensure(h->trace, (size_t)ML * DEC_MAX * VOCAB_PAD * 4));
    // device scratch: [0,R) parents | [R,2R) new tokens | [2R,3R) tok_off | [3R,4R) T | top-k vals [R*8] | top-k idx [R*8] | token table B
    KOCR_TRY(ensure(h->beam_scratch, (size_t)(4 * ML + 16 * ML) * 4 + (size_t)ML * KOCR_TOKENS_LD * 4));
    int* d_par = reinterpret_cast<int*>(h->beam_scratch.p);
    int* d_newtok = d_par + ML;
    int* d_tokoff = d_par + 2 * ML;
    int* d_T = d_par + 3 * ML;
    float* d_vals = reinterpret_cast<float*>(d_par + 4 * ML);
    int* d_idx = d_par + 12 * ML;
    int* d_tokB = d_par + 20 * ML;
    int* tokens = buf<int>(h, "tokens");
    float* base = reinterpret_cast<float*>(h->beam_cache.p);
    h->have_forced = false;
    const int sv_force = h->force_tokens, sv_fused = h->dec_fused;
    h->force_tokens = 0; h->dec_fused = 0;        // the beam step wants the plain logits of every live row (trace), no greedy bookkeeping
    struct Restore { kocr_handle* h; int f, d; ~Restore() { h->force_tokens = f; h->dec_fused = d; } } restore{h, sv_force, sv_fused};

    std::vector<int32_t> tab((size_t)4 * ML, 0);
    int max_T = 0;
    for (int l = 0; l < n_lines; ++l)
        for (int i = 0; i < bw; ++i) {
            const int r = l * bw + i;
            tab[r] = i == 0 ? 0 : -1;                      // position 0: one live hypothesis per line (<sos>)
            tab[(size_t)2 * ML + r] = h->line_first_chunk[l] * TOK_PER_CHUNK;
            tab[(size_t)3 * ML + r] = h->line_T[l];
            max_T = std::max(max_T, h->line_T[l]);
        }
    max_T = (max_T + 127) / 128 * 128;
    KOCR_CUDA(cudaMemcpyAsync(d_par, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
    {   // token table: <sos> at position 0 of every row; dead rows flagged
        std::vector<int32_t> init((size_t)R * KOCR_TOKENS_LD, 0);
        for (int r = 0; r < R; ++r) init[(size_t)r * KOCR_TOKENS_LD] = 2;
        KOCR_CUDA(cudaMemcpyAsync(tokens, init.data(), init.size() * 4, cudaMemcpyHostToDevice, s));
        std::vector<int32_t> fin((size_t)R);
        for (int r = 0; r < R; ++r) fin[r] = (r % bw) == 0 ? 0 : 1;
        KOCR_CUDA(cudaMemcpyAsync(buf<int>(h, "finished"), fin.data(), fin.size() * 4, cudaMemcpyHostToDevice, s));
        KOCR_CUDA(cudaMemsetAsync(buf<int>(h, "step_base"), 0, 64, s));
        KOCR_CUDA(wait_stream(h, s));
    }
    // host state (reference: `beams` = list of (score, seq), `completed`)
    struct Hyp { double score; std::vector<int32_t> seq; };
    std::vector<std::vector<Hyp>> beams(n_lines);
    std::vector<double> best_score(n_lines, -INFINITY);
    std::vector<std::vector<int32_t>> best_seq(n_lines);
    for (int l = 0; l < n_lines; ++l) beams[l].push_back(Hyp{0.0, std::vector<int32_t>{2}});
    std::vector<float> hv((size_t)R * 8);
    std::vector<int32_t> hi((size_t)R * 8), par((size_t)2 * ML);
    int cur = 0, steps_run = 0;
    DecRows rows;
    rows.n_rows = R; rows.layer_stride = layer_stride; rows.tok_off = d_tokoff; rows.T = d_T;
    struct Cand { double s; int hyp; int tok; };
    std::vector<Cand> cand;
    for (int t = 0; t < max_len; ++t) {
        rows.kcache = base + cur * buf_stride;
        rows.vcache = rows.kcache + kv_stride;
        KOCR_TRY(decode_step(h, 0, max_T, s, &rows));
        beam_topk_kernel<<<(R + 3) / 4, 128, 0, s>>>(reinterpret_cast<const float*>(h->trace.p), t, bw, R, d_vals, d_idx);
        KOCR_CUDA(cudaGetLastError());
        ++g_launches;
        KOCR_CUDA(cudaMemcpyAsync(hv.data(), d_vals, (size_t)R * 8 * 4, cudaMemcpyDeviceToHost, s));
        KOCR_CUDA(cudaMemcpyAsync(hi.data(), d_idx, (size_t)R * 8 * 4, cudaMemcpyDeviceToHost, s));
        KOCR_CUDA(wait_stream(h, s));
        bool any_live = false;
        for (int l = 0; l < n_lines; ++l) {
            std::vector<Hyp>& B = beams[l];
            for (int i = 0; i < bw; ++i) { par[l * bw + i] = -1; par[(size_t)ML + l * bw + i] = 0; }
            if (B.empty()) continue;
            cand.clear();
            for (int i = 0; i < (int)B.size(); ++i)
                for (int k = 0; k < bw; ++k)
                    cand.push_back(Cand{B[i].score + (double)hv[(size_t)(l * bw + i) * 8 + k], i, hi[(size_t)(l * bw + i) * 8 + k]});
            std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.s > b.s; });   // list.sort(reverse=True)
            std::vector<Hyp> next;
            for (const Cand& c : cand) {
                if (c.tok == 3) {                                   // <eos>: completed with score / len(seq)
                    const double sc = c.s / (double)(B[c.hyp].seq.size() + 1);
                    if (sc > best_score[l]) {                       // sorted(completed, reverse=True)[0]: first among equals
                        best_score[l] = sc;
                        best_seq[l] = B[c.hyp].seq;
                        best_seq[l].push_back(3);
                    }
                } else if ((int)next.size() < bw) {
                    par[l * bw + (int)next.size()] = l * bw + c.hyp;
                    par[(size_t)ML + l * bw + (int)next.size()] = c.tok;
                    Hyp nh{c.s, B[c.hyp].seq};
                    nh.seq.push_back(c.tok);
                    next.push_back(std::move(nh));
                }
            }
            B.swap(next);
            // Exact early stop.  The reference never leaves its loop before decode_max_len positions (every live hypothesis
            // contributes non-<eos> candidates, so `beams` is never empty) - it keeps extending low-probability alternatives
            // long after the answer is known.  Log-probabilities are <= 0, so a hypothesis with score s can at best complete
            // with a normalised score of s / (max_len + 1) (its sum can only fall, its length is at most max_len + 1); once
            // that bound is below the best completed score for EVERY live hypothesis of the line, nothing the reference
            // would still compute can change its answer (ties go to the earlier completion), and the line stops here.
            if (best_score[l] > -INFINITY) {
                bool can_win = false;
                for (const Hyp& hy : B) can_win = can_win || (hy.score / (double)(max_len + 1) >= best_score[l]);
                if (!can_win) B.clear();
            }
            any_live = any_live || !B.empty();
        }
        steps_run = t + 1;
        int live_lines = 0;
        for (int l = 0; l < n_lines; ++l) live_lines += beams[l].empty() ? 0 : 1;
        // long tail (option "straggler_threshold", like kocr_decode_greedy): alternatives that never emit <eos> ramble on to
        // max_len and hold the whole pass; stop once only a few lines still have live hypotheses - the caller re-submits
        // those lines (kocr_read_unfinished; the search is deterministic)
        if (!any_live || live_lines <= h->straggler_threshold || t + 1 >= max_len) break;
        // device state of position t + 1: caches and prefixes follow their parents
        KOCR_CUDA(cudaMemcpyAsync(d_par, par.data(), (size_t)2 * ML * 4, cudaMemcpyHostToDevice, s));
        beam_rows_kernel<<<R, 64, 0, s>>>(tokens, d_tokB, d_par, d_newtok, buf<int>(h, "finished"), t);
        KOCR_CUDA(cudaGetLastError());
        KOCR_CUDA(cudaMemcpyAsync(tokens, d_tokB, (size_t)R * KOCR_TOKENS_LD * 4, cudaMemcpyDeviceToDevice, s));
        const int nxt = cur ^ 1;
        // (dead rows carry parent -1: clamp to their own row, the copy is harmless)
        beam_reorder_kernel<<<dim3(R, 2, 2), 256, 0, s>>>(base + cur * buf_stride, base + nxt * buf_stride, d_par, R, t + 1, layer_stride, kv_stride);
        KOCR_CUDA(cudaGetLastError());
        g_launches += 2;
        cur = nxt;
        KOCR_TRY(launch_dec_bump(buf<int>(h, "step_base"), 1, s)); ++g_launches;
    }
    for (int l = 0; l < n_lines; ++l) {
        // best completed hypothesis, else the first live one (predictor.py:135); `beams` is never empty in that case: a line
        // without live beams has completed at least one
        const std::vector<int32_t>& seq = !best_seq[l].empty() ? best_seq[l] : beams[l].front().seq;
        const int n = std::min((int)seq.size(), KOCR_TOKENS_LD);
        memset(tokens_out + (size_t)l * KOCR_TOKENS_LD, 0, KOCR_TOKENS_LD * 4);
        memcpy(tokens_out + (size_t)l * KOCR_TOKENS_LD, seq.data(), (size_t)n * 4);
        lengths_out[l] = n;
        h->fin_host[l] = beams[l].empty() ? 1 : 0;   // live hypotheses left: unfinished unless the position budget ran out
    }
    h->last_steps = steps_run; h->last_max_steps = max_len;
    return 0;
}

int kocr_beam_step(kocr_handle* h, int line, int n_rows, const int32_t* parents, const int32_t* prefixes, int t,
                   float* logits_out, void* stream) {
    KOCR_CHECK(h != nullptr, "kocr_beam_step: null handle");
    KOCR_CHECK(n_rows >= 1 && n_rows <= kocr_handle::BEAM_MAX, "kocr_beam_step: %d hypotheses (max %d)", n_rows, kocr_handle::BEAM_MAX);
    int32_t lines[kocr_handle::BEAM_MAX];
    for (int r = 0; r < n_rows; ++r) lines[r] = line;
    return kocr_beam_step_batch(h, n_rows, lines, parents, prefixes, t, logits_out, stream);
}

// ---- teacher-forced batched forward (KhmerOCR.forward, se_model.py:240-289) -----------------------------------
int kocr_forward_teacher_forced(kocr_handle* h, const int32_t* tgt_tokens, int L, float* logits_out, void* stream) {
    KOCR_CHECK(h != nullptr && tgt_tokens != nullptr && logits_out != nullptr, "kocr_forward_teacher_forced: null argument");
    const int B = h->n_lines, Tmax = h->max_T;
    KOCR_CHECK(B > 0 && h->n_tok > 0, "kocr_forward_teacher_forced: run kocr_gather_chunks + kocr_sevgg_encoder_forward on a batch first");
    KOCR_CHECK(L >= 1 && L <= h->dec_max_len, "kocr_forward_teacher_forced: target length %d outside [1, %d]", L, h->dec_max_len);
    const long rows = (long)B * Tmax;
    KOCR_CHECK(rows <= (long)h->max_chunks * TOK_PER_CHUNK, "kocr_forward_teacher_forced: %d lines padded to %d tokens exceed the "
               "handle's %ld memory rows (create it with a larger max_chunks)", B, Tmax, (long)h->max_chunks * TOK_PER_CHUNK);
    for (int b = 0; b < B; ++b)
        KOCR_CHECK(tgt_tokens[(size_t)b * L] == 2, "kocr_forward_teacher_forced: target %d does not start with <sos>", b);
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    // tables: padded token offsets (b * Tmax) and the full length Tmax of every line for the BiLSTM
    std::vector<int32_t> tab((size_t)2 * B);
    for (int b = 0; b < B; ++b) { tab[b] = b * Tmax; tab[(size_t)B + b] = Tmax; }
    KOCR_TRY(ensure(h->tf_tab, tab.size() * 4));
    KOCR_TRY(ensure(h->tf_x, (size_t)rows * D_MODEL * 2));
    int* d_off_pad = reinterpret_cast<int*>(h->tf_tab.p);
    int* d_T_full = d_off_pad + B;
    KOCR_CUDA(cudaMemcpyAsync(d_off_pad, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
    act16_t* xpad = reinterpret_cast<act16_t*>(h->tf_x.p);
    KOCR_TRY(launch_pad_memory(buf<act16_t>(h, "xb"), h->global_pos, h->d_line_tok_off, h->d_line_T, B, Tmax, xpad, s)); ++g_launches;
    const act16_t* memb = xpad;                       // VGG baseline: the padded merged sequence is the memory
    if (h->variant == 0) {
        GemmEpilogue e = ep_none();
        e.bias = h->lstm_b; e.out_f32 = buf<float>(h, "gin"); e.ld_f32 = 8 * LSTM_H;
        KOCR_TRY(gemm_linear(h, xpad, rows, h->lstm_w_ih, 8 * LSTM_H, D_MODEL, e, s));
        // no packing (se_model.py:278-279): both directions run over all Tmax rows, the backward one reads the pads first
        KOCR_TRY(launch_bilstm_mma(buf<float>(h, "gin"), h->lstm_w_hh_mma, d_off_pad, d_T_full, h->d_groups16, h->n_groups16,
                                   buf<float>(h, "mem"), buf<act16_t>(h, "memb"), nullptr, s));
        ++g_launches;
        memb = buf<act16_t>(h, "memb");
    }
    if (h->variant == 0) {
        KOCR_TRY(project_cross_kv(h, rows, buf<float>(h, "mem"), memb, s));
    } else {        // the padded memory of the baselines exists in 16 bits only: plain 16-bit projection
        GemmEpilogue e = ep_none();
        e.bias = h->dec_kv_b; e.out_a16 = buf<act16_t>(h, "kv"); e.ld_a16 = 4 * D_MODEL;
        KOCR_TRY(gemm_linear(h, memb, rows, h->dec_kv_w, 4 * D_MODEL, D_MODEL, e, s));
    }
    // from here on the handle's memory lives in the padded layout: point the decoder at it (the real lengths in d_line_T
    // are the memory_key_padding_mask, se_model.py:282-285); the next kocr_gather_chunks rewrites these tables
    KOCR_CUDA(cudaMemcpyAsync(h->d_line_tok_off, d_off_pad, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    // ... and in the host copy of the table, which kocr_decode_greedy restores from after a row compaction
    memcpy(h->staging_host + (reinterpret_cast<uint8_t*>(h->d_line_tok_off) - h->staging_dev), tab.data(), (size_t)B * 4);
    for (int b = 0; b < B; ++b) h->line_first_chunk[b] = b * Tmax / TOK_PER_CHUNK;
    KOCR_CUDA(wait_stream(h, s));              // `tab` is a host vector of this call
    return run_forced_decode(h, tgt_tokens, B, L, logits_out, VOCAB_PAD, stream, s);
}

// ---- the reference's MODEL protocol (predictor.py:53-78,166-192) on host tensors in the reference's layouts ----------------
namespace {
// tables for a chunk-level stage call: the n chunks are independent there, so they are booked as lines of up to
// max_seq_len / 32 chunks (any n <= max_chunks fits, whatever max_lines is)
int set_chunk_tables(kocr_handle* h, int n, cudaStream_t s) {
    const int per_line = std::max(1, h->max_seq_len / TOK_PER_CHUNK);
    std::vector<int> T;
    for (int left = n; left > 0; left -= per_line) T.push_back(std::min(left, per_line) * TOK_PER_CHUNK);
    return set_token_tables(h, (int)T.size(), T.data(), s);
}
}  // namespace

int kocr_model_cnn(kocr_handle* h, const float* chunks, int n, float* f_out, void* stream) {
    KOCR_CHECK(h != nullptr && chunks != nullptr && f_out != nullptr, "kocr_model_cnn: null argument");
    KOCR_CHECK(n > 0 && n <= h->max_chunks, "kocr_model_cnn: %d chunks exceed the handle's capacity", n);
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    KOCR_TRY(set_chunk_tables(h, n, s));                 // every chunk is its own 32-token "line"
    KOCR_CUDA(cudaMemcpyAsync(buf<float>(h, "chunks"), chunks, (size_t)n * IMG_H * CHUNK_W * 4, cudaMemcpyHostToDevice, s));
    KOCR_TRY(stage_cnn_encoder(h, s, 1));
    std::vector<act16_t> pin((size_t)n * TOK_PER_CHUNK * 1024);
    KOCR_CUDA(cudaMemcpyAsync(pin.data(), buf<act16_t>(h, "patch_in"), pin.size() * 2, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(wait_stream(h, s));
    for (int i = 0; i < n; ++i)                                     // patch_in [n*32 + k][kh*512 + c] -> f (n, 512, 2, 32)
        for (int k = 0; k < TOK_PER_CHUNK; ++k)
            for (int kh = 0; kh < 2; ++kh)
                for (int c = 0; c < 512; ++c)
                    f_out[(((size_t)i * 512 + c) * 2 + kh) * TOK_PER_CHUNK + k] = host_f32(pin[((size_t)i * TOK_PER_CHUNK + k) * 1024 + kh * 512 + c]);
    return 0;
}

int kocr_model_patch(kocr_handle* h, const float* f, int n, float* x_out, void* stream) {
    KOCR_CHECK(h != nullptr && f != nullptr && x_out != nullptr, "kocr_model_patch: null argument");
    KOCR_CHECK(n > 0 && n <= h->max_chunks, "kocr_model_patch: %d chunks exceed the handle's capacity", n);
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    KOCR_TRY(set_chunk_tables(h, n, s));
    std::vector<act16_t> pin((size_t)n * TOK_PER_CHUNK * 1024);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < TOK_PER_CHUNK; ++k)
            for (int kh = 0; kh < 2; ++kh)
                for (int c = 0; c < 512; ++c)
                    pin[((size_t)i * TOK_PER_CHUNK + k) * 1024 + kh * 512 + c] = host_a16(f[(((size_t)i * 512 + c) * 2 + kh) * TOK_PER_CHUNK + k]);
    KOCR_CUDA(cudaMemcpyAsync(buf<act16_t>(h, "patch_in"), pin.data(), pin.size() * 2, cudaMemcpyHostToDevice, s));
    KOCR_TRY(stage_cnn_encoder(h, s, 2));
    KOCR_CUDA(cudaMemcpyAsync(x_out, buf<float>(h, "x"), (size_t)n * TOK_PER_CHUNK * D_MODEL * 4, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(wait_stream(h, s));
    return 0;
}

int kocr_model_enc(kocr_handle* h, const float* p, int n, float* out, void* stream) {
    KOCR_CHECK(h != nullptr && p != nullptr && out != nullptr, "kocr_model_enc: null argument");
    KOCR_CHECK(n > 0 && n <= h->max_chunks, "kocr_model_enc: %d chunks exceed the handle's capacity", n);
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    KOCR_TRY(set_chunk_tables(h, n, s));
    const size_t rows = (size_t)n * TOK_PER_CHUNK;
    std::vector<float> x(rows * D_MODEL);
    std::vector<act16_t> xb(rows * D_MODEL);
    for (int tok = 0; tok < TOK_PER_CHUNK; ++tok)                   // seq-first (32, n, 384) -> chunk-major rows
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < D_MODEL; ++d) {
                const float v = p[((size_t)tok * n + i) * D_MODEL + d];
                x[((size_t)i * TOK_PER_CHUNK + tok) * D_MODEL + d] = v;
                xb[((size_t)i * TOK_PER_CHUNK + tok) * D_MODEL + d] = host_a16(v);
            }
    KOCR_CUDA(cudaMemcpyAsync(buf<float>(h, "x"), x.data(), x.size() * 4, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaMemcpyAsync(buf<act16_t>(h, "xb"), xb.data(), xb.size() * 2, cudaMemcpyHostToDevice, s));
    KOCR_TRY(stage_cnn_encoder(h, s, 4, /*add_global_pos=*/false));
    KOCR_CUDA(cudaMemcpyAsync(x.data(), buf<float>(h, "x"), x.size() * 4, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(wait_stream(h, s));
    for (int tok = 0; tok < TOK_PER_CHUNK; ++tok)
        for (int i = 0; i < n; ++i)
            memcpy(out + ((size_t)tok * n + i) * D_MODEL, x.data() + ((size_t)i * TOK_PER_CHUNK + tok) * D_MODEL, D_MODEL * 4);
    return 0;
}

namespace {
// uploads B sequences of (at most) T rows each into the token-major fp32 buffer `name` (+ its 16-bit twin) at the padded offsets
int upload_sequences(kocr_handle* h, const char* name, const char* name16, const float* src, int B, int T, const int* Ti, cudaStream_t s) {
    std::vector<float> x((size_t)h->n_tok * D_MODEL, 0.f);
    std::vector<act16_t> xb((size_t)h->n_tok * D_MODEL, host_a16(0.f));
    for (int b = 0; b < B; ++b) {
        const size_t off = (size_t)h->line_first_chunk[b] * TOK_PER_CHUNK;
        for (int t = 0; t < Ti[b]; ++t)
            for (int d = 0; d < D_MODEL; ++d) {
                const float v = src[((size_t)b * T + t) * D_MODEL + d];
                x[(off + t) * D_MODEL + d] = v;
                xb[(off + t) * D_MODEL + d] = host_a16(v);
            }
    }
    KOCR_CUDA(cudaMemcpyAsync(buf<float>(h, name), x.data(), x.size() * 4, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaMemcpyAsync(buf<act16_t>(h, name16), xb.data(), xb.size() * 2, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(wait_stream(h, s));              // host vectors of this call
    return 0;
}
}  // namespace

int kocr_model_bilstm(kocr_handle* h, const float* merged, int B, int T, float* out, void* stream) {
    KOCR_CHECK(h != nullptr && merged != nullptr && out != nullptr, "kocr_model_bilstm: null argument");
    KOCR_CHECK(h->variant == 0, "kocr_model_bilstm: this checkpoint family has no context_bilstm");
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    std::vector<int> Ti((size_t)B, T);
    KOCR_TRY(set_token_tables(h, B, Ti.data(), s));
    KOCR_TRY(upload_sequences(h, "x", "xb", merged, B, T, Ti.data(), s));
    KOCR_TRY(stage_memory(h, s, 1));
    std::vector<float> mem((size_t)h->n_tok * D_MODEL);
    KOCR_CUDA(cudaMemcpyAsync(mem.data(), buf<float>(h, "mem"), mem.size() * 4, cudaMemcpyDeviceToHost, s));
    KOCR_CUDA(wait_stream(h, s));
    for (int b = 0; b < B; ++b)
        memcpy(out + (size_t)b * T * D_MODEL, mem.data() + (size_t)h->line_first_chunk[b] * TOK_PER_CHUNK * D_MODEL, (size_t)T * D_MODEL * 4);
    return 0;
}

int kocr_model_dec(kocr_handle* h, const int32_t* tgt, int B, int t, const float* memory, int T, const uint8_t* pad_mask,
                   float* logits_out, void* stream) {
    KOCR_CHECK(h != nullptr && tgt != nullptr && memory != nullptr && logits_out != nullptr, "kocr_model_dec: null argument");
    KOCR_CHECK(t >= 1 && t <= h->dec_max_len, "kocr_model_dec: target length %d outside [1, %d]", t, h->dec_max_len);
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    std::vector<int> Ti((size_t)B, T);
    for (int b = 0; b < B && pad_mask; ++b) {          // memory_key_padding_mask: a suffix of padded positions per line
        int n = T;
        while (n > 0 && pad_mask[(size_t)b * T + n - 1]) --n;
        for (int j = 0; j < n; ++j)
            KOCR_CHECK(!pad_mask[(size_t)b * T + j], "kocr_model_dec: memory_key_padding_mask of line %d is not a suffix mask", b);
        KOCR_CHECK(n > 0, "kocr_model_dec: line %d has every memory position masked", b);
        Ti[b] = n;
    }
    KOCR_TRY(set_token_tables(h, B, Ti.data(), s));
    const bool lstm = h->variant == 0;
    KOCR_TRY(upload_sequences(h, lstm ? "mem" : "x", lstm ? "memb" : "xb", memory, B, T, Ti.data(), s));
    KOCR_TRY(stage_memory(h, s, 2));                   // cross-attention K/V of the given memory
    return run_forced_decode(h, tgt, B, t, logits_out, VOCAB, stream, s);
}

int kocr_crop_lines(kocr_handle* h, const uint8_t* page, int page_h, int page_w, int channels, int page_on_device,
                    const int32_t* boxes, int n_lines, int pad_px, uint8_t* out_pixels_dev, const int64_t* out_offsets,
                    void* stream) {
    KOCR_CHECK(h != nullptr && page != nullptr, "kocr_crop_lines: null argument");
    KOCR_CHECK(n_lines >= 0 && pad_px >= 0 && page_h > 0 && page_w > 0, "kocr_crop_lines: bad sizes");
    KOCR_CHECK(channels == 1 || channels == 3, "kocr_crop_lines: page must be L (1 channel) or RGB (3 channels)");
    if (n_lines == 0) return 0;
    KOCR_CHECK(boxes != nullptr && out_pixels_dev != nullptr && out_offsets != nullptr, "kocr_crop_lines: null argument");
    for (int i = 0; i < n_lines; ++i) {
        const int32_t* b = boxes + 4 * i;
        KOCR_CHECK(b[0] >= 0 && b[1] >= 0 && b[2] <= page_w && b[3] <= page_h && b[2] > b[0] && b[3] > b[1],
                   "kocr_crop_lines: box %d = (%d, %d, %d, %d) is empty or outside the %d x %d page", i, b[0], b[1], b[2], b[3],
                   page_w, page_h);
    }
    KOCR_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
    const uint8_t* d_page = page;
    KOCR_TRY(order_after_default_stream(h, s));       // out_pixels_dev (and a device page) belong to the caller
    if (!page_on_device) {
        const size_t bytes = (size_t)page_h * page_w * channels;
        KOCR_TRY(ensure(h->crop_page, bytes));
        KOCR_CUDA(cudaMemcpyAsync(h->crop_page.p, page, bytes, cudaMemcpyHostToDevice, s));
        d_page = reinterpret_cast<const uint8_t*>(h->crop_page.p);
    }
    const size_t box_bytes = (size_t)n_lines * 16, off_bytes = (size_t)n_lines * 8;
    KOCR_TRY(ensure(h->crop_tab, box_bytes + off_bytes));
    uint8_t* tab = reinterpret_cast<uint8_t*>(h->crop_tab.p);
    KOCR_CUDA(cudaMemcpyAsync(tab, out_offsets, off_bytes, cudaMemcpyHostToDevice, s));
    KOCR_CUDA(cudaMemcpyAsync(tab + off_bytes, boxes, box_bytes, cudaMemcpyHostToDevice, s));
    KOCR_TRY(launch_crop_lines(d_page, page_w, channels, reinterpret_cast<const int*>(tab + off_bytes),
                               reinterpret_cast<const long long*>(tab), n_lines, pad_px, out_pixels_dev, s));
    ++g_launches;
    KOCR_CUDA(wait_stream(h, s));        // boxes / offsets (and a pageable page) are host memory of the caller
    return 0;
}

int kocr_read_unfinished(kocr_handle* h, int32_t* flags_out) {
    KOCR_CHECK(h != nullptr && flags_out != nullptr, "kocr_read_unfinished: null argument");
    // a line is unfinished if it neither emitted <eos> nor used up all decode positions (host-only: the flags
    // were copied to pinned memory at the end of kocr_decode_greedy)
    for (int i = 0; i < h->n_lines; ++i) flags_out[i] = (h->fin_host[i] == 0 && h->last_steps < h->last_max_steps) ? 1 : 0;
    return 0;
}

int kocr_read_kernel_timing(kocr_handle* h, char* text_out, size_t cap) {
    KOCR_CHECK(h != nullptr && text_out != nullptr && cap > 0, "kocr_read_kernel_timing: null argument");
    KOCR_CUDA(cudaSetDevice(h->device));
    KOCR_CUDA(cudaDeviceSynchronize());
    for (auto& p : h->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { h->sites[p.site].ms += ms; h->sites[p.site].count += 1; }
        h->event_pool.push_back(p.a); h->event_pool.push_back(p.b);
    }
    h->pending.clear();
    std::string out;
    char line[256];
    for (auto& st : h->sites) {
        snprintf(line, sizeof line, "%s %.6f %d %.1f\n", st.name.c_str(), st.ms, st.count, st.flops);
        out += line;
    }
    KOCR_CHECK(out.size() + 1 <= cap, "kocr_read_kernel_timing: buffer too small");
    memcpy(text_out, out.c_str(), out.size() + 1);
    return 0;
}

int kocr_debug_read(kocr_handle* h, const char* name, void* dst, size_t dst_bytes, size_t* bytes_out) {
    KOCR_CHECK(h != nullptr && name != nullptr, "kocr_debug_read: null argument");
    KOCR_CUDA(cudaSetDevice(h->device));
    const size_t NC = h->n_chunks, M = h->n_tok;
    const void* src = nullptr;
    size_t bytes = 0;
    std::string n(name);
    auto act = [&](const char* b, const StageGeom& g, int C) { auto it = h->named.find(b); if (it != h->named.end()) { src = it->second.p; bytes = NC * px(g) * C * 2; } };
    if (n == "chunks") { src = h->named["chunks"].p; bytes = NC * IMG_H * CHUNK_W * 4; }
    else if (n == "pool1") act("pool1", G1, 64);
    else if (n == "conv2") act("conv2", G1, 128);
    else if (n == "pool2") act("pool2", G2, 128);
    else if (n == "conv3") act("conv3", G2, 256);
    else if (n == "conv4") act("conv4", G2, 256);            // ResNet baseline only (the SE / VGG conv4 is pooled in its epilogue)
    else if (n == "pool3") act("pool3", G3, 256);
    else if (n == "conv5") act("conv5", G3, 512);
    else if (n == "conv6") act("conv6", G3, 512);            // ResNet baseline only
    else if (n == "pool4") act("pool4", G4, 512);
    else if (n == "conv7") act("conv7", G4, 512);            // ResNet baseline only
    else if (n == "bins7") { src = h->named["bins7"].p; bytes = NC * 25 * 2 * 512 * 2; }
    else if (n == "se_mean3") { src = h->named["se_mean3"].p; bytes = NC * 25 * 256 * 2; }
    else if (n == "se_mean4" || n == "se_mean5") { src = h->named[n].p; bytes = NC * 25 * 512 * 2; }
    else if (n == "patch_in") { src = h->named["patch_in"].p; bytes = M * 1024 * 2; }
    else if (n == "enc") { src = h->named["x"].p; bytes = M * D_MODEL * 4; }
    else if (n == "memory") { src = h->named[h->variant == 0 ? "mem" : "x"].p; bytes = M * D_MODEL * 4; }
    else if (n == "dx" || n == "dy" || n == "daof" || n == "dq") { src = h->named[n].p; bytes = (size_t)h->n_lines * D_MODEL * 4; }
    else if (n == "dqkv") { src = h->named[n].p; bytes = (size_t)h->n_lines * 3 * D_MODEL * 4; }
    else if (n == "dh") { src = h->named[n].p; bytes = (size_t)h->n_lines * 4 * D_MODEL * 4; }
    else if (n == "logits") { src = h->named[n].p; bytes = (size_t)h->n_lines * VOCAB_PAD * 4; }
    else if (n == "logits_trace") { src = h->trace.p; bytes = (size_t)h->n_lines * DEC_MAX * VOCAB_PAD * 4; }
    else if (n == "last_steps") { if (bytes_out) *bytes_out = (size_t)h->last_steps; return 0; }
    else if (n == "host_launch_us") { if (bytes_out) *bytes_out = (size_t)h->host_launch_us; return 0; }
    else if (n == "host_wait_us") { if (bytes_out) *bytes_out = (size_t)h->host_wait_us; return 0; }
    KOCR_CHECK(src != nullptr, "kocr_debug_read: unknown or empty buffer '%s'", name);
    if (bytes_out) *bytes_out = bytes;
    if (dst == nullptr) return 0;
    KOCR_CHECK(dst_bytes >= bytes, "kocr_debug_read: '%s' needs %zu bytes, got %zu", name, bytes, dst_bytes);
    KOCR_CUDA(cudaDeviceSynchronize());
    KOCR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int kocr_test_gemm(int impl, const void* a_a16, int64_t rows_a, const void* w_a16, int m, int n, int taps, int cin,
                   int conv_h, int conv_w, int tile_cols, int col_mode, const float* bias, int relu, float* out_f32,
                   void* out_a16, void* out_pool, void* out_colmean, void* stream) {
    GemmProblem p;
    memset(&p, 0, sizeof p);
    p.M = m; p.N = n; p.taps = taps; p.cin = cin;
    if (taps == 9) {
        KOCR_CHECK(conv_h > 0 && conv_w > 0 && m % (conv_h * conv_w) == 0, "kocr_test_gemm: M = %d is not a whole number of %d x %d images", m, conv_w, conv_h);
        p.conv_H = conv_h; p.conv_W = conv_w; p.n_img = m / (conv_h * conv_w); p.tile_cols = tile_cols;
    }
    p.ep.bias = bias; p.ep.relu = relu;
    p.ep.out_f32 = out_f32; p.ep.ld_f32 = n;
    p.ep.out_a16 = reinterpret_cast<act16_t*>(out_a16); p.ep.ld_a16 = n;
    p.ep.col_mode = col_mode; p.ep.out_pool = reinterpret_cast<act16_t*>(out_pool); p.ep.out_colmean = reinterpret_cast<act16_t*>(out_colmean);
    if (col_mode) p.bn = n % 256 == 0 ? 256 : 128;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    p.tf32 = impl == 2 ? 1 : 0;          // impl 2: fp32 operands consumed as TF32
    if (impl == 1)
        return launch_gemm_simt_check(reinterpret_cast<const act16_t*>(a_a16), rows_a,
                                      reinterpret_cast<const act16_t*>(w_a16), p, s);
    int dev = 0, sms = 148;
    KOCR_CUDA(cudaGetDevice(&dev));
    KOCR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return launch_gemm_tc(a_a16, rows_a, w_a16, p, sms, s);
}

}  // extern "C"
