// Stage 1: Pillow-exact BILINEAR resize to height 48, 48x100/overlap-16 chunk gather, white
// padding and (x/255 - 0.5)/0.5 normalisation, bit-exact against
// ImagePreprocessor.process (reference: netra_ocr/recognition/preprocessor.py:35-58) and
// Pillow's Resample.c (precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
// ImagingResampleVertical_8bpc), which the reference calls at preprocessor.py:49.
//
// HBM-bound byte work: no tensor cores.  Horizontal pass: one thread per output column (its
// coefficients live in registers/local memory, rows are streamed, neighbouring threads read
// neighbouring source bytes).  Vertical pass is fused with the chunk gather: one thread per 4
// output pixels, float4 stores, 256-entry LUT for the two IEEE fp32 ops.
#include "kernels.cuh"

namespace kocr {

static constexpr int KMAX = 64;          // max filter taps (down-scaling factor <= 31)
static constexpr int PRECISION_BITS = 22;

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one output index, BILINEAR filter.
// Every double operation is an explicit round-to-nearest intrinsic so that nvcc cannot contract
// a multiply-add into an FMA (the C original is compiled without FMA).
__device__ __forceinline__ void bilinear_coeffs(int in_size, int out_size, int xx, int& xmin_out, int& cnt_out,
                                                int* kk) {
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                       // 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > KMAX) xmax = KMAX;                              // guarded on the host (ksize <= KMAX)
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        const double v = __dmul_rn(w, (double)(1 << PRECISION_BITS));
        kk[x] = w < 0.0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
    }
    xmin_out = xmin;
    cnt_out = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: src (h, w) u8 -> mid (h, new_w) u8.  grid = (ceil(max_new_w/128), n_lines).
__global__ void __launch_bounds__(128) resize_h_kernel(const uint8_t* __restrict__ pixels, uint8_t* __restrict__ mid,
                                                       const LineDesc* __restrict__ lines) {
    const LineDesc L = lines[blockIdx.y];
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= L.new_w) return;
    const uint8_t* src = pixels + L.src_off;
    uint8_t* dst = mid + L.mid_off;
    if (L.new_w == L.w) {               // Pillow skips the horizontal pass when the width is unchanged
        for (int y = 0; y < L.h; ++y) dst[(long)y * L.new_w + xx] = src[(long)y * L.w + xx];
        return;
    }
    int kk[KMAX];
    int xmin, cnt;
    bilinear_coeffs(L.w, L.new_w, xx, xmin, cnt, kk);
    for (int y = 0; y < L.h; ++y) {
        const uint8_t* row = src + (long)y * L.w + xmin;
        int acc = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < cnt; ++x) acc += (int)row[x] * kk[x];
        dst[(long)y * L.new_w + xx] = clip8(acc);
    }
}

// Vertical coefficients: table [line][48][2 + KMAX] ints.  grid = n_lines, block = 64 (48 active).
__global__ void resize_vcoef_kernel(const LineDesc* __restrict__ lines, int* __restrict__ vtab) {
    const LineDesc L = lines[blockIdx.x];
    const int yy = threadIdx.x;
    if (yy >= IMG_H) return;
    int* t = vtab + ((long)blockIdx.x * IMG_H + yy) * (2 + KMAX);
    int kk[KMAX];
    int ymin, cnt;
    if (L.h == IMG_H) {                 // vertical pass skipped: identity
        ymin = yy; cnt = 1; kk[0] = 1 << PRECISION_BITS;
    } else {
        bilinear_coeffs(L.h, IMG_H, yy, ymin, cnt, kk);
    }
    t[0] = ymin; t[1] = cnt;
    for (int i = 0; i < cnt; ++i) t[2 + i] = kk[i];
}

// Vertical pass + chunk gather + normalise.  One thread per 4 consecutive output pixels.
// out: fp32 (n_chunks, 1, 48, 100).
__global__ void __launch_bounds__(256) resize_v_chunk_kernel(const uint8_t* __restrict__ mid,
                                                             const LineDesc* __restrict__ lines,
                                                             const int* __restrict__ chunk_line,
                                                             const int* __restrict__ vtab, float* __restrict__ out,
                                                             int n_chunks) {
    __shared__ float lut[256];
    // the two IEEE fp32 operations of ToTensor + Normalize (preprocessor.py:50,55)
    lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.0f), 0.5f), 0.5f);
    __syncthreads();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over n_chunks * 48 * 25
    if (idx >= (long)n_chunks * IMG_H * (CHUNK_W / 4)) return;
    const int x4 = (int)(idx % (CHUNK_W / 4));
    const int y = (int)((idx / (CHUNK_W / 4)) % IMG_H);
    const int c = (int)(idx / ((CHUNK_W / 4) * IMG_H));
    const int li = chunk_line[c];
    const LineDesc L = lines[li];
    const int k = c - L.first_chunk;
    const int* t = vtab + ((long)li * IMG_H + y) * (2 + KMAX);
    const int ymin = t[0], cnt = t[1];
    const uint8_t* base = mid + L.mid_off;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xs = k * CHUNK_STRIDE + x4 * 4 + j;
        int v = 255;                                              // white pad (preprocessor.py:25-28)
        if (xs < L.new_w) {
            int acc = 1 << (PRECISION_BITS - 1);
            for (int i = 0; i < cnt; ++i) acc += (int)base[(long)(ymin + i) * L.new_w + xs] * t[2 + i];
            v = clip8(acc);
        }
        r[j] = lut[v];
    }
    reinterpret_cast<float4*>(out)[idx] = make_float4(r[0], r[1], r[2], r[3]);
}

int launch_preprocess(const uint8_t* d_pixels, uint8_t* d_mid, const LineDesc* d_lines, const int* d_chunk_line,
                      int* d_vtab, float* d_chunks, int n_lines, int n_chunks, int max_new_w, cudaStream_t stream) {
    if (n_lines == 0 || n_chunks == 0) return 0;
    dim3 gh((max_new_w + 127) / 128, n_lines);
    resize_h_kernel<<<gh, 128, 0, stream>>>(d_pixels, d_mid, d_lines);
    resize_vcoef_kernel<<<n_lines, 64, 0, stream>>>(d_lines, d_vtab);
    const long total = (long)n_chunks * IMG_H * (CHUNK_W / 4);
    resize_v_chunk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_mid, d_lines, d_chunk_line, d_vtab,
                                                                              d_chunks, n_chunks);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

int preprocess_vtab_ints_per_line() { return IMG_H * (2 + KMAX); }
int preprocess_kmax() { return KMAX; }

// ------------------------------------------------------------------------------------------
// Input side: text-line crops of a page image, the way the reference cuts them before recognition -
// extract_textline_crops (netra_ocr/textline_detection.py:7-53: crop the expanded, clipped box, paste it on a white
// canvas with `padding_px` on every side) and the custom-detector branch of OCREngine (ocr_engine.py:72-76: clipped
// box, no canvas) - followed by `convert('L')` (preprocessor.py:39-41; Pillow rgb2l: (19595 R + 38470 G + 7471 B +
// 0x8000) >> 16).  Pure byte indexing: one CTA row-strides over a line, a warp reads 96 contiguous bytes of a page
// row.  grid = (n_lines, CROP_ROW_CTAS).
// ------------------------------------------------------------------------------------------
static constexpr int CROP_ROW_CTAS = 8;

__global__ void __launch_bounds__(256) crop_lines_kernel(const uint8_t* __restrict__ page, int page_w, int channels,
                                                         const int* __restrict__ boxes, const long long* __restrict__ offsets,
                                                         int pad, uint8_t* __restrict__ out) {
    const int line = blockIdx.x;
    const int x0 = boxes[4 * line], y0 = boxes[4 * line + 1], x1 = boxes[4 * line + 2], y1 = boxes[4 * line + 3];
    const int cw = x1 - x0, chh = y1 - y0, ow = cw + 2 * pad, oh = chh + 2 * pad;
    uint8_t* dst = out + offsets[line];
    for (int r = blockIdx.y; r < oh; r += gridDim.y) {
        const int sy = r - pad;
        const bool row_in = sy >= 0 && sy < chh;
        const uint8_t* src = page + ((long)(y0 + sy) * page_w + x0) * channels;
        for (int c = threadIdx.x; c < ow; c += blockDim.x) {
            const int sx = c - pad;
            uint8_t v = 255;                                    // white canvas (textline_detection.py:44)
            if (row_in && sx >= 0 && sx < cw) {
                if (channels == 1) {
                    v = src[sx];
                } else {
                    const uint8_t* px = src + (long)sx * channels;
                    v = (uint8_t)((19595u * px[0] + 38470u * px[1] + 7471u * px[2] + 0x8000u) >> 16);
                }
            }
            dst[(long)r * ow + c] = v;
        }
    }
}

int launch_crop_lines(const uint8_t* d_page, int page_w, int channels, const int* d_boxes, const long long* d_offsets,
                      int n_lines, int pad, uint8_t* d_out, cudaStream_t stream) {
    if (n_lines == 0) return 0;
    crop_lines_kernel<<<dim3(n_lines, CROP_ROW_CTAS), 256, 0, stream>>>(d_page, page_w, channels, d_boxes, d_offsets, pad, d_out);
    KOCR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace kocr
