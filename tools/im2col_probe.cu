// Probe of TMA im2col-mode loads (cp.async.bulk.tensor.4d...im2col) on sm_100a: checks the semantics the implicit-GEMM
// convolution relies on, against a CPU restatement, before the GEMM kernel is built on them.
//   tensor: NWHC activation [N][W][H][C] declared to TMA as (C, "W" = H, "H" = W, N); 3x3 filter, pad 1
//   one load = `ppc` consecutive output pixels (linear index over n, w, h with h fastest) x 64 channels of ONE filter tap,
//   128-byte swizzle, halo and the tail past the tensor zero-filled by the hardware.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o im2col_probe tools/im2col_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int cw, int ch, int cn, int ow, int oh,
                             uint32_t expect_bytes, uint16_t* out, int* status) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 128 * 128);
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) ((uint16_t*)smem)[i] = 0x5555;   // poison
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(expect_bytes), "r"(smem_u32(bar)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(smem_u32(smem)), "l"((uint64_t)&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(cw), "r"(ch), "r"(cn),
              "h"((uint16_t)ow), "h"((uint16_t)oh)
            : "memory");
        int ok = 0;
        for (long spin = 0; spin < (1L << 22); ++spin) {      // bounded wait: a wrong byte count must not hang the GPU
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
            if (done) { ok = 1; break; }
        }
        *status = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) out[i] = ((uint16_t*)smem)[i];
}

typedef CUresult (*PFN_im2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 5, W = 25, H = 6, C = 128;
    std::vector<uint16_t> act((size_t)N * W * H * C);
    for (size_t i = 0; i < act.size(); ++i) act[i] = (uint16_t)(1 + (i * 2654435761u >> 7) % 0x7BFE);   // finite, non-zero fp16 bit patterns
    uint16_t* d_act; CK(cudaMalloc(&d_act, act.size() * 2));
    CK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
    uint16_t* d_out; CK(cudaMalloc(&d_out, 128 * 64 * 2));
    int* d_status; CK(cudaMalloc(&d_status, 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeIm2col entry point\n"); return 2; }
    PFN_im2col enc = (PFN_im2col)fn;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 2048));
    int fails = 0, total = 0;
    for (int ppc : {128, 126, 120}) {
        CUtensorMap tm;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N};          // (C, "W" = h, "H" = w, N)
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * C * 2, (cuuint64_t)W * H * C * 2};
        int lower[2] = {-1, -1}, upper[2] = {-1, -1};                                               // 3x3, pad 1 (fprop)
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d_act, gdim, gstr, lower, upper, 64, (cuuint32_t)ppc, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("ppc=%d encode rc=%d\n", ppc, (int)r);
        if (r != CUDA_SUCCESS) { ++fails; continue; }
        const int P = N * W * H;
        std::vector<int> starts = {0, ppc, 3 * ppc, 5 * ppc, (P / ppc) * ppc /* tail tile */, 17 /* unaligned */};
        for (int m0 : starts) for (int tap = 0; tap < 9; ++tap) for (int c0 : {0, 64}) {
            if (m0 >= P) continue;
            const int r_ = tap / 3, s_ = tap % 3;              // r_: shift along h (dh = r_ - 1), s_: shift along w (dw = s_ - 1)
            const int n = m0 / (W * H), rem = m0 % (W * H), w = rem / H, h = rem % H;
            CK(cudaMemset(d_status, 0xff, 4));
            probe_kernel<<<1, 128, 128 * 128 + 2048>>>(tm, c0, h - 1, w - 1, n, r_, s_, (uint32_t)ppc * 128, d_out, d_status);
            CK(cudaDeviceSynchronize());
            int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
            std::vector<uint16_t> got(128 * 64);
            CK(cudaMemcpy(got.data(), d_out, got.size() * 2, cudaMemcpyDeviceToHost));
            ++total;
            if (st != 1) { printf("  ppc=%d m0=%d tap=%d c0=%d: TIMEOUT (expect_tx never completed)\n", ppc, m0, tap, c0); ++fails; continue; }
            int bad = 0, first = -1;
            for (int row = 0; row < ppc; ++row) {
                const long m = (long)m0 + row;
                int nn = (int)(m / (W * H)), rr = (int)(m % (W * H)), ww = rr / H, hh = rr % H;
                const int sh = hh + r_ - 1, sw = ww + s_ - 1;
                const bool in = m < P && sh >= 0 && sh < H && sw >= 0 && sw < W;
                for (int j = 0; j < 8; ++j) for (int e = 0; e < 8; ++e) {
                    const uint16_t want = in ? act[(((size_t)nn * W + sw) * H + sh) * C + c0 + j * 8 + e] : 0;
                    const uint16_t g = got[(size_t)row * 64 + ((j ^ (row & 7)) * 8) + e];
                    if (g != want) { if (first < 0) first = row * 64 + j * 8 + e; ++bad; }
                }
            }
            if (bad) { printf("  ppc=%d m0=%d tap=%d c0=%d: %d mismatching elements (first at row %d col %d)\n", ppc, m0, tap, c0, bad, first / 64, first % 64); ++fails; }
        }
    }
    printf("im2col probe: %d cases, %d failed\n", total, fails);
    return fails ? 1 : 0;
}
