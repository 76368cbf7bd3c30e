// Shared helpers for the sm_100a kernels: error plumbing, a16 packing, PTX wrappers for
// mbarrier / TMA / tcgen05 (TMEM + UMMA).  Compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

namespace kocr {

// ------------------------------------------------------------------------------------------
// "a16" = the 16-bit storage format of activations and tensor-core operands on stages 2-5a.
// Default: IEEE fp16 (11 significant bits).  The tensor pipe runs fp16 and bf16 at the same rate
// (tcgen05.mma.kind::f16 takes either), HBM traffic is identical, and the 8x finer rounding is what gets
// greedy-decoded token sequences to agree with the fp32 reference (with bf16 storage 2 of
// 256 lines of the c2 batch flip a near-tie).  Every layer on this path is BatchNorm/LayerNorm-bounded;
// conversions saturate to +-65504 instead of overflowing to inf.  -DKOCR_A16_BF16 builds the bf16 variant.
// ------------------------------------------------------------------------------------------
#ifdef KOCR_A16_BF16
typedef __nv_bfloat16 act16_t;
typedef __nv_bfloat162 act16x2_t;
#define KOCR_A16_FORMAT 0
#define KOCR_MMA_A16 "bf16"
#else
typedef __half act16_t;
typedef __half2 act16x2_t;
#define KOCR_A16_FORMAT 1
#define KOCR_MMA_A16 "f16"
#endif

// ------------------------------------------------------------------------------------------
// Host-side error plumbing: every C-ABI entry returns int (0 ok); the message is thread-local.
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define KOCR_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::kocr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                              cudaGetErrorString(_e));                                       \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

#define KOCR_CHECK(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::kocr::set_error(__VA_ARGS__);                                                  \
            return 2;                                                                        \
        }                                                                                    \
    } while (0)

#define KOCR_TRY(expr)                                                                       \
    do {                                                                                     \
        int _r = (expr);                                                                     \
        if (_r != 0) return _r;                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Inside the decode loop every kernel is launched with the
// programmatic-stream-serialization attribute: it may start (prologue: barrier init, TMEM alloc, descriptor
// prefetch, smem carve-up) while its predecessor is still running, and blocks in pdl_wait() until the
// predecessor has completed and its writes are visible.  Every kernel that can be launched this way calls
// pdl_trigger() first and pdl_wait() before touching global memory; both are no-ops for a normal launch.
// ------------------------------------------------------------------------------------------
bool pdl_active();               // true while the calling thread is enqueueing a PDL chain (decode step)
void pdl_set_active(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_active() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: one flag per (launch site, device), so that a
// process holding handles on several GPUs opts in on each of them.  Racing threads may both set it (harmless).
struct PerDeviceOnce { std::atomic<unsigned char> done[64]; };
template <typename K>
inline cudaError_t opt_in_dynamic_smem(PerDeviceOnce& st, K kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && st.done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && tracked) st.done[dev].store(1, std::memory_order_release);
    return e;
}

// ------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
#ifdef KOCR_A16_BF16
__device__ __forceinline__ uint32_t pack_a16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float a16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float a16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ act16_t to_a16(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float from_a16(act16_t x) { return __bfloat162float(x); }
#else
__device__ __forceinline__ uint32_t pack_a16(float lo, float hi) {      // round to nearest even, saturating
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float a16_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float a16_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
__device__ __forceinline__ act16_t to_a16(float x) { return __ushort_as_half((unsigned short)(pack_a16(x, 0.f) & 0xffffu)); }
__device__ __forceinline__ float from_a16(act16_t x) { return __half2float(x); }
#endif
// fp32 -> TF32 with round-to-nearest (ties away).  tcgen05.mma.kind::tf32 TRUNCATES its fp32 operands to 10 mantissa bits
// (a biased error of up to 2^-10); kernels that produce a TF32 GEMM operand round it here instead, which halves the
// error and removes the bias - 6 of the 8 lines of the 8192-line c3 set that decoded differently from the fp32
// reference came from that truncation (tests/parity/parity_rootcause8.py, profiles/r02/parity_rootcause8_fixes.json).
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t a16x2_max(uint32_t a, uint32_t b) {
    act16x2_t r = __hmax2(*reinterpret_cast<act16x2_t*>(&a), *reinterpret_cast<act16x2_t*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// 16-byte asynchronous global -> shared copy through L2 (no registers, no L1 allocation)
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
        "elect.sync R|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
        : "memory");
}

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2D tiled load: crd0 = innermost (element) coordinate, crd1 = row coordinate (may be negative /
// past the end: out-of-bounds elements are zero-filled and still counted in complete_tx).
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst,
                                            int32_t crd0, int32_t crd1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)),
          "r"(crd0), "r"(crd1)
        : "memory");
}

// 4D im2col-mode load (3x3 convolution operand): the tensor map describes the dense activation as (C, d1, d2, N) with
// a bounding box of base pixels [-1, d - 1) in both spatial dims (pad 1); {crd_c, crd1, crd2, crd_n} is the first base
// pixel (output position - 1), {off1, off2} in [0, 3) the filter tap added to every base pixel.  The hardware walks
// `pixelsPerColumn` consecutive output pixels (d1 fastest, wrapping into d2 and N), zero-fills the halo and the tail past
// the tensor, and always signals pixelsPerColumn * 128 bytes (verified on B200 by tools/im2col_probe.cu).
__device__ __forceinline__ void tma_load_im2col_4d(const void* desc, uint64_t* bar, void* smem_dst, int32_t crd_c,
                                                   int32_t crd1, int32_t crd2, int32_t crd_n, uint16_t off1, uint16_t off2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)),
          "r"(crd_c), "r"(crd1), "r"(crd2), "r"(crd_n), "h"(off1), "h"(off2)
        : "memory");
}

// named barrier over `nthreads` threads of the CTA (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// ---- tcgen05 -----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B (64 a16), 8-row
// atoms 1024 B apart (SBO), LBO unused.  bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1 (sm_100), [49,52) base offset, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>((smem_addr >> 7) & 0x7u) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// Instruction descriptor, A and B K-major, D = f32, M x N tile.  fmt (A and B operand format): 0 = F16, 1 = BF16 (both kind::f16), 2 = TF32 (kind::tf32).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}
static constexpr uint32_t IDESC_FMT_A16 = KOCR_A16_FORMAT == 1 ? 0u : 1u;
__host__ __device__ constexpr uint32_t make_idesc_a16(int M, int N) { return make_idesc(M, N, IDESC_FMT_A16); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void umma_a16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with fp32 operands in shared memory consumed as TF32 (K = 8 per instruction = 32 bytes).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: the inverse of tmem_ld32 (this warp's 32 lanes x 32 consecutive fp32 columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#endif  // __CUDACC__

}  // namespace kocr
