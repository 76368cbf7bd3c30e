// Probe of programmatic-dependent-launch semantics on this driver/GPU (dev tool, not product code).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// stage kernel: out[i] = in[i] + 1 after a delay; optionally triggers early / waits
__global__ void stage(const int* in, int* out, int n, int early_trigger, int do_wait, long long spin) {
    if (early_trigger) pdl_trigger();
    if (do_wait) pdl_wait();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldcg(in + i) + 1;
}

static void launch(const int* in, int* out, int n, int early, int wait, long long spin, bool pdl, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n + 255) / 256); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, stage, in, out, n, early, wait, spin);
}

int main() {
    const int n = 4096, depth = 12;
    int* buf[depth + 1];
    for (int i = 0; i <= depth; ++i) { cudaMalloc(&buf[i], n * 4); cudaMemset(buf[i], 0, n * 4); }
    cudaStream_t s; cudaStreamCreate(&s);
    int* host = new int[n];
    for (int mode = 0; mode < 4; ++mode) {
        const int early = mode & 1, pdl = (mode >> 1) & 1;
        for (int rep = 0; rep < 3; ++rep) {
            for (int i = 1; i <= depth; ++i) cudaMemsetAsync(buf[i], 0xff, n * 4, s);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, s);
            for (int i = 0; i < depth; ++i) launch(buf[i], buf[i + 1], n, early, 1, 20000, pdl, s);
            cudaEventRecord(b, s);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaMemcpy(host, buf[depth], n * 4, cudaMemcpyDeviceToHost);
            int bad = 0; for (int i = 0; i < n; ++i) bad += host[i] != depth;
            printf("pdl=%d early_trigger=%d rep=%d: %.1f us for %d kernels, wrong=%d (err=%s)\n", pdl, early, rep, ms * 1e3, depth, bad,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
