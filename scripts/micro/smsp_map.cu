// Which SM sub-partition (warp scheduler + FP64 pipe) does warp w of a CTA run on, and do two co-resident CTAs share the mapping?
// Warps selected by a mask run a fixed FP64 loop (8 independent chains: one warp alone nearly saturates its sub-partition's FP64
// pipe); the others exit.  Two warps on the same sub-partition take twice as long as two warps on different ones.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/micro/smsp_map scripts/micro/smsp_map.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(256, 2) work(double* out, unsigned* sm_slot, unsigned mask_even, unsigned mask_odd, int iters, int by_slot, unsigned* dump) {
    extern __shared__ double sm[];
    __shared__ unsigned s_par;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_par = atomicAdd(&sm_slot[smid], 1u) & 1u;   // 0 / 1: first / second CTA on this SM
    }
    __syncthreads();
    const unsigned mask = s_par ? mask_odd : mask_even;
    const unsigned warp = threadIdx.x / 32;
    unsigned hwid, smid2;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hwid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid2));
    if (dump && smid2 == 0 && threadIdx.x % 32 == 0) dump[s_par * 8 + warp] = hwid;
    // by_slot: the mask selects sub-partitions (hardware warp slot % 4) instead of warp indices
    if (!(mask >> (by_slot ? (hwid & 3u) : warp) & 1u)) return;
    double a[8];
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = a[k] * 1.0000001;
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = a[k] + 1e-9;
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float run(const char* name, int grid, unsigned me, unsigned mo, int by_slot = 0) {
    static double* d = nullptr; static unsigned* slot = nullptr;
    static unsigned* dump = nullptr;
    if (!d) { cudaMalloc(&d, 8 * 256 * 296); cudaMalloc(&slot, 4 * 256); cudaMalloc(&dump, 64); cudaMemset(dump, 0xff, 64); }
    cudaFuncSetAttribute(work, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        cudaMemset(slot, 0, 4 * 256);
        cudaEventRecord(e0);
        work<<<grid, 256, 100 * 1024>>>(d, slot, me, mo, 20000, by_slot, dump);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-72s: %.3f ms\n", name, best);
    if (by_slot == 2) { unsigned h[16]; cudaMemcpy(h, dump, 64, cudaMemcpyDeviceToHost); printf("  %%warpid on SM 0, first CTA:"); for (int k = 0; k < 8; k++) printf(" %u", h[k]); printf(" | second CTA:"); for (int k = 0; k < 8; k++) printf(" %u", h[8 + k]); printf("\n"); }
    return best;
}

int main() {
    // one CTA per SM
    run("1 CTA/SM, warp 0", 148, 0x01, 0x01);
    run("1 CTA/SM, warps 0,1 (different sub-partitions if w % 4)", 148, 0x03, 0x03);
    run("1 CTA/SM, warps 0,4 (same sub-partition if w % 4)", 148, 0x11, 0x11);
    run("1 CTA/SM, warps 0,5", 148, 0x21, 0x21);
    run("1 CTA/SM, warps 0,1,2,3", 148, 0x0f, 0x0f);
    run("1 CTA/SM, warps 0,4 + 1,5", 148, 0x33, 0x33);
    // two CTAs per SM
    run("2 CTAs/SM, warp 0 in both", 296, 0x01, 0x01);
    run("2 CTAs/SM, warp 0 | warp 1", 296, 0x01, 0x02);
    run("2 CTAs/SM, warp 0 | warp 4", 296, 0x01, 0x10);
    run("2 CTAs/SM, warps 2..7 in both (two staging warps, as shipped)", 296, 0xfc, 0xfc);
    run("2 CTAs/SM, warps 2..7 | warps 0,1,4..7 (balanced if w % 4, no offset)", 296, 0xfc, 0xf3);
    run("2 CTAs/SM, warps 1..7 in both (one staging warp, as shipped)", 296, 0xfe, 0xfe);
    run("2 CTAs/SM, warps 1..7 | warps 0,1,3..7", 296, 0xfe, 0xfb);
    run("2 CTAs/SM, all 8 warps in both", 296, 0xff, 0xff);
    run("2 CTAs/SM, slots % 4 == 0 in both (4 warps on one sub-partition if slot % 4)", 296, 0x1, 0x1, 1);
    run("2 CTAs/SM, slots % 4 == 0 | slots % 4 == 1 (2 + 2)", 296, 0x1, 0x2, 1);
    run("2 CTAs/SM, slots % 4 in {0,1,2} in both (6 + 6 warps, 3 per sub-partition)", 296, 0x7, 0x7, 2);
    return 0;
}
