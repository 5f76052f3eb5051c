// FP64 dependent-issue latency and single-warp throughput on this GPU (informs the exact integrator's tile shape).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/micro/fp64_lat scripts/micro/fp64_lat.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int CH, int KIND>
__global__ void chain(double* out, long long* cyc, int iters, double m, double b) {
    double a[CH];
    for (int k = 0; k < CH; k++) a[k] = threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            if (KIND == 0) a[k] = a[k] * m;                 // dependent DMUL
            else if (KIND == 1) a[k] = a[k] + b;            // dependent DADD
            else a[k] = a[k] + (a[k] * m) * b;              // DMUL, DMUL, DADD (shape of one integrand term)
        }
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < CH; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH, int KIND> void run(const char* name, int warps) {
    double* d; long long* c; cudaMalloc(&d, 8 * 1024 * 148); cudaMalloc(&c, 8);
    const int iters = 4096;
    chain<CH, KIND><<<1, 32 * warps>>>(d, c, iters, 1.0000001, 1e-9);
    chain<CH, KIND><<<1, 32 * warps>>>(d, c, iters, 1.0000001, 1e-9);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const int ops = KIND == 2 ? 3 : 1;
    printf("%-10s chains/thread %2d warps %2d: %.2f cycles per dependent step, %.2f cycles per warp-instruction\n", name, CH, warps, (double)h / iters,
           (double)h / iters / (CH * ops));
    cudaFree(d); cudaFree(c);
}
int main() {
    run<1, 0>("DMUL", 1); run<1, 1>("DADD", 1); run<1, 2>("MUL,MUL,ADD", 1);
    run<4, 0>("DMUL", 1); run<8, 0>("DMUL", 1); run<16, 0>("DMUL", 1);
    run<16, 2>("MUL,MUL,ADD", 1); run<16, 2>("MUL,MUL,ADD", 4); run<16, 2>("MUL,MUL,ADD", 8); run<16, 2>("MUL,MUL,ADD", 16);
    run<8, 2>("MUL,MUL,ADD", 8); run<4, 2>("MUL,MUL,ADD", 8); run<4, 2>("MUL,MUL,ADD", 16); run<4, 2>("MUL,MUL,ADD", 32);
    return 0;
}
