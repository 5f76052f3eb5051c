// Shared-memory load cost by width and lane pattern on this GPU: cycles the LSU data stage spends per warp-level LDS (informs the
// lane -> micro-tile mapping of the persistent integrator).  8 warps of one CTA issue independent loads back to back.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/lds_bw scripts/micro/lds_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// pattern: byte offset of lane l = (l / group) * stride_bytes  (group lanes share one address)
template <int WIDTH>   // bytes per lane: 8 or 16
__global__ void lds(double* out, long long* cyc, int iters, int group, int stride_bytes, int interleave) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int k = threadIdx.x; k < 48 * 1024 / 8; k += blockDim.x) reinterpret_cast<double*>(sm)[k] = k;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + (interleave ? lane % (32 / group) : lane / group) * stride_bytes;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;   // integer accumulation: the FP64 pipe must not be what bounds the loop
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t ad = a + ((i * 8 + u) & 15) * 2048;   // 16 different windows of the 48 KB
            if (WIDTH == 16) {
                uint32_t x, y, z, w;
                asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(ad));
                if (u & 1) { s0 ^= x ^ y; s1 ^= z ^ w; } else { s2 ^= x ^ y; s3 ^= z ^ w; }
            } else {
                uint32_t x, y;
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(ad));
                if (u & 1) { s0 += x; s1 += y; } else { s2 += x; s3 += y; }
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(s0 ^ s1 ^ s2 ^ s3);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int WIDTH> void run(const char* name, int group, int stride_bytes, int interleave = 0) {
    double* d; long long* c; cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 8);
    const int iters = 2048, warps = 16;
    for (int r = 0; r < 2; r++) lds<WIDTH><<<1, 32 * warps, 48 * 1024>>>(d, c, iters, group, stride_bytes, interleave);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("LDS.%-3d %-44s: %.2f cycles per warp-level load\n", WIDTH * 8, name, (double)h / (iters * 8.0 * warps));
    cudaFree(d); cudaFree(c);
}

int main() {
    run<8>("all lanes one address", 32, 0);
    run<8>("lanes consecutive (8 B apart)", 1, 8);
    run<8>("lanes 16 B apart", 1, 16);
    run<8>("lanes 32 B apart", 1, 32);
    run<8>("4 addresses x 8 lanes, 32 B apart", 8, 32);
    run<8>("8 addresses x 4 lanes, 16 B apart", 4, 16);
    run<16>("all lanes one address", 32, 0);
    run<16>("lanes consecutive (16 B apart)", 1, 16);
    run<16>("lanes 32 B apart", 1, 32);
    run<16>("2 addresses x 16 lanes, 32 B apart", 16, 32);
    run<16>("4 addresses x 8 lanes, 32 B apart", 8, 32);
    run<16>("4 addresses x 8 lanes, 16 B apart", 8, 16);
    run<16>("8 addresses x 4 lanes, 16 B apart", 4, 16);
    run<16>("8 addresses x 4 lanes, 32 B apart", 4, 32);
    run<16>("16 addresses x 2 lanes, 16 B apart", 2, 16);
    run<16>("4 addresses, lane % 4, 32 B apart", 8, 32, 1);
    run<16>("8 addresses, lane % 8, 16 B apart", 4, 16, 1);
    run<16>("2 addresses, lane % 2, 32 B apart", 16, 32, 1);
    run<8>("4 addresses, lane % 4, 32 B apart", 8, 32, 1);
    run<8>("8 addresses, lane % 8, 8 B apart", 4, 8, 1);
    return 0;
}
