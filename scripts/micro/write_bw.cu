// Microbenchmark: what a write-dominated kernel of K3's shape can reach on this GPU (ceilings for the scatter kernel's roofline).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/write_bw scripts/micro/write_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void w16(double* a, double* b, size_t n) {   // 2 slots per thread, 16-B stores to two arrays
    size_t s = 2 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    if (s + 1 < n) {
        *reinterpret_cast<double2*>(a + s) = make_double2(1.0 + s, 2.0);
        *reinterpret_cast<double2*>(b + s) = make_double2(3.0, 4.0 + s);
    }
}
__global__ void w8(double* a, double* b, size_t n) {    // 1 slot per thread, 8-B stores
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s < n) { a[s] = 1.0 + s; b[s] = 3.0 + s; }
}
__global__ void w32x4(double* a, double* b, size_t n) { // 4 iterations per thread, 16-B stores, block-strided (K3-RLE shape)
    size_t base = blockIdx.x * (size_t)2048;
    for (int it = 0; it < 4; it++) {
        size_t s = base + it * 512 + 2 * threadIdx.x;
        if (s + 1 < n) {
            *reinterpret_cast<double2*>(a + s) = make_double2(1.0 + s, 2.0);
            *reinterpret_cast<double2*>(b + s) = make_double2(3.0, 4.0 + s);
        }
    }
}
// gather from a small L2-resident table (16 B per slot) + write: K3 with a perfect index stream
__global__ void g16(const double2* __restrict__ V, uint32_t vmask, double* a, double* b, size_t n) {
    size_t s = 2 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    if (s + 1 < n) {
        const uint32_t i0 = (uint32_t)(s * 7) & vmask;
        const double2 v0 = __ldg(&V[i0]), v1 = __ldg(&V[i0 + 1]);
        *reinterpret_cast<double2*>(a + s) = make_double2(v0.x, v1.x);
        *reinterpret_cast<double2*>(b + s) = make_double2(v0.y, v1.y);
    }
}
// same, one slot per lane (coalesced 16-B gathers, 8-B stores)
__global__ void g8(const double2* __restrict__ V, uint32_t vmask, double* a, double* b, size_t n) {
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s < n) {
        const uint32_t i0 = ((uint32_t)((s & ~31ull) * 7) & vmask) + (uint32_t)(s & 31);
        const double2 v0 = __ldg(&V[i0]);
        a[s] = v0.x; b[s] = v0.y;
    }
}
// g8 + a dependent 16-bit index stream in front of the gather (the packed source map)
__global__ void g8i(const double2* __restrict__ V, const uint16_t* __restrict__ idx, const uint32_t* __restrict__ cb, double* a, double* b, size_t n) {
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s < n) {
        const uint32_t i0 = cb[s / 64] + idx[s];
        const double2 v0 = __ldg(&V[i0]);
        a[s] = v0.x; b[s] = v0.y;
    }
}
template <int IT>
__global__ void g8iN(const double2* __restrict__ V, const uint16_t* __restrict__ idx, const uint32_t* __restrict__ cb, double* a, double* b, size_t n) {
    size_t s0 = blockIdx.x * (size_t)blockDim.x * IT + (threadIdx.x / 32) * 32 * IT + threadIdx.x % 32;
    uint32_t i0[IT]; double2 v[IT];
#pragma unroll
    for (int k = 0; k < IT; k++) { size_t s = s0 + k * 32; i0[k] = s < n ? cb[s / 64] + idx[s] : 0; }
#pragma unroll
    for (int k = 0; k < IT; k++) v[k] = __ldg(&V[i0[k]]);
#pragma unroll
    for (int k = 0; k < IT; k++) { size_t s = s0 + k * 32; if (s < n) { a[s] = v[k].x; b[s] = v[k].y; } }
}
// same + L2 prefetch of the index stream PF slots ahead
template <int IT>
__global__ void g8iNp(const double2* __restrict__ V, const uint16_t* __restrict__ idx, const uint32_t* __restrict__ cb, double* a, double* b, size_t n, size_t pf) {
    size_t s0 = blockIdx.x * (size_t)blockDim.x * IT + (threadIdx.x / 32) * 32 * IT + threadIdx.x % 32;
    {   // the warp covers 32*IT slots = IT/2 lines of idx; lanes < IT/2 (at least one) prefetch them
        const size_t w0 = s0 - threadIdx.x % 32 + pf;
        const unsigned lane = threadIdx.x % 32;
        if (lane < (IT + 1) / 2 && w0 + lane * 64 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(idx + w0 + lane * 64));
    }
    uint32_t i0[IT]; double2 v[IT];
#pragma unroll
    for (int k = 0; k < IT; k++) { size_t s = s0 + k * 32; i0[k] = s < n ? cb[s / 64] + idx[s] : 0; }
#pragma unroll
    for (int k = 0; k < IT; k++) v[k] = __ldg(&V[i0[k]]);
#pragma unroll
    for (int k = 0; k < IT; k++) { size_t s = s0 + k * 32; if (s < n) { a[s] = v[k].x; b[s] = v[k].y; } }
}
__global__ void fill_idx(uint16_t* idx, uint32_t* cb, size_t n) {
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s < n) { idx[s] = (uint16_t)(s & 63); if ((s & 63) == 0) cb[s / 64] = (uint32_t)((s / 64) * 7919u) & 0xffffu; }
}
__global__ void copyk(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s < n2) out[s] = in[s];
}

template <class F> float timeit(F f, int reps = 20) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; i++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}

int main() {
    const size_t n = 57557904;
    double *a, *b; double2* V; double2* c;
    cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&c, n * 16);
    const uint32_t vn = 1u << 17; cudaMalloc(&V, (vn + 64) * 16); cudaMemset(V, 0, (vn + 64) * 16);
    const double gb = 16.0 * n / 1e9;
    auto rep = [&](const char* name, float ms, double bytes_gb) { printf("%-28s %8.3f us  %8.1f GB/s\n", name, ms * 1e3, bytes_gb / (ms * 1e-3)); };
    rep("cudaMemset a+b", timeit([&] { cudaMemsetAsync(a, 0, n * 8); cudaMemsetAsync(b, 0, n * 8); }), gb);
    rep("w16 (2 slots/thread)", timeit([&] { w16<<<(unsigned)((n / 2 + 255) / 256), 256>>>(a, b, n); }), gb);
    rep("w8 (1 slot/thread)", timeit([&] { w8<<<(unsigned)((n + 255) / 256), 256>>>(a, b, n); }), gb);
    rep("w32x4 (2048-slot blocks)", timeit([&] { w32x4<<<(unsigned)((n + 2047) / 2048), 256>>>(a, b, n); }), gb);
    rep("g16 gather+write", timeit([&] { g16<<<(unsigned)((n / 2 + 255) / 256), 256>>>(V, vn - 1, a, b, n); }), gb);
    rep("g8 gather+write", timeit([&] { g8<<<(unsigned)((n + 255) / 256), 256>>>(V, vn - 1, a, b, n); }), gb);
    uint16_t* idx; uint32_t* cb; cudaMalloc(&idx, n * 2 + 64); cudaMalloc(&cb, (n / 64 + 2) * 4);
    fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, cb, n);
    rep("g8i index+gather+write", timeit([&] { g8i<<<(unsigned)((n + 255) / 256), 256>>>(V, idx, cb, a, b, n); }), gb);
    rep("g8i x2", timeit([&] { g8iN<2><<<(unsigned)((n + 511) / 512), 256>>>(V, idx, cb, a, b, n); }), gb);
    rep("g8i x4", timeit([&] { g8iN<4><<<(unsigned)((n + 1023) / 1024), 256>>>(V, idx, cb, a, b, n); }), gb);
    rep("g8i x8", timeit([&] { g8iN<8><<<(unsigned)((n + 2047) / 2048), 256>>>(V, idx, cb, a, b, n); }), gb);
    rep("g8i x4 128thr", timeit([&] { g8iN<4><<<(unsigned)((n + 511) / 512), 128>>>(V, idx, cb, a, b, n); }), gb);
    for (size_t pf : {(size_t)1 << 18, (size_t)1 << 20, (size_t)1 << 22, (size_t)1 << 24}) {
        char nm[64]; snprintf(nm, 64, "g8i x4 +L2 prefetch %zuK", pf >> 10);
        rep(nm, timeit([&] { g8iNp<4><<<(unsigned)((n + 1023) / 1024), 256>>>(V, idx, cb, a, b, n, pf); }), gb);
        snprintf(nm, 64, "g8i x2 +L2 prefetch %zuK", pf >> 10);
        rep(nm, timeit([&] { g8iNp<2><<<(unsigned)((n + 511) / 512), 256>>>(V, idx, cb, a, b, n, pf); }), gb);
        snprintf(nm, 64, "g8i x8 +L2 prefetch %zuK", pf >> 10);
        rep(nm, timeit([&] { g8iNp<8><<<(unsigned)((n + 2047) / 2048), 256>>>(V, idx, cb, a, b, n, pf); }), gb);
    }
    rep("copy 921MB->921MB (R+W)", timeit([&] { copyk<<<(unsigned)((n + 255) / 256), 256>>>((const double2*)c, (double2*)a == nullptr ? c : c, 0); copyk<<<(unsigned)((n / 2 + 255) / 256), 256>>>((const double2*)a, (double2*)b, n / 2); }), 2 * 8.0 * n / 1e9);
    rep("cudaMemcpy D2D a->b", timeit([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); }), 2 * 8.0 * n / 1e9);
    return 0;
}
