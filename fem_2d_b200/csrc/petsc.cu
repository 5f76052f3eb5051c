// SparseMatrix -> PETSc AIJ binary image, built on the device from the plan's upper-triangular pattern and a device value array.
// Replaces `impl From<SparseMatrix> for AIJMatrixBinary` + `AIJMatrixBinary::print_to_petsc_binary_file`
// (/root/reference/src/fem_problem/linalg/sparse_matrix.rs:184-264, called by GEP::print_to_petsc_binary_files, linalg.rs:44-52):
//   header   00 12 7B 50 (the literal b"\0\x12{P" = 1211216), then u32 BE rows, cols, nnz
//   i        per-row entry counts of the FULL symmetric matrix, u32 BE          (sparse_matrix.rs:187-197)
//   j        column ids, rows in order, columns ascending inside a row, u32 BE  (sparse_matrix.rs:200-212)
//   a        values in the same order, f64 BE
// Full row r = the mirrored entries (r', r), r' < r, in ascending r' (the column r of the upper triangle), then the row's own upper
// part (r, c >= r) in ascending c.  The upper part sits at known positions (CSR offsets of the pattern); the mirrored part needs the
// transpose of the strictly upper pattern, a stable sort of (column, row) keys that depends on the pattern only: it is built once
// per plan and reused for A and B.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/fem2d.h"
#include "device_plan.hpp"

namespace fem2d {
namespace {

__global__ void lower_keys_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cols, uint32_t nnz, unsigned long long* __restrict__ keys,
                                  uint32_t* __restrict__ slots, uint32_t* __restrict__ lower_cnt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const uint32_t r = rows[k], c = cols[k];
    // diagonal entries have no mirror image: their key sorts behind every real one and is cut off by the count
    keys[k] = r == c ? ~0ull : ((unsigned long long)c << 32 | r);
    slots[k] = k;
    if (r != c) atomicAdd(&lower_cnt[c], 1u);
}

__global__ void full_counts_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ lower_cnt, uint32_t n, uint32_t* __restrict__ counts) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) counts[r] = (row_ptr[r + 1] - row_ptr[r]) + lower_cnt[r];
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// One thread per entry of the full matrix, by source: thread k < nnz handles the upper entry of slot k; thread nnz + q handles the
// q-th mirrored entry (sorted by (column, row)).  Writes j (u32 BE) and a (f64 BE) at the entry's position in the image.
__global__ void aij_fill_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cols, const uint32_t* __restrict__ row_ptr,
                                const uint32_t* __restrict__ lower_cnt, const uint32_t* __restrict__ full_start, const uint32_t* __restrict__ lower_start,
                                const uint32_t* __restrict__ lower_slot, uint32_t nnz, uint32_t n_lower, const double* __restrict__ vals,
                                uint32_t* __restrict__ j_out, uint2* __restrict__ a_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz + n_lower) return;
    uint32_t pos, col, slot;
    if (t < nnz) {                       // (r, c): behind the mirrored part of row r
        slot = t;
        const uint32_t r = rows[t];
        pos = full_start[r] + lower_cnt[r] + (t - row_ptr[r]);
        col = cols[t];
    } else {                             // (c, r) mirrored from slot (r, c): q-th mirrored entry overall, rank inside row c = q - lower_start[c]
        const uint32_t q = t - nnz;
        slot = lower_slot[q];
        const uint32_t c = cols[slot];
        pos = full_start[c] + (q - lower_start[c]);
        col = rows[slot];
    }
    j_out[pos] = bswap32(col);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[slot]);
    a_out[pos] = make_uint2(bswap32((uint32_t)(bits >> 32)), bswap32((uint32_t)bits));
}

__global__ void aij_head_kernel(const uint32_t* __restrict__ counts, uint32_t n, uint32_t nnz_full, uint32_t* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { out[0] = bswap32(1211216u); out[1] = bswap32(n); out[2] = bswap32(n); out[3] = bswap32(nnz_full); }
    if (r < n) out[4 + r] = bswap32(counts[r]);
}

}  // namespace

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return FEM2D_ERR_CUDA; } } while (0)

// Transpose of the strictly upper pattern + full-row offsets (pattern-only data, cached in the plan).
int device_petsc_prepare(Plan& P, std::string& err) {
    if (P.d_aij_arena) return FEM2D_OK;
    CKP(cudaSetDevice(P.device));
    const uint32_t n = P.host.n_dofs, nnz = (uint32_t)P.nnz;
    if (2ull * P.nnz >= (1ull << 32)) { err = "full matrix has more than 2^32 entries"; return FEM2D_ERR_UNSUPPORTED; }
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    const size_t o_lcnt = take(((size_t)n + 1) * 4), o_lstart = take(((size_t)n + 1) * 4), o_cnt = take(((size_t)n + 1) * 4), o_fstart = take(((size_t)n + 1) * 4);
    const size_t o_lslot = take((size_t)nnz * 4);
    void* arena = nullptr;
    CKP(dev_malloc(&arena, o));
    char* base = (char*)arena;
    uint32_t *d_lcnt = (uint32_t*)(base + o_lcnt), *d_lstart = (uint32_t*)(base + o_lstart), *d_cnt = (uint32_t*)(base + o_cnt), *d_fstart = (uint32_t*)(base + o_fstart);
    uint32_t* d_lslot = (uint32_t*)(base + o_lslot);
    // scratch: keys (double buffered), slots (second buffer), cub temp
    int bits = 1; while ((1ull << bits) < (unsigned long long)n) bits++;
    size_t temp_sort = 0, temp_scan = 0;
    {
        cub::DoubleBuffer<unsigned long long> kq(nullptr, nullptr); cub::DoubleBuffer<uint32_t> vq(nullptr, nullptr);
        CKP(cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, kq, vq, (int)std::max(nnz, 1u), 0, 64));
        CKP(cub::DeviceScan::ExclusiveSum(nullptr, temp_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n + 1)));
    }
    size_t temp = std::max(temp_sort, temp_scan);
    void *d_k1 = nullptr, *d_k2 = nullptr, *d_s2 = nullptr, *d_temp = nullptr;
    cudaError_t e = dev_malloc(&d_k1, (size_t)std::max(nnz, 1u) * 8);
    if (e == cudaSuccess) e = dev_malloc(&d_k2, (size_t)std::max(nnz, 1u) * 8);
    if (e == cudaSuccess) e = dev_malloc(&d_s2, (size_t)std::max(nnz, 1u) * 4);
    if (e == cudaSuccess) e = dev_malloc(&d_temp, temp);
    uint32_t h_tail[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemsetAsync(d_lcnt, 0, ((size_t)n + 1) * 4, nullptr);
    if (e == cudaSuccess && nnz) {
        lower_keys_kernel<<<(nnz + 255) / 256, 256>>>(P.d_rows, P.d_cols, nnz, (unsigned long long*)d_k1, d_lslot, d_lcnt);
        e = cudaGetLastError();
        cub::DoubleBuffer<unsigned long long> kb((unsigned long long*)d_k1, (unsigned long long*)d_k2);
        cub::DoubleBuffer<uint32_t> vb(d_lslot, (uint32_t*)d_s2);
        // (column, row) keys: sort on the significant bits of the row, then of the column (stable LSD passes); the all-ones keys of the
        // diagonal entries need the top bits too, so the second pass runs to bit 64
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(d_temp, temp, kb, vb, (int)nnz, 0, bits);
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(d_temp, temp, kb, vb, (int)nnz, 32, 64);
        if (e == cudaSuccess && vb.Current() != d_lslot) e = cudaMemcpyAsync(d_lslot, vb.Current(), (size_t)nnz * 4, cudaMemcpyDeviceToDevice, nullptr);
    }
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_temp, temp, d_lcnt, d_lstart, (int)(n + 1));
    if (e == cudaSuccess) { full_counts_kernel<<<(n + 255) / 256, 256>>>(P.d_row_ptr, d_lcnt, n, d_cnt); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt + n, 0, 4, nullptr);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_temp, temp, d_cnt, d_fstart, (int)(n + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_tail[0], d_lstart + n, 4, cudaMemcpyDeviceToHost, nullptr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_tail[1], d_fstart + n, 4, cudaMemcpyDeviceToHost, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    dev_free(d_k1); dev_free(d_k2); dev_free(d_s2); dev_free(d_temp);
    if (e != cudaSuccess) { dev_free(arena); err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    P.d_aij_arena = arena; P.d_aij_lower_cnt = d_lcnt; P.d_aij_lower_start = d_lstart; P.d_aij_counts = d_cnt; P.d_aij_full_start = d_fstart; P.d_aij_lower_slot = d_lslot;
    P.aij_n_lower = h_tail[0]; P.aij_nnz_full = h_tail[1];
    return FEM2D_OK;
}

// The finished byte image in a device buffer (header | counts | j | a); *d_image is owned by the caller (dev_free).
int device_petsc_image(Plan& P, const double* d_vals, void** d_image, uint64_t* bytes, std::string& err) {
    int st = device_petsc_prepare(P, err);
    if (st != FEM2D_OK) return st;
    const uint32_t n = P.host.n_dofs, nnz = (uint32_t)P.nnz;
    const uint64_t nf = P.aij_nnz_full;
    const uint64_t total = 16 + 4ull * n + 12ull * nf;
    void* img = nullptr;
    CKP(dev_malloc(&img, total + 16));
    uint32_t* head = (uint32_t*)img;
    uint32_t* j_out = head + 4 + n;
    aij_head_kernel<<<(n + 256) / 256, 256>>>(P.d_aij_counts, n, (uint32_t)nf, head);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && nf) {
        // the value block starts at 16 + 4 n + 4 nf bytes, 8-byte aligned iff n + nf is even: the fill kernel stores uint2, so an odd
        // n + nf goes through an aligned scratch block and one device-to-device copy
        char* a_base = (char*)img + 16 + 4ull * n + 4ull * nf;
        const bool aligned = ((uintptr_t)a_base & 7u) == 0;
        uint2* a_dst = (uint2*)a_base;
        void* tmp = nullptr;
        if (!aligned) { e = dev_malloc(&tmp, 8ull * nf); a_dst = (uint2*)tmp; }
        if (e == cudaSuccess) {
            const uint64_t threads = (uint64_t)nnz + P.aij_n_lower;
            aij_fill_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(P.d_rows, P.d_cols, P.d_row_ptr, P.d_aij_lower_cnt, P.d_aij_full_start, P.d_aij_lower_start,
                                                                        P.d_aij_lower_slot, nnz, P.aij_n_lower, d_vals, j_out, a_dst);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && !aligned) e = cudaMemcpyAsync(a_base, tmp, 8ull * nf, cudaMemcpyDeviceToDevice, nullptr);
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
        dev_free(tmp);
    }
    if (e != cudaSuccess) { dev_free(img); err = cudaGetErrorString(e); return FEM2D_ERR_CUDA; }
    *d_image = img; *bytes = total;
    return FEM2D_OK;
}

}  // namespace fem2d
