// K3: deterministic DoF scatter.  The reference inserts every pair into per-Elem BTreeMaps and merges them serially
// (sparse_matrix.rs:68-120, linalg.rs:74-79); here every slot of the fixed pattern *gathers* its precomputed sources from
// the value buffer V in a fixed order: out = V[src1] (+ V[extra] ...).  No atomics, no zero-fill, writes are contiguous.
// A key's first contribution is stored as-is (not added to 0.0) exactly like BTreeMap::insert, so signed zeros survive.
#include <cuda_runtime.h>

#include "device_plan.hpp"

namespace fem2d {
namespace {

__global__ void __launch_bounds__(256) k3_gather_kernel(const uint32_t* __restrict__ src1, const double2* __restrict__ V,
                                                        double* __restrict__ a, double* __restrict__ b,
                                                        unsigned long long begin, unsigned long long end, int vec_ok, int selA, int selB) {
    // two slots per thread, pairs aligned to even slot indices -> 16-byte stores
    const unsigned long long s = (begin & ~1ull) + 2ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
    if (s >= end) return;
    if (vec_ok && s >= begin && s + 1 < end) {
        const uint2 src = *reinterpret_cast<const uint2*>(src1 + s);
        const double2 v0 = __ldg(&V[src.x]), v1 = __ldg(&V[src.y]);
        *reinterpret_cast<double2*>(a + s) = make_double2(selA ? v0.y : v0.x, selA ? v1.y : v1.x);
        *reinterpret_cast<double2*>(b + s) = make_double2(selB ? v0.y : v0.x, selB ? v1.y : v1.x);
    } else {
        for (unsigned long long k = s; k < s + 2; k++) {
            if (k < begin || k >= end) continue;
            const double2 v = __ldg(&V[src1[k]]);
            a[k] = selA ? v.y : v.x; b[k] = selB ? v.y : v.x;
        }
    }
}

// Keys with more than one contribution (shared edges / RBS inter-layer overlaps; at most 2 in practice): the head of each
// run adds the extra sources sequentially, in the plan's fixed order.
__global__ void k3_extras_kernel(const uint32_t* __restrict__ extra_slot, const uint32_t* __restrict__ extra_src, unsigned long long n_extra,
                                 const double2* __restrict__ V, double* __restrict__ a, double* __restrict__ b,
                                 unsigned long long begin, unsigned long long end, int selA, int selB) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n_extra) return;
    const uint32_t slot = extra_slot[i];
    if (slot < begin || slot >= end) return;
    if (i > 0 && extra_slot[i - 1] == slot) return;
    double va = a[slot], vb = b[slot];
    for (unsigned long long k = i; k < n_extra && extra_slot[k] == slot; k++) {
        const double2 v = V[extra_src[k]];
        va = va + (selA ? v.y : v.x); vb = vb + (selB ? v.y : v.x);      // *current_value += value (sparse_matrix.rs:61,115)
    }
    a[slot] = va; b[slot] = vb;
}

}  // namespace

cudaError_t launch_k3_scatter(const Plan& P, uint64_t slot_begin, uint64_t slot_end, double* d_a, double* d_b, int selA, int selB, cudaStream_t st, uint32_t* launches) {
    if (slot_end > P.nnz) slot_end = P.nnz;
    if (slot_begin >= slot_end) return cudaSuccess;
    const unsigned long long n2 = (slot_end - (slot_begin & ~1ull) + 1) / 2;
    const int vec_ok = ((((uintptr_t)d_a) | ((uintptr_t)d_b)) & 15u) == 0;
    k3_gather_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(P.d_src1, P.d_V, d_a, d_b, slot_begin, slot_end, vec_ok, selA, selB);
    if (launches) (*launches)++;
    if (P.n_extra) {
        k3_extras_kernel<<<(unsigned)((P.n_extra + 255) / 256), 256, 0, st>>>(P.d_extra_slot, P.d_extra_src, P.n_extra, P.d_V, d_a, d_b, slot_begin, slot_end, selA, selB);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}

}  // namespace fem2d
