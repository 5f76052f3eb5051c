// K3: deterministic DoF scatter.  The reference inserts every pair into per-Elem BTreeMaps and merges them serially
// (sparse_matrix.rs:68-120, linalg.rs:74-79); here every slot of the fixed pattern *gathers* its precomputed sources from
// the value buffer V in a fixed order: out = V[src1] (+ V[extra] ...).  No atomics, no zero-fill, writes are contiguous.
// A key's first contribution is stored as-is (not added to 0.0) exactly like BTreeMap::insert, so signed zeros survive.
//
// Source map encoding: src1[slot] < 2^31 -> index of the single contribution in V.  Top bit set -> the key has more than one
// contribution (shared edges / RBS inter-layer overlaps; at most 2 in practice): the low 31 bits index the first entry of the
// slot's run in the extras arrays, extra_first[k] is the first contribution, extra_src[k..] the following ones in plan order.
#include <cuda_runtime.h>

#include <algorithm>

#include "device_plan.hpp"

namespace fem2d {
namespace {

struct K3Args {
    const uint32_t* src1; const uint32_t* extra_slot; const uint32_t* extra_src; const uint32_t* extra_first; const double2* V;
    double* a; double* b;
    unsigned long long n_extra;
    unsigned long long begin[MAX_SLOT_RANGES], end[MAX_SLOT_RANGES], first_pair[MAX_SLOT_RANGES + 1];   // slot ranges of this launch
    uint32_t n_ranges;
    int vec_ok, selA, selB;
};

__device__ __forceinline__ double2 k3_value(const K3Args& g, unsigned long long slot, uint32_t s) {
    if (!(s & 0x80000000u)) return __ldg(&g.V[s]);
    unsigned long long k = s & 0x7fffffffu;
    double2 v = __ldg(&g.V[g.extra_first[k]]);
    do {   // *current_value += value (sparse_matrix.rs:61,115), sequentially in the plan's fixed order
        const double2 e = __ldg(&g.V[g.extra_src[k]]);
        v.x = v.x + e.x; v.y = v.y + e.y;
        k++;
    } while (k < g.n_extra && g.extra_slot[k] == slot);
    return v;
}

__global__ void __launch_bounds__(256) k3_gather_kernel(const K3Args g) {
    // two slots per thread, pairs aligned to even slot indices -> 16-byte stores; the launch covers up to MAX_SLOT_RANGES ranges
    const unsigned long long pair = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    uint32_t r = 0;
    while (r + 1 < g.n_ranges && pair >= g.first_pair[r + 1]) r++;
    const unsigned long long begin = g.begin[r], end = g.end[r];
    const unsigned long long s = (begin & ~1ull) + 2ull * (pair - g.first_pair[r]);
    if (s >= end) return;
    if (g.vec_ok && s >= begin && s + 1 < end) {
        const uint2 src = *reinterpret_cast<const uint2*>(g.src1 + s);
        const double2 v0 = k3_value(g, s, src.x), v1 = k3_value(g, s + 1, src.y);
        *reinterpret_cast<double2*>(g.a + s) = make_double2(g.selA ? v0.y : v0.x, g.selA ? v1.y : v1.x);
        *reinterpret_cast<double2*>(g.b + s) = make_double2(g.selB ? v0.y : v0.x, g.selB ? v1.y : v1.x);
    } else {
        for (unsigned long long k = s; k < s + 2; k++) {
            if (k < begin || k >= end) continue;
            const double2 v = k3_value(g, k, g.src1[k]);
            g.a[k] = g.selA ? v.y : v.x; g.b[k] = g.selB ? v.y : v.x;
        }
    }
}

}  // namespace

cudaError_t launch_k3_scatter(const Plan& P, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, double* d_a, double* d_b, int selA, int selB,
                              cudaStream_t st, uint32_t* launches) {
    K3Args g{};
    g.src1 = P.d_src1; g.extra_slot = P.d_extra_slot; g.extra_src = P.d_extra_src; g.extra_first = P.d_extra_first; g.V = P.d_V;
    g.a = d_a; g.b = d_b; g.n_extra = P.n_extra;
    g.vec_ok = ((((uintptr_t)d_a) | ((uintptr_t)d_b)) & 15u) == 0; g.selA = selA; g.selB = selB;
    unsigned long long pairs = 0;
    for (uint32_t k = 0; k < n_ranges && g.n_ranges < MAX_SLOT_RANGES; k++) {
        const uint64_t b = begins[k], e = std::min<uint64_t>(ends[k], P.nnz);
        if (b >= e) continue;
        g.begin[g.n_ranges] = b; g.end[g.n_ranges] = e; g.first_pair[g.n_ranges] = pairs;
        pairs += (e - (b & ~1ull) + 1) / 2;
        g.n_ranges++;
    }
    g.first_pair[g.n_ranges] = pairs;
    if (pairs == 0) return cudaSuccess;
    k3_gather_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(g);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace fem2d
