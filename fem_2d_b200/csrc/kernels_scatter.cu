// K3: deterministic DoF scatter.  The reference inserts every pair into per-Elem BTreeMaps and merges them serially
// (sparse_matrix.rs:68-120, linalg.rs:74-79); here every slot of the fixed pattern *gathers* its precomputed sources from
// the value buffer V in a fixed order: out = V[src1] (+ V[extra] ...).  No atomics, no zero-fill, writes are contiguous.
// A key's first contribution is stored as-is (not added to 0.0) exactly like BTreeMap::insert, so signed zeros survive.
//
// Source map: src1[slot], read here in its packed form (device_plan.cu: 16-bit offsets from a per-chunk base, plain 32-bit for the
// chunks that do not fit).  Encoding of a source: src1[slot] < 2^31 -> index of the single contribution in V.  Top bit set -> the key has more than one
// contribution (shared edges / RBS inter-layer overlaps; at most 2 in practice): the low 31 bits index the first entry of the
// slot's run in the extras arrays, extra_first[k] is the first contribution, extra_src[k..] the following ones in plan order.
#include <cuda_runtime.h>

#include <algorithm>

#include "device_plan.hpp"

namespace fem2d {
namespace {

struct K3Args {
    const uint32_t* chunk_base; const uint16_t* src16; const uint32_t* src1;   // packed source map + plain form (device_plan.cu)
    const uint32_t* extra_slot; const uint32_t* extra_src; const uint32_t* extra_first; const double2* V;
    double* a; double* b;
    unsigned long long n_extra, n_chunks;
    unsigned long long begin[MAX_SLOT_RANGES], end[MAX_SLOT_RANGES];   // slot ranges of this launch
    uint32_t first_block[MAX_SLOT_RANGES + 1];                         // CTA index at which each range starts
    uint32_t n_ranges;
    int selA, selB;
};

// V is written by the integrator grid that, under programmatic dependent launch, can still be running when this grid starts:
// read it with plain loads after cudaGridDependencySynchronize() (L1 is clean at grid start and no V line is touched before the
// wait), never through the non-coherent path, whose read-only contract would not hold.
__device__ __forceinline__ double2 ld_v(const double2* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double2 k3_value(const K3Args& g, unsigned long long slot, uint32_t s) {
    if (!(s & 0x80000000u)) return ld_v(&g.V[s]);
    unsigned long long k = s & 0x7fffffffu;
    double2 v = ld_v(&g.V[g.extra_first[k]]);
    do {   // *current_value += value (sparse_matrix.rs:61,115), sequentially in the plan's fixed order
        const double2 e = ld_v(&g.V[g.extra_src[k]]);
        v.x = v.x + e.x; v.y = v.y + e.y;
        k++;
    } while (k < g.n_extra && g.extra_slot[k] == slot);
    return v;
}

// One CTA per K3_BLOCK_SLOTS-aligned block of slots, K3_ITERS slots per thread, one slot per lane and iteration: a warp's 32
// lanes take 32 consecutive slots, read their 16-bit source offsets (plus the chunk base, a warp-wide broadcast), gather 32
// entries of V -- consecutive 16-byte entries where the pattern row walks along a row of a value tile, i.e. one coalesced
// 512-byte read -- and write 256 contiguous bytes to each value array.  No shared memory, no barriers; the K3_ITERS iterations
// of a thread are independent two-step chains (offset -> V -> store) that the unrolled code keeps in flight together.
// What was measured on B200 to get here (scripts/micro/write_bw.cu, 57.6 M slots): two adjacent slots per lane (16-byte stores)
// halve the sector efficiency of the gathers, 4.1 vs 6.5 TB/s; the dependent index load in front of the gather costs 142 -> 173 us
// at 4 slots per thread (245 us at 1), and an L2 prefetch of the index stream ~1 M slots ahead takes back 13 us of that.
// Slots of plain chunks read src1 instead; the rare multi-contribution slots (shared edges, RBS overlaps; 1.2 % at 1 M DoFs)
// are finished after the streaming part so their dependent loads stay out of it.
__global__ void __launch_bounds__(K3_THREADS) k3_gather_kernel(const K3Args g) {
    static_assert((K3_ITERS * 32) % SRC_CHUNK == 0 && K3_BLOCK_SLOTS == K3_THREADS * K3_ITERS, "block geometry");
    uint32_t r = 0;
    while (r + 1 < g.n_ranges && blockIdx.x >= g.first_block[r + 1]) r++;
    const unsigned long long begin = g.begin[r], end = g.end[r];
    const uint32_t lane = threadIdx.x % 32;
    // warp w covers slots [w * 32 * K3_ITERS, (w + 1) * 32 * K3_ITERS) of the block
    const unsigned long long wbase = (begin / K3_BLOCK_SLOTS + (blockIdx.x - g.first_block[r])) * K3_BLOCK_SLOTS + (threadIdx.x / 32) * (32 * K3_ITERS);
    {   // pull the offsets that the warps K3_PREFETCH_SLOTS further on will need into L2 (one 128-byte line = 64 offsets)
        const unsigned long long p = wbase + K3_PREFETCH_SLOTS + lane * 64ull;
        if (lane < K3_ITERS * 32 / 64 && p < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.src16 + p));
    }
    const unsigned long long base = wbase + lane;
    const unsigned long long chunk0 = wbase / SRC_CHUNK;
    uint32_t cb[K3_ITERS * 32 / SRC_CHUNK];
#pragma unroll
    for (uint32_t c = 0; c < K3_ITERS * 32 / SRC_CHUNK; c++) cb[c] = chunk0 + c < g.n_chunks ? __ldg(&g.chunk_base[chunk0 + c]) : 0u;
    uint32_t src[K3_ITERS]; bool ok[K3_ITERS];
#pragma unroll
    for (uint32_t it = 0; it < K3_ITERS; it++) {
        const unsigned long long s = base + it * 32;
        ok[it] = s >= begin && s < end;
        src[it] = ok[it] ? (uint32_t)__ldg(&g.src16[s]) : 0u;
    }
#pragma unroll
    for (uint32_t it = 0; it < K3_ITERS; it++) {
        const uint32_t b = cb[it * 32 / SRC_CHUNK];
        src[it] += b;
        if (b == SRC_CHUNK_PLAIN) src[it] = ok[it] ? __ldg(&g.src1[base + it * 32]) : 0u;
    }
    // programmatic dependent launch: everything above only reads the plan's source map and may overlap the integrator's tail;
    // V is complete once the preceding grid is
    cudaGridDependencySynchronize();
    double2 v[K3_ITERS];
#pragma unroll
    for (uint32_t it = 0; it < K3_ITERS; it++) v[it] = ld_v(&g.V[(src[it] & 0x80000000u) ? 0u : src[it]]);
    bool any = false;
#pragma unroll
    for (uint32_t it = 0; it < K3_ITERS; it++) {
        const unsigned long long s = base + it * 32;
        const bool multi = (src[it] & 0x80000000u) != 0;
        if (ok[it] && !multi) { g.a[s] = g.selA ? v[it].y : v[it].x; g.b[s] = g.selB ? v[it].y : v[it].x; }
        any |= ok[it] && multi;
    }
    if (!any) return;
#pragma unroll
    for (uint32_t it = 0; it < K3_ITERS; it++) {
        if (!(ok[it] && (src[it] & 0x80000000u))) continue;
        const unsigned long long s = base + it * 32;
        const double2 w = k3_value(g, s, src[it]);
        g.a[s] = g.selA ? w.y : w.x; g.b[s] = g.selB ? w.y : w.x;
    }
}

}  // namespace

cudaError_t launch_k3_scatter(const Plan& P, uint32_t n_ranges, const uint64_t* begins, const uint64_t* ends, double* d_a, double* d_b, int selA, int selB,
                              cudaStream_t st, uint32_t* launches) {
    K3Args g{};
    g.chunk_base = P.d_chunk_base; g.src16 = P.d_src16; g.src1 = P.d_src1;
    g.extra_slot = P.d_extra_slot; g.extra_src = P.d_extra_src; g.extra_first = P.d_extra_first; g.V = P.d_V;
    g.a = d_a; g.b = d_b; g.n_extra = P.n_extra; g.n_chunks = (P.nnz + SRC_CHUNK - 1) / SRC_CHUNK;
    g.selA = selA; g.selB = selB;
    uint64_t blocks = 0;
    for (uint32_t k = 0; k < n_ranges && g.n_ranges < MAX_SLOT_RANGES; k++) {
        const uint64_t b = begins[k], e = std::min<uint64_t>(ends[k], P.nnz);
        if (b >= e) continue;
        g.begin[g.n_ranges] = b; g.end[g.n_ranges] = e; g.first_block[g.n_ranges] = (uint32_t)blocks;
        blocks += (e - 1) / K3_BLOCK_SLOTS - b / K3_BLOCK_SLOTS + 1;
        g.n_ranges++;
    }
    g.first_block[g.n_ranges] = (uint32_t)blocks;
    if (blocks == 0) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(K3_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k3_gather_kernel, g);
    if (launches) (*launches)++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace fem2d
