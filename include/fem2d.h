/* =====================================================================================
 * fem2d.h -- C-ABI of the B200-native Galerkin assembly path.
 *
 * Drop-in boundary for the reference's
 *   galerkin_sample_gep_hcurl::<BSpace, CurlCurl, L2Inner>(&domain, Option<[usize;2]>)
 *       -> Result<GEP, GalerkinSamplingError>            (src/fem_problem/galerkin.rs:33-40)
 * The reference has no FFI of its own (pure Rust); these entry points are what a thin
 * `fem_2d-sys` shim inside `fem_problem::galerkin` would bind (INTEGRATION.md shows it).
 * Plain pointers and sizes only; no C++/torch types; never unwinds across the boundary.
 *
 * Everything below runs on the GPU (sm_100a).  There is no CPU fallback: without a CUDA
 * device every numeric entry point returns FEM2D_ERR_NO_DEVICE.
 * ===================================================================================== */
#ifndef FEM2D_H
#define FEM2D_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  1..3 mirror GalerkinSamplingError (galerkin.rs:191-195) and are decided before
 * any compute, in the reference's order (galerkin.rs:42-59). */
enum {
    FEM2D_OK = 0,
    FEM2D_ERR_WRONG_CONTINUITY = 1, /* GalerkinSamplingError::WrongContinuityCondition */
    FEM2D_ERR_EMPTY_DOF_SET = 2,    /* GalerkinSamplingError::EmptyDOFSet */
    FEM2D_ERR_INVALID_GLQ = 3,      /* GalerkinSamplingError::InvalidGLQSettings (< MIN_GLQ_ORDER = 4, galerkin.rs:13) */
    FEM2D_ERR_BAD_ARGUMENT = 100,
    FEM2D_ERR_NO_DEVICE = 101,      /* no CUDA device / driver: the product path refuses to run */
    FEM2D_ERR_CUDA = 102,
    FEM2D_ERR_UNSUPPORTED = 103,    /* unknown basis / integral kind, order > table limit, nnz >= 2^31 ... */
    FEM2D_ERR_INTERNAL = 104,
    FEM2D_ERR_OUT_OF_MEMORY = 105
};

/* BSpace type parameter (basis.rs:19-44 implementors). */
enum {
    FEM2D_BASIS_HIER_POLY = 0,     /* HierPoly (hierarchical_basis_fns.rs:15) == the "KOLShapeFn" of older releases */
    FEM2D_BASIS_HIER_MAX_ORTHO = 1 /* HierMaxOrtho (hierarchical_basis_fns.rs:260), Legendre based, orders <= 12 */
};
/* AI / BI type parameters (integration.rs:59-90 implementors). */
enum { FEM2D_INTEGRAL_CURL_CURL = 0 /* integrals.rs:13 */, FEM2D_INTEGRAL_L2_INNER = 1 /* integrals.rs:279 */ };
/* ContinuityCondition (domain.rs:19-23). */
enum { FEM2D_CC_HCURL = 0, FEM2D_CC_HDIV = 1, FEM2D_CC_DISCONTINUOUS = 2 };

/* Numeric mode of the per-pair integrator.
 * EXACT replays the reference's floating-point operation order (no FMA contraction): values are bit-identical
 *   to the reference algorithm evaluated with the same GLQ nodes (the parity mode, default).
 * SUMFACT evaluates the separable closed form (1-D Gram matrices); DMMA contracts Phi^T W Phi on FP64 tensor-core
 *   tiles (mma.sync m8n8k4.f64).  Both are mathematically equal but re-ordered: they meet a scale-aware
 *   tolerance (1e-12 relative, floor 1e-14*max|M|), not the literal one. */
enum { FEM2D_MODE_EXACT = 0, FEM2D_MODE_SUMFACT = 1, FEM2D_MODE_DMMA = 2 };

/* HRefLoc codes (h_refinement.rs:211-229) used in fem2d_domain_view::elem_loc. */
enum { FEM2D_LOC_SW = 0, FEM2D_LOC_SE, FEM2D_LOC_NW, FEM2D_LOC_NE, FEM2D_LOC_W, FEM2D_LOC_E, FEM2D_LOC_S, FEM2D_LOC_N,
       FEM2D_LOC_BASE = 255 };

/* Read-only flattened view of a reference `Domain` (domain.rs:42-50).  All arrays are caller-owned HOST memory. */
typedef struct fem2d_domain_view {
    uint32_t n_elems;            /* mesh.elems.len() */
    uint32_t n_elements;         /* mesh.elements.len() */
    uint32_t n_dofs;             /* domain.dofs.len() */
    uint32_t continuity;         /* domain.cc as FEM2D_CC_* (galerkin.rs:42) */
    const uint32_t* elem_element; /* [n_elems] elem.element.id (elem.rs:105) */
    const int32_t* elem_parent;   /* [n_elems] elem.parent_id() or -1 (elem.rs:158) */
    const uint8_t* elem_loc;      /* [n_elems] last HRefLoc of elem.loc_stack() (elem.rs:109,165), FEM2D_LOC_BASE on the base layer */
    const double* element_p0;     /* [n_elements][2] element.points[0].{x,y} (element.rs:38-42) */
    const double* element_p3;     /* [n_elements][2] element.points[3].{x,y} */
    const double* element_eps_re; /* [n_elements] materials.eps_rel.re (integrals.rs:303) */
    const double* element_mu_re;  /* [n_elements] materials.mu_rel.re  (integrals.rs:37) */
    const uint32_t* bs_off;       /* [n_elems+1] CSR offsets into the per-Elem BasisSpec lists (domain.basis_specs, domain.rs:47) */
    const uint8_t* bs_i;          /* BasisSpec::i   (basis_spec.rs:13) */
    const uint8_t* bs_j;          /* BasisSpec::j   (basis_spec.rs:15) */
    const uint8_t* bs_dir;        /* BasisSpec::dir (0 = U, 1 = V)  (basis_spec.rs:17, :138-149) */
    const uint32_t* bs_dof;       /* BasisSpec::dof_id (basis_spec.rs:23) */
    uint32_t i_max, j_max;        /* mesh.max_expansion_orders() (galerkin.rs:65, mesh.rs:626) */
} fem2d_domain_view;

typedef struct fem2d_plan fem2d_plan; /* opaque: fixed sparsity pattern + scatter map + device buffers */

/* Symbolic phase: validates the view (status 1/2 as the reference), enumerates the (Elem, Elem-or-descendant) pair blocks
 * of galerkin.rs:91-178, groups bit-identical blocks into classes, and builds on `device` the sorted unique upper-
 * triangular key set (== BTreeMap<[u32;2]> iteration order, sparse_matrix.rs:16,48-58) plus the per-slot source map.
 * The plan is reusable across numeric calls on the same Domain. `dedupe` != 0 merges blocks whose inputs are
 * bit-identical (geometry, materials, basis-spec sets) so they are integrated once. */
int fem2d_symbolic(const fem2d_domain_view* view, int device, int dedupe, fem2d_plan** out);
void fem2d_plan_free(fem2d_plan* plan);

/* Plan queries. info[]: 0 nnz_upper, 1 n_pairs (galerkin.rs pair count), 2 n_blocks, 3 n_classes, 4 n_value_slots (V buffer
 * entries), 5 n_multi (keys with >1 contribution), 6 max contributions per key, 7 n_tables, 8 n_work_items, 9 n_dofs,
 * 10 n_lists, 11 n_extra (2nd+ contributions), 12 / 13 wall microseconds of the host / device halves of the symbolic phase,
 * 14 micro-tile height chosen for the exact integrator (4: throughput shape, 1: latency shape for small plans) */
int fem2d_plan_info(const fem2d_plan* plan, uint64_t info[16]);
/* Self-check of the exact integrator's work decomposition (host logic, works on host-only plans): every pair the pattern reads
 * (all (a, b) of a local-desc block, a <= b of a local-local block; galerkin.rs:91-127,138-178) lies in exactly one micro-tile of
 * its class, the tile numbering is invertible, every tile of every class is in exactly one work item, same-direction tiles precede
 * cross-direction ones in every item, and the staged function ranges of an item cover the functions its tiles read.
 * out[]: 0 tiles, 1 same-direction tiles, 2 thread slots (tiles + warp-alignment gaps), 3 violations found (0 = consistent). */
int fem2d_plan_check_work_items(const fem2d_plan* plan, uint64_t out[4]);
/* Work of one numeric call of the integrator, counted on the plan's classes (every class is integrated once per call): out[0] same-direction
 * pairs (U-U, V-V: A and B, 8 FP64 operations per pair and quadrature point + 4 per pair and quadrature row in the reference's order,
 * integrals.rs:36-91,302-353, glq.rs:19-32), out[1] cross-direction pairs (U-V, V-U: A only, 3 per point + 2 per row), out[2] / out[3] micro-tiles of
 * the two kinds, out[4] / out[5] pairs per micro-tile of the two kinds, out[6] thread slots of the warps that hold those micro-tiles (whole warps per pack / work item and kind),
 * out[7] staging warps per CTA of the persistent integrator chosen for this plan (1 or 2; 0: the plan runs the latency shape).
 * Host logic; works on host-only plans. */
int fem2d_plan_work_info(const fem2d_plan* plan, uint64_t out[8]);
/* Size of the per-slot source map the scatter kernel reads.  The map is packed per chunk of slots as 16-bit offsets from the
 * chunk's smallest source; chunks that do not fit keep plain 32-bit indices.  info[]: 0 chunks in plain form, 1 slots per chunk,
 * 2 bytes of the map one full scatter reads, 3 bytes of an all-plain map (4 per slot).  Device plans only. */
int fem2d_plan_source_map_info(const fem2d_plan* plan, uint64_t info[4]);
/* Copy the pattern to host: rows[k] <= cols[k], sorted by (row, col). */
int fem2d_plan_pattern(const fem2d_plan* plan, uint32_t* rows, uint32_t* cols);
/* CSR row offsets of the pattern: row_ptr[r] = first slot of row r, row_ptr[n_dofs] = nnz_upper (n_dofs + 1 entries). */
int fem2d_plan_row_offsets(fem2d_plan* plan, uint64_t* row_ptr);
/* What the host-output calls move over PCIe for the pattern instead of rows[] / cols[] (which they expand on host threads):
 * info[0] bytes of the CSR row offsets, info[1] number of runs of consecutive column ids, info[2] bytes of the column runs,
 * info[3] bytes of plain rows[] + cols[].  Builds the column runs on first use.  Device plans only. */
int fem2d_plan_pattern_transfer_info(fem2d_plan* plan, uint64_t info[4]);
/* Device pointers of the pattern (uint32 rows, cols; length nnz_upper). */
int fem2d_plan_pattern_device(const fem2d_plan* plan, const uint32_t** d_rows, const uint32_t** d_cols);

/* Concurrency: a plan supports ONE numeric call in flight on ONE stream at a time (its scratch -- tables, value buffer, restricted work-item
 * list -- is shared by all calls).  The first call with a new set of slot ranges rebuilds the restricted list and synchronises the device. */
/* Numeric phase (same replacement as fem2d_assemble below) into caller-provided DEVICE buffers d_a / d_b (nnz_upper doubles each, on the plan's device),
 * asynchronous on `stream` (a cudaStream_t, may be NULL).  GLQ nodes / weights are HOST inputs so the caller can pass
 * the exact values of gauss_quadrature_points (glq.rs:179-222, basis.rs:83-90).
 * slot_begin/slot_end restrict the scatter to a row-block slice [slot_begin, slot_end) of the pattern (multi-GPU
 * sharding, see fem2d_plan_row_blocks); pass 0 / UINT64_MAX for everything.  Values outside the slice are untouched. */
int fem2d_assemble_device(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode,
                          const double* u_pts, const double* u_w, uint32_t nu,
                          const double* v_pts, const double* v_w, uint32_t nv,
                          uint64_t slot_begin, uint64_t slot_end,
                          double* d_a, double* d_b, void* stream);

/* Same for up to 4 disjoint slot ranges in one call (one integrator launch restricted to what those slots read, one scatter
 * launch).  Multi-GPU use: a rank passes its block of the single-Elem rows and its block of the shared (edge-type) rows, see
 * fem2d_plan_row_blocks_split. */
int fem2d_assemble_device_ranges(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode,
                                 const double* u_pts, const double* u_w, uint32_t nu,
                                 const double* v_pts, const double* v_w, uint32_t nv,
                                 uint32_t n_ranges, const uint64_t* slot_begins, const uint64_t* slot_ends,
                                 double* d_a, double* d_b, void* stream);

/* Numeric phase with HOST outputs (a_vals / b_vals: nnz_upper doubles each; rows / cols may be NULL). Synchronous.
 * Replaces the Rayon loop over Elems with its per-pair AI::integrate / BI::integrate calls (galerkin.rs:73-184, integrals.rs:26-92,
 * 292-354) and the per-Elem insert_group + serial consume_matrix merge (sparse_matrix.rs:68-120, linalg.rs:59-81); the outputs are
 * what SparseMatrix::iter_upper_tri (sparse_matrix.rs:123-127) would yield for gep.a and gep.b. */
int fem2d_assemble(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode,
                   const double* u_pts, const double* u_w, uint32_t nu,
                   const double* v_pts, const double* v_w, uint32_t nv,
                   uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals);

/* Same, restricted to the row-block slice [slot_begin, slot_end): outputs hold slot_end - slot_begin entries. */
int fem2d_assemble_range(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode,
                         const double* u_pts, const double* u_w, uint32_t nu,
                         const double* v_pts, const double* v_w, uint32_t nv,
                         uint64_t slot_begin, uint64_t slot_end,
                         uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals);

/* Same for up to 4 slot ranges; the outputs hold the ranges back to back. */
int fem2d_assemble_ranges(fem2d_plan* plan, int basis_kind, int a_kind, int b_kind, int mode,
                          const double* u_pts, const double* u_w, uint32_t nu,
                          const double* v_pts, const double* v_w, uint32_t nv,
                          uint32_t n_ranges, const uint64_t* slot_begins, const uint64_t* slot_ends,
                          uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals);

/* One-shot equivalent of the reference call: symbolic + numeric + copy-out.  The caller passes its output capacity; when it is too small
 * the call fails with FEM2D_ERR_BAD_ARGUMENT and *nnz_out holds the required size.  Status 1/2/3 exactly as galerkin.rs:42-59. */
int fem2d_galerkin_sample_gep_hcurl(const fem2d_domain_view* view, int device, int basis_kind, int a_kind, int b_kind, int mode,
                                    const double* u_pts, const double* u_w, uint32_t nu,
                                    const double* v_pts, const double* v_w, uint32_t nv,
                                    uint64_t capacity, uint64_t* nnz_out,
                                    uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals);

/* The same one-shot call on `n_devices` GPUs of one process (devices[]: CUDA device indices, normally distinct): what a single Rust caller of the
 * drop-in gets on a multi-GPU box.  The host half of the symbolic phase runs once; one host thread per device builds the pattern there,
 * assembles block r of the two-level row partition (fem2d_plan_row_blocks_split) and copies its slices of rows / cols / A / B to their
 * positions in the caller's arrays.  The result is ONE GEP (galerkin.rs:33-40, linalg.rs:59-81), bit-identical to the single-device call
 * for any device count (no collective: the <= 2 contributions of a key are summed on the device that owns its row). */
int fem2d_galerkin_sample_gep_hcurl_multi(const fem2d_domain_view* view, uint32_t n_devices, const int* devices, int basis_kind, int a_kind, int b_kind,
                                          int mode, const double* u_pts, const double* u_w, uint32_t nu,
                                          const double* v_pts, const double* v_w, uint32_t nv,
                                          uint64_t capacity, uint64_t* nnz_out,
                                          uint32_t* rows, uint32_t* cols, double* a_vals, double* b_vals);

/* Row-block partition of the pattern for `world` ranks: bounds[r]..bounds[r+1] are slot indices aligned to row starts
 * and balanced by nnz (bounds has world+1 entries). */
int fem2d_plan_row_blocks(const fem2d_plan* plan, uint32_t world, uint64_t* bounds);

/* Two-level partition for `world` ranks: the rows of DoFs carried by a single Elem (the reference numbers these Elem-type DoFs
 * first, domain.rs:83-96) and the rows of shared (edge-type) DoFs are each split into `world` row-aligned, nnz-balanced blocks.
 * Rank r assembles [bounds_single[r], bounds_single[r+1]) and [bounds_shared[r], bounds_shared[r+1]): Elem and Edge ids grow
 * together under refinement, so a rank's edge rows mostly read the pair blocks its Elem rows already need, which balances the
 * integrator across ranks. */
int fem2d_plan_row_blocks_split(const fem2d_plan* plan, uint32_t world, uint64_t* bounds_single, uint64_t* bounds_shared);

/* Per-phase timing is opt-in: with `on` != 0 every numeric call records CUDA events between its kernels (sampler, integrator, scatter).
 * Off (the default) the three kernels are chained by programmatic dependent launch and overlap their launch latencies and prologues. */
int fem2d_plan_set_phase_timing(fem2d_plan* plan, int on);

/* Per-phase device timings of the last timed numeric call in milliseconds (CUDA events on the launch stream):
 * ms[0] sampler (K1), ms[1] integrator (K2), ms[2] scatter (K3), ms[3] total; launches[0..2] kernel launch counts. */
int fem2d_plan_last_timing(fem2d_plan* plan, float ms[4], uint32_t launches[4]);
/* Same for the numeric call `calls_back` calls ago (0 = last; the plan keeps the events of its last 64 calls). */
int fem2d_plan_timing(fem2d_plan* plan, uint32_t calls_back, float ms[4], uint32_t launches[4]);

/* Memory kept by the library between calls: freed device blocks are parked in a process-wide cache (at most 8 GB / 96 blocks, served by the
 * library's OWN stream-ordered pool per device -- the application's default pool keeps its settings) and a few pinned staging buffers are
 * kept, so that a one-shot caller (one plan per call) does not pay the driver allocator every time.  fem2d_trim_cache returns all of it to
 * the driver, e.g. before a downstream GPU eigensolver needs the memory.  Call it with no numeric call in flight. */
void fem2d_trim_cache(void);

/* Pinned host memory for outputs (so the D2H copy of a 1 GB value array runs at PCIe speed). */
void* fem2d_host_alloc(size_t bytes);
void fem2d_host_free(void* p);

/* Field evaluation, the caller-side "next" row: UniformFieldSpace::xy_fields (fields.rs:63-127).
 * x_out / y_out: HOST [n_leaves][d][d] (reference indexing quirk: square densities only). leaf_ids: [n_leaves]. */
int fem2d_xy_fields(const fem2d_domain_view* view, int device, int basis_kind, uint32_t density, const double* solution,
                    uint64_t leaf_capacity, uint64_t* n_leaves, uint32_t* leaf_ids, double* x_out, double* y_out);

/* SparseMatrix -> PETSc AIJ binary (the SLEPc route: `impl From<SparseMatrix> for AIJMatrixBinary` + print_to_petsc_binary_file,
 * sparse_matrix.rs:184-264; GEP::print_to_petsc_binary_files, linalg.rs:44-52), emitted from the device arrays: the plan's upper-triangular
 * pattern is mirrored into full symmetric rows on the GPU (per-row counts, sorted columns, big-endian byte swap) and the finished byte
 * image crosses PCIe once.  d_vals: DEVICE array of nnz_upper doubles (the d_a or d_b of fem2d_assemble_device).
 * fem2d_petsc_aij_size: size of the image in bytes and number of entries of the full matrix.  fem2d_petsc_aij_image: the bytes into a host
 * buffer.  fem2d_write_petsc_aij: the same bytes into a file (what `{dir}/tmp/{prefix}_a.dat` holds in the reference). */
int fem2d_petsc_aij_size(fem2d_plan* plan, uint64_t* bytes, uint64_t* nnz_full);
int fem2d_petsc_aij_image(fem2d_plan* plan, const double* d_vals, void* host_image, uint64_t capacity);
int fem2d_write_petsc_aij(fem2d_plan* plan, const double* d_vals, const char* path);

/* FP64 pipe micro-benchmark (roofline denominator for the EXACT integrator): returns achieved GFLOP/s of a DFMA chain
 * (kind 0) or of a DMUL+DADD non-fused chain (kind 1) on `device`. */
int fem2d_fp64_peak(int device, int kind, double* gflops);

int fem2d_device_count(void);
const char* fem2d_status_string(int status);
const char* fem2d_last_error(void); /* thread-local detail message of the last failing call */
const char* fem2d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FEM2D_H */
