/* =====================================================================================
 * fem2d_host.h -- C-ABI of the host-side mirror of the reference's Mesh / Domain API
 * (src/fem_domain/domain/mesh.rs, domain.rs).  A Rust caller does not need these (it owns a
 * real `Domain` and flattens it into fem2d_domain_view itself); they
 * exist so C / Python callers (tests, bench) can build the same Domains -- identical Elem,
 * Edge, Node and DoF ids -- and feed them to include/fem2d.h.  Pure host code, no GPU.
 *
 * Return convention: 0 ok; positive = fem2dh error kind (mirrors HRefError / PRefError /
 * MeshAccessError variants); message via fem2dh_last_error().
 * ===================================================================================== */
#ifndef FEM2D_HOST_H
#define FEM2D_HOST_H
#include <stdint.h>

#include "fem2d.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fem2dh_mesh fem2dh_mesh;     /* Mesh   (mesh.rs:46-51) */
typedef struct fem2dh_domain fem2dh_domain; /* Domain (domain.rs:42-50) */

enum { FEM2DH_HREF_T = 0, FEM2DH_HREF_U = 1, FEM2DH_HREF_V = 2 }; /* HRef (h_refinement.rs:69-77); ext: -1 None, 0/1 Some(k) */
enum {
    FEM2DH_OK = 0,
    FEM2DH_ERR_ELEM_DOES_NOT_EXIST = 1, FEM2DH_ERR_ELEM_NOT_REFINEABLE = 2, FEM2DH_ERR_DUPLICATE_ELEM_IDS = 3,
    FEM2DH_ERR_ELEM_HAS_CHILDREN = 4, FEM2DH_ERR_EDGE_HAS_CHILDREN = 5, FEM2DH_ERR_MIN_EDGE_LENGTH = 6,
    FEM2DH_ERR_EDGE_ON_EQUAL_POINTS = 7, FEM2DH_ERR_BISECTION_IDX_EXCEEDED = 8, FEM2DH_ERR_REFINEMENT_OUT_OF_BOUNDS = 9,
    FEM2DH_ERR_EXCEEDED_MAX_EXPANSION = 10, FEM2DH_ERR_NEG_EXPANSION = 11, FEM2DH_ERR_BAD_MESH_FILE = 12, FEM2DH_ERR_INTERNAL = 13
};

const char* fem2dh_last_error(void);

/* Mesh::from_file (mesh.rs:142), Mesh::unit (mesh.rs:59), clone, drop */
int fem2dh_mesh_from_file(const char* path, fem2dh_mesh** out);
int fem2dh_mesh_from_arrays(uint64_t n_elements, const double* materials4, const int64_t* node_ids4, uint64_t n_nodes, const double* xy, fem2dh_mesh** out);
int fem2dh_mesh_unit(fem2dh_mesh** out);
int fem2dh_mesh_clone(const fem2dh_mesh* m, fem2dh_mesh** out);
void fem2dh_mesh_free(fem2dh_mesh* m);

uint64_t fem2dh_mesh_num_elems(const fem2dh_mesh* m);
uint64_t fem2dh_mesh_num_edges(const fem2dh_mesh* m);
uint64_t fem2dh_mesh_num_nodes(const fem2dh_mesh* m);
uint64_t fem2dh_mesh_num_elements(const fem2dh_mesh* m);
/* out16: nodes[4], edges[4], parent(-1), has_children, ni, nj, h_u, h_v, element id, n_children ; children4: child ids */
int fem2dh_mesh_elem_info(const fem2dh_mesh* m, uint64_t id, int64_t out16[16], int64_t children4[4]);
/* out10: nodes[2], boundary, dir(0 U / 1 V), parent(-1), children[2](-1), active pair[2](-1), child node(-1) */
int fem2dh_mesh_edge_info(const fem2dh_mesh* m, uint64_t id, int64_t out10[10], double* length);
int fem2dh_mesh_node_info(const fem2dh_mesh* m, uint64_t id, double xy[2], int* boundary);
/* parametric_range (from_ancestor < 0) or relative_parametric_range (elem.rs:170-197): [u_min,u_max,v_min,v_max] */
int fem2dh_mesh_elem_range(const fem2dh_mesh* m, uint64_t id, int64_t from_ancestor, double out4[4]);
int64_t fem2dh_mesh_descendant_elems(const fem2dh_mesh* m, uint64_t id, int include_start, int64_t* out, uint64_t cap);
int64_t fem2dh_mesh_ancestor_elems(const fem2dh_mesh* m, uint64_t id, int include_start, int64_t* out, uint64_t cap);
void fem2dh_mesh_max_expansion_orders(const fem2dh_mesh* m, uint32_t out2[2]);
int fem2dh_mesh_elem_is_h_refineable(const fem2dh_mesh* m, uint64_t id); /* 1 / 0 / -1 (does not exist) */

/* h-refinement (mesh.rs:713-914) */
int fem2dh_mesh_global_h_refinement(fem2dh_mesh* m, int kind, int ext);
int fem2dh_mesh_h_refine_elems(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, int kind, int ext);
int fem2dh_mesh_execute_h_refinements(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* kinds, const int32_t* exts);
/* p-refinement (mesh.rs:1265-1665) */
int fem2dh_mesh_global_p_refinement(fem2dh_mesh* m, int di, int dj);
int fem2dh_mesh_p_refine_elems(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, int di, int dj);
int fem2dh_mesh_execute_p_refinements(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* di, const int32_t* dj);
int fem2dh_mesh_set_global_expansion_orders(fem2dh_mesh* m, int ni, int nj);
int fem2dh_mesh_set_expansion_orders(fem2dh_mesh* m, uint64_t n, const uint64_t* ids, const int32_t* ni, const int32_t* nj);

/* Domain::from_mesh (domain.rs:69) -- copies the mesh */
int fem2dh_domain_from_mesh(const fem2dh_mesh* m, int continuity, fem2dh_domain** out);
int fem2dh_domain_blank(int continuity, fem2dh_domain** out);
void fem2dh_domain_free(fem2dh_domain* d);
const fem2dh_mesh* fem2dh_domain_mesh(const fem2dh_domain* d);
uint64_t fem2dh_domain_num_dofs(const fem2dh_domain* d);
uint64_t fem2dh_domain_num_basis_specs(const fem2dh_domain* d, uint64_t elem_id);
/* local_basis_specs (domain.rs:253) in the reference's list order */
int fem2dh_domain_basis_specs(const fem2dh_domain* d, uint64_t elem_id, int32_t* i, int32_t* j, int32_t* dir, int64_t* dof);
/* The flattened view handed to include/fem2d.h; owned by the domain handle, valid until fem2dh_domain_free. */
const fem2d_domain_view* fem2dh_domain_view(fem2dh_domain* d);

/* gauss_quadrature_points(n, false) (glq.rs:179) and default_ngq (basis.rs:172) */
int fem2dh_gauss_quadrature_points(uint32_t n, double* points, double* weights);
uint64_t fem2dh_default_ngq(uint64_t max_order);

/* Caller-side "next" row: SparseMatrix -> PETSc AIJ binary (sparse_matrix.rs:184-264, linalg.rs:44-52) written straight from
 * the sorted upper-triangular arrays (full symmetric rows, big-endian). */
int fem2dh_write_petsc_aij(const char* path, uint64_t dimension, uint64_t nnz_upper, const uint32_t* rows, const uint32_t* cols, const double* values);
/* Cost of rebuilding an ordered map (the stand-in for BTreeMap<[u32;2], f64>, sparse_matrix.rs:16) from sorted output arrays with end hints:
 * what the Rust shim's `SparseMatrix::from_sorted_upper_tri` would pay at best.  Returns seconds (construction + destruction), < 0 on error. */
double fem2dh_ordered_map_rebuild_seconds(uint64_t nnz, const uint32_t* rows, const uint32_t* cols, const double* values);

#ifdef __cplusplus
}
#endif
#endif /* FEM2D_HOST_H */
