"""CPU oracle for the fem_2d Galerkin assembly path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/fem2d_oracle.cpp`` (a C++ restatement of the reference crate; see
that file's header for what pins it).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``fem_2d_b200``) never does.

The classes mirror the reference API names (``Mesh``, ``Domain``, ``galerkin_sample_gep_hcurl``;
/root/reference/src/lib.rs:14-39) so the pinning tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

T, U, V = 0, 1, 2  # HRef kinds (h_refinement.rs:69-77)
HCURL, HDIV, DISCONTINUOUS = 0, 1, 2
HIER_POLY, HIER_MAX_ORTHO = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``).  Building the checker is not using it."""
    src = os.path.join(_HERE, "fem2d_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_mesh_from_arrays", "orc_mesh_unit", "orc_mesh_clone", "orc_domain_from_mesh", "orc_domain_mesh",
                     "orc_assemble", "orc_assemble_elems"):
            getattr(L, name).restype = C.c_void_p
        for name in ("orc_mesh_num_elems", "orc_mesh_num_edges", "orc_mesh_num_nodes", "orc_mesh_descendant_elems",
                     "orc_mesh_ancestor_elems", "orc_domain_num_dofs", "orc_domain_num_basis_specs", "orc_gep_nnz",
                     "orc_default_ngq", "orc_xy_fields", "orc_elem_entries_size"):
            getattr(L, name).restype = C.c_int64
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _err() -> str:
    return lib().orc_last_error().decode()


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


@dataclass
class ElemInfo:
    id: int
    nodes: list
    edges: list
    parent: int
    has_children: bool
    ni: int
    nj: int
    h_u: int
    h_v: int
    element: int
    children: list


@dataclass
class HRef:
    kind: int
    ext: int = -1


class Mesh:
    """mesh.rs:46-51"""

    def __init__(self, handle):
        if not handle:
            raise OracleError(_err())
        self._h = C.c_void_p(handle)
        self._owned = True

    def __del__(self):
        if getattr(self, "_owned", False) and self._h:
            lib().orc_mesh_free(self._h)

    # -- constructors ------------------------------------------------------------------
    @staticmethod
    def unit() -> "Mesh":
        return Mesh(lib().orc_mesh_unit())

    @staticmethod
    def from_file(path: str) -> "Mesh":
        with open(path) as f:
            j = json.load(f)
        mats = np.array([e["materials"] for e in j["Elements"]], dtype=np.float64).reshape(-1)
        nids = np.array([e["node_ids"] for e in j["Elements"]], dtype=np.int64).reshape(-1)
        xy = np.array(j["Nodes"], dtype=np.float64).reshape(-1)
        return Mesh(lib().orc_mesh_from_arrays(len(j["Elements"]), _p(mats, C.c_double), _p(nids, C.c_int64),
                                               len(j["Nodes"]), _p(xy, C.c_double)))

    def clone(self) -> "Mesh":
        return Mesh(lib().orc_mesh_clone(self._h))

    # -- queries -----------------------------------------------------------------------
    @property
    def num_elems(self) -> int:
        return lib().orc_mesh_num_elems(self._h)

    @property
    def num_edges(self) -> int:
        return lib().orc_mesh_num_edges(self._h)

    @property
    def num_nodes(self) -> int:
        return lib().orc_mesh_num_nodes(self._h)

    def elem(self, eid: int) -> ElemInfo:
        out = np.zeros(16, dtype=np.int64)
        ch = np.zeros(4, dtype=np.int64)
        if lib().orc_mesh_elem_info(self._h, C.c_int64(eid), _p(out, C.c_int64), _p(ch, C.c_int64)) != 0:
            raise OracleError(_err())
        return ElemInfo(eid, out[0:4].tolist(), out[4:8].tolist(), int(out[8]), bool(out[9]), int(out[10]), int(out[11]),
                        int(out[12]), int(out[13]), int(out[14]), ch[: int(out[15])].tolist())

    def elems(self) -> list:
        return [self.elem(i) for i in range(self.num_elems)]

    def edge(self, eid: int) -> dict:
        out = np.zeros(10, dtype=np.int64)
        ln = C.c_double()
        if lib().orc_mesh_edge_info(self._h, C.c_int64(eid), _p(out, C.c_int64), C.byref(ln)) != 0:
            raise OracleError(_err())
        return dict(id=eid, nodes=out[0:2].tolist(), boundary=bool(out[2]), dir=int(out[3]), parent=int(out[4]),
                    children=out[5:7].tolist(), active=out[7:9].tolist(), child_node=int(out[9]), length=ln.value)

    def node(self, nid: int) -> tuple:
        xy = np.zeros(2)
        b = lib().orc_mesh_node_xy(self._h, C.c_int64(nid), _p(xy, C.c_double))
        if b < 0:
            raise OracleError(_err())
        return float(xy[0]), float(xy[1]), bool(b)

    def elem_points(self, eid: int) -> list:
        return [self.node(n)[:2] for n in self.elem(eid).nodes]

    def parametric_range(self, eid: int, from_ancestor: int = -1) -> list:
        out = np.zeros(4)
        if lib().orc_mesh_elem_ranges(self._h, C.c_int64(eid), C.c_int64(from_ancestor), _p(out, C.c_double)) != 0:
            raise OracleError(_err())
        return [[out[0], out[1]], [out[2], out[3]]]

    def descendant_elems(self, eid: int, include: bool) -> list:
        cap = self.num_elems + 1
        out = np.zeros(cap, dtype=np.int64)
        n = lib().orc_mesh_descendant_elems(self._h, C.c_int64(eid), int(include), _p(out, C.c_int64), C.c_int64(cap))
        if n < 0:
            raise OracleError(_err())
        return out[:n].tolist()

    def ancestor_elems(self, eid: int, include: bool) -> list:
        cap = 64
        out = np.zeros(cap, dtype=np.int64)
        n = lib().orc_mesh_ancestor_elems(self._h, C.c_int64(eid), int(include), _p(out, C.c_int64), C.c_int64(cap))
        if n < 0:
            raise OracleError(_err())
        return out[:n].tolist()

    def max_expansion_orders(self) -> list:
        out = np.zeros(2, dtype=np.int32)
        lib().orc_mesh_max_expansion_orders(self._h, _p(out, C.c_int32))
        return out.tolist()

    def elem_is_h_refineable(self, eid: int) -> bool:
        r = lib().orc_mesh_elem_is_h_refineable(self._h, C.c_int64(eid))
        if r < 0:
            raise OracleError(_err())
        return bool(r)

    # -- h-refinement (mesh.rs:713-914) ----------------------------------------------------
    def execute_h_refinements(self, refs: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in refs], dtype=np.int64)
        kinds = np.array([r[1].kind if isinstance(r[1], HRef) else r[1] for r in refs], dtype=np.int32)
        exts = np.array([r[1].ext if isinstance(r[1], HRef) else -1 for r in refs], dtype=np.int32)
        if lib().orc_mesh_execute_h_refinements(self._h, C.c_int64(len(refs)), _p(ids, C.c_int64), _p(kinds, C.c_int32),
                                                 _p(exts, C.c_int32)) != 0:
            raise OracleError(_err())

    def global_h_refinement(self, href) -> None:
        self.execute_h_refinements([(e, href) for e in range(self.num_elems) if self.elem_is_h_refineable(e)])

    def h_refine_elems(self, ids: Iterable[int], href) -> None:
        ids = list(ids)
        if len(set(ids)) != len(ids):
            raise OracleError("DuplicateElemIds")
        self.execute_h_refinements([(e, href) for e in ids])

    def h_refine_with_filter(self, filt: Callable[[ElemInfo], Optional[object]]) -> None:
        refs = []
        for e in range(self.num_elems):
            if self.elem_is_h_refineable(e):
                r = filt(self.elem(e))
                if r is not None:
                    refs.append((e, r))
        self.execute_h_refinements(refs)

    # -- p-refinement (mesh.rs:1265-1665) --------------------------------------------------
    def execute_p_refinements(self, refs: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in refs], dtype=np.int64)
        di = np.array([r[1] for r in refs], dtype=np.int32)
        dj = np.array([r[2] for r in refs], dtype=np.int32)
        if lib().orc_mesh_execute_p_refinements(self._h, C.c_int64(len(refs)), _p(ids, C.c_int64), _p(di, C.c_int32),
                                                 _p(dj, C.c_int32)) != 0:
            raise OracleError(_err())

    def global_p_refinement(self, di: int, dj: int) -> None:
        refs = []
        for e in self.elems():  # PRef::constrained_to the valid window (mesh.rs:1265-1281)
            lo_u, hi_u = -(e.ni - 1), 20 - e.ni
            lo_v, hi_v = -(e.nj - 1), 20 - e.nj
            refs.append((e.id, min(max(di, lo_u), hi_u), min(max(dj, lo_v), hi_v)))
        self.execute_p_refinements(refs)

    def p_refine_elems(self, ids: Iterable[int], di: int, dj: int) -> None:
        ids = list(ids)
        if len(set(ids)) != len(ids):
            raise OracleError("DuplicateElemIds")
        self.execute_p_refinements([(e, di, dj) for e in ids])

    def set_expansion_orders(self, orders: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in orders], dtype=np.int64)
        ni = np.array([r[1] for r in orders], dtype=np.int32)
        nj = np.array([r[2] for r in orders], dtype=np.int32)
        if lib().orc_mesh_set_expansion_orders(self._h, C.c_int64(len(orders)), _p(ids, C.c_int64), _p(ni, C.c_int32),
                                                _p(nj, C.c_int32)) != 0:
            raise OracleError(_err())

    def set_global_expansion_orders(self, ni: int, nj: int) -> None:
        self.set_expansion_orders([(e, ni, nj) for e in range(self.num_elems)])


class Domain:
    """domain.rs:42-50"""

    def __init__(self, handle):
        if not handle:
            raise OracleError(_err())
        self._h = C.c_void_p(handle)
        m = Mesh.__new__(Mesh)
        m._h = C.c_void_p(lib().orc_domain_mesh(self._h))
        m._owned = False
        self.mesh = m

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_domain_free(self._h)

    @staticmethod
    def from_mesh(mesh: Mesh, cc: int = HCURL) -> "Domain":
        return Domain(lib().orc_domain_from_mesh(mesh._h, cc))

    @property
    def num_dofs(self) -> int:
        return lib().orc_domain_num_dofs(self._h)

    def local_basis_specs(self, eid: int):
        """(i, j, dir, dof) arrays in the reference's list order (domain.rs:253-259)."""
        n = lib().orc_domain_num_basis_specs(self._h, C.c_int64(eid))
        i = np.zeros(n, dtype=np.int32); j = np.zeros(n, dtype=np.int32); d = np.zeros(n, dtype=np.int32)
        dof = np.zeros(n, dtype=np.int64)
        if lib().orc_domain_basis_specs(self._h, C.c_int64(eid), _p(i, C.c_int32), _p(j, C.c_int32), _p(d, C.c_int32),
                                        _p(dof, C.c_int64)) != 0:
            raise OracleError(_err())
        return i, j, d, dof

    def descendant_basis_specs(self, eid: int):
        return [(d, self.local_basis_specs(d)) for d in self.mesh.descendant_elems(eid, False)]

    def ancestor_basis_specs(self, eid: int):
        return [(d, self.local_basis_specs(d)) for d in self.mesh.ancestor_elems(eid, False)]


def gauss_quadrature_points(n: int):
    """glq.rs:179-222 (include_endpoints=false)."""
    p = np.zeros(n); w = np.zeros(n)
    if lib().orc_glq(n, _p(p, C.c_double), _p(w, C.c_double)) != 0:
        raise OracleError(_err())
    return p, w


def default_ngq(max_order: int) -> int:
    return lib().orc_default_ngq(C.c_int64(max_order))


def basis_tables(kind: int, n_max: int, pts: np.ndarray):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    outs = [np.zeros((n_max + 1, len(pts))) for _ in range(4)]
    if lib().orc_basis_tables(kind, n_max, len(pts), _p(pts, C.c_double), *[_p(o, C.c_double) for o in outs]) != 0:
        raise OracleError(_err())
    return outs  # norm, norm_d1, tang, tang_d1


@dataclass
class GEP:
    """Upper-triangular (row<=col) entries sorted by (row, col) == BTreeMap<[u32;2]> order (sparse_matrix.rs:12-17)."""
    n_dofs: int
    rows: np.ndarray
    cols: np.ndarray
    a: np.ndarray
    b: np.ndarray
    t_integrate: float = 0.0
    t_merge: float = 0.0

    def dense(self):
        A = np.zeros((self.n_dofs, self.n_dofs)); B = np.zeros_like(A)
        A[self.rows, self.cols] = self.a; A[self.cols, self.rows] = self.a
        B[self.rows, self.cols] = self.b; B[self.cols, self.rows] = self.b
        return A, B


class GalerkinSamplingError(Exception):
    """galerkin.rs:191-195"""
    NAMES = {1: "WrongContinuityCondition", 2: "EmptyDOFSet", 3: "InvalidGLQSettings"}

    def __init__(self, code):
        super().__init__(self.NAMES.get(code, str(code)))
        self.code = code


def galerkin_sample_gep_hcurl(domain: Domain, glq_grid_dim=None, basis: int = HIER_POLY, n_threads: int = 1,
                              glq=None) -> GEP:
    """galerkin.rs:33-187.  ``glq`` optionally supplies ((u_pts,u_w),(v_pts,v_w)) (nodes are an input of the path)."""
    if glq is None:
        if glq_grid_dim is None:
            mo = domain.mesh.max_expansion_orders()
            glq_grid_dim = [default_ngq(mo[0]), default_ngq(mo[1])]
        if glq_grid_dim[0] < 4 or glq_grid_dim[1] < 4:
            # the reference checks this before generating points (galerkin.rs:51-57)
            if domain.num_dofs == 0:
                raise GalerkinSamplingError(2)
            raise GalerkinSamplingError(3)
        glq = (gauss_quadrature_points(glq_grid_dim[0]), gauss_quadrature_points(glq_grid_dim[1]))
    (up, uw), (vp, vw) = glq
    up, uw, vp, vw = [np.ascontiguousarray(x, dtype=np.float64) for x in (up, uw, vp, vw)]
    g = lib().orc_assemble(domain._h, basis, _p(up, C.c_double), _p(uw, C.c_double), C.c_int64(len(up)), _p(vp, C.c_double),
                           _p(vw, C.c_double), C.c_int64(len(vp)), n_threads)
    if not g:
        raise OracleError(_err())
    g = C.c_void_p(g)
    try:
        st = lib().orc_gep_status(g)
        if st != 0:
            raise GalerkinSamplingError(st)
        nnz = lib().orc_gep_nnz(g)
        rows = np.zeros(nnz, dtype=np.uint32); cols = np.zeros(nnz, dtype=np.uint32)
        a = np.zeros(nnz); b = np.zeros(nnz)
        lib().orc_gep_copy(g, _p(rows, C.c_uint32), _p(cols, C.c_uint32), _p(a, C.c_double), _p(b, C.c_double))
        t = np.zeros(2)
        lib().orc_gep_times(g, _p(t, C.c_double))
        return GEP(domain.num_dofs, rows, cols, a, b, float(t[0]), float(t[1]))
    finally:
        lib().orc_gep_free(g)


def assemble_elems(domain: Domain, elem_ids, glq, basis: int = HIER_POLY):
    """Per-Elem matrices (galerkin.rs:73-183 closure body) of the selected Elems, NOT merged: (elem, rows, cols, a, b)."""
    (up, uw), (vp, vw) = glq
    up, uw, vp, vw = [np.ascontiguousarray(x, dtype=np.float64) for x in (up, uw, vp, vw)]
    ids = np.ascontiguousarray(elem_ids, dtype=np.int64)
    h = lib().orc_assemble_elems(domain._h, basis, _p(up, C.c_double), _p(uw, C.c_double), C.c_int64(len(up)), _p(vp, C.c_double),
                                 _p(vw, C.c_double), C.c_int64(len(vp)), _p(ids, C.c_int64), C.c_int64(len(ids)))
    if not h:
        raise OracleError(_err())
    h = C.c_void_p(h)
    try:
        n = lib().orc_elem_entries_size(h)
        el = np.zeros(n, dtype=np.int64); rows = np.zeros(n, dtype=np.uint32); cols = np.zeros(n, dtype=np.uint32)
        a = np.zeros(n); b = np.zeros(n)
        lib().orc_elem_entries_copy(h, _p(el, C.c_int64), _p(rows, C.c_uint32), _p(cols, C.c_uint32), _p(a, C.c_double), _p(b, C.c_double))
        return el, rows, cols, a, b
    finally:
        lib().orc_elem_entries_free(h)


def xy_fields(domain: Domain, densities, solution, basis: int = HIER_POLY):
    """fields.rs:63-127.  Returns (leaf_ids, x[leaf][m][n], y[leaf][m][n])."""
    d0, d1 = densities
    sol = np.ascontiguousarray(solution, dtype=np.float64)
    if len(sol) != domain.num_dofs:
        raise OracleError("MismatchedSolutionSize")
    cap = domain.mesh.num_elems
    ids = np.zeros(cap, dtype=np.int64)
    x = np.zeros(cap * d0 * d1); y = np.zeros(cap * d0 * d1)
    n = lib().orc_xy_fields(domain._h, basis, C.c_int64(d0), C.c_int64(d1), _p(sol, C.c_double), C.c_int64(cap), _p(ids, C.c_int64),
                            _p(x, C.c_double), _p(y, C.c_double))
    if n < 0:
        raise OracleError(_err())
    return ids[:n], x[: n * d0 * d1].reshape(n, d1, d0), y[: n * d0 * d1].reshape(n, d1, d0)


def nalgebra_solve_surrogate(gep: GEP, target: float) -> float:
    """nalgebra_solve_gep (nalgebra_solve.rs:14-50) feeds the NON-symmetric B^-1 A to SymmetricEigen, which reads the
    lower triangle only; the returned eigenvalue is that of tril(B^-1 A) symmetrised."""
    A, B = gep.dense()
    L = np.linalg.cholesky(B)
    Binv = np.linalg.inv(L).T @ np.linalg.inv(L)
    M = Binv @ A
    S = np.tril(M) + np.tril(M, -1).T
    ev = np.linalg.eigvalsh(S)
    return float(ev[np.argmin(np.abs(ev - target))])


def petsc_aij_bytes(dimension: int, rows, cols, values) -> bytes:
    """CPU restatement (numpy, test infrastructure) of `impl From<SparseMatrix> for AIJMatrixBinary` and
    `AIJMatrixBinary::print_to_petsc_binary_file` (sparse_matrix.rs:184-264): the bytes the reference writes for a matrix whose
    upper-triangular entries are (rows, cols, values).
      :187-197  row_counts: +1 on the diagonal, +1 for both rows otherwise
      :200-206  full_matrix = {[c, r]: v} of every entry, then .append(entries): keys ascending, an equal key ([r, r]) is overwritten
      :209-212  j = column of every key in that order, a = its value
      :228-262  header b"\\0\\x12{P" (= 1211216 as u32 BE), dim, dim, len(a); counts, j as u32 BE; a as f64 BE"""
    rows = np.ascontiguousarray(rows, dtype=np.int64); cols = np.ascontiguousarray(cols, dtype=np.int64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    counts = np.zeros(dimension, dtype=np.int64)
    np.add.at(counts, rows, 1)
    off = rows != cols
    np.add.at(counts, cols[off], 1)
    # the mirrored map first, the original entries appended after it: on equal keys (the diagonal) the appended value wins -- same value
    key_r = np.concatenate([cols, rows]); key_c = np.concatenate([rows, cols]); val = np.concatenate([values, values])
    order = np.lexsort((np.arange(len(key_r)), key_c, key_r))          # by (row, col), insertion order last
    key_r, key_c, val = key_r[order], key_c[order], val[order]
    last = np.ones(len(key_r), dtype=bool)
    last[:-1] = (key_r[:-1] != key_r[1:]) | (key_c[:-1] != key_c[1:])  # keep the LAST of equal keys (the appended one)
    j, a = key_c[last], val[last]
    head = np.array([1211216, dimension, dimension, len(a)], dtype=">u4").tobytes()
    assert head[:4] == b"\x00\x12{P"
    return head + counts.astype(">u4").tobytes() + j.astype(">u4").tobytes() + a.astype(">f8").tobytes()
