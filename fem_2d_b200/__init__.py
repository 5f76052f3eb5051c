"""fem_2d_b200 -- B200-native Galerkin assembly for the fem_2d H(curl) generalized eigenproblem.

Python face of ``libfem2d_b200.so`` (CUDA, sm_100a).  The names mirror the reference crate's prelude
(/root/reference/src/lib.rs:14-39): ``Mesh``, ``Domain``, ``HRef``, ``PRef``, ``Orders``, ``ContinuityCondition``,
``galerkin_sample_gep_hcurl``, ``GEP``, ``HierPoly``, ``CurlCurl``, ``L2Inner`` ...

There is no CPU fallback: the numeric path raises ``BackendError`` without a CUDA device, and importing this package
raises ``ImportError`` when the native library has not been built (``python -m fem_2d_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEM2D_LIB") or os.path.join(_HERE, "libfem2d_b200.so")   # FEM2D_LIB: tuning builds (build.py FEM2D_VARIANT)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -m fem_2d_b200.build or __graft_entry__.build()). "
        "fem_2d_b200 has no pure-Python / CPU fallback."
    )

_L = C.CDLL(LIB_PATH)

# ---- enums (include/fem2d.h, include/fem2d_host.h) ------------------------------------------------------------------------------
OK, ERR_WRONG_CONTINUITY, ERR_EMPTY_DOF_SET, ERR_INVALID_GLQ = 0, 1, 2, 3
ERR_BAD_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_INTERNAL, ERR_OUT_OF_MEMORY = 100, 101, 102, 103, 104, 105
MODE_EXACT, MODE_SUMFACT, MODE_DMMA = 0, 1, 2
MIN_GLQ_ORDER = 4  # galerkin.rs:13
MAX_DENSE_SIZE = 1000  # nalgebra_solve.rs:6


class HierPoly:  # hierarchical_basis_fns.rs:15
    kind = 0


KOLShapeFn = HierPoly  # pre-rename name used by BASELINE.json / README.md:10


class HierMaxOrtho:  # hierarchical_basis_fns.rs:260
    kind = 1


class CurlCurl:  # integrals.rs:13
    kind = 0


class L2Inner:  # integrals.rs:279
    kind = 1


class ContinuityCondition:  # domain.rs:19-23
    HCurl, HDiv, Discontinuous = 0, 1, 2


class _View(C.Structure):
    _fields_ = [
        ("n_elems", C.c_uint32), ("n_elements", C.c_uint32), ("n_dofs", C.c_uint32), ("continuity", C.c_uint32),
        ("elem_element", C.POINTER(C.c_uint32)), ("elem_parent", C.POINTER(C.c_int32)), ("elem_loc", C.POINTER(C.c_uint8)),
        ("element_p0", C.POINTER(C.c_double)), ("element_p3", C.POINTER(C.c_double)),
        ("element_eps_re", C.POINTER(C.c_double)), ("element_mu_re", C.POINTER(C.c_double)),
        ("bs_off", C.POINTER(C.c_uint32)), ("bs_i", C.POINTER(C.c_uint8)), ("bs_j", C.POINTER(C.c_uint8)),
        ("bs_dir", C.POINTER(C.c_uint8)), ("bs_dof", C.POINTER(C.c_uint32)),
        ("i_max", C.c_uint32), ("j_max", C.c_uint32),
    ]


# Every symbol declared in include/fem2d.h and include/fem2d_host.h (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "fem2d_symbolic", "fem2d_plan_free", "fem2d_plan_info", "fem2d_plan_check_work_items", "fem2d_plan_work_info", "fem2d_plan_source_map_info", "fem2d_plan_row_offsets", "fem2d_plan_pattern_transfer_info", "fem2d_plan_pattern", "fem2d_plan_pattern_device",
    "fem2d_assemble_device", "fem2d_assemble_device_ranges", "fem2d_plan_row_blocks_split", "fem2d_assemble_ranges", "fem2d_assemble", "fem2d_galerkin_sample_gep_hcurl", "fem2d_galerkin_sample_gep_hcurl_multi", "fem2d_plan_row_blocks",
    "fem2d_plan_last_timing", "fem2d_plan_timing", "fem2d_plan_set_phase_timing", "fem2d_assemble_range", "fem2d_host_alloc", "fem2d_host_free", "fem2d_trim_cache", "fem2d_xy_fields", "fem2d_fp64_peak",
    "fem2d_petsc_aij_size", "fem2d_petsc_aij_image", "fem2d_write_petsc_aij", "fem2d_device_count", "fem2d_status_string", "fem2d_last_error", "fem2d_version",
]
HOST_ABI_SYMBOLS = [
    "fem2dh_last_error", "fem2dh_mesh_from_file", "fem2dh_mesh_from_arrays", "fem2dh_mesh_unit", "fem2dh_mesh_clone",
    "fem2dh_mesh_free", "fem2dh_mesh_num_elems", "fem2dh_mesh_num_edges", "fem2dh_mesh_num_nodes", "fem2dh_mesh_num_elements",
    "fem2dh_mesh_elem_info", "fem2dh_mesh_edge_info", "fem2dh_mesh_node_info", "fem2dh_mesh_elem_range",
    "fem2dh_mesh_descendant_elems", "fem2dh_mesh_ancestor_elems", "fem2dh_mesh_max_expansion_orders",
    "fem2dh_mesh_elem_is_h_refineable", "fem2dh_mesh_global_h_refinement", "fem2dh_mesh_h_refine_elems",
    "fem2dh_mesh_execute_h_refinements", "fem2dh_mesh_global_p_refinement", "fem2dh_mesh_p_refine_elems",
    "fem2dh_mesh_execute_p_refinements", "fem2dh_mesh_set_global_expansion_orders", "fem2dh_mesh_set_expansion_orders",
    "fem2dh_domain_from_mesh", "fem2dh_domain_blank", "fem2dh_domain_free", "fem2dh_domain_mesh", "fem2dh_domain_num_dofs",
    "fem2dh_domain_num_basis_specs", "fem2dh_domain_basis_specs", "fem2dh_domain_view", "fem2dh_gauss_quadrature_points",
    "fem2dh_default_ngq", "fem2dh_write_petsc_aij", "fem2dh_ordered_map_rebuild_seconds",
]

for _n in ("fem2d_status_string", "fem2d_last_error", "fem2d_version", "fem2dh_last_error"):
    getattr(_L, _n).restype = C.c_char_p
for _n in ("fem2dh_mesh_num_elems", "fem2dh_mesh_num_edges", "fem2dh_mesh_num_nodes", "fem2dh_mesh_num_elements",
           "fem2dh_domain_num_dofs", "fem2dh_domain_num_basis_specs", "fem2dh_default_ngq"):
    getattr(_L, _n).restype = C.c_uint64
for _n in ("fem2dh_mesh_descendant_elems", "fem2dh_mesh_ancestor_elems"):
    getattr(_L, _n).restype = C.c_int64
_L.fem2dh_ordered_map_rebuild_seconds.restype = C.c_double
_L.fem2dh_domain_mesh.restype = C.c_void_p
_L.fem2dh_domain_view.restype = C.POINTER(_View)
_L.fem2d_host_alloc.restype = C.c_void_p
_L.fem2d_host_alloc.argtypes = [C.c_size_t]
_L.fem2d_host_free.argtypes = [C.c_void_p]
_L.fem2d_plan_free.argtypes = [C.c_void_p]
_L.fem2dh_mesh_free.argtypes = [C.c_void_p]
_L.fem2dh_domain_free.argtypes = [C.c_void_p]


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class BackendError(RuntimeError):
    """CUDA / argument failure of the native library (status >= 100)."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"[{status}] {msg}")
        self.status = status


class GalerkinSamplingError(Exception):
    """galerkin.rs:191-213"""
    WrongContinuityCondition, EmptyDOFSet, InvalidGLQSettings = 1, 2, 3

    def __init__(self, kind: int):
        super().__init__(_L.fem2d_status_string(kind).decode())
        self.kind = kind


class MeshError(Exception):
    """HRefError / PRefError / MeshAccessError (h_refinement.rs:283-296, p_refinement.rs, mesh.rs:1799-1804)."""
    KINDS = {1: "ElemDoesNotExist", 2: "ElemNotRefineable", 3: "DuplicateElemIds", 4: "ElemHasChildren", 5: "EdgeHasChildren",
             6: "MinEdgeLength", 7: "EdgeOnEqualPoints", 8: "BisectionIdxExceeded", 9: "RefinementOutOfBounds",
             10: "ExceededMaxExpansion", 11: "NegExpansion", 12: "BadMeshFile", 13: "Internal"}

    def __init__(self, code: int):
        self.code = code
        self.kind = self.KINDS.get(code, str(code))
        super().__init__(f"{self.kind}: {_L.fem2dh_last_error().decode()}")


def _hck(st: int):
    if st != 0:
        raise MeshError(st)


def _ck(st: int):
    if st == 0:
        return
    if 1 <= st <= 3:
        raise GalerkinSamplingError(st)
    raise BackendError(st, _L.fem2d_last_error().decode() or _L.fem2d_status_string(st).decode())


def device_count() -> int:
    return int(_L.fem2d_device_count())


def version() -> str:
    return _L.fem2d_version().decode()


# ---- HRef / PRef / Orders --------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class HRef:
    """h_refinement.rs:69-117.  kind: 0 T, 1 U, 2 V; ext: -1 None, 0/1 = Some(child index)."""
    kind: int
    ext: int = -1
    T = None  # filled below

    @staticmethod
    def t() -> "HRef":
        return HRef(0)

    @staticmethod
    def u() -> "HRef":
        return HRef(1)

    @staticmethod
    def v() -> "HRef":
        return HRef(2)

    @staticmethod
    def u_extened(child_idx: int) -> "HRef":  # (sic) reference spelling, h_refinement.rs:99
        if child_idx not in (0, 1):
            raise MeshError(8)
        return HRef(1, child_idx)

    @staticmethod
    def v_extened(child_idx: int) -> "HRef":
        if child_idx not in (0, 1):
            raise MeshError(8)
        return HRef(2, child_idx)


HRef.T = HRef(0)


@dataclass(frozen=True)
class PRef:
    di: int
    dj: int

    @staticmethod
    def from_(di: int, dj: int) -> "PRef":
        return PRef(di, dj)


@dataclass(frozen=True)
class Orders:
    ni: int
    nj: int

    @staticmethod
    def new(ni: int, nj: int) -> "Orders":
        return Orders(ni, nj)


@dataclass
class Elem:
    """Snapshot of an Elem (elem.rs:101-110)."""
    id: int
    nodes: list
    edges: list
    parent: int
    children: list
    ni: int
    nj: int
    h_u: int
    h_v: int
    element: int

    def has_children(self) -> bool:
        return bool(self.children)


class Mesh:
    """Mirror of mesh.rs:46-51 over the native host implementation (fem_2d_b200/csrc/host/mesh.hpp)."""

    def __init__(self, handle, owned=True):
        self._h = C.c_void_p(handle)
        self._owned = owned

    def __del__(self, _free=_L.fem2dh_mesh_free):   # bound at definition: module globals may be gone at interpreter exit
        if getattr(self, "_owned", False) and self._h:
            _free(self._h)
            self._h = None

    @staticmethod
    def from_file(path: str) -> "Mesh":
        h = C.c_void_p()
        _hck(_L.fem2dh_mesh_from_file(os.fsencode(path), C.byref(h)))
        return Mesh(h.value)

    @staticmethod
    def unit() -> "Mesh":
        h = C.c_void_p()
        _hck(_L.fem2dh_mesh_unit(C.byref(h)))
        return Mesh(h.value)

    def clone(self) -> "Mesh":
        h = C.c_void_p()
        _hck(_L.fem2dh_mesh_clone(self._h, C.byref(h)))
        return Mesh(h.value)

    # -- queries ---------------------------------------------------------------------------------------------------------------------
    @property
    def num_elems(self) -> int:
        return int(_L.fem2dh_mesh_num_elems(self._h))

    @property
    def num_edges(self) -> int:
        return int(_L.fem2dh_mesh_num_edges(self._h))

    @property
    def num_nodes(self) -> int:
        return int(_L.fem2dh_mesh_num_nodes(self._h))

    def elem(self, eid: int) -> Elem:
        out = np.zeros(16, dtype=np.int64)
        ch = np.zeros(4, dtype=np.int64)
        _hck(_L.fem2dh_mesh_elem_info(self._h, C.c_uint64(eid), _p(out, C.c_int64), _p(ch, C.c_int64)))
        return Elem(eid, out[0:4].tolist(), out[4:8].tolist(), int(out[8]), ch[: int(out[15])].tolist(), int(out[10]), int(out[11]),
                    int(out[12]), int(out[13]), int(out[14]))

    @property
    def elems(self) -> list:
        return [self.elem(e) for e in range(self.num_elems)]

    def edge(self, eid: int) -> dict:
        out = np.zeros(10, dtype=np.int64)
        ln = C.c_double()
        _hck(_L.fem2dh_mesh_edge_info(self._h, C.c_uint64(eid), _p(out, C.c_int64), C.byref(ln)))
        return dict(id=eid, nodes=out[0:2].tolist(), boundary=bool(out[2]), dir=int(out[3]), parent=int(out[4]),
                    children=out[5:7].tolist(), active=out[7:9].tolist(), child_node=int(out[9]), length=ln.value)

    def node(self, nid: int) -> tuple:
        xy = np.zeros(2)
        b = C.c_int()
        _hck(_L.fem2dh_mesh_node_info(self._h, C.c_uint64(nid), _p(xy, C.c_double), C.byref(b)))
        return float(xy[0]), float(xy[1]), bool(b.value)

    def elem_points(self, eid: int) -> list:
        return [self.node(n)[:2] for n in self.elem(eid).nodes]

    def parametric_range(self, eid: int, from_ancestor: int = -1) -> list:
        out = np.zeros(4)
        _hck(_L.fem2dh_mesh_elem_range(self._h, C.c_uint64(eid), C.c_int64(from_ancestor), _p(out, C.c_double)))
        return [[out[0], out[1]], [out[2], out[3]]]

    def descendant_elems(self, eid: int, include_starting_elem: bool) -> list:
        cap = self.num_elems + 1
        out = np.zeros(cap, dtype=np.int64)
        n = _L.fem2dh_mesh_descendant_elems(self._h, C.c_uint64(eid), int(include_starting_elem), _p(out, C.c_int64), C.c_uint64(cap))
        if n < 0:
            raise MeshError(1)
        return out[:n].tolist()

    def ancestor_elems(self, eid: int, include_starting_elem: bool) -> list:
        out = np.zeros(64, dtype=np.int64)
        n = _L.fem2dh_mesh_ancestor_elems(self._h, C.c_uint64(eid), int(include_starting_elem), _p(out, C.c_int64), C.c_uint64(64))
        if n < 0:
            raise MeshError(1)
        return out[:n].tolist()

    def max_expansion_orders(self) -> list:
        out = np.zeros(2, dtype=np.uint32)
        _L.fem2dh_mesh_max_expansion_orders(self._h, _p(out, C.c_uint32))
        return out.tolist()

    def elem_is_h_refineable(self, eid: int) -> bool:
        r = _L.fem2dh_mesh_elem_is_h_refineable(self._h, C.c_uint64(eid))
        if r < 0:
            raise MeshError(1)
        return bool(r)

    # -- h-refinement (mesh.rs:713-914) -----------------------------------------------------------------------------------------------
    def global_h_refinement(self, refinement: HRef) -> None:
        _hck(_L.fem2dh_mesh_global_h_refinement(self._h, refinement.kind, refinement.ext))

    def h_refine_elems(self, elems: Iterable[int], refinement: HRef) -> None:
        ids = np.array(list(elems), dtype=np.uint64)
        _hck(_L.fem2dh_mesh_h_refine_elems(self._h, C.c_uint64(len(ids)), _p(ids, C.c_uint64), refinement.kind, refinement.ext))

    def execute_h_refinements(self, refinements: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in refinements], dtype=np.uint64)
        kinds = np.array([r[1].kind for r in refinements], dtype=np.int32)
        exts = np.array([r[1].ext for r in refinements], dtype=np.int32)
        _hck(_L.fem2dh_mesh_execute_h_refinements(self._h, C.c_uint64(len(ids)), _p(ids, C.c_uint64), _p(kinds, C.c_int32), _p(exts, C.c_int32)))

    def h_refine_with_filter(self, filt: Callable[[Elem], Optional[HRef]]) -> None:
        refs = []
        for e in range(self.num_elems):
            if self.elem_is_h_refineable(e):
                r = filt(self.elem(e))
                if r is not None:
                    refs.append((e, r))
        self.execute_h_refinements(refs)

    # -- p-refinement (mesh.rs:1265-1665) ---------------------------------------------------------------------------------------------
    def global_p_refinement(self, refinement: PRef) -> None:
        _hck(_L.fem2dh_mesh_global_p_refinement(self._h, refinement.di, refinement.dj))

    def p_refine_elems(self, elems: Iterable[int], refinement: PRef) -> None:
        ids = np.array(list(elems), dtype=np.uint64)
        _hck(_L.fem2dh_mesh_p_refine_elems(self._h, C.c_uint64(len(ids)), _p(ids, C.c_uint64), refinement.di, refinement.dj))

    def execute_p_refinements(self, refinements: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in refinements], dtype=np.uint64)
        di = np.array([r[1].di for r in refinements], dtype=np.int32)
        dj = np.array([r[1].dj for r in refinements], dtype=np.int32)
        _hck(_L.fem2dh_mesh_execute_p_refinements(self._h, C.c_uint64(len(ids)), _p(ids, C.c_uint64), _p(di, C.c_int32), _p(dj, C.c_int32)))

    def p_refine_with_filter(self, filt: Callable[[Elem], Optional[PRef]]) -> None:
        refs = []
        for e in self.elems:
            r = filt(e)
            if r is not None:  # constrained to the valid window (mesh.rs:1373-1398)
                refs.append((e.id, PRef(min(max(r.di, -(e.ni - 1)), 20 - e.ni), min(max(r.dj, -(e.nj - 1)), 20 - e.nj))))
        self.execute_p_refinements(refs)

    def set_global_expansion_orders(self, orders: Orders) -> None:
        _hck(_L.fem2dh_mesh_set_global_expansion_orders(self._h, orders.ni, orders.nj))

    def set_expansion_orders(self, poly_orders: Sequence[tuple]) -> None:
        ids = np.array([r[0] for r in poly_orders], dtype=np.uint64)
        ni = np.array([r[1].ni for r in poly_orders], dtype=np.int32)
        nj = np.array([r[1].nj for r in poly_orders], dtype=np.int32)
        _hck(_L.fem2dh_mesh_set_expansion_orders(self._h, C.c_uint64(len(ids)), _p(ids, C.c_uint64), _p(ni, C.c_int32), _p(nj, C.c_int32)))

    def set_expansion_on_elems(self, elems: Iterable[int], orders: Orders) -> None:
        self.set_expansion_orders([(e, orders) for e in elems])

    def set_expansions_with_filter(self, filt: Callable[[Elem], Optional[Orders]]) -> None:
        self.set_expansion_orders([(e.id, o) for e in self.elems for o in [filt(e)] if o is not None])


class DomainView:
    """numpy face of fem2d_domain_view (include/fem2d.h): what a Rust shim would flatten `&Domain` into."""

    def __init__(self, cview: "_View", keepalive=None):
        self.c = cview
        self._keepalive = keepalive
        v = cview
        ne, nel = v.n_elems, v.n_elements
        as_np = np.ctypeslib.as_array
        self.n_elems, self.n_elements, self.n_dofs, self.continuity = ne, nel, v.n_dofs, v.continuity
        self.i_max, self.j_max = v.i_max, v.j_max
        self.elem_element = as_np(v.elem_element, (ne,)) if ne else np.zeros(0, np.uint32)
        self.elem_parent = as_np(v.elem_parent, (ne,)) if ne else np.zeros(0, np.int32)
        self.elem_loc = as_np(v.elem_loc, (ne,)) if ne else np.zeros(0, np.uint8)
        self.element_p0 = as_np(v.element_p0, (nel, 2)) if nel else np.zeros((0, 2))
        self.element_p3 = as_np(v.element_p3, (nel, 2)) if nel else np.zeros((0, 2))
        self.element_eps_re = as_np(v.element_eps_re, (nel,)) if nel else np.zeros(0)
        self.element_mu_re = as_np(v.element_mu_re, (nel,)) if nel else np.zeros(0)
        self.bs_off = as_np(v.bs_off, (ne + 1,)) if v.bs_off else np.zeros(1, np.uint32)
        nbs = int(self.bs_off[-1])
        self.bs_i = as_np(v.bs_i, (nbs,)) if nbs else np.zeros(0, np.uint8)
        self.bs_j = as_np(v.bs_j, (nbs,)) if nbs else np.zeros(0, np.uint8)
        self.bs_dir = as_np(v.bs_dir, (nbs,)) if nbs else np.zeros(0, np.uint8)
        self.bs_dof = as_np(v.bs_dof, (nbs,)) if nbs else np.zeros(0, np.uint32)


class Domain:
    """Mirror of domain.rs:42-50."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self.mesh = Mesh(_L.fem2dh_domain_mesh(self._h), owned=False)
        self._view = None

    def __del__(self, _free=_L.fem2dh_domain_free):
        if getattr(self, "_h", None):
            _free(self._h)
            self._h = None

    @staticmethod
    def from_mesh(mesh: Mesh, cc: int = ContinuityCondition.HCurl) -> "Domain":
        h = C.c_void_p()
        _hck(_L.fem2dh_domain_from_mesh(mesh._h, cc, C.byref(h)))
        d = Domain(h.value)
        d.cc = cc
        return d

    @staticmethod
    def blank(cc: int) -> "Domain":
        h = C.c_void_p()
        _hck(_L.fem2dh_domain_blank(cc, C.byref(h)))
        d = Domain(h.value)
        d.cc = cc
        return d

    @staticmethod
    def unit(cc: int = ContinuityCondition.HCurl) -> "Domain":
        return Domain.from_mesh(Mesh.unit(), cc)

    @property
    def num_dofs(self) -> int:
        return int(_L.fem2dh_domain_num_dofs(self._h))

    def local_basis_specs(self, eid: int):
        """(i, j, dir, dof) arrays in the reference's list order (domain.rs:253-259)."""
        if eid >= self.mesh.num_elems:
            raise MeshError(1)
        n = int(_L.fem2dh_domain_num_basis_specs(self._h, C.c_uint64(eid)))
        i = np.zeros(n, dtype=np.int32); j = np.zeros(n, dtype=np.int32); d = np.zeros(n, dtype=np.int32)
        dof = np.zeros(n, dtype=np.int64)
        _hck(_L.fem2dh_domain_basis_specs(self._h, C.c_uint64(eid), _p(i, C.c_int32), _p(j, C.c_int32), _p(d, C.c_int32), _p(dof, C.c_int64)))
        return i, j, d, dof

    def descendant_basis_specs(self, eid: int):
        return [(d, self.local_basis_specs(d)) for d in self.mesh.descendant_elems(eid, False)]

    def ancestor_basis_specs(self, eid: int):
        return [(d, self.local_basis_specs(d)) for d in self.mesh.ancestor_elems(eid, False)]

    def view(self) -> DomainView:
        if self._view is None:
            self._view = DomainView(_L.fem2dh_domain_view(self._h).contents, keepalive=self)
        return self._view


def gauss_quadrature_points(n: int):
    """gauss_quadrature_points(n, false) (glq.rs:179-222)."""
    p = np.zeros(n); w = np.zeros(n)
    _hck(_L.fem2dh_gauss_quadrature_points(C.c_uint32(n), _p(p, C.c_double), _p(w, C.c_double)))
    return p, w


def default_ngq(max_order: int) -> int:
    return int(_L.fem2dh_default_ngq(C.c_uint64(max_order)))


# ---- results ---------------------------------------------------------------------------------------------------------------------
class SparseMatrix:
    """sparse_matrix.rs:12-17: square symmetric, upper-triangular storage keyed [min,max], (row, col) order, zeros kept."""

    def __init__(self, dimension: int, rows: np.ndarray, cols: np.ndarray, values: np.ndarray):
        self.dimension = dimension
        self.rows, self.cols, self.values = rows, cols, values

    def num_entries(self) -> int:  # sparse_matrix.rs:32-35
        return 2 * len(self.rows) - int(np.count_nonzero(self.rows == self.cols))

    def iter_upper_tri(self):
        for r, c, v in zip(self.rows, self.cols, self.values):
            yield [int(r), int(c)], float(v)

    def to_dense(self) -> np.ndarray:  # sparse_matrix.rs:168-182
        m = np.zeros((self.dimension, self.dimension))
        m[self.rows, self.cols] = self.values
        m[self.cols, self.rows] = self.values
        return m

    def print_to_petsc_binary_file(self, path: str) -> None:  # sparse_matrix.rs:184-264
        r = np.ascontiguousarray(self.rows, dtype=np.uint32); c = np.ascontiguousarray(self.cols, dtype=np.uint32)
        v = np.ascontiguousarray(self.values, dtype=np.float64)
        _hck(_L.fem2dh_write_petsc_aij(os.fsencode(path), C.c_uint64(self.dimension), C.c_uint64(len(r)), _p(r, C.c_uint32), _p(c, C.c_uint32),
                                        _p(v, C.c_double)))


class GEP:
    """linalg.rs:28-42"""

    def __init__(self, a: SparseMatrix, b: SparseMatrix):
        self.a, self.b = a, b

    def to_nalgebra_dense_mats(self):
        return [self.a.to_dense(), self.b.to_dense()]

    def print_to_petsc_binary_files(self, directory: str, prefix: str) -> None:  # linalg.rs:44-52
        self.a.print_to_petsc_binary_file(f"{directory}/tmp/{prefix}_a.dat")
        self.b.print_to_petsc_binary_file(f"{directory}/tmp/{prefix}_b.dat")


class EigenPair:  # linalg.rs:84-97
    def __init__(self, value: float, vector: np.ndarray):
        self.value, self.vector = value, vector

    def normalized_eigenvector(self):
        return self.vector / np.sqrt(np.sum(self.vector ** 2))


class NalgebraGEPError(Exception):
    pass


def nalgebra_solve_gep(gep: GEP, target_eigenvalue: float) -> EigenPair:
    """Downstream dense solve used by the reference's own test (nalgebra_solve.rs:14-50), reproduced with numpy including
    its quirk: SymmetricEigen is fed the non-symmetric B^-1 A and reads the lower triangle only."""
    if gep.a.dimension > MAX_DENSE_SIZE:
        raise NalgebraGEPError("ProblemTooLarge")
    A, B = gep.to_nalgebra_dense_mats()
    try:
        Lc = np.linalg.cholesky(B)
    except np.linalg.LinAlgError:
        raise NalgebraGEPError("FailedToInvertB")
    Li = np.linalg.inv(Lc)
    M = (Li.T @ Li) @ A
    S = np.tril(M) + np.tril(M, -1).T
    w, vecs = np.linalg.eigh(S)
    if np.all(np.abs(w) < 1e-12):
        raise NalgebraGEPError("SpuriouslyConverged")
    k = int(np.argmin(np.abs(w - target_eigenvalue)))
    return EigenPair(float(w[k]), vecs[:, k].copy())


# ---- plan + assembly --------------------------------------------------------------------------------------------------------------
INFO_KEYS = ["nnz_upper", "n_pairs", "n_blocks", "n_classes", "n_values", "n_multi", "max_contrib", "n_tables", "n_work_items",
             "n_dofs", "n_lists", "n_extra", "symbolic_host_us", "symbolic_device_us", "tile_p", "range_tiles_needed"]


class Plan:
    """Symbolic phase result (fem2d_symbolic): fixed pattern + scatter map, reusable across numeric calls."""

    def __init__(self, view: DomainView, device: int = 0, dedupe: bool = True):
        h = C.c_void_p()
        _ck(_L.fem2d_symbolic(C.byref(view.c), int(device), int(bool(dedupe)), C.byref(h)))
        self._h = h
        self._view = view
        self.device = device
        info = (C.c_uint64 * 16)()
        _ck(_L.fem2d_plan_info(self._h, info))
        self.info = {k: int(info[i]) for i, k in enumerate(INFO_KEYS)}
        self.nnz = self.info["nnz_upper"]
        self.n_dofs = self.info["n_dofs"]

    def __del__(self, _free=_L.fem2d_plan_free):
        if getattr(self, "_h", None):
            _free(self._h)
            self._h = None

    def refresh_info(self) -> dict:
        info = (C.c_uint64 * 16)()
        _ck(_L.fem2d_plan_info(self._h, info))
        self.info = {k: int(info[i]) for i, k in enumerate(INFO_KEYS)}
        return self.info

    def check_work_items(self) -> dict:
        """fem2d_plan_check_work_items: self-check of the exact integrator's micro-tile / work-item decomposition."""
        out = (C.c_uint64 * 4)()
        _ck(_L.fem2d_plan_check_work_items(self._h, out))
        return {"tiles": int(out[0]), "same_tiles": int(out[1]), "slots": int(out[2]), "violations": int(out[3])}

    def work_info(self) -> dict:
        """fem2d_plan_work_info: pairs / micro-tiles one numeric call integrates (every class once)."""
        out = (C.c_uint64 * 8)()
        _ck(_L.fem2d_plan_work_info(self._h, out))
        keys = ["same_pairs", "cross_pairs", "same_tiles", "cross_tiles", "pairs_per_same_tile", "pairs_per_cross_tile", "warp_slots", "staging_warps"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def round_fill(self) -> dict:
        """fem2d_debug_round_fill (diagnostic): how full the rounds of the persistent integrator are, and which scales it folds into the weights."""
        out = (C.c_uint64 * 8)()
        _ck(_L.fem2d_debug_round_fill(self._h, out))
        keys = ["tiles", "round_slots", "packs", "rounds", "staged_columns", "rounds_under_half", "fold"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def fp64_lane_ops(self, nu: int, nv: int) -> int:
        """FP64 operations (lane-ops, none of them fusable) of the reference's per-pair quadrature for one numeric call: per same-direction pair
        8 per point (A: 3 products + 1 sum, B: the same) + 4 per quadrature row (solution += inner * u_w, twice); per cross-direction pair 3 per
        point + 2 per row.  Tile padding, slab staging and the final coefficient products are NOT counted."""
        w = self.work_info()
        return w["same_pairs"] * (8 * nu * nv + 4 * nu) + w["cross_pairs"] * (3 * nu * nv + 2 * nu)

    def source_map_info(self) -> dict:
        """fem2d_plan_source_map_info: size of the packed source map the scatter kernel reads."""
        info = (C.c_uint64 * 4)()
        _ck(_L.fem2d_plan_source_map_info(self._h, info))
        return {"plain_chunks": int(info[0]), "chunk_slots": int(info[1]), "map_bytes": int(info[2]), "plain_bytes": int(info[3])}

    def pattern(self):
        rows = np.zeros(self.nnz, dtype=np.uint32); cols = np.zeros(self.nnz, dtype=np.uint32)
        _ck(_L.fem2d_plan_pattern(self._h, _p(rows, C.c_uint32), _p(cols, C.c_uint32)))
        return rows, cols

    def pattern_transfer_info(self) -> dict:
        """fem2d_plan_pattern_transfer_info: bytes the host-output calls move for the pattern (row offsets + column runs)."""
        info = (C.c_uint64 * 4)()
        _ck(_L.fem2d_plan_pattern_transfer_info(self._h, info))
        return {"row_offset_bytes": int(info[0]), "col_runs": int(info[1]), "col_run_bytes": int(info[2]), "plain_bytes": int(info[3])}

    def row_offsets(self) -> np.ndarray:
        """fem2d_plan_row_offsets: CSR row offsets of the upper-triangular pattern (n_dofs + 1 entries)."""
        rp = np.zeros(self.n_dofs + 1, dtype=np.uint64)
        _ck(_L.fem2d_plan_row_offsets(self._h, _p(rp, C.c_uint64)))
        return rp

    def row_blocks(self, world: int) -> np.ndarray:
        b = np.zeros(world + 1, dtype=np.uint64)
        _ck(_L.fem2d_plan_row_blocks(self._h, C.c_uint32(world), _p(b, C.c_uint64)))
        return b

    def row_blocks_split(self, world: int):
        """(bounds of the single-Elem rows, bounds of the shared/edge-type rows), see fem2d_plan_row_blocks_split."""
        b1 = np.zeros(world + 1, dtype=np.uint64); b2 = np.zeros(world + 1, dtype=np.uint64)
        _ck(_L.fem2d_plan_row_blocks_split(self._h, C.c_uint32(world), _p(b1, C.c_uint64), _p(b2, C.c_uint64)))
        return b1, b2

    def assemble_device_ranges(self, glq, d_a: int, d_b: int, ranges, basis=HierPoly, a=CurlCurl, b=L2Inner, mode: int = MODE_EXACT, stream: int = 0):
        """fem2d_assemble_device_ranges: `ranges` = up to 4 (slot_begin, slot_end) pairs handled by one integrator + one scatter launch."""
        up, uw, vp, vw = self._glq_args(glq)
        bg = np.array([r[0] for r in ranges], dtype=np.uint64); en = np.array([r[1] for r in ranges], dtype=np.uint64)
        _ck(_L.fem2d_assemble_device_ranges(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                                            _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)), C.c_uint32(len(ranges)),
                                            _p(bg, C.c_uint64), _p(en, C.c_uint64), C.c_void_p(d_a), C.c_void_p(d_b), C.c_void_p(stream or None)))

    @staticmethod
    def _glq_args(glq):
        (up, uw), (vp, vw) = glq
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (up, uw, vp, vw)]
        return arrs

    def assemble(self, glq, basis=HierPoly, a=CurlCurl, b=L2Inner, mode: int = MODE_EXACT, with_pattern: bool = True):
        """fem2d_assemble: host outputs.  Returns (rows, cols, a_vals, b_vals)."""
        up, uw, vp, vw = self._glq_args(glq)
        rows = np.zeros(self.nnz, dtype=np.uint32) if with_pattern else None
        cols = np.zeros(self.nnz, dtype=np.uint32) if with_pattern else None
        av = np.zeros(self.nnz); bv = np.zeros(self.nnz)
        _ck(_L.fem2d_assemble(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                              _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)),
                              _p(rows, C.c_uint32) if with_pattern else None, _p(cols, C.c_uint32) if with_pattern else None,
                              _p(av, C.c_double), _p(bv, C.c_double)))
        return rows, cols, av, bv

    def assemble_into(self, glq, a_ptr: int, b_ptr: int, host_rows_ptr: int = 0, host_cols_ptr: int = 0, basis=HierPoly, a=CurlCurl, b=L2Inner,
                      mode: int = MODE_EXACT):
        """fem2d_assemble with raw HOST pointers (e.g. pinned buffers from host_alloc)."""
        up, uw, vp, vw = self._glq_args(glq)
        _ck(_L.fem2d_assemble(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                              _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)), C.c_void_p(host_rows_ptr or None),
                              C.c_void_p(host_cols_ptr or None), C.c_void_p(a_ptr), C.c_void_p(b_ptr)))

    def assemble_device(self, glq, d_a: int, d_b: int, basis=HierPoly, a=CurlCurl, b=L2Inner, mode: int = MODE_EXACT, slot_begin: int = 0,
                        slot_end: int = 2 ** 64 - 1, stream: int = 0):
        """fem2d_assemble_device: d_a / d_b are raw DEVICE pointers (e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        up, uw, vp, vw = self._glq_args(glq)
        _ck(_L.fem2d_assemble_device(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                                     _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)), C.c_uint64(slot_begin), C.c_uint64(slot_end),
                                     C.c_void_p(d_a), C.c_void_p(d_b), C.c_void_p(stream or None)))

    def assemble_range_into(self, glq, slot_begin: int, slot_end: int, a_ptr: int, b_ptr: int, rows_ptr: int = 0, cols_ptr: int = 0,
                            basis=HierPoly, a=CurlCurl, b=L2Inner, mode: int = MODE_EXACT):
        """fem2d_assemble_range with raw HOST pointers: numeric phase + D2H of the row-block slice."""
        up, uw, vp, vw = self._glq_args(glq)
        _ck(_L.fem2d_assemble_range(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                                    _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)), C.c_uint64(slot_begin), C.c_uint64(slot_end),
                                    C.c_void_p(rows_ptr or None), C.c_void_p(cols_ptr or None), C.c_void_p(a_ptr), C.c_void_p(b_ptr)))

    def assemble_ranges_into(self, glq, ranges, a_ptr: int, b_ptr: int, rows_ptr: int = 0, cols_ptr: int = 0, basis=HierPoly, a=CurlCurl, b=L2Inner,
                             mode: int = MODE_EXACT):
        """fem2d_assemble_ranges with raw HOST pointers: numeric phase + D2H of the ranges, back to back."""
        up, uw, vp, vw = self._glq_args(glq)
        bg = np.array([r[0] for r in ranges], dtype=np.uint64); en = np.array([r[1] for r in ranges], dtype=np.uint64)
        _ck(_L.fem2d_assemble_ranges(self._h, basis.kind, a.kind, b.kind, int(mode), _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)),
                                     _p(vp, C.c_double), _p(vw, C.c_double), C.c_uint32(len(vp)), C.c_uint32(len(ranges)), _p(bg, C.c_uint64),
                                     _p(en, C.c_uint64), C.c_void_p(rows_ptr or None), C.c_void_p(cols_ptr or None), C.c_void_p(a_ptr), C.c_void_p(b_ptr)))

    def petsc_aij_image(self, d_vals: int) -> np.ndarray:
        """fem2d_petsc_aij_image: the PETSc AIJ binary image (sparse_matrix.rs:184-264) of the matrix whose nnz_upper values sit at the DEVICE
        pointer `d_vals`, built on the GPU; returns the bytes."""
        nbytes = C.c_uint64(); nf = C.c_uint64()
        _ck(_L.fem2d_petsc_aij_size(self._h, C.byref(nbytes), C.byref(nf)))
        img = np.empty(nbytes.value, dtype=np.uint8)
        _ck(_L.fem2d_petsc_aij_image(self._h, C.c_void_p(d_vals), img.ctypes.data_as(C.c_void_p), C.c_uint64(nbytes.value)))
        return img

    def write_petsc_aij(self, d_vals: int, path: str) -> None:
        """fem2d_write_petsc_aij: the same image straight into a file."""
        _ck(_L.fem2d_write_petsc_aij(self._h, C.c_void_p(d_vals), os.fsencode(path)))

    def set_phase_timing(self, on: bool = True):
        """fem2d_plan_set_phase_timing: record per-phase CUDA events in the following numeric calls (off by default)."""
        _ck(_L.fem2d_plan_set_phase_timing(self._h, int(bool(on))))

    def last_timing(self, calls_back: int = 0):
        ms = (C.c_float * 4)(); ln = (C.c_uint32 * 4)()
        _ck(_L.fem2d_plan_timing(self._h, C.c_uint32(calls_back), ms, ln))
        return dict(sampler_ms=ms[0], integrator_ms=ms[1], scatter_ms=ms[2], total_ms=ms[3], launches=int(ln[3]),
                    launches_by_phase=[int(ln[0]), int(ln[1]), int(ln[2])])


def galerkin_sample_gep_hcurl(domain: Domain, glq_grid_dim=None, basis=HierPoly, a=CurlCurl, b=L2Inner, device: int = 0,
                              mode: int = MODE_EXACT, glq=None) -> GEP:
    """galerkin_sample_gep_hcurl::<BSpace, AI, BI>(&domain, Option<[usize; 2]>) -> Result<GEP, GalerkinSamplingError>
    (galerkin.rs:33-187).  `glq` optionally supplies ((u_pts,u_w),(v_pts,v_w)) -- the nodes are an input of the native path."""
    cc = getattr(domain, "cc", ContinuityCondition.HCurl)
    if cc != ContinuityCondition.HCurl:
        raise GalerkinSamplingError(GalerkinSamplingError.WrongContinuityCondition)
    if domain.num_dofs == 0:
        raise GalerkinSamplingError(GalerkinSamplingError.EmptyDOFSet)
    if glq is None:
        if glq_grid_dim is not None:
            if glq_grid_dim[0] < MIN_GLQ_ORDER or glq_grid_dim[1] < MIN_GLQ_ORDER:
                raise GalerkinSamplingError(GalerkinSamplingError.InvalidGLQSettings)
            nu, nv = glq_grid_dim
        else:
            mo = domain.mesh.max_expansion_orders()
            nu, nv = default_ngq(mo[0]), default_ngq(mo[1])  # basis.rs:83-90
        glq = (gauss_quadrature_points(nu), gauss_quadrature_points(nv))
    plan = Plan(domain.view(), device=device)
    rows, cols, av, bv = plan.assemble(glq, basis, a, b, mode)
    n = domain.num_dofs
    return GEP(SparseMatrix(n, rows, cols, av), SparseMatrix(n, rows, cols, bv))


def galerkin_sample_gep_hcurl_multi(domain_or_view, glq, devices, basis=HierPoly, a=CurlCurl, b=L2Inner, mode: int = MODE_EXACT, out=None):
    """fem2d_galerkin_sample_gep_hcurl_multi: the one-shot call on several GPUs of this process; returns (rows, cols, a_vals, b_vals) of the ONE
    assembled GEP.  `out` = optional (rows_ptr, cols_ptr, a_ptr, b_ptr, capacity) raw HOST pointers (e.g. pinned buffers from host_alloc)."""
    view = domain_or_view.view() if isinstance(domain_or_view, Domain) else domain_or_view
    (up, uw), (vp, vw) = glq
    up, uw, vp, vw = [np.ascontiguousarray(x, dtype=np.float64) for x in (up, uw, vp, vw)]
    devs = np.ascontiguousarray(list(devices), dtype=np.int32)
    nnz = C.c_uint64()
    if out is None:
        # size probe with a zero capacity: the call reports the required size and fails with BAD_ARGUMENT
        st = _L.fem2d_galerkin_sample_gep_hcurl_multi(C.byref(view.c), C.c_uint32(len(devs)), _p(devs, C.c_int32), basis.kind, a.kind, b.kind, int(mode),
                                                      _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)), _p(vp, C.c_double), _p(vw, C.c_double),
                                                      C.c_uint32(len(vp)), C.c_uint64(0), C.byref(nnz), None, None, None, None)
        if st != ERR_BAD_ARGUMENT or nnz.value == 0:
            _ck(st)
        n = nnz.value
        rows = np.zeros(n, dtype=np.uint32); cols = np.zeros(n, dtype=np.uint32); av = np.zeros(n); bv = np.zeros(n)
        ptrs = (rows.ctypes.data, cols.ctypes.data, av.ctypes.data, bv.ctypes.data, n)
    else:
        rows = cols = av = bv = None
        ptrs = out
    _ck(_L.fem2d_galerkin_sample_gep_hcurl_multi(C.byref(view.c), C.c_uint32(len(devs)), _p(devs, C.c_int32), basis.kind, a.kind, b.kind, int(mode),
                                                 _p(up, C.c_double), _p(uw, C.c_double), C.c_uint32(len(up)), _p(vp, C.c_double), _p(vw, C.c_double),
                                                 C.c_uint32(len(vp)), C.c_uint64(ptrs[4]), C.byref(nnz), C.c_void_p(ptrs[0] or None), C.c_void_p(ptrs[1] or None),
                                                 C.c_void_p(ptrs[2]), C.c_void_p(ptrs[3])))
    return (rows, cols, av, bv) if out is None else int(nnz.value)


def fp64_peak(device: int = 0, kind: int = 0) -> float:
    """Measured FP64 pipe throughput in GFLOP/s: kind 0 = DFMA chain, 1 = non-fused DMUL+DADD chain."""
    g = C.c_double()
    _ck(_L.fem2d_fp64_peak(int(device), int(kind), C.byref(g)))
    return g.value


def host_alloc(nbytes: int) -> int:
    p = _L.fem2d_host_alloc(C.c_size_t(nbytes))
    if not p:
        raise BackendError(ERR_OUT_OF_MEMORY, "cudaMallocHost failed")
    return p


def trim_cache() -> None:
    """fem2d_trim_cache: hand the cached device blocks and pinned staging buffers back to the driver."""
    _L.fem2d_trim_cache()


def host_free(ptr: int) -> None:
    _L.fem2d_host_free(C.c_void_p(ptr))


# ---- caller-side "next" row: field evaluation (fields.rs) ------------------------------------------------------------------------------
class UniformFieldError(Exception):
    """fields.rs:412-433"""


class UniformFieldSpace:
    """Mirror of UniformFieldSpace (fields.rs:17-22): field quantities on a uniform grid over every leaf Elem.
    `xy_fields` runs on the GPU (fem2d_xy_fields); quantities are dicts {leaf elem id: ndarray[d][d]}."""

    def __init__(self, domain: Domain, densities):
        if densities[0] != densities[1]:
            # the reference allocates [densities[1]][densities[0]] but indexes [m < d0][n < d1] (fields.rs:83-84,102-112)
            raise UniformFieldError("non-square densities index out of bounds in the reference; only square grids are supported")
        self.domain = domain
        self.densities = list(densities)
        self.quantities = {}

    def xy_fields(self, vector_name: str, solution, basis=HierPoly, device: int = 0):
        """fields.rs:63-127: returns [f"{vector_name}_x", f"{vector_name}_y"]."""
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        if len(sol) != self.domain.num_dofs:
            raise UniformFieldError(f"Domain size ({self.domain.num_dofs}) does not match solution size ({len(sol)})!")
        d = self.densities[0]
        cap = self.domain.mesh.num_elems
        ids = np.zeros(cap, dtype=np.uint32)
        x = np.zeros((cap, d, d)); y = np.zeros((cap, d, d))
        n = C.c_uint64()
        _ck(_L.fem2d_xy_fields(C.byref(self.domain.view().c), int(device), basis.kind, C.c_uint32(d), _p(sol, C.c_double), C.c_uint64(cap),
                               C.byref(n), _p(ids, C.c_uint32), _p(x, C.c_double), _p(y, C.c_double)))
        n = n.value
        xn, yn = f"{vector_name}_x", f"{vector_name}_y"
        self.quantities[xn] = {int(ids[k]): x[k] for k in range(n)}
        self.quantities[yn] = {int(ids[k]): y[k] for k in range(n)}
        return [xn, yn]

    def map_to_quantity(self, name: str, result_name: str, operator) -> None:           # fields.rs:256-279
        if name not in self.quantities:
            raise UniformFieldError(f"Missing quantity '{name}', Cannot apply operation!")
        self.quantities[result_name] = {e: np.vectorize(operator)(v) for e, v in self.quantities[name].items()}

    def expression_2arg(self, operand_names, result_name: str, expression) -> None:      # fields.rs:301-339
        a, b = operand_names
        for nm in (a, b):
            if nm not in self.quantities:
                raise UniformFieldError(f"Missing quantity '{nm}', Cannot apply operation!")
        qa, qb = self.quantities[a], self.quantities[b]
        self.quantities[result_name] = {e: np.vectorize(expression)(qa[e], qb[e]) for e in qa}
