"""numpy/scipy restatement of the reference's FEA solve (TEST INFRASTRUCTURE).

Follows ``datagen/fea_analysis.py`` (reference) line by line where the logic
is in-repo, and SURVEY.md Appendix A (sfepy 2023.3 semantics) where the
arithmetic lives in the un-vendored dependency.  Every function cites what it
restates.  Direct solver: ``scipy.sparse.linalg.spsolve`` (SuperLU/COLAMD),
which is what sfepy's ``ScipyDirect({})`` resolves to without scikits.umfpack
(``fea_analysis.py:371-375``, A-12).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------
# material / geometry primitives
# --------------------------------------------------------------------------
def plane_strain_D(young: float, poisson: float) -> np.ndarray:
    """``stiffness_from_youngpoisson(dim=2, young, poisson)`` with sfepy's default
    ``plane='strain'`` (``fea_analysis.py:263-265``).

    Strain vector ordering (e11, e22, 2 e12).
    """
    lam = young * poisson / ((1.0 + poisson) * (1.0 - 2.0 * poisson))
    mu = young / (2.0 * (1.0 + poisson))
    o = np.array([1.0, 1.0, 0.0])
    return lam * np.outer(o, o) + mu * np.diag([2.0, 2.0, 1.0])


def signed_area2(coors: np.ndarray, conn: np.ndarray) -> np.ndarray:
    """Twice the signed area of each cell (shoelace), any k."""
    x = coors[conn, 0]
    y = coors[conn, 1]
    xn = np.roll(x, -1, axis=1)
    yn = np.roll(y, -1, axis=1)
    return (x * yn - xn * y).sum(axis=1)


def fix_orientation(coors: np.ndarray, conn: np.ndarray) -> Tuple[np.ndarray, int]:
    """A-2: clockwise triangles get local vertices 1<->2 swapped (sfepy's
    'bad element orientation ... corrected', ``test_nbs/generateapplication.ipynb:112``).
    Quads (unpinned, F11) are reversed to (0,3,2,1)."""
    conn = np.array(conn, dtype=np.int32, copy=True)
    neg = signed_area2(coors, conn) < 0
    if conn.shape[1] == 3:
        conn[neg] = conn[neg][:, [0, 2, 1]]
    else:
        conn[neg] = conn[neg][:, [0, 3, 2, 1]]
    return conn, int(neg.sum())


# --------------------------------------------------------------------------
# region selectors (the two numpy expressions ARE the spec)
# --------------------------------------------------------------------------
def points_on_edge(coors: np.ndarray, bounding_tags: Tuple[int, int]) -> np.ndarray:
    """``FEAnalysis._get_points_on_edge`` (``fea_analysis.py:182-188``): vertices on the
    infinite line through the two 1-based tagged vertices, |cross| < 1e-14."""
    c0 = coors[bounding_tags[0] - 1]
    c1 = coors[bounding_tags[1] - 1]
    x1, y1 = c1[0] - c0[0], c1[1] - c0[1]
    x2, y2 = coors[:, 0] - c0[0], coors[:, 1] - c0[1]
    return np.where(np.abs(x1 * y2 - x2 * y1) < 1e-14)[0]


def points_in_list(coors: np.ndarray, region_coordinates) -> np.ndarray:
    """``FEAnalysis._get_points_in_list`` (``fea_analysis.py:190-194``): per-scalar
    membership of x and of y in the flattened x-union-y value set (A-18)."""
    return np.where(np.isin(coors, region_coordinates).all(axis=1))[0]


def mesh_edges(conn: np.ndarray) -> np.ndarray:
    """Unique undirected mesh edges (facets of a 2-D mesh), (n_edge, 2) sorted pairs."""
    k = conn.shape[1]
    e = np.concatenate([conn[:, [a, (a + 1) % k]] for a in range(k)], axis=0)
    e = np.sort(e, axis=1)
    return np.unique(e, axis=0)


def facet_region_vertices(conn: np.ndarray, v0: np.ndarray, n_v: int,
                          edges: Optional[np.ndarray] = None) -> np.ndarray:
    """A-7 (facet kind, UNPINNED): vertices of the mesh edges whose both endpoints
    are in ``v0``; isolated members of ``v0`` are dropped."""
    if edges is None:
        edges = mesh_edges(conn)
    m = np.zeros(n_v, dtype=bool)
    m[v0] = True
    sel = edges[m[edges[:, 0]] & m[edges[:, 1]]]
    return np.unique(sel)


def complete_cells(conn: np.ndarray, v0: np.ndarray, n_v: int) -> np.ndarray:
    """A-7 (cell kind, pinned by nnz 270712): cells with ALL vertices in ``v0``."""
    m = np.zeros(n_v, dtype=bool)
    m[v0] = True
    return m[conn].all(axis=1)


# --------------------------------------------------------------------------
# element stiffness  (dw_lin_elastic, order-2 integral; A-4/A-5)
# --------------------------------------------------------------------------
def element_stiffness(coors: np.ndarray, conn: np.ndarray, D: np.ndarray) -> np.ndarray:
    """K_e for every cell; ``D`` is (3,3) or (n_cell,3,3).

    Local DOF order is node-major interleaved: local dof = 2*a + comp, matching
    the global numbering dof = 2*vertex + comp (A-3).
    P1 (k=3): K_e = area * B^T D B, B constant (A-4; the 3-point rule of A-5
    integrates a constant exactly).  Q1 (k=4, UNPINNED F11): bilinear basis on
    the [0,1]^2 reference cell, 2x2 Gauss-Legendre.
    """
    nc, k = conn.shape
    D = np.broadcast_to(np.asarray(D, dtype=np.float64), (nc, 3, 3))
    X = coors[conn]  # (nc,k,2)
    if k == 3:
        x, y = X[:, :, 0], X[:, :, 1]
        b = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], axis=1)
        c = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)
        det = x[:, 0] * b[:, 0] + x[:, 1] * b[:, 1] + x[:, 2] * b[:, 2]  # 2*signed area
        gx = b / det[:, None]
        gy = c / det[:, None]
        B = np.zeros((nc, 3, 6))
        B[:, 0, 0::2] = gx
        B[:, 1, 1::2] = gy
        B[:, 2, 0::2] = gy
        B[:, 2, 1::2] = gx
        area = 0.5 * np.abs(det)
        return area[:, None, None] * np.einsum("eji,ejk,ekl->eil", B, D, B)
    if k == 4:
        g = 0.5 / np.sqrt(3.0)
        pts = [(0.5 - g, 0.5 - g), (0.5 + g, 0.5 - g), (0.5 + g, 0.5 + g), (0.5 - g, 0.5 + g)]
        Ke = np.zeros((nc, 8, 8))
        for (xi, eta) in pts:
            dN = np.array([[-(1 - eta), (1 - eta), eta, -eta],
                           [-(1 - xi), -xi, xi, (1 - xi)]])  # (2,4) d/dxi, d/deta
            J = np.einsum("ia,eaj->eij", dN, X)  # (nc,2,2): J[i,j] = d x_j / d xi_i
            det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
            Ji = np.empty_like(J)
            Ji[:, 0, 0] = J[:, 1, 1] / det
            Ji[:, 0, 1] = -J[:, 0, 1] / det
            Ji[:, 1, 0] = -J[:, 1, 0] / det
            Ji[:, 1, 1] = J[:, 0, 0] / det
            G = np.einsum("eij,ja->eia", Ji, dN)  # (nc,2,4) physical gradients
            B = np.zeros((nc, 3, 8))
            B[:, 0, 0::2] = G[:, 0]
            B[:, 1, 1::2] = G[:, 1]
            B[:, 2, 0::2] = G[:, 1]
            B[:, 2, 1::2] = G[:, 0]
            Ke += (0.25 * np.abs(det))[:, None, None] * np.einsum("eji,ejk,ekl->eil", B, D, B)
        return Ke
    raise ValueError("cells must be triangles or quads")


# --------------------------------------------------------------------------
# strain / stress post-process hook  (fea_analysis.py:397-416)
# --------------------------------------------------------------------------
def cell_strain_stress(coors: np.ndarray, conn: np.ndarray, u: np.ndarray, D: np.ndarray
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """``calculate_stress_strain`` (``fea_analysis.py:397-416``): the post-process hook evaluates
    ``ev_cauchy_strain.2.Omega(u)`` and ``ev_cauchy_stress.2.Omega(m.D, u)`` in ``mode='el_avg'``,
    i.e. per cell (integral over the cell) / (cell volume), strain in sfepy's symmetric-storage order
    (e11, e22, 2 e12) and stress = D strain.  ``u`` is (n_v,2); ``D`` is (3,3) -- the ONE material the
    hook names ``m``: every region's material is created under the name ``m`` (``:286-290``) and the
    expression is evaluated over Omega, so the first material's D is applied to every cell (recalled
    sfepy behaviour, UNPINNED: no artefact of the reference holds these fields for a multi-material
    plate) -- or (n_cell,3,3) for a per-cell D.
    P1: B is constant, so the average is B u_e.  Q1 (UNPINNED, F11): 2x2 Gauss average weighted by |J|.
    Returns (strain (n_cell,3), stress (n_cell,3))."""
    nc, k = conn.shape
    D = np.broadcast_to(np.asarray(D, dtype=np.float64), (nc, 3, 3))
    X = coors[conn]
    U = np.asarray(u, dtype=np.float64)[conn]          # (nc,k,2)
    if k == 3:
        x, y = X[:, :, 0], X[:, :, 1]
        b = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], axis=1)
        c = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)
        det = x[:, 0] * b[:, 0] + x[:, 1] * b[:, 1] + x[:, 2] * b[:, 2]
        gx, gy = b / det[:, None], c / det[:, None]
        strain = np.stack([(gx * U[:, :, 0]).sum(1), (gy * U[:, :, 1]).sum(1),
                           (gy * U[:, :, 0] + gx * U[:, :, 1]).sum(1)], axis=1)
    elif k == 4:
        g = 0.5 / np.sqrt(3.0)
        pts = [(0.5 - g, 0.5 - g), (0.5 + g, 0.5 - g), (0.5 + g, 0.5 + g), (0.5 - g, 0.5 + g)]
        acc = np.zeros((nc, 3))
        vol = np.zeros(nc)
        for (xi, eta) in pts:
            dN = np.array([[-(1 - eta), (1 - eta), eta, -eta], [-(1 - xi), -xi, xi, (1 - xi)]])
            J = np.einsum("ia,eaj->eij", dN, X)
            det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
            Ji = np.empty_like(J)
            Ji[:, 0, 0], Ji[:, 0, 1] = J[:, 1, 1] / det, -J[:, 0, 1] / det
            Ji[:, 1, 0], Ji[:, 1, 1] = -J[:, 1, 0] / det, J[:, 0, 0] / det
            G = np.einsum("eij,ja->eia", Ji, dN)
            w = 0.25 * np.abs(det)
            e = np.stack([(G[:, 0] * U[:, :, 0]).sum(1), (G[:, 1] * U[:, :, 1]).sum(1),
                          (G[:, 1] * U[:, :, 0] + G[:, 0] * U[:, :, 1]).sum(1)], axis=1)
            acc += w[:, None] * e
            vol += w
        strain = acc / vol[:, None]
    else:
        raise ValueError("cells must be triangles or quads")
    stress = np.einsum("eij,ej->ei", D, strain)
    return strain, stress


# --------------------------------------------------------------------------
# assembly  (A-9 .. A-11)
# --------------------------------------------------------------------------
def equation_map(fixed_vertex: np.ndarray) -> Tuple[np.ndarray, int]:
    """A-9: eq[dof] = rank among non-fixed DOFs in ascending DOF order, -1 if fixed."""
    fixed_dof = np.repeat(np.asarray(fixed_vertex, dtype=bool), 2)
    eq = np.full(fixed_dof.size, -1, dtype=np.int64)
    act = np.where(~fixed_dof)[0]
    eq[act] = np.arange(act.size)
    return eq, int(act.size)


def assemble_csr(n_v: int, conn: np.ndarray, Ke: np.ndarray, cell_region: np.ndarray,
                 fixed_vertex: np.ndarray) -> sp.csr_matrix:
    """Reduced stiffness matrix over active DOFs as canonical scipy CSR.

    Cells with ``cell_region < 0`` contribute nothing (F4: seam cells own no
    stiffness).  Rows/cols of fixed DOFs are dropped (A-10).  The result has
    sorted, duplicate-free column indices with explicit zeros kept (A-11); rows
    exist for all active DOFs even when empty (A-18).
    """
    eq, n = equation_map(fixed_vertex)
    k = conn.shape[1]
    use = np.asarray(cell_region) >= 0
    c = conn[use]
    ke = Ke[use]
    dofs = (2 * c[:, :, None] + np.arange(2)[None, None, :]).reshape(len(c), 2 * k)
    e = eq[dofs]
    rows = np.repeat(e[:, :, None], 2 * k, axis=2).ravel()
    cols = np.repeat(e[:, None, :], 2 * k, axis=1).ravel()
    vals = ke.reshape(-1)
    ok = (rows >= 0) & (cols >= 0)
    A = sp.coo_matrix((vals[ok], (rows[ok], cols[ok])), shape=(n, n)).tocsr()
    A.sort_indices()
    return A


# --------------------------------------------------------------------------
# the problem object: restates FEAnalysis.__init__ + calculate()
# --------------------------------------------------------------------------
@dataclass
class OracleProblem:
    """CPU restatement of ``FEAnalysis`` set-up (``fea_analysis.py:32-164``) and
    ``calculate()`` (``:418-461``) on in-memory mesh arrays.

    ``coors`` (n_v,2) f64, ``conn`` (n_cell,k) i32 as read from the file (orientation
    is fixed here).  Tags are the reference's 1-based gmsh point tags.
    """
    coors: np.ndarray
    conn: np.ndarray
    force_vertex_tags_magnitudes: Sequence = ()
    force_edges_tags_magnitudes: Sequence = ()
    constraints_vertex_tags: Sequence = ()
    constraints_edges_tags: Sequence = ()
    num_steps: int = 11
    material_properties_to_vertices: Optional[Dict] = None
    youngs_modulus: float = 210000
    poisson_ratio: float = 0.3
    # derived
    n_flipped: int = 0
    fixed_vertex: np.ndarray = field(default=None, repr=False)
    cell_region: np.ndarray = field(default=None, repr=False)
    D: np.ndarray = field(default=None, repr=False)
    n_regions: int = 1
    load: np.ndarray = field(default=None, repr=False)
    magnitudes_lines: List[str] = field(default_factory=list)
    materials_lines: List[str] = field(default_factory=list)
    region_vertices: Dict[str, np.ndarray] = field(default_factory=dict)

    def __post_init__(self):
        self.coors = np.ascontiguousarray(self.coors, dtype=np.float64)
        self.conn, self.n_flipped = fix_orientation(self.coors, self.conn)
        n_v = len(self.coors)
        edges = None
        # --- forces (fea_analysis.py:76-124); load = sum of per-vertex magnitudes m (A-8)
        m = np.zeros((n_v, 2))
        for i, (tag, mag) in enumerate(self.force_vertex_tags_magnitudes):
            v = np.array([tag - 1])  # 'vertex {tag-1}' (fea_analysis.py:198, A-6)
            self.region_vertices["VertexForce%d" % i] = v
            m[v] += np.asarray(mag, dtype=np.float64)
            self.magnitudes_lines.append("VertexForce%d:%s" % (i, str(mag)))
        for i, (tags, mag) in enumerate(self.force_edges_tags_magnitudes):
            if edges is None:
                edges = mesh_edges(self.conn)
            v = facet_region_vertices(self.conn, points_on_edge(self.coors, tags), n_v, edges)
            self.region_vertices["EdgeForce%d" % i] = v
            nv = max(len(v), 1)  # fea_analysis.py:99
            per = tuple(component / nv for component in mag)
            m[v] += np.asarray(per, dtype=np.float64)
            self.magnitudes_lines.append("EdgeForce%d:%s" % (i, str(per)))
        # --- constraints (fea_analysis.py:127-138): union of region vertices, u.all = 0
        fixed = np.zeros(n_v, dtype=bool)
        for i, tag in enumerate(self.constraints_vertex_tags):
            v = np.array([tag - 1])
            self.region_vertices["VertexConstraint%d" % i] = v
            fixed[v] = True
        for i, tags in enumerate(self.constraints_edges_tags):
            if edges is None:
                edges = mesh_edges(self.conn)
            v = facet_region_vertices(self.conn, points_on_edge(self.coors, tags), n_v, edges)
            self.region_vertices["EdgeConstraint%d" % i] = v
            fixed[v] = True
        self.fixed_vertex = fixed
        # --- materials (fea_analysis.py:140-161, 268-311)
        n_cell = len(self.conn)
        if self.material_properties_to_vertices is not None:
            items = list(self.material_properties_to_vertices.items())
            self.n_regions = len(items)
            self.cell_region = np.full(n_cell, -1, dtype=np.int8)
            self.cell_regions_all = []
            Ds = []
            for i, ((E, nu), verts) in enumerate(items):
                self.materials_lines.append("MaterialRegion%d:%s" % (i, str((E, nu))))
                v0 = points_in_list(self.coors, verts)
                cm = complete_cells(self.conn, v0, n_v)
                self.region_vertices["MaterialRegion%d" % i] = np.unique(self.conn[cm])
                self.cell_regions_all.append(cm)
                # a cell complete in several regions (A-18 round-number overlap) is summed
                # over all of them by sfepy; cell_region keeps the first, extras listed apart.
                first = cm & (self.cell_region < 0)
                self.cell_region[first] = i
                Ds.append(plane_strain_D(E, nu))
            self.D = np.stack(Ds) if Ds else np.zeros((0, 3, 3))
        else:
            self.n_regions = 1
            self.cell_region = np.zeros(n_cell, dtype=np.int8)
            self.cell_regions_all = [np.ones(n_cell, dtype=bool)]
            self.D = plane_strain_D(self.youngs_modulus, self.poisson_ratio)[None]
        # F2/F3: solved system is K u = N_regions * t * m
        self.load = self.n_regions * m

    # ------------------------------------------------------------------
    def overlap_cells(self) -> int:
        """Cells that are complete in more than one material region (A-18)."""
        return int((np.sum(self.cell_regions_all, axis=0) > 1).sum())

    def stiffness(self) -> sp.csr_matrix:
        """Sum over the LHS terms (one per material region, complete cells only), assembled as
        one COO so that explicit zeros survive (A-11) -- sparse '+' would prune them."""
        n_v = len(self.coors)
        eq, n = equation_map(self.fixed_vertex)
        k = self.conn.shape[1]
        R, Cc, V = [], [], []
        for i, cm in enumerate(self.cell_regions_all):
            c = self.conn[cm]
            if not len(c):
                continue
            ke = element_stiffness(self.coors, c, self.D[i])
            dofs = (2 * c[:, :, None] + np.arange(2)[None, None, :]).reshape(len(c), 2 * k)
            e = eq[dofs]
            R.append(np.repeat(e[:, :, None], 2 * k, axis=2).ravel())
            Cc.append(np.repeat(e[:, None, :], 2 * k, axis=1).ravel())
            V.append(ke.reshape(-1))
        if not R:
            return sp.csr_matrix((n, n))
        rows, cols, vals = np.concatenate(R), np.concatenate(Cc), np.concatenate(V)
        ok = (rows >= 0) & (cols >= 0)
        A = sp.coo_matrix((vals[ok], (rows[ok], cols[ok])), shape=(n, n)).tocsr()
        A.sort_indices()
        return A

    def rhs_final(self) -> np.ndarray:
        """Reduced right-hand side at t = 1."""
        eq, n = equation_map(self.fixed_vertex)
        b = np.zeros(n)
        f = self.load.reshape(-1)
        act = eq >= 0
        b[eq[act]] = f[act]
        return b

    def classify(self) -> Dict[str, int]:
        """A-19 (derived): parts of the stiffness mesh (cells connected through shared edges; parts
        touching in one vertex are hinged) that carry fewer than two fixed vertices, and empty rows."""
        import scipy.sparse.csgraph as csg
        n_v = len(self.coors)
        c = self.conn[self.cell_region >= 0].astype(np.int64)
        touched = np.zeros(n_v, dtype=bool)
        touched[c.reshape(-1)] = True
        empty = int((~touched & ~self.fixed_vertex).sum())
        n_c, k = c.shape
        floating = 0
        if n_c:
            ends = np.stack([c, np.roll(c, -1, axis=1)], axis=2).reshape(-1, 2)
            ends.sort(axis=1)
            owner = np.repeat(np.arange(n_c), k)
            edge_id = np.unique(ends, axis=0, return_inverse=True)[1].reshape(-1)
            order = np.argsort(edge_id, kind="stable")
            e, o = edge_id[order], owner[order]
            twin = e[1:] == e[:-1]
            G = sp.coo_matrix((np.ones(int(twin.sum())), (o[:-1][twin], o[1:][twin])), shape=(n_c, n_c))
            ncomp, lab = csg.connected_components(G, directed=False)
            for part in range(ncomp):
                verts = np.unique(c[lab == part])
                floating += int(self.fixed_vertex[verts].sum() < 2)
        return dict(floating_components=floating, empty_rows=empty,
                    well_posed=int(floating == 0 and empty == 0))

    def solve(self, mode: str = "reference") -> np.ndarray:
        return solve_load_steps(self.stiffness(), self.rhs_final(), self.fixed_vertex,
                                self.num_steps, mode=mode)

    def strain_stress(self, u_final: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Final-step ``cauchy_strain`` / ``cauchy_stress`` cell fields of ``domain.<k>.vtk``
        (``fea_analysis.py:397-416``): the hook's ``m.D`` is the first material (see
        ``cell_strain_stress``); step k's fields are t_k times these (F5)."""
        return cell_strain_stress(self.coors, self.conn, u_final, self.D[0])

    def ranges_lines(self, u_steps: np.ndarray) -> List[str]:
        """``ranges.txt`` lines (custom_plotter.py:181-188, A-17): x_1, y_1, x_2, ..."""
        out = []
        for k in range(1, self.num_steps):
            for c, name in enumerate(("displacement_x", "displacement_y")):
                col = u_steps[k][:, c]
                out.append("%s_%d:%s" % (name, k, str((float(col.min()), float(col.max())))))
        return out


def solve_load_steps(A: sp.csr_matrix, b_final: np.ndarray, fixed_vertex: np.ndarray,
                     num_steps: int, mode: str = "reference") -> np.ndarray:
    """Newton(i_max=1) + ScipyDirect + SimpleTimeSteppingSolver (A-12..A-14).

    ``mode='reference'``: for each step k>=1 evaluate r = K u - t_k b and call
    ``spsolve`` afresh (re-factorises every step, as the reference does).
    ``mode='best'``: factor once, one solve, scale by t_k (F5).
    Returns u (num_steps, n_v, 2) with zeros at fixed DOFs (A-15).
    """
    eq, n = equation_map(fixed_vertex)
    act = eq >= 0
    n_v = len(fixed_vertex)
    times = np.linspace(0.0, 1.0, num_steps)
    out = np.zeros((num_steps, 2 * n_v))
    A = sp.csc_matrix(A)
    if mode == "best":
        x = spla.splu(A).solve(b_final) if n else np.zeros(0)
        for k in range(1, num_steps):
            out[k, act] = times[k] * x
    else:
        u = np.zeros(n)
        for k in range(1, num_steps):
            r = A @ u - times[k] * b_final
            if np.linalg.norm(r) >= 1e-10:  # eps_a (A-13)
                du = spla.spsolve(A, r) if n else np.zeros(0)
                u = u - du
            out[k, act] = u
    return out.reshape(num_steps, n_v, 2)
