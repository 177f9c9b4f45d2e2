"""Host-side problem set-up: everything ``FEAnalysis.__init__`` (reference
``datagen/fea_analysis.py:32-164``) derives from the mesh file and the condition dict before any
arithmetic happens -- mesh parsing, region selection, Dirichlet set, material cells, load vector.

These are O(n) numpy operations on exact fp64 values (set membership, collinearity with a fixed
1e-14 threshold), kept on the host so that no FMA contraction can change which vertices a region
selects.  The output is a ``solver.Sample`` ready for the CUDA path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .solver import Sample


# --------------------------------------------------------------------------
# mesh file
# --------------------------------------------------------------------------
def read_mesh(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Parse a MEDIT ``.mesh`` file as written by gmsh (reference ``mesh_generator.py:304-309``)
    into (coors (n_v,2) f64, conn (n_cell,k) i32).  z == 0 is dropped; ``Edges`` are ignored
    (sfepy keeps only the highest-dimension cells, SURVEY A-1)."""
    with open(path, "r") as f:
        words = f.read().split()
    pos, dim = 0, 3
    coors = None
    cells = {}
    sizes = {"Edges": 2, "Triangles": 3, "Quadrilaterals": 4}
    while pos < len(words):
        w = words[pos]
        if w == "Dimension":
            dim = int(words[pos + 1])
            pos += 2
        elif w == "Vertices":
            n = int(words[pos + 1])
            block = np.asarray(words[pos + 2: pos + 2 + n * (dim + 1)], dtype=np.float64)
            coors = block.reshape(n, dim + 1)[:, :dim]
            pos += 2 + n * (dim + 1)
        elif w in sizes:
            k = sizes[w]
            n = int(words[pos + 1])
            block = np.asarray(words[pos + 2: pos + 2 + n * (k + 1)], dtype=np.int64)
            cells[w] = block.reshape(n, k + 1)[:, :k] - 1
            pos += 2 + n * (k + 1)
        elif w == "End":
            break
        else:
            pos += 1
    if coors is None:
        raise ValueError("%s: no Vertices section" % path)
    flat = [a for a in range(dim) if np.ptp(coors[:, a]) > 1e-15]
    if len(flat) != 2:
        raise ValueError("%s: expected a planar mesh" % path)
    tri, quad = cells.get("Triangles"), cells.get("Quadrilaterals")
    if tri is None and quad is None:
        raise ValueError("%s: no triangles or quadrilaterals" % path)
    conn = tri if quad is None or (tri is not None and len(tri) >= len(quad)) else quad
    return np.ascontiguousarray(coors[:, flat]), np.ascontiguousarray(conn, dtype=np.int32)


# --------------------------------------------------------------------------
# materials
# --------------------------------------------------------------------------
def stiffness_plane_strain(young: float, poisson: float) -> np.ndarray:
    """3x3 D of ``stiffness_from_youngpoisson(dim=2, ...)`` -- sfepy's default is plane STRAIN
    (reference ``fea_analysis.py:263-265``), strain order (e11, e22, 2e12)."""
    lam = young * poisson / ((1.0 + poisson) * (1.0 - 2.0 * poisson))
    mu = young / (2.0 * (1.0 + poisson))
    return np.array([[lam + 2.0 * mu, lam, 0.0], [lam, lam + 2.0 * mu, 0.0], [0.0, 0.0, mu]])


# --------------------------------------------------------------------------
# region selection
# --------------------------------------------------------------------------
def vertices_on_line(coors: np.ndarray, tags: Tuple[int, int]) -> np.ndarray:
    """Vertices on the infinite line through the two 1-based tagged vertices
    (``FEAnalysis._get_points_on_edge``, ``fea_analysis.py:182-188``)."""
    p0 = coors[tags[0] - 1]
    p1 = coors[tags[1] - 1]
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    rx, ry = coors[:, 0] - p0[0], coors[:, 1] - p0[1]
    return np.flatnonzero(np.abs(dx * ry - rx * dy) < 1e-14)


def vertices_in_list(coors: np.ndarray, listed) -> np.ndarray:
    """``FEAnalysis._get_points_in_list`` (``fea_analysis.py:190-194``): a vertex is selected when
    its x AND its y both occur anywhere among the listed coordinate values (SURVEY A-18)."""
    return np.flatnonzero(np.isin(coors, listed).all(axis=1))


class MeshTopology:
    """Edges of a mesh, built once per mesh and shared by all facet-kind regions."""

    def __init__(self, conn: np.ndarray, n_v: int):
        k = conn.shape[1]
        a = conn.reshape(-1)
        b = np.roll(conn, -1, axis=1).reshape(-1)
        lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
        key = np.unique(lo * n_v + hi)
        self.e0 = (key // n_v).astype(np.int64)
        self.e1 = (key % n_v).astype(np.int64)
        self.n_v = n_v
        self.k = k

    def facet_vertices(self, selected: np.ndarray) -> np.ndarray:
        """sfepy 'facet'-kind region (SURVEY A-7): vertices of mesh edges whose two endpoints are
        both selected; selected vertices with no such edge are dropped."""
        mask = np.zeros(self.n_v, dtype=bool)
        mask[selected] = True
        keep = mask[self.e0] & mask[self.e1]
        out = np.zeros(self.n_v, dtype=bool)
        out[self.e0[keep]] = True
        out[self.e1[keep]] = True
        return np.flatnonzero(out)


def cells_inside(conn: np.ndarray, selected: np.ndarray, n_v: int) -> np.ndarray:
    """sfepy 'cell'-kind region (SURVEY A-7, F4): cells whose vertices are ALL selected."""
    mask = np.zeros(n_v, dtype=bool)
    mask[selected] = True
    return mask[conn].all(axis=1)


# --------------------------------------------------------------------------
# the condition -> Sample
# --------------------------------------------------------------------------
class ProblemSetup:
    """Result of restating ``FEAnalysis.__init__`` on host arrays: the CUDA ``Sample`` plus the
    text lines and region vertex sets the reference writes / renders."""

    def __init__(self, coors: np.ndarray, conn: np.ndarray,
                 force_vertex_tags_magnitudes: Sequence = (),
                 force_edges_tags_magnitudes: Sequence = (),
                 constraints_vertex_tags: Sequence = (),
                 constraints_edges_tags: Sequence = (),
                 material_properties_to_vertices: Optional[Dict] = None,
                 youngs_modulus: float = 210000, poisson_ratio: float = 0.3,
                 topology: Optional[MeshTopology] = None):
        coors = np.ascontiguousarray(coors, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        n_v, n_cell = len(coors), len(conn)
        self.coors, self.conn = coors, conn
        self.regions: Dict[str, np.ndarray] = {}
        self.magnitudes_lines: List[str] = []
        self.materials_lines: List[str] = []
        topo = topology

        def edge_region(tags):
            nonlocal topo
            if topo is None:
                topo = MeshTopology(conn, n_v)
            return topo.facet_vertices(vertices_on_line(coors, tags))

        # loads: dw_point_load adds the magnitude at every region vertex (fea_analysis.py:76-124)
        load = np.zeros((n_v, 2))
        for i, (tag, mag) in enumerate(force_vertex_tags_magnitudes):
            v = np.array([tag - 1], dtype=np.int64)
            self.regions["VertexForce%d" % i] = v
            load[v] += np.asarray(mag, dtype=np.float64)
            self.magnitudes_lines.append("VertexForce%d:%s" % (i, str(mag)))
        for i, (tags, mag) in enumerate(force_edges_tags_magnitudes):
            v = edge_region(tags)
            self.regions["EdgeForce%d" % i] = v
            count = max(len(v), 1)
            per_vertex = tuple(component / count for component in mag)
            load[v] += np.asarray(per_vertex, dtype=np.float64)
            self.magnitudes_lines.append("EdgeForce%d:%s" % (i, str(per_vertex)))
        # Dirichlet set: u.all = 0 on every constraint region (fea_analysis.py:127-138, 362-369)
        fixed = np.zeros(n_v, dtype=bool)
        for i, tag in enumerate(constraints_vertex_tags):
            v = np.array([tag - 1], dtype=np.int64)
            self.regions["VertexConstraint%d" % i] = v
            fixed[v] = True
        for i, tags in enumerate(constraints_edges_tags):
            v = edge_region(tags)
            self.regions["EdgeConstraint%d" % i] = v
            fixed[v] = True
        # material regions own complete cells only (fea_analysis.py:235-252, 268-311; F4).
        # K_e is linear in D, so a cell that is complete in several regions (A-18 overlap of
        # round coordinates) gets the SUM of their D matrices as an extra table entry.
        if material_properties_to_vertices is not None:
            items = list(material_properties_to_vertices.items())
            n_terms = len(items)
            Ds = []
            member = np.zeros((n_cell, max(n_terms, 1)), dtype=bool)
            for i, ((E, nu), verts) in enumerate(items):
                self.materials_lines.append("MaterialRegion%d:%s" % (i, str((E, nu))))
                inside = cells_inside(conn, vertices_in_list(coors, verts), n_v)
                member[:, i] = inside
                self.regions["MaterialRegion%d" % i] = np.unique(conn[inside])
                Ds.append(stiffness_plane_strain(E, nu))
            count = member.sum(axis=1)
            cell_region = np.full(n_cell, -1, dtype=np.int64)
            single = count == 1
            cell_region[single] = member[single].argmax(axis=1)
            combos = {}
            for c in np.flatnonzero(count > 1):
                key = tuple(np.flatnonzero(member[c]))
                if key not in combos:
                    combos[key] = len(Ds)
                    Ds.append(sum(Ds[r] for r in key))
                cell_region[c] = combos[key]
            if len(Ds) > 127:
                raise ValueError("too many material combinations for int8 region ids")
            D = np.stack(Ds) if Ds else np.zeros((0, 3, 3))
        else:
            n_terms = 1
            cell_region = np.zeros(n_cell, dtype=np.int64)
            D = stiffness_plane_strain(youngs_modulus, poisson_ratio)[None]
        # one Equation per LHS term, each carrying ALL load terms (fea_analysis.py:348-359): the
        # assembled system is K u = n_terms * t * m  
        self.n_terms = n_terms
        self.sample = Sample(coors=coors, conn=conn, cell_region=cell_region.astype(np.int8), D=D,
                             fixed=fixed, rhs=n_terms * load)

    def bbox(self):
        c = self.coors
        return float(c[:, 0].min()), float(c[:, 1].min()), float(c[:, 0].max()), float(c[:, 1].max())


def floating_components(sample: Sample) -> Tuple[int, int]:
    """Derived well-posedness check (SURVEY A-19): (#parts of the stiffness mesh with fewer than two
    fixed vertices, #active vertices touching no stiffness cell).  A *part* is a set of stiffness
    cells connected through shared EDGES: two parts that touch in a single vertex form a hinge, and
    the one without constraints of its own is a mechanism (K singular), so every part must carry
    two fixed vertices itself.  (Conservative: a hinged part held by the hinge plus one fixed
    vertex is stable but is rejected as well.)"""
    import scipy.sparse as sp
    import scipy.sparse.csgraph as csg
    n_v = len(sample.coors)
    c = np.asarray(sample.conn)[np.asarray(sample.cell_region) >= 0].astype(np.int64)
    fixed = np.asarray(sample.fixed, dtype=bool)
    touched = np.zeros(n_v, dtype=bool)
    touched[c.reshape(-1)] = True
    n_c, k = c.shape
    if n_c == 0:
        return 0, int((~fixed).sum())
    a = c.reshape(-1)
    b = np.roll(c, -1, axis=1).reshape(-1)
    key = np.minimum(a, b) * n_v + np.maximum(a, b)
    cell = np.repeat(np.arange(n_c), k)
    order = np.argsort(key, kind="stable")
    ks, cs = key[order], cell[order]
    same = ks[1:] == ks[:-1]
    g = sp.coo_matrix((np.ones(int(same.sum()), dtype=np.int8), (cs[:-1][same], cs[1:][same])), shape=(n_c, n_c))
    ncomp, lab = csg.connected_components(g, directed=False)
    pairs = np.unique(np.repeat(lab, k).astype(np.int64) * n_v + a)          # (part, vertex), once each
    part, vert = pairs // n_v, pairs % n_v
    nfix = np.bincount(part[fixed[vert]], minlength=ncomp)
    return int((nfix < 2).sum()), int((~touched & ~fixed).sum())
