"""Deterministic synthetic plate generator: geometry sampler, in-house triangular mesher and
condition sampler.

The reference builds its inputs with shapely + gmsh + sklearn (``datagen/mesh_generator.py``),
none of which exist in this image, and north_star keeps meshing on the host outside the timed
path.  This module produces inputs of the same character for the benchmark configurations of
BASELINE.json ("M-plate(seed)"):

* geometry  -- union of 1-3 convex polygons (3-8 points drawn in the reference's "diversity"
  half-boxes, ``mesh_generator.py:125-196``) with 0-3 convex holes, normalised to unit extent
  (``:84-93``);
* mesh      -- boundary points every ``mesh_size``, jittered hexagonal interior lattice,
  scipy Delaunay, triangles outside the domain dropped: 2.5-10 k vertices, valence ~6 like
  gmsh's frontal mesh at ``mesh_size=1e-2``; geometry points come first so that the 1-based point
  tag equals vertex index + 1 (SURVEY A-6);
* conditions -- a restatement of ``sample_conditions`` (``:397-521``) driven by one seeded
  ``random.Random``, in the reference's draw order.  Material regions (``:319-385``), selected by
  ``region_method``: ``"reference"`` restates the reference's two methods on scikit-learn (KMeans of the
  vertices into 5-20 clusters merged by a second KMeans over the FLATTENED centres, or agglomerative
  clustering with complete / average / ward linkage -- quirks included, see ``_regions_kmeans``);
  ``"lloyd"`` is the fast stand-in the benchmark workloads were defined with in round 1 (Lloyd clusters
  merged into bands).  The reference draws the method, cluster counts and materials from the UNSEEDED
  global ``random`` module and leaves KMeans unseeded, so its own runs are not reproducible; here every
  one of those draws comes from the seeded stream.  Materials from the reference's 18-entry table
  (``:33-55``), forces uniform integers in +-[1, 1000] (``:66, 493-519``).

Parity is always "same mesh through both paths", so the mesher does not need to match gmsh.
"""
from __future__ import annotations

import math
import random
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
from scipy.spatial import ConvexHull, Delaunay, cKDTree

# (name, E [MPa], nu) -- reference mesh_generator.py:33-55
MATERIALS = [
    ("Steel", 210000, 0.3), ("Aluminum", 68900, 0.33), ("Copper", 117000, 0.34), ("Brass", 97000, 0.33),
    ("Titanium", 105000, 0.34), ("Stainless Steel", 195000, 0.3), ("Nickel", 207000, 0.31),
    ("Zinc", 100000, 0.25), ("Lead", 17500, 0.44), ("Magnesium", 46500, 0.35), ("Concrete", 30000, 0.2),
    ("Fibre Glass", 84700, 0.26), ("Carbon Fibre A4S", 225000, 0.25), ("Bronze", 120000, 0.34),
    ("Tungsten", 411000, 0.28), ("Silver", 83000, 0.37), ("Gold", 78000, 0.44), ("Platinum", 168000, 0.38),
]


_TREE_CACHE = None


def _tree_cache():
    """joblib.Memory in a per-process temporary directory (removed at exit)."""
    global _TREE_CACHE
    if _TREE_CACHE is None:
        import atexit
        import shutil
        import tempfile
        from joblib import Memory
        d = tempfile.mkdtemp(prefix="fea_agglomerative_")
        atexit.register(shutil.rmtree, d, True)
        _TREE_CACHE = Memory(location=d, verbose=0)
    return _TREE_CACHE


class GeometryRejected(Exception):
    pass


# --------------------------------------------------------------------------
# polygons
# --------------------------------------------------------------------------
def _area2(ring: np.ndarray) -> float:
    x, y = ring[:, 0], ring[:, 1]
    return float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _hull(points) -> np.ndarray:
    pts = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    if len(pts) < 3:
        raise GeometryRejected("degenerate hull")
    try:
        h = ConvexHull(pts)
    except Exception as e:  # collinear input
        raise GeometryRejected(str(e))
    ring = pts[h.vertices]  # counter-clockwise in 2-D
    if abs(_area2(ring)) < 1e-4:
        raise GeometryRejected("sliver hull")
    return ring


def points_in_ring(ring: np.ndarray, p: np.ndarray) -> np.ndarray:
    """Even-odd ray casting, vectorised over points p (n,2)."""
    x, y = p[:, 0], p[:, 1]
    inside = np.zeros(len(p), dtype=bool)
    a, b = ring, np.roll(ring, -1, axis=0)
    for (x0, y0), (x1, y1) in zip(a, b):
        if y0 == y1:
            continue
        cond = (y0 > y) != (y1 > y)
        xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
        inside ^= cond & (x < xi)
    return inside


def _seg_intersections(p, q, ring):
    """Parameters t in (0,1) where segment p->q properly crosses edges of ring."""
    a, b = ring, np.roll(ring, -1, axis=0)
    d = q - p
    e = b - a
    den = d[0] * e[:, 1] - d[1] * e[:, 0]
    ok = np.abs(den) > 1e-14
    w = a - p
    t = np.where(ok, (w[:, 0] * e[:, 1] - w[:, 1] * e[:, 0]) / np.where(ok, den, 1), -1)
    s = np.where(ok, (w[:, 0] * d[1] - w[:, 1] * d[0]) / np.where(ok, den, 1), -1)
    hit = ok & (t > 1e-12) & (t < 1 - 1e-12) & (s >= 0) & (s <= 1)
    return t[hit]


def union_exterior(polys: List[np.ndarray]) -> np.ndarray:
    """Exterior ring (CCW) of the union of convex CCW polygons; rejects disconnected unions
    (the reference's unary_union then yields a MultiPolygon and the plate is redrawn,
    ``generate.py:57-60``).  Holes enclosed by the union are filled, as the reference only keeps
    ``geometry.exterior`` (``mesh_generator.py:184``)."""
    if len(polys) == 1:
        return polys[0]
    segs = []
    for i, P in enumerate(polys):
        others = [Q for j, Q in enumerate(polys) if j != i]
        for p, q in zip(P, np.roll(P, -1, axis=0)):
            ts = [0.0, 1.0]
            for Q in others:
                ts.extend(_seg_intersections(p, q, Q).tolist())
            ts = sorted(set(ts))
            for t0, t1 in zip(ts[:-1], ts[1:]):
                if t1 - t0 < 1e-12:
                    continue
                mid = (p + 0.5 * (t0 + t1) * (q - p))[None]
                if any(points_in_ring(Q, mid)[0] for Q in others):
                    continue
                segs.append((p + t0 * (q - p), p + t1 * (q - p)))
    key = lambda v: (round(float(v[0]), 9), round(float(v[1]), 9))
    nxt = {}
    for a, b in segs:
        if key(a) in nxt:
            raise GeometryRejected("degenerate union (shared start point)")
        nxt[key(a)] = (a, b)
    loops = []
    used = set()
    for k0 in list(nxt):
        if k0 in used:
            continue
        loop, k = [], k0
        while k not in used:
            if k not in nxt:
                raise GeometryRejected("open boundary chain")
            used.add(k)
            a, b = nxt[k]
            loop.append(a)
            k = key(b)
        if k != k0:
            raise GeometryRejected("boundary chain does not close")
        loops.append(np.array(loop))
    outer = [l for l in loops if _area2(l) > 0]
    if len(outer) != 1:
        raise GeometryRejected("union is not connected")
    return outer[0]


def _rings_cross(A: np.ndarray, B: np.ndarray) -> bool:
    for p, q in zip(A, np.roll(A, -1, axis=0)):
        if len(_seg_intersections(p, q, B)):
            return True
    return False


@dataclass
class Geometry:
    exterior: np.ndarray        # (n,2) CCW
    holes: List[np.ndarray]     # each (m,2) CW

    def bounds(self):
        e = self.exterior
        return float(e[:, 0].min()), float(e[:, 1].min()), float(e[:, 0].max()), float(e[:, 1].max())

    def contains(self, p: np.ndarray) -> np.ndarray:
        m = points_in_ring(self.exterior, p)
        for h in self.holes:
            m &= ~points_in_ring(h, p)
        return m

    def rings(self) -> List[np.ndarray]:
        return [self.exterior] + list(self.holes)


class PlateGenerator:
    """Seeded stand-in for the reference's ``MeshGenerator`` (same constructor defaults)."""

    def __init__(self, num_polygons_range=(1, 3), points_per_polygon_range=(3, 8), holes_per_polygon_range=(0, 3),
                 points_per_hole_range=(3, 4), num_regions=(1, 5), force_magnitude_range=(1, 1000), random_seed=None,
                 region_method: str = "reference"):
        self.num_polygons_range = num_polygons_range
        self.points_per_polygon_range = points_per_polygon_range
        self.holes_per_polygon_range = holes_per_polygon_range
        self.points_per_hole_range = points_per_hole_range
        self.num_regions = num_regions
        self.force_magnitude_range = force_magnitude_range
        if region_method not in ("reference", "lloyd"):
            raise ValueError("region_method must be 'reference' or 'lloyd'")
        self.region_method = region_method
        self.random = random.Random(random_seed)
        self.np_rng = np.random.default_rng(random_seed)
        self.mesh: Optional[Tuple[np.ndarray, np.ndarray]] = None

    # ---- geometry (mesh_generator.py:102-196) ---------------------------------
    def _random_float(self) -> float:
        return float(self.random.randint(0, 1000)) / 1000

    def _random_coordinates(self, n, bounds=(0, 0, 1, 1)):
        return [(bounds[0] + self._random_float() * (bounds[2] - bounds[0]),
                 bounds[1] + self._random_float() * (bounds[3] - bounds[1])) for _ in range(n)]

    def _random_convex(self) -> np.ndarray:
        n = self.random.randint(*self.points_per_polygon_range)
        boxes = [[0.5, 0, 1, 1], [0, 0, 0.5, 1], [0, 0.5, 1, 1], [0, 0, 1, 0.5]]
        self.random.shuffle(boxes)
        pts = (self._random_coordinates(n // 3, boxes[0]) + self._random_coordinates(n // 3, boxes[1])
               + self._random_coordinates(n - 2 * n // 3, boxes[2]))
        return _hull(pts)

    def generate_geometry(self) -> Geometry:
        polys = [self._random_convex() for _ in range(self.random.randint(*self.num_polygons_range))]
        ext = union_exterior(polys)
        holes: List[np.ndarray] = []
        bx = (ext[:, 0].min(), ext[:, 1].min(), ext[:, 0].max(), ext[:, 1].max())
        for _ in range(self.random.randint(*self.holes_per_polygon_range)):
            n = self.random.randint(*self.points_per_hole_range)
            for _attempt in range(200):
                try:
                    h = _hull(self._random_coordinates(n, bx))
                except GeometryRejected:
                    continue
                if not points_in_ring(ext, h).all() or _rings_cross(h, ext):
                    continue
                if any(_rings_cross(h, o) or points_in_ring(o, h).any() or points_in_ring(h, o).any() for o in holes):
                    continue
                holes.append(h[::-1].copy())
                break
        return Geometry(ext, holes)

    @staticmethod
    def normalize_geometry(g: Geometry) -> Geometry:
        x0, y0, x1, y1 = g.bounds()
        s = 1.0 / max(x1 - x0, y1 - y0)
        f = lambda r: (r - np.array([x0, y0])) * s
        return Geometry(f(g.exterior), [f(h) for h in g.holes])

    # ---- mesh (stand-in for gmsh, mesh_generator.py:246-317) -------------------
    def generate_mesh(self, g: Geometry, mesh_size: float = 1e-2):
        """Returns (polygons_ptags, polygons_ltag_ptags) like the reference and stores
        ``self.mesh = (coors, conn)``.  Holes are numbered before the exterior, as the reference
        creates the internal gmsh polygons first (``:268-285``)."""
        h = mesh_size
        order = list(g.holes) + [g.exterior]
        corners = np.concatenate(order)
        ptags, ltags, tag, ltag = [], [], 1, 1
        for ring in order:
            t = list(range(tag, tag + len(ring)))
            d = OrderedDict()
            for i in range(len(t)):
                d[ltag] = (t[i], t[(i + 1) % len(t)])
                ltag += 1
            ptags.append(t)
            ltags.append(d)
            tag += len(ring)
        edge_pts, bnd_edges, nxt_id = [], [], len(corners)
        base = 0
        for ring in order:
            m = len(ring)
            for i in range(m):
                a, b = ring[i], ring[(i + 1) % m]
                n = max(1, int(math.ceil(np.linalg.norm(b - a) / h - 1e-9)))
                ids = [base + i] + list(range(nxt_id, nxt_id + n - 1)) + [base + (i + 1) % m]
                for k in range(1, n):
                    edge_pts.append(a + (b - a) * (k / n))
                nxt_id += n - 1
                bnd_edges.extend(zip(ids[:-1], ids[1:]))
            base += m
        bpts = np.concatenate([corners, np.array(edge_pts).reshape(-1, 2)])
        x0, y0, x1, y1 = g.bounds()
        dy = h * math.sqrt(3) / 2
        ox, oy = self.np_rng.uniform(0, h), self.np_rng.uniform(0, dy)
        ys = np.arange(y0 - dy + oy, y1 + dy, dy)
        xs = np.arange(x0 - h + ox, x1 + h, h)
        X, Y = np.meshgrid(xs, ys)
        X = X + (np.arange(len(ys))[:, None] % 2) * (h / 2)
        lat = np.stack([X.ravel(), Y.ravel()], axis=1)
        lat += self.np_rng.uniform(-0.05 * h, 0.05 * h, size=lat.shape)
        lat = lat[g.contains(lat)]
        dist, _ = cKDTree(bpts).query(lat)
        lat = lat[dist >= 0.75 * h]
        pts = np.concatenate([bpts, lat])
        tri = Delaunay(pts).simplices.astype(np.int64)
        cen = pts[tri].mean(axis=1)
        tri = tri[g.contains(cen)]
        P = pts[tri]
        a2 = ((P[:, 1, 0] - P[:, 0, 0]) * (P[:, 2, 1] - P[:, 0, 1]) - (P[:, 2, 0] - P[:, 0, 0]) * (P[:, 1, 1] - P[:, 0, 1]))
        tri = tri[np.abs(a2) > 1e-3 * h * h]
        # conformity: every boundary edge must be used by exactly one kept triangle
        e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
        keys, cnt = np.unique(e[:, 0] * len(pts) + e[:, 1], return_counts=True)
        be = np.sort(np.array(bnd_edges, dtype=np.int64), axis=1)
        bkeys = be[:, 0] * len(pts) + be[:, 1]
        pos = np.searchsorted(keys, bkeys)
        ok = (pos < len(keys)) & (keys[np.minimum(pos, len(keys) - 1)] == bkeys)
        if not ok.all() or not (cnt[pos[ok]] == 1).all() or (cnt > 2).any():
            raise GeometryRejected("triangulation does not conform to the boundary")
        if (cnt == 1).sum() != len(bkeys):
            raise GeometryRejected("triangulation has spurious boundary edges")
        used = np.zeros(len(pts), dtype=bool)
        used[tri.reshape(-1)] = True
        if not used[:len(corners)].all():
            raise GeometryRejected("geometry point unused")
        new = np.cumsum(used) - 1
        self.mesh = (np.ascontiguousarray(pts[used]), np.ascontiguousarray(new[tri].astype(np.int32)))
        return ptags, ltags

    # ---- material regions (mesh_generator.py:319-395) --------------------------
    def _create_regions_randomly(self) -> List[np.ndarray]:
        """``_create_regions_randomly`` (``:379-386``): one of the two clustering methods, drawn first."""
        if self.region_method == "lloyd":
            return self._regions_lloyd()
        method = self.random.choice(["kmeans", "agglomerative"])
        if method == "kmeans":
            return self._regions_kmeans()
        return self._regions_agglomerative(self.random.choice(["complete", "average", "ward"]))

    @staticmethod
    def _sklearn():
        try:
            from sklearn.cluster import AgglomerativeClustering, KMeans
        except ImportError as e:   # no silent change of the sampled distribution
            raise ImportError("region_method='reference' needs scikit-learn (the reference's own dependency, "
                              "mesh_generator.py:16); pass region_method='lloyd' for the stand-in") from e
        return AgglomerativeClustering, KMeans

    def _regions_kmeans(self) -> List[np.ndarray]:
        """``_create_regions_with_kmeans`` (``:319-352``).  KMeans of the mesh points (3-D, z = 0, as pyvista
        reads them) into 5-20 clusters; the clusters are merged into ``num_regions`` regions by a second KMeans
        that the reference runs on ``cluster_centers_.reshape(-1, 1)`` -- the FLATTENED centres
        (x0, y0, z0, x1, ...) -- and indexes with the cluster number: cluster i takes the label of scalar i of
        that list, i.e. of coordinate i % 3 of centre i // 3.  Restated as written."""
        _, KMeans = self._sklearn()
        coors = self.mesh[0]
        pts = np.concatenate([coors, np.zeros((len(coors), 1))], axis=1)
        n_clusters = self.random.randint(5, 20)
        km = KMeans(n_clusters=n_clusters, n_init=1, random_state=self.random.randrange(2 ** 31))
        lab = km.fit_predict(pts)
        n_regions = self.random.randint(*self.num_regions)
        km2 = KMeans(n_clusters=n_regions, n_init=1, random_state=self.random.randrange(2 ** 31))
        lab2 = km2.fit_predict(km.cluster_centers_.reshape(-1, 1))
        regions: List[List[np.ndarray]] = [[] for _ in range(n_regions)]
        for i in range(n_clusters):
            regions[int(lab2[i])].append(coors[lab == i])
        return [np.concatenate(r) if r else np.zeros((0, 2)) for r in regions]

    def _regions_agglomerative(self, link: str) -> List[np.ndarray]:
        """``_create_regions_with_agglomerative_clustering`` (``:354-377``): the vertices clustered straight
        into ``num_regions`` regions."""
        Agglomerative, _ = self._sklearn()
        coors = self.mesh[0]
        pts = np.concatenate([coors, np.zeros((len(coors), 1))], axis=1)
        n_regions = self.random.randint(*self.num_regions)
        # the merge tree depends on the points and the linkage only, the number of regions only on where it is cut:
        # scikit-learn's own `memory` option keeps the tree (O(n^2) to build: seconds for a 10 k-vertex plate)
        # between the candidate conditions of a plate -- same labels, one tree per (plate, linkage)
        lab = Agglomerative(n_clusters=n_regions, linkage=link, memory=_tree_cache()).fit_predict(pts)
        return [coors[lab == r] for r in range(n_regions)]

    def _regions_lloyd(self) -> List[np.ndarray]:
        """Stand-in (round-1 benchmark workloads): Lloyd clusters merged into bands along a random direction."""
        coors = self.mesh[0]
        n_clusters = self.random.randint(5, 20)
        n_regions = self.random.randint(*self.num_regions)
        centres = coors[self.np_rng.choice(len(coors), n_clusters, replace=False)].copy()
        for _ in range(6):  # Lloyd
            lab = cKDTree(centres).query(coors)[1]
            for c in range(n_clusters):
                m = lab == c
                if m.any():
                    centres[c] = coors[m].mean(axis=0)
        lab = cKDTree(centres).query(coors)[1]
        # merge clusters into regions along a random direction (spatially coherent bands)
        ang = self.random.uniform(0, math.pi)
        proj = centres @ np.array([math.cos(ang), math.sin(ang)])
        order = np.argsort(proj)
        region_of_cluster = np.empty(n_clusters, dtype=np.int64)
        for r, chunk in enumerate(np.array_split(order, n_regions)):
            region_of_cluster[chunk] = r
        reg = region_of_cluster[lab]
        return [coors[reg == r] for r in range(n_regions)]

    def _assign_materials(self, regions: List[np.ndarray]) -> Dict[Tuple[float, float], np.ndarray]:
        out = {}
        for region in regions:
            if len(region) > 0:
                _, E, nu = self.random.choice(MATERIALS)
                out[(float(E), float(nu))] = region
        return out

    # ---- conditions (restates mesh_generator.py:397-521) -----------------------
    def sample_conditions(self, polygons_ptags, polygons_ltag_ptags, num_conditions: int = 4) -> List[Dict]:
        conditions = []
        all_ptags = [p for ptags in polygons_ptags for p in ptags]
        all_edges = [e for d in polygons_ltag_ptags for e in d.values()]
        n_v = len(self.mesh[0])
        while len(conditions) < num_conditions:
            ptags, edges = list(all_ptags), list(all_edges)
            sampled_edges = self.random.sample(edges, self.random.randint(1, len(edges) - 1))
            on_sampled = set()
            for e in sampled_edges:
                on_sampled.update(e)
            edges_to_constrain = self.random.sample(sampled_edges, self.random.randint(1, len(sampled_edges)))
            vertices_to_constrain = set(on_sampled)
            for e in edges_to_constrain:
                vertices_to_constrain.discard(e[0])
                vertices_to_constrain.discard(e[1])
            for e in edges_to_constrain:
                edges.remove(e)
            for v in on_sampled:
                ptags.remove(v)
            try:
                point_forces = self.random.sample(ptags, self.random.randint(1, len(ptags)))
            except ValueError:
                point_forces = []
            edge_forces = self.random.sample(edges, self.random.randint(0 if len(point_forces) >= 1 else 1, len(edges)))
            regions = self._create_regions_randomly()
            if sum(len(r) for r in regions) != n_v:
                continue
            materials = self._assign_materials(regions)
            if sum(len(r) for r in materials.values()) != n_v:
                continue  # two regions drew the same material (mesh_generator.py:476-480)
            cond = {"material_regions": materials, "point_constraints": list(vertices_to_constrain),
                    "edge_constraints": list(edges_to_constrain), "point_forces": list(point_forces),
                    "edge_forces": list(edge_forces)}
            conditions.append(cond)
        sign = [-1, 1]
        lo, hi = self.force_magnitude_range
        for cond in conditions:
            cond["point_forces"] = [(p, (self.random.randint(lo, hi) * self.random.choice(sign),
                                         self.random.randint(lo, hi) * self.random.choice(sign)))
                                    for p in cond["point_forces"]]
            cond["edge_forces"] = [(e, (self.random.randint(lo, hi) * self.random.choice(sign),
                                        self.random.randint(lo, hi) * self.random.choice(sign)))
                                   for e in cond["edge_forces"]]
        return conditions


def make_plate(seed: int, mesh_size: float = 1e-2, max_tries: int = 50, region_method: str = "lloyd"):
    """One plate: (generator with .mesh set, ptags, ltags).  Deterministic in ``seed``.  The benchmark
    workloads keep the round-1 ``"lloyd"`` regions (their seeds, iteration counts and the ill-conditioned
    plate the parity tests pin are defined on them); the dataset generator passes ``"reference"``."""
    gen = PlateGenerator(random_seed=seed, region_method=region_method)
    for _ in range(max_tries):
        try:
            g = gen.normalize_geometry(gen.generate_geometry())
            ptags, ltags = gen.generate_mesh(g, mesh_size)
            return gen, ptags, ltags
        except GeometryRejected:
            continue
    raise RuntimeError("no valid plate after %d tries (seed %d)" % (max_tries, seed))


def condition_kwargs(cond: Dict) -> Dict:
    """Condition dict -> FEAnalysis keyword arguments (reference generate.py:88-107)."""
    return dict(force_vertex_tags_magnitudes=cond["point_forces"], force_edges_tags_magnitudes=cond["edge_forces"],
                constraints_vertex_tags=cond["point_constraints"], constraints_edges_tags=cond["edge_constraints"],
                material_properties_to_vertices=cond["material_regions"])
