"""Mesh file I/O for the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

* ``read_medit``  -- MEDIT ``.mesh`` text files as gmsh 4.11 writes them and
  sfepy reads them at ``datagen/fea_analysis.py:61`` (reference); rules
  SURVEY.md A-1 (z == 0 dropped, only highest-dimension cells kept).
* ``read_vtk_legacy`` / ``write_vtk_legacy`` -- legacy-VTK 4.2 BINARY files in
  the layout meshio 4.4.6 emits for sfepy; the reader is
  how the golden ``u`` of ``applications/*/**.vtk`` is loaded.
"""
from __future__ import annotations

import numpy as np

_MEDIT_CELLS = {"Edges": 2, "Triangles": 3, "Quadrilaterals": 4, "Tetrahedra": 4, "Hexahedra": 8}


def read_medit(path):
    """Return dict(coors (n_v,2) f64, conn (n_cell,k) i32 0-based, k, vertex_ref, cell_ref).

    Only the highest-dimension 2-D cell block is kept (``Edges`` ignored);
    if both Triangles and Quadrilaterals exist the larger block wins.
    """
    with open(path, "r") as f:
        tok = f.read().split()
    i = 0
    dim = 3
    verts = None
    blocks = {}
    n = len(tok)
    while i < n:
        t = tok[i]
        if t == "MeshVersionFormatted":
            i += 2
        elif t == "Dimension":
            dim = int(tok[i + 1])
            i += 2
        elif t == "Vertices":
            nv = int(tok[i + 1])
            w = dim + 1
            arr = np.array(tok[i + 2 : i + 2 + nv * w], dtype=np.float64).reshape(nv, w)
            verts = arr
            i += 2 + nv * w
        elif t in _MEDIT_CELLS:
            k = _MEDIT_CELLS[t]
            nc = int(tok[i + 1])
            arr = np.array(tok[i + 2 : i + 2 + nc * (k + 1)], dtype=np.int64).reshape(nc, k + 1)
            blocks[t] = arr
            i += 2 + nc * (k + 1)
        elif t == "End":
            break
        else:
            i += 1
    if verts is None:
        raise ValueError("no Vertices section in %s" % path)
    xyz = verts[:, :dim]
    vref = verts[:, dim].astype(np.int64)
    # A-1: effective dimension = axes with non-zero extent
    ext = xyz.max(axis=0) - xyz.min(axis=0)
    keep = [a for a in range(dim) if ext[a] > 1e-15]
    if len(keep) != 2:
        raise ValueError("oracle handles planar meshes only (got %d non-flat axes)" % len(keep))
    coors = np.ascontiguousarray(xyz[:, keep])
    cand = [(name, blocks[name]) for name in ("Triangles", "Quadrilaterals") if name in blocks]
    if not cand:
        raise ValueError("no 2-D cells in %s" % path)
    name, blk = max(cand, key=lambda p: len(p[1]))
    conn = np.ascontiguousarray(blk[:, :-1] - 1).astype(np.int32)
    cref = blk[:, -1].copy()
    return dict(coors=coors, conn=conn, k=conn.shape[1], vertex_ref=vref, cell_ref=cref)


_VTK_DT = {"double": ">f8", "float": ">f4", "long": ">i8", "int": ">i4", "vtktypeint64": ">i8",
           "vtktypeint32": ">i4", "unsigned_char": ">u1"}


class _Cur:
    def __init__(self, buf):
        self.b = buf
        self.p = 0

    def line(self):
        e = self.b.index(b"\n", self.p)
        s = self.b[self.p:e].decode("ascii", "replace").strip()
        self.p = e + 1
        return s

    def nonempty(self):
        while self.p < len(self.b):
            s = self.line()
            if s:
                return s
        return None

    def raw(self, dtype, count):
        dt = np.dtype(dtype)
        nb = dt.itemsize * count
        a = np.frombuffer(self.b, dtype=dt, count=count, offset=self.p)
        self.p += nb
        return a


def read_vtk_legacy(path):
    """Read a BINARY legacy-VTK unstructured grid (meshio 4.4.6 layout).

    Returns dict(points (n,3), cells (n_cell,k) i32, cell_types, point_data{}, cell_data{}).
    """
    with open(path, "rb") as f:
        cur = _Cur(f.read())
    cur.line()
    cur.line()
    if cur.line().upper() != "BINARY":
        raise ValueError("only BINARY legacy VTK supported")
    cur.line()
    out = dict(point_data={}, cell_data={})
    target = None
    while True:
        s = cur.nonempty()
        if s is None:
            break
        w = s.split()
        key = w[0].upper()
        if key == "POINTS":
            n = int(w[1])
            out["points"] = cur.raw(_VTK_DT[w[2].lower()], 3 * n).astype(np.float64).reshape(n, 3)
        elif key == "CELLS":
            nc, size = int(w[1]), int(w[2])
            flat = cur.raw(">i4", size).astype(np.int32)
            k = size // nc - 1
            out["cells"] = flat.reshape(nc, k + 1)[:, 1:].copy()
        elif key == "CELL_TYPES":
            out["cell_types"] = cur.raw(">i4", int(w[1])).astype(np.int32)
        elif key == "POINT_DATA":
            target = out["point_data"]
        elif key == "CELL_DATA":
            target = out["cell_data"]
        elif key == "FIELD":
            for _ in range(int(w[2])):
                h = cur.nonempty().split()
                name, ncomp, ntup, dt = h[0], int(h[1]), int(h[2]), h[3].lower()
                a = cur.raw(_VTK_DT[dt], ncomp * ntup)
                target[name] = a.astype(a.dtype.newbyteorder("=")).reshape(ntup, ncomp)
        else:
            raise ValueError("unhandled VTK section %r" % s)
    return out


def write_vtk_legacy(path, coors2d, conn, point_data=None, cell_data=None):
    """Write the meshio-4.4.6 style file sfepy produces for a 2-D P1/Q1 mesh.

    ``point_data`` / ``cell_data``: ordered dict name -> (n, ncomp) array;
    float arrays are written as ``double``, integer arrays as ``long``.
    Vector point data with 2 components is zero-padded to 3 (A-15).
    """
    coors2d = np.asarray(coors2d, dtype=np.float64)
    conn = np.asarray(conn)
    n, nc, k = len(coors2d), len(conn), conn.shape[1]
    pts = np.zeros((n, 3), dtype=">f8")
    pts[:, :2] = coors2d
    cells = np.empty((nc, k + 1), dtype=">i4")
    cells[:, 0] = k
    cells[:, 1:] = conn
    ctype = 5 if k == 3 else 9

    def _field(f, data):
        f.write(("FIELD FieldData %d\n" % len(data)).encode())
        for name, a in data.items():
            a = np.asarray(a)
            if a.ndim == 1:
                a = a[:, None]
            if np.issubdtype(a.dtype, np.floating):
                if a.shape[1] == 2:
                    a = np.concatenate([a, np.zeros((len(a), 1))], axis=1)
                dt, tag = ">f8", "double"
            else:
                dt, tag = ">i8", "long"
            f.write(("%s %d %d %s\n" % (name, a.shape[1], a.shape[0], tag)).encode())
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
            f.write(b"\n")

    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 4.2\nwritten by meshio v4.4.6\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(("POINTS %d double\n" % n).encode())
        f.write(pts.tobytes())
        f.write(b"\n")
        f.write(("CELLS %d %d\n" % (nc, nc * (k + 1))).encode())
        f.write(cells.tobytes())
        f.write(b"\n")
        f.write(("CELL_TYPES %d\n" % nc).encode())
        f.write(np.full(nc, ctype, dtype=">i4").tobytes())
        f.write(b"\n")
        if point_data:
            f.write(("POINT_DATA %d\n" % n).encode())
            _field(f, point_data)
        if cell_data:
            f.write(("CELL_DATA %d\n" % nc).encode())
            _field(f, cell_data)
