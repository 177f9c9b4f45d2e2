"""Legacy-VTK (4.2, BINARY, big-endian) unstructured-grid files in the exact byte layout the
reference's outputs have: sfepy 2023.3 writes ``domain.<k>.vtk`` / ``regions.vtk`` through
meshio 4.4.6 (reference ``datagen/fea_analysis.py:377-381, 434-439``; layout pinned by
``applications/cantilever/cantilever.vtk``).  Consumers:
``metrics/calculate_accuracy.py:48-58`` and ``test_scripts/setscale.py:14`` read ``u`` back.

Section order: POINTS (double, z = 0) / CELLS (int32: k a b c ..) / CELL_TYPES (5 triangle,
9 quad) / POINT_DATA + FIELD / CELL_DATA + FIELD; every binary block is followed by "\\n";
float fields are ``double``, integer fields ``long`` (int64); 2-component vectors are padded to 3.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

_HEADER = b"# vtk DataFile Version 4.2\nwritten by meshio v4.4.6\nBINARY\nDATASET UNSTRUCTURED_GRID\n"
_TYPES = {"double": ">f8", "float": ">f4", "long": ">i8", "int": ">i4", "vtktypeint64": ">i8",
          "vtktypeint32": ">i4", "unsigned_char": ">u1"}


def _field_block(data: Dict[str, np.ndarray]) -> bytes:
    parts = [b"FIELD FieldData %d\n" % len(data)]
    for name, values in data.items():
        a = np.asarray(values)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if np.issubdtype(a.dtype, np.floating):
            if a.shape[1] == 2:  # vectors of a 2-D problem carry a zero z component (A-15)
                a = np.column_stack([a, np.zeros(len(a))])
            kind, dt = b"double", ">f8"
        else:
            kind, dt = b"long", ">i8"
        parts.append(b"%s %d %d %s\n" % (name.encode("ascii"), a.shape[1], a.shape[0], kind))
        parts.append(np.ascontiguousarray(a, dtype=dt).tobytes())
        parts.append(b"\n")
    return b"".join(parts)


def vtk_bytes(coors: np.ndarray, conn: np.ndarray, point_data: Optional[Dict[str, np.ndarray]] = None,
              cell_data: Optional[Dict[str, np.ndarray]] = None) -> bytes:
    coors = np.asarray(coors, dtype=np.float64)
    conn = np.asarray(conn)
    n_v, (n_cell, k) = len(coors), conn.shape
    xyz = np.zeros((n_v, 3), dtype=">f8")
    xyz[:, :coors.shape[1]] = coors
    rows = np.empty((n_cell, k + 1), dtype=">i4")
    rows[:, 0] = k
    rows[:, 1:] = conn
    out = [_HEADER,
           b"POINTS %d double\n" % n_v, xyz.tobytes(), b"\n",
           b"CELLS %d %d\n" % (n_cell, n_cell * (k + 1)), rows.tobytes(), b"\n",
           b"CELL_TYPES %d\n" % n_cell, np.full(n_cell, 5 if k == 3 else 9, dtype=">i4").tobytes(), b"\n"]
    if point_data:
        out += [b"POINT_DATA %d\n" % n_v, _field_block(point_data)]
    if cell_data:
        out += [b"CELL_DATA %d\n" % n_cell, _field_block(cell_data)]
    return b"".join(out)


def write_vtk(path: str, coors, conn, point_data=None, cell_data=None) -> None:
    with open(path, "wb") as f:
        f.write(vtk_bytes(coors, conn, point_data, cell_data))


def read_vtk(path: str) -> Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """(points (n,3), cells (n_cell,k), point_data, cell_data) of a file written by ``write_vtk``
    or by the reference (uniform cell type only)."""
    buf = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        while True:
            end = buf.index(b"\n", pos)
            text = buf[pos:end].decode("ascii", "replace").strip()
            pos = end + 1
            if text:
                return text

    def block(dtype, count):
        nonlocal pos
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes
        return a

    line(), line()
    if line().upper() != "BINARY":
        raise ValueError("%s: only BINARY legacy VTK is supported" % path)
    line()
    points = cells = None
    pdata: Dict[str, np.ndarray] = {}
    cdata: Dict[str, np.ndarray] = {}
    target = pdata
    while pos < len(buf) and buf[pos:].strip():
        w = line().split()
        key = w[0].upper()
        if key == "POINTS":
            points = block(_TYPES[w[2].lower()], 3 * int(w[1])).astype(np.float64).reshape(-1, 3)
        elif key == "CELLS":
            n, total = int(w[1]), int(w[2])
            cells = block(">i4", total).astype(np.int32).reshape(n, total // max(n, 1))[:, 1:]
        elif key == "CELL_TYPES":
            block(">i4", int(w[1]))
        elif key == "POINT_DATA":
            target = pdata
        elif key == "CELL_DATA":
            target = cdata
        elif key == "FIELD":
            for _ in range(int(w[2])):
                name, ncomp, ntup, kind = line().split()
                a = block(_TYPES[kind.lower()], int(ncomp) * int(ntup))
                target[name] = a.astype(a.dtype.newbyteorder("=")).reshape(int(ntup), int(ncomp))
        else:
            raise ValueError("%s: unsupported section %r" % (path, " ".join(w)))
    return points, cells, pdata, cdata


def domain_filename(step: int, num_steps: int) -> str:
    """``domain.{k}.vtk`` with the step suffix as wide as the digits of num_steps - 1 (sfepy's
    time-stepping output, A-14; the reference re-derives the same names at fea_analysis.py:473-476,
    586-589)."""
    return "domain.%0*d.vtk" % (len(str(num_steps - 1)), step)
