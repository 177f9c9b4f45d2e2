"""Synthetic workloads for bench.py and the large-solve tests: BASELINE.json's configurations
built from the seeded plate generator."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from . import imaging
from .host import MeshTopology, ProblemSetup, floating_components
from .plates import condition_kwargs, make_plate


@dataclass
class WorkItem:
    plate: int
    condition: int
    setup: ProblemSetup
    kwargs: Dict                 # FEAnalysis-style arguments (for the CPU oracle)
    window: int                  # modified_image_size of the plate
    bounds: Tuple[int, int, int, int]
    affine: np.ndarray           # world -> crop pixel
    size: int                    # crop size (pixels)


def plate_conditions(seed: int, conditions_per_plate: int, image_size: int, mesh_size: float = 1e-2,
                     well_posed: bool = True, max_draws: int = 400):
    """All conditions of one plate.  With ``well_posed`` the condition sampler is redrawn until
    every stiffness-connected component carries >= 2 fixed vertices and no active vertex is
    isolated (SURVEY A-19); the number of rejected draws is returned for the report."""
    gen, ptags, ltags = make_plate(seed, mesh_size)
    coors, conn = gen.mesh
    topo = MeshTopology(conn, len(coors))
    bbox = (float(coors[:, 0].min()), float(coors[:, 1].min()), float(coors[:, 0].max()), float(coors[:, 1].max()))
    window, bounds = imaging.plate_window(bbox, image_size)
    affine = imaging.crop_affine(bbox, window, bounds)
    items: List[WorkItem] = []
    rejected = 0
    draws = 0
    while len(items) < conditions_per_plate:
        if draws >= max_draws:
            raise RuntimeError("plate %d: no well-posed condition in %d draws" % (seed, max_draws))
        cond = gen.sample_conditions(ptags, ltags, 1)[0]
        draws += 1
        kw = condition_kwargs(cond)
        setup = ProblemSetup(coors, conn, topology=topo, **kw)
        if well_posed and floating_components(setup.sample) != (0, 0):
            rejected += 1
            continue
        items.append(WorkItem(seed, len(items), setup, kw, window, bounds, affine, bounds[2] - bounds[0]))
    return items, rejected


def build_workload(n_plates: int, conditions_per_plate: int = 4, image_size: int = 64, seed0: int = 0,
                   mesh_size: float = 1e-2, well_posed: bool = True):
    items: List[WorkItem] = []
    rejected = 0
    for p in range(n_plates):
        it, rj = plate_conditions(seed0 + p, conditions_per_plate, image_size, mesh_size, well_posed)
        items.extend(it)
        rejected += rj
    return items, rejected


def refine_uniform(coors: np.ndarray, conn: np.ndarray, levels: int):
    """Uniform red refinement (each triangle -> 4; midpoints appended after the existing
    vertices), the mesh family of BASELINE config 3 (SURVEY C-5; the reference's
    ``refine_mesh`` hook, applications/cantilever/cantilever.py:26-27)."""
    coors = np.asarray(coors, dtype=np.float64)
    conn = np.asarray(conn, dtype=np.int64)
    for _ in range(levels):
        n = len(coors)
        e = np.concatenate([conn[:, [0, 1]], conn[:, [1, 2]], conn[:, [2, 0]]])
        key = np.minimum(e[:, 0], e[:, 1]) * n + np.maximum(e[:, 0], e[:, 1])
        uniq, inv = np.unique(key, return_inverse=True)
        mid = 0.5 * (coors[uniq // n] + coors[uniq % n])
        m = (n + inv).reshape(3, -1).T  # midpoints of edges (01, 12, 20) per cell
        a, b, c = conn[:, 0], conn[:, 1], conn[:, 2]
        ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
        conn = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1),
                               np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)])
        coors = np.concatenate([coors, mid])
    return np.ascontiguousarray(coors), np.ascontiguousarray(conn.astype(np.int32))
