"""Drop-in for the reference's ``datagen.fea_analysis.FEAnalysis`` (reference
``datagen/fea_analysis.py:31-613``) over the CUDA library.

Same constructor, attributes and methods, same files written (``magnitudes.txt``,
``materials.txt``, ``ranges.txt``, ``domain.<k>.vtk``, ``regions.vtk``, the PNGs), so
``datagen/generate.py`` drives it unchanged.  What differs is where the work happens: the
reference builds an sfepy ``Problem`` and lets SuperLU / VTK do the arithmetic; here
``__init__`` hands the tags, magnitudes and material coordinate lists to the library, which selects
the regions and derives the Dirichlet mask, material cells, D and the load vector on the device
(``fea_batch_create_from_conditions``) and assembles the system; ``calculate()`` runs the Jacobi-PCG
solve through the C-ABI, and the image methods call the CUDA rasteriser.  All load steps come from one solve
(u_k = t_k * u_final).

One process-wide ``Context`` (GPU 0 unless ``FEA_B200_DEVICE`` says otherwise) is shared by all
instances; pass ``context=`` to use another.  The library is required: there is no CPU fallback.
"""
from __future__ import annotations

import math
import os
from os import path
from typing import Dict, List, Optional, Tuple

import numpy as np
from PIL import Image

from .. import imaging
from .._capi import SAMPLE_CONVERGED, SAMPLE_EMPTY_ROW
from ..host import read_mesh
from ..solver import Context, PackedConditions, Sample
from .vtk_io import domain_filename, write_vtk

_shared_context: Optional[Context] = None
_mesh_cache: Dict[Tuple[str, int, int], Tuple[np.ndarray, np.ndarray]] = {}


class DeviceSetupView:
    """What ``FEAnalysis.__init__`` derived for one condition, as downloaded from the device: same
    attributes as ``host.ProblemSetup`` (coors, conn, regions, sample, bbox)."""

    def __init__(self, coors, conn, names, dev):
        self.coors, self.conn = coors, conn
        self.regions = {nm: np.flatnonzero(dev.region_flags[0][i]) for i, nm in enumerate(names)}
        self.region_count = dev.region_count[0]
        self.sample = Sample(coors=coors, conn=conn, cell_region=dev.cell_region, D=dev.D[0],
                             fixed=dev.fixed.astype(bool), rhs=dev.rhs)

    def bbox(self):
        c = self.coors
        return float(c[:, 0].min()), float(c[:, 1].min()), float(c[:, 0].max()), float(c[:, 1].max())


def default_context() -> Context:
    global _shared_context
    if _shared_context is None or _shared_context.h is None:
        _shared_context = Context(int(os.environ.get("FEA_B200_DEVICE", "0")))
    return _shared_context


def _load_mesh(filepath: str):
    """Mesh arrays, cached per (file, mtime, size): every condition of a plate re-opens the same
    ``part.mesh`` (reference generate.py:88-90)."""
    st = os.stat(filepath)
    key = (path.abspath(filepath), st.st_mtime_ns, st.st_size)
    hit = _mesh_cache.get(key)
    if hit is None:
        coors, conn = read_mesh(filepath)
        hit = (coors, conn)
        if len(_mesh_cache) >= 8:
            _mesh_cache.clear()
        _mesh_cache[key] = hit
    return hit


class FEAnalysis:
    def __init__(
        self,
        filename: str,
        data_dir: str,
        condition_dir: str,
        force_vertex_tags_magnitudes: List[Tuple[int, Tuple[float, float]]],
        force_edges_tags_magnitudes: List[Tuple[Tuple[int, int], Tuple[int, int]]],
        constraints_vertex_tags: List[int],
        constraints_edges_tags: List[Tuple[int, int]],
        num_steps: int = 11,
        save_meshes: bool = False,
        material_properties_to_vertices: Optional[Dict[Tuple[float, float], List[Tuple[float, float]]]] = None,
        youngs_modulus: Optional[float] = 210000,
        poisson_ratio: Optional[float] = 0.3,
        *,
        context: Optional[Context] = None,
        rtol: float = 1e-10,
        max_iter: int = 50000,
        strict: bool = True,
    ):
        """Positional/keyword arguments as in the reference (``fea_analysis.py:32-48``).
        Keyword-only extras: ``rtol`` / ``max_iter`` of the PCG solve; ``strict`` makes
        ``calculate()`` return False for singular (floating-region) conditions too --
        the reference writes SuperLU noise for those and only rejects NaN."""
        self.data_dir = data_dir
        self.region_filename = "regions"
        self.save_meshes = save_meshes
        self.condition_dir = condition_dir
        self.initial_image_size = math.ceil(512 / 0.685546875)
        self.image_size = self.initial_image_size
        self.bounds = (0, 0, self.initial_image_size, self.initial_image_size)
        self.common_config = ("-2 --color-map binary --no-scalar-bars --no-axes --window-size {},{} --off-screen"
                              .format(self.initial_image_size, self.initial_image_size))
        if material_properties_to_vertices is None:
            assert youngs_modulus is not None and poisson_ratio is not None, (
                "If material_properties_to_vertices is not provided, youngs_modulus and poisson_ratio must be provided")
        self.num_steps = num_steps
        self.context = context
        self.rtol, self.max_iter, self.strict = rtol, max_iter, strict

        coors, conn = _load_mesh(path.join(data_dir, filename))
        kw = dict(force_vertex_tags_magnitudes=force_vertex_tags_magnitudes,
                  force_edges_tags_magnitudes=force_edges_tags_magnitudes,
                  constraints_vertex_tags=constraints_vertex_tags, constraints_edges_tags=constraints_edges_tags,
                  material_properties_to_vertices=material_properties_to_vertices,
                  youngs_modulus=youngs_modulus, poisson_ratio=poisson_ratio)
        # region selection, Dirichlet mask, material cells, D and load on the device; the batch (one
        # sample) stays assembled until calculate()
        self._packed = PackedConditions([(coors, conn)], [(0, kw)])
        self._batch = self._ctx().create_batch_from_conditions(self._packed)
        self._batch.assemble()
        self.setup = DeviceSetupView(coors, conn, self._packed.names[0], self._batch.setup())
        # side effects of the reference constructor (fea_analysis.py:87-91, 108-115, 278-282)
        from ..dataset import text_lines
        magnitudes, materials = text_lines(kw, self.setup.region_count)
        for name, text in (("magnitudes.txt", magnitudes), ("materials.txt", materials)):
            if text:
                with open(path.join(self.condition_dir, name), "a+") as f:
                    f.write(text)
        self.displacement: Optional[np.ndarray] = None   # (num_steps, n_v, 2) after calculate()
        self.cell_strain: Optional[np.ndarray] = None    # (n_cell, 3) final step
        self.cell_stress: Optional[np.ndarray] = None
        self.status: Optional[int] = None
        self.iterations: Optional[int] = None

    # ---- small helpers --------------------------------------------------------------------
    def _ctx(self) -> Context:
        return self.context if self.context is not None else default_context()

    def _append_line(self, filename: str, line: str):
        with open(path.join(self.condition_dir, filename), "a+") as f:
            f.write(line + "\n")

    def _append_region_value_to_file(self, filename: str, region_name: str, value):
        self._append_line(filename, "{}:{}".format(region_name, str(value)))

    def _output_dir(self) -> str:
        return self.condition_dir if self.save_meshes else self.data_dir

    @property
    def times(self) -> np.ndarray:
        return np.linspace(0.0, 1.0, self.num_steps)

    def __del__(self):
        b = getattr(self, "_batch", None)
        if b is not None:
            try:
                b.destroy()
            except Exception:
                pass

    def clear_condition_dir(self):
        for file in os.listdir(self.condition_dir):
            os.remove(path.join(self.condition_dir, file))

    @staticmethod
    def crop_image(image_path, bounds):
        image = Image.open(image_path)
        image = image.crop(bounds)
        image.save(image_path)

    # the two selectors are part of the reference's surface (static methods); kept for callers
    @staticmethod
    def _get_points_on_edge(coords, bounding_tags, **_):
        from ..host import vertices_on_line
        return vertices_on_line(np.asarray(coords), bounding_tags)

    @staticmethod
    def _get_points_in_list(coords, region_coordinates, **_):
        from ..host import vertices_in_list
        return vertices_in_list(np.asarray(coords), region_coordinates)

    # ---- solve ----------------------------------------------------------------------------
    def calculate(self) -> bool:
        """Replaces Problem + Newton + ScipyDirect + SimpleTimeSteppingSolver
        (``fea_analysis.py:418-461``).  Returns False when the final displacement has NaN (the
        reference's only failure signal) or, with ``strict``, when the system is singular."""
        smp = self.setup.sample
        b = self._batch
        if b is None or b.h is None:                 # calculate() called again: set the problem up anew
            b = self._ctx().create_batch_from_conditions(self._packed)
            b.assemble()
        try:
            b.solve(self.rtol, self.max_iter)
            res = b.download()
            stress_region = 0 if len(smp.D) else -1
            strain, stress = b.cell_strain_stress(stress_region)
        finally:
            b.destroy()
            self._batch = None
        self.status, self.iterations = int(res.status[0]), int(res.iters[0])
        self.relres = float(res.relres[0])
        u = res.u
        self.displacement = self.times[:, None, None] * u[None]
        self.cell_strain, self.cell_stress = strain, stress
        self._save_regions()
        self._save_steps()
        if np.isnan(u).any() or self.status == SAMPLE_EMPTY_ROW:
            return False
        if self.strict and self.status != SAMPLE_CONVERGED:
            return False
        return True

    def _region_flags(self) -> Dict[str, np.ndarray]:
        n_v = len(self.setup.coors)
        out = {"Omega": np.ones(n_v)}
        for name, verts in self.setup.regions.items():
            f = np.zeros(n_v)
            f[verts] = 1.0
            out[name] = f
        return out

    def _save_regions(self):
        """``problem.save_regions_as_groups`` (``fea_analysis.py:377-381``): one 0/1 vertex field
        per region in ``regions.vtk``."""
        directory = self._output_dir()
        target = path.join(directory, self.region_filename + ".vtk")
        if path.isfile(target):
            os.remove(target)
        n_cell = len(self.setup.conn)
        write_vtk(target, self.setup.coors, self.oriented_conn(), point_data=self._region_flags(),
                  cell_data={"mat_id": np.zeros(n_cell, dtype=np.int64)})

    def oriented_conn(self) -> np.ndarray:
        """Connectivity with clockwise cells fixed the way sfepy does on load (A-2)."""
        co, cn = self.setup.coors, self.setup.conn.copy()
        x, y = co[cn, 0], co[cn, 1]
        k = cn.shape[1]
        area2 = sum(x[:, a] * y[:, (a + 1) % k] - x[:, (a + 1) % k] * y[:, a] for a in range(k))
        cw = area2 < 0
        if k == 3:
            cn[cw] = cn[cw][:, [0, 2, 1]]
        else:
            cn[cw] = cn[cw][:, [0, 3, 2, 1]]
        return cn

    def _save_steps(self):
        """One ``domain.<k>.vtk`` per load step with u, node_groups, mat_id, cauchy_strain,
        cauchy_stress (``problem.solve(save_results=True, post_process_hook=...)``,
        ``fea_analysis.py:436-439``)."""
        directory = self._output_dir()
        co, cn = self.setup.coors, self.oriented_conn()
        n_v, n_cell = len(co), len(cn)
        groups = np.zeros(n_v, dtype=np.int64)
        mat_id = np.zeros(n_cell, dtype=np.int64)
        for k, t in enumerate(self.times):
            write_vtk(path.join(directory, domain_filename(k, self.num_steps)), co, cn,
                      point_data={"u": self.displacement[k], "node_groups": groups},
                      cell_data={"cauchy_strain": t * self.cell_strain, "cauchy_stress": t * self.cell_stress,
                                 "mat_id": mat_id})

    # ---- images ---------------------------------------------------------------------------
    def update_image_size_or_bounds(self, image_size=None, bounds=None):
        if image_size is not None:
            self.image_size = image_size
            self.common_config = ("-2 --color-map binary --no-scalar-bars --no-axes --window-size {},{} --off-screen"
                                  .format(image_size, image_size))
        if bounds is not None:
            self.bounds = bounds

    def _window_affine(self) -> np.ndarray:
        return imaging.crop_affine(self.setup.bbox(), self.image_size, (0, 0, self.image_size, self.image_size))

    def _render(self, fields: np.ndarray, clim, cell_fields: bool = False) -> np.ndarray:
        return self._ctx().rasterize_fields(self.setup.coors, self.setup.conn, fields, clim, self._window_affine(),
                                            self.image_size, cell_fields=cell_fields)

    @staticmethod
    def _write_png(gray: np.ndarray, filepath: str):
        Image.fromarray(np.repeat(gray[:, :, None], 3, axis=2)).save(filepath)

    def save_input_image(self, filepath, input_filepath=None, outline=False, crop=True):
        """``fields=[("1", "vs")]`` render (``fea_analysis.py:472-506``): the plate in black, or --
        with ``outline`` -- the bounding-box outline that ``find_image_bounds`` measures."""
        size = self.image_size
        if outline:
            gray = np.full((size, size), 255, np.uint8)
            left, top, right, bottom = imaging.outline_bounds(size, self.setup.bbox())
            l, t, r, b = (max(0, min(size - 1, v)) for v in (left, top, right, bottom))
            for w in (0, 1):  # 2 px lines, drawn inwards; the outer row/column is what the bounds see
                gray[min(t + w, b), l:r + 1] = 186
                gray[max(b - w, t), l:r + 1] = 186
                gray[t:b + 1, min(l + w, r)] = 186
                gray[t:b + 1, max(r - w, l)] = 186
        else:
            gray = self._render(np.ones((1, len(self.setup.coors))), [(0.0, 1.0)])[0]
        self._write_png(gray, filepath)
        if crop:
            self.crop_image(filepath, self.bounds)

    def save_region_images(self, filepathroot, crop=True):
        """``regions_<Region>.png`` for every region except Omega (``fea_analysis.py:508-524``)."""
        flags = self._region_flags()
        names = [n for n in flags if "Omega" not in n]
        if not names:
            return
        images = self._render(np.stack([flags[n] for n in names]), [(0.0, 1.0)] * len(names))
        for name, gray in zip(names, images):
            filepath = "{}_{}.png".format(filepathroot, name)
            self._write_png(gray, filepath)
            if crop:
                self.crop_image(filepath, self.bounds)

    def save_output_images(self, filepathroot, save_displacement=True, save_stress=True, save_strain=True, crop=True):
        """Step-1 images + ``ranges.txt`` lines for every step and type
        (``fea_analysis.py:526-613``, ``custom_plotter.py:121-193``)."""
        if self.displacement is None:
            raise RuntimeError("save_output_images() before calculate()")
        kinds = []  # (type name, final-step values, per-cell?)
        u = self.displacement[-1]
        if save_displacement:
            kinds += [("displacement_x", u[:, 0], False), ("displacement_y", u[:, 1], False)]
        if save_stress:
            kinds += [("stress_x", self.cell_stress[:, 0], True), ("stress_y", self.cell_stress[:, 1], True)]
        if save_strain:
            kinds += [("strain_x", self.cell_strain[:, 0], True), ("strain_y", self.cell_strain[:, 1], True)]
        times = self.times
        lines = []
        for step in range(1, self.num_steps):
            for name, values, _ in kinds:
                lo, hi = float((times[step] * values).min()), float((times[step] * values).max())
                lines.append("{}_{}:{}".format(name, step, str((lo, hi))))
        with open(path.join(self.condition_dir, "ranges.txt"), "a+") as f:
            f.write("".join(line + "\n" for line in lines))
        for per_cell in (False, True):
            group = [(n, v) for n, v, c in kinds if c == per_cell]
            if not group:
                continue
            fields = np.stack([times[1] * v for _, v in group])
            clim = [(float(f.min()), float(f.max())) for f in fields]
            images = self._render(fields, clim, cell_fields=per_cell)
            for (name, _), gray in zip(group, images):
                filepath = "{}_{}.png".format(filepathroot, name)
                self._write_png(gray, filepath)
                if crop:
                    self.crop_image(filepath, self.bounds)
