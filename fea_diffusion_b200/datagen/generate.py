"""``generate_data`` of the datagen package.

The plates x conditions loop (reference ``datagen/generate.py:56-164``) is the CALLER of the hot
path, not part of it, and is not restated here.  Two ways to drive the CUDA path with the
reference's call:

* **the reference's own file, unchanged** -- ``load_reference_generate(reference_dir)`` executes a
  checkout's ``datagen/generate.py`` as a module of THIS package, so that its relative imports
  (``.fea_analysis``, ``.mesh_generator``, ``.utils``; reference generate.py:7-9) bind to the drop-in
  classes here.  ``tests/test_caller_surface.py`` checks every call that file makes against the
  drop-in signatures.  ``generate_data`` uses it when ``FEA_REFERENCE_DIR`` names a checkout.
* **the batched generator** -- otherwise ``generate_data`` forwards the reference's keyword
  arguments to ``fea_diffusion_b200.dataset.generate_dataset`` (same dataset tree, conditions of a
  plate solved together, plates sharded over GPUs).
"""
import importlib.util
import os
import sys
from typing import Optional, Tuple


def load_reference_generate(reference_dir: str):
    """The reference's ``generate_data`` function, from its own source file, bound to this package's
    ``FEAnalysis`` / ``MeshGenerator`` / ``find_image_bounds`` / ``verify_directory``."""
    path = os.path.join(reference_dir, "datagen", "generate.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    name = __package__ + "._reference_generate"
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod.generate_data


def generate_data(
    data_dir: str = "data/",
    image_size: int = 512,
    num_plates: int = 1,
    start_plate: Optional[int] = None,
    conditions_per_plate: int = 4,
    mesh_size: float = 1e-2,
    num_polygons_range: Tuple[int, int] = (1, 3),
    points_per_polygon_range: Tuple[int, int] = (3, 8),
    holes_per_polygon_range: Tuple[int, int] = (0, 3),
    points_per_hole_range: Tuple[int, int] = (3, 4),
    num_regions: Tuple[int, int] = (1, 5),
    save_displacement: bool = True,
    save_strain: bool = False,
    save_stress: bool = False,
    num_steps_per_condition: int = 11,
    save_meshes: bool = False,
    wandb_inject_function=None,
    **batched_options,
):
    """Keyword arguments of the reference (generate.py:12-31); ``batched_options`` (seed, rank, world,
    plates_per_batch, workers, ...) go to ``dataset.generate_dataset``."""
    ref = os.environ.get("FEA_REFERENCE_DIR")
    if ref:
        return load_reference_generate(ref)(
            data_dir=data_dir, image_size=image_size, num_plates=num_plates, start_plate=start_plate,
            conditions_per_plate=conditions_per_plate, mesh_size=mesh_size, num_polygons_range=num_polygons_range,
            points_per_polygon_range=points_per_polygon_range, holes_per_polygon_range=holes_per_polygon_range,
            points_per_hole_range=points_per_hole_range, num_regions=num_regions, save_displacement=save_displacement,
            save_strain=save_strain, save_stress=save_stress, num_steps_per_condition=num_steps_per_condition,
            save_meshes=save_meshes, wandb_inject_function=wandb_inject_function)
    from ..dataset import generate_dataset
    defaults = ((1, 3), (3, 8), (0, 3), (3, 4), (1, 5))
    given = (num_polygons_range, points_per_polygon_range, holes_per_polygon_range, points_per_hole_range, num_regions)
    if tuple(map(tuple, given)) != defaults:
        raise NotImplementedError("the batched generator draws plates with the reference's default geometry ranges")
    start = max(0, (start_plate - 1) if start_plate is not None else 0)   # reference generate.py:50
    progress = None
    if wandb_inject_function is not None:
        progress = lambda done, total: wandb_inject_function(done - 1, 0.0, 0.0)   # (plate_index, total_time, remaining)
    st = generate_dataset(data_dir, num_plates, conditions_per_plate, image_size, num_steps_per_condition, mesh_size,
                          start_plate=start, save_meshes=save_meshes, save_displacement=save_displacement,
                          save_stress=save_stress, save_strain=save_strain, progress=progress, **batched_options)
    return st["gpu_s"]
