"""``generate_data`` with the reference's signature and on-disk results (reference
``datagen/generate.py:12-167``): plates x conditions loop, one ``FEAnalysis`` per condition, retry
of a condition whose solve fails, window sizing from the outline render of the first condition,
``TIME:`` bracket around ``calculate()`` only.  The loop is the reference's control flow restated
over the CUDA-backed ``FEAnalysis``; the throughput path for whole datasets is
``fea_diffusion_b200.dataset`` (batched, sharded over GPUs)."""
import os
from timeit import default_timer as timer
from typing import Optional, Tuple

from .fea_analysis import FEAnalysis
from .mesh_generator import MeshGenerator
from .utils import find_image_bounds, verify_directory


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
    random_seed: Optional[int] = None,
    verbose: bool = True,
    max_retries_per_condition: int = 200,
):
    assert num_steps_per_condition > 1, "Must have at least 2 steps per condition."
    verify_directory(data_dir)
    generator = MeshGenerator(
        num_polygons_range=num_polygons_range, points_per_polygon_range=points_per_polygon_range,
        holes_per_polygon_range=holes_per_polygon_range, points_per_hole_range=points_per_hole_range,
        num_regions=num_regions, random_seed=random_seed)
    assert num_plates >= 1
    assert conditions_per_plate >= 1
    if start_plate is not None:
        assert start_plate < num_plates and start_plate >= 0
    plate_index = (start_plate - 1) if (start_plate is not None) else 0
    plate_image_size, plate_bounds = None, None
    say = print if verbose else (lambda *a, **k: None)
    total_time = 0.0
    started = timer()
    while plate_index < num_plates:
        try:
            geometry = generator.normalize_geometry(generator.generate_geometry())
            polygons_ptags, polygons_ltag_ptags = generator.generate_mesh(
                geometry, os.path.join(data_dir, "part"), mesh_size=mesh_size, view_mesh=False)
        except Exception:
            continue
        conditions = generator.sample_conditions(polygons_ptags, polygons_ltag_ptags,
                                                 num_conditions=conditions_per_plate)
        plate_dir = os.path.join(data_dir, str(plate_index + 1))
        verify_directory(plate_dir)
        condition_index, retries = 0, 0
        while condition_index < len(conditions):
            condition_dir = os.path.join(plate_dir, str(condition_index + 1))
            verify_directory(condition_dir)
            cond = conditions[condition_index]
            analyzer = FEAnalysis(
                filename="part.mesh", data_dir=data_dir, condition_dir=condition_dir,
                force_vertex_tags_magnitudes=cond["point_forces"],
                force_edges_tags_magnitudes=cond["edge_forces"],
                constraints_vertex_tags=cond["point_constraints"],
                constraints_edges_tags=cond["edge_constraints"],
                material_properties_to_vertices=cond["material_regions"],
                num_steps=num_steps_per_condition, save_meshes=save_meshes)
            start = timer()
            success = analyzer.calculate()
            end = timer()
            if not success:
                retries += 1
                if retries > max_retries_per_condition:
                    raise RuntimeError("plate %d: no solvable condition in %d draws" % (plate_index + 1, retries))
                say("Failed to calculate for plate {} condition {}".format(plate_index + 1, condition_index + 1))
                say("Regenerating condition")
                analyzer.clear_condition_dir()
                conditions[condition_index] = generator.sample_conditions(
                    polygons_ptags, polygons_ltag_ptags, num_conditions=1)[0]
                continue
            retries = 0
            say("TIME:", end - start)
            total_time += end - start
            if condition_index == 0:
                outline_dir = os.path.join(plate_dir, "outline.png")
                analyzer.save_input_image(outline_dir, outline=True, crop=False)
                left, top, right, bottom = find_image_bounds(outline_dir)
                max_size = max(right - left, bottom - top)
                modified_image_size = round(image_size / (max_size / analyzer.initial_image_size))
                analyzer.update_image_size_or_bounds(image_size=modified_image_size)
                analyzer.save_input_image(outline_dir, outline=True, crop=False)
                left, top, right, bottom = find_image_bounds(outline_dir)
                lbound, ubound = (left, right) if right > bottom else (top, bottom)
                bounds = (lbound, lbound, ubound, ubound)
                analyzer.update_image_size_or_bounds(bounds=bounds)
                plate_image_size, plate_bounds = modified_image_size, bounds
                analyzer.save_input_image(os.path.join(plate_dir, "input.png"))
            else:
                analyzer.update_image_size_or_bounds(image_size=plate_image_size, bounds=plate_bounds)
            analyzer.save_region_images(os.path.join(condition_dir, "regions"))
            analyzer.save_output_images(os.path.join(condition_dir, "outputs"), save_displacement=save_displacement,
                                        save_strain=save_strain, save_stress=save_stress)
            condition_index += 1
        plate_index += 1
        if wandb_inject_function is not None:
            elapsed = timer() - started
            remaining = elapsed / plate_index * (num_plates - plate_index) if plate_index else 0
            wandb_inject_function(plate_index - 1, total_time, remaining)
        say("PLATE TIME:", total_time)
    say("TOTAL TIME:", total_time)
    return total_time


if __name__ == "__main__":
    generate_data()
