"""The drop-in ``FEAnalysis`` / ``generate_data`` surface on the GPU, checked against the CPU oracle
and the reference's committed text files."""
import os

import numpy as np
import pytest
from PIL import Image

import cases
from fea_diffusion_b200 import imaging
from fea_diffusion_b200.datagen import FEAnalysis, generate_data
from fea_diffusion_b200.datagen.mesh_generator import write_medit
from fea_diffusion_b200.datagen.vtk_io import read_vtk
from fea_diffusion_b200.host import read_mesh
from oracle import raster_oracle as ro
from oracle.fea_oracle import OracleProblem

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_feanalysis_cantilever_files_and_fields(tmp_path, golden):
    data_dir, cond_dir = str(tmp_path), str(tmp_path / "1")
    os.makedirs(cond_dir)
    write_medit(os.path.join(data_dir, "part.mesh"), golden["cantilever_coors"], golden["cantilever_conn"])
    kw = dict(force_vertex_tags_magnitudes=[(4, (0, -1000))], force_edges_tags_magnitudes=[((3, 4), (120, 0))],
              constraints_vertex_tags=[], constraints_edges_tags=[(1, 2)])
    an = FEAnalysis("part.mesh", data_dir, cond_dir, num_steps=5, save_meshes=True, **kw)
    assert an.initial_image_size == 747 and an.image_size == 747 and an.bounds == (0, 0, 747, 747)
    assert an.calculate() is True
    orc = OracleProblem(golden["cantilever_coors"], golden["cantilever_conn"], num_steps=5, **kw)
    u = orc.solve("reference")
    for k in range(5):
        if k:
            assert rel(an.displacement[k], u[k]) <= 1e-8
        pts, cells, pd, cd = read_vtk(os.path.join(cond_dir, "domain.%d.vtk" % k))
        assert np.array_equal(pd["u"][:, :2], an.displacement[k]) and np.all(pd["u"][:, 2] == 0)
        assert np.array_equal(cells, orc.conn) and np.array_equal(pts[:, :2], orc.coors)
        assert set(cd) == {"cauchy_strain", "cauchy_stress", "mat_id"}
    # el_avg strain of a P1 cell = B u_e; checked against a direct numpy evaluation
    co, cn, uf = orc.coors, orc.conn, an.displacement[-1]
    x, y = co[cn, 0], co[cn, 1]
    det = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    gx = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], 1) / det[:, None]
    gy = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], 1) / det[:, None]
    ux, uy = uf[cn, 0], uf[cn, 1]
    strain = np.stack([(gx * ux).sum(1), (gy * uy).sum(1), (gy * ux + gx * uy).sum(1)], 1)
    assert np.abs(an.cell_strain - strain).max() <= 1e-12 * np.abs(strain).max()
    D = an.setup.sample.D[0]
    assert np.abs(an.cell_stress - strain @ D.T).max() <= 1e-12 * np.abs(strain @ D.T).max()
    assert os.path.exists(os.path.join(cond_dir, "regions.vtk"))
    _, _, rp, _ = read_vtk(os.path.join(cond_dir, "regions.vtk"))
    assert set(rp) == {"Omega", "VertexForce0", "EdgeForce0", "EdgeConstraint0"}
    assert int(rp["EdgeConstraint0"].sum()) == 21 and int(rp["VertexForce0"].sum()) == 1
    assert open(os.path.join(cond_dir, "magnitudes.txt")).read() == \
        "VertexForce0:(0, -1000)\nEdgeForce0:(%s, 0.0)\n" % repr(120 / 21)
    # images: two-pass window sizing as generate.py:129-145 does it
    out = os.path.join(data_dir, "outline.png")
    an.save_input_image(out, outline=True, crop=False)
    from fea_diffusion_b200.datagen.utils import find_image_bounds
    l, t, r, b = find_image_bounds(out)
    W2 = round(64 / (max(r - l, b - t) / an.initial_image_size))
    an.update_image_size_or_bounds(image_size=W2)
    an.save_input_image(out, outline=True, crop=False)
    l, t, r, b = find_image_bounds(out)
    lo, hi = (l, r) if r > b else (t, b)
    an.update_image_size_or_bounds(bounds=(lo, lo, hi, hi))
    assert (W2, (lo, lo, hi, hi)) == imaging.plate_window(an.setup.bbox(), 64)
    an.save_input_image(os.path.join(data_dir, "input.png"))
    an.save_region_images(os.path.join(cond_dir, "regions"))
    an.save_output_images(os.path.join(cond_dir, "outputs"), save_displacement=True, save_stress=True, save_strain=False)
    size = hi - lo
    inp = np.array(Image.open(os.path.join(data_dir, "input.png")))
    assert inp.shape == (size, size, 3) and (inp == 0).any() and (inp == 255).any()
    lines = open(os.path.join(cond_dir, "ranges.txt")).read().splitlines()
    assert len(lines) == 4 * 4 and lines[0].startswith("displacement_x_1:(") and lines[3].startswith("stress_y_1:(")
    lo_hi = eval(lines[4 + 1].split(":", 1)[1])          # displacement_y_2
    assert abs(lo_hi[0] - u[2][:, 1].min()) <= 1e-8 * abs(u[2][:, 1].min())
    aff = imaging.crop_affine(an.setup.bbox(), W2, (lo, lo, hi, hi))
    for c, name in enumerate(["displacement_x", "displacement_y"]):
        img = np.array(Image.open(os.path.join(cond_dir, "outputs_%s.png" % name)))
        assert img.shape == (size, size, 3) and np.array_equal(img[:, :, 0], img[:, :, 2])
        ref = ro.rasterize_scalar(orc.coors, orc.conn, u[1][:, c], size, aff)
        assert np.abs(ref.astype(int) - img[:, :, 0].astype(int)).max() <= 1
    assert Image.open(os.path.join(cond_dir, "regions_EdgeConstraint0.png")).size == (size, size)
    for name in ("stress_x", "stress_y"):
        assert os.path.exists(os.path.join(cond_dir, "outputs_%s.png" % name))
    # the 0/1 region flag decays over one 0.01-wide cell: sub-pixel at 64 px, a dark band at 512 px
    an.update_image_size_or_bounds(image_size=512, bounds=(0, 0, 512, 512))
    an.save_region_images(os.path.join(cond_dir, "big"))
    reg = np.array(Image.open(os.path.join(cond_dir, "big_EdgeConstraint0.png")))[:, :, 0]
    assert reg.shape == (512, 512) and reg[:, :128].min() < 64 and reg[:, 128:].min() == 255
    frc = np.array(Image.open(os.path.join(cond_dir, "big_EdgeForce0.png")))[:, :, 0]
    assert frc[:, 384:].min() < 64 and frc[:, :384].min() == 255          # edge (3, 4) is the right edge
    an.clear_condition_dir()
    assert os.listdir(cond_dir) == []


def test_feanalysis_text_files_match_the_committed_composite_condition(tmp_path, golden_meta):
    """applications/composite/{magnitudes,materials}.txt (reference) for the notebook's condition;
    that condition floats (F4), which strict mode reports instead of writing solver noise."""
    co, cn, kw = cases.composite_args(well_posed=False)
    kw = dict(force_edges_tags_magnitudes=[], constraints_edges_tags=[], **kw)   # required positionals
    data_dir, cond_dir = str(tmp_path), str(tmp_path / "c")
    os.makedirs(cond_dir)
    write_medit(os.path.join(data_dir, "part.mesh"), co, cn)
    an = FEAnalysis("part.mesh", data_dir, cond_dir, num_steps=2, max_iter=2000, **kw)
    assert open(os.path.join(cond_dir, "magnitudes.txt")).read() == golden_meta["composite_magnitudes"]
    assert open(os.path.join(cond_dir, "materials.txt")).read() == golden_meta["composite_materials"]
    assert an.calculate() is False
    lenient = FEAnalysis("part.mesh", data_dir, cond_dir, num_steps=2, max_iter=2000, strict=False, **kw)
    assert lenient.calculate() is True and np.isfinite(lenient.displacement).all()


def test_generate_data_tree(tmp_path):
    """generate_data (the reference's keyword surface, datagen/generate.py:12-31) end to end: directory
    tree and file contents the training set reader expects (model/diffusion.py:134-141, 174-217,
    359-378), incl. the stress / strain images and their ranges.txt lines (fea_analysis.py:539-558)."""
    d = str(tmp_path / "data")
    generate_data(data_dir=d, image_size=64, num_plates=2, conditions_per_plate=2, mesh_size=4e-2,
                  num_steps_per_condition=5, save_meshes=True, save_stress=True, save_strain=True, seed=11, workers=2)
    assert sorted(os.listdir(d)) == ["1", "2"]
    kinds = ["displacement_x", "displacement_y", "stress_x", "stress_y", "strain_x", "strain_y"]
    for plate in ("1", "2"):
        pd = os.path.join(d, plate)
        assert {"1", "2", "input.png", "outline.png"} <= set(os.listdir(pd))
        size = Image.open(os.path.join(pd, "input.png")).size
        assert abs(size[0] - 64) <= 2 and size[0] == size[1]
        for cond in ("1", "2"):
            cd = os.path.join(pd, cond)
            files = set(os.listdir(cd))
            assert {"outputs_%s.png" % k for k in kinds} | {"magnitudes.txt", "materials.txt", "ranges.txt", "regions.vtk"} <= files
            assert "status.txt" not in files                      # only converged solutions are written
            assert {"domain.%d.vtk" % k for k in range(5)} <= files
            assert any(f.startswith("regions_MaterialRegion") for f in files)
            for k in kinds:
                assert Image.open(os.path.join(cd, "outputs_%s.png" % k)).size == size
            lines = open(os.path.join(cd, "ranges.txt")).read().splitlines()
            assert [l.split(":")[0] for l in lines] == ["%s_%d" % (k, st) for st in range(1, 5) for k in kinds]
            _, _, p4, c4 = read_vtk(os.path.join(cd, "domain.4.vtk"))
            _, _, p1, _ = read_vtk(os.path.join(cd, "domain.1.vtk"))
            assert np.allclose(p1["u"], 0.25 * p4["u"], rtol=1e-14, atol=0)
            final = dict(l.split(":", 1) for l in lines[-6:])
            lo, hi = eval(final["displacement_x_4"])
            assert lo == p4["u"][:, 0].min() and hi == p4["u"][:, 0].max()
            lo, hi = eval(final["stress_y_4"])
            assert lo == c4["cauchy_stress"][:, 1].min() and hi == c4["cauchy_stress"][:, 1].max()
            lo, hi = eval(final["strain_x_4"])
            assert lo == c4["cauchy_strain"][:, 0].min() and hi == c4["cauchy_strain"][:, 0].max()
            img = np.array(Image.open(os.path.join(cd, "outputs_stress_x.png")))[:, :, 0]
            assert img.min() < 128 and (img == 255).any()         # own range: dark at the extreme cells, white background


def test_batched_generator_matches_the_dropin_loop_sample_by_sample(tmp_path):
    """fea_diffusion_b200.dataset.generate_dataset (batched, pipelined) against FEAnalysis on the
    same plates: same files, same pixels, same ranges; shards of a 2-rank run are identical to
    the 1-rank run."""
    from fea_diffusion_b200.dataset import generate_dataset
    from fea_diffusion_b200.workload import plate_conditions
    kw = dict(conditions_per_plate=2, image_size=64, num_steps=5, mesh_size=5e-2, seed=7, plates_per_batch=2,
              workers=2, save_meshes=True, save_stress=True, save_strain=True, region_method="lloyd")   # as plate_conditions
    d1 = str(tmp_path / "one")
    st = generate_dataset(d1, 3, **kw)
    assert st["plates"] == 3 and st["samples"] == 6
    d2 = str(tmp_path / "two")
    for rank in range(2):
        generate_dataset(d2, 3, rank=rank, world=2, device=0, **kw)
    for plate in ("1", "2", "3"):
        for root, _, files in os.walk(os.path.join(d1, plate)):
            for f in files:
                a = open(os.path.join(root, f), "rb").read()
                b = open(os.path.join(root.replace(d1, d2), f), "rb").read()
                assert a == b, (root, f)
    # plate 2 through the drop-in class
    items, _ = plate_conditions(7 + 1, 2, 64, mesh_size=5e-2)
    data_dir = str(tmp_path / "ref")
    os.makedirs(data_dir)
    write_medit(os.path.join(data_dir, "part.mesh"), items[0].setup.coors, items[0].setup.conn)
    for ci, it in enumerate(items):
        cond_dir = os.path.join(data_dir, str(ci + 1))
        os.makedirs(cond_dir)
        an = FEAnalysis("part.mesh", data_dir, cond_dir, num_steps=5, **it.kwargs)
        assert an.calculate()
        an.update_image_size_or_bounds(image_size=it.window, bounds=it.bounds)
        an.save_region_images(os.path.join(cond_dir, "regions"))
        an.save_output_images(os.path.join(cond_dir, "outputs"), save_stress=True, save_strain=True)
        got = os.path.join(d1, "2", str(ci + 1))
        for f in sorted(os.listdir(cond_dir)):
            if f.endswith(".png"):
                x = np.array(Image.open(os.path.join(cond_dir, f))).astype(int)
                y = np.array(Image.open(os.path.join(got, f))).astype(int)
                assert x.shape == y.shape and np.abs(x - y).max() <= 1, f
            elif f in ("magnitudes.txt", "materials.txt"):
                assert open(os.path.join(cond_dir, f)).read() == open(os.path.join(got, f)).read()
        ra = [eval(l.split(":", 1)[1]) for l in open(os.path.join(cond_dir, "ranges.txt"))]
        rb = [eval(l.split(":", 1)[1]) for l in open(os.path.join(got, "ranges.txt"))]
        assert np.allclose(ra, rb, rtol=1e-12, atol=0)
        _, _, pd, cd = read_vtk(os.path.join(got, "domain.4.vtk"))
        assert np.allclose(pd["u"][:, :2], an.displacement[-1], rtol=1e-9, atol=1e-18)
        assert np.allclose(cd["cauchy_strain"], an.cell_strain, rtol=1e-9, atol=1e-18)
    assert np.array_equal(np.array(Image.open(os.path.join(d1, "2", "input.png")).size), [items[0].size] * 2)


def test_region_and_input_images_match_the_committed_composite_dataset(tmp_path, golden):
    """FEAnalysis.save_input_image / save_region_images on the GPU against the reference's own
    dataset images (applications/composite/*.png): same size, >= 99.4 % pixels bit-equal."""
    co, cn, kw = cases.composite_args(well_posed=False)
    kw = dict(force_edges_tags_magnitudes=[], constraints_edges_tags=[], **kw)
    data_dir, cond_dir = str(tmp_path), str(tmp_path / "c")
    os.makedirs(cond_dir)
    write_medit(os.path.join(data_dir, "part.mesh"), co, cn)
    an = FEAnalysis("part.mesh", data_dir, cond_dir, num_steps=2, strict=False, max_iter=50, **kw)
    W, b = imaging.plate_window(an.setup.bbox(), 512)
    an.update_image_size_or_bounds(image_size=W, bounds=b)
    an.save_input_image(os.path.join(data_dir, "input.png"))
    an.save_region_images(os.path.join(cond_dir, "regions"))
    got = {"input": os.path.join(data_dir, "input.png")}
    for nm in ("MaterialRegion0", "MaterialRegion1", "VertexForce0", "VertexConstraint0"):
        got["regions_" + nm] = os.path.join(cond_dir, "regions_%s.png" % nm)
    for name, path_ in got.items():
        mine = np.array(Image.open(path_))[:, :, 0].astype(int)
        ref = golden["composite_png_" + name].astype(int)
        assert mine.shape == ref.shape == (512, 512)
        assert (mine == ref).mean() >= 0.994 and ((mine < 128) == (ref < 128)).mean() >= 0.996, name
