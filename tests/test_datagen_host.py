"""CPU-side tests of the drop-in ``datagen`` layer: file formats and host helpers (no GPU)."""
import hashlib

import numpy as np
import pytest
from PIL import Image

from fea_diffusion_b200 import host, imaging
from fea_diffusion_b200.datagen import utils
from fea_diffusion_b200.datagen.mesh_generator import MeshGenerator, write_medit
from fea_diffusion_b200.datagen.vtk_io import domain_filename, read_vtk, vtk_bytes, write_vtk
from oracle import mesh_io


@pytest.mark.parametrize("name", ["cantilever", "shearblade"])
def test_vtk_writer_is_byte_identical_to_the_reference_output(golden, golden_meta, name):
    """sha256 of applications/<name>/<name>.vtk (reference; written by sfepy via meshio 4.4.6)."""
    b = vtk_bytes(golden[name + "_vtk_points"][:, :2], golden[name + "_vtk_cells"],
                  {"u": golden[name + "_u"], "node_groups": golden[name + "_node_groups"]},
                  {"mat_id": golden[name + "_mat_id"]})
    assert hashlib.sha256(b).hexdigest()[:16] == golden_meta[name]["vtk_sha"]


def test_vtk_reader_matches_oracle_reader(tmp_path, golden):
    p = str(tmp_path / "d.vtk")
    u2 = golden["cantilever_u"][:, :2]
    write_vtk(p, golden["cantilever_coors"], golden["cantilever_conn"], {"u": u2},
              {"cauchy_strain": np.ones((len(golden["cantilever_conn"]), 3)), "mat_id": golden["cantilever_mat_id"]})
    pts, cells, pd, cd = read_vtk(p)
    o = mesh_io.read_vtk_legacy(p)
    assert np.array_equal(pts, o["points"]) and np.array_equal(cells, o["cells"])
    assert np.array_equal(pd["u"], o["point_data"]["u"]) and pd["u"].shape[1] == 3 and np.all(pd["u"][:, 2] == 0)
    assert np.array_equal(pd["u"][:, :2], u2)
    assert set(cd) == {"cauchy_strain", "mat_id"} and cd["mat_id"].dtype == np.int64


def test_domain_file_names():
    # reference fea_analysis.py:473-476, 586-589
    assert domain_filename(0, 11) == "domain.00.vtk" and domain_filename(10, 11) == "domain.10.vtk"
    assert domain_filename(3, 10) == "domain.3.vtk" and domain_filename(4, 5) == "domain.4.vtk"


def _bounds_reference_scan(path):
    """The reference's pixel loops (datagen/utils.py:18-56), kept literal for small images."""
    image = Image.open(path)
    px = image.load()
    W, H = image.size
    left, right, top, bottom = 0, W, 0, H
    for x in range(W):
        for y in range(H):
            if px[x, y] != (255, 255, 255):
                left = x
                break
        if left != 0:
            break
    for x in range(W - 1, -1, -1):
        for y in range(H):
            if px[x, y] != (255, 255, 255):
                right = x
                break
        if right != W:
            break
    for y in range(H):
        for x in range(W):
            if px[x, y] != (255, 255, 255):
                top = y
                break
        if top != 0:
            break
    for y in range(H - 1, -1, -1):
        for x in range(W):
            if px[x, y] != (255, 255, 255):
                bottom = y
                break
        if bottom != H:
            break
    return left, top, right, bottom


@pytest.mark.parametrize("box", [(5, 9, 40, 30), (0, 0, 47, 35), (0, 3, 10, 35), (12, 0, 47, 20), None])
def test_find_image_bounds_matches_reference_scan(tmp_path, box):
    img = np.full((36, 48, 3), 255, np.uint8)
    if box is not None:
        l, t, r, b = box
        img[t:b + 1, l:r + 1] = 100
        img[t + 2:b - 1, l + 2:r - 1] = 255
    p = str(tmp_path / "o.png")
    Image.fromarray(img).save(p)
    assert utils.find_image_bounds(p) == _bounds_reference_scan(p)


def test_find_image_bounds_on_the_committed_outline_renders(tmp_path, golden):
    # SURVEY A-16: cantilever l=13,r=526; shearblade l=39,r=551 in the reference's outline.png
    for name, (l, r) in {"cantilever": (13, 526), "shearblade": (39, 551)}.items():
        p = str(tmp_path / (name + ".png"))
        Image.fromarray(golden[name + "_png_outline"]).save(p)
        left, top, right, bottom = utils.find_image_bounds(p)
        assert (left, right) == (l, r)
        co = golden[name + "_coors"]
        bbox = (co[:, 0].min(), co[:, 1].min(), co[:, 0].max(), co[:, 1].max())
        W = golden[name + "_png_outline"].shape[0]
        assert imaging.outline_bounds(W, bbox) == (left, top, right, bottom)


def test_medit_writer_round_trips_bit_exactly(tmp_path):
    gen = MeshGenerator(random_seed=5)
    g = gen.normalize_geometry(gen.generate_geometry())
    ptags, ltags = gen.generate_mesh(g, str(tmp_path / "part"), mesh_size=4e-2)
    co, cn = host.read_mesh(str(tmp_path / "part.mesh"))
    assert np.array_equal(co, gen.mesh[0]) and np.array_equal(cn, gen.mesh[1])
    m = mesh_io.read_medit(str(tmp_path / "part.mesh"))
    assert np.array_equal(co, m["coors"]) and np.array_equal(cn, m["conn"])
    # geometry points come first: 1-based tag == vertex index + 1 (A-6)
    assert sorted(t for ring in ptags for t in ring) == list(range(1, 1 + sum(len(r) for r in ptags)))
    conds = gen.sample_conditions(ptags, ltags, 2)
    assert len(conds) == 2 and set(conds[0]) == {"material_regions", "point_constraints", "edge_constraints",
                                                 "point_forces", "edge_forces"}
    write_medit(str(tmp_path / "q.mesh"), np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]]), np.array([[0, 1, 2, 3]]))
    co, cn = host.read_mesh(str(tmp_path / "q.mesh"))
    assert cn.shape == (1, 4)


def test_reference_region_methods_partition_the_vertices_and_are_seeded():
    """Material regions of the condition sampler by the reference's own methods
    (mesh_generator.py:319-385, restated on scikit-learn in plates.py): every draw partitions the
    mesh vertices into 1..5 non-empty regions, both methods and all three linkages occur, the
    KMeans merge follows the reference's flattened-centre indexing, and a seed reproduces its stream."""
    pytest.importorskip("sklearn")
    from sklearn.cluster import KMeans
    from fea_diffusion_b200.plates import make_plate
    gen, ptags, ltags = make_plate(3, mesh_size=4e-2, region_method="reference")
    coors = gen.mesh[0]
    seen = set()
    orig_km, orig_ag = gen._regions_kmeans, gen._regions_agglomerative
    gen._regions_kmeans = lambda: (seen.add("kmeans"), orig_km())[1]
    gen._regions_agglomerative = lambda link: (seen.add(link), orig_ag(link))[1]
    conds = gen.sample_conditions(ptags, ltags, 24)
    assert seen == {"kmeans", "complete", "average", "ward"}
    for c in conds:
        regs = list(c["material_regions"].values())
        assert 1 <= len(regs) <= 5 and all(len(r) for r in regs)
        allp = np.concatenate(regs)
        assert len(allp) == len(coors)
        assert np.array_equal(np.unique(allp, axis=0), np.unique(coors, axis=0))
    # same seed, same stream
    gen2, p2, l2 = make_plate(3, mesh_size=4e-2, region_method="reference")
    c2 = gen2.sample_conditions(p2, l2, 3)
    for a, b in zip(conds[:3], c2):
        assert list(a["material_regions"]) == list(b["material_regions"])
        assert all(np.array_equal(x, y) for x, y in zip(a["material_regions"].values(), b["material_regions"].values()))
        # (magnitudes are drawn after ALL conditions of a call, mesh_generator.py:493-519: compare the tags)
        assert [t for t, _ in a["point_forces"]] == [t for t, _ in b["point_forces"]]
        assert a["edge_constraints"] == b["edge_constraints"]
    # the merge rule: cluster i takes the label of scalar i of the flattened (x0, y0, z0, x1, ...) centres
    gen3, _, _ = make_plate(3, mesh_size=4e-2, region_method="reference")
    state = gen3.random.getstate()
    regs = gen3._regions_kmeans()
    gen3.random.setstate(state)
    pts = np.concatenate([coors, np.zeros((len(coors), 1))], axis=1)
    nc = gen3.random.randint(5, 20)
    km = KMeans(n_clusters=nc, n_init=1, random_state=gen3.random.randrange(2 ** 31))
    lab = km.fit_predict(pts)
    nr = gen3.random.randint(1, 5)
    lab2 = KMeans(n_clusters=nr, n_init=1, random_state=gen3.random.randrange(2 ** 31)).fit_predict(km.cluster_centers_.reshape(-1, 1))
    for r in range(nr):
        want = np.concatenate([coors[lab == i] for i in range(nc) if lab2[i] == r] or [np.zeros((0, 2))])
        assert np.array_equal(regs[r], want)
    # the mesh does not depend on the region method
    lloyd, _, _ = make_plate(3, mesh_size=4e-2, region_method="lloyd")
    assert np.array_equal(lloyd.mesh[0], coors) and np.array_equal(lloyd.mesh[1], gen.mesh[1])
