"""Property tests on seeded random meshes (the reference has no tests; SURVEY section 4 asks for
these): the assembled operator is symmetric and annihilates rigid-body motions before boundary
conditions, the solve matches the CPU oracle's direct solve, a batch of different meshes matches
its members solved alone, and the displacement is linear in the load."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from fea_diffusion_b200 import Context, Sample, pack
from fea_diffusion_b200._capi import SAMPLE_CONVERGED
from fea_diffusion_b200.host import ProblemSetup
from oracle.fea_oracle import OracleProblem

pytestmark = pytest.mark.gpu


def random_case(seed):
    """Delaunay mesh of a jittered point cloud in a random rectangle; Dirichlet on one random side,
    1-3 material bands chosen through the reference's coordinate-list selector, random vertex loads;
    about half of the cells are handed over clockwise."""
    rng = np.random.default_rng(seed)
    w, h = rng.uniform(0.5, 1.5), rng.uniform(0.3, 1.0)
    nx, ny = int(rng.integers(8, 30)), int(rng.integers(6, 20))
    xs, ys = np.meshgrid(np.linspace(0, w, nx), np.linspace(0, h, ny))
    co = np.stack([xs.ravel(), ys.ravel()], 1)
    inner = (co[:, 0] > 0) & (co[:, 0] < w) & (co[:, 1] > 0) & (co[:, 1] < h)
    co[inner] += rng.uniform(-0.3, 0.3, (int(inner.sum()), 2)) * [w / nx, h / ny]
    tri = Delaunay(co).simplices.astype(np.int32)
    flip = rng.random(len(tri)) < 0.5
    tri[flip] = tri[flip][:, [0, 2, 1]]
    side = int(rng.integers(0, 4))
    on_side = [co[:, 0] == 0, co[:, 0] == w, co[:, 1] == 0, co[:, 1] == h][side]
    ends = np.flatnonzero(on_side)
    key = co[ends, 1] if side < 2 else co[ends, 0]
    a, b = ends[np.argmin(key)], ends[np.argmax(key)]
    n_mat = int(rng.integers(1, 4))
    cuts = np.sort(rng.uniform(0.2 * w, 0.8 * w, n_mat - 1))
    band = np.searchsorted(cuts, co[:, 0])
    table = [(210000, 0.3), (68900, 0.33), (30000, 0.2)]
    mats = {table[m]: co[band == m] for m in range(n_mat) if (band == m).any()}
    loads = [(int(v) + 1, (float(rng.integers(-900, 900)), float(rng.integers(-900, 900))))
             for v in rng.choice(len(co), int(rng.integers(1, 5)), replace=False)]
    kw = dict(force_vertex_tags_magnitudes=loads, constraints_edges_tags=[(int(a) + 1, int(b) + 1)],
              material_properties_to_vertices=mats)
    return co, tri, kw


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


SEEDS = list(range(10))


@pytest.fixture(scope="module")
def solved(ctx):
    out = {}
    for seed in SEEDS:
        co, tri, kw = random_case(seed)
        setup = ProblemSetup(co, tri, **kw)
        orc = OracleProblem(co, tri, num_steps=3, **kw)
        with ctx.create_batch(pack([setup.sample])) as b:
            r = b.assemble().solve(1e-11, 50000).download()
            K = b.csr(0)
        out[seed] = (setup, orc, r, K)
    return out


@pytest.mark.parametrize("seed", SEEDS)
def test_matches_oracle_direct_solve(solved, seed):
    setup, orc, r, K = solved[seed]
    info = orc.classify()
    A = orc.stiffness()
    assert np.array_equal(K.indptr, A.indptr) and np.array_equal(K.indices, A.indices)
    assert abs(K - K.T).max() <= 1e-12 * abs(K).max()
    if not info["well_posed"]:       # straddling cells can detach a band: nothing to compare
        assert r.status[0] != SAMPLE_CONVERGED or np.isfinite(r.u).all()
        return
    u = orc.solve("best")[-1]
    assert r.status[0] == SAMPLE_CONVERGED
    assert np.linalg.norm(r.u - u) / np.linalg.norm(u) <= 1e-8


def test_rigid_body_motions_are_in_the_null_space_before_bcs(ctx):
    co, tri, kw = random_case(3)
    setup = ProblemSetup(co, tri, material_properties_to_vertices=None)   # one material, nothing fixed
    with ctx.create_batch(pack([setup.sample])) as b:
        b.assemble()
        modes = [np.tile([1.0, 0.0], len(co)), np.tile([0.0, 1.0], len(co)),
                 np.stack([-co[:, 1], co[:, 0]], 1).reshape(-1)]
        scale = abs(b.csr(0)).max()
        for m in modes:
            y = b.spmv(0, m)
            assert np.abs(y).max() <= 1e-10 * scale * np.abs(m).max()


def test_batch_of_random_meshes_equals_its_members(ctx, solved):
    good = [s for s in SEEDS if solved[s][1].classify()["well_posed"]]
    with ctx.create_batch(pack([solved[s][0].sample for s in good])) as b:
        r = b.assemble().solve(1e-11, 50000).download()
        us = b.packed.split_vertices(r.u)
    for i, s in enumerate(good):
        assert np.array_equal(us[i], solved[s][2].u) and r.iters[i] == solved[s][2].iters[0]


def test_linearity_in_the_load(ctx, solved):
    seed = next(s for s in SEEDS if solved[s][1].classify()["well_posed"])
    smp = solved[seed][0].sample
    tripled = Sample(smp.coors, smp.conn, smp.cell_region, smp.D, smp.fixed, 3.0 * smp.rhs)
    with ctx.create_batch(pack([tripled])) as b:
        r3 = b.assemble().solve(1e-11, 50000).download()
    u = solved[seed][2].u
    assert np.linalg.norm(r3.u - 3.0 * u) / np.linalg.norm(3.0 * u) <= 1e-9


def test_result_does_not_depend_on_the_input_vertex_numbering(ctx):
    """The solver sorts every sample's rows spatially when the input numbering is not local
    (row_order auto): a randomly renumbered mesh gives the same displacements (to rounding) as the
    original, and the exported CSR pattern still follows the numbering it was given (A-9/A-11)."""
    from fea_diffusion_b200.host import floating_components
    seed = next(s for s in range(5, 40) if floating_components(ProblemSetup(*random_case(s)[:2], **random_case(s)[2]).sample) == (0, 0))
    co, tri, kw = random_case(seed)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(co))
    inv = np.empty(len(co), np.int64)
    inv[perm] = np.arange(len(co))
    kw2 = dict(kw)
    kw2["force_vertex_tags_magnitudes"] = [(int(inv[t - 1]) + 1, m) for t, m in kw["force_vertex_tags_magnitudes"]]
    kw2["constraints_edges_tags"] = [(int(inv[a - 1]) + 1, int(inv[b - 1]) + 1) for a, b in kw["constraints_edges_tags"]]
    kw2["material_properties_to_vertices"] = kw["material_properties_to_vertices"]
    s1 = ProblemSetup(co, tri, **kw)
    s2 = ProblemSetup(co[perm], inv[tri].astype(np.int32), **kw2)
    assert np.array_equal(np.asarray(s2.sample.fixed), np.asarray(s1.sample.fixed)[perm])
    res = {}
    for order in (0, 1, 2, 3):
        ctx.set_option("row_order", order)
        with ctx.create_batch(pack([s1.sample, s2.sample])) as b:
            r = b.assemble().solve(1e-11, 50000).download()
            us = b.packed.split_vertices(r.u)
            K2 = b.csr(1)
        assert (r.status == SAMPLE_CONVERGED).all()
        assert np.linalg.norm(us[1] - us[0][perm]) / np.linalg.norm(us[0]) <= 1e-9
        res[order] = us[0]
        A2 = OracleProblem(co[perm], inv[tri].astype(np.int32), num_steps=2, **kw2).stiffness()
        assert np.array_equal(K2.indptr, A2.indptr) and np.array_equal(K2.indices, A2.indices)
    ctx.set_option("row_order", 3)
    for order in (1, 2, 3):
        assert np.linalg.norm(res[order] - res[0]) / np.linalg.norm(res[0]) <= 1e-9
