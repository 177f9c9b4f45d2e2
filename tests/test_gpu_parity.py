"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the reference's golden
vectors.  Integer/index work is bit-exact; floating point within the tolerance north_star states
(rel-L2 displacement <= 1e-8, ranges within 1e-8 relative, images within 1 LSB)."""
import numpy as np
import pytest

import cases
from fea_diffusion_b200 import Context, FeaError, Sample, pack
from fea_diffusion_b200 import imaging
from fea_diffusion_b200._capi import (SAMPLE_BREAKDOWN, SAMPLE_CONVERGED, SAMPLE_EMPTY_ROW,
                                      SAMPLE_MAX_ITER, SAMPLE_STAGNATED)
from oracle import raster_oracle as ro
from oracle.fea_oracle import element_stiffness

pytestmark = pytest.mark.gpu

RTOL_SOLVER = 1e-11          # PCG stopping tolerance used by the tests
TOL_U = 1e-8                 # north_star: rel-L2 displacement error vs the direct solve


@pytest.fixture(scope="module", params=["onchip", "stream"])
def ctx(request):
    """Both solver paths: systems that fit a thread-block cluster stay on chip (k_pcg_cluster);
    "stream" forces the lock-step streaming kernels (what ~1M-DOF systems always use)."""
    c = Context(0)
    c.set_option("pcg_path", 1 if request.param == "stream" else 0)
    c.path = request.param
    yield c
    c.close()


CASES = {"cantilever": cases.cantilever, "shearblade": cases.shearblade, "gusset": cases.gusset,
         "composite": cases.composite, "quads": cases.quad_plate}


@pytest.fixture(scope="module")
def built(ctx):
    """Each case assembled on the GPU once (single-sample batches)."""
    out = {}
    for name, fn in CASES.items():
        setup, orc = fn()
        b = ctx.create_batch(pack([setup.sample])).assemble()
        out[name] = (setup, orc, b)
    yield out
    for _, _, b in out.values():
        b.destroy()


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("name", list(CASES))
def test_orientation_fix_bit_exact(built, name):
    setup, orc, b = built[name]
    conn, flips = b.conn()
    assert np.array_equal(conn, orc.conn)
    assert int(flips[0]) == orc.n_flipped
    if name == "shearblade":
        assert orc.n_flipped == len(orc.conn)


@pytest.mark.parametrize("name", list(CASES))
def test_element_stiffness(built, name):
    setup, orc, b = built[name]
    ke = b.element_stiffness()
    Dc = setup.sample.D[np.maximum(setup.sample.cell_region, 0)]
    ref = element_stiffness(orc.coors, orc.conn, Dc)
    ref[setup.sample.cell_region < 0] = 0.0
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    scale[scale == 0] = 1.0
    assert np.abs((ke - ref) / scale).max() <= 1e-13


@pytest.mark.parametrize("name,n,nnz", [("cantilever", 4844, 65896), ("shearblade", 10466, 144084),
                                        ("gusset", 19996, 276864), ("composite", None, None),
                                        ("quads", None, None)])
def test_pattern_bit_exact_and_values(built, name, n, nnz):
    """CSR pattern identical to sfepy's (A-11: ascending active-DOF numbering, sorted unique
    columns, full node blocks, explicit zeros) -- pinned by the log nnz figures -- and values
    equal to the oracle assembly."""
    setup, orc, b = built[name]
    A = orc.stiffness()
    G = b.csr(0)
    if n is not None:
        assert G.shape == (n, n) and G.nnz == nnz
    assert G.shape == A.shape
    assert np.array_equal(G.indptr, A.indptr)
    assert np.array_equal(G.indices, A.indices)
    assert np.abs(G.data - A.data).max() <= 1e-12 * np.abs(A.data).max()
    a, z = b.sample_sizes()
    assert (a[0], z[0]) == (A.shape[0], A.nnz)
    info = b.info()
    assert info["nnz"] == A.nnz and info["n_active_dofs"] == A.shape[0]


@pytest.mark.parametrize("name", list(CASES))
def test_spmv_kernel(built, name):
    setup, orc, b = built[name]
    A = orc.stiffness()
    x = np.random.default_rng(1).standard_normal(A.shape[0])
    y = b.spmv(0, x)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("name", ["cantilever", "shearblade"])
def test_solve_matches_golden(built, name):
    """sfepy's own fp64 output (applications/*/*.vtk in the reference)."""
    setup, orc, b = built[name]
    r = b.solve(RTOL_SOLVER, 20000).download()
    g = cases.fixtures()[name + "_u"][:, :2]
    assert r.status[0] == SAMPLE_CONVERGED
    assert rel(r.u, g) <= TOL_U
    assert np.all(r.u[setup.sample.fixed.astype(bool)] == 0.0)  # A-15


@pytest.mark.parametrize("name", list(CASES))
def test_cell_strain_stress_matches_oracle(built, name):
    """The post-process hook's cauchy_strain / cauchy_stress cell fields (fea_analysis.py:397-416)
    against oracle.cell_strain_stress: with the oracle's own displacement the two agree to rounding
    (1e-12 of the field's largest entry); with the GPU solve to the solver tolerance.  Covers the
    two-material composite plate (the hook applies the FIRST material's D to every cell) and Q1 cells.
    UNPINNED in the reference (no artefact holds these fields), stated in the oracle."""
    setup, orc, b = built[name]
    r = b.solve(RTOL_SOLVER, 40000).download()
    region = 0 if len(setup.sample.D) else -1
    strain, stress = b.cell_strain_stress(region)
    e_ref, s_ref = orc.strain_stress(r.u)              # same u: isolates the hook's arithmetic
    assert np.abs(strain - e_ref).max() <= 1e-12 * np.abs(e_ref).max()
    assert np.abs(stress - s_ref).max() <= 1e-12 * np.abs(s_ref).max()
    e_or, s_or = orc.strain_stress(orc.solve("best")[-1])   # end to end against the direct solve
    assert rel(strain, e_or) <= 1e-7 and rel(stress, s_or) <= 1e-7
    if name == "composite":   # per-cell D (stress_region -1) differs from the hook's single-material rule
        _, s_cell = b.cell_strain_stress(-1)
        Dc = orc.D[np.maximum(orc.cell_region, 0)]
        ref = np.einsum("eij,ej->ei", Dc, e_ref)
        ref[orc.cell_region < 0] = 0.0
        assert np.abs(s_cell - ref).max() <= 1e-12 * np.abs(ref).max()
        assert np.abs(s_cell - stress).max() > 1e-3 * np.abs(stress).max()


@pytest.mark.parametrize("name", list(CASES))
def test_solve_matches_oracle_and_ranges(built, name):
    setup, orc, b = built[name]
    assert orc.classify()["well_posed"] == 1
    r = b.solve(RTOL_SOLVER, 40000).download()
    u = orc.solve("reference")
    assert r.status[0] == SAMPLE_CONVERGED
    assert rel(r.u, u[-1]) <= TOL_U
    # load steps are t_k multiples of the final state (F5): ranges.txt values within 1e-8
    times = np.linspace(0, 1, orc.num_steps)
    for k in range(1, orc.num_steps):
        for c in range(2):
            lo, hi = times[k] * r.ranges[0, 2 * c], times[k] * r.ranges[0, 2 * c + 1]
            ref_lo, ref_hi = u[k][:, c].min(), u[k][:, c].max()
            span = max(abs(ref_lo), abs(ref_hi))
            assert abs(lo - ref_lo) <= 1e-8 * span and abs(hi - ref_hi) <= 1e-8 * span
    assert np.array_equal(r.ranges[0], [r.u[:, 0].min(), r.u[:, 0].max(), r.u[:, 1].min(), r.u[:, 1].max()])


def test_lockstep_batch_is_bitwise_independent_of_batching(ctx, built):
    """Samples solved together give exactly the bytes they give alone (deterministic
    reductions; required for 1-vs-N GPU sharding to be byte-identical)."""
    names = ["cantilever", "composite", "shearblade", "gusset", "cantilever"]
    singles = {n: built[n][2].solve(RTOL_SOLVER, 40000).download() for n in set(names)}
    with ctx.create_batch(pack([built[n][0].sample for n in names])) as b:
        r = b.assemble().solve(RTOL_SOLVER, 40000).download()
        us = b.packed.split_vertices(r.u)
    for i, n in enumerate(names):
        assert np.array_equal(us[i], singles[n].u), n
        assert r.iters[i] == singles[n].iters[0]
        assert np.array_equal(r.ranges[i], singles[n].ranges[0])
    assert (r.status == SAMPLE_CONVERGED).all()
    assert len(set(r.iters.tolist())) > 1  # systems really finish at different iterations


@pytest.mark.parametrize("name,image_size", [("cantilever", 64), ("shearblade", 64), ("composite", 64),
                                             ("shearblade", 512), ("quads", 96)])
def test_raster(built, name, image_size):
    setup, orc, b = built[name]
    bbox = setup.bbox()
    W, bounds = imaging.plate_window(bbox, image_size)
    assert (W, bounds) == ro.closed_form_window(bbox, image_size)
    size = bounds[2] - bounds[0]
    assert abs(size - image_size) <= 2
    aff = imaging.crop_affine(bbox, W, bounds)
    assert np.array_equal(aff, np.array(ro.pixel_affine(bbox, W, bounds)))
    t1 = np.linspace(0, 1, orc.num_steps)[1]
    b.solve(RTOL_SOLVER, 40000).rasterize(size, aff[None], t1)
    r = b.download(images=True)
    u_or = orc.solve("best")
    for c in range(2):
        # same input -> bit-exact pixels (coverage rule and LUT arithmetic identical)
        same = ro.rasterize_scalar(orc.coors, orc.conn, t1 * r.u[:, c], size, aff)
        assert np.array_equal(r.images[0, c], same)
        # oracle end to end (direct solve): within 1 LSB
        ref = ro.rasterize_scalar(orc.coors, orc.conn, u_or[1][:, c], size, aff)
        d = np.abs(ref.astype(int) - r.images[0, c].astype(int))
        assert d.max() <= 1
        assert (r.images[0, c] != 255).sum() > 0.05 * size * size * min(1.0, (bbox[3] - bbox[1]) / (bbox[2] - bbox[0]))


def test_both_solver_paths_agree(built, ctx):
    """On-chip and streaming PCG differ only in summation order: same iterates to rounding."""
    other = Context(0)
    other.set_option("pcg_path", 0 if ctx.path == "stream" else 1)
    try:
        for name in ("cantilever", "gusset"):
            setup, orc, b = built[name]
            r1 = b.solve(RTOL_SOLVER, 40000).download()
            with other.create_batch(pack([setup.sample])) as b2:
                r2 = b2.assemble().solve(RTOL_SOLVER, 40000).download()
                st2 = b2.stats()
            st1 = b.stats()
            assert rel(r1.u, r2.u) <= 1e-9 and abs(int(r1.iters[0]) - int(r2.iters[0])) <= 3
            assert (st1["cluster_systems"], st2["cluster_systems"]) == ((1, 0) if ctx.path == "onchip" else (0, 1))
    finally:
        other.close()


def test_singular_samples_are_reported(ctx):
    """(i) the reference's committed composite condition floats (F4): CG must not claim
    convergence; (ii) an active vertex with no stiffness cell is the reference's NaN path."""
    setup, orc = cases.composite(well_posed=False)
    assert orc.classify()["well_posed"] == 0
    s2, _ = cases.cantilever()
    smp = s2.sample
    cr = smp.cell_region.copy()
    touching = (smp.conn == 100).any(axis=1)   # strip every cell around vertex 100
    cr[touching] = -1
    lonely = Sample(smp.coors, smp.conn, cr, smp.D, smp.fixed, smp.rhs)
    good, _ = cases.cantilever()
    with ctx.create_batch(pack([setup.sample, lonely, good.sample])) as b:
        r = b.assemble().solve(1e-10, 3000).download()
        us = b.packed.split_vertices(r.u)
    assert r.status[0] in (SAMPLE_BREAKDOWN, SAMPLE_MAX_ITER, SAMPLE_STAGNATED)
    assert r.status[1] == SAMPLE_EMPTY_ROW and np.isnan(us[1]).all()
    assert r.status[2] == SAMPLE_CONVERGED and np.isfinite(us[2]).all()


def test_edge_cases(ctx):
    setup, _ = cases.cantilever()
    smp = setup.sample
    zero = Sample(smp.coors, smp.conn, smp.cell_region, smp.D, smp.fixed, np.zeros_like(smp.rhs))
    allfixed = Sample(smp.coors, smp.conn, smp.cell_region, smp.D, np.ones(len(smp.coors), bool), smp.rhs)
    tiny = Sample(np.array([[0, 0], [1, 0], [0, 1.0]]), np.array([[0, 2, 1]], np.int32), np.zeros(1, np.int8),
                  smp.D, np.array([1, 1, 0], bool), np.array([[0, 0], [0, 0], [0, -1.0]]))
    with ctx.create_batch(pack([zero, allfixed, tiny])) as b:
        r = b.assemble().solve(1e-12, 100).download()
        us = b.packed.split_vertices(r.u)
        a, z = b.sample_sizes()
    assert list(a) == [2 * (len(smp.coors) - int(smp.fixed.sum())), 0, 2] and z[1] == 0 and z[2] == 4
    assert (r.status == SAMPLE_CONVERGED).all()
    assert r.iters[0] == 0 and np.all(us[0] == 0)
    assert np.all(us[1] == 0)
    K = element_stiffness(tiny.coors, np.array([[0, 1, 2]], np.int32), smp.D[0])[0][4:, 4:]
    assert np.allclose(us[2][2], np.linalg.solve(K, [0, -1.0]), rtol=1e-10)


def test_bad_arguments_fail_loudly(ctx):
    setup, _ = cases.cantilever()
    smp = setup.sample
    bad = Sample(smp.coors, smp.conn + 5, smp.cell_region, smp.D, smp.fixed, smp.rhs)  # index overflow
    with pytest.raises(FeaError):
        ctx.create_batch(pack([bad])).assemble()
    with ctx.create_batch(pack([smp])) as b:
        with pytest.raises(FeaError):
            b.solve()      # before assemble
        b.assemble()
        with pytest.raises(FeaError):
            b.download()   # before solve


def test_one_call_host_path(ctx, built):
    """fea_solve_batch (the e2e entry) equals the staged path byte for byte."""
    setup, orc, b = built["cantilever"]
    staged = b.solve(RTOL_SOLVER, 20000).download()
    bbox = setup.bbox()
    W, bounds = imaging.plate_window(bbox, 64)
    size = bounds[2] - bounds[0]
    aff = imaging.crop_affine(bbox, W, bounds)[None]
    r = ctx.solve_batch(pack([setup.sample]), RTOL_SOLVER, 20000, image_size=size, affine=aff, value_scale=0.1)
    assert np.array_equal(r.u, staged.u) and np.array_equal(r.ranges, staged.ranges)
    assert r.images.shape == (1, 2, size, size) and (r.images != 255).any()
    assert r.stats["kernel_launches"] > 0
    assert (r.stats["cluster_ms"] > 0 and r.stats["cluster_systems"] == 1) if ctx.path == "onchip" else r.stats["spmv_ms_avg"] > 0


def hinged_plates():
    """Two triangulated squares that touch in ONE vertex; the left one is clamped, the right one
    only hangs on the hinge and carries a load: a mechanism (rotation about the hinge), K singular
    and the load not in its range.  Vertex-connected, so a component count by vertices misses it."""
    n = 12
    xs, ys = np.meshgrid(np.linspace(0, 1, n + 1), np.linspace(0, 1, n + 1))
    sq = np.stack([xs.ravel(), ys.ravel()], 1)
    idx = lambda i, j: j * (n + 1) + i
    tri = np.array([[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)] for j in range(n) for i in range(n)] +
                   [[idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)] for j in range(n) for i in range(n)], dtype=np.int32)
    right = sq + [1.0, 1.0]                      # shares only the corner (1, 1)
    co = np.concatenate([sq, right[1:]])         # right[0] == sq[idx(n, n)]
    remap = np.concatenate([[idx(n, n)], len(sq) + np.arange(len(right) - 1)])
    cn = np.concatenate([tri, remap[tri]]).astype(np.int32)
    fixed = np.zeros(len(co), bool)
    fixed[[idx(0, j) for j in range(n + 1)]] = True
    rhs = np.zeros((len(co), 2))
    rhs[len(co) - 1] = (0.0, -100.0)             # far corner of the hanging square
    D = np.array([[[1.35, 0.58, 0], [0.58, 1.35, 0], [0, 0, 0.38]]]) * 2e5
    return Sample(co, cn, np.zeros(len(cn), np.int8), D, fixed, rhs)


def test_mechanism_is_stopped_early_not_at_max_iter(ctx):
    from fea_diffusion_b200.host import floating_components
    smp = hinged_plates()
    assert floating_components(smp) == (1, 0)          # the hinged part has no constraint of its own
    good, _ = cases.cantilever()
    with ctx.create_batch(pack([smp, good.sample])) as b:
        r = b.assemble().solve(1e-10, 200000).download()
    assert r.status[1] == SAMPLE_CONVERGED
    assert r.status[0] in (SAMPLE_STAGNATED, SAMPLE_BREAKDOWN)
    assert r.iters[0] <= 8192                          # true-residual monitor, not the iteration cap


@pytest.mark.parametrize("nx,ny,cl", [(48, 30, 1), (70, 42, 2), (90, 55, 3), (105, 66, 4), (118, 76, 5),
                                      (130, 84, 6), (140, 92, 7), (150, 102, 8)])
def test_every_cluster_size_agrees_with_the_streaming_path(nx, ny, cl):
    """The on-chip solver picks 1..8 CTAs per cluster from the system size: one Q1 plate per
    class, solved on chip and by the streaming kernels (which the goldens pin)."""
    setup, _ = cases.quad_plate(nx, ny)
    rows = -(-((nx + 1) * (ny + 1)) // 128) * 128
    assert -(-rows // 2048) == cl
    out = {}
    for path in (0, 1):
        c = Context(0)
        c.set_option("pcg_path", path)
        try:
            with c.create_batch(pack([setup.sample])) as b:
                r = b.assemble().solve(1e-11, 200000).download()
                st = b.stats()
        finally:
            c.close()
        assert r.status[0] == SAMPLE_CONVERGED
        assert (st["cluster_systems"], st["cluster_size"] if path == 0 else cl) == (1 - path, cl)
        out[path] = r
    assert rel(out[0].u, out[1].u) <= 1e-9
    assert abs(int(out[0].iters[0]) - int(out[1].iters[0])) <= max(3, out[1].iters[0] // 200)


def test_cluster_hands_a_system_back_when_its_halo_does_not_fit():
    """On chip every CTA receives the rows it gathers from its peers ("halo") by pushes into its own
    shared memory; a system whose halo does not fit is handed back to the streaming kernels.  Forced
    here with the halo cap knob: the result must be the streaming path's, bit for bit, and a
    single-CTA system (no halo) of the same batch must still be solved on chip."""
    big, _ = cases.quad_plate(90, 55)      # 3-CTA cluster
    small, _ = cases.quad_plate(48, 30)    # 1 CTA: no halo
    res = {}
    for name, opts in (("capped", {"cluster_halo_cap": 0}), ("stream", {"pcg_path": 1}), ("onchip", {})):
        c = Context(0)
        for k, v in opts.items():
            c.set_option(k, v)
        try:
            with c.create_batch(pack([big.sample, small.sample])) as b:
                res[name] = b.assemble().solve(1e-11, 200000).download()
        finally:
            c.close()
        assert (res[name].status == SAMPLE_CONVERGED).all()
    ub = [pack([big.sample, small.sample]).split_vertices(res[k].u) for k in ("capped", "stream", "onchip")]
    assert np.array_equal(ub[0][0], ub[1][0]) and res["capped"].iters[0] == res["stream"].iters[0]
    assert np.array_equal(ub[0][1], ub[2][1]) and res["capped"].iters[1] == res["onchip"].iters[1]
    assert rel(ub[0][0], ub[2][0]) <= 1e-9


def test_cluster_with_a_halo_of_thousands_of_rows():
    """A randomly renumbered 3-CTA plate solved in INPUT order (row_order 0): nearly every row a CTA
    gathers belongs to a peer, so halo and send lists hold thousands of entries (several passes of
    the push loop).  Same displacements as the lattice numbering, and it still runs on chip."""
    setup, _ = cases.quad_plate(90, 55)
    smp = setup.sample
    nv = len(smp.coors)
    rng = np.random.default_rng(3)
    perm = rng.permutation(nv)                 # new index of old vertex i
    inv = np.empty(nv, np.int64)
    inv[perm] = np.arange(nv)
    shuffled = Sample(smp.coors[inv], perm[smp.conn].astype(np.int32), smp.cell_region, smp.D,
                      np.asarray(smp.fixed)[inv], np.asarray(smp.rhs).reshape(nv, 2)[inv].reshape(np.shape(smp.rhs)))
    c = Context(0)
    c.set_option("row_order", 0)
    try:
        with c.create_batch(pack([smp, shuffled])) as b:
            r = b.assemble().solve(1e-11, 200000).download()
            us = b.packed.split_vertices(r.u)
            st = b.stats()
    finally:
        c.close()
    assert (r.status == SAMPLE_CONVERGED).all() and st["cluster_systems"] == 2
    assert rel(us[1][perm], us[0]) <= 1e-9


@pytest.mark.parametrize("cmin", [2, 5, 8])
def test_cluster_min_runs_small_systems_on_oversized_clusters(cmin):
    """``cluster_min`` forces larger clusters than a system needs: CTAs that own no rows at all (a
    128-row system on 8 CTAs) still take part in every hand-over of the iteration.  Same result as
    the default cluster to rounding."""
    tiny, _ = cases.quad_plate(12, 8)          # 117 vertices: one 32-row slice per CTA at most
    mid, _ = cases.quad_plate(70, 42)          # 2-CTA system by default
    small, _ = cases.quad_plate(48, 30)        # 1 CTA by default
    smp = [tiny.sample, mid.sample, small.sample]
    res = {}
    for k in (1, cmin):
        c = Context(0)
        c.set_option("cluster_min", k)
        try:
            with c.create_batch(pack(smp)) as b:
                res[k] = b.assemble().solve(1e-11, 200000).download()
                st = b.stats()
        finally:
            c.close()
        assert (res[k].status == SAMPLE_CONVERGED).all() and st["cluster_systems"] == 3
    assert rel(res[cmin].u, res[1].u) <= 1e-9
    assert np.abs(res[cmin].iters.astype(int) - res[1].iters.astype(int)).max() <= 5
