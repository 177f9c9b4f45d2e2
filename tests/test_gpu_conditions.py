"""Device-side problem set-up (fea_batch_create_from_conditions) against the host restatement
(host.ProblemSetup) and the CPU oracle: region vertex sets, Dirichlet mask, material cells, D, load
vector must be BIT-IDENTICAL (they decide set membership and the assembled system; reference
fea_analysis.py:76-146, 182-194, 235-252), the classifier counts equal (SURVEY A-19), and a batch
made from conditions must solve to the same bytes as the batch made from the derived arrays."""
import numpy as np
import pytest

import cases
from fea_diffusion_b200 import Context, PackedConditions, ProblemSetup, pack
from fea_diffusion_b200.host import MeshTopology, floating_components
from fea_diffusion_b200.workload import build_workload
from oracle.fea_oracle import OracleProblem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def check_against_host(ctx, meshes, samples, solve=False):
    """samples: [(mesh index, kwargs)].  Returns the device set-up and the host set-ups."""
    pc = PackedConditions(meshes, samples)
    setups = []
    topo = {}
    for m, kw in samples:
        co, cn = meshes[m]
        if m not in topo:
            topo[m] = MeshTopology(np.asarray(cn, np.int32), len(co))
        setups.append(ProblemSetup(co, cn, topology=topo[m], **kw))
    with ctx.create_batch_from_conditions(pc) as b:
        b.assemble()
        dev = b.setup()
        fl, em = b.classify()
        info = b.info()
        if solve:
            b.solve(1e-10, 50000)
            res = b.download()
    fixed = pc.split_vertices(dev.fixed)
    rhs = pc.split_vertices(dev.rhs)
    for s, st in enumerate(setups):
        smp = st.sample
        c0, c1 = pc.cell_off[s], pc.cell_off[s + 1]
        assert np.array_equal(fixed[s].astype(bool), np.asarray(smp.fixed, bool)), "fixed mask of sample %d" % s
        assert np.array_equal(dev.cell_region[c0:c1], smp.cell_region), "cell_region of sample %d" % s
        assert dev.rhs.dtype == np.float64 and np.array_equal(rhs[s].view(np.int64), np.asarray(smp.rhs).view(np.int64)), \
            "load vector of sample %d (bitwise)" % s
        assert dev.D[s].shape == smp.D.shape and np.array_equal(dev.D[s].view(np.int64), smp.D.view(np.int64)), \
            "material table of sample %d (bitwise)" % s
        names = pc.names[s]
        assert names == list(st.regions)
        for i, nm in enumerate(names):
            want = np.zeros(len(smp.coors), np.uint8)
            want[st.regions[nm]] = 1
            assert np.array_equal(dev.region_flags[s][i], want), (s, nm)
            assert dev.region_count[s][i] == len(st.regions[nm])
        assert (int(fl[s]), int(em[s])) == floating_components(smp), "classifier of sample %d" % s
    if solve:
        with ctx.create_batch(pack([st.sample for st in setups])) as b2:
            b2.assemble().solve(1e-10, 50000)
            res2 = b2.download()
            assert b2.info() == info
        assert np.array_equal(res.status, res2.status) and np.array_equal(res.iters, res2.iters)
        assert np.array_equal(res.u.view(np.int64), res2.u.view(np.int64))      # same inputs -> same bytes
    return dev, setups, (fl, em)


def test_composite_conditions_match_host_and_the_sfepy_log_pins(ctx):
    """applications/composite (reference datagenapplication.ipynb:262-316): np.isin selector, complete
    cells, F2.  The committed condition pins shape 19672 / nnz 270712 through the C-ABI."""
    co, cn, kw_bad = cases.composite_args(False)
    _, _, kw_ok = cases.composite_args(True)
    dev, setups, (fl, em) = check_against_host(ctx, [(co, cn)], [(0, kw_bad), (0, kw_ok)], solve=False)
    assert fl[0] > 0 and fl[1] == 0 and em[0] == 0 and em[1] == 0      # the committed condition floats (F4)
    pc = PackedConditions([(co, cn)], [(0, kw_bad)])
    with ctx.create_batch_from_conditions(pc) as b:
        b.assemble()
        n, nnz = b.sample_sizes()
        A = b.csr(0, values=False)
    assert int(n[0]) == 19672 and int(nnz[0]) == 270712 and A.shape == (19672, 19672) and A.nnz == 270712
    orc = OracleProblem(co, cn, **kw_bad)
    assert np.array_equal(A.indptr, orc.stiffness().indptr)


def test_overlap_combinations_and_tag_wraparound(ctx):
    """A-18: a cell complete in two regions gets the summed D as an extra table entry; tag 0 selects
    the last vertex like numpy's coors[-1]."""
    co = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [2, 0], [2, 1.0]])
    cn = np.array([[0, 1, 2], [0, 2, 3], [1, 4, 5], [1, 5, 2]], np.int32)
    mats = {(100.0, 0.3): [(0, 0), (1, 0), (1, 1), (0, 1)], (50.0, 0.2): [(1, 0), (2, 0), (2, 1), (1, 1), (0, 0)]}
    kw = dict(force_vertex_tags_magnitudes=[(6, (1, 2)), (0, (-3, 0.5))], constraints_vertex_tags=[1, 4],
              material_properties_to_vertices=mats)
    kw2 = dict(force_edges_tags_magnitudes=[((2, 3), (7, -9)), ((5, 6), (1, 1))], constraints_edges_tags=[(1, 4)],
               youngs_modulus=1234.5, poisson_ratio=0.31)
    dev, setups, _ = check_against_host(ctx, [(co, cn)], [(0, kw), (0, kw2)], solve=True)
    assert len(dev.D[0]) == 3 and (dev.cell_region[:4] >= 2).any()      # the overlap entry is in use


def test_quads_and_negative_zero(ctx):
    s, _ = cases.quad_plate()
    co, cn = s.coors.copy(), s.conn
    nx, ny = 24, 16
    idx = lambda i, j: j * (nx + 1) + i
    left = co[co[:, 0] <= 0.5]
    right = co[co[:, 0] > 0.5]
    left = left.copy()
    left[left[:, 0] == 0.0, 0] = -0.0                    # -0.0 must match the 0.0 abscissae (np.isin equality)
    assert np.isin(0.0, left[:, 0]) and np.signbit(left[:, 0]).any()
    kw = dict(force_vertex_tags_magnitudes=[(idx(nx, ny) + 1, (50, -400))],
              force_edges_tags_magnitudes=[((idx(nx, 0) + 1, idx(nx, ny) + 1), (30, 11))],
              constraints_edges_tags=[(idx(0, 0) + 1, idx(0, ny) + 1)],
              material_properties_to_vertices={(68900.0, 0.33): left, (117000.0, 0.34): right})
    check_against_host(ctx, [(co, cn)], [(0, kw)], solve=True)


def test_workload_batch_from_conditions(ctx):
    """48 plate-conditions of the bench workload (12 plates x 4): every derived array bit-identical
    to the host set-up, classifier equal, solve identical to the host-packed batch."""
    items, _ = build_workload(12, 4, 64, seed0=31)
    meshes, samples, index = [], [], {}
    for it in items:
        if it.plate not in index:
            index[it.plate] = len(meshes)
            meshes.append((it.setup.coors, it.setup.conn))
        samples.append((index[it.plate], it.kwargs))
    dev, setups, (fl, em) = check_against_host(ctx, meshes, samples, solve=True)
    assert not fl.any() and not em.any()                 # the workload keeps well-posed draws only


def test_as_sampled_conditions_classifier(ctx):
    """The reference's own condition distribution (a third singular by construction, F4): device
    classifier == host classifier == oracle classifier on every draw."""
    items, _ = build_workload(3, 8, 64, seed0=77, well_posed=False)
    meshes, samples, index = [], [], {}
    for it in items:
        if it.plate not in index:
            index[it.plate] = len(meshes)
            meshes.append((it.setup.coors, it.setup.conn))
        samples.append((index[it.plate], it.kwargs))
    dev, setups, (fl, em) = check_against_host(ctx, meshes, samples)
    assert (fl > 0).any()
    for s, it in enumerate(items[:8]):
        cl = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs).classify()
        assert (cl["floating_components"], cl["empty_rows"]) == (int(fl[s]), int(em[s]))


def test_region_images_from_device_flags(ctx):
    """fea_batch_rasterize_regions (device flags) == fea_batch_rasterize_flags (host flags)."""
    items, _ = build_workload(2, 2, 64, seed0=5)
    meshes, samples, index = [], [], {}
    for it in items:
        if it.plate not in index:
            index[it.plate] = len(meshes)
            meshes.append((it.setup.coors, it.setup.conn))
        samples.append((index[it.plate], it.kwargs))
    pc = PackedConditions(meshes, samples)
    size = max(it.size for it in items)
    affine = np.stack([it.affine for it in items])
    with ctx.create_batch_from_conditions(pc) as b:
        b.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1)
        mask = [i % 2 == 0 for i in range(len(items))]
        imgs = b.rasterize_regions(mask)
        flags = []
        for i, it in enumerate(items):
            f = np.zeros((len(it.setup.regions) + (1 if mask[i] else 0), len(it.setup.coors)), np.uint8)
            for k, nm in enumerate(it.setup.regions):
                f[k, it.setup.regions[nm]] = 1
            if mask[i]:
                f[-1] = 1
            flags.append(f)
        ref = b.rasterize_flags(flags)
    for a, r in zip(imgs, ref):
        assert a.shape == r.shape and np.array_equal(a, r)


def test_staged_outputs_fetched_on_a_copy_context(ctx):
    """fea_batch_stage_outputs / fea_batch_fetch_outputs: the read-back of batch j on a second context's stream, from
    another host thread, while the first context already solves batch j + 1 -- same bytes as the synchronous calls."""
    import threading
    items, _ = build_workload(3, 2, 64, seed0=9)
    meshes, samples, index = [], [], {}
    for it in items:
        if it.plate not in index:
            index[it.plate] = len(meshes)
            meshes.append((it.setup.coors, it.setup.conn))
        samples.append((index[it.plate], it.kwargs))
    pc = PackedConditions(meshes, samples)
    size = max(it.size for it in items)
    affine = np.stack([it.affine for it in items])
    mask = [it.condition == 0 for it in items]
    with ctx.create_batch_from_conditions(pc) as b:
        b.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1)
        want = b.download(images=True)
        want_reg = [r.copy() for r in b.rasterize_regions(mask)]
        want_cls = b.classify()
    copy_ctx = Context(0)
    got = {}
    b1 = ctx.create_batch_from_conditions(pc)
    b1.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1).stage_outputs(mask)
    th = threading.Thread(target=lambda: got.update(r=b1.fetch_outputs(copy_ctx)))
    th.start()
    with ctx.create_batch_from_conditions(pc) as b2:       # the next batch on the first context meanwhile
        b2.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1)
        again = b2.download(images=True)
    th.join()
    b1.destroy()
    # the big input arrays uploaded ahead of time on another context (fea_device_upload): the batch is created from
    # device pointers and must come out the same
    from fea_diffusion_b200.solver import PinnedArena
    arena = PinnedArena(ctx, pc.h2d_bytes + (1 << 16))
    pc_dev = PackedConditions(meshes, samples, alloc=arena.empty)
    arena.upload(copy_ctx)
    pc_dev.use_device_copy(arena)
    assert pc_dev.desc.xy != pc_dev.xy.ctypes.data and arena.spilled == 0
    with ctx.create_batch_from_conditions(pc_dev) as b4:
        b4.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1)
        r4 = b4.download(images=True)
    arena.release_device()
    copy_ctx.close()
    assert np.array_equal(r4.u, want.u, equal_nan=True) and np.array_equal(r4.images, want.images)
    res, reg, cls = got["r"]
    for r in (res, again):
        assert np.array_equal(r.u, want.u, equal_nan=True) and np.array_equal(r.status, want.status)
        assert np.array_equal(r.ranges, want.ranges, equal_nan=True) and np.array_equal(r.iters, want.iters)
        assert np.array_equal(r.images, want.images)
    assert np.array_equal(res.relres, want.relres, equal_nan=True)
    assert len(reg) == len(want_reg) and all(np.array_equal(a, w) for a, w in zip(reg, want_reg))
    assert np.array_equal(cls[0], want_cls[0]) and np.array_equal(cls[1], want_cls[1])
    # a batch made from arrays stages too (no region images, no classifier)
    with ctx.create_batch(pack([it.setup.sample for it in items])) as b3:
        b3.assemble().solve(1e-10, 50000).rasterize(size, affine, 0.1).stage_outputs()
        r3, reg3, cls3 = b3.fetch_outputs()
    assert reg3 is None and cls3 is None and np.array_equal(r3.u, want.u, equal_nan=True)


def test_classifier_on_samples_too_large_for_shared_memory(ctx):
    """The classifier keeps a plate-sized sample's union-find in shared memory (k_cc_sample); a sample with more
    than ~25 k cells goes through the global-memory kernels (k_cc_union ...).  Both against the host classifier, on
    a refined cantilever (77 k cells) held properly, not held at all, and with a second, unheld copy beside it."""
    from fea_diffusion_b200.workload import large_case
    setup, _ = large_case("cantilever", 2)
    s = setup.sample
    assert len(s.conn) > 60000
    import copy
    free = copy.deepcopy(s)
    free.fixed[:] = False
    two = copy.deepcopy(s)                       # the mesh twice in one sample, the copy shifted and not constrained
    nv = len(s.coors)
    two.coors = np.concatenate([s.coors, s.coors + np.array([10.0, 0.0])])
    two.conn = np.concatenate([s.conn, s.conn + nv]).astype(np.int32)
    two.cell_region = np.concatenate([s.cell_region, s.cell_region])
    two.fixed = np.concatenate([s.fixed, np.zeros(nv, dtype=s.fixed.dtype)])
    two.rhs = np.concatenate([s.rhs, np.zeros_like(s.rhs)])
    samples = [s, free, two]
    with ctx.create_batch(pack(samples)) as b:
        b.assemble()
        fl, em = b.classify()
    want = [floating_components(x) for x in samples]
    assert [(int(a), int(z)) for a, z in zip(fl, em)] == [tuple(int(v) for v in w) for w in want]
    assert [int(a) for a in fl] == [0, 1, 1] and not em.any()


def test_bad_tags_are_rejected(ctx):
    from fea_diffusion_b200 import FeaError
    co = np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]])
    cn = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    for kw in (dict(force_vertex_tags_magnitudes=[(9, (1, 1))]), dict(constraints_edges_tags=[(1, -7)])):
        with pytest.raises(FeaError):
            ctx.create_batch_from_conditions(PackedConditions([(co, cn)], [(0, kw)]))
