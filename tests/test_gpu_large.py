"""BASELINE config 3: single large systems (uniformly refined cantilever, up to ~1.2 M DOFs) through
the streaming PCG kernels with two-level reductions.  The level-2 mesh is checked against the CPU
oracle's direct solve; level 4 is beyond what the oracle finishes in seconds, so it is checked through
size-independent properties: true residual of the exported CSR, symmetry, linearity in the load,
and convergence of the tip deflection under refinement."""
import numpy as np
import pytest

import cases
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200._capi import SAMPLE_CONVERGED
from fea_diffusion_b200.host import MeshTopology, ProblemSetup
from fea_diffusion_b200.workload import refine_uniform
from oracle.fea_oracle import OracleProblem

pytestmark = pytest.mark.gpu


def refined_cantilever(levels):
    """applications/cantilever/cantilever.py:43-52 (reference) on the `levels` times refined mesh
    (its `refine_mesh` hook, :26-27): fix x < 0.01, load (0, -1000) at vertex 3."""
    F = cases.fixtures()
    co, cn = refine_uniform(F["cantilever_coors"], F["cantilever_conn"], levels)
    setup = ProblemSetup(co, cn)
    fixed = MeshTopology(setup.conn, len(co)).facet_vertices(np.flatnonzero(co[:, 0] < 0.01))
    setup.sample.fixed[:] = False
    setup.sample.fixed[fixed] = True
    setup.sample.rhs[:] = 0
    setup.sample.rhs[3] = (0.0, -1000.0)
    return setup, fixed


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("level", [2, 3])
def test_refined_levels_match_oracle_direct_solve(ctx, level):
    """Level 3 (300 k DOFs) takes the oracle's sparse LU ~15 s; level 4 takes minutes (run once for
    DESIGN.md: rel-L2 2.0e-11 against scipy's direct solve)."""
    setup, fixed = refined_cantilever(level)
    assert len(setup.coors) > 16384            # too large for a cluster: streaming kernels
    orc = OracleProblem(setup.coors, setup.conn, num_steps=2)
    orc.fixed_vertex[:] = False
    orc.fixed_vertex[fixed] = True
    orc.load[:] = 0
    orc.load[3] = (0.0, -1000.0)
    u = orc.solve("best")[-1]
    with ctx.create_batch(pack([setup.sample])) as b:
        r = b.assemble().solve(1e-10, 200000).download()
        st = b.stats()
    assert r.status[0] == SAMPLE_CONVERGED and st["cluster_systems"] == 0 and st["spmv_launches_timed"] > 0
    assert np.linalg.norm(r.u - u) / np.linalg.norm(u) <= 1e-8


def test_level4_million_dof_solve_properties(ctx):
    setup, fixed = refined_cantilever(4)
    smp = setup.sample
    n_v = len(smp.coors)
    with ctx.create_batch(pack([smp])) as b:
        b.assemble()
        a, z = b.sample_sizes()
        assert a[0] == 2 * (n_v - len(fixed)) and a[0] > 1_150_000 and z[0] > 16_000_000   # SURVEY C3 sizes
        r = b.solve(1e-10, 400000).download()
        st = b.stats()
        K = b.csr(0)
    assert r.status[0] == SAMPLE_CONVERGED
    assert st["cluster_systems"] == 0 and st["refined_systems"] >= 1     # 12 k iterations: residual replacement kicks in
    active = ~smp.fixed.astype(bool)
    f = smp.rhs[active].reshape(-1)
    ua = r.u[active].reshape(-1)
    # true residual in the Jacobi-scaled norm the solver controls, and in the plain 2-norm
    res = f - K @ ua
    d = K.diagonal()
    true_rel = np.linalg.norm(res / np.sqrt(d)) / np.linalg.norm(f / np.sqrt(d))
    assert true_rel <= 3e-10 and abs(r.relres[0] - true_rel) <= 0.25 * true_rel   # relres reports the TRUE residual
    assert abs(K - K.T).max() <= 1e-9 * abs(K).max()
    assert np.all(r.u[~active] == 0.0)
    # linearity in the load (the load steps of a condition are multiples of one solve, F5)
    half = smp.__class__(smp.coors, smp.conn, smp.cell_region, smp.D, smp.fixed, 0.5 * smp.rhs)
    with ctx.create_batch(pack([half])) as b2:
        r2 = b2.assemble().solve(1e-10, 400000).download()
    assert np.linalg.norm(r2.u - 0.5 * r.u) / np.linalg.norm(r.u) <= 1e-9
    # refinement: the deflection under the load grows towards the (log-singular) point-load limit
    # by a shrinking amount; level 0 is sfepy's golden output
    g = cases.fixtures()["cantilever_u"][:, :2]
    assert g[3, 1] < 0 and r.u[3, 1] < g[3, 1] and abs(r.u[3, 1] - g[3, 1]) < 0.1 * abs(g[3, 1])
    # away from the singularity the field has converged: mid-span vertex of the coarse mesh
    mid = int(np.argmin(np.abs(setup.coors[:len(g), 0] - 0.5) + np.abs(setup.coors[:len(g), 1] - 0.5)))
    assert abs(r.u[mid, 1] - g[mid, 1]) <= 0.01 * abs(g[mid, 1])


@pytest.mark.parametrize("name,level", [("cantilever", 4), ("gusset", 3)])
def test_config3_meshes_match_the_stored_sparse_lu_solution(ctx, name, level):
    """BASELINE config 3 at full size (1.19 M / 1.28 M DOFs; reference applications/cantilever/cantilever.py:26-52,
    applications/gusset/gusset.py:39-94): the GPU solve against scipy's sparse LU + one refinement step, computed
    offline (tools/make_c3_reference.py, minutes of CPU) and stored at the coarse mesh's vertices, which keep
    their indices under uniform refinement (tests/golden/c3_lu.npz)."""
    import os
    from fea_diffusion_b200.workload import large_case
    ref = np.load(os.path.join(cases.ROOT, "tests", "golden", "c3_lu.npz"))
    setup, n0 = large_case(name, level)
    with ctx.create_batch(pack([setup.sample])) as b:
        b.assemble()
        a, _ = b.sample_sizes()
        assert int(a[0]) == int(ref["%s_L%d_n_dofs" % (name, level)])
        r = b.solve(1e-10, 400000).download()
        st = b.stats()
    assert r.status[0] == SAMPLE_CONVERGED and st["cluster_systems"] == 0
    g = ref["%s_L%d_u_coarse" % (name, level)]
    err = float(np.linalg.norm(r.u[:n0] - g) / np.linalg.norm(g))
    assert err <= 1e-8, err
