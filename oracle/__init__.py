"""CPU oracle for the fea-diffusion FEA data-synthesis hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
algorithm the reference delegates to sfepy 2023.3 + scipy SuperLU + VTK inside
``datagen/fea_analysis.py`` (reference) -- see SURVEY.md Appendix A for the
rule-by-rule specification.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``; the product path (``fea_diffusion_b200``) never routes
through it and fails loudly when the CUDA library is missing.

Parity pins (checked by ``tests/test_oracle_golden.py``):
  * ``applications/cantilever/cantilever.vtk``  fp64 ``u``  (rel-L2 <= 1e-10)
  * ``applications/shearblade/shearblade.vtk``  fp64 ``u``  (rel-L2 <= 1e-10)
  * sfepy log pins: shape 10466 / nnz 144084 / ||r0|| 3001.666 (shearblade),
    shape 19672 / nnz 270712 / ||r0|| 800 / 19307 cells (composite).
Unpinned (no reference artefact exists): Q1 quads, edge-force facet rule,
well-posed multi-material solves, auto-range PNGs.  Those say so where they
are implemented.
"""
from .mesh_io import read_medit, read_vtk_legacy, write_vtk_legacy  # noqa: F401
from .fea_oracle import (  # noqa: F401
    OracleProblem,
    assemble_csr,
    complete_cells,
    element_stiffness,
    fix_orientation,
    plane_strain_D,
    points_in_list,
    points_on_edge,
    facet_region_vertices,
    solve_load_steps,
)
from .raster_oracle import (  # noqa: F401
    camera_scale,
    closed_form_window,
    gray_from_scalar,
    rasterize_scalar,
)
