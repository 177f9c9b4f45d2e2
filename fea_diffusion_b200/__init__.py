"""fea_diffusion_b200 -- B200-native batched finite-element solver for the data-synthesis hot path
of namanxkumar/fea-diffusion (the solve + displacement rasterisation behind
``datagen/fea_analysis.py``).  CUDA kernels live in ``csrc/`` behind the C-ABI of
``include/fea_b200.h``; this package is the thin ctypes host layer plus the drop-in
``datagen.fea_analysis.FEAnalysis`` surface.  (The directory is named with an underscore so it is
importable; it is the ``fea-diffusion_b200`` package of the project layout.)
"""
from ._capi import FeaError, load_library  # noqa: F401
from .solver import Batch, BatchResult, Context, PackedBatch, PackedConditions, Sample, pack  # noqa: F401
from .host import ProblemSetup, read_mesh, stiffness_plane_strain  # noqa: F401

__all__ = ["FeaError", "load_library", "Batch", "BatchResult", "Context", "PackedBatch", "Sample",
           "pack", "ProblemSetup", "read_mesh", "stiffness_plane_strain"]
