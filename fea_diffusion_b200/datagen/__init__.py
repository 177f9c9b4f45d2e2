"""Mirror of the reference's ``datagen`` package for the FEA path: ``FEAnalysis`` and
``generate_data`` with the reference's call surface, backed by the CUDA library."""
from .fea_analysis import FEAnalysis  # noqa: F401
from .generate import generate_data  # noqa: F401
from .utils import find_image_bounds, verify_directory  # noqa: F401
