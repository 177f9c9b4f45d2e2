"""Host-side image geometry: the camera of the reference's off-screen render and the two-pass
window sizing + crop of ``datagen/generate.py:129-145`` (reference), in closed form.

pyvista's ``view_xy`` + VTK's bounding-sphere camera reset with a 30 degree view angle give, for a
square window of W pixels and a mesh bounding box (w, h):  S = W cos(15 deg) / sqrt(w^2 + h^2)
pixels per world unit, centred.  ``find_image_bounds`` (reference
``datagen/utils.py:18-56``) then sees the outline render's non-white box, which is the geometric
box grown by the outline's line thickness; 0.7 px per side reproduces every integer observable in
the reference's committed renders (any value in [0.45, 0.95] does).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

INITIAL_IMAGE_SIZE = math.ceil(512 / 0.685546875)  # reference fea_analysis.py:54
_COS15 = math.cos(math.pi / 12.0)
_OUTLINE_GROW = 0.7


def pixels_per_unit(window: int, bbox) -> float:
    w, h = bbox[2] - bbox[0], bbox[3] - bbox[1]
    return window * _COS15 / math.hypot(w, h)


def outline_bounds(window: int, bbox) -> Tuple[int, int, int, int]:
    """(left, top, right, bottom) that ``find_image_bounds`` reports for the outline render."""
    s = pixels_per_unit(window, bbox)
    hw, hh = 0.5 * (bbox[2] - bbox[0]) * s, 0.5 * (bbox[3] - bbox[1]) * s
    c = window / 2.0
    return (math.floor(c - hw - _OUTLINE_GROW), math.floor(c - hh - _OUTLINE_GROW),
            math.floor(c + hw + _OUTLINE_GROW), math.floor(c + hh + _OUTLINE_GROW))


def plate_window(bbox, image_size: int, initial: int = INITIAL_IMAGE_SIZE):
    """Restates generate.py:129-145: (modified_image_size, bounds=(l, l, u, u))."""
    left, top, right, bottom = outline_bounds(initial, bbox)
    max_size = max(right - left, bottom - top)
    modified = round(image_size / (max_size / initial))
    left, top, right, bottom = outline_bounds(modified, bbox)
    lo, hi = (left, right) if right > bottom else (top, bottom)
    return modified, (lo, lo, hi, hi)


def crop_affine(bbox, window: int, bounds) -> np.ndarray:
    """(ax, bx, ay, by) mapping world coordinates to pixel coordinates of the cropped image:
    px = ax * x + bx, py = ay * y + by (y points down in the image)."""
    s = pixels_per_unit(window, bbox)
    cx, cy = 0.5 * (bbox[0] + bbox[2]), 0.5 * (bbox[1] + bbox[3])
    c = window / 2.0
    return np.array([s, c - cx * s - bounds[0], -s, c + cy * s - bounds[1]], dtype=np.float64)
