"""Helpers of the reference's ``datagen/utils.py`` that the orchestration loop uses."""
import os

import numpy as np
from PIL import Image


def verify_directory(directory):
    if not os.path.exists(directory):
        os.makedirs(directory)


def find_image_bounds(image_path):
    """(left, top, right, bottom) of the non-white content, with the reference's scan semantics
    (``datagen/utils.py:18-56``): each side is the first row/column holding a non-white pixel;
    a side whose scan ends on its initial value keeps it (left/top 0, right/bottom = size)."""
    rgb = np.asarray(Image.open(image_path).convert("RGB"))
    ink = (rgb != 255).any(axis=2)
    h, w = ink.shape
    cols = np.flatnonzero(ink.any(axis=0))
    rows = np.flatnonzero(ink.any(axis=1))
    if len(cols) == 0:
        return 0, 0, w, h
    # the reference's loops only stop at a non-zero hit, so content touching column/row 0 makes
    # them run on to the next inked column/row after it
    left = int(cols[cols > 0][0]) if (cols > 0).any() else 0
    top = int(rows[rows > 0][0]) if (rows > 0).any() else 0
    return left, top, int(cols[-1]), int(rows[-1])
