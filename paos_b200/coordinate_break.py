"""Coordinate break of the paraxial chief-ray vectors (host scalars; reference
``paos/core/coordinateBreak.py:7-72``): decenter, then intrinsic XYZ rotation, then re-intersection with the
new z = 0 plane."""
import numpy as np


def coordinate_break(vt, vs, xdec, ydec, xrot, yrot, zrot, order=0):
    from scipy.spatial.transform import Rotation

    if order != 0:
        raise ValueError("Coordinate break orders other than 0 not implemented yet")
    xdec, ydec, xrot, yrot, zrot = (v if np.isfinite(v) else 0.0 for v in (xdec, ydec, xrot, yrot, zrot))
    to_new_frame = Rotation.from_euler("xyz", [xrot, yrot, zrot], degrees=True).inv()
    point = to_new_frame.apply([vs[0] - xdec, vt[0] - ydec, 0.0])
    direction = to_new_frame.apply([vs[1], vt[1], 1])
    direction /= direction[2]
    hit = point - direction * point[2] / direction[2]
    return np.array([hit[1], direction[1]]), np.array([hit[0], direction[0]])
