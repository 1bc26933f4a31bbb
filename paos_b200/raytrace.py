"""Paraxial chief-ray trace through an optical chain (diagnostic; host scalars only).

Mirror of ``paos.core.raytrace.raytrace`` (``paos/core/raytrace.py:7-60``): the same two-vector walk that ``run`` performs
for the aperture centres (``paos/core/run.py:69-70, :213-214``), reported per surface as text.
"""
import numpy as np

from .coordinate_break import coordinate_break

_LINE = "S{:02d} - {:15s} y:{:7.3f}mm ut:{:10.3e} rad x:{:7.3f}mm us:{:10.3e} rad"


def raytrace(field, opt_chain, x=0.0, y=0.0):
    """Trace the ray that starts at ``(x, y)`` with slopes ``field = {'ut': .., 'us': ..}``; returns one formatted line per
    surface (height in mm and slope in rad, tangential then sagittal plane), in the reference's format."""
    tangential = np.array([y, field["ut"]])
    sagittal = np.array([x, field["us"]])
    lines = []
    for num, item in opt_chain.items():
        if item["type"] == "Coordinate Break":
            tangential, sagittal = coordinate_break(tangential, sagittal, item["xdec"], item["ydec"], item["xrot"],
                                                    item["yrot"], 0.0)
        tangential = item["ABCDt"]() @ tangential
        sagittal = item["ABCDs"]() @ sagittal
        lines.append(_LINE.format(num, item["name"], 1000 * tangential[0], tangential[1], 1000 * sagittal[0], sagittal[1]))
    return lines
