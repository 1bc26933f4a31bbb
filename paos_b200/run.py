"""Chain driver: drop-in for ``paos.core.run.run`` (reference ``paos/core/run.py:30-228``).

The per-surface logic (coordinate break -> aperture -> stop -> aberration -> snapshot -> magnification ->
medium -> lens -> propagation -> ray/ABCD update) is the reference's; what differs is *when* arrays exist:
the wavefront stays in HBM for the whole chain, the library records the surface operations and fuses them
into a few sweeps, and a snapshot (``amplitude``, ``phase``, ``wfo``) is read back only for surfaces with
``save=True`` -- the reference computes all three at every surface and throws most of them away
(``run.py:179``).  Nothing on the host ever looks at wavefront data to take a decision, so this is exact.
"""
from copy import deepcopy

import numpy as np

from .abcd import ABCD
from .coordinate_break import coordinate_break
from .wfo import WFO

ARRAY_KEYS = ("amplitude", "phase", "wfo")


def push_results(wfo, keys=None):
    """Snapshot of the current surface (``run.py:12-27``).  ``keys`` limits which of the three N x N arrays
    (``amplitude``, ``phase``, ``wfo``) are read back from the device; ``None`` reads all, like the reference."""
    want = ARRAY_KEYS if keys is None else tuple(k for k in ARRAY_KEYS if k in keys)
    out = {}
    for k in want:
        out[k] = getattr(wfo, k)
    out.update(
        wz=wfo.wz, distancetofocus=wfo.distancetofocus, fratio=wfo.fratio, dx=wfo.dx, dy=wfo.dy, wl=wfo.wl,
        extent=wfo.extent, propagator=wfo.propagator,
    )
    return out


class AsyncSnapshots:
    """Saved surfaces without stalling the chain (the device half of the reference's output path, ``saveOutput.py`` /
    ``pipeline.py:72-84``): every requested array of a ``save=True`` surface is read out on the device
    (``paos_wfo_read_device``: fused into the pass that produces the surface) and copied to pinned host memory
    asynchronously on the wavefront's stream, while the host goes on recording the next surfaces.  The arrays handed out
    are numpy views of those pinned buffers; they hold valid data once :meth:`finish` has returned."""

    _DEVICE_READ = {"amplitude": "amplitude_device", "phase": "phase_device", "wfo": "wfo_device"}

    def __init__(self, wfo):
        self.wfo = wfo

    def take(self, keys=None):
        import torch

        wfo = self.wfo
        want = ARRAY_KEYS if keys is None else tuple(k for k in ARRAY_KEYS if k in keys)
        out = {}
        for k in want:
            dev = getattr(wfo, self._DEVICE_READ[k])()
            host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
            with torch.cuda.stream(wfo._stream):
                host.copy_(dev, non_blocking=True)
            out[k] = host.numpy()
        out.update(
            wz=wfo.wz, distancetofocus=wfo.distancetofocus, fratio=wfo.fratio, dx=wfo.dx, dy=wfo.dy, wl=wfo.wl,
            extent=wfo.extent, propagator=wfo.propagator,
        )
        return out

    def finish(self):
        self.wfo.sync()


def _surface_aperture(wfo, item, vt, vs):
    ap = item["aperture"]
    xc = ap["xc"] if np.isfinite(ap["xc"]) else vs[0]
    yc = ap["yc"] if np.isfinite(ap["yc"]) else vt[0]
    xrad = ap["xrad"] * np.sqrt(1 / (vs[1] ** 2 + 1))
    yrad = ap["yrad"] * np.sqrt(1 / (vt[1] ** 2 + 1))
    if not np.all(np.isfinite([xrad, yrad])):
        return None
    return wfo.aperture(xc - vs[0], yc - vt[0], hx=xrad, hy=yrad, shape=ap["shape"],
                        obscuration=ap["type"] != "aperture")


def run(pupil_diameter, wavelength, gridsize, zoom, field, opt_chain, *, device=0, dtype="complex128",
        stream=None, keys=None, psd_noise=None, wfo=None, snapshot=None, async_snapshots=False):
    """Run the physical-optics propagation of one wavelength / field through ``opt_chain``.

    Positional parameters and the returned ``{surface_num: {...}}`` dictionary are the reference's.  Keyword-only
    extras: ``device``/``dtype``/``stream`` are passed to :class:`WFO`; ``keys`` limits the arrays read back per
    saved surface (e.g. ``("amplitude",)``, the reference pipeline's ``store_keys``); ``psd_noise`` is a callable
    ``(surface_num, shape) -> (n1, n2)`` injecting the PSD noise draws (bit-parity mode); ``wfo`` re-uses an
    existing :class:`WFO` (its buffer and stream) instead of allocating one; ``snapshot`` replaces
    :func:`push_results` for saved surfaces, e.g. to keep read-outs on the device (``snapshot(wfo, item) -> dict``);
    ``async_snapshots`` streams the saved arrays to pinned host memory while the chain continues (:class:`AsyncSnapshots`;
    same values, the result dictionaries hold views of pinned buffers).
    """
    assert isinstance(opt_chain, dict), "opt_chain must be a dict"
    results = {}
    vt = np.array([0.0, field["ut"]])
    vs = np.array([0.0, field["us"]])
    total_t, total_s = ABCD(), ABCD()
    if wfo is None:
        wfo = WFO(pupil_diameter, wavelength, gridsize, zoom, device=device, dtype=dtype, stream=stream)
    else:
        assert wfo.grid_size == gridsize, "the re-used WFO has a different grid size"
        wfo.reset(pupil_diameter, wavelength, zoom)
    streamer = AsyncSnapshots(wfo) if (async_snapshots and snapshot is None) else None

    for item in opt_chain.values():
        if item["type"] == "Coordinate Break":
            vt, vs = coordinate_break(vt, vs, item["xdec"], item["ydec"], item["xrot"], item["yrot"], 0.0)

        save = bool(item["save"])
        snap = {"aperture": None}
        if "aperture" in item:
            snap["aperture"] = _surface_aperture(wfo, item, vt, vs)
        if item["is_stop"]:
            wfo.make_stop()

        kind = item["type"]
        if kind == "Zernike":
            radius = item["Zradius"] if np.isfinite(item["Zradius"]) else wfo.wz
            zmask = False
            if item["Zorthonorm"]:
                assert "aperture" in item, "Zorthonorm requires aperture"
                zmask = ~snap["aperture"].to_mask(method="exact").to_image((wfo.grid_size, wfo.grid_size)).astype(bool)
            wfe = wfo.zernikes(item["Zindex"], item["Z"], item["Zordering"], item["Znormalize"], radius,
                               origin=item["Zorigin"], orthonorm=item["Zorthonorm"], mask=zmask, return_wfe=save)
            if save:
                snap["wfe"] = wfe
        elif kind == "Grid Sag":
            snap["wfe"] = wfo.grid_sag(item["grid_sag"], item["nx"], item["ny"], item["delx"], item["dely"],
                                       item["xdec"], item["ydec"])
        elif kind == "PSD":
            noise = psd_noise(item["num"], (wfo.grid_size, wfo.grid_size)) if psd_noise is not None else None
            wfe = wfo.psd(item["A"], item["B"], item["C"], item["fknee"], item["fmin"], item["fmax"], item["SR"],
                          item["units"], noise=noise, return_wfe=save)
            if save:
                snap["wfe"] = wfe

        if save:
            if streamer is not None:
                snap.update(streamer.take(keys))
            else:
                snap.update(push_results(wfo, keys) if snapshot is None else snapshot(wfo, item))

        abcd_t, abcd_s = item["ABCDt"], item["ABCDs"]
        Ms, Mt = abcd_s.M, abcd_t.M
        fl = np.inf if abcd_t.power == 0 else abcd_t.cout / abcd_t.power
        thickness = abcd_t.cout * abcd_t.thickness
        n1n2 = abcd_t.n1n2
        if Mt != 1.0 or Ms != 1.0:
            wfo.Magnification(Mt, Ms)
        if np.abs(n1n2) != 1.0:
            wfo.ChangeMedium(n1n2)
        if np.isfinite(fl):
            wfo.lens(fl)
        if np.isfinite(thickness) and np.abs(thickness) > 1e-10:
            wfo.propagate(thickness)

        vt = abcd_t() @ vt
        vs = abcd_s() @ vs
        total_t = abcd_t * total_t
        total_s = abcd_s * total_s
        if save:
            snap["ABCDt"], snap["ABCDs"] = total_t, total_s
            results[item["num"]] = deepcopy(snap) if (snapshot is None and streamer is None) else snap

    if streamer is not None:
        streamer.finish()
    return results
