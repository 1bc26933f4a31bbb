"""Native chain execution: compile an ``opt_chain`` dictionary into ``paos_surface`` records and run the whole
per-surface loop inside ``libpaos_b200.so`` (``paos_chain_run``), so that a sweep over hundreds of wavelengths is
not limited by the Python interpreter.  The records carry exactly the numbers ``paos_b200.run`` would read from
the dictionary (reference data contract: ``paos/core/parseConfig.py:155-394``); results are those of ``run``.
"""
import ctypes as C

import numpy as np

from . import _lib
from .zernike import j2mn, zernike_norms

SURF_GENERIC, SURF_COORDBREAK, SURF_ZERNIKE, SURF_SCREEN, SURF_PSD, SURF_GRIDSAG = 0, 1, 2, 3, 4, 5
_TYPES = {"Standard": SURF_GENERIC, "Paraxial Lens": SURF_GENERIC, "ABCD": SURF_GENERIC,
          "Coordinate Break": SURF_COORDBREAK, "Zernike": SURF_ZERNIKE, "Grid Sag": SURF_GRIDSAG, "PSD": SURF_PSD}
_SHAPES = {"elliptical": _lib.SHAPE_ELLIPSE, "rectangular": _lib.SHAPE_RECT}


class Surface(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("is_stop", C.c_int), ("save", C.c_int), ("has_aperture", C.c_int), ("ap_shape", C.c_int),
        ("ap_obscuration", C.c_int), ("read_what", C.c_int), ("zernike_terms", C.c_int), ("zernike_origin", C.c_int),
        ("screen_on_device", C.c_int), ("read_discard", C.c_int), ("reserved", C.c_int),
        ("ap_xrad", C.c_double), ("ap_yrad", C.c_double), ("ap_xc", C.c_double), ("ap_yc", C.c_double),
        ("abcd_t", C.c_double * 4), ("abcd_s", C.c_double * 4), ("cout_t", C.c_double),
        ("xdec", C.c_double), ("ydec", C.c_double), ("xrot", C.c_double), ("yrot", C.c_double),
        ("zernike_radius", C.c_double), ("screen_dx", C.c_double), ("screen_dy", C.c_double), ("psd", C.c_double * 8), ("psd_seed", C.c_uint64),
        ("zernike_m", C.POINTER(C.c_int)), ("zernike_n", C.POINTER(C.c_int)), ("zernike_coef", C.POINTER(C.c_double)),
        ("screen", C.POINTER(C.c_double)), ("psd_noise1", C.POINTER(C.c_double)), ("psd_noise2", C.POINTER(C.c_double)),
        ("read_dst", C.c_void_p),
        ("sag", C.POINTER(C.c_double)), ("sag_mask", C.POINTER(C.c_ubyte)), ("sag_nx", C.c_int), ("sag_ny", C.c_int),
        ("sag_delx", C.c_double), ("sag_dely", C.c_double), ("sag_xdec", C.c_double), ("sag_ydec", C.c_double),
        ("sag_key", C.c_uint64),
    ]


class Snapshot(C.Structure):
    _fields_ = [
        ("surface", C.c_int), ("propagator", C.c_char * 4),
        ("wl", C.c_double), ("z", C.c_double), ("w0", C.c_double), ("zw0", C.c_double), ("zr", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("C", C.c_double), ("fratio", C.c_double), ("wz", C.c_double),
        ("distancetofocus", C.c_double), ("vt", C.c_double * 2), ("vs", C.c_double * 2),
    ]

    def as_dict(self, n):
        return dict(wz=self.wz, distancetofocus=self.distancetofocus, fratio=self.fratio, dx=self.dx, dy=self.dy, wl=self.wl,
                    extent=(-n // 2 * self.dx, (n // 2 - 1) * self.dx, -n // 2 * self.dy, (n // 2 - 1) * self.dy),
                    propagator=self.propagator.decode())


class ChainArgs(C.Structure):
    """``paos_chain_args`` of ``include/paos_b200.h``: the arguments of ``paos_chain_run`` for one item of a batch."""

    _fields_ = [
        ("pupil_diameter", C.c_double), ("wavelength", C.c_double), ("zoom", C.c_double), ("us", C.c_double), ("ut", C.c_double),
        ("surfaces", C.c_void_p), ("n_surfaces", C.c_int), ("max_snapshots", C.c_int), ("snapshots", C.c_void_p),
        ("n_snapshots", C.c_void_p), ("final_state", C.c_void_p),
    ]


_NAN = float("nan")
_ZTABLES = {}


def _zernike_tables(K, ordering, normalize):
    """``(m, n, norms)`` of the first K polynomials as contiguous int32 / float64 arrays, computed once per (K, ordering,
    normalize): a Monte-Carlo sweep compiles the same tables for every realization."""
    key = (int(K), str(ordering), str(normalize))
    hit = _ZTABLES.get(key)
    if hit is None:
        m, n = j2mn(K, ordering)
        hit = (np.ascontiguousarray(m, dtype=np.int32), np.ascontiguousarray(n, dtype=np.int32),
               np.ascontiguousarray(zernike_norms(m, n, normalize), dtype=np.float64))
        if len(_ZTABLES) < 256:
            _ZTABLES[key] = hit
    return hit


# numpy view of the two 2 x 2 matrices inside the paos_surface records: the ctypes array is filled in two vectorised writes
# instead of two ctypes array constructions per surface (which were 60 % of the time of building the records)
_ABCD_VIEW = np.dtype({"names": ["abcd_t", "abcd_s"], "formats": [("f8", (4,)), ("f8", (4,))],
                       "offsets": [Surface.abcd_t.offset, Surface.abcd_s.offset], "itemsize": C.sizeof(Surface)})


class CompiledChain:
    """``paos_surface`` array of one job plus the host arrays it points to (kept alive here)."""

    def __init__(self, opt_chain, gridsize, pupil_diameter, zoom, psd_seed=0, psd_noise=None, device=None, screen_cache=None):
        items = list(opt_chain.values())
        self.n = int(gridsize)
        self.count = len(items)
        self.array = (Surface * max(self.count, 1))()
        self.keep = []
        self.nums = []
        self.saved = []
        for i, item in enumerate(items):
            s = self.array[i]
            kind = item["type"]
            if kind not in _TYPES:
                raise ValueError(f"Surface Type not recognised: {kind}")
            s.type = _TYPES[kind]
            s.is_stop = 1 if item["is_stop"] else 0
            s.save = 1 if item["save"] else 0
            s.read_what = -1
            if s.save:
                self.saved.append(i)
            self.nums.append(item["num"])
            if "aperture" in item:
                ap = item["aperture"]
                if ap["shape"] not in _SHAPES:
                    raise ValueError(f"Aperture {ap['shape']} not supported by the native chain runner")
                s.has_aperture = 1
                s.ap_shape = _SHAPES[ap["shape"]]
                s.ap_obscuration = 0 if ap["type"] == "aperture" else 1
                s.ap_xrad, s.ap_yrad, s.ap_xc, s.ap_yc = float(ap["xrad"]), float(ap["yrad"]), float(ap["xc"]), float(ap["yc"])
            s.cout_t = float(item["ABCDt"]._cout)
            s.zernike_radius = _NAN
            if s.type == SURF_COORDBREAK:
                s.xdec, s.ydec, s.xrot, s.yrot = (float(item[k]) for k in ("xdec", "ydec", "xrot", "yrot"))
            elif s.type == SURF_ZERNIKE:
                if item["Zorthonorm"]:
                    raise NotImplementedError("PolyOrthoNorm screens are not on the device path yet")
                if item["Zorigin"] not in ("x", "y"):
                    raise ValueError(f"Origin {item['Zorigin']} not recognised. Origin shall be either x or y")
                index = np.asarray(item["Zindex"])
                assert not np.any(np.diff(index) - 1), "Zernike sequence should be continuous"
                K = len(index)
                m32, n32, norms = _zernike_tables(K, item["Zordering"], item["Znormalize"])
                coef = np.ascontiguousarray(np.asarray(item["Z"], dtype=np.float64) * norms)
                self.keep += [coef, m32, n32]
                s.zernike_terms = K
                s.zernike_origin = 0 if item["Zorigin"] == "x" else 1
                s.zernike_m = m32.ctypes.data_as(C.POINTER(C.c_int))
                s.zernike_n = n32.ctypes.data_as(C.POINTER(C.c_int))
                s.zernike_coef = coef.ctypes.data_as(C.POINTER(C.c_double))
                s.zernike_radius = float(item["Zradius"])
            elif s.type == SURF_GRIDSAG:
                # the raw map travels as the lens file gives it; the library resamples it on the device at the pitch the
                # beam has at this surface (wfo.py:848-862), and jobs that carry the same map share the prepared screen
                data, mask, key = _raw_sag(item, screen_cache)
                self.keep += [data, mask]
                s.sag = data.ctypes.data_as(C.POINTER(C.c_double))
                if mask is not None:
                    s.sag_mask = mask.ctypes.data_as(C.POINTER(C.c_ubyte))
                s.sag_nx, s.sag_ny = int(item["nx"]), int(item["ny"])
                s.sag_delx, s.sag_dely = float(item["delx"]), float(item["dely"])
                s.sag_xdec, s.sag_ydec = float(item["xdec"]), float(item["ydec"])
                s.sag_key = key
            elif s.type == SURF_PSD:
                from .wfo import _unit_to_m

                vals = [item["A"], item["B"], item["C"], item["fknee"], item["fmin"], item["fmax"], item["SR"], _unit_to_m(item["units"])]
                s.psd[:] = [float(v) for v in vals]
                s.psd_seed = int(psd_seed)
                if psd_noise is not None:
                    n1, n2 = psd_noise(item["num"], (self.n, self.n))
                    n1 = np.ascontiguousarray(n1, dtype=np.float64)
                    n2 = np.ascontiguousarray(n2, dtype=np.float64)
                    self.keep += [n1, n2]
                    s.psd_noise1 = n1.ctypes.data_as(C.POINTER(C.c_double))
                    s.psd_noise2 = n2.ctypes.data_as(C.POINTER(C.c_double))
        if self.count:
            view = np.frombuffer(self.array, dtype=_ABCD_VIEW, count=self.count)
            view["abcd_t"] = np.array([item["ABCDt"]._ABCD for item in items], dtype=np.float64).reshape(self.count, 4)
            view["abcd_s"] = np.array([item["ABCDs"]._ABCD for item in items], dtype=np.float64).reshape(self.count, 4)
        self.snapshots = (Snapshot * max(len(self.saved), 1))()
        self.final = Snapshot()
        self.nsnap = C.c_int(0)

    def set_readout(self, surface_index, what, dev_ptr, final=False):
        """``final``: the read-out is the last use of the wavefront (honoured for the last surface of the chain only)."""
        s = self.array[surface_index]
        s.read_what = int(what)
        s.read_dst = dev_ptr
        s.read_discard = 1 if final else 0


def _raw_sag(item, cache):
    """``(data, mask, key)`` of a Grid Sag item: contiguous float64 samples (masked ones as 0), the explicit mask of a masked
    array as bytes (None: the library masks non-finite and zero samples, wfo.py:757-765) and a 64-bit content key.  The
    digest is computed once per map object (``cache``: id of the array -> result, kept by the Sweep)."""
    import hashlib

    sag = item["grid_sag"]
    memo = cache.get(("raw", id(sag))) if cache is not None else None
    if memo is not None and memo[3] is sag:
        return memo[0], memo[1], memo[2]
    assert sag.ndim == 2, "sag shall be a 2D array"
    assert sag.shape == (int(item["ny"]), int(item["nx"]))
    mask = None
    if isinstance(sag, np.ma.MaskedArray):
        mask = np.ascontiguousarray(np.ma.getmaskarray(sag), dtype=np.uint8)
        data = np.ascontiguousarray(sag.filled(0.0), dtype=np.float64)
    else:
        data = np.ascontiguousarray(sag, dtype=np.float64)
    h = hashlib.sha1(data.tobytes())
    if mask is not None:
        h.update(mask.tobytes())
    key = int.from_bytes(h.digest()[:8], "little") or 1
    if cache is not None:
        cache[("raw", id(sag))] = (data, mask, key, sag)
    return data, mask, key


def compile_job(job, psd_noise=None, device=None, screen_cache=None):
    """Compile (and cache on the job dict) the native surface records of a job.  With ``device`` the grid-sag maps are
    uploaded once (shared through ``screen_cache``) instead of travelling from host memory on every run."""
    cc = job.get("_compiled")
    sig = (None if device is None else int(device), int(job["gridsize"]), float(job["zoom"]), float(job["pupil_diameter"]))
    if cc is not None and getattr(cc, "signature", None) != sig:
        cc = None  # compiled for another GPU (raw device pointers) or another geometry: rebuild
    if cc is None or psd_noise is not None:
        cc = CompiledChain(job["opt_chain"], job["gridsize"], job["pupil_diameter"], job["zoom"],
                           psd_seed=job.get("psd_seed", 0), psd_noise=psd_noise, device=device, screen_cache=screen_cache)
        cc.signature = sig
        if psd_noise is None:
            job["_compiled"] = cc
    return cc


def run_compiled(wfo, job, cc):
    """Enqueue one compiled chain on ``wfo`` (asynchronous); returns the list of snapshot dicts of saved surfaces."""
    _lib.check(_lib.lib.paos_chain_run(
        wfo._handle, float(job["pupil_diameter"]), float(job["wavelength"]), float(job["zoom"]), float(job["field"]["us"]),
        float(job["field"]["ut"]), cc.array, cc.count, cc.snapshots, len(cc.saved), C.byref(cc.nsnap), C.byref(cc.final)))
    wfo._sync_scalars(cc.final)
    return [cc.snapshots[k].as_dict(cc.n) for k in range(min(cc.nsnap.value, len(cc.saved)))]


def run_compiled_batch(wfos, jobs, ccs):
    """Enqueue the compiled chains of ``jobs`` on ``wfos`` (same stream, grid size and precision) as ONE batch
    (``paos_batch_chain_run``): the same-axis passes of all items share a kernel launch.  Returns one snapshot list per job."""
    nb = len(jobs)
    args = (ChainArgs * nb)()
    handles = (C.c_void_p * nb)()
    for b, (wfo, job, cc) in enumerate(zip(wfos, jobs, ccs)):
        handles[b] = wfo._handle
        a = args[b]
        a.pupil_diameter, a.wavelength, a.zoom = float(job["pupil_diameter"]), float(job["wavelength"]), float(job["zoom"])
        a.us, a.ut = float(job["field"]["us"]), float(job["field"]["ut"])
        a.surfaces = C.cast(cc.array, C.c_void_p)
        a.n_surfaces = cc.count
        a.max_snapshots = len(cc.saved)
        a.snapshots = C.cast(cc.snapshots, C.c_void_p)
        a.n_snapshots = C.cast(C.pointer(cc.nsnap), C.c_void_p)
        a.final_state = C.cast(C.pointer(cc.final), C.c_void_p)
    _lib.check(_lib.lib.paos_batch_chain_run(handles, nb, C.cast(args, C.c_void_p)))
    out = []
    for wfo, cc in zip(wfos, ccs):
        wfo._sync_scalars(cc.final)
        out.append([cc.snapshots[k].as_dict(cc.n) for k in range(min(cc.nsnap.value, len(cc.saved)))])
    return out
