"""Shared helpers of the parity tests: error metric, matched oracle/device wavefront pairs."""
import numpy as np

TOL = {"complex128": 1e-10, "complex64": 1e-4}  # BASELINE.json north_star tolerances


def relerr(got, ref):
    """max|got - ref| / max|ref| -- the parity metric of SURVEY.md section 7.3 item 3."""
    ref = np.asarray(ref)
    got = np.asarray(got)
    denom = np.max(np.abs(ref))
    if denom == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref)) / denom)


def random_field(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex128)


class Pair:
    """An oracle WFO and a device WFO driven with the same calls."""

    def __init__(self, D, wl, n, zoom, dtype="complex128", field=None):
        from oracle import paos_np
        import paos_b200

        self.o = paos_np.WFO(D, wl, n, zoom)
        self.d = paos_b200.WFO(D, wl, n, zoom, dtype=dtype)
        self.dtype = dtype
        if field is not None:
            self.o._wfo = field.copy()
            self.d.wfo = field

    def call(self, name, *a, **k):
        ro = getattr(self.o, name)(*a, **k)
        rd = getattr(self.d, name)(*a, **k)
        return ro, rd

    def check_scalars(self):
        for k in ("wl", "z", "w0", "zw0", "zr", "dx", "dy", "C", "fratio", "wz", "distancetofocus"):
            a, b = getattr(self.o, k), getattr(self.d, k)
            assert a == b or (np.isnan(a) and np.isnan(b)), (k, a, b)

    def check(self, tol=None, what=("wfo", "amplitude")):
        tol = TOL[self.dtype] if tol is None else tol
        self.check_scalars()
        errs = {}
        for k in what:
            ref = self.o._wfo if k == "wfo" else getattr(self.o, k)
            errs[k] = relerr(getattr(self.d, k), ref)
            assert errs[k] <= tol, (k, errs[k], tol)
        return errs
