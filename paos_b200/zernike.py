"""Host-side Zernike bookkeeping: index conversions and normalisation factors.

Only the O(K) integer/scalar part of ``paos/classes/zernike.py`` lives on the host (``j2mn :132-176``,
``mn2j :179-207``, the norms of ``:77-83``); the polynomials themselves are evaluated per pixel on the device
(``csrc/aux_kernels.cu``, ``zernike_kernel``).
"""
import numpy as np

ORDERINGS = ("ansi", "standard", "noll", "fringe")


def j2mn(N, ordering):
    """Azimuthal and radial numbers ``(m, n)`` of the first ``N`` polynomials of an ordering."""
    if ordering not in ORDERINGS:
        raise NameError("Ordering not supported.")
    j = np.arange(N, dtype=int)
    if ordering in ("ansi", "standard"):
        n = np.ceil((-3.0 + np.sqrt(9.0 + 8.0 * j)) / 2.0).astype(int)
        m = 2 * j - n * (n + 2)
        if ordering == "standard":
            m = -m
        return m, n
    idx = j + 1
    if ordering == "noll":
        n = ((0.5 * (np.sqrt(8 * idx - 7) - 3)) + 1).astype(int)
        base = n * (n + 1) // 2 + 1
        par = n % 2  # even radial order: m = 0, 2, 2, 4, 4...; odd: m = 1, 1, 3, 3...
        m = np.where(par == 0, (idx - base + 1) // 2 * 2, (idx - base) // 2 * 2 + 1)
        m = np.where(idx % 2 == 0, m, -m)
        return m.astype(int), n
    # fringe
    half = np.ceil(np.sqrt(idx)) - 1          # (n + |m|) / 2
    first = half**2 + 1                       # first index of the group
    n = half + np.floor((idx - first) / 2)
    m = (2 * half - n) * (1 - np.mod(idx - first, 2) * 2)
    return m.astype(int), n.astype(int)


def mn2j(m, n, ordering):
    """Index of the polynomial ``(m, n)``: 0-based for 'ansi'/'standard', 1-based (as published) for 'noll'
    and 'fringe' -- the conventions of ``zernike.py:179-207``."""
    m = np.atleast_1d(np.asarray(m, dtype=int))
    n = np.atleast_1d(np.asarray(n, dtype=int))
    if ordering == "ansi":
        return (n * (n + 2) + m) // 2
    if ordering == "standard":
        return (n * (n + 2) - m) // 2
    if ordering == "fringe":
        half = (n + np.abs(m)) // 2
        return (half + 1) ** 2 - 2 * np.abs(m) + (m < 0).astype(int)
    if ordering == "noll":
        low = np.isin(n % 4, (0, 1))
        # published Noll rule: within a radial order the even index goes to m > 0
        p = np.where(m == 0, 1, np.where((m > 0) == low, 0, 1))
        return n * (n + 1) // 2 + np.abs(m) + p
    raise NameError("Ordering not supported.")


def zernike_norms(m, n, normalize):
    """``sqrt(n+1)`` (m = 0) or ``sqrt(2(n+1))`` when ``normalize`` else ones (``zernike.py:77-83``)."""
    m = np.asarray(m)
    n = np.asarray(n)
    if not normalize:
        return np.ones(len(m), dtype=np.float64)
    return np.array([np.sqrt(nn + 1) if mm == 0 else np.sqrt(2.0 * (nn + 1)) for mm, nn in zip(m, n)], dtype=np.float64)


def _points_call(device, m, n, norm, rho, phi, mask, mat=None, want_stack=True, want_cov=False):
    """One call of ``paos_zernike_points``: the polynomials are evaluated on the device (no CPU fallback)."""
    import ctypes as C

    from . import _lib

    K = len(m)
    if K > 64:
        raise ValueError("at most 64 polynomials per stack")
    shape = np.shape(rho)
    rho_c = np.ascontiguousarray(np.ma.getdata(rho), dtype=np.float64).ravel()
    phi_c = np.ascontiguousarray(np.ma.getdata(phi), dtype=np.float64).ravel()
    if phi_c.size != rho_c.size:
        raise ValueError("phi must have the shape of rho")
    mask_c = None if mask is None else np.ascontiguousarray(np.broadcast_to(mask, shape), dtype=np.uint8).ravel()
    m32, n32 = np.ascontiguousarray(m, dtype=np.int32), np.ascontiguousarray(n, dtype=np.int32)
    norm_c = np.ascontiguousarray(norm, dtype=np.float64)
    mat_c = None if mat is None else np.ascontiguousarray(mat, dtype=np.float64)
    out = np.empty((K,) + tuple(shape), dtype=np.float64) if want_stack else None
    cov = np.empty((K, K), dtype=np.float64) if want_cov else None
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib.paos_zernike_points(
        int(device), K, m32.ctypes.data_as(C.POINTER(C.c_int)), n32.ctypes.data_as(C.POINTER(C.c_int)), ptr(norm_c), ptr(mat_c),
        ptr(rho_c), ptr(phi_c), ptr(mask_c), C.c_size_t(rho_c.size), ptr(out), ptr(cov)))
    return out, cov


class Zernike:
    """Zernike polynomials at arbitrary points, evaluated on the device: the drop-in of ``paos.Zernike``
    (``paos/classes/zernike.py:5-109``; same constructor, ``__call__``, ``j2mn``, ``mn2j``, ``cov``).

    ``Z`` is a masked ``(N,) + rho.shape`` stack (masked where ``rho > 1`` or where ``rho`` itself is masked; the data
    under the mask is 0).  Like the reference, a masked ``rho`` has its mask extended in place."""

    j2mn = staticmethod(j2mn)
    mn2j = staticmethod(mn2j)

    def __init__(self, N, rho, phi, ordering="ansi", normalize=False, device=0):
        assert ordering in ("ansi", "noll", "fringe", "standard"), "Unrecognised ordering scheme."
        assert N > 0, "N shall be a positive integer"
        self.ordering = ordering
        self.N = N
        self.device = device
        self.m, self.n = j2mn(N, ordering)
        self.norm = zernike_norms(self.m, self.n, normalize)
        outside = np.ma.getdata(rho) > 1.0
        if isinstance(rho, np.ma.MaskedArray):
            rho.mask |= outside
            self._mask = np.ma.getmaskarray(rho).copy()
        else:
            self._mask = np.asarray(outside)
        self._rho, self._phi = rho, phi
        data, _ = _points_call(device, self.m, self.n, self.norm, rho, phi, self._mask)
        self.Z = np.ma.MaskedArray(data, mask=np.broadcast_to(self._mask, data.shape).copy(), fill_value=0.0)

    def __call__(self, j=None):
        return self.Z if j is None else self.Z[j]

    def cov(self):
        """``M[i, j] = mean(Z[i] * Z[j])`` over the unmasked points, entries below 1e-10 set to 0 (``zernike.py:293-317``)."""
        _, cov = _points_call(self.device, self.m, self.n, self.norm, self._rho, self._phi, self._mask, want_stack=False,
                              want_cov=True)
        cov[np.abs(cov) < 1e-10] = 0.0
        return cov


class PolyOrthoNorm(Zernike):
    """Polynomials orthonormal on the pupil given by the masks: ``U = inv(chol(cov)) @ Z`` (``zernike.py:320-402``).
    The covariance reduction and the product with the stack run on the device, the ``N x N`` factorisation on the host."""

    def __init__(self, N, rho, phi, ordering="ansi", normalize=False, mask=False, device=0):
        super().__init__(N, rho, phi, ordering=ordering, normalize=normalize, device=device)
        cov = self.cov()
        self.Qt = np.linalg.cholesky(cov)
        self.M = np.linalg.inv(self.Qt)
        self.M[np.abs(self.M) < 1.0e-10] = 0.0
        full_mask = np.ma.getmaskarray(self.Z) | mask
        data, _ = _points_call(device, self.m, self.n, self.norm, self._rho, self._phi, self._mask, mat=self.M)
        self.Z = np.ma.MaskedArray(data=data, mask=full_mask, fill_value=0.0)

    def toZernike(self, coeff):
        """Zernike coefficients ``c = M.T @ coeff`` of a field given in the orthonormal base (``zernike.py:404-420``)."""
        return np.dot(self.M.T, coeff)
