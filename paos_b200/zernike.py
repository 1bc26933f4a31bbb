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
