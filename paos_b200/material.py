"""Refractive index of the optical glasses a lens file may name (host scalars, once per wavelength).

Restates ``paos/util/material.py:36-167`` (Sellmeier-1 dispersion relative to air at the glass reference
temperature, the Kohlrausch air index, the ``D0`` thermal model) and the published catalogue constants of
``paos/util/lib.py``.  Used only by :func:`paos_b200.parse_config.parse_config`.
"""
import numpy as np

#            K1              L1              K2              L2              K3              L3              D0          Tref
_CATALOGUE = {
    "CAF2":     (5.67588800e-1, 2.52643000e-3, 4.71091400e-1, 1.00783330e-2, 3.84847230e0, 1.20055600e3, -2.6600e-5, 20.0),
    "SAPPHIRE": (1.023798000e0, 3.775880000e-3, 1.058264000e0, 1.225440000e-2, 5.280792000e0, 3.213616000e2, 1.8000e-5, 20.0),
    "ZNSE":     (4.29801490e0, 3.68881960e-2, 6.27765570e-1, 1.43476258e-1, 2.89556330e0, 2.20849196e3, 5.5400e-5, 20.0),
    "BK7":      (1.03961212e0, 6.00069867e-3, 2.31792344e-1, 2.00179144e-2, 1.01046945e0, 1.03560653e2, 1.8600e-6, 20.0),
    "SF6":      (1.724484820e0, 1.348719470e-2, 3.901048890e-1, 5.693180950e-2, 1.045728580e0, 1.185571850e2, 6.6900e-6, 20.0),
    "SF11":     (1.73848403e0, 1.36068604e-2, 3.11168974e-1, 6.15960463e-2, 1.17490871e0, 1.21922711e2, 1.1200e-5, 20.0),
    "BAF2":     (6.43356000e-1, 3.34000000e-3, 5.06762000e-1, 1.20300000e-2, 3.82610000e0, 2.15169810e3, -4.4600e-5, 20.0),
}

materials = {
    name: {"Tref": c[7], "sellmeier": dict(zip(("K1", "L1", "K2", "L2", "K3", "L3"), c[:6])), "Tmodel": {"D0": c[6]}}
    for name, c in _CATALOGUE.items()
}


class Material:
    """Glass library evaluated at wavelength(s) ``wl`` [micron], ambient temperature [C] and pressure [atm]."""

    def __init__(self, wl, Tambient=-218.0, Pambient=1.0, materials=None):
        self.wl = wl
        self.Tambient = Tambient
        self.Pambient = Pambient
        self.materials = globals()["materials"] if materials is None else materials

    def sellmeier(self, par):
        wl2 = self.wl**2
        acc = par["K1"] * wl2 / (wl2 - par["L1"])
        acc += par["K2"] * wl2 / (wl2 - par["L2"])
        acc += par["K3"] * wl2 / (wl2 - par["L3"])
        return np.sqrt(acc + 1.0)

    @staticmethod
    def nT(n, D0, delta_T):
        return n + (n**2 - 1.0) / (2.0 * n) * D0 * delta_T

    def nair(self, T, P=1.0):
        wl2 = self.wl**2
        nref = 1.0 + 1.0e-8 * (6432.8 + 2949810.0 * wl2 / (146.0 * wl2 - 1.0) + 25540.0 * wl2 / (41.0 * wl2 - 1.0))
        return 1.0 + (nref - 1.0) * P / (1.0 + 3.4785e-3 * (T - 15))

    def nmat(self, name):
        """(index at the glass reference temperature, index at ``Tambient``), both relative to ambient air."""
        glass = self.materials[name.upper()]
        n_ref = self.sellmeier(glass["sellmeier"]) * self.nair(T=glass["Tref"], P=self.Pambient)
        return n_ref, self.nT(n_ref, glass["Tmodel"]["D0"], self.Tambient - glass["Tref"])
