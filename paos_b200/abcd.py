"""Paraxial ABCD ray-transfer matrix (host scalars; reference ``paos/classes/abcd.py:6-164``).

The matrix is built as translation @ refraction @ magnification and then *decomposed* again on access, so the
properties can differ from the constructor arguments exactly as they do in the reference.
"""
import numpy as np


class ABCD:
    def __init__(self, thickness=0.0, curvature=0.0, n1=1.0, n2=1.0, M=1.0):
        if n1 == 0 or n2 == 0 or M == 0:
            raise ValueError("Refractive index and magnification shall not be zero")
        translate = np.array([[1.0, thickness], [0, 1.0]])
        if n1 == n2:  # thin lens of focal length 1/curvature
            refract = np.array([[1.0, 0.0], [-curvature, 1.0]])
        else:  # dioptre or mirror
            refract = np.array([[1.0, 0.0], [-(1 - n1 / n2) * curvature, n1 / n2]])
        magnify = np.array([[M, 0.0], [0.0, 1.0 / M]])
        self._ABCD = translate @ refract @ magnify
        self._cin = np.sign(n1)
        self._cout = np.sign(n2)

    def _abcd(self):
        return self._ABCD[0, 0], self._ABCD[0, 1], self._ABCD[1, 0], self._ABCD[1, 1]

    @property
    def thickness(self):
        _, b, _, d = self._abcd()
        return b / d

    @property
    def M(self):
        a, b, c, d = self._abcd()
        return (a * d - b * c) / d

    @property
    def n1n2(self):
        return self._ABCD[1, 1] * self.M

    @property
    def power(self):
        return -self._ABCD[1, 0] / self.M

    @property
    def f_eff(self):
        return 1 / (self.power * self.M)

    @property
    def cin(self):
        return self._cin

    @cin.setter
    def cin(self, c):
        self._cin = c

    @property
    def cout(self):
        return self._cout

    @cout.setter
    def cout(self, c):
        self._cout = c

    @property
    def ABCD(self):
        return self._ABCD

    @ABCD.setter
    def ABCD(self, matrix):
        self._ABCD = matrix.copy()

    def __call__(self):
        return self._ABCD

    def __mul__(self, other):
        out = ABCD()
        out.ABCD = self._ABCD @ other()
        out.cin = other.cin
        out.cout = other.cout
        return out
