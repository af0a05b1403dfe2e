"""astropy.constants (>= 7: CODATA 2018 / IAU 2015), SI values as astropy defines them — stub, test
infrastructure only.  `e.gauss` is what agnpy/synchrotron/synchrotron.py:12 uses."""
import numpy as np

from .. import units as u


class Constant(u.Quantity):
    def __new__(cls, name, value, unit):
        obj = u.Quantity.__new__(cls, value, unit)
        obj.name = name
        return obj

    def __array_finalize__(self, obj):
        super().__array_finalize__(obj)
        self.name = getattr(obj, "name", None)

    # arithmetic on a constant gives a plain Quantity, as in astropy
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(x.view(u.Quantity) if isinstance(x, Constant) else x for x in inputs)
        return u.Quantity.__array_ufunc__(inputs[0] if isinstance(inputs[0], u.Quantity) else
                                          next(x for x in inputs if isinstance(x, u.Quantity)),
                                          ufunc, method, *inputs, **kwargs)

    def __eq__(self, other):
        if isinstance(other, Constant):
            return self.name == other.name
        return super().__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def to(self, *a, **k):
        return self.view(u.Quantity).to(*a, **k)

    @property
    def cgs(self):
        return self.view(u.Quantity).cgs

    @property
    def si(self):
        return self.view(u.Quantity).si


class EMConstant(Constant):
    @property
    def gauss(self):
        # astropy: e.gauss = 4.803204712570263e-10 Fr
        return u.Quantity(4.803204712570263e-10, u.Fr)

    @property
    def esu(self):
        return self.gauss

    @property
    def cgs(self):
        raise TypeError("Constant 'e' does not have a single CGS representation; use e.gauss / e.esu")


c = Constant("c", 299792458.0, "m / s")
h = Constant("h", 6.62607015e-34, "J s")
hbar = Constant("hbar", 6.62607015e-34 / (2 * np.pi), "J s")
k_B = Constant("k_B", 1.380649e-23, "J / K")
G = Constant("G", 6.6743e-11, "m3 / (kg s2)")
m_e = Constant("m_e", 9.1093837015e-31, "kg")
m_p = Constant("m_p", 1.67262192369e-27, "kg")
m_n = Constant("m_n", 1.67492749804e-27, "kg")
e = EMConstant("e", 1.602176634e-19, "C")
sigma_T = Constant("sigma_T", 6.6524587321e-29, "m2")
sigma_sb = Constant("sigma_sb", 5.670374419e-8, "W / (K4 m2)")
mu0 = Constant("mu0", 1.25663706212e-6, "N / A2")
eps0 = Constant("eps0", 8.8541878128e-12, "F / m")
M_sun = Constant("M_sun", 1.988409870698051e30, "kg")
L_sun = Constant("L_sun", 3.828e26, "W")
R_sun = Constant("R_sun", 695700000.0, "m")
pc = Constant("pc", 3.0856775814913673e16, "m")
au = Constant("au", 1.49597870700e11, "m")
