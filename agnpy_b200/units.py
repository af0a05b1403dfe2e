"""Quantity boundary of the drop-in API.

The reference API takes and returns `astropy.units.Quantity` objects.  When astropy is
installed, those are used as they are (`to_value` strips them, results are wrapped into
astropy Quantities).  astropy is not installable in the build / GPU image, so this module
also carries a small CGS-only stand-in (`Quantity`, `Unit`, the `u` namespace) that
implements the subset of the protocol the hot path touches: `.to_value(unit)`, `.to(unit)`,
`.unit`, `.value`, scalar/array arithmetic.  Unit stripping mirrors
agnpy/utils/conversion.py:30-46: everything ends up as CGS (Gaussian) doubles.
"""
import functools
import re
from fractions import Fraction

import numpy as np

try:  # pragma: no cover - astropy is absent in the build image
    import astropy as _astropy
    if getattr(_astropy, "IS_AGNPY_B200_STUB", False):   # the test suite's stand-in is not astropy
        raise ImportError("astropy stand-in of the test infrastructure")
    import astropy.units as _astropy_units
    HAVE_ASTROPY = True
except Exception:  # ImportError or a stub without the API
    _astropy_units = None
    HAVE_ASTROPY = False

_BASES = ("cm", "g", "s", "K")


def _dim(**kw):
    return tuple(Fraction(kw.get(b, 0)) for b in _BASES)


# name -> (scale to CGS, dimension)
_UNITS = {
    "": (1.0, _dim()),
    "cm": (1.0, _dim(cm=1)), "m": (1e2, _dim(cm=1)), "km": (1e5, _dim(cm=1)),
    "pc": (3.0856775814913673e18, _dim(cm=1)), "kpc": (3.0856775814913673e21, _dim(cm=1)),
    "Mpc": (3.0856775814913673e24, _dim(cm=1)), "Gpc": (3.0856775814913673e27, _dim(cm=1)),
    "AA": (1e-8, _dim(cm=1)), "Angstrom": (1e-8, _dim(cm=1)),
    "g": (1.0, _dim(g=1)), "kg": (1e3, _dim(g=1)), "M_sun": (1.988409870698051e33, _dim(g=1)),
    "s": (1.0, _dim(s=1)), "d": (86400.0, _dim(s=1)), "h": (3600.0, _dim(s=1)),
    "yr": (31557600.0, _dim(s=1)),
    "Hz": (1.0, _dim(s=-1)), "GHz": (1e9, _dim(s=-1)),
    "erg": (1.0, _dim(g=1, cm=2, s=-2)), "J": (1e7, _dim(g=1, cm=2, s=-2)),
    "eV": (1.602176634e-12, _dim(g=1, cm=2, s=-2)), "keV": (1.602176634e-9, _dim(g=1, cm=2, s=-2)),
    "MeV": (1.602176634e-6, _dim(g=1, cm=2, s=-2)), "GeV": (1.602176634e-3, _dim(g=1, cm=2, s=-2)),
    "TeV": (1.602176634, _dim(g=1, cm=2, s=-2)),
    "W": (1e7, _dim(g=1, cm=2, s=-3)),
    # Gauss in Gaussian-cgs base units, agnpy/utils/conversion.py:11-12
    "G": (1.0, _dim(cm=Fraction(-1, 2), g=Fraction(1, 2), s=-1)),
    "K": (1.0, _dim(K=1)),
}


_TOKEN = re.compile(r"\s*(/|\*|1(?![0-9])|[A-Za-z_]+)\s*(\(\s*[-+]?\d+(?:/\d+)?\s*\)|[-+]?\d+)?")


@functools.lru_cache(maxsize=512)
def _parse(text):
    """'erg cm-2 s-1', 'erg/s', 'g / s', 'cm(-1/2) g(1/2) s-1' -> (scale, dimension)"""
    text = text.strip()
    scale, dim = 1.0, list(_dim())
    sign, pos = 1, 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise ValueError(f"cannot parse unit {text!r}")
        pos = m.end()
        name, p = m.group(1), m.group(2)
        if name == "/":
            sign = -1
            continue
        if name in ("*", "1"):
            continue
        if name not in _UNITS:
            raise ValueError(f"unknown unit {name!r} in {text!r}")
        power = Fraction(p.strip("() ")) if p else Fraction(1)
        s, d = _UNITS[name]
        power = sign * power
        scale *= s ** float(power)
        dim = [a + power * b for a, b in zip(dim, d)]
    return scale, tuple(dim)


class UnitConversionError(ValueError):
    pass


class Unit:
    __array_ufunc__ = None  # numpy scalars / arrays defer to __rmul__ / __rtruediv__

    def __init__(self, text="", _parsed=None):
        if isinstance(text, Unit):
            text, _parsed = text.text, (text.scale, text.dim)
        self.text = text
        self.scale, self.dim = _parsed if _parsed is not None else _parse(text)

    def __repr__(self):
        return f'Unit("{self.text}")'

    def __str__(self):
        return self.text

    def __eq__(self, other):
        other = Unit(other) if isinstance(other, str) else other
        return isinstance(other, Unit) and self.dim == other.dim and np.isclose(self.scale, other.scale, rtol=1e-14)

    def __hash__(self):
        return hash(self.dim)

    def _combine(self, other, sign):
        d = tuple(a + sign * b for a, b in zip(self.dim, other.dim))
        s = self.scale * other.scale ** sign
        op = " " if sign > 0 else " / "
        return Unit(f"{self.text}{op}({other.text})" if other.text else self.text, (s, d))

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, 1)
        return Quantity(other, self)

    __rmul__ = lambda self, other: Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit("")._combine(self, -1))

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(12)
        return Unit(f"({self.text})^{p}", (self.scale ** float(p), tuple(a * p for a in self.dim)))

    def factor_to(self, other):
        if isinstance(other, str):
            if other == self.text:
                return 1.0
            o_scale, o_dim = _parse(other)
            if self.dim != o_dim:
                raise UnitConversionError(f"'{self.text}' and '{other}' are not convertible")
            return self.scale / o_scale
        if self.dim != other.dim:
            raise UnitConversionError(f"'{self.text}' and '{other.text}' are not convertible")
        return self.scale / other.scale


class Quantity:
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, value, unit=""):
        if isinstance(value, Quantity):
            unit = value.unit._combine(Unit(unit), 1) if unit else value.unit
            value = value.value
        nd = getattr(value, "ndim", None)
        if nd is None:
            nd = np.ndim(value)
        self.value = np.asarray(value, dtype=np.float64) if nd else float(value)
        self.unit = unit if isinstance(unit, Unit) else Unit(unit)

    # ---- the protocol the hot path uses ---------------------------------------------------
    def to_value(self, unit=None):
        if unit is None:
            return self.value
        f = self.unit.factor_to(unit)
        return self.value if f == 1.0 else self.value * f

    def to(self, unit, equivalencies=None):
        unit = unit if isinstance(unit, Unit) else Unit(unit)
        return Quantity(self.to_value(unit), unit)

    @property
    def cgs(self):
        return Quantity(self.value * self.unit.scale, Unit("cgs", (1.0, self.unit.dim)))

    # ---- light arithmetic --------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value * other.value, self.unit._combine(other.unit, 1))
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit._combine(other, 1))
        return Quantity(self.value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value / other.value, self.unit._combine(other.unit, -1))
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit._combine(other, -1))
        return Quantity(self.value / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / self.value, Unit("")._combine(self.unit, -1))

    def __add__(self, other):
        return Quantity(self.value + Quantity(other).to_value(self.unit), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.value - Quantity(other).to_value(self.unit), self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def __getitem__(self, idx):
        return Quantity(self.value[idx], self.unit)

    def __len__(self):
        return len(self.value)

    def __iter__(self):
        return (Quantity(v, self.unit) for v in self.value)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __float__(self):
        return float(self.value)

    @property
    def shape(self):
        return np.shape(self.value)

    @property
    def size(self):
        return np.size(self.value)

    def max(self):
        return Quantity(np.max(self.value), self.unit)

    def argmax(self):
        return int(np.argmax(self.value))

    def __repr__(self):
        return f"<Quantity {self.value!r} {self.unit.text}>"


class _Namespace:
    """`u.cm`, `u.Hz`, `u.Unit("erg s-1")`, `u.Quantity`"""
    Unit = Unit
    Quantity = Quantity

    def __getattr__(self, name):
        if name in _UNITS:
            return Unit(name)
        raise AttributeError(name)


u = _astropy_units if HAVE_ASTROPY else _Namespace()


# ---- stripping / wrapping used by the evaluate_* boundary ---------------------------------------
def strip(x, unit, name="argument"):
    """Quantity -> float64 value(s) in `unit`.  Plain numbers are taken to be already in
    that (CGS) unit.  Wrong physical dimensions raise."""
    if hasattr(x, "to_value"):
        try:
            v = x.to_value(unit)
        except Exception as exc:
            raise TypeError(f"{name}: cannot convert {getattr(x, 'unit', '?')} to '{unit}'") from exc
        nd = getattr(v, "ndim", None)  # np.ndim() is the slow way to ask an ndarray or a float
        if nd is None:
            nd = np.ndim(v)
        return np.asarray(v, dtype=np.float64) if nd else float(v)
    if isinstance(x, (int, float, np.floating, np.integer)):
        return float(x)
    if isinstance(x, (np.ndarray, list, tuple)):
        return np.asarray(x, dtype=np.float64)
    raise TypeError(f"{name}: expected a Quantity [{unit}] or a number, got {type(x).__name__}")


def wrap(values, unit):
    """float64 array -> Quantity (astropy's when available)"""
    if HAVE_ASTROPY:
        return values * _astropy_units.Unit(unit)
    return Quantity(values, unit)
