"""astropy.coordinates.Distance(z=..., cosmology=...) — stub, test infrastructure only."""
from .. import units as u
from ..cosmology import default_cosmology


class Distance(u.Quantity):
    def __new__(cls, value=None, unit=None, z=None, cosmology=None, **kw):
        if z is not None:
            if value is not None:
                raise ValueError("give either a value or a redshift")
            cosmo = default_cosmology if cosmology is None else cosmology
            q = cosmo.luminosity_distance(z)
            if unit is not None:
                q = q.to(unit)
        else:
            q = u.Quantity(value, unit)
        obj = q.view(cls)
        obj._unit = q.unit
        return obj

    @property
    def z(self):
        """redshift at which the default cosmology has this luminosity distance (astropy: z_at_value)"""
        import numpy as np
        from scipy.optimize import brentq

        target = float(self.to_value(u.Mpc))
        f = lambda zz: float(default_cosmology.luminosity_distance(zz).value) - target
        return brentq(f, 1e-10, 1e3, xtol=1e-14, rtol=1e-13)
