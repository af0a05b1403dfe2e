"""astropy.modeling.models.BlackBody().evaluate(x, temperature, scale) — stub, test infrastructure only;
the same sequence of operations as astropy's physical_models.BlackBody.evaluate."""
import numpy as np

from .. import constants as const
from .. import units as u


class BlackBody:
    _native_units = u.Unit("erg cm-2 s-1 Hz-1 sr-1")

    def __init__(self, temperature=5000.0 * u.K, scale=1.0):
        self.temperature = temperature
        self.scale = scale

    def evaluate(self, x, temperature, scale):
        in_temp = temperature if isinstance(temperature, u.Quantity) else u.Quantity(temperature, u.K)
        in_x = x if isinstance(x, u.Quantity) else u.Quantity(x, u.Hz)
        freq = in_x.to(u.Hz, equivalencies=u.spectral())
        temp = in_temp.to(u.K)
        log_boltz = const.h * freq / (const.k_B * temp)
        boltzm1 = np.expm1(log_boltz)
        bb_nu = 2.0 * const.h * freq ** 3 / (const.c ** 2 * boltzm1) / u.sr
        if hasattr(scale, "unit"):
            scale = scale.to(u.dimensionless_unscaled).value
        y = scale * bb_nu.to(self._native_units)
        return y if hasattr(temperature, "unit") else y.value

    def __call__(self, x):
        return self.evaluate(x, self.temperature, self.scale)
