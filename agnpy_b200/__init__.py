"""agnpy_b200: B200-native (sm_100a, FP64) emission-integral hot path of agnpy behind the
reference's own Python API.  See DESIGN.md / INTEGRATION.md."""
from ._lib import strict_reference, strict_reference_mode
from .absorption import Absorption
from .compton import ExternalCompton, SynchrotronSelfCompton
from .ebl import EBL
from .grids import (gamma_e_to_integrate, mu_to_integrate, nu_to_integrate, phi_to_integrate,
                    trapz_loglog, trapz_prefix)
from .spectra import (BrokenPowerLaw, ExpCutoffBrokenPowerLaw, ExpCutoffPowerLaw,
                      InterpolatedDistribution, LogParabola, PowerLaw)
from .synchrotron import ProtonSynchrotron, Synchrotron
from .targets import CMB, Blob, PointSourceBehindJet, RingDustTorus, SphericalShellBLR, SSDisk
from .units import Quantity, u

__all__ = ["Synchrotron", "ProtonSynchrotron", "SynchrotronSelfCompton", "ExternalCompton", "Absorption", "EBL",
           "Blob", "CMB", "PointSourceBehindJet", "SSDisk", "SphericalShellBLR", "RingDustTorus",
           "PowerLaw", "BrokenPowerLaw", "LogParabola", "ExpCutoffPowerLaw", "ExpCutoffBrokenPowerLaw",
           "InterpolatedDistribution",
           "trapz_loglog", "trapz_prefix", "gamma_e_to_integrate", "nu_to_integrate", "mu_to_integrate",
           "phi_to_integrate", "strict_reference", "strict_reference_mode", "Quantity", "u"]
__version__ = "0.2.0"
