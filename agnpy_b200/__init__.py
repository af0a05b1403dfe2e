"""agnpy_b200: B200-native (sm_100a, FP64) emission-integral hot path of agnpy behind the
reference's own Python API.  See DESIGN.md / INTEGRATION.md."""
from .absorption import Absorption
from .compton import ExternalCompton, SynchrotronSelfCompton
from .ebl import EBL
from .grids import (gamma_e_to_integrate, mu_to_integrate, nu_to_integrate, phi_to_integrate,
                    trapz_fp32, trapz_loglog, trapz_prefix)
from .synchrotron import ProtonSynchrotron, Synchrotron
from .units import Quantity, u

__all__ = ["Synchrotron", "ProtonSynchrotron", "SynchrotronSelfCompton", "ExternalCompton", "Absorption", "EBL",
           "trapz_loglog", "trapz_prefix", "trapz_fp32", "gamma_e_to_integrate", "nu_to_integrate", "mu_to_integrate",
           "phi_to_integrate", "Quantity", "u"]
__version__ = "0.1.0"
