"""EBL absorption (agnpy/absorption/ebl_absorption.py) on the device.

`EBL(model)` mirrors the reference class: it loads one of the XSPEC-style table models the
reference ships under `agnpy/data/ebl_models/` and `absorption(nu, z)` interpolates it linearly
in (energy, redshift), NaN outside the table.  The interpolation runs in `agnpy_b200_ebl_absorption`;
`apply(sed, nu, z)` multiplies a batch of SEDs [n_models][n_nu] by the table in place (one launch,
per-model redshifts).  astropy.io.fits is replaced by a 40-line reader of the three binary tables.

The table files are data of the reference, not part of this package: they are looked up in
`path` / `data_dir`, then in $AGNPY_EBL_DIR, then in an installed agnpy's `data/ebl_models`."""
import gzip
import importlib.util
import os
from pathlib import Path

import numpy as np

from . import _core
from .units import strip

__all__ = ["EBL", "ebl_files_dict", "read_ebl_table"]

# ebl_absorption.py:12-18
ebl_files_dict = {
    "franceschini_2008": "ebl_franceschini.fits.gz",
    "franceschini_2017": "ebl_franceschini_2017.fits.gz",
    "dominguez_2011": "ebl_dominguez11.fits.gz",
    "finke_2010": "frd_abs.fits.gz",
    "saldana_lopez_2021": "ebl_saldana-lopez_2021.fits.gz",
}


def _bintables(path):
    raw = gzip.open(path, "rb").read() if str(path).endswith(".gz") else Path(path).read_bytes()
    pos, tables = 0, {}
    while pos < len(raw):
        hdr, done = {}, False
        while not done:
            block = raw[pos:pos + 2880]
            pos += 2880
            if len(block) < 2880:
                raise ValueError(f"{path}: truncated FITS header")
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode("ascii", "replace")
                if card[:8].strip() == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    hdr[card[:8].strip()] = card[10:].split("/")[0].strip().strip("'").strip()
        naxis, size = int(hdr.get("NAXIS", "0")), 0
        if naxis:
            size = abs(int(hdr["BITPIX"])) // 8
            for a in range(1, naxis + 1):
                size *= int(hdr[f"NAXIS{a}"])
            size += int(hdr.get("PCOUNT", "0"))
        if hdr.get("XTENSION") == "BINTABLE":
            tables[hdr.get("EXTNAME")] = (hdr, raw[pos:pos + size])
        pos += (size + 2879) // 2880 * 2880
    return tables


def _f32_table(hdr, data, n_cols_expected=None):
    rows, width = int(hdr["NAXIS2"]), int(hdr["NAXIS1"])
    n_fields = int(hdr["TFIELDS"])
    if any(not hdr[f"TFORM{i}"].rstrip().endswith("E") for i in range(1, n_fields + 1)):
        raise ValueError("EBL table: only float32 ('E') columns are supported")
    return np.frombuffer(data[:rows * width], dtype=">f4").reshape(rows, width // 4).astype(np.float32)


def read_ebl_table(path, drop_row=None):
    """(energy_ref [keV], z_ref, values_ref[n_z][n_e]) exactly as the reference holds them
    (float32; ebl_absorption.py:44-57)"""
    t = _bintables(path)
    en = _f32_table(*t["ENERGIES"])
    sp = _f32_table(*t["SPECTRA"])
    energy_ref = np.sqrt(en[:, 0] * en[:, 1])
    z_ref, values_ref = sp[:, 0].copy(), sp[:, 1:].copy()
    if drop_row is not None:
        z_ref, values_ref = np.delete(z_ref, drop_row), np.delete(values_ref, drop_row, axis=0)
    return energy_ref, z_ref, values_ref


def _find(fname, data_dir):
    cands = []
    if data_dir:
        cands.append(Path(data_dir) / fname)
    if os.environ.get("AGNPY_EBL_DIR"):
        cands.append(Path(os.environ["AGNPY_EBL_DIR"]) / fname)
    spec = importlib.util.find_spec("agnpy")
    if spec and spec.origin:
        cands.append(Path(spec.origin).parent / "data" / "ebl_models" / fname)
    for c in cands:
        if c.exists():
            return c
    raise FileNotFoundError(f"EBL table {fname} not found (looked in {[str(c) for c in cands]}); "
                            "pass data_dir= or set AGNPY_EBL_DIR")


class EBL:
    """ebl_absorption.py:21-73"""

    def __init__(self, model="dominguez_2011", data_dir=None, path=None):
        if model not in ebl_files_dict:
            raise ValueError("No EBL model for the reference you specified")
        self.model_name = model
        self.model_file = Path(path) if path else _find(ebl_files_dict[model], data_dir)
        # the franceschini_2008 file repeats the z = 1.001 row (ebl_absorption.py:53-57)
        drop = 1001 if model == "franceschini_2008" else None
        self.energy_ref, self.z_ref, self.values_ref = read_ebl_table(self.model_file, drop)
        self._values_ez = np.ascontiguousarray(self.values_ref.T, dtype=np.float64)

    def absorption(self, nu, z):
        """attenuation exp(-tau_EBL) at the frequencies `nu` for a source at redshift `z`;
        plain ndarray of nu's shape (ebl_absorption.py:66-73)"""
        nu_hz = strip(nu, "Hz", "nu")
        out = _core.ebl_absorption_host(float(z), np.ravel(nu_hz), self.energy_ref, self.z_ref,
                                        self._values_ez)
        return out[0].reshape(np.shape(nu_hz))

    def apply(self, sed, nu, z):
        """multiply a batch of SEDs [n_models][n_nu] (float64, C-contiguous, plain CGS values) by
        the table in place; `z` is a scalar or one redshift per model"""
        nu_hz = strip(nu, "Hz", "nu")
        z = np.broadcast_to(np.asarray(z, dtype=float), (sed.shape[0],))
        return _core.ebl_absorption_host(z, np.ravel(nu_hz), self.energy_ref, self.z_ref,
                                         self._values_ez, sed=sed)
