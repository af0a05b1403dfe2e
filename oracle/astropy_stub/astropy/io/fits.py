"""The part of astropy.io.fits the reference's EBL reader uses (binary tables of an XSPEC-style
table model, `agnpy/absorption/ebl_absorption.py:44-56`) — stub, test infrastructure only."""
import gzip

import numpy as np

_TFORM = {"E": ">f4", "D": ">f8", "J": ">i4", "I": ">i2", "K": ">i8", "B": "u1", "L": "u1"}


class _HDU:
    def __init__(self, header, data):
        self.header, self.data, self.name = header, data, header.get("EXTNAME", "PRIMARY")


class HDUList(list):
    def __getitem__(self, key):
        if isinstance(key, str):
            for h in self:
                if h.name == key:
                    return h
            raise KeyError(key)
        return list.__getitem__(self, key)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _header(buf, pos):
    hdr = {}
    while True:
        block = buf[pos:pos + 2880]
        pos += 2880
        done = False
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:].split("/")[0].strip() if not card[10:].lstrip().startswith("'") else \
                card[10:].strip()[1:].split("'")[0].strip()
            try:
                val = int(val)
            except ValueError:
                try:
                    val = float(val)
                except ValueError:
                    val = {"T": True, "F": False}.get(val, val)
            hdr[key] = val
        if done:
            return hdr, pos


def open(path, *a, **k):
    opener = gzip.open if str(path).endswith(".gz") else __builtins__["open"]
    with opener(path, "rb") as f:
        buf = f.read()
    hdus, pos = HDUList(), 0
    while pos < len(buf):
        hdr, pos = _header(buf, pos)
        naxis = hdr.get("NAXIS", 0)
        size = 0
        if naxis:
            size = abs(hdr.get("BITPIX", 8)) // 8
            for i in range(1, naxis + 1):
                size *= hdr[f"NAXIS{i}"]
            size += hdr.get("PCOUNT", 0)
        data = None
        if hdr.get("XTENSION", "").strip() == "BINTABLE":
            fields = []
            for i in range(1, hdr["TFIELDS"] + 1):
                tform = str(hdr[f"TFORM{i}"]).strip()
                if "P" in tform:        # variable-length column: (count, heap offset) descriptor
                    fields.append((str(hdr[f"TTYPE{i}"]).strip(), ">i4", (2,)))
                    continue
                count = int(tform[:-1]) if tform[:-1] else 1
                code = tform[-1]
                dt = f"S{count}" if code == "A" else _TFORM[code]
                fields.append((str(hdr[f"TTYPE{i}"]).strip(), dt) if code == "A" or count == 1
                              else (str(hdr[f"TTYPE{i}"]).strip(), dt, (count,)))
            rec = np.frombuffer(buf, dtype=np.dtype(fields), count=hdr["NAXIS2"], offset=pos)
            data = {n: np.ascontiguousarray(rec[n]).astype(rec[n].dtype.newbyteorder("="))
                    for n in rec.dtype.names}
        hdus.append(_HDU(hdr, data))
        pos += (size + 2879) // 2880 * 2880
    return hdus
