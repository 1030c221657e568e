"""A small FITS reader for the tests (astropy is not installed): primary HDU + IMAGE extensions, any BITPIX."""
import numpy as np

_DT = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}


def read_fits(path):
    """-> list of (header dict, data array or None).  Image arrays are returned as [NAXIS2, NAXIS1] (FITS order)."""
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0, "a FITS file is a whole number of 2880-byte blocks"
    hdus, pos = [], 0
    while pos < len(raw):
        header, done = {}, False
        while not done:
            block = raw[pos:pos + 2880].decode("ascii", "replace")
            pos += 2880
            for i in range(0, 2880, 80):
                card = block[i:i + 80]
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if key == "HIERARCH":
                    name, _, rest = card[9:].partition("=")
                    key, card = name.strip(), "        =" + rest
                if card[8:10] != "= ":
                    continue
                val = card[10:].split(" /")[0].strip() if not card[10:].lstrip().startswith("'") else card[10:]
                if val.lstrip().startswith("'"):
                    s = val.lstrip()[1:]
                    val = s[:s.index("'")].rstrip() if "'" in s else s.rstrip()
                elif val in ("T", "F"):
                    val = val == "T"
                else:
                    try:
                        val = int(val)
                    except ValueError:
                        try:
                            val = float(val)
                        except ValueError:
                            pass
                header[key] = val
        naxis = header.get("NAXIS", 0)
        data = None
        if naxis:
            shape = [header[f"NAXIS{i}"] for i in range(naxis, 0, -1)]
            n = int(np.prod(shape))
            dt = np.dtype(_DT[header["BITPIX"]])
            data = np.frombuffer(raw, dtype=dt, count=n, offset=pos).reshape(shape).astype(dt.newbyteorder("="))
            pos += ((n * dt.itemsize + 2879) // 2880) * 2880
        hdus.append((header, data))
    return hdus


def by_extname(hdus):
    return {h.get("EXTNAME"): d for h, d in hdus if h.get("EXTNAME")}
