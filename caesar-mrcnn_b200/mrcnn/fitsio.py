"""Minimal FITS primary-HDU reader (astropy is not a dependency of this build).

Covers what `utils.read_fits` needs (mrcnn/utils.py:1033-1091 of the reference): the primary image
of a SIMPLE FITS file with NAXIS 2 or 4, BITPIX 8/16/32/-32/-64, BSCALE/BZERO/BLANK, returned as a
native-endian numpy array plus an ordered header mapping.
"""
import collections

import numpy as np

_BLOCK = 2880
_CARD = 80
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


class FitsError(IOError):
    pass


class Header(collections.OrderedDict):
    """Ordered keyword -> value mapping; COMMENT/HISTORY cards are collected in lists."""
    pass


def _parse_value(text):
    text = text.strip()
    if not text:
        return None
    if text[0] == "'":
        # quoted string, '' is an escaped quote; trailing blanks are not significant
        out, i = [], 1
        while i < len(text):
            ch = text[i]
            if ch == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(ch)
            i += 1
        return "".join(out).rstrip()
    token = text.split("/", 1)[0].strip()
    if token in ("T", "F"):
        return token == "T"
    try:
        return int(token)
    except ValueError:
        pass
    try:
        return float(token.replace("D", "E").replace("d", "e"))
    except ValueError:
        return token


def read_header(buf, offset=0):
    """Parses header blocks starting at `offset`; returns (Header, offset_of_data)."""
    hdr = Header()
    pos = offset
    while True:
        block = buf[pos:pos + _BLOCK]
        if len(block) < _BLOCK:
            raise FitsError("truncated FITS header")
        pos += _BLOCK
        for i in range(0, _BLOCK, _CARD):
            card = block[i:i + _CARD].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return hdr, pos
            if not key:
                continue
            if key in ("COMMENT", "HISTORY"):
                hdr.setdefault(key, []).append(card[8:].rstrip())
            elif card[8:10] == "= ":
                hdr[key] = _parse_value(card[10:])


def read_primary(path):
    """Returns (data, header) of the primary HDU; data is None when NAXIS == 0."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:6] != b"SIMPLE":
        raise FitsError("%s is not a FITS file" % path)
    hdr, pos = read_header(buf)
    naxis = int(hdr.get("NAXIS", 0))
    if naxis == 0:
        return None, hdr
    bitpix = int(hdr["BITPIX"])
    if bitpix not in _DTYPES:
        raise FitsError("unsupported BITPIX %r" % bitpix)
    shape = tuple(int(hdr["NAXIS%d" % (k + 1)]) for k in range(naxis))[::-1]
    count = int(np.prod(shape))
    raw = np.frombuffer(buf, dtype=_DTYPES[bitpix], count=count, offset=pos).reshape(shape)
    bscale = hdr.get("BSCALE", 1.0)
    bzero = hdr.get("BZERO", 0.0)
    if bitpix > 0:
        blank = hdr.get("BLANK")
        if bscale != 1.0 or bzero != 0.0 or blank is not None:
            data = raw.astype(np.float64) * float(bscale) + float(bzero)
            if blank is not None:
                data[raw == blank] = np.nan
            data = data.astype(np.float32) if bitpix <= 16 else data
        else:
            data = raw.astype(raw.dtype.newbyteorder("="))
    else:
        data = raw.astype(raw.dtype.newbyteorder("="))
        if bscale != 1.0 or bzero != 0.0:
            data = data * type(data.flat[0])(bscale) + type(data.flat[0])(bzero)
    return data, hdr
