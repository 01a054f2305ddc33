"""Minimal FITS primary-HDU reader (astropy is not a dependency of this build).

Covers what `utils.read_fits` needs (mrcnn/utils.py:1033-1091 of the reference): the primary image
of a SIMPLE FITS file with NAXIS 2 or 4, BITPIX 8/16/32/-32/-64, BSCALE/BZERO/BLANK, returned as a
native-endian numpy array plus an ordered header mapping.

For survey-sized images (the tile driver, mrcnn/sfinder.py) three more entry points avoid touching the whole file:
`read_header_only` (header blocks only: utils.get_fits_header / get_fits_size), `open_primary` (the raw big-endian
image as a read-only numpy.memmap) and `read_subimage` (one [ymin:ymax, xmin:xmax] window of the first plane, scaled
like read_primary). `write_primary` writes a minimal primary HDU (tools and tests; the reference only reads).
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


def read_header_only(path):
    """Header of the primary HDU without reading the data; returns (Header, offset_of_data)."""
    with open(path, "rb") as f:
        buf = f.read(_BLOCK)
        if buf[:6] != b"SIMPLE":
            raise FitsError("%s is not a FITS file" % path)
        while True:
            try:
                return read_header(buf)
            except FitsError:
                more = f.read(_BLOCK)
                if not more:
                    raise
                buf += more


def _shape_and_dtype(hdr):
    naxis = int(hdr.get("NAXIS", 0))
    bitpix = int(hdr["BITPIX"])
    if bitpix not in _DTYPES:
        raise FitsError("unsupported BITPIX %r" % bitpix)
    return tuple(int(hdr["NAXIS%d" % (k + 1)]) for k in range(naxis))[::-1], _DTYPES[bitpix]


def open_primary(path):
    """(memmap, header): the unscaled big-endian image of the primary HDU, mapped read-only (None when NAXIS == 0)."""
    hdr, pos = read_header_only(path)
    if int(hdr.get("NAXIS", 0)) == 0:
        return None, hdr
    shape, dtype = _shape_and_dtype(hdr)
    return np.memmap(path, dtype=dtype, mode="r", offset=pos, shape=shape), hdr


def _scale(raw, hdr):
    """BSCALE / BZERO / BLANK handling of read_primary applied to a raw big-endian array."""
    bitpix = int(hdr["BITPIX"])
    bscale = hdr.get("BSCALE", 1.0)
    bzero = hdr.get("BZERO", 0.0)
    if bitpix > 0:
        blank = hdr.get("BLANK")
        if bscale != 1.0 or bzero != 0.0 or blank is not None:
            data = raw.astype(np.float64) * float(bscale) + float(bzero)
            if blank is not None:
                data[raw == blank] = np.nan
            return data.astype(np.float32) if bitpix <= 16 else data
        return raw.astype(raw.dtype.newbyteorder("="))
    data = raw.astype(raw.dtype.newbyteorder("="))
    if bscale != 1.0 or bzero != 0.0:
        data = data * type(data.flat[0])(bscale) + type(data.flat[0])(bzero)
    return data


def read_subimage(path, xmin, xmax, ymin, ymax):
    """(data[ymin:ymax, xmin:xmax], header) of the first 2-D plane (NAXIS 2, or [0,0] of NAXIS 4) — the window
    utils.read_fits cuts with xmin..ymax (mrcnn/utils.py:1061-1078) — reading only the rows involved."""
    mm, hdr = open_primary(path)
    if mm is None or mm.ndim not in (2, 4):
        raise FitsError("%s: no 2-D / 4-D primary image" % path)
    plane = mm[0, 0] if mm.ndim == 4 else mm
    return _scale(np.array(plane[ymin:ymax, xmin:xmax]), hdr), hdr


def write_primary(path, data, cards=()):
    """Writes `data` (2-D or 4-D; uint8, int16, int32, int64, float32 or float64) as a minimal primary HDU; `cards`
    is an iterable of extra (keyword, value) pairs."""
    data = np.asarray(data)
    bitpix = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}.get(data.dtype.str[1:])
    if bitpix is None or data.ndim not in (2, 4):
        raise FitsError("write_primary: unsupported array %s %s" % (data.dtype, data.shape))

    def card(key, value):
        if isinstance(value, bool):
            text = "%20s" % ("T" if value else "F")
        elif isinstance(value, (int, np.integer)):
            text = "%20d" % value
        elif isinstance(value, (float, np.floating)):
            text = "%20s" % repr(float(value)).upper()
        else:
            text = "'%-8s'" % str(value).replace("'", "''")
        return ("%-8s= %s" % (key, text)).ljust(_CARD)

    lines = [card("SIMPLE", True), card("BITPIX", bitpix), card("NAXIS", data.ndim)]
    lines += [card("NAXIS%d" % (k + 1), n) for k, n in enumerate(data.shape[::-1])]
    lines += [card(k, v) for k, v in cards] + ["END".ljust(_CARD)]
    header = "".join(lines)
    header = header.ljust((len(header) + _BLOCK - 1) // _BLOCK * _BLOCK)
    payload = np.ascontiguousarray(data, dtype=_DTYPES[bitpix]).tobytes()
    payload += b"\0" * ((_BLOCK - len(payload) % _BLOCK) % _BLOCK)
    with open(path, "wb") as f:
        f.write(header.encode("ascii") + payload)

