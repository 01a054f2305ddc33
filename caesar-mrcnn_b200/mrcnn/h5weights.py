"""Keras-HDF5 weight files without h5py / libhdf5 (neither exists in this build environment).

`read_keras_weights(path)` is what `MaskRCNN.load_weights` (reference: mrcnn/model.py:2197-2239 ->
keras.engine.saving.load_weights_from_hdf5_group_by_name) needs: {layer_name: [arrays in
layer.weights order]}.  `write_keras_weights(path, weights)` writes the same layout Keras 2.2.4 +
h5py <= 2.10 produce with default settings: superblock v0, version-1 object
headers, old-style groups (local heap + v1 B-tree + symbol-table nodes), attributes as header
messages, contiguous little-endian float32 datasets.

Reader coverage: superblock v0/v1; v1 object headers with continuation blocks; symbol-table groups;
attribute messages v1-v3 with fixed-length or variable-length strings and numeric scalars/arrays;
contiguous and compact datasets; chunked datasets without filters or with deflate/shuffle.
Validated on self-written files only (the real 256 MB file is an unresolved Git-LFS pointer).
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


# --------------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------------

class _Datatype(object):
    def __init__(self, cls, size, np_dtype=None, vlen_string=False, strpad=0, base=None):
        self.cls, self.size, self.np_dtype, self.vlen_string, self.strpad, self.base = cls, size, np_dtype, vlen_string, strpad, base


class H5File(object):
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        base = -1
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise H5Error("%s: not an HDF5 file" % path)
        b = self.buf
        ver = b[base + 8]
        if ver not in (0, 1):
            raise H5Error("HDF5 superblock version %d is not supported (only 0/1, as written by h5py<=2.10 defaults)" % ver)
        self.so, self.sl = b[base + 13], b[base + 14]
        if self.so != 8 or self.sl != 8:
            raise H5Error("only 8-byte offsets/lengths are supported")
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_addr = self._u(p, 8)
        p += 8 * 4
        # root group symbol table entry
        self.root = self._read_symbol_entry(p)

    # -- primitives -----------------------------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _read_symbol_entry(self, p):
        name_off, ohdr, cache = self._u(p, 8), self._u(p + 8, 8), self._u(p + 16, 4)
        btree = heap = None
        if cache == 1:
            btree, heap = self._u(p + 24, 8), self._u(p + 32, 8)
        return {"name_off": name_off, "ohdr": ohdr, "btree": btree, "heap": heap}

    # -- object headers -------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, data_bytes)] of a version-1 object header (following continuations)."""
        b = self.buf
        addr += self.base_addr
        if b[addr] != 1:
            if b[addr:addr + 4] == b"OHDR":
                raise H5Error("version-2 object headers (libver='latest' files) are not supported")
            raise H5Error("bad object header version %d" % b[addr])
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    blocks.append((self._u_bytes(data, 0, 8) + self.base_addr, self._u_bytes(data, 8, 8)))
                out.append((mtype, flags, data))
        return out

    @staticmethod
    def _u_bytes(data, off, n):
        return int.from_bytes(data[off:off + n], "little")

    # -- groups -----------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        h = heap_addr + self.base_addr
        if self.buf[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = self._u(h + 24, 8) + self.base_addr
        end = self.buf.index(b"\0", data + off)
        return self.buf[data + off:end].decode("utf-8")

    def _walk_btree(self, addr, heap, out):
        a = addr + self.base_addr
        if self.buf[a:a + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        level, used = self.buf[a + 5], self._u(a + 6, 2)
        p = a + 8 + 16
        for i in range(used):
            child = self._u(p + 8, 8)          # key (8) then child (8)
            p += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                s = child + self.base_addr
                if self.buf[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol table node")
                n = self._u(s + 6, 2)
                q = s + 8
                for _ in range(n):
                    ent = self._read_symbol_entry(q)
                    out[self._heap_name(heap, ent["name_off"])] = ent["ohdr"]
                    q += 40

    def group_links(self, ohdr):
        """{name: object header address} of a group."""
        links = {}
        for mtype, _, data in self.messages(ohdr):
            if mtype == 0x11:      # symbol table message
                self._walk_btree(self._u_bytes(data, 0, 8), self._u_bytes(data, 8, 8), links)
            elif mtype == 0x06:    # link message (compact new-style group)
                ver, fl = data[0], data[1]
                p = 2
                ltype = 0
                if fl & 0x08:
                    ltype = data[p]
                    p += 1
                if fl & 0x04:
                    p += 8
                if fl & 0x10:
                    p += 1
                lsz = 1 << (fl & 3)
                nlen = self._u_bytes(data, p, lsz)
                p += lsz
                name = data[p:p + nlen].decode("utf-8")
                p += nlen
                if ltype == 0:
                    links[name] = self._u_bytes(data, p, 8)
        return links

    def is_group(self, ohdr):
        return any(m[0] in (0x11, 0x02, 0x06) for m in self.messages(ohdr))

    # -- datatypes / dataspaces ------------------------------------------------------------------
    def _parse_datatype(self, data, p=0):
        cv = data[p]
        cls, bits0 = cv & 0x0F, data[p + 1]
        size = self._u_bytes(data, p + 4, 4)
        if cls == 0:     # fixed point
            end = ">" if bits0 & 1 else "<"
            signed = "i" if bits0 & 8 else "u"
            return _Datatype(cls, size, np.dtype("%s%s%d" % (end, signed, size))), p + 8 + 4
        if cls == 1:     # float
            end = ">" if bits0 & 1 else "<"
            return _Datatype(cls, size, np.dtype("%sf%d" % (end, size))), p + 8 + 12
        if cls == 3:     # fixed-length string
            return _Datatype(cls, size, np.dtype("S%d" % size), strpad=bits0 & 0x0F), p + 8
        if cls == 9:     # variable length
            is_str = (bits0 & 0x0F) == 1
            base, q = self._parse_datatype(data, p + 8)
            return _Datatype(cls, size, None, vlen_string=is_str, base=base), q
        raise H5Error("unsupported HDF5 datatype class %d" % cls)

    def _parse_dataspace(self, data):
        ver, rank, flags = data[0], data[1], data[2]
        p = 8 if ver == 1 else 4
        if ver == 2 and data[3] == 2:
            return None      # null dataspace
        return tuple(self._u_bytes(data, p + 8 * i, 8) for i in range(rank))

    def _global_heap_object(self, addr, index):
        a = addr + self.base_addr
        if self.buf[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = self._u(a + 8, 8)
        p, end = a + 16, a + size
        while p + 16 <= end:
            idx, osz = self._u(p, 2), self._u(p + 8, 8)
            if idx == 0:
                break
            if idx == index:
                return self.buf[p + 16:p + 16 + osz]
            p += 16 + ((osz + 7) // 8) * 8
        raise H5Error("global heap object %d not found" % index)

    def _decode(self, raw, dt, shape):
        count = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            items = []
            for i in range(count):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                obj = self._global_heap_object(addr, idx)[:ln * (1 if dt.vlen_string else dt.base.size)]
                items.append(obj if dt.vlen_string else np.frombuffer(obj, dtype=dt.base.np_dtype))
            arr = np.array(items, dtype=object)
            return arr.reshape(shape) if shape else arr[0]
        arr = np.frombuffer(raw, dtype=dt.np_dtype, count=count)
        if dt.cls != 3:
            arr = arr.astype(dt.np_dtype.newbyteorder("="))
        return arr.reshape(shape) if shape else arr[0]

    # -- attributes -------------------------------------------------------------------------------
    def attributes(self, ohdr):
        out = {}
        for mtype, _, data in self.messages(ohdr):
            if mtype != 0x0C:
                continue
            ver = data[0]
            nsz, dtsz, dssz = self._u_bytes(data, 2, 2), self._u_bytes(data, 4, 2), self._u_bytes(data, 6, 2)
            p = 8 + (1 if ver == 3 else 0)
            pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            name = data[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += pad(nsz)
            dt, _ = self._parse_datatype(data, p)
            p += pad(dtsz)
            shape = self._parse_dataspace(data[p:p + dssz])
            p += pad(dssz)
            if shape is None:
                out[name] = None
                continue
            out[name] = self._decode(data[p:], dt, shape)
        return out

    # -- datasets ---------------------------------------------------------------------------------
    def read_dataset(self, ohdr):
        dt = shape = layout = None
        filters = []
        for mtype, _, data in self.messages(ohdr):
            if mtype == 0x03:
                dt, _ = self._parse_datatype(data)
            elif mtype == 0x01:
                shape = self._parse_dataspace(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._parse_filters(data)
        if dt is None or shape is None or layout is None:
            raise H5Error("object is not a dataset")
        nbytes = int(np.prod(shape)) * dt.size if shape else dt.size
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 0:
                sz = self._u_bytes(layout, 2, 2)
                raw = layout[4:4 + sz]
            elif cls == 1:
                addr = self._u_bytes(layout, 2, 8)
                raw = b"\0" * nbytes if addr == UNDEF else self.buf[addr + self.base_addr:addr + self.base_addr + nbytes]
            elif cls == 2:
                ndim = layout[2]
                btree = self._u_bytes(layout, 3, 8)
                cdims = [self._u_bytes(layout, 11 + 4 * i, 4) for i in range(ndim)]
                return self._read_chunked(btree, cdims[:-1], shape, dt, filters)
            else:
                raise H5Error("unsupported layout class %d" % cls)
        elif ver in (1, 2):
            ndim, cls = layout[1], layout[2]
            p = 8
            if cls == 1:
                addr = self._u_bytes(layout, p, 8)
                raw = self.buf[addr + self.base_addr:addr + self.base_addr + nbytes]
            elif cls == 2:
                btree = self._u_bytes(layout, p, 8)
                cdims = [self._u_bytes(layout, p + 8 + 4 * i, 4) for i in range(ndim)]
                return self._read_chunked(btree, cdims[:-1], shape, dt, filters)
            else:
                raise H5Error("unsupported v%d layout class %d" % (ver, cls))
        else:
            raise H5Error("unsupported data layout version %d" % ver)
        return self._decode(raw, dt, shape)

    def _parse_filters(self, data):
        ver, n = data[0], data[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = self._u_bytes(data, p, 2)
            if ver == 1 or fid >= 256:
                nlen = self._u_bytes(data, p + 2, 2)
                p += 4
            else:
                nlen = 0
                p += 2
            ncv = self._u_bytes(data, p + 2, 2)
            p += 4
            if ver == 1:
                nlen = (nlen + 7) // 8 * 8
            p += nlen
            cvals = [self._u_bytes(data, p + 4 * i, 4) for i in range(ncv)]
            p += 4 * ncv
            if ver == 1 and ncv % 2:
                p += 4
            out.append((fid, cvals))
        return out

    def _read_chunked(self, btree, cdims, shape, dt, filters):
        out = np.zeros(shape, dtype=dt.np_dtype.newbyteorder("="))
        rank = len(shape)

        def walk(addr):
            a = addr + self.base_addr
            if self.buf[a:a + 4] != b"TREE":
                raise H5Error("bad chunk B-tree node")
            level, used = self.buf[a + 5], self._u(a + 6, 2)
            p = a + 24
            ksz = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, fmask = self._u(p, 4), self._u(p + 4, 4)
                offs = [self._u(p + 8 + 8 * i, 8) for i in range(rank)]
                child = self._u(p + ksz, 8)
                p += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = self.buf[child + self.base_addr:child + self.base_addr + csize]
                for k, (fid, cv) in reversed(list(enumerate(filters))):
                    if fmask & (1 << k):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cv[0] if cv else dt.size
                        arr = np.frombuffer(raw, dtype=np.uint8)
                        raw = arr.reshape(es, -1).T.tobytes()
                    else:
                        raise H5Error("unsupported HDF5 filter %d" % fid)
                chunk = np.frombuffer(raw, dtype=dt.np_dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                out[sl_out] = chunk[sl_in]
        if btree != UNDEF:
            walk(btree)
        return out


def _as_str(x):
    if isinstance(x, bytes):
        return x.split(b"\0")[0].decode("utf-8")
    if isinstance(x, np.bytes_):
        return bytes(x).split(b"\0")[0].decode("utf-8")
    return str(x)


def read_keras_weights(path):
    """{layer_name: [float32 arrays in layer.weights order]} from a Keras `save_weights` file (or a
    full-model file, whose weights sit under /model_weights — reference: mrcnn/model.py:2218-2219).
    Weights of nested models (the `rpn_model` layer) are filed under their own layer names
    (`rpn_conv_shared`, `rpn_class_raw`, `rpn_bbox_pred`), which is what by-name loading needs."""
    f = H5File(path)
    root = f.root["ohdr"]
    attrs = f.attributes(root)
    links = f.group_links(root)
    if "layer_names" not in attrs and "model_weights" in links:
        root = links["model_weights"]
        attrs = f.attributes(root)
        links = f.group_links(root)
    if "layer_names" not in attrs:
        raise H5Error("%s: no 'layer_names' attribute — not a Keras weight file" % path)
    layer_names = [_as_str(n) for n in np.atleast_1d(attrs["layer_names"])]
    out = {}
    for lname in layer_names:
        if lname not in links:
            continue
        g = links[lname]
        gattrs = f.attributes(g)
        wnames = [_as_str(n) for n in np.atleast_1d(gattrs.get("weight_names", []))] if gattrs.get("weight_names") is not None else []
        prefixes = {wn.split("/")[0] for wn in wnames if "/" in wn}
        nested = len(prefixes) > 1          # a nested Model (e.g. rpn_model): file under the inner layer names
        for wn in wnames:
            node = g
            for part in wn.split("/"):
                node = f.group_links(node)[part]
            arr = np.ascontiguousarray(f.read_dataset(node), dtype=np.float32)
            out.setdefault(wn.split("/")[0] if nested else lname, []).append(arr)
    return out


# --------------------------------------------------------------------------------------------
# writer (Keras 2.2.4 + h5py<=2.10 default layout)
# --------------------------------------------------------------------------------------------

class _Writer(object):
    LEAF_K = 64        # symbol-table node holds up to 2*LEAF_K entries
    INTERNAL_K = 16

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)   # superblock v0 placeholder (8+8+4+4 + 4*8 + 40 = 96)

    def _align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def alloc(self, data):
        self._align()
        addr = len(self.buf)
        self.buf += data
        return addr

    # -- messages ---------------------------------------------------------------------------------
    @staticmethod
    def _msg(mtype, data, flags=0):
        data = bytes(data)
        pad = (-len(data)) % 8
        return struct.pack("<HHB3x", mtype, len(data) + pad, flags) + data + b"\0" * pad

    @staticmethod
    def _dt_f32():
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dt_str(n):
        return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)       # null-terminated ASCII, size n

    @staticmethod
    def _ds(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)

    def _attr(self, name, dt, ds, payload):
        nb = name.encode() + b"\0"
        p8 = lambda b: b + b"\0" * ((-len(b)) % 8)  # noqa: E731
        body = struct.pack("<BxHHH", 1, len(nb), len(dt), len(ds)) + p8(nb) + p8(dt) + p8(ds) + payload
        return self._msg(0x0C, body)

    def attr_strings(self, name, strings):
        enc = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
        n = max([len(e) for e in enc] + [1])
        payload = b"".join(e.ljust(n, b"\0") for e in enc)
        return self._attr(name, self._dt_str(n), self._ds((len(enc),)), payload)

    def attr_string_scalar(self, name, s):
        e = s.encode()
        return self._attr(name, self._dt_str(max(len(e), 1)), self._ds(()), e or b"\0")

    def object_header(self, msgs):
        body = b"".join(msgs)
        hdr = struct.pack("<BxHII4x", 1, len(msgs), 1, len(body))
        return self.alloc(hdr + body)

    # -- objects ----------------------------------------------------------------------------------
    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        return self.object_header([self._msg(0x01, self._ds(arr.shape)), self._msg(0x03, self._dt_f32(), 1),
                                   self._msg(0x08, layout)])

    def group(self, children, attr_msgs=()):
        """children: {name: object header address}. Returns the group's object header address."""
        names = sorted(children, key=lambda s: s.encode())
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode() + b"\0"
            while len(heap_data) % 8:
                heap_data.append(0)
        heap_data += b"\0" * 16                                  # a free block to keep libhdf5 happy
        free_off = len(heap_data) - 16
        heap_data[free_off:free_off + 16] = struct.pack("<QQ", 1, 16)
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        per = 2 * self.LEAF_K
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), per):
            chunk = names[i:i + per]
            body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            body += b"\0" * (40 * (per - len(chunk)))
            snods.append(self.alloc(body))
            keys.append(offs[chunk[-1]] if chunk else 0)
        if len(snods) > 2 * self.INTERNAL_K:
            raise H5Error("too many links for a single-level group B-tree")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, s_addr in enumerate(snods):
            node += struct.pack("<QQ", keys[i], s_addr)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\0" * (16 * (2 * self.INTERNAL_K - len(snods)))
        btree_addr = self.alloc(node)
        ohdr = self.object_header([self._msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + list(attr_msgs))
        return ohdr, btree_addr, heap_addr

    def finish(self, root_ohdr, root_btree, root_heap):
        self._align()
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, weights, layer_order=None, keras_version="2.2.4", backend="tensorflow", nest_rpn=True):
    """Writes {layer_name: [arrays]} in the Keras `save_weights` layout. With nest_rpn the three RPN
    layers are stored the way Keras stores the nested `rpn_model` (one group, six weights)."""
    w = _Writer()
    kinds = {2: ["kernel:0", "bias:0"], 4: ["gamma:0", "beta:0", "moving_mean:0", "moving_variance:0"]}
    names = list(layer_order) if layer_order else list(weights)
    rpn = [n for n in ("rpn_conv_shared", "rpn_class_raw", "rpn_bbox_pred") if n in weights] if nest_rpn else []
    top_children, layer_names = {}, []

    def layer_group(lname, members):
        """members: [(inner_layer_name, arrays)] -> group with nested <inner>/<weight> datasets."""
        sub, wnames = {}, []
        for inner, arrays in members:
            leaf = {}
            for a, suffix in zip(arrays, kinds[len(arrays)]):
                leaf[suffix] = w.dataset(a)
                wnames.append("%s/%s" % (inner, suffix))
            sub[inner] = w.group(leaf)[0]
        return w.group(sub, [w.attr_strings("weight_names", wnames)])[0]

    done_rpn = False
    for lname in names:
        if lname in rpn:
            if not done_rpn:
                top_children["rpn_model"] = layer_group("rpn_model", [(n, weights[n]) for n in rpn])
                layer_names.append("rpn_model")
                done_rpn = True
            continue
        top_children[lname] = layer_group(lname, [(lname, weights[lname])])
        layer_names.append(lname)
    attrs = [w.attr_strings("layer_names", layer_names), w.attr_string_scalar("backend", backend),
             w.attr_string_scalar("keras_version", keras_version)]
    root, bt, hp = w.group(top_children, attrs)
    with open(path, "wb") as f:
        f.write(w.finish(root, bt, hp))
