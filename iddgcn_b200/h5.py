"""Minimal pure-Python reader/writer for the Keras 2.x HDF5 weight files the reference saves/loads.

The reference checkpoints with `model.save_weights(<name>.h5)` (prediction/IDDGCN.py:181-199,396)
and restores with `model.load_weights` (IDDGCN_eval.py:46, GnnExplainer.py:121, IDDGCN_explain.py:172).
h5py is not available here, so this module parses exactly the subset those files use: superblock
v0, v1 object headers, symbol-table groups (TREE/HEAP/SNOD), contiguous little-endian datasets,
v1 attribute messages with fixed-length or variable-length strings.

`read_keras_weights(path)` -> list of (layer_name, [(weight_name, ndarray), ...]) in `layer_names`
order (the order Keras' topological loader uses); weights are loaded BY ORDER, never by name
(the shipped fold-3 file names its layers rgcn__layer_3..5).

`write_keras_weights(path, layers)` writes the same structure back (readable by h5py/Keras).
"""
from __future__ import annotations

import struct
from typing import Sequence

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise ValueError(f"unsupported superblock version {buf[8]}")
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("only 8-byte offsets/lengths are supported")
        self.base = struct.unpack_from("<Q", buf, 24)[0]
        self.root_header = struct.unpack_from("<Q", buf, 56 + 8)[0]
        self._gcol_cache: dict[int, dict[int, bytes]] = {}

    # -- object headers ---------------------------------------------------------------------
    def messages(self, addr: int):
        b = self.b
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError(f"unsupported object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((coff, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- groups -----------------------------------------------------------------------------
    def group_entries(self, header_addr: int) -> dict[str, int]:
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.b, body)
                return self._walk_btree(btree, self._heap_data(heap))
        raise ValueError("object is not a symbol-table group")

    def _heap_data(self, heap_addr: int) -> int:
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return struct.unpack_from("<Q", self.b, heap_addr + 24)[0]

    def _cstr(self, off: int) -> str:
        end = self.b.index(b"\0", off)
        return self.b[off:end].decode("utf-8")

    def _walk_btree(self, addr: int, heap_data: int) -> dict[str, int]:
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        _ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        out: dict[str, int] = {}
        pos = addr + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level > 0:
                out.update(self._walk_btree(child, heap_data))
            else:
                if b[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol node signature")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for s in range(nsym):
                    e = child + 8 + s * 40
                    name_off, obj = struct.unpack_from("<QQ", b, e)
                    out[self._cstr(heap_data + name_off)] = obj
        return out

    # -- attributes -------------------------------------------------------------------------
    def attributes(self, header_addr: int) -> dict[str, object]:
        out = {}
        for mtype, body, _ in self.messages(header_addr):
            if mtype != 0x0C:
                continue
            b = self.b
            ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", b, body)
            if ver != 1:
                raise ValueError(f"unsupported attribute message version {ver}")
            pad = lambda n: (n + 7) & ~7
            p = body + 8
            name = b[p:p + nsz].split(b"\0")[0].decode()
            p += pad(nsz)
            tclass, tsize = self._datatype(p)
            tpos = p
            p += pad(tsz)
            dims = self._dataspace(p)
            p += pad(ssz)
            n = int(np.prod(dims)) if dims else 1
            if tclass == 3:  # fixed strings
                vals = [b[p + i * tsize:p + (i + 1) * tsize].split(b"\0")[0].decode() for i in range(n)]
            elif tclass == 9:  # variable-length (strings)
                vals = []
                for i in range(n):
                    ln, gaddr, gidx = struct.unpack_from("<IQI", b, p + i * 16)
                    vals.append(self._gcol(gaddr)[gidx][:ln].decode())
            elif tclass in (0, 1):
                dt = self._np_dtype(tpos)
                vals = np.frombuffer(b, dt, n, p).copy()
            else:
                vals = None
            out[name] = vals if dims else (vals[0] if vals is not None and len(vals) else vals)
        return out

    def _gcol(self, addr: int) -> dict[int, bytes]:
        if addr in self._gcol_cache:
            return self._gcol_cache[addr]
        b = self.b
        if b[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap signature")
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        objs = {}
        p = addr + 16
        while p + 16 <= addr + size:
            idx, _ref, _, osz = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            objs[idx] = b[p + 16:p + 16 + osz]
            p += 16 + ((osz + 7) & ~7)
        self._gcol_cache[addr] = objs
        return objs

    def _datatype(self, p: int):
        cv = self.b[p]
        size = struct.unpack_from("<I", self.b, p + 4)[0]
        return cv & 0x0F, size

    def _np_dtype(self, p: int):
        tclass, size = self._datatype(p)
        bits0 = self.b[p + 1]
        order = ">" if (bits0 & 1) else "<"
        if tclass == 1:
            return np.dtype(f"{order}f{size}")
        if tclass == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        raise ValueError(f"unsupported datatype class {tclass}")

    def _dataspace(self, p: int):
        ver, rank, flags = struct.unpack_from("<BBB", self.b, p)
        if ver == 1:
            off = p + 8
        elif ver == 2:
            off = p + 4
        else:
            raise ValueError(f"unsupported dataspace version {ver}")
        return [struct.unpack_from("<Q", self.b, off + 8 * i)[0] for i in range(rank)]

    # -- datasets ---------------------------------------------------------------------------
    def dataset(self, header_addr: int) -> np.ndarray:
        dims = dt = None
        addr = size = None
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x01:
                dims = self._dataspace(body)
            elif mtype == 0x03:
                dt = self._np_dtype(body)
            elif mtype == 0x08:
                ver, cls = struct.unpack_from("<BB", self.b, body)
                if ver != 3 or cls != 1:
                    raise ValueError("only contiguous (layout v3 class 1) datasets are supported")
                addr, size = struct.unpack_from("<QQ", self.b, body + 2)
            elif mtype == 0x0B:
                raise ValueError("filtered datasets are not supported")
        if dims is None or dt is None or addr is None:
            raise ValueError("object is not a simple dataset")
        n = int(np.prod(dims)) if dims else 1
        if addr == UNDEF:
            return np.zeros(dims, dt.newbyteorder("="))
        arr = np.frombuffer(self.b, dt, n, self.base + addr).reshape(dims)
        return arr.astype(dt.newbyteorder("="), copy=True)

    def resolve(self, group_addr: int, path: str) -> int:
        addr = group_addr
        for part in path.split("/"):
            if part:
                addr = self.group_entries(addr)[part]
        return addr


def read_keras_weights(path) -> list[tuple[str, list[tuple[str, np.ndarray]]]]:
    with open(path, "rb") as f:
        r = _Reader(f.read())
    root = r.root_header
    layer_names = r.attributes(root).get("layer_names")
    if layer_names is None:
        raise ValueError("file has no layer_names attribute (not a Keras weights file)")
    groups = r.group_entries(root)
    out = []
    for lname in layer_names:
        gaddr = groups[lname]
        wnames = r.attributes(gaddr).get("weight_names")
        if wnames is None or isinstance(wnames, np.ndarray):
            wnames = [] if wnames is None else [w.decode() if isinstance(w, bytes) else str(w) for w in wnames]
        weights = [(w, r.dataset(r.resolve(gaddr, w))) for w in wnames]
        out.append((lname, weights))
    return out


def read_weight_list(path) -> list[np.ndarray]:
    """Flat list of weight arrays in load order (what `model.set_weights` would consume)."""
    return [arr for _, ws in read_keras_weights(path) for _, arr in ws]


# ------------------------------------------------------------------------------------------------
# writer: same subset (superblock v0, symbol-table groups, contiguous f32 datasets, fixed strings)
# ------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.buf = bytearray(b"\0" * 96)  # superblock (56) + root symbol table entry (40)

    def tell(self) -> int:
        return len(self.buf)

    def alloc(self, data: bytes, align: int = 8) -> int:
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
        body = body + b"\0" * ((-len(body)) % 8)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, msgs: Sequence[bytes]) -> int:
        payload = b"".join(msgs)
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(payload))
        return self.alloc(hdr + payload)

    @staticmethod
    def _dataspace(dims) -> bytes:
        return struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)

    @staticmethod
    def _dtype_f32() -> bytes:
        # class 1 (float) v1, little-endian, IEEE: bit fields 0x20 0x1f 0x00; size 4; props
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dtype_str(n: int) -> bytes:
        return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)  # null-terminated ASCII

    def attr_strings(self, name: str, values: Sequence[str]) -> bytes:
        width = max([len(v.encode()) for v in values] + [0]) + 1
        nm = name.encode() + b"\0"
        dt = self._dtype_str(width)
        ds = self._dataspace([len(values)])
        pad = lambda x: x + b"\0" * ((-len(x)) % 8)
        data = b"".join(v.encode().ljust(width, b"\0") for v in values)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + pad(nm) + pad(dt) + pad(ds) + data
        return self._msg(0x0C, body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        daddr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)
        msgs = [self._msg(0x01, self._dataspace(arr.shape)), self._msg(0x03, self._dtype_f32(), 1),
                self._msg(0x08, layout)]
        return self.object_header(msgs)

    def group(self, entries: dict[str, int], attrs: Sequence[bytes] = ()) -> int:
        names = sorted(entries)  # B-tree keys must be ordered
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += n.encode() + b"\0"
            while len(heap) % 8:
                heap.append(0)
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        # one SNOD per <= 8 symbols (2*leaf_k with leaf_k=4), a single level-0 TREE node over them
        snods = []
        for i in range(0, max(len(names), 1), 8):
            chunk = names[i:i + 8]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
            body += b"\0" * (40 * (8 - len(chunk)))
            snods.append((self.alloc(body), chunk))
        if len(snods) > 32:
            raise ValueError("too many entries for a single B-tree node")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        tree += struct.pack("<Q", 0)
        for addr, chunk in snods:
            tree += struct.pack("<QQ", addr, offs[chunk[-1]] if chunk else 0)
        tree += b"\0" * (16 * (32 - len(snods)))
        tree_addr = self.alloc(tree)
        msgs = [self._msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + list(attrs)
        return self.object_header(msgs)

    def finish(self, root_addr: int) -> bytes:
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII16x", 0, root_addr, 0, 0)
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, layers: Sequence[tuple[str, Sequence[tuple[str, np.ndarray]]]],
                        keras_version: str = "2.7.0", backend: str = "tensorflow") -> None:
    """Write `layers` = [(layer_name, [(weight_name like 'layer/kernel:0', array), ...]), ...]."""
    w = _Writer()
    root_entries = {}
    for lname, weights in layers:
        tree: dict = {}
        for wname, arr in weights:
            node = tree
            parts = wname.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = np.asarray(arr)

        def emit(node) -> int:
            ent = {k: (emit(v) if isinstance(v, dict) else w.dataset(v)) for k, v in node.items()}
            return w.group(ent)

        ent = {k: (emit(v) if isinstance(v, dict) else w.dataset(v)) for k, v in tree.items()}
        attrs = [w.attr_strings("weight_names", [wn for wn, _ in weights])] if weights else \
            [w.attr_strings("weight_names", [])]
        root_entries[lname] = w.group(ent, attrs)
    attrs = [w.attr_strings("layer_names", [ln for ln, _ in layers]), w.attr_strings("backend", [backend]),
             w.attr_strings("keras_version", [keras_version])]
    root = w.group(root_entries, attrs)
    with open(path, "wb") as f:
        f.write(w.finish(root))
