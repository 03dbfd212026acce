"""MAT-file loading for the LRS-PnP data path (no h5py dependency).

The reference opens its cubes two ways: ``scipy.io.loadmat`` for v5 files
(masks, dictionary, ``low_rank_sparsity_noisy.mat``; main_LRS_PnP.py:159,183)
and ``h5py.File`` for v7.3 files (clean / noisy_imgN cubes;
main_LRS_PnP.py:170-180).  h5py is not available here, so v7.3 files are read
with a minimal HDF5 parser that covers exactly the structures those files use:
superblock v0 behind the 512-byte MATLAB user block, v1 object headers, one
chunked (v1 B-tree, deflate) or contiguous dataset per variable.

``loadmat_any`` sniffs the header, so callers do not need to know which flavour
a file is (the DIP scripts call h5py on a file that is in fact v5,
main_LRS_PnP_DIP_pro.py:278).
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict

import numpy as np

_HDF5_SIG = b"\x89HDF\r\n\x1a\n"


class _H5:
    def __init__(self, buf: bytes):
        self.buf = buf
        sb = buf.find(_HDF5_SIG)
        if sb < 0:
            raise ValueError("not an HDF5 file")
        self.sb = sb
        ver = buf[sb + 8]
        if ver != 0:
            raise ValueError(f"unsupported HDF5 superblock version {ver}")
        if buf[sb + 13] != 8 or buf[sb + 14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        self.base = struct.unpack_from("<Q", buf, sb + 24)[0]
        # addresses in the file are relative to the base address; MATLAB writes
        # base == 512 for its user block, but be tolerant of base == 0 files
        # whose superblock sits at 512.
        if self.base == 0 and sb != 0:
            self.base = sb
        root = sb + 24 + 32  # root group symbol-table entry
        self.root_ohdr = struct.unpack_from("<Q", buf, root + 8)[0]
        cache_type = struct.unpack_from("<I", buf, root + 16)[0]
        if cache_type != 1:
            raise ValueError("root group without cached symbol table")
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", buf, root + 24)

    def a(self, addr: int) -> int:
        return addr + self.base

    # ---- groups -----------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        o = self.a(heap_addr)
        assert self.buf[o:o + 4] == b"HEAP"
        return struct.unpack_from("<Q", self.buf, o + 24)[0]

    def _name(self, heap_data: int, off: int) -> str:
        o = self.a(heap_data) + off
        e = self.buf.index(b"\0", o)
        return self.buf[o:e].decode()

    def _walk_group(self, btree: int, heap_data: int, out: Dict[str, int]):
        o = self.a(btree)
        assert self.buf[o:o + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", self.buf, o + 4)
        assert ntype == 0
        p = o + 24
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, p + 8)[0]
            p += 16
            if level > 0:
                self._walk_group(child, heap_data, out)
            else:
                s = self.a(child)
                assert self.buf[s:s + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", self.buf, s + 6)[0]
                for j in range(nsym):
                    e = s + 8 + 40 * j
                    noff, ohdr = struct.unpack_from("<QQ", self.buf, e)
                    out[self._name(heap_data, noff)] = ohdr

    def members(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        self._walk_group(self.root_btree, self._heap_data(self.root_heap), out)
        return out

    # ---- object headers ---------------------------------------------------
    def _messages(self, ohdr: int):
        o = self.a(ohdr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.buf, o)
        if ver != 1:
            raise ValueError(f"unsupported object header version {ver}")
        blocks = [(o + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, p)
                body = p + 8
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from("<QQ", self.buf, body)
                    blocks.append((self.a(coff), clen))
                msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def dataset(self, ohdr: int) -> np.ndarray:
        shape = dtype = layout = None
        deflate = False
        for mtype, b, msize in self._messages(ohdr):
            if mtype == 0x0001:
                ver, rank, flags = struct.unpack_from("<BBB", self.buf, b)
                doff = b + (8 if ver == 1 else 4)
                shape = struct.unpack_from(f"<{rank}Q", self.buf, doff)
            elif mtype == 0x0003:
                cv = self.buf[b]
                cls = cv & 0x0F
                bits0 = self.buf[b + 1]
                size = struct.unpack_from("<I", self.buf, b + 4)[0]
                if bits0 & 1:
                    raise ValueError("big-endian data not supported")
                if cls == 1:
                    dtype = np.dtype(f"<f{size}")
                elif cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                else:
                    raise ValueError(f"unsupported datatype class {cls}")
            elif mtype == 0x0008:
                ver, cls = struct.unpack_from("<BB", self.buf, b)
                if ver != 3:
                    raise ValueError(f"unsupported layout version {ver}")
                if cls == 1:
                    addr, size = struct.unpack_from("<QQ", self.buf, b + 2)
                    layout = ("contiguous", addr, size)
                elif cls == 2:
                    ndim = self.buf[b + 2]
                    addr = struct.unpack_from("<Q", self.buf, b + 3)[0]
                    cdims = struct.unpack_from(f"<{ndim}I", self.buf, b + 11)
                    layout = ("chunked", addr, cdims)
                else:
                    raise ValueError("compact layout not supported")
            elif mtype == 0x000B:
                ver, nf = struct.unpack_from("<BB", self.buf, b)
                p = b + 8
                for _ in range(nf):
                    fid, nlen, _fl, ncv = struct.unpack_from("<HHHH", self.buf, p)
                    p += 8 + ((nlen + 7) // 8) * 8 + 4 * (ncv + (ncv & 1))
                    if fid == 1:
                        deflate = True
                    else:
                        raise ValueError(f"unsupported HDF5 filter {fid}")
        if shape is None or dtype is None or layout is None:
            raise ValueError("object is not a simple dataset")
        out = np.zeros(shape, dtype=dtype)
        if layout[0] == "contiguous":
            _, addr, size = layout
            o = self.a(addr)
            out[...] = np.frombuffer(self.buf, dtype, count=out.size, offset=o).reshape(shape)
            return out
        _, addr, cdims = layout
        self._read_chunks(addr, cdims[:-1], out, deflate)
        return out

    def _read_chunks(self, btree: int, cshape, out: np.ndarray, deflate: bool):
        o = self.a(btree)
        assert self.buf[o:o + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", self.buf, o + 4)
        assert ntype == 1
        rank = out.ndim
        ksz = 8 + 8 * (rank + 1)
        p = o + 24
        for _ in range(used):
            csize, _fmask = struct.unpack_from("<II", self.buf, p)
            offs = struct.unpack_from(f"<{rank}Q", self.buf, p + 8)
            child = struct.unpack_from("<Q", self.buf, p + ksz)[0]
            p += ksz + 8
            if level > 0:
                self._read_chunks(child, cshape, out, deflate)
                continue
            raw = self.buf[self.a(child):self.a(child) + csize]
            if deflate:
                raw = zlib.decompress(raw)
            chunk = np.frombuffer(raw, out.dtype, count=int(np.prod(cshape))).reshape(cshape)
            sl_out, sl_in = [], []
            for d in range(rank):
                n = min(cshape[d], out.shape[d] - offs[d])
                sl_out.append(slice(offs[d], offs[d] + n))
                sl_in.append(slice(0, n))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]


def load_v73(path: str) -> Dict[str, np.ndarray]:
    """Read every top-level dataset of a MATLAB v7.3 file, h5py-style (i.e. the
    axes come back reversed w.r.t. MATLAB, as the reference expects:
    'received 36 36 128 1', main_LRS_PnP.py:171)."""
    with open(path, "rb") as f:
        buf = f.read()
    h5 = _H5(buf)
    out = {}
    for name, ohdr in h5.members().items():
        if name.startswith("#"):
            continue
        try:
            out[name] = h5.dataset(ohdr)
        except ValueError:
            continue
    return out


def loadmat_any(path: str) -> Dict[str, np.ndarray]:
    """v5 → scipy.io.loadmat ; v7.3 → mini HDF5 reader.  Keys without the
    scipy ``__header__`` bookkeeping entries."""
    with open(path, "rb") as f:
        head = f.read(128)
    if head.startswith(b"MATLAB 7.3"):
        return load_v73(path)
    from scipy.io import loadmat

    return {k: v for k, v in loadmat(path).items() if not k.startswith("__")}


def load_cube(path: str, key: str | None = None) -> np.ndarray:
    """Load a hyperspectral cube as the reference's ``[1, bands, d2, d3]`` f32
    tensor layout (main_LRS_PnP.py:174 for v7.3; the v5 file already arrives
    in that layout)."""
    d = loadmat_any(path)
    if key is None:
        cands = [k for k in ("masked_image", "clean_image") if k in d]
        if not cands:
            raise KeyError(f"no cube variable in {path}: {list(d)}")
        key = cands[0]
    arr = np.asarray(d[key], dtype=np.float32)
    with open(path, "rb") as f:
        v73 = f.read(10).startswith(b"MATLAB 7.3")
    if v73:
        arr = arr.transpose((-1, 2, 1, 0))
    return np.ascontiguousarray(arr)


def unfold_cube(cube: np.ndarray) -> np.ndarray:
    """``[1,B,d2,d3]`` → ``Y[d3*d2, B]`` with row = a*d2 + b for cube[0,c,b,a]
    (main_LRS_PnP.py:209)."""
    _, B, d2, d3 = cube.shape
    return np.ascontiguousarray(cube.reshape(B, d2, d3).transpose(2, 1, 0).reshape(d3 * d2, B))


def fold_cube(Y: np.ndarray, d2: int, d3: int) -> np.ndarray:
    """Inverse of :func:`unfold_cube` (main_LRS_PnP.py:369,398)."""
    B = Y.shape[1]
    return np.ascontiguousarray(Y.reshape(d3, d2, B).transpose(2, 1, 0).reshape(1, B, d2, d3))


def unfold_mask(msk: np.ndarray, bands: int) -> np.ndarray:
    """``msk (1,1,d2,d3)`` → band-replicated ``mask[d3*d2, bands]`` f32
    (main_LRS_PnP.py:188-192)."""
    single = np.asarray(msk, dtype=np.float32).transpose(0, 1, 3, 2).reshape(-1)
    return np.ascontiguousarray(np.repeat(single[:, None], bands, axis=1))
