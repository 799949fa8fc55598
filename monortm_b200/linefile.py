"""TAPE3 line-file support for the harness: the in-memory line store exactly as
lnfl_mod holds it (src/lnfl_mod.f90:9-13), a reader that goes through the C++
host helper (mrtm_host_get_lnfl == GET_LNFL, src/lnfl_mod.f90:22-133) and a
writer of the binary format (src/struct_types.f90:27-43,
src/lnfl_mod.f90:250-252) used to make synthetic line files -- the reference
ships no TAPE3 (run/in/TAPE3_* is a dangling symlink).
"""
import ctypes as C
import struct

import numpy as np

from . import _capi

MXMOL = 39
NLINEREC = 250
BLOCK_WORDS = 9750  # INPUT_BLOCK = 39000 bytes


class LineStore:
    """Module arrays of lnfl_mod, Fortran (column-major) order."""

    def __init__(self, iim):
        self.iim = int(iim)
        f = dict(order="F")
        self.nblm = np.zeros(MXMOL, np.int64)
        self.iso = np.zeros((MXMOL, iim), np.int64, **f)
        for name in ("xnu0", "deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep"):
            setattr(self, name, np.zeros((MXMOL, iim), np.float64, **f))
        self.brd_mol_flg = np.zeros((7, 7, iim), np.int32, **f)
        self.brd_mol_tmp = np.zeros((7, 7, iim), np.float64, **f)
        self.brd_mol_hw = np.zeros((7, 7, iim), np.float64, **f)
        self.brd_mol_shft = np.zeros((7, 7, iim), np.float64, **f)

    # order of the array arguments shared by mrtm_stage_lines / mrtm_host_get_lnfl / the oracle
    ARRAY_ORDER = ("iso", "xnu0", "deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep",
                   "brd_mol_flg", "brd_mol_tmp", "brd_mol_hw", "brd_mol_shft")

    def pointers(self):
        return [getattr(self, n).ctypes.data_as(C.c_void_p) for n in self.ARRAY_ORDER]

    def n_records(self):
        return int(self.nblm.sum())


def read_tape3(path, v1, v2, iim=None):
    """GET_LNFL through the C++ host helper.  Returns a LineStore."""
    lib = _capi.load_library()
    if iim is None:
        import os
        iim = max(16, (os.path.getsize(path) // 39024 + 1) * NLINEREC + 8)
    ls = LineStore(iim)
    rc = lib.mrtm_host_get_lnfl(str(path).encode(), float(v1), float(v2), ls.iim,
                                ls.nblm.ctypes.data_as(C.c_void_p), *ls.pointers())
    if rc != 0:
        raise RuntimeError("mrtm_host_get_lnfl(%s) failed: %s" % (path, lib.mrtm_strerror(rc).decode()))
    return ls


# --------------------------------------------------------------------------- writer
REC_DTYPE = np.dtype([
    ("vnu", "<f8"), ("sp", "<f4"), ("alfa", "<f4"), ("epp", "<f4"), ("mol", "<i4"),
    ("hwhm", "<f4"), ("tmpalf", "<f4"), ("pshift", "<f4"), ("iflg", "<i4"),
    ("brd_flg", "<i4", (7,)), ("brd_dat", "<f4", (21,)), ("sdep", "<f4"),
])


def _f32_bits_as_i32(x):
    return np.array([x], "<f4").view("<i4")[0]


def coupling_record(y, g, iflg):
    """One coupling-coefficient pseudo-record: Y and G at 200/250/296/340 K sit in the fields
    VNU,SP,ALFA,EPP,MOL(bits),HWHM,TMPALF,PSHIFT (src/modm.f90:331-338)."""
    r = np.zeros((), REC_DTYPE)
    r["vnu"], r["sp"] = y[0], g[0]
    r["alfa"], r["epp"] = y[1], g[1]
    r["mol"], r["hwhm"] = _f32_bits_as_i32(y[2]), g[2]
    r["tmpalf"], r["pshift"] = y[3], g[3]
    r["iflg"] = -abs(int(iflg))
    return r


def _fortran_record(payload):
    n = struct.pack("<i", len(payload))
    return n + payload + n


def write_tape3(path, records):
    """Write `records` (structured array of REC_DTYPE, file order: ascending line centre, every
    line with IFLG in {1,3,5} immediately followed by its coefficient record(s)) as a TAPE3.
    Blocks hold at most 250 records and never separate a line from its coefficient records."""
    records = np.asarray(records, REC_DTYPE)
    n = len(records)
    is_line = records["iflg"] >= 0
    # ---- file header (1664 bytes), src/lnfl_mod.f90:250-252
    hlinid = [b"SYNTHTP3", b"monortm_", b"b200 syn", b"thetic  ", b"line fil", b"e       ", b"        ",
              b"        ", b"        ", b"LNFL 91I"]  # 8th char of HLINID(10) must be 'I' (:297-302)
    hdr = b"".join(h.ljust(8)[:8] for h in hlinid)
    hdr += b"".join(b"        " for _ in range(64))          # BMOLID
    hdr += np.zeros(64 * 3, "<i4").tobytes()                  # MOLCNT, MCNTLC, MCNTNL
    hdr += np.zeros(64, "<f4").tobytes()                      # SUMSTR
    lines_v = records["vnu"][is_line]
    flo = float(lines_v.min()) if len(lines_v) else 0.0
    fhi = float(lines_v.max()) if len(lines_v) else 0.0
    hdr += struct.pack("<iffiiiii", 39, flo, fhi, int(is_line.sum()), 0, 0, 0, 0)
    hdr += b"        " * 2                                    # HID1
    assert len(hdr) == 1664, len(hdr)
    with open(path, "wb") as f:
        f.write(_fortran_record(hdr))
        i = 0
        while i < n:
            j = min(i + NLINEREC, n)
            # do not start the next block with a coefficient record
            while j < n and j > i and records["iflg"][j] < 0:
                j -= 1
            if j == i:
                raise ValueError("coupling group longer than a block")
            blk = records[i:j]
            nrec = len(blk)
            lv = blk["vnu"][blk["iflg"] >= 0]
            vmin, vmax = float(lv.min()), float(lv.max())
            f.write(_fortran_record(struct.pack("<ddii", vmin, vmax, nrec, BLOCK_WORDS)))
            buf = bytearray(39000)

            def put(off, arr, width):
                a = np.zeros(NLINEREC * width, arr.dtype)
                a[:nrec * width] = arr.reshape(-1)
                buf[off:off + a.nbytes] = a.tobytes()

            put(0, blk["vnu"], 1)
            put(2000, blk["sp"], 1)
            put(3000, blk["alfa"], 1)
            put(4000, blk["epp"], 1)
            put(5000, blk["mol"], 1)
            put(6000, blk["hwhm"], 1)
            put(7000, blk["tmpalf"], 1)
            put(8000, blk["pshift"], 1)
            put(9000, blk["iflg"], 1)
            put(10000, blk["brd_flg"], 7)
            put(17000, blk["brd_dat"], 21)
            put(38000, blk["sdep"], 1)
            f.write(_fortran_record(bytes(buf)))
            i = j
    return n
