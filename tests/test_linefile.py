"""TAPE3 reader/writer (harness side): the C++ host reader (mrtm_host_get_lnfl) and the
oracle's restatement of GET_LNFL must produce identical lnfl_mod arrays from the same file."""
import os
import struct
import tempfile

import numpy as np
import pytest

import harness
from monortm_b200 import linefile, synth


@pytest.fixture(scope="module")
def tape(tmp_path_factory):
    recs = synth.synthetic_records(700, seed=7, n_co2=6, n_sdep=5, n_generic_lc=5, brd_fraction=0.3)
    path = str(tmp_path_factory.mktemp("t3") / "TAPE3")
    linefile.write_tape3(path, recs)
    return path, recs


def _same(a, b):
    assert np.array_equal(a.nblm, b.nblm)
    for n in linefile.LineStore.ARRAY_ORDER:
        assert np.array_equal(getattr(a, n), getattr(b, n)), n


def test_cpp_reader_equals_oracle_reader(tape):
    path, recs = tape
    for v1, v2 in ((0.2, 55.0), (10.0, 12.0), (0.0, 0.5), (60.0, 70.0)):
        a = linefile.read_tape3(path, v1, v2, iim=1200)
        b = harness.oracle_read_tape3(path, v1, v2, 1200)
        _same(a, b)
        assert a.n_records() > 0


def test_full_read_recovers_every_record_and_field_semantics(tape):
    path, recs = tape
    ls = linefile.read_tape3(path, 0.0, 80.0, iim=1200)
    assert ls.n_records() == len(recs)
    # molecule bucket = mod(mol,100) of the line (coefficient records follow their parent, lnfl_mod.f90:46-64)
    parent_mol = None
    counts = np.zeros(39, int)
    for r in recs:
        if r["iflg"] >= 0:
            parent_mol = int(r["mol"]) % 100
        counts[parent_mol - 1] += 1
    assert np.array_equal(counts, ls.nblm)
    # first H2O record
    h2o = recs[(recs["iflg"] >= 0) & (recs["mol"] % 100 == 1)][0]
    assert ls.xnu0[0, 0] == h2o["vnu"] and ls.s0[0, 0] == np.float64(h2o["sp"])
    assert ls.alps[0, 0] == np.float64(h2o["hwhm"]) and ls.iso[0, 0] == (int(h2o["mol"]) % 1000) // 100
    # O2: air->foreign width correction with rvmr=0.21 (lnfl_mod.f90:98-104)
    o2 = recs[(recs["iflg"] >= 0) & (recs["mol"] % 100 == 7)]
    j = 0
    k = 0
    while k < ls.nblm[6] and ls.xg[6, k] != 0:   # find first uncoupled O2 line in the store
        k += 2
    first_plain = [r for r in o2 if r["iflg"] == 0][0]
    idx = [i for i in range(int(ls.nblm[6])) if ls.xnu0[6, i] == first_plain["vnu"]][0]
    exp = (np.float64(first_plain["alfa"]) - 0.21 * np.float64(first_plain["hwhm"])) / (1.0 - 0.21)
    assert ls.alpf[6, idx] == exp
    # XG = -|iflg| for lines and coefficient records alike (:75-79); RMOL carries the float bits of MOL
    lc = [i for i in range(int(ls.nblm[6])) if ls.xg[6, i] == -1.0]
    assert len(lc) >= 2 and len(lc) % 2 == 0
    crec = [r for r in recs if r["iflg"] == -1][0]
    ci = [i for i in range(int(ls.nblm[6])) if ls.xnu0[6, i] == crec["vnu"] and ls.xg[6, i] == -1.0][0]
    assert ls.rmol[6, ci] == np.float64(np.array([crec["mol"]], "<i4").view("<f4")[0])


def test_block_granular_selection(tape):
    path, recs = tape
    # panels wholly below v1-25 are skipped; reading stops after the block whose last VNU > v2+25 (:116,161-168)
    full = linefile.read_tape3(path, 0.0, 80.0, iim=1200)
    part = linefile.read_tape3(path, 0.2, 1.2, iim=1200)
    assert 0 < part.n_records() < full.n_records()
    assert part.n_records() % 1 == 0
    high = linefile.read_tape3(path, 70.0, 75.0, iim=1200)
    assert high.xnu0[high.xnu0 > 0].min() > 0.0
    assert high.n_records() < full.n_records()


def test_blocks_never_split_a_coupled_group(tape):
    path, _ = tape
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    n = struct.unpack_from("<i", data, off)[0]
    off += 8 + n
    nblocks = 0
    while off < len(data):
        n = struct.unpack_from("<i", data, off)[0]
        assert n == 24
        vmin, vmax, nrec, nwds = struct.unpack_from("<ddii", data, off + 4)
        off += 8 + n
        n2 = struct.unpack_from("<i", data, off)[0]
        assert n2 == 39000 and nwds == 9750 and 1 <= nrec <= 250
        iflg = np.frombuffer(data, "<i4", 250, off + 4 + 9000)
        assert iflg[0] >= 0            # a block never starts with a coefficient record
        assert vmin <= vmax
        off += 8 + n2
        nblocks += 1
    assert nblocks >= 3


def test_reader_errors(tmp_path):
    from monortm_b200 import _capi
    lib = _capi.load_library()
    with pytest.raises(RuntimeError):
        linefile.read_tape3(str(tmp_path / "missing"), 0, 1, iim=10)
    # header without the isotope flag 'I' -> STOP ' PRLNHD - NO ISOTOPE INFO ON LINFIL ' (lnfl_mod.f90:297-302)
    recs = synth.synthetic_records(8, seed=1, with_physical=False)
    p = str(tmp_path / "bad")
    linefile.write_tape3(p, recs)
    raw = bytearray(open(p, "rb").read())
    raw[4 + 9 * 8 + 7] = ord("X")
    open(p, "wb").write(raw)
    with pytest.raises(RuntimeError):
        linefile.read_tape3(p, 0, 80, iim=64)
    # unknown coupling flag -> 'LC flag not recongnized' STOP (:61-63)
    recs = synth.synthetic_records(8, seed=1, with_physical=False)
    recs["iflg"][3] = -7
    linefile.write_tape3(p, recs)
    with pytest.raises(RuntimeError):
        linefile.read_tape3(p, 0, 80, iim=64)
    # too small a line store
    recs = synth.synthetic_records(64, seed=1, with_physical=False)
    linefile.write_tape3(p, recs)
    with pytest.raises(RuntimeError):
        linefile.read_tape3(p, 0, 80, iim=4)
