"""The C-ABI library loads without a GPU, exports every symbol include/monortm_b200.h declares, fails
loudly when no device exists, and its host helpers (TIPS, PROF reader, tables) behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import harness
from monortm_b200 import _capi, api, profio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "monortm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrtm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
        assert n in _capi.SIGNATURES, "ctypes signature missing for " + n
    assert b"monortm_b200" in lib.mrtm_version()
    assert b"no CPU fallback" in lib.mrtm_strerror(1)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.MonortmError) as e:
        api.Session(0)
    assert e.value.code == 1      # MRTM_ENODEV


def test_tips_scor():
    s296 = api.tips_2003(22, 296.0)
    for mol, niso in ((1, 6), (2, 9), (3, 9), (7, 3), (22, 1)):
        assert np.allclose(s296[mol - 1, :niso], 1.0, rtol=1e-14)
    # independent Lagrange evaluation on the committed table for H2O 161 at 250 K
    txt = open(os.path.join(ROOT, "monortm_b200", "csrc", "tables", "tips_tables.inc")).read()
    q = np.array([float(x) for x in re.search(r"TIPS_QOFT\[\d+\] = \{(.*?)\};", txt, re.S).group(1).split(",") if x.strip()])
    tdat = np.array([float(x) for x in re.search(r"TIPS_TDAT\[\d+\] = \{(.*?)\};", txt, re.S).group(1).split(",") if x.strip()])
    q161 = q[:119]

    def lag4(t):
        i = int(np.searchsorted(tdat, t))       # first tdat >= t  (0-based) -> Fortran I = i+1
        idx = [i - 2, i - 1, i, i + 1]
        out = 0.0
        for a in idx:
            w = 1.0
            for b in idx:
                if a != b:
                    w *= (t - tdat[b]) / (tdat[a] - tdat[b])
            out += w * q161[a]
        return out
    s = api.tips_2003(22, 250.0)
    assert abs(s[0, 0] - lag4(296.0) / lag4(250.0)) < 1e-13
    assert s[0, 0] > 1.0 and api.tips_2003(22, 320.0)[0, 0] < 1.0      # Q grows with T
    with pytest.raises(api.MonortmError):
        api.tips_2003(22, 60.0)                                          # outside 70-3000 K -> STOP
    # molecules 34..39: the reference does not stop at atomic oxygen -- it sets gi=1, QT=1
    # (tips_2003.f90:233-238) so scor(34,1)=1; 35-38 come from their tables; for CH3OH (39) both qt_296 and qt_temp end
    # up as the stale QT of the previous call (:260-268, 288-292), i.e. scor(39,1)=1 as well
    s39 = api.tips_2003(39, 250.0)
    assert s39[33, 0] == 1.0 and s39[38, 0] == 1.0
    assert np.all(s39[34:38, 0] > 1.0) and s39[34, 1] > 1.0 and s39[36, 1] > 1.0 and s39[37, 1] > 1.0
    assert np.array_equal(s39[:33], api.tips_2003(33, 250.0)[:33])
    with pytest.raises(api.MonortmError):
        api.tips_2003(40, 250.0)


def test_tables_spot_values():
    # literal values typed from the reference DATA statements
    inc = open(os.path.join(ROOT, "monortm_b200", "csrc", "tables", "mtckd_tables.inc")).read()

    def arr(name):
        m = re.search(name + r"\[(\d+)\] = \{(.*?)\};", inc, re.S)
        v = np.array([float(x) for x in m.group(2).split(",") if x.strip()])
        assert len(v) == int(m.group(1))
        return v
    s296 = arr("MTCKD_SH2O_296")
    assert len(s296) == 2003 and s296[0] == 2.731e-01 and s296[2] == 2.877E-01 and s296[-1] == 9.558e-12  # contnm.f90:1493-1497,1934
    s260 = arr("MTCKD_SH2O_260")
    assert s260[0] == 5.998e-01 and s260[1] == 6.382e-01                                                 # contnm.f90:2000
    fh = arr("MTCKD_FH2O")
    assert fh[0] == 1.205e-02 and fh[1] == 1.126e-02                                                     # contnm.f90:2509
    co2 = arr("MTCKD_FCO2")
    assert len(co2) == 5003 and co2[0] == 8.391e-13 and co2[2] == 8.345e-13                              # contnm.f90:3050-3052
    n2 = arr("MTCKD_N2RT_296")
    assert len(n2) == 73 and n2[0] == 0.4303E-06 and n2[-1] == 0.2428E-09                                # contnm.f90:4246,4260
    assert arr("MTCKD_N2RT_296_SF")[0] == 1.3534
    x = arr("MTCKD_XFAC_RHU")
    assert len(x) == 63 and x[0] == 0.7620 and x[-1] == 1.0450                                           # contnm.f90:186-202
    sm = open(os.path.join(ROOT, "monortm_b200", "csrc", "tables", "smass_table.inc")).read()
    v = [float(t) for t in re.search(r"\{(.*?)\};", sm, re.S).group(1).split(",") if t.strip()]
    assert v[0] == 18.01 and v[9 * 6] == 31.99 and v[9 * 21] == 28.01                                    # isotope.incl


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_tables_regenerate_identically(tmp_path):
    import subprocess
    import sys
    import shutil
    tabs = os.path.join(ROOT, "monortm_b200", "csrc", "tables")
    keep = str(tmp_path / "keep")
    shutil.copytree(tabs, keep)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py")], stdout=subprocess.DEVNULL)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_tables_ir.py")], stdout=subprocess.DEVNULL)
    for f in os.listdir(keep):
        assert open(os.path.join(keep, f)).read() == open(os.path.join(tabs, f)).read(), f


def test_prof_reader_on_reference_fixtures():
    g = os.path.join(ROOT, "tests", "golden")
    a = profio.read_prof_in(os.path.join(g, "MONORTM_PROF.IN_sav"))[0]
    b = profio.read_prof_in(os.path.join(g, "MONORTM_PROF.IN_liquid_cloud"))[0]
    assert a["nlay"] == 19 and a["nmol"] == 22 and a["irt"] == 3          # ANGLE=0 < 90 -> downwelling
    assert a["p"][0, 0] == 972.2109 and a["t"][0, 0] == 285.94 and a["tz"][0, 0] == 288.20 and a["tz"][1, 0] == 283.65
    assert a["wkl"][0, 0, 0] == 1.2207059E+22 and a["wbrodl"][0, 0] == 1.5169132E+22
    assert a["wkl"][21, 0, 0] == 1.3375841E+24
    assert np.all(a["clw"] == 0)
    assert list(b["clw"][2:5, 0]) == [0.03, 0.04, 0.03] and b["clw"].sum() == pytest.approx(0.10)
    for k in ("p", "t", "tz", "wkl", "wbrodl"):
        assert np.array_equal(a[k], b[k])
    assert np.all(np.diff(a["p"][:, 0]) < 0)                               # surface-most layer first (IDU=1)


def test_oracle_on_reference_profile_fixture():
    """C1: the shipped IATM=0 profile x the 4 MWR channels through the oracle, cloudy vs clear."""
    from monortm_b200 import synth
    g = os.path.join(ROOT, "tests", "golden")
    wn = synth.freq_c1_channels()
    ls = harness.synthetic_store(512, v1=float(wn[0]), v2=float(wn[-1]))
    out = {}
    for name in ("MONORTM_PROF.IN_sav", "MONORTM_PROF.IN_liquid_cloud"):
        pr = profio.read_prof_in(os.path.join(g, name))[0]
        scor = api.scor_for_layers(22, pr["t"])
        case = dict(ls=ls, wn=wn, dvset=0.0, prof=pr, scor=scor, irt=pr["irt"], cntnm=(1.,) * 7, ibrd=0, nmol=22,
                    tmpsfc=2.75, emiss=np.ones(4), reflc=np.zeros(4))
        out[name] = harness.run_oracle(case)
    clear, cloudy = out["MONORTM_PROF.IN_sav"], out["MONORTM_PROF.IN_liquid_cloud"]
    assert np.all(clear["tb"] > 5) and np.all(clear["tb"] < 150)           # downwelling TBs (synthetic filler lines add to the real ones)
    assert np.all(cloudy["tb"] > clear["tb"] + 1.0)                        # 0.1 mm of liquid adds several K
    assert np.all(cloudy["o_clw"][:, 2:5] > 0) and np.all(cloudy["o_clw"][:, :2] == 0)


def test_mrtm_opts_layout_matches_the_header_the_ctypes_mirror_and_the_fortran_shim(tmp_path):
    """sizeof / offsetof of mrtm_opts as the C compiler lays it out for include/monortm_b200.h against the ctypes mirror
    (monortm_b200/_capi.py) and the member order of the bind(c) type in the Fortran shim (which cannot be compiled here)."""
    import re
    import subprocess
    src = tmp_path / "lay.c"
    names = [n for n, _ in _capi.MrtmOpts._fields_]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "monortm_b200.h"\nint main(void){printf("%zu", sizeof(mrtm_opts));' +
                   "".join('printf(" %%zu", offsetof(mrtm_opts, %s));' % n for n in names) +
                   'printf(" %zu", sizeof(mrtm_xs_region));return 0;}\n')
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    vals = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == C.sizeof(_capi.MrtmOpts)
    assert vals[1:-1] == [getattr(_capi.MrtmOpts, n).offset for n in names]
    assert vals[-1] == C.sizeof(_capi.MrtmXsRegion)
    shim = open(os.path.join(ROOT, "monortm_b200", "shim", "monortm_gpu_shim.f90")).read()
    body = shim[shim.index("type, bind(c) :: mrtm_opts"):shim.index("end type mrtm_opts")]
    members = []
    for line in body.split("\n")[1:]:
        m = re.match(r"\s*(integer\(c_int32_t\)|integer\(c_int64_t\)|real\(c_double\)|type\(c_ptr\))\s*::\s*(.*?)(!.*)?$", line)
        if m:
            members += [x.split("=")[0].strip() for x in m.group(2).split(",")]
    assert members == names


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_generated_tables_agree_with_a_second_extraction_route():
    """The oracle and the product share csrc/tables/*.inc (tools/gen_tables.py parses the DATA statements with regular
    expressions).  A second, independent route -- the mechanical translator executes the BLOCK DATA units and the numbers are
    read out of the COMMON storage the accessor routines index -- must give the same values, so an extraction error cannot
    hide as a common-mode error of oracle and product."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_exec
    ns = ref_exec.load()

    def inc(fname, name):
        src = open(os.path.join(ROOT, "monortm_b200", "csrc", "tables", fname)).read()
        m = re.search(r"static const double %s\[(\d+)\] = \{(.*?)\};" % name, src, re.S)
        v = np.array([float(x) for x in m.group(2).replace("\n", " ").split(",") if x.strip()])
        assert len(v) == int(m.group(1))
        return v

    for blk, members in (("sh2o", [("MTCKD_SH2O_296", 2003)]), ("s260", [("MTCKD_SH2O_260", 2003)]), ("fh2o", [("MTCKD_FH2O", 2003)]),
                         ("fco2", [("MTCKD_FCO2", 5003)]), ("n2rt296", [("MTCKD_N2RT_296", 73), ("MTCKD_N2RT_296_SF", 73)]),
                         ("n2rt220", [("MTCKD_N2RT_220", 73), ("MTCKD_N2RT_220_SF", 73)])):
        s = ns["C_" + blk].s
        off = 4
        for name, n in members:
            assert np.array_equal(inc("mtckd_tables.inc", name), np.array([float(x) for x in s[off:off + n]])), name
            off += n
    smass = np.array([float(x) for x in ns["C_isvect"].s[39:39 + 39 * 9]]).reshape((39, 9), order="F")
    assert np.array_equal(inc("smass_table.inc", "ISO_SMASS").reshape(39, 9), smass)
