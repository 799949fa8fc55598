"""End to end through the stand-in executable: MONORTM.IN + MONORTM_PROF.IN + TAPE3 -> MONORTM.OUT,
checked against the oracle's spectra written with an independent statement of STOREOUT's formats."""
import os
import subprocess

import numpy as np
import pytest

from monortm_b200 import api, driver, linefile, synth

import harness
import storeout_ref as sref
from test_driver_host import rec12

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "monortm_b200", "bin", "monortm_b200")


def _prof_lines(name, angle=None):
    src = open(os.path.join(GOLD, name)).read().split("\n")
    while src and not src[-1].strip():
        src.pop()
    if angle is not None:
        src[0] = src[0][:65] + "%8.3f" % angle + src[0][73:]
    return src


def _expected(workdir, ctrl, nprof, scale=None, xs=None):
    """Oracle spectra of every profile of workdir/MONORTM_PROF.IN -> the text STOREOUT would write.
    xs = (regions, xamnt): cross sections through the oracle's MONORTM_XSEC_SUB (IXSECT=1)."""
    wn = ctrl["wn"]
    ls = linefile.read_tape3(os.path.join(workdir, "TAPE3"), float(wn[0]), float(wn[-1]))
    emiss, reflc = driver.emiss_reflec(ctrl, wn)
    tmpsfc = ctrl["tmpbnd"]
    text, ids, layers = [], None, []
    for ip in range(nprof):
        pr = driver.read_profile(os.path.join(workdir, "MONORTM_PROF.IN"), ip)
        nlay, nmol = pr["nlay"], pr["nmol"]
        wkl = pr["wkl"][:, :, 0].copy(order="F")
        if scale:
            for m, f in scale.items():
                wkl[m - 1, :] = wkl[m - 1, :] * f
        scor = api.scor_for_layers(nmol, pr["t"][:, 0])
        odx_in = harness.oracle_xsec(xs[0], wn, pr["p"][:, 0], pr["t"][:, 0], xs[1]) if xs else None
        m = harness.oracle_modm(ls, wn, ctrl["dvset"], pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], nmol, wkl,
                                pr["wbrodl"][:, 0], scor, cntnm=ctrl["cntnm"], ibrd=ctrl["ibrd"], selection=False, odxsec_in=odx_in)
        tmr = harness.oracle_calctmr(wn, pr["t"][:, 0], pr["tz"][:, 0], m["o"])
        r = harness.oracle_rtm(ctrl["iplot"], pr["irt"], wn, pr["t"][:, 0], pr["tz"][:, 0], m["o"], tmpsfc, reflc, emiss)
        tmpsfc = r["tmpsfc"]                                   # RTM leaves 2.75 behind for IRT 2,3 (RTMmono.f90:122)
        if ids is None:
            ids = sref.id_mols(wkl, pr["wbrodl"][:, 0], nmol)
        otot, obm, odx = sref.layer_sums(m["o"], m["o_by_mol"], m["oc"], odx_in)
        wv = 0.0
        for l in range(nlay):
            wv = wv + wkl[0, l]
        clwc = 0.0
        for l in range(nlay):
            clwc = clwc + pr["clw"][l, 0]
        text += sref.header(len(wn), wn, ids)
        text += sref.rows(ip + 1, wn, r["tb"], tmr, r["rad"], r["trtot"], wv * 2.99150e-23, clwc, tmpsfc, emiss, reflc,
                          pr["angle"], otot, obm, odx, ids)
        layers.append(m["o"])
    return text, layers


def _same_records(got, want, rel=2e-9):
    """Same records: identical layout, numbers equal to the printed precision (one unit in the last place tolerated,
    the GPU and the oracle differ by ~1e-14 relative before rounding)."""
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if g == w:
            continue
        assert len(g) == len(w), (g, w)
        gt, wt = g.split(), w.split()
        assert len(gt) == len(wt), (g, w)
        for a, b in zip(gt, wt):
            if a != b:
                fa, fb = float(a), float(b)
                digits = len(b.split("E")[0].split(".")[1]) if "." in b else 0
                ulp = 10.0 ** (-digits) * (10.0 ** int(b.split("E")[1]) if "E" in b else 1.0)
                assert abs(fa - fb) <= 1.01 * ulp and abs(fa - fb) <= max(rel * abs(fb), 1.01 * ulp), (a, b, g)


def _write_tape3(workdir, n_filler=300):
    linefile.write_tape3(os.path.join(workdir, "TAPE3"), synth.synthetic_records(n_filler, seed=77))


def test_executable_reproduces_oracle_monortm_out_for_the_reference_fixtures(tmp_path):
    """run/in/MONORTM.IN_IATM0_dn (4 MWR channels) over three profiles: the fixture profile looking up, the liquid-cloud
    fixture, and the fixture looking down -- which inherits TBOUND = 2.75 K from the earlier RTM calls like the reference."""
    wd = str(tmp_path)
    with open(os.path.join(wd, "MONORTM.IN"), "w") as f:
        f.write(open(os.path.join(GOLD, "MONORTM.IN_IATM0_dn")).read())
    prof = _prof_lines("MONORTM_PROF.IN_sav") + _prof_lines("MONORTM_PROF.IN_liquid_cloud") + _prof_lines("MONORTM_PROF.IN_sav", 180.0)
    with open(os.path.join(wd, "MONORTM_PROF.IN"), "w") as f:
        f.write("\n".join(prof) + "\n")
    _write_tape3(wd)
    res = subprocess.run([EXE, "-C", wd], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "PROCESSING PROFILE NUMBER:    3" in res.stdout
    ctrl = driver.read_control(os.path.join(wd, "MONORTM.IN"))
    want, _ = _expected(wd, ctrl, 3)
    got = open(os.path.join(wd, "MONORTM.OUT")).read().split("\n")
    assert got[-1] == ""
    _same_records(got[:-1], want)
    rows = [g for g in got if g.startswith("    3")]
    assert len(rows) == 4 and all(r[83:91] == "    2.75" and r[107:116] == "  180.000" for r in rows)
    assert os.path.exists(os.path.join(wd, "MONORTM.LOG"))


def test_library_driver_gridded_scaled_profile_with_layer_files(tmp_path):
    """Gridded record 1.3 (V1, V2, DVSET), ICNTNM=3, IOD=1 and a profile-scaling record (H2O x1.2, CO2 x0.5)."""
    wd = str(tmp_path)
    r13 = "     0.700     1.100" + " " * 10 + "     0.050" + " " * 60 + "    2"
    with open(os.path.join(wd, "MONORTM.IN"), "w") as f:
        f.write("\n".join(["$ gridded", rec12(icntnm=3, iod=1), r13, "11", "  1.2000000E+00  5.0000000E-01",
                           "   285.00       0.8       0.1       0.0       0.2", "%"]) + "\n")
    with open(os.path.join(wd, "MONORTM_PROF.IN"), "w") as f:
        f.write("\n".join(_prof_lines("MONORTM_PROF.IN_liquid_cloud", 180.0)) + "\n")
    _write_tape3(wd)
    driver.run_monortm(wd)
    ctrl = driver.read_control(os.path.join(wd, "MONORTM.IN"))
    assert len(ctrl["wn"]) == 9 and ctrl["cntnm"][1] == 0.0 and ctrl["iod"] == 1
    want, layers = _expected(wd, ctrl, 1, scale={1: 1.2, 2: 0.5})
    got = open(os.path.join(wd, "MONORTM.OUT")).read().split("\n")
    _same_records(got[:-1], want)
    od = open(os.path.join(wd, "ODmono_prf0001_lay0003")).read().split("\n")
    assert od[0] == "NWN :       9" and len(od) == 12
    vals = np.array([float(x.split()[1]) for x in od[2:11]])
    assert np.allclose(vals, layers[0][:, 2], rtol=6e-4)             # e12.4 keeps four digits


def test_driver_with_cross_sections(tmp_path):
    """IXSECT=1 end to end: the cross-section block of MONORTM_PROF.IN (monortm.f90:491-527), XSREAD on
    FSCDXS, MONORTM_XSEC_SUB inside the step, the XSEC_OD column of MONORTM.OUT."""
    from monortm_b200 import xsfile
    wd = str(tmp_path)
    xsfile.synthetic_set(wd)
    r13 = "     3.000    11.000" + " " * 10 + "     1.000"
    with open(os.path.join(wd, "MONORTM.IN"), "w") as f:
        f.write("\n".join(["$ cross sections", rec12(ixsect=1), r13, "   288.20       0.9       0.0       0.0       0.1", "%"]) + "\n")
    prof = _prof_lines("MONORTM_PROF.IN_sav", 180.0)
    nlay = 19
    rng = np.random.default_rng(5)
    xamnt = np.zeros((38, nlay), order="F")
    xamnt[0] = 10.0 ** rng.uniform(14, 16, nlay)
    xamnt[1] = 10.0 ** rng.uniform(12, 14, nlay)
    prof += ["%5d%5s%5d" % (2, "", 0), "%-10s%-10s" % ("HNO3", "CFC11"),
             " 1%3d%5d  1.000000XS AMOUNTS      H1=    0.00 H2=   20.00 ANG= 180.000 LEN= 0" % (nlay, 2)]
    for l in range(nlay):
        prof.append("  0.0000000E+00    0.0000    0.0000   0")
        prof.append("".join("%15.7E" % v for v in list(xamnt[:7, l]) + [0.0]))
        xamnt[:, l] = [float("%15.7E" % v) for v in xamnt[:, l]]
    with open(os.path.join(wd, "MONORTM_PROF.IN"), "w") as f:
        f.write("\n".join(prof) + "\n")
    _write_tape3(wd)
    driver.run_monortm(wd)
    ctrl = driver.read_control(os.path.join(wd, "MONORTM.IN"))
    assert ctrl["ixsect"] == 1 and len(ctrl["wn"]) == 9
    regs = xsfile.read_regions(wd, ["HNO3", "CFC11"], 3.0, 11.0)
    assert [(r["ixmol"], r["v1fx"]) for r in regs] == [(0, 2.0)]                  # the F11 regions lie outside 3-11 cm-1
    want, _ = _expected(wd, ctrl, 1, xs=(regs, xamnt))
    got = open(os.path.join(wd, "MONORTM.OUT")).read().split("\n")
    _same_records(got[:-1], want)
    assert all(float(g.split()[-1]) > 0 for g in got[4:-1])                       # XSEC_OD column


def test_driver_stops_like_the_reference(tmp_path):
    wd = str(tmp_path)
    with open(os.path.join(wd, "MONORTM.IN"), "w") as f:
        f.write(open(os.path.join(GOLD, "MONORTM.IN_MDL_ATM_up")).read())
    with pytest.raises(api.MonortmError, match="IATM=1"):
        driver.run_monortm(wd)
    with open(os.path.join(wd, "MONORTM.IN"), "w") as f:
        f.write(open(os.path.join(GOLD, "MONORTM.IN_IATM0_dn")).read())
    with pytest.raises(api.MonortmError, match="GETPROFNUMBER"):
        driver.run_monortm(wd)
    with open(os.path.join(wd, "MONORTM_PROF.IN"), "w") as f:
        f.write("\n".join(_prof_lines("MONORTM_PROF.IN_sav")) + "\n")
    with pytest.raises(api.MonortmError, match="GET_LNFL"):             # no TAPE3
        driver.run_monortm(wd)
    res = subprocess.run([EXE, "-C", wd, "-q"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 1 and "STOP" in res.stderr
