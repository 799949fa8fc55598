"""Host driver pieces, CPU only: RDLBLINP, MONORTM_PROF.IN reader, EMISS_REFLEC, STOREOUT formats."""
import os

import numpy as np
import pytest

from monortm_b200 import driver, profio
from monortm_b200.api import MonortmError

import storeout_ref as sref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rdlblinp_on_reference_control_files():
    c = driver.read_control(os.path.join(GOLD, "MONORTM.IN_IATM0_dn"))
    assert (c["ihirac"], c["icntnm"], c["iemit"], c["iplot"], c["iatm"], c["iod"], c["ixsect"], c["ispd"], c["ibrd"]) == \
        (1, 1, 1, 1, 0, 0, 0, 0, 0)
    assert c["cntnm"] == (1.0,) * 7 and c["dvset"] == 0.0 and c["v1"] == -0.2 and c["v2"] == 8.8
    assert np.array_equal(c["wn"], [0.789344, 0.79828, 1.043027, 1.051763])
    assert c["tmpbnd"] == 0.0 and c["bndemi"] == (1.0, 0.0, 0.0) and c["bndrfl"] == (0.0, 0.0, 0.0)
    c = driver.read_control(os.path.join(GOLD, "MONORTM.IN_MDL_ATM_up"))
    assert c["iatm"] == 1 and c["tmpbnd"] == 290.0 and c["bndemi"][0] == 0.6 and c["bndrfl"][0] == 0.4


def _control(tmp_path, rec12, rec13, extra=(), rec14="   288.20       0.9       0.0       0.0       0.1"):
    p = tmp_path / "MONORTM.IN"
    p.write_text("\n".join(["comment", "$ test", rec12] + [rec13] + list(extra) + [rec14, "%%%"]) + "\n")
    return str(p)


R13_GRID = "     0.200     1.200" + " " * 10 + "     0.010"


def rec12(ihirac=1, icntnm=1, iemit=1, iplot=1, iatm=0, iod=0, ixsect=0, ispd=0, ibrd=0):
    """record 1.2, FORMAT 925 (4X,I1,9X,I1,9X,I1,14X,I1,9X,I1,14X,I1,4X,I1,16X,I4,I4)"""
    r = [" "] * 94
    for col, v in ((5, ihirac), (15, icntnm), (25, iemit), (40, iplot), (50, iatm), (65, iod), (70, ixsect)):
        r[col - 1] = str(v)
    return "".join(r[:86]) + "%4d%4d" % (ispd, ibrd)


REC12 = rec12()


def test_rdlblinp_grid_modes_and_stops(tmp_path):
    # gridded: NWN = NINT((V2-V1)/DVSET + 1), WN(J) = V1+(J-1)*DVSET (monortm_sub.F90:279-290)
    f = _control(tmp_path, REC12, R13_GRID)
    c = driver.read_control(f)
    assert len(c["wn"]) == 101 and c["wn"][0] == 0.2 and c["wn"][100] == 0.2 + 100 * 0.01 and c["dvset"] == 0.01
    # implied decimal point of E10.3: a field without '.' is scaled by 1e-3
    f = _control(tmp_path, REC12, "       200      1200" + " " * 10 + "        10")
    assert np.array_equal(driver.read_control(f)["wn"], c["wn"])
    # single frequency
    f = _control(tmp_path, REC12, "     0.750     0.750")
    assert np.array_equal(driver.read_control(f)["wn"], [0.75])
    # STOPs
    with pytest.raises(MonortmError, match="AMBIGUITY"):
        driver.read_control(_control(tmp_path, REC12, "     0.000     0.850"))
    with pytest.raises(MonortmError, match="POSITIVE DVSET"):
        driver.read_control(_control(tmp_path, REC12, "     0.750     0.850"))
    with pytest.raises(MonortmError, match="EXCEEDS LIMIT"):
        driver.read_control(_control(tmp_path, REC12, "     0.100    55.000" + " " * 10 + "    0.0001"))
    assert len(driver.read_control(_control(tmp_path, REC12, "     0.100    55.000" + " " * 10 + "    0.0001"), nwnmx=10**6)["wn"]) == 549001
    with pytest.raises(MonortmError, match="ILNFLG"):
        driver.read_control(_control(tmp_path, REC12, R13_GRID + " " * 44 + "1"))
    with pytest.raises(MonortmError, match="BNDEMI OUTSIDE"):
        driver.read_control(_control(tmp_path, REC12, R13_GRID, rec14="   288.20       1.9"))
    with pytest.raises(MonortmError, match="DERIVATIVES"):
        driver.read_control(_control(tmp_path, rec12(iemit=3), R13_GRID))
    with pytest.raises(MonortmError, match="ERROR READING"):
        driver.read_control(_control(tmp_path, REC12, "     0.2x0     1.200" + " " * 10 + "     0.010"))
    with pytest.raises(MonortmError, match="ERROR OPENING"):
        driver.read_control(str(tmp_path / "missing"))


def test_rdlblinp_continuum_factors_and_scaling_records(tmp_path):
    # ICNTNM 0..5 -> applyCntnmCombo (CntnmFactors.f90:143-186); 6 -> record 1.2a, list-directed
    want = {0: (0,) * 7, 1: (1,) * 7, 2: (0, 1, 1, 1, 1, 1, 1), 3: (1, 0, 1, 1, 1, 1, 1), 4: (0, 0, 1, 1, 1, 1, 1), 5: (1, 1, 1, 1, 1, 1, 0)}
    for ic, w in want.items():
        r12 = rec12(icntnm=ic)
        assert driver.read_control(_control(tmp_path, r12, "     0.750     0.750"))["cntnm"] == tuple(float(x) for x in w)
    r12 = rec12(icntnm=6)
    p = tmp_path / "MONORTM.IN"
    p.write_text("\n".join(["$ x", r12, " 1.1, 0.9 1.0d0", " 0.5 0.25 2 3", "     0.750     0.750", "   288.20       0.9", "%"]) + "\n")
    assert driver.read_control(str(p))["cntnm"] == (1.1, 0.9, 1.0, 0.5, 0.25, 2.0, 3.0)
    # profile scaling: nmol_scal in cols 101-105, then (64a1) and (7e15.7,/,(8e15.7,/))
    r13 = "     0.750     0.750" + " " * 80 + "    2"
    f = _control(tmp_path, REC12, r13, extra=["1P", "  1.2000000E+00  2.5000000E+00"])
    c = driver.read_control(f)
    assert c["nmol_scal"] == 2 and c["hmol_scal"] == "1P" and c["xmol_scal"] == (1.2, 2.5) and c["tmpbnd"] == 288.2
    # exactly 7 factors: format control passes the '/' and swallows one more record before record 1.4
    r13 = "     0.750     0.750" + " " * 80 + "    7"
    f = _control(tmp_path, REC12, r13, extra=["1111111", "".join("%15.7E" % (1 + 0.1 * k) for k in range(7)), "swallowed"])
    c = driver.read_control(f)
    assert c["nmol_scal"] == 7 and c["xmol_scal"][6] == 1.6 and c["tmpbnd"] == 288.2


def test_prof_reader_equals_python_reader_on_reference_fixtures():
    for name in ("MONORTM_PROF.IN_sav", "MONORTM_PROF.IN_liquid_cloud"):
        f = os.path.join(GOLD, name)
        assert driver.count_profiles(f) == 1
        a, b = driver.read_profile(f, 0), profio.read_prof_in(f)[0]
        for k in ("p", "t", "clw", "wbrodl", "tz", "pz", "altz", "wkl"):
            assert np.array_equal(a[k], b[k]), k
        assert (a["nlay"], a["nmol"], a["irt"], a["angle"]) == (19, 22, 3, 0.0)
    with pytest.raises(MonortmError, match="beyond the end"):
        driver.read_profile(f, 1)
    cloud = driver.read_profile(f, 0)
    assert np.array_equal(cloud["clw"][2:5, 0], [0.03, 0.04, 0.03])


def test_prof_reader_two_profiles_mixing_ratios_and_stops(tmp_path):
    src = open(os.path.join(GOLD, "MONORTM_PROF.IN_sav")).read().split("\n")
    while src and not src[-1].strip():
        src.pop()
    up = list(src)
    up[0] = up[0][:65] + "%8.3f" % 180.0 + up[0][73:]            # second profile looks down: IRT=1
    # express CO2 of layer 1 as a mixing ratio: WKL < 1 is multiplied by the dry-air column (monortm.f90:423-483)
    first = driver.read_profile(os.path.join(GOLD, "MONORTM_PROF.IN_sav"), 0)
    w = first["wkl"][:, 0, 0]
    dry = first["wbrodl"][0, 0] + w[2:22].sum()
    vmr = w[1] / (dry + w[1])                                       # so that vmr * dry/(1-vmr) == w[1]
    up[2] = up[2][:15] + "%15.7E" % vmr + up[2][30:]
    f = tmp_path / "MONORTM_PROF.IN"
    f.write_text("\n".join(src + up) + "\n")
    assert driver.count_profiles(str(f)) == 2
    b = driver.read_profile(str(f), 1)
    assert b["irt"] == 1 and b["angle"] == 180.0
    assert abs(b["wkl"][1, 0, 0] / w[1] - 1) < 2e-7                  # 8 printed digits of the mixing ratio
    assert np.array_equal(b["wkl"][0, :, 0], first["wkl"][0, :, 0])
    bad = list(src)
    bad[2] = bad[2][:105] + "%15.7E" % 0.5                          # WBRODL < 1 and non-zero
    f.write_text("\n".join(bad) + "\n")
    with pytest.raises(MonortmError, match="WBRODL"):
        driver.read_profile(str(f), 0)
    bad = list(src)
    bad[1] = bad[1][:3] + "x" + bad[1][4:]
    f.write_text("\n".join(bad) + "\n")
    with pytest.raises(MonortmError, match="ERROR READING"):
        driver.read_profile(str(f), 0)


def test_emiss_reflec_polynomial_and_tables(tmp_path):
    wn = np.array([0.5, 1.0, 2.25])
    c = dict(bndemi=(0.9, 0.01, 0.002), bndrfl=(0.1, 0.0, 0.0))
    e, r = driver.emiss_reflec(c, wn)
    assert np.array_equal(e, 0.9 + 0.01 * wn + 0.002 * wn * wn) and np.array_equal(r, [0.1] * 3)
    # tables (READEM/READRF + LINTCO): V1, V2, DV, NLIM then one value per record
    os.makedirs(tmp_path / "in")
    z = [0.5, 0.6, 0.7, 0.8, 0.9]
    (tmp_path / "in" / "EMISSION").write_text(" 0.000E+00 4.000E+00 1.000E+00         5\n" + "".join("%15.7E\n" % v for v in z))
    (tmp_path / "in" / "REFLECTION").write_text(" 0.000E+00 4.000E+00 1.000E+00         5\n" + "".join("%15.7E\n" % (1 - v) for v in z))
    c = dict(bndemi=(-1.0, 0, 0), bndrfl=(-1.0, 0, 0))
    e, r = driver.emiss_reflec(c, np.array([1.5, 2.25]), dir=str(tmp_path) + "/")
    # NELMNT = INT((VI-V1)/DV) indexes ZEMIS(NELMNT), ZEMIS(NELMNT+1) (1-based) between V1+DV*NELMNT and +1
    assert abs(e[0] - 0.55) < 1e-12 and abs(e[1] - 0.625) < 1e-12 and abs(r[0] - 0.45) < 1e-12
    with pytest.raises(MonortmError, match="EMISFN|REFLFN"):
        driver.emiss_reflec(c, np.array([0.5]), dir=str(tmp_path) + "/")      # NELMNT <= 0 -> STOP


def _fake_results(nwn, nlay, seed=3):
    rng = np.random.default_rng(seed)
    F = dict(order="F")
    o = np.asfortranarray(rng.uniform(1e-4, 0.3, (nwn, nlay)))
    obm = np.zeros((nwn, 39, nlay), **F)
    oc = np.zeros((nwn, 39, nlay), **F)
    for m in (0, 1, 2, 6, 21):
        obm[:, m, :] = rng.uniform(0, 1e-2, (nwn, nlay))
        oc[:, m, :] = rng.uniform(0, 1e-3, (nwn, nlay))
    wkl = np.zeros((39, nlay), **F)
    wkl[[0, 1, 2, 6], :] = rng.uniform(1e15, 1e22, (4, nlay))
    return dict(o=o, obm=obm, oc=oc, wkl=wkl, wbrodl=rng.uniform(1e22, 1e23, nlay), rad=rng.uniform(1e-14, 1e-11, nwn),
                tb=rng.uniform(3, 300, nwn), tmr=rng.uniform(200, 300, nwn), trtot=rng.uniform(0, 1, nwn),
                emiss=np.full(nwn, 0.6), reflc=np.full(nwn, 0.4))


def test_storeout_records_match_the_reference_formats(tmp_path):
    nwn, nlay, nmol = 5, 7, 7
    wn = np.array([0.789344, 0.79828, 1.043027, 1.051763, 5.0])
    d = _fake_results(nwn, nlay)
    d["tb"][0], d["rad"][1], d["trtot"][2] = 2.5e-7, 1.234567891e-101, 0.99999999
    out = tmp_path / "MONORTM.OUT"
    wkl = d["wkl"].copy(order="F")
    driver.storeout(str(out), False, wn, wkl, d["wbrodl"], d["rad"], d["tb"], d["trtot"], 1, d["o"], d["obm"], d["oc"], None,
                    d["tmr"], 2.3456, 0.1, 2.75, d["reflc"], d["emiss"], nmol, 0.0)
    assert np.array_equal(wkl[21], d["wbrodl"])                     # nmol < 22: WKL(22,:) = WBRODL (:600)
    ids = sref.id_mols(d["wkl"], d["wbrodl"], nmol)
    assert ids == [1, 2, 3, 7, 22]
    otot, obm, odx = sref.layer_sums(d["o"], d["obm"], d["oc"])
    want = sref.header(nwn, wn, ids) + sref.rows(1, wn, d["tb"], d["tmr"], d["rad"], d["trtot"], 2.3456, 0.1, 2.75, d["emiss"],
                                                 d["reflc"], 0.0, otot, obm, odx, ids)
    got = out.read_text().split("\n")
    assert got[-1] == "" and got[:-1] == want
    # literal spot checks of the edit descriptors (gfortran behaviour)
    assert got[2] == "NWN :       5" + " " * 101 + " " * 14 + "Molecular Optical Depths -->"
    assert got[3] == ("PROF FREQ(GHz)      BT(K)      TMR(K)  RAD(W/cm2_ster_cm-1)   TRANS     PWV     CLW  TBOUND    EMIS    REFL    ANGLE"
                      "    TOTAL_OD      H2O         CO2          O3          O2          N2        XSEC_OD")
    assert got[4].startswith("    1    23.664    0.00000")               # f10.3 of 0.789344*c/1e9, f11.5 of 2.5e-7
    assert "     1.234567891-101" in got[5]                              # 1pE21.9 with a three-digit exponent
    assert got[6][58:67] == "  1.00000"                                  # f9.5 rounds 0.99999999 up
    assert got[4][67:83] == "  2.3456  0.1000" and got[4][83:107] == "    2.75    0.60    0.40"
    # second profile appends, keeps the molecule columns SAVEd at the first call, prints PROF = 2
    driver.storeout(str(out), True, wn, wkl, d["wbrodl"], d["rad"], d["tb"], d["trtot"], 2, d["o"], d["obm"], d["oc"], None,
                    d["tmr"], 2.3456, 0.1, 2.75, d["reflc"], d["emiss"], nmol, 0.0)
    got2 = out.read_text().split("\n")
    assert len(got2) == 2 * len(got) - 1 and got2[len(got) - 1 + 3] == got[3] and got2[len(got) - 1 + 4].startswith("    2")
    # cm-1 units when wn(1) >= 100, width overflow prints asterisks
    wn2 = wn + 100.0
    d["tb"][0] = 123456.7
    driver.storeout(str(out), False, wn2, wkl, d["wbrodl"], d["rad"], d["tb"], d["trtot"], 1, d["o"], d["obm"], d["oc"], None,
                    d["tmr"], 2.3456, 0.1, 288.2, d["reflc"], d["emiss"], nmol, 180.0)
    g = out.read_text().split("\n")
    assert g[3].startswith("PROF FREQ(cm-1)") and g[4].startswith("    1   100.789***********") and g[4][107:116] == "  180.000"


def test_storeout_layer_optical_depth_files(tmp_path):
    nwn, nlay = 3, 2
    wn = np.array([0.5, 1.0, 1.5])
    d = _fake_results(nwn, nlay)
    d["o"][0, 0], d["o"][1, 0], d["o"][2, 0] = 0.123449, 1.0, 0.0
    driver.storeout(str(tmp_path / "MONORTM.OUT"), False, wn, d["wkl"].copy(order="F"), d["wbrodl"], d["rad"], d["tb"], d["trtot"], 7,
                    d["o"], d["obm"], d["oc"], None, d["tmr"], 1.0, 0.0, 2.75, d["reflc"], d["emiss"], 7, 0.0, iod=1)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["MONORTM.OUT", "ODmono_prf0007_lay0001", "ODmono_prf0007_lay0002"]
    g = (tmp_path / "ODmono_prf0007_lay0001").read_text().split("\n")
    assert g[0] == "NWN :       3" and g[1] == "FREQ(GHz)   LAYER_OD"
    assert g[2] == "    14.990  0.1234E+00" and g[3] == "    29.979  0.1000E+01" and g[4] == "    44.969  0.0000E+00"
