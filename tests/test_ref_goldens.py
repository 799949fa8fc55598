"""The oracle (oracle/monortm_oracle.c) and the product's host helpers against vectors produced by EXECUTING THE REFERENCE'S
OWN FORTRAN TEXT (tests/golden/ref_*.npz, written by tools/gen_ref_goldens.py through the mechanical Fortran->Python
translator tools/f90fn.py -- this image has no Fortran compiler).  These are the reference-derived pins of the oracle:
scalar routines to a few ulp, whole MODM (TIPS_2003 + CONTNM + LINES inside) + CALCTMR + RTM cases to 1e-13.
Nothing here reads /root/reference: the vectors are committed."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

import harness
import ref_cases
from monortm_b200 import api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FN = np.load(os.path.join(GOLD, "ref_functions.npz"))
CASE_FILES = sorted(glob.glob(os.path.join(GOLD, "ref_case_*.npz")))


def _rel(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ok = np.isfinite(b)
    assert np.array_equal(ok, np.isfinite(a))
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), floor))) if ok.any() else 0.0


def _lib():
    lib = harness.oracle_lib()
    D, I, P = C.c_double, C.c_int64, C.c_void_p
    lib.orc_lsf_lortz.restype = D
    lib.orc_lsf_lortz.argtypes = [D] * 8 + [I]
    lib.orc_lsf_sdvoigt.restype = D
    lib.orc_lsf_sdvoigt.argtypes = [D] * 9 + [I, D]
    lib.orc_halfwhm_d.restype = D
    lib.orc_halfwhm_d.argtypes = [I, I, D, D]
    lib.orc_intens.restype = D
    lib.orc_intens.argtypes = [D] * 7
    lib.orc_xlorentz.restype = D
    lib.orc_xlorentz.argtypes = [D]
    lib.orc_xint.restype = None
    lib.orc_xint.argtypes = [D, D, D, P, D, D, D, P, I, I]
    return lib


def test_humlicek_w4_and_sd_humlicek_match_the_reference_text():
    lib = _lib()
    re, im = C.c_double(), C.c_double()
    got = []
    for x, y in zip(FN["w4_x"], FN["w4_y"]):
        lib.orc_w4(float(x), float(y), C.byref(re), C.byref(im))
        got.append((re.value, im.value))
    got = np.array(got)
    scale = np.hypot(FN["w4_re"], FN["w4_im"])
    assert np.max(np.abs(got[:, 0] - FN["w4_re"]) / scale) < 4e-16 and np.max(np.abs(got[:, 1] - FN["w4_im"]) / scale) < 4e-16
    got = []
    for a in zip(FN["sdh_x1"], FN["sdh_y1"], FN["sdh_x2"], FN["sdh_y2"]):
        lib.orc_sd_humlicek(*[float(v) for v in a], C.byref(re), C.byref(im))
        got.append((re.value, im.value))
    got = np.array(got)
    # a difference of two W values: compare to the size of the terms, not of the (cancelling) result
    scale = np.maximum(np.hypot(FN["sdh_re"], FN["sdh_im"]), 1e-3)
    assert np.max(np.abs(got[:, 0] - FN["sdh_re"]) / scale) < 1e-15 and np.max(np.abs(got[:, 1] - FN["sdh_im"]) / scale) < 1e-15


def test_sdvoigt_incl_speed_dependence_matches_the_reference_text():
    lib = _lib()
    err = C.c_int()
    got = []
    for a in zip(FN["sdv_dn"], FN["sdv_al"], FN["sdv_ad"], FN["sdv_sdep"]):
        v = lib.orc_sdvoigt(*[float(x) for x in a], C.byref(err))
        got.append(np.nan if err.value else v)           # the reference STOPs when REAL(v) < 0 (modm.f90:1062)
    assert (FN["sdv_sdep"] > 0).sum() > 100 and np.isfinite(FN["sdv"]).sum() > 400
    assert _rel(got, FN["sdv"]) < 2e-15


def test_line_shape_case_trees_match_the_reference_text():
    lib = _lib()
    A = FN["lsf_args"]
    lor = [lib.orc_lsf_lortz(*[float(v) for v in r[:8]], int(r[9])) for r in A]
    sdv = [lib.orc_lsf_sdvoigt(*[float(v) for v in r[:9]], int(r[9]), float(r[10])) for r in A]
    assert _rel(lor, FN["lsf_lortz"], floor=1e-30) < 4e-16
    # the reference STOPs on a negative SD-Voigt real part; the oracle flags it instead and returns a number
    ok = np.isfinite(FN["lsf_sdvoigt"])
    assert ok.sum() > 0.9 * len(ok)
    assert np.max(np.abs(np.array(sdv)[ok] - FN["lsf_sdvoigt"][ok]) / np.maximum(np.abs(FN["lsf_sdvoigt"][ok]), 1e-30)) < 4e-15
    xl = [lib.orc_xlorentz(float(z)) for z in FN["xl_z"]]
    assert _rel(xl, FN["xl"]) < 3e-16


def test_doppler_width_intensity_radfn_planck_cloud_match_the_reference_text():
    lib = _lib()
    hd = FN["halfwhm_d"]
    assert len(hd) > 80
    got = [lib.orc_halfwhm_d(int(m), int(i), float(x), float(t)) for m, i, x, t, _ in hd]
    assert _rel(got, hd[:, 4]) < 3e-16
    radct = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16
    got = [lib.orc_intens(float(t), float(s), float(e), radct, 296.0, float(x), float(q))
           for t, s, e, x, q in zip(FN["int_t"], FN["int_s0"], FN["int_es"], FN["int_xnu"], FN["int_q"])]
    assert _rel(got, FN["intens"]) < 4e-16
    got = [lib.orc_radfn(float(a), float(b)) for a, b in zip(FN["radfn_vi"], FN["radfn_xkt"])]
    assert _rel(got, FN["radfn"]) < 3e-16
    got = [lib.orc_bb_fn(float(a), float(b)) for a, b in zip(FN["bb_v"], FN["bb_fbeta"])]
    assert _rel(got, FN["bb"]) < 3e-16
    # CloudOptProp carries d0 literals: binary128 in the oracle (as in the parity build), binary64 in the translator
    got = [lib.orc_odclw(float(a), float(b), float(c)) for a, b, c in zip(FN["clw_wn"], FN["clw_t"], FN["clw_amt"])]
    assert _rel(got, FN["odclw"]) < 1e-13


def test_xint_matches_the_reference_text():
    lib = _lib()
    a = np.ascontiguousarray(FN["xint_a"])
    r3 = np.zeros(64)
    lib.orc_xint(-20.0, 370.0, 10.0, harness._p(a), 1.0, -3.0, 1.0, harness._p(r3), 1, 64)
    assert np.max(np.abs(r3 - FN["xint_grid"])) < 1e-15
    for vft, want in zip(FN["xint_pts"], FN["xint_single"]):
        one = np.zeros(1)
        lib.orc_xint(-3.0, 60.0, 1.0, harness._p(r3), 1.0, float(vft), 1.0, harness._p(one), 1, 1)
        assert abs(one[0] - want) < 1e-15


def test_host_tips_2003_matches_the_reference_text_for_all_39_molecules():
    for t, want in zip(FN["tips_t"], FN["tips_scor"]):
        got = api.tips_2003(39, float(t))
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 2e-15, t
        orc = harness.oracle_tips_2003(39, float(t))          # the oracle's own TIPS (CPU legs of bench.py)
        assert np.max(np.abs(orc - want) / np.maximum(np.abs(want), 1e-300)) < 2e-15, t
    with pytest.raises(RuntimeError):
        harness.oracle_tips_2003(22, 60.0)
    assert FN["tips_scor"][2, 33, 0] == 1.0 and FN["tips_scor"][2, 38, 0] == 1.0      # O and CH3OH: scor = 1


@pytest.mark.parametrize("path", CASE_FILES, ids=[os.path.basename(p)[9:-4] for p in CASE_FILES])
def test_oracle_reproduces_the_reference_text_end_to_end(path):
    """MODM (TIPS_2003, CONTNM, LINES, ODCLW) + CALCTMR + RTM: oracle vs the executed reference text."""
    g = np.load(path)
    case = ref_cases.build_case(json.loads(str(g["spec"])))
    assert ref_cases.inputs_digest(case) == str(g["digest"]), "the regenerated inputs differ from the ones the golden was made with"
    assert int(g["oob_reads"]) == 0
    # TIPS: what the harness feeds both sides is what the reference's TIPS_2003 gives
    assert np.max(np.abs(case["scor"][:, :, :, 0] - g["scor"]) / np.maximum(np.abs(g["scor"]), 1e-300)) < 2e-15
    ref = harness.run_oracle(case)
    scale = np.abs(g["o"])[:, None, :]
    assert harness.rel_diff(ref["o"], g["o"]) < 1e-13
    assert np.max(np.abs(ref["o_by_mol"] - g["o_by_mol"]) / scale) < 1e-13
    assert np.max(np.abs(ref["oc"] - g["oc"]) / scale) < 1e-13
    assert np.max(np.abs(ref["o_clw"] - g["o_clw"]) / np.abs(g["o"])) < 1e-13
    for k in ("rad", "rup", "rdn", "trtot"):
        assert harness.rel_diff(ref[k], g[k], floor=1e-300) < 1e-12, k
    assert np.max(np.abs(ref["tb"] - g["tb"])) < 1e-9 and np.max(np.abs(ref["tmr"] - g["tmr"])) < 1e-9
    assert ref["tmpsfc"] == float(g["tmpsfc"])


def test_the_goldens_cover_the_branches_they_claim():
    g = np.load(os.path.join(GOLD, "ref_case_voigt_zone_cov.npz"))
    case = ref_cases.build_case(json.loads(str(g["spec"])))
    br = harness.run_oracle(case)["branches"]
    assert br["sdep"] > 100 and br["co2"] > 50 and br["co2_lc1"] > 30 and br["generic_lc"] > 50 and br["o2_lc"] > 10 and br["neg_res"] > 50


def test_cross_sections_match_the_reference_text(tmp_path):
    """MONORTM_XSEC_SUB + convolve (monortm_sub.F90:1540-1834): oracle vs the executed reference text on the seeded synthetic
    cross-section set (two molecules, three regions, 1-4 temperature files, a TORR file; Lorentz convolution and the
    interpolation shortcut; table edges; frequencies outside every region)."""
    g = np.load(os.path.join(GOLD, "ref_xsec_synth.npz"))
    spec = json.loads(str(g["spec"]))
    regs, wn, p, t, xamnt = harness.xsec_case(spec, str(tmp_path))
    assert [(r["ixmol"], r["v1fx"], len(r["files"])) for r in regs] == [(0, 2.0, 4), (0, 40.0, 2), (1, 18.0, 1)]   # F11 300-310 skipped
    od = harness.oracle_xsec(regs, wn, p, t, xamnt)
    assert np.array_equal(od == 0, g["odxsec"] == 0) and (g["odxsec"] != 0).sum() >= 60
    assert _rel(od, g["odxsec"]) < 1e-14
    for (pd, tt, pp), ref in zip(g["conv_args"], g["conv"]):
        got = harness.oracle_convolve(g["conv_tab"], 5.0, 6.0, 0.0025, pd, 1.1e-5, tt, pp, g["conv_wn"])
        assert np.array_equal(got == 0, ref == 0) and _rel(got, ref) < 1e-14
