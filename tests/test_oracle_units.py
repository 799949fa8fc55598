"""CPU tests that pin the oracle (oracle/monortm_oracle.c) with analytic identities and independent
implementations.  The reference ships no golden outputs, so these identities --
plus the committed self-generated regression vectors in tests/golden -- are what anchors it."""
import ctypes as C

import numpy as np
import pytest
from scipy.special import wofz

import harness

RADCN1 = 1.191042722E-12
RADCN2 = 1.4387752


def _w4(x, y):
    lib = harness.oracle_lib()
    re, im = C.c_double(), C.c_double()
    lib.orc_w4(x, y, C.byref(re), C.byref(im))
    return complex(re.value, im.value)


def test_w4_matches_faddeeva_within_stated_accuracy():
    # "MAXIMUM RELATIVE ERROR OF BOTH REAL AND IMAGINARY PARTS IS <1*10**(-4)", modm.f90:1094
    rng = np.random.default_rng(1)
    pts = [(0.0, 1e-3), (0.3, 0.05), (1.0, 0.1), (3.0, 0.5), (5.4, 0.2), (6.0, 0.01), (10.0, 4.0), (20.0, 0.5),
           (0.5, 20.0), (2.0, 0.15), (4.0, 0.6)]
    pts += [(float(x), float(y)) for x, y in zip(rng.uniform(0, 25, 200), 10 ** rng.uniform(-3, 1.3, 200))]
    for x, y in pts:
        w = _w4(x, y)
        ref = wofz(complex(x, y))
        assert abs(w.real - ref.real) <= 1.2e-4 * abs(ref.real) + 1e-12, (x, y, w, ref)


def test_w4_region_boundaries_follow_reference():
    # regions switch at |x|+y = 15 and 5.5 and at y = 0.195|x|-0.176 (modm.f90:1105,1109,1114)
    for x, y in ((14.9999, 0.0001), (15.0, 0.0), (5.4999, 0.0001), (5.5, 0.0), (3.0, 0.195 * 3 - 0.176 - 1e-9)):
        w = _w4(x, y)
        assert np.isfinite(w.real) and np.isfinite(w.imag)
    # region I formula: t*.5641896/(.5+t*t) with t = y - ix
    x, y = 20.0, 1.0
    t = complex(y, -x)
    assert abs(_w4(x, y) - t * .5641896 / (.5 + t * t)) < 1e-15


def test_sd_humlicek_reduces_to_w4_difference():
    lib = harness.oracle_lib()
    re, im = C.c_double(), C.c_double()
    for (x1, y1, x2, y2) in ((0.2, 3.0, 0.2, 9.0), (1.0, 0.8, 1.0, 2.0), (7.0, 9.0, 7.0, 12.0)):
        lib.orc_sd_humlicek(x1, y1, x2, y2, C.byref(re), C.byref(im))
        d = wofz(complex(x1, y1)) - wofz(complex(x2, y2))
        assert abs(re.value - d.real) < 3e-4 * max(abs(d.real), 1e-3)


def test_sdvoigt_limits():
    lib = harness.oracle_lib()
    err = C.c_int(0)
    # alphad = 0 -> zeta == 1 -> pure Lorentz shortcut (modm.f90:1017)
    v = lib.orc_sdvoigt(0.03, 0.05, 0.0, 0.0, C.byref(err))
    assert abs(v - 0.05 / (3.1415926535898 * (0.05 ** 2 + 0.03 ** 2))) < 1e-15
    # Voigt ~ Gaussian core when alphal << alphad: peak = sqrt(ln2/pi)/alphad * Re w(i y)
    ad, al = 1e-3, 1e-6
    v = lib.orc_sdvoigt(0.0, al, ad, 0.0, C.byref(err))
    y = np.sqrt(np.log(2.0)) * al / ad
    assert abs(v - np.sqrt(np.log(2) / 3.1415926535898) / ad * wofz(1j * y).real) < 2e-4 * v
    # area of the Voigt profile ~ 1
    d = np.linspace(-0.05, 0.05, 20001)
    prof = np.array([lib.orc_sdvoigt(float(x), 2e-4, 1e-3, 0.0, C.byref(err)) for x in d])
    assert abs(np.trapezoid(prof, d) - 1.0) < 5e-3
    # speed-dependent branch stays positive and close to the plain Voigt for small SDEP effects
    vs = lib.orc_sdvoigt(2e-4, 2e-4, 1e-3, 0.08, C.byref(err))
    v0 = lib.orc_sdvoigt(2e-4, 2e-4, 1e-3, 0.0, C.byref(err))
    assert err.value == 0 and vs > 0 and abs(vs - v0) < 0.2 * v0


def test_radfn_three_regimes():
    lib = harness.oracle_lib()
    xkt = 250.0 / RADCN2
    assert lib.orc_radfn(1.0, xkt) == 0.5 * (1.0 / xkt) * 1.0                  # x <= 0.01
    v = 50.0
    x = v / xkt
    assert abs(lib.orc_radfn(v, xkt) - v * np.tanh(x / 2)) < 1e-13 * v          # (1-e)/(1+e) == tanh(x/2)
    assert lib.orc_radfn(5000.0, xkt) == 5000.0                                 # x > 10
    assert lib.orc_radfn(3.0, 0.0) == 3.0                                       # XKT <= 0


def test_planck_roundtrip_and_isothermal_layer_identity():
    # one isothermal layer, levels at the layer temperature: bb == bba so the Pade term collapses and
    # RDN = B(T)(1-exp(-tau)) (RTMmono.f90:215-216); TB inverts Planck (:150-151)
    lib = harness.oracle_lib()
    wn = np.array([0.7, 2.0, 6.1, 30.0])
    T = 271.3
    tau = np.array([1e-3, 0.4, 2.5, 30.0])
    o = np.asfortranarray(tau.reshape(4, 1))
    r = harness.oracle_rtm(1, 3, wn, [T], [T, T], o, 300.0, np.zeros(4), np.ones(4))
    B = np.array([lib.orc_bb_fn(float(v), RADCN2 / T) for v in wn])
    assert np.allclose(r["rdn"], B * (1 - np.exp(-tau)), rtol=1e-14)
    assert np.allclose(r["trtot"], np.exp(-tau), rtol=1e-15)
    assert r["tmpsfc"] == 2.75                                                 # RTMmono.f90:122
    cosmic = np.array([lib.orc_bb_fn(float(v), RADCN2 / 2.75) for v in wn])
    assert np.allclose(r["rad"], r["rdn"] + np.exp(-tau) * cosmic, rtol=1e-14)
    # opaque layer: TB -> T
    assert abs(r["tb"][3] - T) < 1e-9
    # mean radiating temperature of an isothermal atmosphere is T
    tmr = harness.oracle_calctmr(wn, [T], [T, T], o)
    assert np.allclose(tmr, T, atol=1e-9)


def test_rtm_upwelling_limits():
    wn = np.array([1.0, 5.0])
    o = np.asfortranarray(np.full((2, 3), 1e-12))
    t, tz = [250., 240., 230.], [255., 245., 235., 225.]
    r = harness.oracle_rtm(1, 1, wn, t, tz, o, 290.0, np.zeros(2), np.ones(2))
    # transparent atmosphere, emissivity 1: TB = surface temperature
    assert np.allclose(r["tb"], 290.0, atol=1e-6)
    assert r["tmpsfc"] == 290.0
    with pytest.raises(RuntimeError):
        harness.oracle_rtm(1, 1, wn, t, tz, o, 290.0, np.zeros(2), np.ones(2), idu=0)   # RTMmono.f90:173 STOP


def test_cloud_liquid_od():
    lib = harness.oracle_lib()
    assert lib.orc_odclw(1.0, 280.0, 0.0) == 0.0
    a1 = lib.orc_odclw(1.0, 280.0, 0.1)
    assert a1 > 0 and abs(lib.orc_odclw(1.0, 280.0, 0.2) - 2 * a1) < 1e-15
    # mass absorption coefficient at 31.4 GHz, 0 C is roughly 0.1 m2/kg per (kg/m2)  (Turner et al.)
    a = lib.orc_odclw(31.4 / 29.9792458, 273.15, 1.0)
    assert 0.05 < a < 0.3
    # absorption grows with frequency in the microwave and with supercooling at 31 GHz
    assert lib.orc_odclw(3.0, 280.0, 1.0) > lib.orc_odclw(1.0, 280.0, 1.0)
    assert lib.orc_odclw(1.0, 253.15, 1.0) > lib.orc_odclw(1.0, 293.15, 1.0)


def _contnm(im, factors, pave=800., tave=270., v1=0.2, v2=30.):
    lib = harness.oracle_lib()
    wk = np.zeros(60)
    wk[0], wk[1], wk[6], wk[21] = 3e22, 6e20, 3.5e23, 1.3e24
    v1abs = float(int(v1)) - 3.
    v2abs = float(int(v2 + 3.5))
    npt = int((v2abs - v1abs) + 1.5)
    ab = np.zeros(npt + 2)
    f = np.array(factors, dtype=np.float64)
    rc = lib.orc_contnm_one(im, f.ctypes.data_as(C.c_void_p), pave, tave, wk.ctypes.data_as(C.c_void_p), 1.3e24, 22,
                            v1, v2, v1abs, v2abs, npt, ab.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.orc_last_error()
    return ab[:npt]


def test_continuum_scaling_and_selectors():
    one = (1.,) * 7
    h = _contnm(1, one)
    assert np.all(h[3:-3] > 0)
    # self + foreign add (modm.f90:213 selects both for H2O); each scales linearly with its factor
    hs = _contnm(1, (1., 0., 1., 1., 1., 1., 1.))
    hf = _contnm(1, (0., 1., 1., 1., 1., 1., 1.))
    assert np.allclose(hs + hf, h, rtol=1e-13)
    assert np.allclose(_contnm(1, (2., 0., 1., 1., 1., 1., 1.)), 2 * hs, rtol=1e-13)
    # species without a microwave continuum give zero (gates contnm.f90:536,657)
    assert np.all(_contnm(3, one) == 0) and np.all(_contnm(7, one) == 0) and np.all(_contnm(99, one) == 0)
    n2 = _contnm(22, one)
    co2 = _contnm(2, one)
    assert np.all(n2[3:-3] > 0) and np.all(co2[3:-3] > 0)
    # N2 CIA scales with density squared: doubling pressure quadruples amagat*rho... tau_fac ~ amagat
    assert np.allclose(_contnm(22, one, pave=400.) * 2, n2, rtol=1e-12)
    # a request beyond the microwave runs the branches above 820 cm-1 (round 1 refused it): the 100-900 cm-1 H2O continuum is
    # finite and positive, Rayleigh (selector 99) switches on at V2 >= 820 (contnm.f90:1107), O3 has nothing below 8920
    lib = harness.oracle_lib()
    wk = np.zeros(60)
    wk[0], wk[6] = 1e22, 2e23
    f = np.ones(7)
    for im, expect in ((1, True), (99, True), (3, False)):
        ab = np.zeros(2000)
        rc = lib.orc_contnm_one(im, f.ctypes.data_as(C.c_void_p), 800., 270., wk.ctypes.data_as(C.c_void_p), 1e24, 22,
                                100., 900., 97., 904., 808, ab.ctypes.data_as(C.c_void_p))
        assert rc == 0 and np.all(np.isfinite(ab)) and (np.all(ab[3:805] > 0) if expect else np.all(ab == 0))


def test_modm_selection_counts_and_totals():
    # every O2 line passes the cutoff (modm.f90:384 has I.NE.7), others only inside 25 cm-1
    case = harness.make_case(n_filler=256, nlay=4, wn=np.array([0.5, 10.0, 40.0, 54.0]), irt=3)
    ref = harness.run_oracle(case)
    ls = case["ls"]
    pr = case["prof"]
    # logical lines per molecule (coupling records are not lines)
    for iw, w in enumerate(case["wn"]):
        for k in range(4):
            ratio = (pr["p"][k, 0] / (1.3806503E-16 * pr["t"][k, 0])) * 1e3 / ((1013.25 / (1.3806503E-16 * 296.)) * 1e3)
            cnt = 0
            for i in range(39):
                n = int(ls.nblm[i])
                if n == 0 or pr["wkl"][i, k, 0] == 0:
                    continue
                j = 0
                while j < n:
                    xg = ls.xg[i, j]
                    xnu = ls.xnu0[i, j] + ls.deltnu[i, j] * ratio
                    if i + 1 == 7 or not abs(w - xnu) > 25.0:
                        cnt += 1
                    j += 2 if xg in (-1., -3., -5.) else 1
            assert ref["sel_count"][iw, k] == cnt
    # O = sum of parts (modm.f90:265-269)
    tot = ref["o_by_mol"].sum(axis=1) + ref["oc"].sum(axis=1) + ref["o_clw"] + ref["odxsec"]
    assert np.allclose(tot, ref["o"], rtol=1e-13)
    assert np.all(ref["o"] > 0)


def test_modm_zero_amount_molecule_is_skipped_and_line_coupling_scales():
    wn = np.linspace(1.6, 2.4, 9)
    case = harness.make_case(n_filler=128, nlay=3, wn=wn, irt=3)
    base = harness.run_oracle(case)
    case0 = dict(case)
    prof = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in case["prof"].items()}
    prof["wkl"][6] = 0.0                       # no O2: W_SPECIES == 0 -> skipped (modm.f90:318-321)
    case0["prof"] = prof
    no_o2 = harness.run_oracle(case0)
    assert np.all(no_o2["o_by_mol"][:, 6, :] == 0)
    assert np.all(no_o2["sel_count"] < base["sel_count"])
    # SCLCPL scales the first-order coupling of the O2 band (modm.f90:358-361)
    c2 = dict(case)
    c2["sclcpl"] = 0.0
    nolc = harness.run_oracle(c2)
    assert not np.allclose(nolc["o_by_mol"][:, 6, 0], base["o_by_mol"][:, 6, 0], rtol=1e-6)


def test_oracle_o0_and_o2_builds_agree():
    case = harness.make_case(n_filler=128, nlay=6, wn=np.linspace(0.3, 20.0, 24), irt=1, clw=True)
    a = harness.run_oracle(case, opt="O0")
    b = harness.run_oracle(case, opt="O2")
    assert np.array_equal(a["sel_hash"], b["sel_hash"])
    assert harness.rel_diff(a["o"], b["o"]) < 1e-13
    assert np.max(np.abs(a["tb"] - b["tb"])) < 1e-9


def test_isolated_line_matches_textbook_formula():
    """One isolated N2O-like line, one layer: the oracle's line optical depth against the textbook expression written
    from scratch in numpy -- HITRAN intensity scaling S(T) = S0 Q0/Q(T) exp(-c2 E (1/T - 1/T0)) (1-e^{-c2 v0/T})/(1-e^{-c2 v0/T0}),
    the radiation-field form k(v) = W S(T) [v tanh(c2 v/2T)] / [v0 tanh(c2 v0/2T)] (L(v-v0) + L(v+v0)) with Lorentzians cut at
    25 cm-1 and their value there subtracted (the negative-frequency partner only while v+v0 <= 25), the half width
    mixed from air and self broadening with the number-density ratio, the centre shifted by delta times that ratio.
    Independent of the oracle's code path: anchors units, strength conversion, radiation term, shape and cutoff."""
    import os
    import tempfile
    from monortm_b200 import api, linefile, synth

    c2 = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16
    t0, p0 = 296.0, 1013.25
    v0, s0, g_air, g_self, epp, xexp, delta = 12.3456, 3.0e-23, 0.08, 0.11, 350.0, 0.7, -2.0e-3
    rec = synth._line(v0, s0, g_air, g_self, epp, xexp, delta, 4, 1)
    recs = np.zeros(1, synth.REC_DTYPE)
    recs[0] = rec
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, 0.0, 55.0)
    finally:
        os.unlink(path)
    # the file keeps float32 parameters: take them back from the store so both sides see the same numbers
    g_air, g_self = float(ls.alpf[3, 0]), float(ls.alps[3, 0])
    epp, xexp, delta = float(ls.e[3, 0]), float(ls.x[3, 0]), float(ls.deltnu[3, 0])
    s0 = float(ls.s0[3, 0]) * (v0 * (1.0 - np.exp(-c2 * v0 / t0)))          # back from the LBLRTM form on the file
    wn = np.array([0.5, 5.0, 12.0, 12.3, 12.3456, 12.4, 20.0, 30.0, 37.3, 37.4, 40.0])
    for t, p in ((296.0, 500.0), (250.0, 800.0), (296.0, 1013.25)):
        w_n2o, w_n2 = 3.0e17, 1.0e23
        wkl = np.zeros((39, 1), order="F")
        wkl[3, 0] = w_n2o
        scor = api.scor_for_layers(7, np.array([[t]]))[:, :, :, 0]
        m = harness.oracle_modm(ls, wn, 0.0, np.array([p]), np.array([t]), np.array([0.0]), 7, wkl, np.array([w_n2]), scor)
        got = m["o_by_mol"][:, 3, 0]
        qratio = scor[3, 0, 0]                                                 # Q(296)/Q(T) of N2O isotopologue 1 (TIPS)
        rho = (p / t) / (p0 / t0)                                              # number-density ratio
        x_self = w_n2o / (w_n2o + w_n2)
        vc = v0 + delta * rho
        gam = (g_air * (1.0 - x_self) + g_self * x_self) * rho * (t / t0) ** xexp
        s_t = s0 * qratio * np.exp(-c2 * epp * (1.0 / t - 1.0 / t0)) * (1 - np.exp(-c2 * vc / t)) / (1 - np.exp(-c2 * vc / t0))

        def lor(d):
            return gam / (np.pi * (d * d + gam * gam))
        shape = np.where(np.abs(wn - vc) <= 25.0, lor(wn - vc) - lor(25.0) + np.where(wn + vc <= 25.0, lor(wn + vc) - lor(25.0), 0.0), 0.0)
        want = w_n2o * s_t * (wn * np.tanh(c2 * wn / (2 * t))) / (vc * np.tanh(c2 * vc / (2 * t))) * shape
        assert np.all((want == 0) == (got == 0))
        nz = want != 0
        assert np.max(np.abs(got[nz] / want[nz] - 1.0)) < 2e-9, (t, p)     # 13-digit pi in the reference: 1e-13; float32 S: exact here


def test_isolated_line_voigt_branch_matches_scipy_voigt_profile():
    """The same isolated line at 0.3 hPa (Doppler width ~ Lorentz width): inside 100 Doppler half-widths the reference takes
    the Voigt branch (modm.f90:427-431).  Against scipy's Voigt profile with the textbook Doppler half-width
    v0/c sqrt(2 ln2 kT/m), within the 1e-4 the Humlicek routine claims (modm.f90:1094).  Pins the Doppler width, the
    isotopologue mass table and the normalisation of the Voigt branch independently of the oracle's code."""
    import os
    import tempfile
    from scipy.special import voigt_profile
    from monortm_b200 import api, linefile, synth

    c2 = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16
    kb, clight, amu = 1.3806503E-16, 2.99792458E+10, 1.0 / 6.02214199E+23
    t0, p0 = 296.0, 1013.25
    v0 = 25.4321
    recs = np.zeros(1, synth.REC_DTYPE)
    recs[0] = synth._line(v0, 2.0e-22, 0.07, 0.10, 120.0, 0.7, 0.0, 4, 1)            # N2O 446, 44.0 amu
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, 0.0, 55.0)
    finally:
        os.unlink(path)
    g_air, g_self, epp, xexp = float(ls.alpf[3, 0]), float(ls.alps[3, 0]), float(ls.e[3, 0]), float(ls.x[3, 0])
    s0 = float(ls.s0[3, 0]) * (v0 * (1.0 - np.exp(-c2 * v0 / t0)))
    t, p = 250.0, 0.3
    mass = 44.0128                                                                  # 14N2 16O, g/mol
    a_d = v0 / clight * np.sqrt(2 * np.log(2.0) * kb * t / (mass * amu))
    wn = v0 + a_d * np.array([-60.0, -20.0, -5.0, -1.5, -0.4, 0.0, 0.3, 1.0, 3.0, 12.0, 45.0, 90.0])
    w_n2o, w_n2 = 3.0e15, 1.0e21
    wkl = np.zeros((39, 1), order="F")
    wkl[3, 0] = w_n2o
    scor = api.scor_for_layers(7, np.array([[t]]))[:, :, :, 0]
    m = harness.oracle_modm(ls, wn, 0.0, np.array([p]), np.array([t]), np.array([0.0]), 7, wkl, np.array([w_n2]), scor)
    assert m["n_voigt"] == len(wn)                                                   # every pair took the Voigt branch
    got = m["o_by_mol"][:, 3, 0]
    rho = (p / t) / (p0 / t0)
    x_self = w_n2o / (w_n2o + w_n2)
    gam = (g_air * (1.0 - x_self) + g_self * x_self) * rho * (t / t0) ** xexp
    assert 0.05 < gam / a_d < 20.0                                                   # a genuine Voigt regime
    s_t = s0 * scor[3, 0, 0] * np.exp(-c2 * epp * (1.0 / t - 1.0 / t0)) * (1 - np.exp(-c2 * v0 / t)) / (1 - np.exp(-c2 * v0 / t0))
    sigma = a_d / np.sqrt(2 * np.log(2.0))
    shape = voigt_profile(wn - v0, sigma, gam) - voigt_profile(25.0, sigma, gam)
    want = w_n2o * s_t * (wn * np.tanh(c2 * wn / (2 * t))) / (v0 * np.tanh(c2 * v0 / (2 * t))) * shape
    assert np.max(np.abs(got / want - 1.0)) < 2e-4


def test_rtm_matches_exact_linear_in_tau_layers():
    """RAD_UP_DN / RTM (RTMmono.f90:13-221) against an independent numpy integration of the same atmosphere in which
    the Planck function varies linearly with optical depth inside each layer (layer mean = B(T), value at the boundary
    facing the observer = B(TZ)) and every layer is integrated exactly.  The reference's Pade weight 0.193 tau + 0.013 tau^2
    approximates exactly that model, so radiances agree to a few 1e-4 relative (0.05 K): a structural pin of the
    direction of the loops, the boundary terms, the surface reflection and the cosmic background."""
    rng = np.random.default_rng(3)
    nwn, nlay = 40, 24
    wn = np.linspace(0.7, 30.0, nwn)
    tz = np.linspace(291.0, 215.0, nlay + 1)
    t = 0.5 * (tz[:-1] + tz[1:]) + rng.normal(0, 0.3, nlay)
    o = np.asfortranarray(10.0 ** rng.uniform(-3.5, 0.3, (nwn, nlay)))
    em, rf, tsfc = np.full(nwn, 0.8), np.full(nwn, 0.2), 293.0

    def planck(v, temp):
        return RADCN1 * v ** 3 / np.expm1(RADCN2 * v / temp)

    def layer_emission(tau, b_mean, b_near):
        # integral of B(tau') exp(-tau') dtau' over the layer, B linear in tau' from b_near at the observer's side with mean b_mean
        return b_near * (-np.expm1(-tau)) + (b_mean - b_near) * (2.0 / tau) * (1.0 - (1.0 + tau) * np.exp(-tau))

    rup, rdn = np.zeros(nwn), np.zeros(nwn)
    for iw in range(nwn):
        tau = o[iw]
        above = np.concatenate([np.cumsum(tau[::-1])[::-1][1:], [0.0]])      # optical depth between layer l and space
        below = np.concatenate([[0.0], np.cumsum(tau)[:-1]])                 # ... between layer l and the surface
        for l in range(nlay):
            rup[iw] += np.exp(-above[l]) * layer_emission(tau[l], planck(wn[iw], t[l]), planck(wn[iw], tz[l + 1]))
            rdn[iw] += np.exp(-below[l]) * layer_emission(tau[l], planck(wn[iw], t[l]), planck(wn[iw], tz[l]))
    trtot = np.exp(-o.sum(axis=1))
    cosmic = planck(wn, 2.75)
    up = rup + trtot * (em * planck(wn, tsfc) + rf * (rdn + trtot * cosmic))
    dn = rdn + trtot * cosmic
    r1 = harness.oracle_rtm(1, 1, wn, t, tz, o, tsfc, rf, em)
    r3 = harness.oracle_rtm(1, 3, wn, t, tz, o, tsfc, rf, em)
    assert np.max(np.abs(r1["rup"] / rup - 1)) < 1e-3 and np.max(np.abs(r1["rdn"] / rdn - 1)) < 1e-3
    assert np.max(np.abs(r1["rad"] / up - 1)) < 1e-3 and np.max(np.abs(r3["rad"] / dn - 1)) < 1e-3
    assert np.allclose(r1["trtot"], trtot, rtol=1e-12)
    tb_up = RADCN2 * wn / np.log(RADCN1 * wn ** 3 / up + 1.0)
    assert np.max(np.abs(r1["tb"] - tb_up)) < 0.1


def test_o2_line_with_first_order_mixing_matches_textbook_formula():
    """One O2 line with an IFLG=1 coefficient record (Y and G equal at the four tabulated temperatures, so the temperature
    interpolation is trivial): against the first-order line-mixing expression written from scratch,
    k(v) = W S~ v tanh(c2 v/2T) (1/pi) { [g(1+G p^2) + Y p (v-v0)]/((v-v0)^2+g^2) + [g(1+G p^2) - Y p (v+v0)]/((v+v0)^2+g^2) },
    p = P/P0, no cutoff and no pedestal for O2 (modm.f90:384, :757-786).  Pins the sign convention of the mixing term on the
    negative-frequency resonance, the pressure scalings and the record layout of the coefficients."""
    import os
    import tempfile
    from monortm_b200 import api, linefile, synth

    c2 = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16
    t0, p0 = 296.0, 1013.25
    v0, yv, gv = 2.0843, 0.35, -0.02
    recs = np.zeros(2, synth.REC_DTYPE)
    recs[0] = synth._line(v0, 1.0e-25, 0.045, 0.047, 2.08, 0.8, 0.0, 7, 1, iflg=1)
    recs[1] = linefile.coupling_record([yv] * 4, [gv] * 4, 1)
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, 0.0, 55.0)
    finally:
        os.unlink(path)
    g_air, g_self, epp, xexp = float(ls.alpf[6, 0]), float(ls.alps[6, 0]), float(ls.e[6, 0]), float(ls.x[6, 0])
    yv, gv = float(np.float32(yv)), float(np.float32(gv))
    s0 = float(ls.s0[6, 0]) * (v0 * (1.0 - np.exp(-c2 * v0 / t0)))
    wn = np.array([0.2, 1.5, 2.0, 2.0843, 2.2, 3.9, 10.0, 30.0, 54.0])         # the last ones lie beyond 25 cm-1: O2 is never cut
    for t, p in ((296.0, 1013.25), (296.0, 300.0), (270.0, 700.0)):
        w_o2, w_n2 = 4.0e23, 1.5e24
        wkl = np.zeros((39, 1), order="F")
        wkl[6, 0] = w_o2
        scor = api.scor_for_layers(7, np.array([[t]]))[:, :, :, 0]
        m = harness.oracle_modm(ls, wn, 0.0, np.array([p]), np.array([t]), np.array([0.0]), 7, wkl, np.array([w_n2]), scor)
        got = m["o_by_mol"][:, 6, 0]
        rho, pr = (p / t) / (p0 / t0), p / p0
        x_self = w_o2 / (w_o2 + w_n2)
        # lnfl_mod.f90:98-113: the O2 air width on the file is corrected to a foreign width with the 0.21 mixing ratio
        g_for = (float(np.float32(0.045)) - 0.21 * g_self) / (1.0 - 0.21)
        assert abs(g_for - g_air) < 1e-7
        gam = (g_air * (1.0 - x_self) + g_self * x_self) * rho * (t / t0) ** xexp
        s_t = s0 * scor[6, 0, 0] * np.exp(-c2 * epp * (1.0 / t - 1.0 / t0)) * (1 - np.exp(-c2 * v0 / t)) / (1 - np.exp(-c2 * v0 / t0))
        dm, dp = wn - v0, wn + v0
        shape = ((gam * (1 + gv * pr ** 2) + yv * pr * dm) / (dm ** 2 + gam ** 2) + (gam * (1 + gv * pr ** 2) - yv * pr * dp) / (dp ** 2 + gam ** 2)) / np.pi
        want = w_o2 * s_t * (wn * np.tanh(c2 * wn / (2 * t))) / (v0 * np.tanh(c2 * v0 / (2 * t))) * shape
        assert np.max(np.abs(got / want - 1.0)) < 5e-9, (t, p, got / want)


def _table(name):
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "monortm_b200", "csrc", "tables", "mtckd_tables.inc")
    txt = open(path).read()
    m = re.search(r"%s\[\d+\]\s*=\s*\{(.*?)\};" % name, txt, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()])


def test_h2o_self_continuum_at_table_nodes_matches_the_mt_ckd_expression():
    """MT_CKD water-vapour self continuum (contnm.f90:300-371) at frequencies that are nodes of the 10 cm-1 coefficient
    grid, where both interpolation stages return the tabulated value exactly: optical depth =
    W_h2o * C_s(296) (C_s(260)/C_s(296))^((T-296)/(260-296)) * x_h2o (P/1013)(296/T) 1e-20 * v tanh(c2 v/2T), written
    from the tables in numpy.  Pins the units, the density factor, the temperature exponent, the grid alignment and
    the radiation term of the continuum path."""
    from monortm_b200 import api, linefile, synth
    import os
    import tempfile

    s296, s260 = _table("MTCKD_SH2O_296"), _table("MTCKD_SH2O_260")
    assert len(s296) == 2003 and len(s260) == 2003
    recs = np.zeros(1, synth.REC_DTYPE)
    recs[0] = synth._line(70.0, 1.0e-30, 0.05, 0.05, 100.0, 0.7, 0.0, 4, 1)        # a line far from the frequencies used
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, 0.0, 80.0)
    finally:
        os.unlink(path)
    wn = np.array([10.0, 20.0, 30.0])
    for t, p in ((270.0, 800.0), (296.0, 1013.0), (245.0, 400.0)):
        w_h2o, w_o2, w_n2 = 2.0e22, 4.0e23, 1.5e24
        wkl = np.zeros((39, 1), order="F")
        wkl[0, 0], wkl[6, 0] = w_h2o, w_o2
        scor = api.scor_for_layers(7, np.array([[t]]))[:, :, :, 0]
        m = harness.oracle_modm(ls, wn, 0.0, np.array([p]), np.array([t]), np.array([0.0]), 7, wkl, np.array([w_n2]), scor,
                                cntnm=(1., 0., 0., 0., 0., 0., 0.))                 # self continuum only
        got = m["oc"][:, 0, 0]
        idx = ((wn + 20.0) / 10.0).astype(int)                                      # table starts at -20 cm-1, step 10
        cs = s296[idx] * (s260[idx] / s296[idx]) ** ((t - 296.0) / (260.0 - 296.0))
        x_h2o = w_h2o / (w_h2o + w_o2 + w_n2)
        radfn = wn * np.tanh(RADCN2 * wn / (2.0 * t))
        want = w_h2o * cs * x_h2o * (p / 1013.0) * (296.0 / t) * 1.0e-20 * radfn
        assert np.max(np.abs(got / want - 1.0)) < 1e-12, (t, p, got / want)
