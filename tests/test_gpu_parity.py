"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): selected-line set per (frequency, layer) bit-exact; layer optical
depths within 1e-9 relative; brightness temperatures within 1e-5 K.
"""
import numpy as np
import pytest

import harness
from monortm_b200 import synth

pytestmark = pytest.mark.gpu

OD_RTOL = 1e-9      # north_star: layer optical depths within 1e-9 relative
TB_ATOL = 1e-5      # north_star: brightness temperatures within 1e-5 K


def check_case(case, od_rtol=OD_RTOL):
    ref = harness.run_oracle(case)
    # the direct mode (every in-window triple evaluated) must meet the same bars and agree with the default
    # mode (far-field expansion) far inside the tolerance
    direct = harness.run_gpu(case, line_mode=1)
    _check_against(ref, direct, od_rtol)
    gpu = harness.run_gpu(case)
    _check_against(ref, gpu, od_rtol)
    assert np.array_equal(gpu["sel_hash"], direct["sel_hash"])
    assert harness.rel_diff(gpu["o"], direct["o"]) < 1e-10
    assert np.max(np.abs(gpu["tb"] - direct["tb"])) < 1e-7
    assert direct["stats"]["far_expansions"] == 0
    # the production variants: no selection instrumentation, with and without per-molecule outputs
    # (one sum over all molecules, column amounts folded into the line strengths)
    for by_mol in (True, False):
        fast = harness.run_gpu(case, by_mol=by_mol, selection=False)
        assert harness.rel_diff(fast["o"], ref["o"]) < od_rtol
        assert harness.rel_diff(fast["o"], gpu["o"]) < 1e-10
        assert np.max(np.abs(fast["tb"] - ref["tb"])) < TB_ATOL
        if by_mol:
            scale = np.abs(ref["o"])[:, None, :]
            assert np.max(np.abs(fast["o_by_mol"] - ref["o_by_mol"]) / scale) < od_rtol
    return ref, gpu


def _check_against(ref, gpu, od_rtol):
    # selected line set: bit exact (count and order-independent hash of (molecule, record) pairs)
    assert np.array_equal(gpu["sel_count"], ref["sel_count"])
    assert np.array_equal(gpu["sel_hash"], ref["sel_hash"])
    assert harness.rel_diff(gpu["o"], ref["o"]) < od_rtol
    # per-molecule line and continuum optical depths: relative to the layer total where tiny
    scale = np.abs(ref["o"])[:, None, :]
    assert np.max(np.abs(gpu["o_by_mol"] - ref["o_by_mol"]) / scale) < od_rtol
    assert np.max(np.abs(gpu["oc"] - ref["oc"]) / scale) < od_rtol
    assert np.max(np.abs(gpu["o_clw"] - ref["o_clw"]) / np.abs(ref["o"])) < od_rtol
    assert np.max(np.abs(gpu["tb"] - ref["tb"])) < TB_ATOL
    assert np.max(np.abs(gpu["tmr"] - ref["tmr"])) < TB_ATOL
    for k in ("rad", "rup", "rdn", "trtot"):
        assert harness.rel_diff(gpu[k], ref[k], floor=1e-300) < 1e-8, k
    assert gpu["tmpsfc"] == ref["tmpsfc"]
    assert gpu["stats"]["kernel_launches"] >= 4


def test_c1_channels_downwelling():
    case = harness.make_case(n_filler=512, nlay=19, wn=synth.freq_c1_channels(), irt=3)
    ref, gpu = check_case(case)
    assert gpu["tmpsfc"] == 2.75      # RTMmono.f90:122


def test_c1_sweep_gridded_upwelling():
    wn, dv = synth.freq_c1_sweep()
    case = harness.make_case(n_filler=512, nlay=25, wn=wn, dvset=dv, irt=1, tmpsfc=290.0, emis=0.6)
    check_case(case)


def test_c2_sounder_channels_up():
    case = harness.make_case(n_filler=1024, nlay=40, wn=synth.freq_c2_sounder(), irt=1, tmpsfc=285.0, emis=0.9)
    check_case(case)


def test_cloudy_up_and_down():
    wn = np.linspace(0.0055, 55.0, 160)
    for irt in (1, 3):
        case = harness.make_case(n_filler=768, nlay=30, wn=wn, irt=irt, clw=True)
        ref, gpu = check_case(case)
        assert np.any(ref["o_clw"] > 0)


def test_voigt_and_dense_grid():
    # a dense grid around the 22 GHz H2O line and the O2 band, upper layers take the Voigt branch
    wn = np.concatenate([0.741691 + np.linspace(-2e-4, 2e-4, 97), 2.0 + np.linspace(-0.05, 0.05, 160)])
    wn.sort()
    case = harness.make_case(n_filler=256, nlay=60, wn=wn, irt=3)
    ref, gpu = check_case(case)
    assert ref["n_voigt"] > 0


def test_coverage_lines_co2_lc_sdep_ibrd():
    wn = np.linspace(0.5, 50.0, 200)
    kw = dict(n_co2=12, n_sdep=8, n_generic_lc=8, brd_fraction=0.2)
    for ibrd in (0, 1):
        case = harness.make_case(n_filler=384, nlay=24, wn=wn, irt=1, ibrd=ibrd, line_kw=kw)
        check_case(case)


def test_continuum_factors_and_scalings():
    wn = np.linspace(0.2, 30.0, 64)
    case = harness.make_case(n_filler=256, nlay=16, wn=wn, irt=3, cntnm=(0.7, 1.2, 1.0, 1.0, 1.0, 0.0, 1.0),
                             sclcpl=0.87, sclhw=1.1, y0res=0.01)
    check_case(case)


def test_large_frequency_tiles():
    # exercises the F=4 kernel variant and tile edges (nwn not a multiple of the tile)
    wn = 5.5e-5 * np.arange(1, 2600 + 1) * 300.0
    case = harness.make_case(n_filler=192, nlay=6, wn=wn, irt=1)
    check_case(case)


def test_dense_grid_far_field_expansion():
    """C3-like dense grid (5.5e-5 cm-1 spacing): most in-window lines take the far-field Taylor path; the
    result must still match the oracle to 1e-9 and the direct mode to 1e-10 (the 14-term series at ratio 6 truncates at 1.3e-11), selection bit-exact."""
    for i0 in (1, 13400, 400000):          # next to zero frequency, across the 22 GHz line, mid-band
        wn = 5.5e-5 * np.arange(i0, i0 + 1536)
        case = harness.make_case(n_filler=1536, nlay=6, wn=wn, irt=1, line_kw=dict(n_co2=4, n_generic_lc=4))
        ref, gpu = check_case(case)
        assert gpu["stats"]["far_expansions"] > 0
        assert gpu["stats"]["direct_evals"] < 0.2 * ref["sel_count"].sum()


@pytest.mark.parametrize("env", [
    {"MRTM_FARW_MIN": "1", "MRTM_FF_S": "2", "MRTM_FF_LEVELS": "4"},      # far_warp_kernel on every level, 4 levels
    {"MRTM_FARW_MIN": "1000000000"},                                      # far_kernel (one CTA per tile and layer) only
    {"MRTM_SIDE_STREAM": "0"},                                            # everything on one stream
    {"MRTM_VOIGT_SIDE": "1"},                                             # Voigt branch on the side stream (scratch plane)
    {"MRTM_NEAR2": "0"},                                                  # near_kernel for every tile
])
def test_kernel_variants_on_the_dense_grid(env, monkeypatch):
    """The library reads its tuning switches when a context is created: run the dense-grid case (13 layers, not a
    multiple of far_warp_kernel's four layers per CTA) under each variant against the oracle and the direct mode."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setattr(harness, "_session", None)
    wn = 5.5e-5 * np.arange(13400, 13400 + 2304)
    case = harness.make_case(n_filler=1536, nlay=13, wn=wn, irt=1, line_kw=dict(n_co2=4, n_generic_lc=4))
    ref = harness.run_oracle(case)
    direct = harness.run_gpu(case, line_mode=1)
    for by_mol in (True, False):
        gpu = harness.run_gpu(case, by_mol=by_mol, selection=by_mol)
        assert harness.rel_diff(gpu["o"], ref["o"]) < OD_RTOL
        assert harness.rel_diff(gpu["o"], direct["o"]) < 1e-10
        assert np.max(np.abs(gpu["tb"] - ref["tb"])) < TB_ATOL
        assert gpu["stats"]["far_expansions"] > 0
        if by_mol:
            assert np.array_equal(gpu["sel_count"], ref["sel_count"])
            assert np.array_equal(gpu["sel_hash"], ref["sel_hash"])
    monkeypatch.setattr(harness, "_session", None)     # the next test creates its context with the default switches


def test_unsorted_frequency_list():
    """List mode (MONORTM.IN record 1.3.1, monortm_sub.F90:264-276) takes the frequencies in any order: a shuffled
    dense list (far field, Voigt zones and window edges inside every tile) must give the oracle's numbers."""
    rng = np.random.default_rng(11)
    wn = np.concatenate([5.5e-5 * np.arange(13400, 13400 + 700), np.linspace(0.3, 54.0, 300), 2.0 + np.linspace(-0.02, 0.02, 100)])
    rng.shuffle(wn)
    case = harness.make_case(n_filler=1024, nlay=14, wn=wn, irt=1)
    ref = harness.run_oracle(case)
    gpu = harness.run_gpu(case)
    _check_against(ref, gpu, OD_RTOL)


def test_gpu_isolated_line_against_the_textbook_expression():
    """The CUDA path itself (not only the oracle) against the from-scratch numpy expression for one isolated N2O-like
    line: Lorentz regime at three (T, p) points to 2e-9 (HITRAN intensity scaling, radiation-field term, 25 cm-1 cutoff
    with pedestal, negative-frequency partner, width mixing, density-scaled shift) -- see
    tests/test_oracle_units.py::test_isolated_line_matches_textbook_formula for the derivation."""
    import os
    import tempfile
    from monortm_b200 import api, linefile

    c2 = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16
    t0, p0 = 296.0, 1013.25
    v0 = 12.3456
    recs = np.zeros(1, synth.REC_DTYPE)
    recs[0] = synth._line(v0, 3.0e-23, 0.08, 0.11, 350.0, 0.7, -2.0e-3, 4, 1)
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, 0.0, 55.0)
    finally:
        os.unlink(path)
    g_air, g_self = float(ls.alpf[3, 0]), float(ls.alps[3, 0])
    epp, xexp, delta = float(ls.e[3, 0]), float(ls.x[3, 0]), float(ls.deltnu[3, 0])
    s0 = float(ls.s0[3, 0]) * (v0 * (1.0 - np.exp(-c2 * v0 / t0)))
    wn = np.array([0.5, 5.0, 12.0, 12.3, 12.3456, 12.4, 20.0, 30.0, 37.3, 37.4, 40.0])
    s = harness.session()
    s.stage_lines(ls)
    for t, p in ((296.0, 500.0), (250.0, 800.0), (296.0, 1013.25)):
        w_n2o, w_n2 = 3.0e17, 1.0e23
        wkl = np.zeros((39, 1), order="F")
        wkl[3, 0] = w_n2o
        scor = api.scor_for_layers(7, np.array([[t]]))[:, :, :, 0]
        for mode in (0, 1):
            m = s.modm(wn, 0.0, np.array([p]), np.array([t]), np.array([0.0]), 7, wkl, np.array([w_n2]), scor,
                       want_by_mol=True, selection=False, line_mode=mode)
            got = m["o_by_mol"][:, 3, 0]
            rho = (p / t) / (p0 / t0)
            x_self = w_n2o / (w_n2o + w_n2)
            vc = v0 + delta * rho
            gam = (g_air * (1.0 - x_self) + g_self * x_self) * rho * (t / t0) ** xexp
            s_t = s0 * scor[3, 0, 0] * np.exp(-c2 * epp * (1.0 / t - 1.0 / t0)) * (1 - np.exp(-c2 * vc / t)) / (1 - np.exp(-c2 * vc / t0))

            def lor(d):
                return gam / (np.pi * (d * d + gam * gam))
            shape = np.where(np.abs(wn - vc) <= 25.0,
                             lor(wn - vc) - lor(25.0) + np.where(wn + vc <= 25.0, lor(wn + vc) - lor(25.0), 0.0), 0.0)
            want = w_n2o * s_t * (wn * np.tanh(c2 * wn / (2 * t))) / (vc * np.tanh(c2 * vc / (2 * t))) * shape
            assert np.all((want == 0) == (got == 0))
            nz = want != 0
            assert np.max(np.abs(got[nz] / want[nz] - 1.0)) < 2e-9, (t, p, mode)
