"""GPU parity on the branches that ordinary grids never reach ("What's weak" 1):
speed-dependent Voigt (modm.f90:1022-1066 -> SD_Humlicek), CO2 (uncoupled and XF=-1) and coupled lines of
other molecules inside the Voigt zone (:588-615, :659), line flags outside {0,1,3,5}, IRT=2 (RTMmono.f90:142-143),
a far-field hierarchy with at least four tiles on every level, and the radiance recurrence at the 600-layer scale
with opaque, transparent and negative optical depths.

Bars (BASELINE.json north_star): selected-line set bit-exact, layer optical depths 1e-9 relative, TB 1e-5 K.
"""
import numpy as np
import pytest

import harness
from test_gpu_parity import OD_RTOL, TB_ATOL, _check_against, check_case

pytestmark = pytest.mark.gpu

LC = (-1.0, -3.0, -5.0)


def logical_lines(ls):
    """(molecule index 0-based, record index, XG) of every logical line: the record walk of LINES (modm.f90:324-434)."""
    out = []
    for m in range(39):
        n, j = int(ls.nblm[m]), 0
        while j < n:
            xg = float(ls.xg[m, j])
            out.append((m, j, xg))
            j += 2 if xg in LC else 1
    return out


def zone_frequencies(centres, mult=(0.0, 0.4, 4.0, 40.0, 130.0)):
    """Frequencies at the centres and at +-mult*1e-6*centre: the Doppler HWHM is (0.8..1.3)e-6*centre for the molecules
    used here at 220-290 K, so the offsets sit at ~0.3, 3, 30 HWHM_D (inside the 100*HWHM_D zone of modm.f90:427) and
    just outside it."""
    offs = np.array(sorted({s * m for m in mult for s in (1.0, -1.0)})) * 1e-6
    c = np.asarray(centres, dtype=np.float64)
    return np.unique((c[:, None] * (1.0 + offs[None, :])).ravel())


KW = dict(n_co2=12, n_sdep=8, n_generic_lc=8, brd_fraction=0.2)


def _special_centres(n_filler, kw):
    ls = harness.synthetic_store(n_filler, v1=0.0, v2=55.0, **kw)
    L = logical_lines(ls)
    sd = [float(ls.xnu0[m, j]) for m, j, xg in L if abs(ls.sdep[m, j]) > 1e-4]
    co2 = [float(ls.xnu0[m, j]) for m, j, xg in L if m == 1]
    glc = [float(ls.xnu0[m, j]) for m, j, xg in L if m not in (1, 6) and xg in LC]
    o2lc = [float(ls.xnu0[m, j]) for m, j, xg in L if m == 6 and xg == -1.0][:6]
    return sd, co2, glc, o2lc


@pytest.mark.parametrize("ibrd", [0, 1])
def test_voigt_zone_speed_dependence_co2_and_coupled_lines(ibrd):
    sd, co2, glc, o2lc = _special_centres(384, KW)
    assert len(sd) == 8 and len(co2) == 12 and len(glc) == 8 and len(o2lc) > 0
    wn = zone_frequencies(sd + co2 + glc + o2lc)
    wn = wn[(wn > 0.4) & (wn < 55.0)]
    case = harness.make_case(n_filler=384, nlay=24, wn=wn, irt=1, ibrd=ibrd, line_kw=KW)
    ref, gpu = check_case(case)
    br = ref["branches"]
    # the oracle's own count of what the run reached: every one of these is a branch the GPU result was compared on
    assert br["sdep"] > 500, br            # SDVOIGT speed-dependent branch -> SD_Humlicek (modm.f90:1022-1066)
    assert br["co2"] > 300 and br["co2_lc1"] > 100, br     # CO2 on the Voigt branch: uncoupled (:640-648) and XF=-1 (:659-686)
    assert br["generic_lc"] > 200, br      # coupled lines of other molecules (:588-604)
    assert br["o2_lc"] > 50, br            # coupled O2 lines (:652-667)
    assert br["neg_res"] > 100, br         # negative-frequency resonance on the Voigt branch (:594, 609, 632)


def test_line_flags_outside_1_3_5_are_uncoupled_lines():
    """GET_LNFL stores XG = -IFLG for any IFLG in 0..100 (lnfl_mod.f90:44-45, 73-77); LINES and the LSF routines only know
    -1, -3, -5, so such a line runs as an uncoupled line -- including on the Voigt branch and for CO2 and O2."""
    kw = dict(n_co2=16)
    base = harness.synthetic_store(256, v1=0.0, v2=55.0, **kw)
    L = logical_lines(base)
    pick = {}
    for m, j, xg in L:                       # four uncoupled lines each of H2O, CO2, O3, O2
        if xg == 0.0 and m in (0, 1, 2, 6) and 0.6 < base.xnu0[m, j] < 50.0 and len(pick.setdefault(m, [])) < 4:
            pick[m].append(j)
    assert all(len(v) >= 2 for v in pick.values()) and len(pick) == 4
    cen = [float(base.xnu0[m, j]) for m, js in pick.items() for j in js]
    wn = np.unique(np.concatenate([zone_frequencies(cen), np.linspace(0.5, 50.0, 40)]))
    case = harness.make_case(n_filler=256, nlay=18, wn=wn, irt=1, line_kw=kw)
    ls = harness.copy_store(case["ls"])
    flags = (-2.0, -4.0, -7.0, -100.0)
    nset = 0
    for m, js in pick.items():
        for j0, fl in zip(js, flags):
            # the same physical lines in the store of this case (loaded for wn[0]-25 .. wn[-1]+25)
            hit = np.nonzero(ls.xnu0[m, :int(ls.nblm[m])] == base.xnu0[m, j0])[0]
            assert len(hit) == 1 and ls.xg[m, hit[0]] == 0.0
            ls.xg[m, hit[0]] = fl
            nset += 1
    assert nset >= 8
    flagged = dict(case, ls=ls)
    ref = harness.run_oracle(flagged)
    assert ref["branches"]["other_flag"] > 0
    for mode in (1, 0):
        gpu = harness.run_gpu(flagged, line_mode=mode)
        _check_against(ref, gpu, OD_RTOL)
    # and the flag really is inert: same numbers as with IFLG=0
    plain = harness.run_oracle(case)
    assert np.array_equal(plain["o"], ref["o"])


def test_irt2_limb():
    """IRT=2 (RTMmono.f90:113-123, 142-143): TMPSFC := 2.75, RAD = RUP + TRTOT*(RDN + TRTOT*B_cosmic)."""
    wn = np.linspace(0.3, 30.0, 96)
    case = harness.make_case(n_filler=384, nlay=22, wn=wn, irt=2, clw=True)
    ref, gpu = check_case(case)
    assert gpu["tmpsfc"] == 2.75 and ref["tmpsfc"] == 2.75
    up = harness.run_oracle(dict(case, irt=1))
    assert np.max(np.abs(up["rad"] / ref["rad"] - 1.0)) > 0.1      # the limb formula is not the upwelling one


def test_far_field_hierarchy_four_tiles_on_every_level():
    """131072 + 700 dense frequencies (5.5e-5 cm-1): 257 / 65 / 17 / 5 tiles on the four hierarchy levels (512 * 4^l
    frequencies per tile), a ragged last tile on each.  The oracle runs on a 1/64 subsample (plus both ends); optical
    depths, TB, TMR and the selected line set are compared there, default mode and direct mode."""
    n = 131072 + 700
    i0 = 13000                                  # the block crosses the 22 GHz H2O line (0.7417 cm-1 = index 13485)
    wn = 5.5e-5 * np.arange(i0, i0 + n)
    case = harness.make_case(n_filler=4096, nlay=4, wn=wn, irt=1, line_kw=dict(n_co2=4, n_generic_lc=4))
    idx = np.unique(np.concatenate([np.arange(0, n, 64), [n - 1]]))
    sub = dict(case, wn=wn[idx], emiss=case["emiss"][idx], reflc=case["reflc"][idx])
    ref = harness.run_oracle(sub)
    for mode in (0, 1):
        gpu = harness.run_gpu(case, by_mol=False, line_mode=mode)
        assert np.array_equal(gpu["sel_count"][idx], ref["sel_count"]), mode
        assert np.array_equal(gpu["sel_hash"][idx], ref["sel_hash"]), mode
        assert harness.rel_diff(gpu["o"][idx], ref["o"]) < OD_RTOL, mode
        assert np.max(np.abs(gpu["tb"][idx] - ref["tb"])) < TB_ATOL, mode
        assert np.max(np.abs(gpu["tmr"][idx] - ref["tmr"])) < TB_ATOL, mode
        if mode == 0:
            assert gpu["stats"]["far_expansions"] > 0
            assert gpu["stats"]["direct_evals"] < 0.05 * float(ref["sel_count"].sum()) * 64
        else:
            assert gpu["stats"]["far_expansions"] == 0


def test_radiance_recurrence_at_600_layers_with_negative_optical_depths():
    """rt_kernel carries the path transmittances by recurrence (one exponential per layer instead of the reference's
    EXP(-ODT) per layer, RTMmono.f90:196-216).  603 layers (MXLAY) with opaque, transparent and slightly negative layer
    optical depths (the pedestal subtraction can produce them) against the oracle: rounding must not accumulate."""
    rng = np.random.default_rng(5)
    nlay, nwn = 603, 64
    wn = np.linspace(0.5, 50.0, nwn)
    tz = 288.0 - 70.0 * np.linspace(0.0, 1.0, nlay + 1) + rng.normal(0.0, 0.5, nlay + 1)
    t = 0.5 * (tz[:-1] + tz[1:])
    o = np.zeros((nwn, nlay), order="F")
    for i in range(nwn):
        kind = i % 4
        if kind == 0:      # transparent column with negative layers
            o[i] = rng.uniform(-2e-6, 5e-6, nlay)
        elif kind == 1:    # moderate, total ~ 3
            o[i] = rng.uniform(0.0, 1e-2, nlay)
        elif kind == 2:    # opaque: total ~ 900, a few very thick layers
            o[i] = rng.uniform(0.0, 3.0, nlay)
            o[i, rng.integers(0, nlay, 5)] = 300.0
        else:              # alternating sign, thin
            o[i] = 1e-4 * np.where(np.arange(nlay) % 2 == 0, 1.0, -0.7)
    s = harness.session()
    for irt in (1, 3, 2):
        ref = harness.oracle_rtm(1, irt, wn, t, tz, o, 290.0, np.full(nwn, 0.3), np.full(nwn, 0.7))
        gpu = s.rtm(1, irt, wn, t, tz, o, 290.0, np.full(nwn, 0.3), np.full(nwn, 0.7))
        for k in ("rad", "rup", "rdn", "trtot"):
            assert harness.rel_diff(gpu[k], ref[k], floor=1e-300) < 1e-10, (irt, k)
        assert np.max(np.abs(gpu["tb"] - ref["tb"])) < 1e-7, irt
    tmr_ref = harness.oracle_calctmr(wn, t, tz, o)
    tmr_gpu = s.calctmr(wn, t, tz, o)
    ok = np.isfinite(tmr_ref)
    assert np.array_equal(ok, np.isfinite(tmr_gpu))
    assert np.max(np.abs(tmr_gpu[ok] - tmr_ref[ok])) < TB_ATOL


@pytest.mark.parametrize("i0,nlay,lb", [(1, 7, 0), (13100, 9, 4), (400000, 5, 2), (453000, 6, 1), (909000, 3, 16)])
def test_production_near_field_kernel_on_dense_grids(i0, nlay, lb, monkeypatch):
    """near3_kernel (plan-driven lists, a group of layers per CTA) is what runs when neither per-molecule outputs nor the
    selection instrumentation are requested: dense 5.5e-5 cm-1 grids next to zero frequency (both resonances everywhere),
    across the 22 GHz line and the window edges of the strong lines, with layer groups that do not divide the layer
    count, a ragged last tile, CO2 / coupled lines in the list.  Against the oracle and the direct mode."""
    if lb:
        monkeypatch.setenv("MRTM_NEAR3_LB", str(lb))
    monkeypatch.setattr(harness, "_session", None)
    n = 4096 + 300
    wn = 5.5e-5 * np.arange(i0, i0 + n)
    case = harness.make_case(n_filler=4096, nlay=nlay, wn=wn, irt=1, line_kw=dict(n_co2=4, n_generic_lc=4, n_sdep=4))
    idx = np.unique(np.concatenate([np.arange(0, n, 16), [n - 1]]))
    sub = dict(case, wn=wn[idx], emiss=case["emiss"][idx], reflc=case["reflc"][idx])
    ref = harness.run_oracle(sub)
    direct = harness.run_gpu(case, by_mol=False, selection=False, line_mode=1)
    fast = harness.run_gpu(case, by_mol=False, selection=False)
    assert fast["stats"]["far_expansions"] > 0
    assert harness.rel_diff(fast["o"][idx], ref["o"]) < OD_RTOL
    assert harness.rel_diff(fast["o"], direct["o"]) < 1e-10
    assert np.max(np.abs(fast["tb"][idx] - ref["tb"])) < TB_ATOL
    assert np.max(np.abs(fast["tb"] - direct["tb"])) < 1e-7
    # the instrumented path (near2_kernel) gives the same numbers and the oracle's selection
    inst = harness.run_gpu(case, by_mol=False, selection=True)
    assert harness.rel_diff(inst["o"], fast["o"]) < 1e-10
    assert np.array_equal(inst["sel_hash"][idx], ref["sel_hash"])
    monkeypatch.setattr(harness, "_session", None)


def test_production_near_field_kernel_unsorted_and_voigt_zone():
    rng = np.random.default_rng(3)
    sd, co2, glc, o2lc = _special_centres(384, KW)
    wn = np.concatenate([zone_frequencies(sd + co2 + glc + o2lc), 5.5e-5 * np.arange(13400, 13400 + 2200), np.linspace(0.3, 54.0, 400)])
    wn = wn[(wn > 0.05) & (wn < 55.0)]
    rng.shuffle(wn)
    case = harness.make_case(n_filler=384, nlay=11, wn=wn, irt=1, line_kw=KW)
    ref = harness.run_oracle(case)
    import os
    os.environ["MRTM_LINES_F"] = "4"
    try:
        harness._session = None
        fast = harness.run_gpu(case, by_mol=False, selection=False)
    finally:
        del os.environ["MRTM_LINES_F"]
        harness._session = None
    assert harness.rel_diff(fast["o"], ref["o"]) < OD_RTOL
    assert np.max(np.abs(fast["tb"] - ref["tb"])) < TB_ATOL


def _run_with_env(case, env, **kw):
    """One run of `case` on a fresh library context created under `env` (the library reads its MRTM_* switches in mrtm_init)."""
    import os
    from monortm_b200 import api
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    saved = harness._session
    try:
        harness._session = api.Session(0)
        return harness.run_gpu(case, **kw)
    finally:
        harness._session = saved
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("F", ["4", "2"])
def test_transposed_voigt_kernel_on_every_voigt_branch(F, variant="1"):
    """voigtT_kernel (dense tiles: lines over warps, lanes over each line's run of frequencies) on the Voigt-zone branch
    case -- speed dependence, CO2, coupled lines of other molecules, coupled O2, the negative-frequency resonance -- plus a
    dense stretch across the 22 GHz line, with the dense-tile kernels forced (MRTM_LINES_F): against the oracle (1e-9, both
    with per-molecule outputs and with one sum over all molecules) and against voigt_kernel on the same tiles (MRTM_VOIGT_T=0)."""
    sd, co2, glc, o2lc = _special_centres(384, KW)
    wn = zone_frequencies(sd + co2 + glc + o2lc)
    wn = np.unique(np.concatenate([wn[(wn > 0.4) & (wn < 55.0)], 0.741691 + np.linspace(-3e-4, 3e-4, 700)]))
    case = harness.make_case(n_filler=384, nlay=24, wn=wn, irt=1, line_kw=KW)
    ref = harness.run_oracle(case)
    assert ref["branches"]["sdep"] > 500 and ref["branches"]["co2_lc1"] > 100 and ref["branches"]["generic_lc"] > 200
    scale = np.abs(ref["o"])[:, None, :]
    new = _run_with_env(case, {"MRTM_LINES_F": F, "MRTM_VOIGT_T": variant})
    _check_against(ref, new, OD_RTOL)
    old = _run_with_env(case, {"MRTM_LINES_F": F, "MRTM_VOIGT_T": "0"})
    _check_against(ref, old, OD_RTOL)
    assert harness.rel_diff(new["o"], old["o"]) < 1e-12
    assert np.max(np.abs(new["o_by_mol"] - old["o_by_mol"]) / scale) < 1e-12
    fast = _run_with_env(case, {"MRTM_LINES_F": F, "MRTM_VOIGT_T": variant}, by_mol=False, selection=False)
    assert harness.rel_diff(fast["o"], ref["o"]) < OD_RTOL
    assert harness.rel_diff(fast["o"], new["o"]) < 1e-10
    assert np.max(np.abs(fast["tb"] - ref["tb"])) < TB_ATOL


@pytest.mark.parametrize("tile", ["128", "64", "32"])
def test_channel_tile_sizes_agree(tile):
    """Channel lists run with one frequency per thread and tiles of MRTM_COARSE_TILE channels (32 by default: one warp per
    (tile, layer), the level-0 far field takes most of the window).  A 200-channel log-spaced list (ragged last tile) with
    cloud, against the oracle under each tile size."""
    wn = np.exp(np.linspace(np.log(0.1), np.log(30.0), 200))
    case = harness.make_case(n_filler=1024, nlay=20, wn=wn, irt=1, clw=True, tmpsfc=285.0, emis=0.85)
    ref = harness.run_oracle(case)
    gpu = _run_with_env(case, {"MRTM_COARSE_TILE": tile})
    _check_against(ref, gpu, OD_RTOL)
    fast = _run_with_env(case, {"MRTM_COARSE_TILE": tile}, by_mol=False, selection=False)
    assert harness.rel_diff(fast["o"], ref["o"]) < OD_RTOL
    assert np.max(np.abs(fast["tb"] - ref["tb"])) < TB_ATOL
    if tile != "128":
        assert fast["stats"]["far_expansions"] > 0


def test_transposed_voigt_kernel_on_an_unsorted_dense_list():
    """voigtT_kernel brackets each line's run of frequencies by binary search on an ascending tile; a tile that is not
    ascending keeps the whole tile as the run and relies on the per-pair test of the reference."""
    rng = np.random.default_rng(5)
    wn = np.concatenate([0.741691 + np.linspace(-3e-4, 3e-4, 1500), 2.0 + np.linspace(-0.02, 0.02, 548)])
    rng.shuffle(wn)
    case = harness.make_case(n_filler=256, nlay=40, wn=wn, irt=3)
    ref = harness.run_oracle(case)
    assert ref["n_voigt"] > 0
    gpu = _run_with_env(case, {"MRTM_LINES_F": "4", "MRTM_VOIGT_T": "1"})
    _check_against(ref, gpu, OD_RTOL)
