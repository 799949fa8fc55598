"""GPU tests of the fused / batched / sharded entry points and of the error behaviour of the C ABI."""
import numpy as np
import pytest

import harness
from monortm_b200 import api, sharding, synth

pytestmark = pytest.mark.gpu


def _fused(case, **kw):
    s = harness.session()
    s.stage_lines(case["ls"])
    return s.profiles(case["wn"], case["dvset"], case["prof"], case["scor"], case["irt"], case["tmpsfc"],
                      case["emiss"], case["reflc"], cntnm=case["cntnm"], ibrd=case["ibrd"], **kw)


def test_fused_profiles_equal_three_operator_calls_and_oracle():
    wn = np.linspace(0.4, 40.0, 300)
    case = harness.make_case(n_filler=384, nlay=18, wn=wn, irt=1, nprof=3, clw=True)
    out = _fused(case, want_o=True, want_otot_by_mol=True, selection=True)
    for ip in range(3):
        ref = harness.run_oracle(case, ip=ip)
        sep = harness.run_gpu(case, ip=ip)
        assert np.array_equal(out["sel_hash"][:, :, ip], ref["sel_hash"])
        assert harness.rel_diff(out["o"][:, :, ip], ref["o"]) < 1e-9
        assert np.array_equal(out["o"][:, :, ip], sep["o"])                  # same kernels, same bits
        for k in ("tb", "tmr"):
            assert np.max(np.abs(out[k][:, ip] - ref[k])) < 1e-5
            assert np.array_equal(out[k][:, ip], sep[k])
        for k in ("rad", "rup", "rdn", "trtot"):
            assert harness.rel_diff(out[k][:, ip], ref[k], floor=1e-300) < 1e-8
        # STOREOUT's per-molecule column optical depths (monortm_sub.F90:649-656)
        col = (ref["o_by_mol"] + ref["oc"]).sum(axis=2).T                     # (39, nwn)
        tot = ref["o"].sum(axis=1)[None, :]
        assert np.max(np.abs(out["otot_by_mol"][:, :, ip] - col) / tot) < 1e-9


def test_device_tips_matches_host_scor():
    # scor=None -> partition sums evaluated on the device from the TIPS tables
    wn = np.linspace(0.5, 30.0, 64)
    case = harness.make_case(n_filler=256, nlay=12, wn=wn, irt=3, nprof=2)
    a = _fused(case, want_o=True)
    c2 = dict(case)
    c2["scor"] = None
    b = _fused(c2, want_o=True)
    assert harness.rel_diff(b["o"], a["o"]) < 1e-13
    assert np.max(np.abs(b["tb"] - a["tb"])) < 1e-9


def test_frequency_shards_reproduce_the_unsharded_run():
    wn, dv = synth.freq_c1_sweep()
    for dvset in (dv, 0.0):
        case = harness.make_case(n_filler=256, nlay=10, wn=wn, dvset=dvset, irt=1)
        full = _fused(case, want_o=True)
        parts = []
        for r in range(3):
            wl, gr = sharding.freq_shard(wn, r, 3)
            c = dict(case)
            c["wn"] = wl
            c["emiss"], c["reflc"] = case["emiss"][gr[2]:gr[2] + len(wl)], case["reflc"][gr[2]:gr[2] + len(wl)]
            parts.append(_fused(c, want_o=True, global_range=gr))
        # the frequency tiling (hence the grouping of the per-line sums) differs between the runs, so
        # agreement is to rounding, not bitwise
        for k in ("rad", "tb", "tmr", "trtot", "o"):
            assert harness.rel_diff(np.concatenate([p[k] for p in parts], axis=0), full[k]) < 1e-10, (k, dvset)


def test_profile_batches_and_memory_budget_split(monkeypatch):
    # many profiles: exercise batching over the plane budget (grid z) and equality with one-by-one runs
    wn = synth.freq_c4_channels(200)
    case = harness.make_case(n_filler=192, nlay=8, wn=wn, irt=3, nprof=9)
    full = _fused(case)
    for ip in (0, 4, 8):
        one = harness.run_gpu(case, ip=ip, by_mol=False)
        assert np.array_equal(full["tb"][:, ip], one["tb"])
    assert np.all(full["tmpsfc"] == 2.75)


def test_properties_at_bench_size_without_the_oracle():
    """Size-independent properties at a size the oracle cannot reach: linearity of optical depth in the
    column amount at fixed broadening state, TB bounded by the profile temperatures, monotone transmittance."""
    wn = 5.5e-5 * np.arange(400000, 400000 + 20000)
    case = harness.make_case(n_filler=1024, nlay=30, wn=wn, irt=3)
    a = _fused(case, want_o=True)
    assert np.all(np.isfinite(a["o"])) and np.all(a["o"] > 0)
    assert np.all(a["trtot"] > 0) and np.all(a["trtot"] <= 1)
    tmax = case["prof"]["tz"].max()
    assert np.all(a["tb"] > 2.7) and np.all(a["tb"] < tmax + 1e-6)
    assert np.all(a["tmr"] > case["prof"]["t"].min() - 1e-6) and np.all(a["tmr"] < tmax + 1e-6)
    # halving the O3 column (trace gas: the broadening state is unchanged to ~1e-6) halves its line OD
    s = harness.session()
    pr = case["prof"]
    m1 = s.modm(wn[:512], 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0],
                case["scor"][:, :, :, 0])
    w2 = pr["wkl"][:, :, 0].copy(order="F")
    w2[2] *= 0.5
    m2 = s.modm(wn[:512], 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, w2, pr["wbrodl"][:, 0],
                case["scor"][:, :, :, 0])
    assert harness.rel_diff(2 * m2["o_by_mol"][:, 2, :], m1["o_by_mol"][:, 2, :]) < 1e-4
    assert np.array_equal(m1["sel_hash"] if "sel_hash" in m1 else 0, m2["sel_hash"] if "sel_hash" in m2 else 0)


def test_error_behaviour_mirrors_reference_stops():
    s = api.Session(0)
    wn = np.linspace(1.0, 2.0, 4)
    case = harness.make_case(n_filler=64, nlay=3, wn=wn, irt=1)
    pr = case["prof"]
    args = (wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0],
            case["scor"][:, :, :, 0])
    with pytest.raises(api.MonortmError) as e:
        s.modm(*args)                                            # no line list staged
    assert e.value.code == 4
    s.stage_lines(case["ls"])
    s.modm(*args)
    r = s.modm(np.array([100.0, 900.0]), *args[1:])              # beyond the microwave: every MT_CKD branch is built (round 1: code 6)
    assert np.all(np.isfinite(r["o"])) and np.all(r["o"] > 0)
    o = np.asfortranarray(np.full((4, 3), 0.1))
    with pytest.raises(api.MonortmError) as e:                   # RTMmono.f90:173 STOP
        s.rtm(1, 1, wn, pr["t"][:, 0], pr["tz"][:, 0], o, 290.0, np.zeros(4), np.ones(4), idu=0)
    assert e.value.code == 8
    bad = pr["t"][:, 0].copy()
    bad[1] = 50.0                                                # TIPS range 70-3000 K -> STOP
    with pytest.raises(api.MonortmError) as e:
        s.modm(wn, 0.0, pr["p"][:, 0], bad, pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0], None)
    assert e.value.code == 11
    # a malformed line store (an XG that GET_LNFL cannot produce: it stores -IFLG for IFLG in 0..100 and the negative
    # flags -1, -3, -5 themselves, lnfl_mod.f90:44-63, 73-77) is refused at staging
    ls = harness.copy_store(case["ls"])
    ls.xg[0, 0] = 2.0
    with pytest.raises(api.MonortmError) as e:
        s.stage_lines(ls)
    assert e.value.code == 5
    s.close()


def test_ragged_and_single_element_inputs():
    # nwn = 1, nlay = 1, and a molecule list shorter than 22 (N2 continuum then uses WBROAD, modm.f90:209)
    case = harness.make_case(n_filler=64, nlay=1, wn=np.array([1.043027]), irt=3, nmol=7)
    ref = harness.run_oracle(case)
    gpu = harness.run_gpu(case)
    assert np.array_equal(gpu["sel_hash"], ref["sel_hash"])
    assert harness.rel_diff(gpu["o"], ref["o"]) < 1e-9 and abs(gpu["tb"][0] - ref["tb"][0]) < 1e-5
    assert ref["oc"][0, 21, 0] > 0
    assert harness.rel_diff(gpu["oc"][:, 21, :], ref["oc"][:, 21, :]) < 1e-9
    # empty line window: frequencies far from every staged line still get continuum + O2
    case = harness.make_case(n_filler=0, nlay=4, wn=np.array([0.05, 0.06]), irt=1, line_kw=dict(with_physical=False))
    ref = harness.run_oracle(case)
    gpu = harness.run_gpu(case)
    assert np.all(ref["sel_count"] == 0) and np.array_equal(gpu["sel_count"], ref["sel_count"])
    assert harness.rel_diff(gpu["o"], ref["o"]) < 1e-9


def test_caller_owned_output_buffers_and_tile_size_by_spectral_width():
    """profiles(out=...) writes the spectra into the caller's arrays (pinned host memory in bench.py); a coarse grid
    (wide tiles would defeat the far field) and a dense grid of the same length give oracle-exact results whichever
    tile size the library picks."""
    for wn in (np.linspace(0.3, 50.0, 2600), 5.5e-5 * np.arange(20000, 22600)):
        case = harness.make_case(n_filler=512, nlay=8, wn=wn, irt=1, nprof=2)
        ref = _fused(case)
        nwn = len(wn)
        out = {k: np.full((nwn, 2), np.nan, order="F") for k in ("rad", "tb", "tmr", "trtot", "rup", "rdn")}
        got = _fused(case, out=out)
        for k in out:
            assert got[k] is out[k]
            assert np.array_equal(out[k], ref[k]), k
        for ip in range(2):
            orc = harness.run_oracle(case, ip=ip)
            assert np.max(np.abs(out["tb"][:, ip] - orc["tb"])) < 1e-5
    with pytest.raises(ValueError):
        _fused(case, out={k: np.zeros((3, 2), order="F") for k in ("rad", "tb", "tmr", "trtot", "rup", "rdn")})


def test_rtm_and_calctmr_on_opaque_and_transparent_columns():
    """RTM / CALCTMR on hand-made optical depths: transparent, ordinary and opaque layers (exp(-tau) underflows, the
    column total exceeds 745) -- rt_kernel forms the path transmittances by recurrence and must re-anchor them where a
    factor underflows (RTMmono.f90:157-221)."""
    rng = np.random.default_rng(7)
    nwn, nlay = 257, 31
    wn = np.sort(rng.uniform(0.05, 55.0, nwn))
    t = np.linspace(288.0, 210.0, nlay)
    tz = np.linspace(290.0, 209.0, nlay + 1)
    o = 10.0 ** rng.uniform(-9, 0.5, (nwn, nlay))
    o[::7, 3] = 900.0                    # one opaque layer low in the column
    o[1::7, nlay - 2] = 2000.0           # one near the top
    o[2::7, :] = 60.0                    # opaque everywhere: total 1860
    o[3::7, :] = 1e-12                   # transparent
    o = np.asfortranarray(o)
    em, rf = np.full(nwn, 0.85), np.full(nwn, 0.15)
    s = harness.session()
    for irt in (1, 3):
        ref = harness.oracle_rtm(1, irt, wn, t, tz, o, 285.0, rf, em)
        got = s.rtm(1, irt, wn, t, tz, o, 285.0, rf, em)
        for k in ("rad", "rup", "rdn", "trtot"):
            assert harness.rel_diff(got[k], ref[k], floor=1e-300) < 1e-10, (k, irt)
        assert np.max(np.abs(got["tb"] - ref["tb"])) < 1e-7
    tmr_ref = harness.oracle_calctmr(wn, t, tz, o)
    tmr = s.calctmr(wn, t, tz, o)
    assert np.max(np.abs(tmr - tmr_ref)) < 1e-7
