"""Several GPUs behind one context (mrtm_init_multi): the in-library split by profile and by frequency against
the single-GPU result.  On a one-GPU box the partitioning logic still runs: MRTM_MULTI_CONTEXTS_PER_DEVICE puts several
contexts on the same device; with `gpurun --gpus N` the contexts sit on different GPUs."""
import os

import numpy as np
import pytest

import harness
from monortm_b200 import api, synth

pytestmark = pytest.mark.gpu


def _multi_session():
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        return api.Session(device_mask=(1 << min(n, 4)) - 1), min(n, 4)
    os.environ["MRTM_MULTI_CONTEXTS_PER_DEVICE"] = "3"
    try:
        return api.Session(device_mask=1), 3
    finally:
        del os.environ["MRTM_MULTI_CONTEXTS_PER_DEVICE"]


def test_in_library_multi_gpu_matches_single_gpu():
    multi, n = _multi_session()
    assert multi.num_devices() == n
    single = harness.session()
    ls = harness.synthetic_store(1024, v1=0.0, v2=20.0)
    assert multi.stage_lines(ls) == single.stage_lines(ls)
    # ---- by profile: 7 profiles over n devices (uneven blocks), layer optical depths and per-molecule totals included
    wn = np.linspace(0.5, 18.0, 300)
    prof = synth.synthetic_profiles(7, 11, seed0=4000, clw_layers=True, nmol=22)
    scor = api.scor_for_layers(22, prof["t"])
    em, rf = np.linspace(0.5, 0.9, 300), np.linspace(0.5, 0.1, 300)
    kw = dict(want_o=True, want_otot_by_mol=True, selection=True)
    a = single.profiles(wn, 0.0, prof, scor, 1, 285.0, em, rf, **kw)
    b = multi.profiles(wn, 0.0, prof, scor, 1, 285.0, em, rf, **kw)
    for k in ("rad", "tb", "tmr", "trtot", "rup", "rdn", "o", "otot_by_mol", "sel_count", "sel_hash", "tmpsfc"):
        assert np.array_equal(a[k], b[k]), k
    # ---- by frequency: fewer profiles than devices; a dense sweep so that the far-field tiles differ between the two runs
    wn = 2.0 + 5.5e-4 * np.arange(3 * 2048 + 700)
    one = synth.synthetic_profiles(1, 9, seed0=4100, clw_layers=False, nmol=22)
    scor = api.scor_for_layers(22, one["t"])
    em, rf = np.full(len(wn), 0.8), np.full(len(wn), 0.2)
    for irt in (1, 3):
        a = single.profiles(wn, 0.0, one, scor, irt, 285.0, em, rf, **kw)
        for rep in range(3):                         # the second and third call run on the rebalanced partition
            b = multi.profiles(wn, 0.0, one, scor, irt, 285.0, em, rf, **kw)
            assert np.array_equal(a["sel_count"], b["sel_count"]) and np.array_equal(a["sel_hash"], b["sel_hash"])
            assert harness.rel_diff(b["o"], a["o"]) < 1e-10 and harness.rel_diff(b["otot_by_mol"], a["otot_by_mol"], floor=1e-300) < 1e-10
            assert np.max(np.abs(a["tb"] - b["tb"])) < 1e-8 and np.max(np.abs(a["tmr"] - b["tmr"])) < 1e-8
            for k in ("rad", "trtot", "rup", "rdn"):
                assert harness.rel_diff(b[k], a[k], floor=1e-300) < 1e-10, k
            assert b["tmpsfc"][0] == a["tmpsfc"][0]
    # gridded mode (DVSET != 0): the interpolation origin stays that of the whole list (modm.f90:218-219)
    wn = 1.0 + 0.002 * np.arange(2 * 2048 + 100)
    em, rf = np.full(len(wn), 0.8), np.full(len(wn), 0.2)
    a = single.profiles(wn, 0.002, one, scor, 1, 285.0, em, rf, want_o=True)
    b = multi.profiles(wn, 0.002, one, scor, 1, 285.0, em, rf, want_o=True)
    assert harness.rel_diff(b["o"], a["o"]) < 1e-10
    st = multi.stats()
    assert st["kernel_launches"] > 0 and st["lines_staged"] == single.stats()["lines_staged"]
    multi.close()


def test_device_resident_call_is_asynchronous_and_reports_deferred_errors():
    import torch
    s = harness.session()
    ls = harness.synthetic_store(256, v1=0.0, v2=20.0)
    s.stage_lines(ls)
    nwn, nlay = 4096, 8
    wn = 1.0 + 1e-3 * np.arange(nwn)
    pr = synth.synthetic_profiles(1, nlay, seed0=4200, clw_layers=False, nmol=22)
    dev = torch.device("cuda", 0)

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).to(dev)
    d = {k: dv(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    d["wn"], d["emiss"], d["reflc"] = dv(wn), dv(np.full(nwn, 0.9)), dv(np.full(nwn, 0.1))
    d["tmpsfc"] = torch.tensor([288.0], dtype=torch.float64, device=dev)
    outs = torch.zeros(6, nwn, dtype=torch.float64, device=dev)
    ptrs = {k: v.data_ptr() for k, v in d.items()}
    for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
        ptrs[k] = outs[i].data_ptr()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        s.profiles_dev(1, nwn, nlay, 22, 0.0, ptrs, float(wn[0]), float(wn[-1]), 0, 1, stream=st.cuda_stream)   # warm up (allocations)
        s.sync()
        s.profiles_dev(1, nwn, nlay, 22, 0.0, ptrs, float(wn[0]), float(wn[-1]), 0, 1, stream=st.cuda_stream)
        s.sync()
    ref = s.profiles(wn, 0.0, pr, None, 1, 288.0, np.full(nwn, 0.9), np.full(nwn, 0.1))
    assert np.max(np.abs(outs[1].cpu().numpy() - ref["tb"][:, 0])) < 1e-9
    assert s.stats()["last_lines_kernel_ms"] > 0
    # a device-detected error (layer temperature outside the TIPS range, tips_2003.f90:271) surfaces at mrtm_sync
    bad = pr["t"].copy()
    bad[2, 0] = 20.0
    d["t"].copy_(dv(bad))
    with torch.cuda.stream(st):
        s.profiles_dev(1, nwn, nlay, 22, 0.0, ptrs, float(wn[0]), float(wn[-1]), 0, 1, stream=st.cuda_stream)
        with pytest.raises(api.MonortmError) as e:
            s.sync()
    assert e.value.code == 11
    s.sync()                                                        # reported once
