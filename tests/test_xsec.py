"""Cross sections: the C++ XSREAD / table reader against its Python mirror on the seeded synthetic set, the
real FSCDXS when the reference tree is present, and -- on the GPU -- mrtm_xsec and MODM with IXSECT=1 against the oracle
and the vectors produced by executing the reference text (tests/golden/ref_xsec_synth.npz)."""
import json
import os

import numpy as np
import pytest

import harness
from monortm_b200 import api, xsfile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x["ixmol"], x["v1fx"], x["v2fx"], len(x["files"])) == (y["ixmol"], y["v1fx"], y["v2fx"], len(y["files"]))
        assert abs(x["xdoplr"] - y["xdoplr"]) <= 2e-16 * abs(y["xdoplr"])
        for f, g in zip(x["files"], y["files"]):
            assert (f["v1x"], f["v2x"], f["npts"], f["t"]) == (g["v1x"], g["v2x"], g["npts"], g["t"])
            assert abs(f["pres"] - g["pres"]) <= 2e-16 * abs(g["pres"])
            assert np.array_equal(f["data"], g["data"])


def test_cpp_xsread_matches_the_python_mirror_on_the_synthetic_set(tmp_path):
    xsfile.synthetic_set(str(tmp_path))
    for names, rng in ((["HNO3", "F11"], (1.5, 50.0)), (["CFC11"], (19.0, 305.0)), (["HNO3"], (13.0, 39.0))):
        _same(api.host_xsread(str(tmp_path), names, *rng), xsfile.read_regions(str(tmp_path), names, *rng))
    assert api.host_xsread(str(tmp_path), ["HNO3"], 13.0, 39.0) == []          # no region overlaps: XSREAD selects nothing
    with pytest.raises(api.MonortmError) as e:                                 # XSREAD :1332-1333 STOP
        api.host_xsread(str(tmp_path), ["XYZ"], 1.0, 50.0)
    assert "IS NOT ONE OF THE CROSS SECTION MOLECULES" in str(e.value)
    with pytest.raises(api.MonortmError) as e:                                 # :1398-1404 STOP (on the alias table, absent from FSCDXS)
        api.host_xsread(str(tmp_path), ["CCL4"], 1.0, 50.0)
    assert "IS NOT FOUND ON FILE FSCDXS" in str(e.value)


@pytest.mark.skipif(not os.path.exists("/root/reference/cross-sections/FSCDXS"), reason="reference tree not present")
def test_cpp_xsread_on_the_reference_fscdxs():
    regs = api.host_xsread("/root/reference/cross-sections", ["HNO3", "F12"], 0.1, 900.0)
    _same(regs, xsfile.read_regions("/root/reference/cross-sections", ["HNO3", "F12"], 0.1, 900.0))
    assert [(r["ixmol"], r["v1fx"], r["v2fx"], len(r["files"])) for r in regs] == \
        [(0, 0.0, 109.0, 6), (0, 332.0, 984.0, 6), (1, 867.013, 936.987, 6)]
    assert [f["t"] for f in regs[0]["files"]] == [203., 213., 233., 253., 273., 296.] and regs[0]["files"][0]["npts"] == 21801


@pytest.mark.skipif(not os.path.exists("/root/reference/cross-sections/FSCDXS"), reason="reference tree not present")
def test_oracle_on_the_shipped_hno3_microwave_table():
    """The one microwave band of the shipped data (HNO3 0-109 cm-1).  Its table starts at 0 cm-1 where the radiation term
    MONORTM_XSEC_SUB divides out vanishes (0/0): near the surface the resampling step equals the table step and the NaN point
    is never touched; in thinner layers it is, the running sum turns NaN and the reference's GOTO loop never ends
    (monortm_sub.F90:1800-1821) -- the oracle (and the product, MRTM_EXSEC) report that instead of hanging."""
    regs = xsfile.read_regions("/root/reference/cross-sections", ["HNO3"], 0.5, 100.0)
    wn = np.array([0.7417, 30.0])                      # 22.235 GHz and the band maximum
    xamnt = np.zeros((xsfile.MX_XS, 1))
    xamnt[0] = [1e16]
    od = harness.oracle_xsec(regs, wn, np.array([1000.]), np.array([288.]), xamnt)
    assert np.all(np.isfinite(od)) and np.all(od > 0) and od[1, 0] > od[0, 0]
    with pytest.raises(RuntimeError, match="rc=53"):  # a frequency near 0 cm-1 in a thin layer: the sum walks into the NaN point
        harness.oracle_xsec(regs, wn[:1], np.array([300.]), np.array([270.]), xamnt)


@pytest.mark.gpu
def test_gpu_cross_sections_against_the_reference_text_and_the_oracle(tmp_path):
    g = np.load(os.path.join(GOLD, "ref_xsec_synth.npz"))
    spec = json.loads(str(g["spec"]))
    regs, wn, p, t, xamnt = harness.xsec_case(spec, str(tmp_path))
    s = harness.session()
    s.stage_xsec(api.host_xsread(str(tmp_path), spec["names"], float(wn.min()), float(wn.max())))
    od = s.xsec(wn, p, t, xamnt)
    assert np.array_equal(od == 0, g["odxsec"] == 0)
    assert harness.rel_diff(od, g["odxsec"]) < 1e-9                           # bar: 1e-9 (measured ~1e-15)
    # a denser list through the oracle
    wn2 = np.sort(np.concatenate([np.linspace(1.9, 12.1, 300), np.linspace(17.9, 21.1, 100), np.linspace(39.9, 44.1, 100)]))
    ref = harness.oracle_xsec(regs, wn2, p, t, xamnt)
    od2 = s.xsec(wn2, p, t, xamnt)
    assert np.array_equal(od2 == 0, ref == 0) and harness.rel_diff(od2, ref) < 1e-9
    # MODM with IXSECT=1 (modm.f90:197-198, :268): odxsec is computed inside and enters the total
    case = harness.make_case(n_filler=64, nlay=len(p), wn=wn, irt=1)
    pr = case["prof"]
    s.stage_lines(case["ls"])
    a = s.modm(wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0],
               case["scor"][:, :, :, 0], ixsect=1, xamnt=xamnt, want_by_mol=False)
    b = s.modm(wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0],
               case["scor"][:, :, :, 0], want_by_mol=False)
    odx = harness.oracle_xsec(regs, wn, pr["p"][:, 0], pr["t"][:, 0], xamnt)
    assert harness.rel_diff(a["odxsec"], odx, floor=1e-300) < 1e-9 and np.all(b["odxsec"] == 0)
    assert np.max(np.abs(a["o"] - (b["o"] + odx)) / np.abs(a["o"])) < 1e-13
    # the reference's array bound xspd_int(0:10000000) (monortm_sub.F90:1755): a 100 cm-1 wide region in a layer far below
    # the table pressure needs 1.5e7 resampled points -- refused by the product and by the oracle alike
    wide = [dict(ixmol=0, v1fx=10., v2fx=110., xdoplr=1e-5,
                 files=[dict(v1x=10., v2x=110., npts=1001, t=250., pres=250., data=np.full(1001, 1e-20))])]
    s.stage_xsec(wide)
    assert np.all(s.xsec(wn, np.array([900.0]), np.array([250.0]), xamnt[:, :1])[wn >= 10.] > 0)
    with pytest.raises(api.MonortmError) as e:
        s.xsec(wn, np.array([5.0]), np.array([250.0]), xamnt[:, :1])
    assert e.value.code == 12
    with pytest.raises(RuntimeError, match="rc=52"):
        harness.oracle_xsec(wide, wn, np.array([5.0]), np.array([250.0]), xamnt[:, :1])
    s.stage_xsec([])
    assert np.all(s.xsec(wn, p, t, xamnt) == 0)
