"""The CUDA path (through the C ABI) against the vectors produced by executing the reference's own Fortran text
(tests/golden/ref_case_*.npz; tools/gen_ref_goldens.py).  Bars of BASELINE.json north_star: layer optical depths within
1e-9 relative, brightness temperatures within 1e-5 K."""
import glob
import json
import os

import numpy as np
import pytest

import harness
import ref_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_FILES = sorted(glob.glob(os.path.join(GOLD, "ref_case_*.npz")))


@pytest.mark.parametrize("path", CASE_FILES, ids=[os.path.basename(p)[9:-4] for p in CASE_FILES])
def test_gpu_reproduces_the_reference_text(path):
    g = np.load(path)
    case = ref_cases.build_case(json.loads(str(g["spec"])))
    assert ref_cases.inputs_digest(case) == str(g["digest"])
    scale = np.abs(g["o"])[:, None, :]
    for mode in (0, 1):
        for scor_on_device in (False, True):
            c = dict(case)
            if scor_on_device:
                c["scor"] = np.zeros_like(case["scor"])     # ignored: run_gpu passes scor=None below
            gpu = run(c, mode, scor_on_device)
            assert harness.rel_diff(gpu["o"], g["o"]) < 1e-9, (mode, scor_on_device)
            assert np.max(np.abs(gpu["o_by_mol"] - g["o_by_mol"]) / scale) < 1e-9
            assert np.max(np.abs(gpu["oc"] - g["oc"]) / scale) < 1e-9
            assert np.max(np.abs(gpu["o_clw"] - g["o_clw"]) / np.abs(g["o"])) < 1e-9
            assert np.max(np.abs(gpu["tb"] - g["tb"])) < 1e-5
            assert np.max(np.abs(gpu["tmr"] - g["tmr"])) < 1e-5
            for k in ("rad", "rup", "rdn", "trtot"):
                assert harness.rel_diff(gpu[k], g[k], floor=1e-300) < 1e-8, k
            assert gpu["tmpsfc"] == float(g["tmpsfc"])


def run(case, line_mode, scor_on_device):
    s = harness.session()
    s.stage_lines(case["ls"])
    pr = case["prof"]
    m = s.modm(case["wn"], case["dvset"], pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], case["nmol"], pr["wkl"][:, :, 0],
               pr["wbrodl"][:, 0], None if scor_on_device else case["scor"][:, :, :, 0], cntnm=case["cntnm"], ibrd=case["ibrd"],
               want_by_mol=True, line_mode=line_mode, sclcpl=case.get("sclcpl", 1.), sclhw=case.get("sclhw", 1.),
               y0res=case.get("y0res", 0.))
    m["tmr"] = s.calctmr(case["wn"], pr["t"][:, 0], pr["tz"][:, 0], m["o"])
    m.update(s.rtm(1, case["irt"], case["wn"], pr["t"][:, 0], pr["tz"][:, 0], m["o"], case["tmpsfc"], case["reflc"], case["emiss"]))
    return m
