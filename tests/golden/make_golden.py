#!/usr/bin/env python3
"""Generate tests/golden/oracle_case_*.npz: outputs of the CPU oracle (-O0) on seeded synthetic inputs.
They pin the oracle against accidental edits and give the GPU tests a committed target; they are NOT
outputs of the Fortran reference (which cannot be built here).  Re-run only when the oracle is
deliberately changed:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import harness  # noqa: E402
from monortm_b200 import synth  # noqa: E402

CASES = {
    "c1_channels_dn": dict(n_filler=512, nlay=19, wn=synth.freq_c1_channels(), irt=3),
    "c2_sounder_up": dict(n_filler=768, nlay=32, wn=synth.freq_c2_sounder(), irt=1, tmpsfc=285.0, emis=0.9),
    "c5_cloudy_up": dict(n_filler=512, nlay=24, wn=np.linspace(0.0055, 55.0, 96), irt=1, clw=True),
}


def build(name):
    return harness.make_case(**CASES[name])


if __name__ == "__main__":
    for name in CASES:
        ref = harness.run_oracle(build(name))
        np.savez_compressed(os.path.join(HERE, "oracle_case_%s.npz" % name),
                            o=ref["o"], tb=ref["tb"], tmr=ref["tmr"], rad=ref["rad"], trtot=ref["trtot"],
                            rup=ref["rup"], rdn=ref["rdn"], sel_count=ref["sel_count"], sel_hash=ref["sel_hash"],
                            o_by_mol_sum=ref["o_by_mol"].sum(axis=1), oc_sum=ref["oc"].sum(axis=1))
        print(name, "written; max TB %.3f K" % ref["tb"].max())
