#!/usr/bin/env python3
"""One step of the bench workload in the direct formulation (mrtm_opts.line_mode=1: every in-window
(line, layer, frequency) triple evaluated per frequency) -- for an ncu capture of near_kernel's streamed loops.
usage: python tools/direct_once.py [nwn]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    from monortm_b200 import api
    nwn = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    inp = bench.build_inputs(nwn, 0, 1, bench.N_FILLER)
    sess = api.Session(0)
    sess.stage_lines(inp["ls"])
    pr = inp["prof"]
    prof = dict(nlay=bench.NLAY, nprof=1, nmol=22, p=pr["p"], t=pr["t"], tz=pr["tz"], clw=pr["clw"], wbrodl=pr["wbrodl"], wkl=pr["wkl"])
    for mode in (1, 1):
        sess.reset_stats()
        r = sess.profiles(inp["wn"], 0.0, prof, inp["scor"], inp["irt"], inp["tmpsfc"], inp["emiss"], inp["reflc"],
                          global_range=(inp["v1"], inp["v2"], inp["iw0"]), line_mode=mode)
        st = sess.stats()
        print("line_mode", mode, "lines group ms", st["last_lines_kernel_ms"], "tb[0]", float(np.asarray(r["tb"]).ravel()[0]))


if __name__ == "__main__":
    main()
