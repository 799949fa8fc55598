#!/usr/bin/env python3
"""Timings of the BASELINE configs that are not the bench.py headline (C2 sounder channels, C4 retrieval
ensemble, C5 cloudy 300-layer case) through the public host-buffer API (Session.profiles: H2D + kernels +
D2H inside the timed region).  One JSON line per config; radiance spectra/s is the figure of merit.

usage: python tools/bench_configs.py [--nprof 256] [--n-filler 65536] [--reps 3] [--configs c2,c4,c5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def xs_case(sess):
    """cross-section optical depths (mrtm_xsec, host buffers): the seeded synthetic set, 4000 frequencies x 100 layers"""
    import tempfile
    import harness  # noqa: F401
    from monortm_b200 import api, synth, xsfile
    with tempfile.TemporaryDirectory() as d:
        xsfile.synthetic_set(d)
        wn = np.linspace(2.0, 44.0, 4000)
        sess.stage_xsec(api.host_xsread(d, ["HNO3", "F11"], float(wn[0]), float(wn[-1])))
        prof = synth.synthetic_profiles(1, 100, seed0=1000, clw_layers=False, nmol=22)
        keep = prof["p"][:, 0] > 260.0                     # above the table pressures: the reference's usable domain
        p, t = prof["p"][keep, 0], prof["t"][keep, 0]
        xamnt = np.zeros((38, len(p)), order="F")
        xamnt[0], xamnt[1] = 1e15, 1e13
        sess.xsec(wn, p, t, xamnt)
        t0 = time.time()
        od = sess.xsec(wn, p, t, xamnt)
        dt = time.time() - t0
        print(json.dumps({"config": "xs", "what": "mrtm_xsec: 3 regions (2 molecules) x %d frequencies x %d layers, host buffers" % (len(wn), len(p)),
                          "s_per_call": dt, "frequency_layer_pairs_per_s": len(wn) * len(p) / dt, "max_od": float(od.max())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nprof", type=int, default=256)
    ap.add_argument("--n-filler", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="c2,c4,c5")
    ap.add_argument("--line-mode", type=int, default=0)
    args = ap.parse_args()
    import harness
    from monortm_b200 import api, synth

    sess = api.Session(0)
    ls = harness.synthetic_store(args.n_filler, v1=0.0, v2=55.0)
    nlines = sess.stage_lines(ls)
    cases = {
        "c2": dict(wn=synth.freq_c2_sounder(), nlay=100, nprof=args.nprof, irt=1, clw=False,
                   what="C2 sounder channels x 100 layers x %d profiles, IRT=1" % args.nprof),
        "c4": dict(wn=synth.freq_c4_channels(1000), nlay=100, nprof=args.nprof, irt=1, clw=False,
                   what="C4 ensemble: 1000 log-spaced channels 0.1-30 cm-1 x 100 layers x %d profiles, IRT=1" % args.nprof),
        "c5": dict(wn=synth.freq_c5(10000), nlay=300, nprof=1, irt=1, clw=True,
                   what="C5 cloudy: 10000 frequencies x 300 layers (3 cloud layers), IRT=1 then IRT=3"),
    }
    for name in args.configs.split(","):
        if name == "xs":
            xs_case(sess)
            continue
        c = cases[name]
        wn = np.asarray(c["wn"], dtype=np.float64)
        prof = synth.synthetic_profiles(c["nprof"], c["nlay"], seed0=1000, clw_layers=c["clw"], nmol=22)
        nwn = len(wn)
        em, rf = np.full(nwn, 0.9), np.full(nwn, 0.1)
        irts = (1, 3) if name == "c5" else (c["irt"],)

        def step():
            for irt in irts:
                sess.profiles(wn, 0.0, prof, None, irt, 288.2, em, rf, line_mode=args.line_mode)     # scor=None: TIPS on the device
        step()
        ts, kms = [], []
        for _ in range(args.reps):
            t0 = time.time()
            step()
            ts.append(time.time() - t0)
            st = sess.stats()
            kms.append((st["last_derive_kernel_ms"], st["last_lines_kernel_ms"], st["last_rt_kernel_ms"]))
        dt = min(ts)
        nspec = c["nprof"] * len(irts)
        k = kms[int(np.argmin(ts))]
        print(json.dumps({"config": name, "what": c["what"], "logical_lines": nlines, "line_mode": args.line_mode,
                          "nwn": nwn, "nlay": c["nlay"], "nprof": c["nprof"], "s_per_call": dt,
                          "spectra_per_s": nspec / dt, "nominal_evals_per_s": float(nlines) * c["nlay"] * nwn * nspec / dt,
                          "last_call_kernel_ms": {"derive": k[0], "lines": k[1], "rt": k[2]}}))
    sess.close()


if __name__ == "__main__":
    main()
