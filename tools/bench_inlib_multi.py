#!/usr/bin/env python3
"""One host process, N GPUs behind one context (mrtm_init_multi): the C3 sweep through the host-buffer call mrtm_profiles
(H2D + kernels + D2H inside the timed region), N x 125000 frequencies for N GPUs (weak scaling), frequency blocks chosen by
the library (time feedback).  One JSON line per N.   usage: python tools/bench_inlib_multi.py [--max-gpus 8] [--steps 10]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gpus", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--nwn-per-gpu", type=int, default=125000)
    ap.add_argument("--n-filler", type=int, default=65536)
    args = ap.parse_args()
    import torch
    import harness
    from monortm_b200 import api, synth
    ndev = min(torch.cuda.device_count(), args.max_gpus)
    dv = 5.5e-5
    ls = harness.synthetic_store(args.n_filler, v1=dv, v2=dv * 1000000)
    prof = synth.synthetic_profiles(1, 100, seed0=1000, clw_layers=False, nmol=22)
    scor = api.scor_for_layers(22, prof["t"])
    n = 1
    ref_tb = None
    while n <= ndev:
        nwn = args.nwn_per_gpu * n
        i0 = (1000000 - nwn) // 2
        wn = dv * np.arange(i0 + 1, i0 + nwn + 1, dtype=np.float64)
        em, rf = np.full(nwn, 0.9), np.full(nwn, 0.1)
        sess = api.Session(device_mask=(1 << n) - 1)
        nlines = sess.stage_lines(ls)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hw, he, hr = pin(wn), pin(em), pin(rf)
        hout_t = torch.zeros(6, nwn, dtype=torch.float64).pin_memory()
        hout = {k: hout_t[i].numpy().reshape(nwn, 1, order="F") for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn"))}
        gr = (dv, dv * 1000000, i0)
        ts = []
        for it in range(args.steps + 4):            # the first calls settle the partition
            t0 = time.time()
            sess.profiles(hw, 0.0, prof, scor, 1, 288.2, he, hr, global_range=gr, out=hout)
            ts.append(time.time() - t0)
        ms = 1e3 * float(np.mean(ts[4:]))
        line = {"what": "in-library multi-GPU (mrtm_init_multi), one host process, host buffers in and out", "n_gpus": n, "nwn": nwn,
                "ms_per_step_e2e": ms, "evals_per_s_e2e": float(nlines) * 100 * nwn / (ms * 1e-3), "first_calls_ms": [1e3 * x for x in ts[:4]],
                "logical_lines": nlines}
        if n == 1:
            ref_tb = hout["tb"][:, 0].copy()
            ref_i0 = i0
        else:                                       # the single-GPU block is a sub-range of this one: same spectra
            a = ref_i0 - i0
            line["max_dtb_vs_1gpu_K"] = float(np.max(np.abs(hout["tb"][a:a + len(ref_tb), 0] - ref_tb)))
        print(json.dumps(line), flush=True)
        sess.close()
        n *= 2


if __name__ == "__main__":
    main()
