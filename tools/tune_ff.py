#!/usr/bin/env python3
"""Far-field parameter tuning: for the current library / environment (MRTM_LIB, MRTM_FF_RATIO, MRTM_FF_S,
MRTM_FF_LEVELS, MRTM_LINES_F) report (a) the accuracy of the default path against the direct path
(line_mode=1, every in-window triple evaluated) on two 8192-frequency blocks of the bench workload and
(b) the device time of the bench step.  One JSON line on stdout.  GPU only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402


def main():
    import torch
    from monortm_b200 import api
    nwn = int(os.environ.get("TUNE_NWN", 125000))
    steps = int(os.environ.get("TUNE_STEPS", 8))
    inp = bench.build_inputs(nwn, 0, 1, bench.N_FILLER)
    sess = api.Session(0)
    nlines = sess.stage_lines(inp["ls"])
    pr = inp["prof"]
    acc = {}
    for name, i0 in (("low", 0), ("mid", 500000)):
        wn = bench.DV * np.arange(i0 + 1, i0 + 8192 + 1, dtype=np.float64)
        res = []
        for mode in (0, 1):
            r = sess.modm(wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0], pr["wbrodl"][:, 0],
                          inp["scor"][:, :, :, 0], want_by_mol=False, selection=False, line_mode=mode,
                          global_range=(inp["v1"], inp["v2"], i0))
            res.append(r["o"])
        acc[name] = float(np.max(np.abs(res[0] - res[1]) / np.abs(res[1])))
    dev = torch.device("cuda", 0)

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).to(dev)
    d = {k: dv(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    d["wn"], d["scor"] = dv(inp["wn"]), dv(inp["scor"])
    d["emiss"], d["reflc"] = dv(inp["emiss"]), dv(inp["reflc"])
    d["tmpsfc"] = torch.tensor([inp["tmpsfc"]], dtype=torch.float64, device=dev)
    outs = torch.zeros(6, nwn, dtype=torch.float64, device=dev)
    ptrs = {k: v.data_ptr() for k, v in d.items()}
    for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
        ptrs[k] = outs[i].data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        sess.profiles_dev(1, nwn, bench.NLAY, 22, 0.0, ptrs, inp["v1"], inp["v2"], inp["iw0"], inp["irt"], stream=stream.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sess.reset_stats()
    tot, lk, rt, de = 0.0, [], [], []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        st = sess.stats()
        lk.append(st["last_lines_kernel_ms"]); rt.append(st["last_rt_kernel_ms"]); de.append(st["last_derive_kernel_ms"])
    st = sess.stats()
    env = {k: v for k, v in os.environ.items() if k.startswith("MRTM_")}
    print(json.dumps({"env": env, "ms_per_step": tot / steps, "lines_ms": float(np.mean(lk)), "rt_ms": float(np.mean(rt)),
                      "derive_ms": float(np.mean(de)), "far_expansions": st["far_expansions"], "direct_evals": st["direct_evals"],
                      "max_rel_far_vs_direct": acc, "evals_per_s": float(nlines) * bench.NLAY * nwn / (tot / steps * 1e-3)}))


if __name__ == "__main__":
    main()
