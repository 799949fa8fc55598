#!/usr/bin/env python3
"""bench.py -- headline benchmark of the monoRTM hot path on B200 (contract: see DESIGN.md section 6).

Workload (config.workload): the per-GPU shard of BASELINE config 3, the dense monochromatic sweep
0-55 cm-1 (wn_i = 5.5e-5*i) x 100 layers, sharded by frequency.  Each GPU takes a contiguous block of
`--nwn-per-gpu` frequencies of the global grid (weak scaling: N GPUs cover N blocks; 8 x 125000 is the
full 1e6-point config), runs MODM + CALCTMR + RTM for it and the ranks all-gather the six spectra
with NCCL.  Line list: TAPE3-synth "fast-like" (4096 filler + physical seed lines).

metric  : line x layer x frequency evaluations per second (nominal triples)
value   : device-resident (inputs in HBM), CUDA-event timed, max over ranks
e2e     : the same through the public host-buffer C ABI call (mrtm_profiles), H2D + D2H inside
--impl reference : the CPU oracle (the C restatement of the reference; the Fortran reference cannot
          be compiled in this image) farmed over all host cores on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NLAY = 100
N_FILLER = 65536                  # TAPE3-synth "full-like" list; --n-filler 4096 = "fast-like"
DV = 5.5e-5
NWN_GLOBAL_FULL = 1000000
FLOP_PER_INWINDOW_EVAL = 12.0     # Lorentz, no coupling, hoisted per-(line,layer) terms
FLOP_PER_FAR_EXPANSION = 59.0     # 11 set-up + 12 x (mul + fma + add) for the 14-term Taylor series


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nwn-per-gpu", type=int, default=int(os.environ.get("MRTM_BENCH_NWN", 125000)))
    ap.add_argument("--n-filler", type=int, default=N_FILLER, help="synthetic filler lines (65536 full-like, 4096 fast-like)")
    ap.add_argument("--cpu-sample-nwn", type=int, default=32)
    ap.add_argument("--direct-steps", type=int, default=3, help="extra steps timed with line_mode=1 (direct evaluation)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: equal frequency counts per GPU instead of equal measured cost")
    ap.add_argument("--config", default="c3", choices=["c3", "c4"],
                    help="c3 (default, the headline): dense sweep sharded by frequency; c4: retrieval ensemble "
                         "(1000 channels x 100 layers x --nprof-per-gpu profiles) sharded by profile")
    ap.add_argument("--nprof-per-gpu", type=int, default=1024, help="profiles per GPU of --config c4")
    ap.add_argument("--nchan", type=int, default=1000, help="channels of --config c4 (log-spaced 0.1-30 cm-1)")
    return ap.parse_args()


def oracle_only_store(n_filler, v1, v2):
    """TAPE3-synth -> file -> the ORACLE's GET_LNFL: no product code on this path (reference arm)."""
    import tempfile
    import harness
    from monortm_b200 import linefile, synth          # pure-Python writer of the synthetic TAPE3; loads no library
    recs = synth.synthetic_records(n_filler)
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        return harness.oracle_read_tape3(path, v1, v2, iim=len(recs) + 8)
    finally:
        os.unlink(path)


def build_inputs(nwn, rank, world, n_filler=N_FILLER, oracle_only=False, iw0=None):
    """Synthetic C3 shard: a contiguous block of `nwn` frequencies of the GLOBAL 1e6-point grid, centred
    in this rank's 1/world-th of the grid; one 100-layer profile.  v1, v2 are the global range.
    oracle_only: line file and TIPS through the oracle's own readers (the reference arm loads no product library)."""
    import harness
    from monortm_b200 import synth
    v1, v2 = DV * 1, DV * NWN_GLOBAL_FULL
    part = NWN_GLOBAL_FULL // world
    if iw0 is None:
        iw0 = rank * part + max(0, (part - nwn) // 2)
    wn = DV * np.arange(iw0 + 1, iw0 + nwn + 1, dtype=np.float64)
    prof = synth.synthetic_profiles(1, NLAY, seed0=1000, clw_layers=False, nmol=22)
    if oracle_only:
        ls = oracle_only_store(n_filler, v1, v2)
        scor = harness.oracle_scor_for_layers(22, prof["t"])
    else:
        from monortm_b200 import api
        ls = harness.synthetic_store(n_filler, v1=v1, v2=v2)
        scor = api.scor_for_layers(22, prof["t"])
    return dict(wn=wn, ls=ls, prof=prof, scor=scor, v1=v1, v2=v2, iw0=iw0,
                emiss=np.full(nwn, 0.9), reflc=np.full(nwn, 0.1), tmpsfc=288.2, irt=1)


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed region through NVML (10 ms period;
    `nvidia-smi` polling is too slow for a sub-second timed region)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.err = gpu, [], False, None
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[0].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            R = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(mhz), pw, [k for k, bit in R.items() if rs & bit]))
                time.sleep(0.01)
        except Exception as e:          # fall back to nvidia-smi polling
            self.err = repr(e)
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            while not self.stop_flag:
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                    f = [x.strip() for x in out.split(",")]
                    self.max_mhz = float(f[1])
                    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                    self.rows.append((float(f[0]), float(f[2]), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
                except Exception:
                    pass
                time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"], "sampler_error": self.err}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(r[1] for r in self.rows)}


def cpu_sample(inp, nwn_sample, opt="O0", keep=None):
    """Time the CPU oracle on `nwn_sample` evenly spaced frequencies of this shard (all layers, all lines)."""
    import harness
    idx = np.linspace(0, len(inp["wn"]) - 1, nwn_sample).astype(int)
    wn = inp["wn"][idx]
    pr = inp["prof"]
    t0 = time.time()
    m = harness.oracle_modm(inp["ls"], wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0],
                            pr["wbrodl"][:, 0], inp["scor"][:, :, :, 0], opt=opt)
    tmr = harness.oracle_calctmr(wn, pr["t"][:, 0], pr["tz"][:, 0], m["o"])
    r = harness.oracle_rtm(1, inp["irt"], wn, pr["t"][:, 0], pr["tz"][:, 0], m["o"], inp["tmpsfc"], inp["reflc"][idx], inp["emiss"][idx])
    dt = time.time() - t0
    nlines = logical_lines(inp["ls"])
    if keep is not None:
        keep.update(idx=idx, o=m["o"], sel_count=m["sel_count"], sel_hash=m["sel_hash"], tmr=tmr, tb=r["tb"], rad=r["rad"])
    return dt, float(nlines) * NLAY * nwn_sample, float(m["sel_count"].sum())


def parity_check(sess, inp, hprof, hscor, oracle):
    """The bench workload itself against the oracle (outside the timed region): the full shard runs once more through
    the public host-buffer call with layer optical depths and the selection instrumentation switched on, in the default
    mode (far-field expansion) and in the direct mode, and is compared with the oracle at the cpu_baseline sample
    frequencies.  Bars (BASELINE.json north_star): selected-line set bit-exact, OD 1e-9 relative, TB (and TMR) 1e-5 K."""
    idx = oracle["idx"]
    out = {"n_freq": int(len(idx)), "nlay": NLAY, "line_modes": [0, 1], "od_rtol_bar": 1e-9, "tb_atol_K_bar": 1e-5,
           "max_rel_od": 0.0, "max_dtb_K": 0.0, "max_dtmr_K": 0.0, "max_rel_rad": 0.0, "sel_exact": True,
           "what": "GPU (mrtm_profiles, whole shard, both line modes) vs oracle at the cpu_baseline sample frequencies x all layers"}
    for mode in (0, 1):
        g = sess.profiles(inp["wn"], 0.0, hprof, hscor, inp["irt"], inp["tmpsfc"], inp["emiss"], inp["reflc"],
                          global_range=(inp["v1"], inp["v2"], inp["iw0"]), want_o=True, selection=True, line_mode=mode)
        o = g["o"][idx, :, 0]
        out["max_rel_od"] = max(out["max_rel_od"], float(np.max(np.abs(o - oracle["o"]) / np.abs(oracle["o"]))))
        out["max_dtb_K"] = max(out["max_dtb_K"], float(np.max(np.abs(g["tb"][idx, 0] - oracle["tb"]))))
        out["max_dtmr_K"] = max(out["max_dtmr_K"], float(np.max(np.abs(g["tmr"][idx, 0] - oracle["tmr"]))))
        out["max_rel_rad"] = max(out["max_rel_rad"], float(np.max(np.abs(g["rad"][idx, 0] / oracle["rad"] - 1.0))))
        out["sel_exact"] = bool(out["sel_exact"] and np.array_equal(g["sel_count"][idx, :, 0], oracle["sel_count"])
                                and np.array_equal(g["sel_hash"][idx, :, 0], oracle["sel_hash"]))
        del g
    out["selected_pairs_compared"] = int(oracle["sel_count"].sum()) * 2
    out["pass"] = bool(out["sel_exact"] and out["max_rel_od"] < 1e-9 and out["max_dtb_K"] < 1e-5 and out["max_dtmr_K"] < 1e-5)
    return out


def logical_lines(ls):
    """Logical lines = records that are not coupling-coefficient records (SURVEY 8d unit of work)."""
    n = 0
    for i in range(39):
        k = int(ls.nblm[i])
        xg = ls.xg[i, :k]
        j = 0
        while j < k:
            n += 1
            j += 2 if xg[j] in (-1.0, -3.0, -5.0) else 1
    return n


_FARM = {}          # inputs prepared once in the parent; the forked workers inherit them


def _farm_worker(args):
    lo, hi, opt = args
    inp = _FARM["inp"]
    sub = dict(inp)
    sub["wn"] = inp["wn"][lo:hi]
    sub["emiss"], sub["reflc"] = inp["emiss"][lo:hi], inp["reflc"][lo:hi]
    dt, nominal, inwin = cpu_sample(sub, hi - lo, opt=opt)
    return dt, nominal, inwin


def farm_rate(pool, chunks, opt, reps):
    """evals/s of `len(chunks)` worker processes running disjoint frequency chunks at once (wall clock of the map)"""
    best, nominal = None, 0.0
    for _ in range(reps):
        t0 = time.time()
        res = pool.map(_farm_worker, [(lo, hi, opt) for lo, hi in chunks])
        dt = time.time() - t0
        nominal = sum(r[1] for r in res)
        best = dt if best is None else min(best, dt)
    return nominal / best, best, [r[0] for r in res]


def run_reference(args):
    """Reference arm: the oracle (C restatement of the reference, pinned against the executed Fortran text; -O0 like
    linuxGNUdbl) farmed over all host cores.  Each step = `cores` disjoint chunks of frequencies of the rank-0 shard.
    Inputs (TAPE3-synth through the oracle's GET_LNFL, profile, the oracle's TIPS_2003) are prepared ONCE in the parent,
    outside the timed region; no product library is loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    nwn = args.nwn_per_gpu
    per = max(2, args.cpu_sample_nwn // 4)
    stride = max(per, (nwn - per) // max(cores, 1))
    chunks = [((c * stride) % (nwn - per), (c * stride) % (nwn - per) + per) for c in range(cores)]
    _FARM["inp"] = build_inputs(nwn, 0, args.gpus, args.n_filler, oracle_only=True)
    import harness
    harness.oracle_lib("O0"), harness.oracle_lib("O2")          # dlopen before forking
    times = []
    nominal = 0.0
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.time()
            res = pool.map(_farm_worker, [(lo, hi, "O0") for lo, hi in chunks])
            dt = time.time() - t0
            if it >= args.warmup:
                times.append(dt)
                nominal = sum(r[1] for r in res)
        in_worker = float(np.mean([r[0] for r in res]))
        o2_farm, _, _ = farm_rate(pool, chunks, "O2", 2)
    # single-process figures on the same chunk size (what one core does when it has the machine to itself)
    single = {}
    for opt in ("O0", "O2"):
        sub = dict(_FARM["inp"])
        lo, hi = chunks[0]
        sub["wn"], sub["emiss"], sub["reflc"] = sub["wn"][lo:hi], sub["emiss"][lo:hi], sub["reflc"][lo:hi]
        dt, nom, _ = cpu_sample(sub, hi - lo, opt=opt)
        single[opt] = nom / dt
    ms = 1e3 * float(np.mean(times))
    val = nominal / (ms * 1e-3)
    sample = "%d processes x %d frequencies x %d layers x all lines per step (oracle -O0, process farm)" % (cores, per, NLAY)
    line = {"metric": "line x layer x frequency evaluations/s", "value": val, "unit": "evals/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                             "single_process_O0": single["O0"], "single_process_O2": single["O2"], "farm_O2": o2_farm,
                             "per_core_in_farm_O0": val / cores, "mean_worker_seconds": in_worker,
                             "note": "the per-core rate inside the farm is below the single-process rate when the cores are SMT "
                                     "siblings / share memory bandwidth: os.cpu_count() counts hardware threads; input preparation "
                                     "is outside the timed region; product library not loaded: %s" %
                                     (not any("libmonortm_b200" in l for l in open("/proc/self/maps")))},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "C3 dense monochromatic sweep 0-55 cm-1 (wn_i=5.5e-5*i), frequency-sharded: %d frequencies/GPU "
                        "(a contiguous block centred in each rank's 1/N of the global 1e6-point grid) x %d layers x TAPE3-synth "
                        "%s line list (%d filler + physical seed lines), IRT=1, MODM+CALCTMR+RTM per step"
                        % (args.nwn_per_gpu, NLAY, "full-like" if args.n_filler >= 65536 else "fast-like", args.n_filler),
            "n_filler": args.n_filler,
            "nwn_per_gpu": args.nwn_per_gpu, "nlay": NLAY, "sharding": "frequency", "parallelism": "freq-shard x%d" % args.gpus,
            "l2": "L2 flushed (256 MiB write) between timed steps"}


def ncu_traffic(args):
    """DRAM bytes the line-path kernels move per step, from the committed ncu capture of the default workload."""
    try:
        if args.nwn_per_gpu != 125000 or args.n_filler != N_FILLER:
            return None
        with open(os.path.join(ROOT, "profiles", "r02_v34_traffic.json")) as f:
            return float(json.load(f)["lines_group_bytes_per_step"])
    except Exception:
        return None


# =====================================================================================================================
# --config c4: the retrieval ensemble of BASELINE config 4 (profile axis; monortm.f90:357 profile loop)
# =====================================================================================================================
C4_V1, C4_V2 = 0.1, 30.0


def c4_config(args):
    return {"workload": "C4 retrieval ensemble: %d log-spaced channels 0.1-30 cm-1 x %d layers x %d profiles/GPU (profile-sharded, "
                        "profiles-synth seeds 1000+global index) x TAPE3-synth %s line list (%d filler + physical seed lines), IRT=1, "
                        "MODM+CALCTMR+RTM per step, partition sums on the device (scor=NULL)"
                        % (args.nchan, NLAY, args.nprof_per_gpu, "full-like" if args.n_filler >= 65536 else "fast-like", args.n_filler),
            "n_filler": args.n_filler, "nwn": args.nchan, "nlay": NLAY, "nprof_per_gpu": args.nprof_per_gpu,
            "sharding": "profile", "parallelism": "profile-shard x%d" % args.gpus,
            "l2": "inputs + derived planes of one step (> 1 GB) exceed L2; no flush needed"}


def c4_inputs(args, rank, oracle_only=False, nprof=None):
    import harness
    from monortm_b200 import synth
    wn = synth.freq_c4_channels(args.nchan)
    nprof = args.nprof_per_gpu if nprof is None else nprof
    prof = synth.synthetic_profiles(nprof, NLAY, seed0=1000 + rank * args.nprof_per_gpu, clw_layers=False, nmol=22)
    if oracle_only:
        ls = oracle_only_store(args.n_filler, float(wn[0]), float(wn[-1]))
    else:
        ls = harness.synthetic_store(args.n_filler, v1=float(wn[0]), v2=float(wn[-1]))
    return dict(wn=wn, ls=ls, prof=prof, emiss=np.full(len(wn), 0.9), reflc=np.full(len(wn), 0.1), tmpsfc=288.2, irt=1)


def c4_oracle_sample(inp, ips, idx, opt="O0", keep=None):
    """The oracle on profiles `ips` x channels `idx` x all layers x all lines; scor from the oracle's own TIPS_2003."""
    import harness
    pr = inp["prof"]
    wn = inp["wn"][idx]
    t0 = time.time()
    res = []
    for ip in ips:
        scor = harness.oracle_scor_for_layers(22, pr["t"][:, ip])
        m = harness.oracle_modm(inp["ls"], wn, 0.0, pr["p"][:, ip], pr["t"][:, ip], pr["clw"][:, ip], 22, pr["wkl"][:, :, ip],
                                pr["wbrodl"][:, ip], scor, opt=opt)
        tmr = harness.oracle_calctmr(wn, pr["t"][:, ip], pr["tz"][:, ip], m["o"])
        r = harness.oracle_rtm(1, inp["irt"], wn, pr["t"][:, ip], pr["tz"][:, ip], m["o"], inp["tmpsfc"], inp["reflc"][idx], inp["emiss"][idx])
        res.append(dict(o=m["o"], sel_count=m["sel_count"], sel_hash=m["sel_hash"], tmr=tmr, tb=r["tb"], rad=r["rad"]))
    dt = time.time() - t0
    if keep is not None:
        keep["res"] = res
    return dt, float(logical_lines(inp["ls"])) * NLAY * len(idx) * len(ips)


def _c4_farm_worker(a):
    ip, lo, hi, opt = a
    dt, nominal = c4_oracle_sample(_FARM["inp"], [ip], np.arange(lo, hi), opt=opt)
    return dt, nominal


def run_c4_reference(args):
    """Reference arm of --config c4: the oracle farmed over the host cores, one profile x a few channels per worker."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    import harness
    cores = os.cpu_count() or 1
    per = 8
    _FARM["inp"] = c4_inputs(args, 0, oracle_only=True, nprof=min(cores, args.nprof_per_gpu))
    harness.oracle_lib("O0"), harness.oracle_lib("O2")
    npf = _FARM["inp"]["prof"]["nprof"]
    jobs = [(c % npf, (c * 37) % (args.nchan - per), (c * 37) % (args.nchan - per) + per, "O0") for c in range(cores)]
    times, nominal = [], 0.0
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.time()
            res = pool.map(_c4_farm_worker, jobs)
            dt = time.time() - t0
            if it >= args.warmup:
                times.append(dt)
                nominal = sum(r[1] for r in res)
        t0 = time.time()
        res2 = pool.map(_c4_farm_worker, [j[:3] + ("O2",) for j in jobs])
        o2_farm = sum(r[1] for r in res2) / (time.time() - t0)
    ms = 1e3 * float(np.mean(times))
    val = nominal / (ms * 1e-3)
    per_spec = float(logical_lines(_FARM["inp"]["ls"])) * NLAY * args.nchan
    line = {"metric": "line x layer x frequency evaluations/s", "value": val, "unit": "evals/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": c4_config(args),
            "spectra_per_s": val / per_spec,
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "farm_O2": o2_farm,
                             "sample": "%d processes x (1 profile x %d channels x %d layers x all lines) per step (oracle -O0, "
                                       "process farm; inputs prepared once in the parent; no product library loaded: %s)"
                                       % (cores, per, NLAY, not any("libmonortm_b200" in l for l in open("/proc/self/maps")))},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_c4(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib_path = os.environ.get("MRTM_LIB", os.path.join(ROOT, "monortm_b200", "lib", "libmonortm_b200.so"))
    if not os.path.exists(lib_path) and world == 1:
        import __graft_entry__
        __graft_entry__.build()
    from monortm_b200 import api
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    inp = c4_inputs(args, rank)
    wn, pr = inp["wn"], inp["prof"]
    nwn, nprof = len(wn), pr["nprof"]
    sess = api.Session(local)
    nlines = sess.stage_lines(inp["ls"])

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).to(dev)
    d = {k: dv(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    d["wn"], d["emiss"], d["reflc"] = dv(wn), dv(inp["emiss"]), dv(inp["reflc"])
    d["tmpsfc"] = torch.full((nprof,), inp["tmpsfc"], dtype=torch.float64, device=dev)
    outs = torch.zeros(6, nprof, nwn, dtype=torch.float64, device=dev)       # each spectrum block is (nwn, nprof) column-major
    gathered = torch.zeros(world, 6, nprof, nwn, dtype=torch.float64, device=dev) if world > 1 else None
    ptrs = {k: v.data_ptr() for k, v in d.items()}
    for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
        ptrs[k] = outs[i].data_ptr()
    stream = torch.cuda.Stream(device=dev)       # the library call is asynchronous on the stream it is given: time on that stream
    torch.cuda.set_stream(stream)

    def step_dev():
        sess.profiles_dev(nprof, nwn, NLAY, 22, 0.0, ptrs, float(wn[0]), float(wn[-1]), 0, inp["irt"], stream=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sess.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    lines_ms, rt_ms, derive_ms, far_exp, direct_ev = [], [], [], [], []
    barrier()
    for i in range(args.steps):
        ev[i][0].record(stream)
        step_dev()
        ev[i][1].record(stream)
        st = sess.stats()
        lines_ms.append(st["last_lines_kernel_ms"]); rt_ms.append(st["last_rt_kernel_ms"]); derive_ms.append(st["last_derive_kernel_ms"])
        far_exp.append(st["far_expansions"]); direct_ev.append(st["direct_evals"])
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = sess.stats()["kernel_launches"]
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    nominal_per_step = float(nlines) * NLAY * nwn * nprof * world
    value = nominal_per_step / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers through mrtm_profiles; the spectra land in pinned host memory
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).pin_memory()
    hp = {k: pin(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    hp["wn"], hp["emiss"], hp["reflc"] = pin(wn), pin(inp["emiss"]), pin(inp["reflc"])
    hprof = dict(nlay=NLAY, nprof=nprof, nmol=22,
                 p=hp["p"].numpy().reshape(NLAY, nprof, order="F"), t=hp["t"].numpy().reshape(NLAY, nprof, order="F"),
                 tz=hp["tz"].numpy().reshape(NLAY + 1, nprof, order="F"), clw=hp["clw"].numpy().reshape(NLAY, nprof, order="F"),
                 wbrodl=hp["wbrodl"].numpy().reshape(NLAY, nprof, order="F"), wkl=hp["wkl"].numpy().reshape(39, NLAY, nprof, order="F"))
    h2d_bytes = sum(hp[k].numel() * 8 for k in hp)
    d2h_bytes = 6 * nwn * nprof * 8
    hout_t = torch.zeros(6, nprof, nwn, dtype=torch.float64).pin_memory()
    hout = {k: hout_t[i].numpy().reshape(-1).reshape(nwn, nprof, order="F") for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn"))}

    def step_e2e():
        return sess.profiles(hp["wn"].numpy(), 0.0, hprof, None, inp["irt"], inp["tmpsfc"], hp["emiss"].numpy(), hp["reflc"].numpy(), out=hout)
    step_e2e()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.time()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = 1e3 * (time.time() - t0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nominal_per_step / (float(te.item()) * 1e-3)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    failed = False
    if rank == 0:
        fp64_peak = sess.fp64_peak_tflops()
        lk_ms = float(np.mean(lines_ms))
        far_n, dir_n = float(np.mean(far_exp)), float(np.mean(direct_ev))
        achieved = (far_n * FLOP_PER_FAR_EXPANSION + dir_n * FLOP_PER_INWINDOW_EVAL) / (lk_ms * 1e-3) / 1e12
        line = {"metric": "line x layer x frequency evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": c4_config(args),
                "spectra_per_s": nprof * world / (ms_per_step * 1e-3), "logical_lines": nlines,
                "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": float(te.item()), "spectra_per_s": nprof * world / (float(te.item()) * 1e-3), "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(),
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": achieved / fp64_peak if fp64_peak else None, "traffic": None,
                             "kernel": "plan+far+near+voigt+final (the line path of one step; near_kernel's streamed direct loops dominate "
                                       "on sparse channels: a 128-channel tile spans several cm-1, so nearly every in-window line is near)",
                             "peak_source": "measured live: mrtm_fp64_peak DFMA probe (MEASURED_PEAKS.json has no FP64 figure)",
                             "flop_per_direct_eval": FLOP_PER_INWINDOW_EVAL, "flop_per_far_expansion": FLOP_PER_FAR_EXPANSION,
                             "far_expansions_per_launch": far_n, "direct_evals_per_launch": dir_n,
                             "kernel_ms": lk_ms, "share_of_step": lk_ms / ms_per_step},
                "derive_kernel_ms": float(np.mean(derive_ms)), "rt_kernel_ms": float(np.mean(rt_ms))}
        if not args.no_cpu_baseline:
            # parity on a profile x channel subsample, then the CPU baseline figures
            ips = sorted({0, nprof // 2, nprof - 1})[:3]
            idx = np.linspace(0, nwn - 1, 12).astype(int)
            keep = {}
            dt, nominal = c4_oracle_sample(inp, ips, idx, keep=keep)
            sub = dict(pr)
            for k in ("p", "t", "tz", "clw", "wbrodl"):
                sub[k] = np.asfortranarray(pr[k][:, ips])
            sub["wkl"] = np.asfortranarray(pr["wkl"][:, :, ips])
            sub["nprof"] = len(ips)
            pc = {"profiles": [int(i) for i in ips], "n_channels": int(len(idx)), "nlay": NLAY, "od_rtol_bar": 1e-9, "tb_atol_K_bar": 1e-5,
                  "max_rel_od": 0.0, "max_dtb_K": 0.0, "max_dtmr_K": 0.0, "sel_exact": True, "line_modes": [0, 1],
                  "what": "GPU (mrtm_profiles on the full channel list, device TIPS, both line modes) vs oracle (own TIPS_2003) on a "
                          "profile x channel subsample x all layers"}
            for mode in (0, 1):
                g = sess.profiles(wn, 0.0, sub, None, inp["irt"], inp["tmpsfc"], inp["emiss"], inp["reflc"], want_o=True, selection=True,
                                  line_mode=mode)
                for j, r in enumerate(keep["res"]):
                    o = g["o"][idx, :, j]
                    pc["max_rel_od"] = max(pc["max_rel_od"], float(np.max(np.abs(o - r["o"]) / np.abs(r["o"]))))
                    pc["max_dtb_K"] = max(pc["max_dtb_K"], float(np.max(np.abs(g["tb"][idx, j] - r["tb"]))))
                    pc["max_dtmr_K"] = max(pc["max_dtmr_K"], float(np.max(np.abs(g["tmr"][idx, j] - r["tmr"]))))
                    pc["sel_exact"] = bool(pc["sel_exact"] and np.array_equal(g["sel_count"][idx, :, j], r["sel_count"])
                                           and np.array_equal(g["sel_hash"][idx, :, j], r["sel_hash"]))
                del g
            # the batched device-resident result of the timed steps against the per-profile call
            pc["pass"] = bool(pc["sel_exact"] and pc["max_rel_od"] < 1e-9 and pc["max_dtb_K"] < 1e-5 and pc["max_dtmr_K"] < 1e-5)
            line["parity_check"] = pc
            dt2, nominal2 = c4_oracle_sample(inp, ips[:1], idx, opt="O2")
            line["cpu_baseline"] = {"value": nominal / dt, "unit": "evals/s", "cores": 1, "kind": "port", "value_O2": nominal2 / dt2,
                                    "sample": "%d profiles x %d channels x %d layers x all %d lines, oracle -O0 (C restatement; the Fortran "
                                              "reference cannot be compiled here), %.1f s" % (len(ips), len(idx), NLAY, nlines, dt)}
            failed = not pc["pass"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        raise SystemExit("bench.py: parity_check failed (see the JSON line): the throughput above is not valid")


def balance_shards(args, world, rank, dev, sess, inp0, torch, dist):
    """Per-rank frequency counts of equal measured cost (see the call site).  Collective: every rank returns the same list."""
    counts = [args.nwn_per_gpu] * world
    total = args.nwn_per_gpu * world
    part = NWN_GLOBAL_FULL // world
    pr = inp0["prof"]

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).to(dev)
    d = {k: dv(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    d["scor"] = dv(inp0["scor"])
    d["tmpsfc"] = torch.tensor([inp0["tmpsfc"]], dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    for rnd in range(3):
        n = counts[rank]
        iw0 = sum(counts[:rank]) if total == NWN_GLOBAL_FULL else rank * part + max(0, (part - n) // 2)
        wn = DV * np.arange(iw0 + 1, iw0 + n + 1, dtype=np.float64)
        d["wn"], d["emiss"], d["reflc"] = dv(wn), dv(np.full(n, 0.9)), dv(np.full(n, 0.1))
        outs = torch.zeros(6, n, dtype=torch.float64, device=dev)
        ptrs = {k: v.data_ptr() for k, v in d.items()}
        for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
            ptrs[k] = outs[i].data_ptr()
        with torch.cuda.stream(stream):
            for _ in range(2):
                sess.profiles_dev(1, n, NLAY, 22, 0.0, ptrs, inp0["v1"], inp0["v2"], iw0, inp0["irt"], stream=stream.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                sess.profiles_dev(1, n, NLAY, 22, 0.0, ptrs, inp0["v1"], inp0["v2"], iw0, inp0["irt"], stream=stream.cuda_stream)
            e1.record(stream)
        stream.synchronize()
        sess.sync()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        ts = [float(x.item()) for x in ts]
        rate = [c / max(x, 1e-9) for c, x in zip(counts, ts)]              # frequencies per ms on each rank's part of the grid
        share = [r / sum(rate) for r in rate]
        new = [max(512, int(round(0.5 * (c + s_ * total) / 512.0)) * 512) for c, s_ in zip(counts, share)]      # damped
        new[-1] += total - sum(new)                                        # keep the sum
        counts = new
    return counts


def main():
    args = parse()
    if args.config == "c4":
        return run_c4_reference(args) if args.impl == "reference" else run_c4(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the library normally travels with the tree (built by __graft_entry__.build()); build it here only if it is
    # missing (one rank, before anybody loads it) -- there is no other code path: without it the bench fails loudly
    lib_path = os.environ.get("MRTM_LIB", os.path.join(ROOT, "monortm_b200", "lib", "libmonortm_b200.so"))
    if not os.path.exists(lib_path) and world == 1:
        import __graft_entry__
        __graft_entry__.build()
    from monortm_b200 import api
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    nwn = args.nwn_per_gpu
    inp = build_inputs(nwn, rank, world, args.n_filler)
    sess = api.Session(local)
    nlines = sess.stage_lines(inp["ls"])
    pr = inp["prof"]
    shard_counts = [nwn] * world
    if world > 1 and not args.no_balance:
        # ---- shards of equal COST, not equal count (frequencies are independent; the cost per frequency depends
        # on the spectral position -- more far-field expansions where window edges and negative-frequency resonances fall
        # inside the coarse tiles).  Untimed: every rank times three steps of its block, the times are all-gathered and the
        # block sizes (multiples of 512, their sum stays world x nwn-per-gpu) are rescaled; three rounds.
        shard_counts = balance_shards(args, world, rank, dev, sess, inp, torch, dist)
        part = NWN_GLOBAL_FULL // world
        if sum(shard_counts) == NWN_GLOBAL_FULL:                      # the blocks tile the whole grid
            iw0 = sum(shard_counts[:rank])
        else:                                                         # each block centred in its 1/world-th of the grid
            iw0 = rank * part + max(0, (part - shard_counts[rank]) // 2)
        nwn = shard_counts[rank]
        inp = build_inputs(nwn, rank, world, args.n_filler, iw0=iw0)
    nmax = max(shard_counts)

    # ---- device-resident buffers (torch is only the allocator / stream / NCCL plumbing)
    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).to(dev)
    d = {k: dv(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    d["wn"], d["scor"] = dv(inp["wn"]), dv(inp["scor"])
    d["emiss"], d["reflc"] = dv(inp["emiss"]), dv(inp["reflc"])
    d["tmpsfc"] = torch.tensor([inp["tmpsfc"]], dtype=torch.float64, device=dev)
    outs = torch.zeros(6, nmax, dtype=torch.float64, device=dev)           # rad,tb,tmr,trtot,rup,rdn (rows padded to the largest shard)
    gathered = torch.zeros(world, 6, nmax, dtype=torch.float64, device=dev) if world > 1 else None
    ptrs = {k: v.data_ptr() for k, v in d.items()}
    for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
        ptrs[k] = outs[i].data_ptr()
    # N > 1: the spectra of step i are all-gathered on a second stream while step i+1 computes (two output / gather buffers);
    # a step waits for the gather that last used its buffer, and the last gathers are added to the timed total
    outs_b = torch.zeros_like(outs) if world > 1 else None
    gathered_b = torch.zeros_like(gathered) if world > 1 else None
    ptrs_b = dict(ptrs)
    if world > 1:
        for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn")):
            ptrs_b[k] = outs_b[i].data_ptr()
    gstream = torch.cuda.Stream(device=dev) if world > 1 else None
    gdone = [None, None]
    step_no = [0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # everything timed runs on ONE explicit stream: the library call is asynchronous on the stream it is given (a NULL /
    # legacy-default handle would send it to the context's own stream, outside the events below)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step_dev(line_mode=0):
        b = step_no[0] & 1
        step_no[0] += 1
        if world > 1 and gdone[b] is not None:
            stream.wait_event(gdone[b])                       # the gather that read this output buffer two steps ago
        sess.profiles_dev(1, nwn, NLAY, 22, 0.0, ptrs_b if (world > 1 and b) else ptrs, inp["v1"], inp["v2"], inp["iw0"], inp["irt"],
                          stream=stream.cuda_stream, line_mode=line_mode)
        if world > 1:
            e = torch.cuda.Event()
            e.record(stream)
            gstream.wait_event(e)
            with torch.cuda.stream(gstream):
                dist.all_gather_into_tensor(gathered_b if b else gathered, outs_b if b else outs)
                gdone[b] = torch.cuda.Event()
                gdone[b].record(gstream)

    def drain_gathers():
        """the compute stream waits for the gathers still in flight (timed by the caller)"""
        for g in gdone:
            if g is not None:
                stream.wait_event(g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps (per-step CUDA events, L2 flushed between steps)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    drain_gathers()
    barrier()
    sess.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    lines_ms, rt_ms, derive_ms, far_exp, direct_ev = [], [], [], [], []
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        step_dev()
        ev[i][1].record(stream)
        st = sess.stats()
        lines_ms.append(st["last_lines_kernel_ms"]); rt_ms.append(st["last_rt_kernel_ms"]); derive_ms.append(st["last_derive_kernel_ms"])
        far_exp.append(st["far_expansions"]); direct_ev.append(st["direct_evals"])
    tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    tail[0].record(stream)
    drain_gathers()
    tail[1].record(stream)
    barrier()
    t_wall = time.time() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) + tail[0].elapsed_time(tail[1])
    launches = sess.stats()["kernel_launches"]
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    # every rank's own kernel time per step (line path + derive + RT, without the gather): shows what is left of the imbalance
    own = torch.tensor([float(np.mean(lines_ms)) + float(np.mean(derive_ms)) + float(np.mean(rt_ms))], dtype=torch.float64, device=dev)
    owns = [torch.zeros_like(own) for _ in range(world)]
    if world > 1:
        dist.all_gather(owns, own)
    rank_kernel_ms = [float(x.item()) for x in owns] if world > 1 else [float(own.item())]
    nominal_per_step = float(nlines) * NLAY * float(sum(shard_counts))
    value = nominal_per_step / (ms_per_step * 1e-3)

    # ---- the same steps with every in-window triple evaluated directly (line_mode=1): the classic
    # per-(line,layer,frequency) kernel, reported beside the default path
    direct = None
    if args.direct_steps > 0:
        step_dev(1)
        barrier()
        evd = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.direct_steps)]
        dl = []
        for i in range(args.direct_steps):
            flush.zero_()
            evd[i][0].record(stream)
            step_dev(1)
            evd[i][1].record(stream)
            dl.append(sess.stats()["last_lines_kernel_ms"])
        barrier()
        dms = torch.tensor([sum(a.elapsed_time(b) for a, b in evd) / args.direct_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        direct = {"ms_per_step": float(dms.item()), "lines_kernel_ms": float(np.mean(dl)),
                  "value": nominal_per_step / (float(dms.item()) * 1e-3)}

    # ---- e2e: host buffers through mrtm_profiles (pinned inputs, H2D + D2H inside the timed region)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1, order="F"))).pin_memory()
        return t
    hp = {k: pin(pr[k]) for k in ("p", "t", "tz", "clw", "wkl", "wbrodl")}
    hp["wn"], hp["scor"], hp["emiss"], hp["reflc"] = pin(inp["wn"]), pin(inp["scor"]), pin(inp["emiss"]), pin(inp["reflc"])
    hprof = dict(nlay=NLAY, nprof=1, nmol=22,
                 p=hp["p"].numpy().reshape(NLAY, 1, order="F"), t=hp["t"].numpy().reshape(NLAY, 1, order="F"),
                 tz=hp["tz"].numpy().reshape(NLAY + 1, 1, order="F"), clw=hp["clw"].numpy().reshape(NLAY, 1, order="F"),
                 wbrodl=hp["wbrodl"].numpy().reshape(NLAY, 1, order="F"), wkl=hp["wkl"].numpy().reshape(39, NLAY, 1, order="F"))
    hscor = hp["scor"].numpy().reshape(42, 9, NLAY, 1, order="F")
    h2d_bytes = sum(hp[k].numel() * 8 for k in hp)
    d2h_bytes = 6 * nwn * 8

    # the six spectra land in pinned host memory owned by the caller (reused every step)
    hout_t = torch.zeros(6, nmax, dtype=torch.float64).pin_memory()
    hout = {k: hout_t[i, :nwn].numpy().reshape(nwn, 1, order="F") for i, k in enumerate(("rad", "tb", "tmr", "trtot", "rup", "rdn"))}

    def step_e2e():
        return sess.profiles(hp["wn"].numpy(), 0.0, hprof, hscor, inp["irt"], inp["tmpsfc"], hp["emiss"].numpy(),
                             hp["reflc"].numpy(), global_range=(inp["v1"], inp["v2"], inp["iw0"]), out=hout)
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(1, args.steps)
    t0 = time.time()
    for _ in range(e2e_steps):
        r = step_e2e()
        if world > 1:
            outs.copy_(hout_t)
            dist.all_gather_into_tensor(gathered, outs)
    barrier()
    e2e_ms = 1e3 * (time.time() - t0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nominal_per_step / (float(te.item()) * 1e-3)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    parity_failed = False
    if rank == 0:
        # in-window fraction: exact counts (selection-instrumented kernel) on a 1/64 frequency subsample
        sub = np.arange(0, nwn, 64)
        rsel = sess.modm(inp["wn"][sub], 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22, pr["wkl"][:, :, 0],
                         pr["wbrodl"][:, 0], inp["scor"][:, :, :, 0], want_by_mol=False, selection=True,
                         global_range=(inp["v1"], inp["v2"], 0))
        inwin_frac = float(rsel["sel_count"].sum()) / (float(nlines) * NLAY * len(sub))
        fp64_peak = sess.fp64_peak_tflops()
        lk_ms = float(np.mean(lines_ms))
        inwin_per_launch = inwin_frac * float(nlines) * NLAY * nwn
        # algorithmic flops of one launch of the default path: far-field expansions + directly evaluated triples
        far_n, dir_n = float(np.mean(far_exp)), float(np.mean(direct_ev))
        flops_launch = far_n * FLOP_PER_FAR_EXPANSION + dir_n * FLOP_PER_INWINDOW_EVAL
        achieved = flops_launch / (lk_ms * 1e-3) / 1e12
        equiv_direct = inwin_per_launch * FLOP_PER_INWINDOW_EVAL / (lk_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        rt_bytes = (8.0 * NLAY + 48.0 + 8.0) * nwn
        rt_k_ms = float(np.mean(rt_ms))
        line = {
            "metric": "line x layer x frequency evaluations/s", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), shard_frequencies=[int(c) for c in shard_counts],
                           gather="none (one GPU)" if world == 1 else "NCCL all-gather of the six spectra of step i on a second stream while step i+1 "
                                  "computes (two buffers); the gathers still in flight after the last step are inside the timed total",
                           shard_balance="blocks of equal measured cost (three untimed feedback rounds), sum = n_gpus x nwn_per_gpu"
                           if (world > 1 and not args.no_balance) else "equal counts"),
            "spectra_per_s": world / (ms_per_step * 1e-3), "rank_kernel_ms": rank_kernel_ms,
            "inwindow_evals_per_s": value * inwin_frac, "inwindow_fraction": inwin_frac, "logical_lines": nlines,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": float(te.item())},
            "gpu_launches": int(launches),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None, "traffic": ncu_traffic(args),
                         "traffic_source": "profiles/r02_v34_traffic.json: dram bytes read+written by the line-path kernels of one step "
                                           "(ncu --set full, default workload only; null otherwise)",
                         "kernel": "plan+far+near2+voigt+final (the line path of one step)",
                         "peak_source": "measured live: mrtm_fp64_peak DFMA probe (MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_direct_eval": FLOP_PER_INWINDOW_EVAL, "flop_per_far_expansion": FLOP_PER_FAR_EXPANSION,
                         "far_expansions_per_launch": far_n, "direct_evals_per_launch": dir_n,
                         "inwindow_evals_per_launch": inwin_per_launch,
                         "equivalent_direct_tflops": equiv_direct,
                         "note": "achieved counts the flops the expansion algorithm needs (DESIGN.md 3); equivalent_direct_tflops "
                                 "is 12 flop x every in-window triple / kernel time, i.e. what a per-triple kernel would have to "
                                 "sustain for the same time (may exceed the FP64 peak)",
                         "kernel_ms": lk_ms, "share_of_step": lk_ms / ms_per_step},
            "roofline_rt": {"bound": "hbm", "kernel": "rt_kernel", "achieved": rt_bytes / (rt_k_ms * 1e-3) / 1e9 if rt_k_ms else None,
                            "peak": hbm_peak, "unit": "GB/s", "frac": (rt_bytes / (rt_k_ms * 1e-3) / 1e9) / hbm_peak if rt_k_ms else None,
                            "kernel_ms": rt_k_ms, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "derive_kernel_ms": float(np.mean(derive_ms)),
        }
        if direct:
            d_ach = inwin_per_launch * FLOP_PER_INWINDOW_EVAL / (direct["lines_kernel_ms"] * 1e-3) / 1e12
            line["direct_path"] = {"what": "same step with mrtm_opts.line_mode=1: every in-window (line,layer,frequency) triple "
                                           "evaluated per frequency, no far-field expansion", "steps": args.direct_steps,
                                   "value": direct["value"], "unit": "evals/s", "ms_per_step": direct["ms_per_step"],
                                   "lines_kernel_ms": direct["lines_kernel_ms"],
                                   "roofline": {"bound": "fp64", "achieved": d_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                                "frac": d_ach / fp64_peak if fp64_peak else None,
                                                "flop_per_inwindow_eval": FLOP_PER_INWINDOW_EVAL}}
        if not args.no_cpu_baseline:
            orc = {}
            dt, nominal, inwin = cpu_sample(inp, args.cpu_sample_nwn, keep=orc)
            line["parity_check"] = parity_check(sess, inp, hprof, hscor, orc)
            dt2, nominal2, _ = cpu_sample(inp, args.cpu_sample_nwn, opt="O2")
            line["cpu_baseline"] = {"value": nominal / dt, "unit": "evals/s", "cores": 1, "kind": "port", "value_O2": nominal2 / dt2,
                                    "sample": "%d evenly spaced frequencies of the shard x %d layers x all %d lines, oracle -O0 "
                                              "(C restatement; the Fortran reference cannot be compiled here), %.1f s" %
                                              (args.cpu_sample_nwn, NLAY, nlines, dt)}
        print(json.dumps(line))
        if "parity_check" in line and not line["parity_check"]["pass"]:
            parity_failed = True
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_failed:
        raise SystemExit("bench.py: parity_check failed (see the JSON line): the throughput above is not valid")


if __name__ == "__main__":
    main()
