"""Shared test harness: ctypes access to the CPU oracle (oracle/, TEST INFRASTRUCTURE) and helpers
that run one case through the oracle and through the product's C ABI on the GPU."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

from monortm_b200 import api, linefile, synth  # noqa: E402

_orc = {}


def oracle_lib(opt="O0"):
    if opt in _orc:
        return _orc[opt]
    name = "libmonortm_oracle.so" if opt == "O0" else "libmonortm_oracle_O2.so"
    path = os.path.join(ORACLE_DIR, "_build", name)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    lib = C.CDLL(path)
    P, D, I = C.c_void_p, C.c_double, C.c_int64
    lib.orc_lines_alloc.restype = P
    lib.orc_lines_alloc.argtypes = [I]
    lib.orc_lines_free.argtypes = [P]
    lib.orc_get_lnfl.restype = C.c_int
    lib.orc_get_lnfl.argtypes = [C.c_char_p, D, D, P]
    lib.orc_modm.restype = C.c_int
    lib.orc_modm.argtypes = [I, P, D, I, P, P, P, P, P, P, P, P, I, P, P, D, D, D, P, I, P, I, P, P, P, P, P]
    lib.orc_calctmr.restype = C.c_int
    lib.orc_calctmr.argtypes = [I, I, P, P, P, P, P]
    lib.orc_rtm.restype = C.c_int
    lib.orc_rtm.argtypes = [I, I, I, P, I, P, P, P, P, P, P, P, P, P, P, P, I]
    lib.orc_w4.argtypes = [D, D, P, P]
    lib.orc_sd_humlicek.argtypes = [D, D, D, D, P, P]
    lib.orc_sdvoigt.restype = D
    lib.orc_sdvoigt.argtypes = [D, D, D, D, P]
    lib.orc_radfn.restype = D
    lib.orc_radfn.argtypes = [D, D]
    lib.orc_odclw.restype = D
    lib.orc_odclw.argtypes = [D, D, D]
    lib.orc_bb_fn.restype = D
    lib.orc_bb_fn.argtypes = [D, D]
    lib.orc_contnm_one.restype = C.c_int
    lib.orc_contnm_one.argtypes = [I, P, D, D, P, D, I, D, D, D, D, I, P]
    lib.orc_line_key.restype = C.c_uint64
    lib.orc_line_key.argtypes = [I, I]
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_branch_counts.argtypes = [P]
    lib.orc_tips_2003.restype = C.c_int
    lib.orc_tips_2003.argtypes = [I, D, P]
    _orc[opt] = lib
    return lib


class OrcLinesStruct(C.Structure):
    _fields_ = [("iim", C.c_int64), ("nblm", C.c_int64 * 39), ("iso", C.c_void_p)] + \
               [(n, C.c_void_p) for n in ("xnu0", "deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep",
                                          "brd_mol_flg", "brd_mol_tmp", "brd_mol_hw", "brd_mol_shft")]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def oracle_lines_from_store(ls):
    """Wrap a LineStore's numpy arrays in the oracle's orc_lines struct (no copy; the oracle applies
    the H2O self-width fix-up in place, modm.f90:841, so pass a copy when the store is reused)."""
    s = OrcLinesStruct()
    s.iim = ls.iim
    for i in range(39):
        s.nblm[i] = int(ls.nblm[i])
    s.iso = _p(ls.iso)
    for n in ("xnu0", "deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep",
              "brd_mol_flg", "brd_mol_tmp", "brd_mol_hw", "brd_mol_shft"):
        setattr(s, n, _p(getattr(ls, n)))
    return s


def oracle_read_tape3(path, v1, v2, iim):
    """orc_get_lnfl into a fresh LineStore (arrays owned by numpy)."""
    lib = oracle_lib()
    ls = linefile.LineStore(iim)
    s = oracle_lines_from_store(ls)
    rc = lib.orc_get_lnfl(str(path).encode(), float(v1), float(v2), C.byref(s))
    if rc:
        raise RuntimeError("orc_get_lnfl: %s" % lib.orc_last_error().decode())
    for i in range(39):
        ls.nblm[i] = s.nblm[i]
    return ls


def copy_store(ls):
    out = linefile.LineStore(ls.iim)
    out.nblm[:] = ls.nblm
    for n in ("iso",) + tuple(x for x in linefile.LineStore.ARRAY_ORDER if x != "iso"):
        getattr(out, n)[...] = getattr(ls, n)
    return out


def oracle_modm(ls, wn, dvset, p, t, clw, nmol, wkl, wbrodl, scor, cntnm=(1.,) * 7, sclcpl=1., sclhw=1.,
                y0res=0., ibrd=0, odxsec_in=None, opt="O0", selection=True):
    lib = oracle_lib(opt)
    F = dict(order="F")
    wn = np.asfortranarray(wn, dtype=np.float64)
    nwn, nlay = len(wn), len(p)
    o = np.zeros((nwn, nlay), **F)
    obm = np.zeros((nwn, 39, nlay), **F)
    oc = np.zeros((nwn, 39, nlay), **F)
    oclw = np.zeros((nwn, nlay), **F)
    odx = np.zeros((nwn, nlay), **F)
    selc = np.zeros((nwn, nlay), np.int64, **F) if selection else None
    selh = np.zeros((nwn, nlay), np.uint64, **F) if selection else None
    nv = C.c_int64(0)
    arrs = [np.asfortranarray(a, dtype=np.float64) for a in (p, t, clw, wkl, wbrodl, scor)]
    c7 = np.array(cntnm, dtype=np.float64)
    lsw = copy_store(ls)
    s = oracle_lines_from_store(lsw)
    odin = None if odxsec_in is None else np.asfortranarray(odxsec_in, dtype=np.float64)
    rc = lib.orc_modm(nwn, _p(wn), float(dvset), nlay, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
                      _p(o), _p(obm), _p(oc), _p(oclw), _p(odx), int(nmol), _p(arrs[3]), _p(arrs[4]),
                      float(sclcpl), float(sclhw), float(y0res), _p(c7), 1 if odin is not None else 0, _p(odin),
                      int(ibrd), _p(arrs[5]), C.byref(s), _p(selc), _p(selh), C.byref(nv))
    if rc:
        raise RuntimeError("orc_modm rc=%d: %s" % (rc, lib.orc_last_error().decode()))
    br = np.zeros(8, np.int64)
    lib.orc_branch_counts(_p(br))
    return dict(o=o, o_by_mol=obm, oc=oc, o_clw=oclw, odxsec=odx, sel_count=selc, sel_hash=selh, n_voigt=nv.value,
                branches=dict(zip(("voigt", "sdep", "co2", "co2_lc1", "generic_lc", "o2_lc", "other_flag", "neg_res"), br.tolist())))


def oracle_tips_2003(mol_max, temp, opt="O0"):
    lib = oracle_lib(opt)
    scor = np.zeros((42, 9), order="F")
    rc = lib.orc_tips_2003(int(mol_max), float(temp), _p(scor))
    if rc:
        raise RuntimeError("orc_tips_2003 rc=%d: %s" % (rc, lib.orc_last_error().decode()))
    return scor


def oracle_scor_for_layers(nmol, t):
    """scor(42,9,nlay[,nprof]) from the oracle's own TIPS_2003 (no product code involved)"""
    t = np.asarray(t, dtype=np.float64)
    out = np.zeros((42, 9) + t.shape, order="F")
    cache = {}
    for idx in np.ndindex(t.shape):
        key = float(t[idx])
        if key not in cache:
            cache[key] = oracle_tips_2003(nmol, key)
        out[(slice(None), slice(None)) + idx] = cache[key]
    return out


def oracle_calctmr(wn, t, tz, o):
    lib = oracle_lib()
    wn = np.asfortranarray(wn, dtype=np.float64)
    t, tz, o = (np.asfortranarray(a, dtype=np.float64) for a in (t, tz, o))
    tmr = np.zeros(len(wn))
    rc = lib.orc_calctmr(len(t), len(wn), _p(wn), _p(t), _p(tz), _p(o), _p(tmr))
    assert rc == 0
    return tmr


def oracle_rtm(iout, irt, wn, t, tz, o, tmpsfc, reflc, emiss, idu=1):
    lib = oracle_lib()
    wn = np.asfortranarray(wn, dtype=np.float64)
    t, tz, o, reflc, emiss = (np.asfortranarray(a, dtype=np.float64) for a in (t, tz, o, reflc, emiss))
    n = len(wn)
    ts = C.c_double(float(tmpsfc))
    rup, trtot, rdn, rad, tb = (np.zeros(n) for _ in range(5))
    rc = lib.orc_rtm(int(iout), int(irt), n, _p(wn), len(t), _p(t), _p(tz), _p(o), C.cast(C.byref(ts), C.c_void_p),
                     _p(rup), _p(trtot), _p(rdn), _p(reflc), _p(emiss), _p(rad), _p(tb), int(idu))
    if rc:
        raise RuntimeError("orc_rtm rc=%d: %s" % (rc, lib.orc_last_error().decode()))
    return dict(rad=rad, tb=tb, rup=rup, rdn=rdn, trtot=trtot, tmpsfc=ts.value)


# --------------------------------------------------------------------------------------- cross sections
class OrcXsRegion(C.Structure):
    _fields_ = [("ixmol", C.c_int32), ("ntemp", C.c_int32), ("npts", C.c_int64), ("v1fx", C.c_double), ("v2fx", C.c_double),
                ("v1x", C.c_double), ("v2x", C.c_double), ("xdoplr", C.c_double), ("tx", C.c_double * 6), ("pdx", C.c_double * 6),
                ("xsdat", C.c_void_p * 6)]


def xs_region_array(regs, cls=OrcXsRegion):
    """regions as monortm_b200.xsfile.read_regions returns them -> C array (+ the numpy arrays that must stay alive)"""
    arr = (cls * max(len(regs), 1))()
    keep = []
    for i, r in enumerate(regs):
        a, last = arr[i], r["files"][-1]
        a.ixmol, a.ntemp, a.npts = r["ixmol"], len(r["files"]), last["npts"]
        a.v1fx, a.v2fx, a.v1x, a.v2x, a.xdoplr = r["v1fx"], r["v2fx"], last["v1x"], last["v2x"], r["xdoplr"]
        for j, f in enumerate(r["files"]):
            a.tx[j], a.pdx[j] = f["t"], f["pres"]
            d = np.ascontiguousarray(f["data"], dtype=np.float64)
            keep.append(d)
            a.xsdat[j] = d.ctypes.data
    return arr, keep


def oracle_xsec(regs, wn, p, t, xamnt):
    lib = oracle_lib()
    lib.orc_xsec_sub.restype = C.c_int
    lib.orc_xsec_sub.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p]
    wn, p, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (wn, p, t))
    xa = np.asfortranarray(xamnt, dtype=np.float64)
    od = np.zeros((len(wn), len(p)), order="F")
    arr, keep = xs_region_array(regs)
    rc = lib.orc_xsec_sub(len(wn), _p(wn), len(p), _p(p), _p(t), len(regs), C.addressof(arr), xa.shape[0], _p(xa), _p(od))
    if rc:
        raise RuntimeError("orc_xsec_sub rc=%d: %s" % (rc, lib.orc_last_error().decode()))
    return od


def oracle_convolve(xspd, v1x, v2x, delvx, pd, hwdop, tave, pave, wn):
    lib = oracle_lib()
    lib.orc_convolve.restype = C.c_int
    lib.orc_convolve.argtypes = [C.c_void_p, C.c_int64] + [C.c_double] * 7 + [C.c_void_p, C.c_int64, C.c_void_p]
    xspd, wn = np.ascontiguousarray(xspd, dtype=np.float64), np.ascontiguousarray(wn, dtype=np.float64)
    out = np.zeros(len(wn))
    rc = lib.orc_convolve(_p(xspd), len(xspd), v1x, v2x, delvx, pd, hwdop, tave, pave, _p(wn), len(wn), _p(out))
    if rc:
        raise RuntimeError("orc_convolve rc=%d: %s" % (rc, lib.orc_last_error().decode()))
    return out


def xsec_case(spec, directory):
    """the seeded synthetic cross-section case of tests/golden/ref_xsec_synth.npz: (regions, wn, p, t, xamnt)"""
    from monortm_b200 import xsfile
    wn, p, t = np.array(spec["wn"]), np.array(spec["p"]), np.array(spec["t"])
    xamnt = np.zeros((xsfile.MX_XS, len(p)), order="F")
    xamnt[:len(spec["names"])] = np.array(spec["xamnt"])
    xsfile.synthetic_set(directory, seed=spec["seed"])
    regs = xsfile.read_regions(directory, spec["names"], float(wn.min()), float(wn.max()))
    return regs, wn, p, t, xamnt


# --------------------------------------------------------------------------------------- cases
_tape_cache = {}


def synthetic_store(n_filler=512, seed=20260101, v1=0.0, v2=55.0, **kw):
    """TAPE3-synth -> file -> GET_LNFL (C++ host reader) -> LineStore, cached per parameter set."""
    key = (n_filler, seed, v1, v2, tuple(sorted(kw.items())))
    if key in _tape_cache:
        return _tape_cache[key]
    recs = synth.synthetic_records(n_filler, seed=seed, **kw)
    with tempfile.NamedTemporaryFile(suffix=".tape3", delete=False) as f:
        path = f.name
    try:
        linefile.write_tape3(path, recs)
        ls = linefile.read_tape3(path, v1, v2)
    finally:
        os.unlink(path)
    _tape_cache[key] = ls
    return ls


def make_case(n_filler=512, nlay=20, wn=None, dvset=0.0, irt=3, clw=False, nprof=1, seed0=1000, nmol=22,
              cntnm=(1.,) * 7, ibrd=0, tmpsfc=290.0, emis=0.6, line_kw=None, **kw):
    wn = np.asarray(wn, dtype=np.float64)
    ls = synthetic_store(n_filler, v1=float(wn[0]), v2=float(wn[-1]), **(line_kw or {}))
    prof = synth.synthetic_profiles(nprof, nlay, seed0=seed0, clw_layers=clw, nmol=nmol)
    scor = api.scor_for_layers(nmol, prof["t"])
    n = len(wn)
    return dict(ls=ls, wn=wn, dvset=dvset, prof=prof, scor=scor, irt=irt, cntnm=cntnm, ibrd=ibrd, nmol=nmol,
                tmpsfc=tmpsfc, emiss=np.full(n, emis), reflc=np.full(n, 1.0 - emis), **kw)


def run_oracle(case, ip=0, opt="O0"):
    pr = case["prof"]
    m = oracle_modm(case["ls"], case["wn"], case["dvset"], pr["p"][:, ip], pr["t"][:, ip], pr["clw"][:, ip],
                    case["nmol"], pr["wkl"][:, :, ip], pr["wbrodl"][:, ip], case["scor"][:, :, :, ip],
                    cntnm=case["cntnm"], ibrd=case["ibrd"], opt=opt,
                    sclcpl=case.get("sclcpl", 1.), sclhw=case.get("sclhw", 1.), y0res=case.get("y0res", 0.))
    tmr = oracle_calctmr(case["wn"], pr["t"][:, ip], pr["tz"][:, ip], m["o"])
    r = oracle_rtm(1, case["irt"], case["wn"], pr["t"][:, ip], pr["tz"][:, ip], m["o"], case["tmpsfc"],
                   case["reflc"], case["emiss"])
    m.update(r)
    m["tmr"] = tmr
    return m


_session = None


def session():
    global _session
    if _session is None:
        _session = api.Session(0)
    return _session


def run_gpu(case, ip=0, by_mol=True, line_mode=0, selection=True):
    """MODM, CALCTMR and RTM through the C ABI (host buffers), one profile.
    line_mode 0 = default (far-field expansion where applicable), 1 = direct evaluation of every triple."""
    s = session()
    s.stage_lines(case["ls"])
    pr = case["prof"]
    s.reset_stats()
    m = s.modm(case["wn"], case["dvset"], pr["p"][:, ip], pr["t"][:, ip], pr["clw"][:, ip], case["nmol"],
               pr["wkl"][:, :, ip], pr["wbrodl"][:, ip], case["scor"][:, :, :, ip], cntnm=case["cntnm"],
               ibrd=case["ibrd"], want_by_mol=by_mol, selection=selection, line_mode=line_mode,
               sclcpl=case.get("sclcpl", 1.), sclhw=case.get("sclhw", 1.), y0res=case.get("y0res", 0.))
    tmr = s.calctmr(case["wn"], pr["t"][:, ip], pr["tz"][:, ip], m["o"])
    r = s.rtm(1, case["irt"], case["wn"], pr["t"][:, ip], pr["tz"][:, ip], m["o"], case["tmpsfc"],
              case["reflc"], case["emiss"])
    m.update(r)
    m["tmr"] = tmr
    m["stats"] = s.stats()
    return m


def rel_diff(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    den = np.maximum(np.abs(b), floor)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, np.where(a == b, 0.0, np.inf))
    return float(np.max(r)) if r.size else 0.0
