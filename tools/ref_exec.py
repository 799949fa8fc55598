"""ref_exec.py -- run the reference's own Fortran text on the CPU through the mechanical translator tools/f90fn.py.
TEST INFRASTRUCTURE: used in THIS container only (it reads /root/reference/src) by tools/gen_ref_goldens.py to produce the
committed vectors tests/golden/ref_*.npz.  Nothing under tests/, bench.py or the product reads /root/reference at run time.

Translated units (no hand-written numerics): tips_2003.f90 (TIPS_2003 and every QT_* table routine), modm.f90 (MODM, LINES,
INITI, INTENS, HALFWHM_C/D, LSF_LORTZ, LSF_SDVOIGT, SDVOIGT, W4, SD_Humlicek, XLORENTZ, chi_fn), contnm.f90 (CONTNM and all
its table accessors / BLOCK DATA), CntnmFactors.f90, lblrtm_sub.f90 (XINT, RADFN), CloudOptProp.f90, RTMmono.f90 (RTM,
RAD_UP_DN, bb_fn, calctmr), PhysConstants.f90, PlanetEarth.f90, lblparams.f90, isotope.incl.
Not translated: GET_LNFL (binary TAPE3 I/O) -- the module arrays it fills are loaded from a LineStore that the C++ and the
oracle readers produce identically (tests/test_linefile.py) -- and MONORTM_XSEC_SUB (ixsect=0 in every golden).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90fn  # noqa: E402
import f90rt as rt  # noqa: E402

REF = "/root/reference/src/"
FILES = [os.path.join(HERE, "f90_stubs", "lnfl_mod_decl.f90")] + [REF + f for f in (
    "PhysConstants.f90", "PlanetEarth.f90", "lblparams.f90", "lblrtm_sub.f90", "RTMmono.f90", "CloudOptProp.f90",
    "CntnmFactors.f90", "tips_2003.f90", "contnm.f90", "modm.f90")]
F = rt.FArr
_ns = None


def load(dump=None):
    """Translate and load the reference once per process."""
    global _ns
    if _ns is None:
        ext = {"get_lnfl": lambda *a: (None,), "monortm_xsec_sub": lambda *a: (None,)}
        _ns = f90fn.load(FILES, externals=ext, dump=dump)
    return _ns


def set_lines(ls):
    """Fill LNFL_MOD's public arrays (lnfl_mod.f90:9-13) from a LineStore (what GET_LNFL leaves behind)."""
    ns = load()
    M = ns["M_lnfl_mod"]
    iim = M.xnu0.a.shape[1]
    n = int(max(ls.nblm))
    if n > iim:
        raise ValueError("line store needs IIM >= %d (stub module has %d)" % (n, iim))
    for name in ("deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep", "xnu0"):
        a = getattr(M, name).a
        a[...] = 0.0
        a[:, :n] = getattr(ls, name)[:, :n]
    M.iso.a[...] = 0
    M.iso.a[:, :n] = ls.iso[:, :n]
    M.nblm.a[...] = ls.nblm
    for name in ("brd_mol_flg", "brd_mol_tmp", "brd_mol_hw", "brd_mol_shft"):
        a = getattr(M, name).a
        a[...] = 0
        a[:, :, :n] = getattr(ls, name)[:, :, :n]
    ns["S_modm"].init = False          # GET_LNFL "has been called" (modm.f90:187-190)


def _padl(a, n):
    out = np.zeros(n)
    out[:len(a)] = a
    return F(out)


def modm(ls, wn, dvset, p, t, clw, nmol, wkl, wbrodl, cntnm=(1.,) * 7, sclcpl=1., sclhw=1., y0res=0., ibrd=0):
    """CALL MODM(...) (modm.f90:21-25) with TIPS_2003 called inside, exactly as the reference does (modm.f90:250)."""
    ns = load()
    set_lines(ls)
    mxlay = ns["M_lblparams"].mxlay
    wn = np.asarray(wn, dtype=np.float64)
    nwn, nlay = len(wn), len(p)
    o, oclw, odx = F.zeros((nwn, nlay)), F.zeros((nwn, nlay)), F.zeros((nwn, nlay))
    obm, oc = F.zeros((nwn, 39, nlay)), F.zeros((nwn, 39, nlay))
    wklp = np.zeros((39, mxlay), order="F")
    wklp[:, :nlay] = wkl
    cf = ns["T_cntnmfactors_t"](*[float(x) for x in cntnm])
    rt.oob_reads[0] = 0
    ns["modm"](0, 1, nwn, F(wn.copy()), float(dvset), nlay, _padl(p, mxlay), _padl(t, mxlay), _padl(clw, mxlay), o, obm, oc,
               oclw, odx, int(nmol), F(wklp), _padl(wbrodl, mxlay), float(sclcpl), float(sclhw), float(y0res), "TAPE3", cf,
               0, int(ibrd))
    return dict(o=o.a, o_by_mol=obm.a, oc=oc.a, o_clw=oclw.a, odxsec=odx.a, oob_reads=rt.oob_reads[0])


def calctmr(wn, t, tz, o):
    ns = load()
    mxlay = ns["M_lblparams"].mxlay
    nwn, nlay = len(wn), len(t)
    tmr = F.zeros(nwn)
    tzp = np.zeros(mxlay + 1)
    tzp[:nlay + 1] = tz
    ns["calctmr"](nlay, nwn, F(np.array(wn, dtype=np.float64)), _padl(t, mxlay), F(tzp, (0,)), F(np.asfortranarray(o).copy(order="F")), tmr)
    return tmr.a


def rtm(iout, irt, wn, t, tz, o, tmpsfc, reflc, emiss, idu=1):
    ns = load()
    mxlay = ns["M_lblparams"].mxlay
    nwn, nlay = len(wn), len(t)
    outs = {k: F.zeros(nwn) for k in ("rup", "trtot", "rdn", "rad", "tb")}
    tzp = np.zeros(mxlay + 1)
    tzp[:nlay + 1] = tz
    r = ns["rtm"](int(iout), int(irt), nwn, F(np.array(wn, dtype=np.float64)), nlay, _padl(t, mxlay), F(tzp, (0,)),
                  F(np.asfortranarray(o).copy(order="F")), float(tmpsfc), outs["rup"], outs["trtot"], outs["rdn"],
                  F(np.array(reflc, dtype=np.float64)), F(np.array(emiss, dtype=np.float64)), outs["rad"], outs["tb"], int(idu))
    res = {k: v.a for k, v in outs.items()}
    res["tmpsfc"] = r[1]            # TMPSFC is modified for IRT 2, 3 (RTMmono.f90:122)
    return res


def tips_2003(mol_max, temp):
    ns = load()
    scor = F.zeros((42, 9))
    ns["tips_2003"](int(mol_max), float(temp), scor)
    return scor.a


def fn(name):
    """a translated scalar routine; returns the Python callable (result tuple: [0] = function value, then the modified
    scalar dummies in argument order)"""
    return load()[name]
