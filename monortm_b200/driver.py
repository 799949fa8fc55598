"""Python mirror of the host driver entry points of libmonortm_b200.so: the pieces of
PROGRAM MONORTM around the hot path for layer input (IATM=0) -- RDLBLINP (src/monortm_sub.F90:33-423),
the MONORTM_PROF.IN reader (src/monortm.f90:380-488), EMISS_REFLEC (:506-516), STOREOUT (:519-787) and
the per-profile loop (src/monortm.f90:357-588).  Everything is done by the C++ library; this module
only marshals arguments.  run_monortm needs a GPU (no CPU path)."""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import MrtmControl
from .api import MonortmError, MXMOL, _f, _ptr

MXLAY = 603   # src/lblparams.f90:28


def _lib():
    return _capi.load_library()


def _check(rc):
    if rc:
        lib = _lib()
        raise MonortmError(rc, (lib.mrtm_host_last_error() or b"").decode() or lib.mrtm_strerror(rc).decode())


def read_control(filein, nwnmx=0):
    """RDLBLINP records 1.1-1.4 -> dict (wn as numpy array)."""
    lib = _lib()
    c = MrtmControl()
    _check(lib.mrtm_host_read_control(os.fsencode(filein), int(nwnmx), C.byref(c)))
    try:
        out = {k: getattr(c, k) for k in ("ihirac", "icntnm", "iemit", "iplot", "iatm", "iod", "ixsect", "ispd", "ibrd",
                                          "v1", "v2", "dvset", "tmpbnd", "nmol_scal")}
        out["cntnm"] = tuple(c.cntnm)
        out["bndemi"], out["bndrfl"] = tuple(c.bndemi), tuple(c.bndrfl)
        out["wn"] = np.array([c.wn[i] for i in range(c.nwn)], dtype=np.float64)
        out["hmol_scal"] = c.hmol_scal[:c.nmol_scal].decode("latin1") if c.nmol_scal else ""
        out["xmol_scal"] = tuple(c.xmol_scal[i] for i in range(c.nmol_scal))
    finally:
        lib.mrtm_host_free_control(C.byref(c))
    return out


def emiss_reflec(ctrl, wn, dir=""):
    """EMISS_REFLEC for the control dict of read_control -> (emiss, reflc)."""
    lib = _lib()
    c = MrtmControl()
    for i in range(3):
        c.bndemi[i], c.bndrfl[i] = ctrl["bndemi"][i], ctrl["bndrfl"][i]
    wn = _f(wn)
    emiss, reflc = np.zeros(len(wn)), np.zeros(len(wn))
    d = os.fsencode(dir) if dir else None
    _check(lib.mrtm_host_emiss_reflec(C.byref(c), d, len(wn), _ptr(wn), _ptr(emiss), _ptr(reflc)))
    return emiss, reflc


def count_profiles(fileprof, ixsect=0):
    n = C.c_int64(0)
    _check(_lib().mrtm_host_count_profiles(os.fsencode(fileprof), int(ixsect), C.byref(n)))
    return n.value


def read_profile(fileprof, index=0):
    """One profile of MONORTM_PROF.IN as the dict Session.profiles() takes (trailing profile dim 1)."""
    lib = _lib()
    iv = [C.c_int64(0) for _ in range(4)]
    dv = [C.c_double(0) for _ in range(4)]
    p, t, clw, wbrodl = (np.zeros(MXLAY) for _ in range(4))
    altz, pz, tz = (np.zeros(MXLAY + 1) for _ in range(3))
    wkl = np.zeros((MXMOL, MXLAY), order="F")
    _check(lib.mrtm_host_read_profile(os.fsencode(fileprof), int(index), MXLAY, *[C.byref(x) for x in iv],
                                      *[C.byref(x) for x in dv], _ptr(p), _ptr(t), _ptr(clw), _ptr(wbrodl),
                                      _ptr(altz), _ptr(pz), _ptr(tz), _ptr(wkl)))
    iform, nlay, nmol, irt = (x.value for x in iv)
    secnt0, h1, h2, angle = (x.value for x in dv)
    F = dict(order="F")
    return dict(iform=iform, nlay=nlay, nprof=1, nmol=nmol, irt=irt, secnt0=secnt0, h1=h1, h2=h2, angle=angle,
                p=np.asarray(p[:nlay].reshape(nlay, 1), **F), t=np.asarray(t[:nlay].reshape(nlay, 1), **F),
                clw=np.asarray(clw[:nlay].reshape(nlay, 1), **F), wbrodl=np.asarray(wbrodl[:nlay].reshape(nlay, 1), **F),
                tz=np.asarray(tz[:nlay + 1].reshape(nlay + 1, 1), **F), pz=pz[:nlay + 1].copy(), altz=altz[:nlay + 1].copy(),
                wkl=np.asarray(wkl[:, :nlay].reshape(MXMOL, nlay, 1), **F))


def storeout(fileout, append, wn, wkl, wbrodl, rad, tb, trtot, npr, o, o_by_mol, oc, odxsec, tmr, wvcolmn, clwcolmn,
             tmpsfc, reflc, emiss, nmol, angle, iod=0):
    """STOREOUT with the reference's argument list; wkl (39,nlay) is modified in place like the reference's."""
    wn = _f(wn)
    nwn = len(wn)
    nlay = np.asarray(wbrodl).shape[0]
    if not (isinstance(wkl, np.ndarray) and wkl.flags.f_contiguous and wkl.dtype == np.float64 and wkl.shape == (MXMOL, nlay)):
        raise ValueError("wkl must be a Fortran-ordered float64 (39,nlay) array (it is IN/OUT)")
    a = [_f(x, (nwn,)) for x in (rad, tb, trtot, tmr, reflc, emiss)]
    o = _f(o, (nwn, nlay))
    obm, occ = _f(o_by_mol, (nwn, MXMOL, nlay)), _f(oc, (nwn, MXMOL, nlay))
    odx = None if odxsec is None else _f(odxsec, (nwn, nlay))
    wb = _f(wbrodl, (nlay,))
    _check(_lib().mrtm_host_storeout(os.fsencode(fileout), int(bool(append)), nwn, _ptr(wn), _ptr(wkl), _ptr(wb),
                                     _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), int(npr), _ptr(o), _ptr(obm), _ptr(occ),
                                     _ptr(odx), _ptr(a[3]), float(wvcolmn), float(clwcolmn), float(tmpsfc),
                                     _ptr(a[4]), _ptr(a[5]), nlay, int(nmol), float(angle), int(iod)))


def run_monortm(workdir, device=0, nwnmx=0, verbose=False):
    """PROGRAM MONORTM for IATM=0 in `workdir` (MONORTM.IN, MONORTM_PROF.IN, TAPE3 -> MONORTM.OUT)."""
    _check(_lib().mrtm_host_run_monortm(os.fsencode(workdir), int(device), int(nwnmx), int(bool(verbose))))
