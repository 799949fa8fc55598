"""Host-side mirror of the reference's operator interface for the hot path, on top of the C ABI.

The three operators keep the reference's names, argument meaning and error behaviour:
  Session.modm(...)     <-> SUBROUTINE MODM     (src/modm.f90:21-25)
  Session.calctmr(...)  <-> SUBROUTINE CALCTMR  (src/RTMmono.f90:239)
  Session.rtm(...)      <-> SUBROUTINE RTM      (src/RTMmono.f90:13-14)
plus Session.profiles(...), the fused batched path.  Arrays are numpy float64 in Fortran order
exactly as the Fortran driver dimensions them.  A reference STOP becomes MonortmError.
All computation happens in libmonortm_b200.so on the GPU; nothing here computes spectra.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MrtmOpts, MrtmStats

MXMOL = 39
CNTNM_ALL_ONE = (1.0,) * 7   # ICNTNM=1 (src/CntnmFactors.f90:160-161)


class MonortmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("monortm_b200 error %d: %s" % (code, msg))
        self.code = code


def _f(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tips_2003(mol_max, temp):
    """TIPS_2003 (src/tips_2003.f90:2-298) through the C++ host helper: scor(42,9)."""
    lib = _capi.load_library()
    scor = np.zeros((42, 9), order="F")
    rc = lib.mrtm_host_tips_2003(int(mol_max), float(temp), _ptr(scor))
    if rc:
        raise MonortmError(rc, lib.mrtm_strerror(rc).decode())
    return scor


def host_xsread(directory, names, xv1, xv2):
    """XSREAD + the table READs through the C++ host helper; returns regions in the layout of xsfile.read_regions."""
    lib = _capi.load_library()
    n = C.c_int64()
    regs = C.POINTER(_capi.MrtmXsRegion)()
    buf = b"".join(nm.encode().ljust(10)[:10] for nm in names)
    rc = lib.mrtm_host_xsread(str(directory).encode(), len(names), buf, float(xv1), float(xv2), C.byref(n), C.byref(regs))
    if rc:
        raise MonortmError(rc, lib.mrtm_host_last_error().decode())
    out = []
    try:
        for i in range(n.value):
            g = regs[i]
            files = [dict(v1x=g.v1x, v2x=g.v2x, npts=int(g.npts), t=g.tx[k], pres=g.pdx[k],
                          data=np.ctypeslib.as_array(C.cast(g.xsdat[k], _capi.c_double_p), shape=(int(g.npts),)).copy())
                     for k in range(g.ntemp)]
            out.append(dict(ixmol=int(g.ixmol), v1fx=g.v1fx, v2fx=g.v2fx, xdoplr=g.xdoplr, files=files))
    finally:
        lib.mrtm_host_xs_free(regs, n.value)
    return out


def scor_for_layers(nmol, t):
    """scor(42,9,nlay[,nprof]) for layer temperatures t (what the Fortran host passes per layer)."""
    t = np.asarray(t, dtype=np.float64)
    out = np.zeros((42, 9) + t.shape, order="F")
    cache = {}
    for idx in np.ndindex(t.shape):
        key = float(t[idx])
        if key not in cache:
            cache[key] = tips_2003(int(nmol), key)
        out[(slice(None), slice(None)) + idx] = cache[key]
    return out


class Session:
    """One mrtm_ctx (one GPU).  Not re-entrant, like the reference's process-global state."""

    def __init__(self, device=0, device_mask=None):
        """device: one GPU.  device_mask (int, bit d = CUDA device d, 0 = all): one context over several GPUs of the node
        (mrtm_init_multi): profiles() splits by profile or by frequency inside the library."""
        self.lib = _capi.load_library()
        h = C.c_void_p()
        if device_mask is not None:
            rc = self.lib.mrtm_init_multi(int(device_mask), C.byref(h))
        else:
            rc = self.lib.mrtm_init(int(device), C.byref(h))
        if rc:
            msg = self.lib.mrtm_last_error(h if h else None)
            if h:
                self.lib.mrtm_free(h)
            raise MonortmError(rc, (msg or b"").decode() or self.lib.mrtm_strerror(rc).decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mrtm_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            detail = self.lib.mrtm_last_error(self.h)
            raise MonortmError(rc, (detail or b"").decode() or self.lib.mrtm_strerror(rc).decode())

    # ---- line store ---------------------------------------------------------------------------
    def stage_lines(self, ls):
        """ls: monortm_b200.linefile.LineStore (the arrays GET_LNFL fills)."""
        self._check(self.lib.mrtm_stage_lines(self.h, _ptr(ls.nblm), ls.iim, *ls.pointers()))
        return int(self.lib.mrtm_num_lines(self.h))

    # ---- cross sections -----------------------------------------------------------
    def stage_xsec(self, regs):
        """regs: regions as monortm_b200.xsfile.read_regions / host_xsread return them (list of dicts), ordered by molecule."""
        arr = (_capi.MrtmXsRegion * max(len(regs), 1))()
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
        self._check(self.lib.mrtm_stage_xsec(self.h, len(regs), C.addressof(arr)))

    def xsec(self, wn, p, t, xamnt):
        """MONORTM_XSEC_SUB (src/monortm_sub.F90:1540): xamnt (ld, nlay) column amounts of the staged molecules -> odxsec (nwn, nlay)."""
        wn = _f(wn)
        nwn, nlay = wn.shape[0], np.asarray(p).shape[0]
        p, t = _f(p, (nlay,)), _f(t, (nlay,))
        xamnt = np.asfortranarray(xamnt, dtype=np.float64)
        od = np.zeros((nwn, nlay), order="F")
        self._check(self.lib.mrtm_xsec(self.h, nwn, _ptr(wn), nlay, _ptr(p), _ptr(t), xamnt.shape[0], _ptr(xamnt), _ptr(od)))
        return od

    def num_devices(self):
        return int(self.lib.mrtm_num_devices(self.h))

    def sync(self):
        """wait for the last asynchronous profiles_dev call; raises its deferred error"""
        self._check(self.lib.mrtm_sync(self.h))

    def stats(self):
        st = MrtmStats()
        self._check(self.lib.mrtm_get_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in MrtmStats._fields_}

    def reset_stats(self):
        self._check(self.lib.mrtm_reset_stats(self.h))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.mrtm_fp64_peak(self.h, C.byref(v)))
        return v.value

    @staticmethod
    def _opts(v1=None, v2=None, iw0=0, sel=None, line_mode=0, xamnt=None):
        o = MrtmOpts()
        o.line_mode = int(line_mode)
        if xamnt is not None:
            o.xamnt = xamnt.ctypes.data_as(_capi.c_double_p)
            o.ld_xamnt = int(xamnt.shape[0])
        if v1 is not None:
            o.use_global_range = 1
            o.v1_global, o.v2_global, o.iw0 = float(v1), float(v2), int(iw0)
        if sel is not None:
            o.sel_count = sel[0].ctypes.data_as(_capi.c_int64_p)
            o.sel_hash = sel[1].ctypes.data_as(_capi.c_uint64_p)
        return o

    # ---- MODM ---------------------------------------------------------------------------------
    def modm(self, wn, dvset, p, t, clw, nmol, wkl, wbrodl, scor, cntnm=CNTNM_ALL_ONE,
             sclcpl=1.0, sclhw=1.0, y0res=0.0, ixsect=0, odxsec=None, ibrd=0,
             want_by_mol=True, selection=False, global_range=None, line_mode=0, xamnt=None):
        """Returns dict(o, o_by_mol, oc, o_clw, odxsec[, sel_count, sel_hash]); shapes as in
        src/monortm.f90:352-353 with mxlay -> nlay.  ixsect=1 with xamnt (ld, nlay): MONORTM_XSEC_SUB runs inside, as in MODM."""
        wn = _f(wn)
        nwn, nlay = wn.shape[0], np.asarray(p).shape[0]
        p, t, clw, wbrodl = _f(p, (nlay,)), _f(t, (nlay,)), _f(clw, (nlay,)), _f(wbrodl, (nlay,))
        wkl = _f(wkl, (MXMOL, nlay))
        scor = None if scor is None else _f(scor, (42, 9, nlay))
        o = np.zeros((nwn, nlay), order="F")
        o_clw = np.zeros((nwn, nlay), order="F")
        odx = np.zeros((nwn, nlay), order="F") if odxsec is None else _f(odxsec, (nwn, nlay)).copy(order="F")
        obm = np.zeros((nwn, MXMOL, nlay), order="F") if want_by_mol else None
        oc = np.zeros((nwn, MXMOL, nlay), order="F") if want_by_mol else None
        sel = None
        if selection:
            sel = (np.zeros((nwn, nlay), np.int64, order="F"), np.zeros((nwn, nlay), np.uint64, order="F"))
        gr = global_range or (None, None, 0)
        xa = None if xamnt is None else np.asfortranarray(xamnt, dtype=np.float64)
        opts = self._opts(gr[0], gr[1], gr[2], sel, line_mode, xa)
        c7 = _f(np.array(cntnm, dtype=np.float64), (7,))
        self._check(self.lib.mrtm_modm(self.h, nwn, _ptr(wn), float(dvset), nlay, _ptr(p), _ptr(t), _ptr(clw),
                                       _ptr(o), _ptr(obm), _ptr(oc), _ptr(o_clw), _ptr(odx), int(nmol),
                                       _ptr(wkl), _ptr(wbrodl), float(sclcpl), float(sclhw), float(y0res),
                                       _ptr(c7), int(ixsect), int(ibrd), _ptr(scor), C.byref(opts)))
        out = dict(o=o, o_by_mol=obm, oc=oc, o_clw=o_clw, odxsec=odx)
        if sel:
            out["sel_count"], out["sel_hash"] = sel
        return out

    # ---- CALCTMR / RTM --------------------------------------------------------------------------
    def calctmr(self, wn, t, tz, o):
        wn = _f(wn)
        nwn, nlay = wn.shape[0], np.asarray(t).shape[0]
        t, tz, o = _f(t, (nlay,)), _f(tz, (nlay + 1,)), _f(o, (nwn, nlay))
        tmr = np.zeros(nwn)
        self._check(self.lib.mrtm_calctmr(self.h, nlay, nwn, _ptr(wn), _ptr(t), _ptr(tz), _ptr(o), _ptr(tmr)))
        return tmr

    def rtm(self, iout, irt, wn, t, tz, o, tmpsfc, reflc, emiss, idu=1):
        """Returns dict(rad, tb, rup, rdn, trtot, tmpsfc) -- tmpsfc comes back as 2.75 for irt 2,3."""
        wn = _f(wn)
        nwn, nlay = wn.shape[0], np.asarray(t).shape[0]
        t, tz, o = _f(t, (nlay,)), _f(tz, (nlay + 1,)), _f(o, (nwn, nlay))
        reflc, emiss = _f(reflc, (nwn,)), _f(emiss, (nwn,))
        ts = C.c_double(float(tmpsfc))
        rup, trtot, rdn, rad, tb = (np.zeros(nwn) for _ in range(5))
        self._check(self.lib.mrtm_rtm(self.h, int(iout), int(irt), nwn, _ptr(wn), nlay, _ptr(t), _ptr(tz), _ptr(o),
                                      C.cast(C.byref(ts), C.c_void_p), _ptr(rup), _ptr(trtot), _ptr(rdn),
                                      _ptr(reflc), _ptr(emiss), _ptr(rad), _ptr(tb), int(idu)))
        return dict(rad=rad, tb=tb, rup=rup, rdn=rdn, trtot=trtot, tmpsfc=ts.value)

    # ---- fused batched path ---------------------------------------------------------------------
    def profiles(self, wn, dvset, prof, scor, irt, tmpsfc, emiss, reflc, cntnm=CNTNM_ALL_ONE, iout=1, idu=1,
                 sclcpl=1.0, sclhw=1.0, y0res=0.0, ibrd=0, want_o=False, want_otot_by_mol=False,
                 selection=False, global_range=None, line_mode=0, out=None):
        """prof: dict from synth.synthetic_profiles / profio (arrays with trailing profile dim).
        out: optional dict of caller-owned float64 Fortran-ordered (nwn, nprof) arrays rad, tb, tmr, trtot, rup, rdn
        that receive the spectra (e.g. views of pinned host memory, reused from call to call)."""
        wn = _f(wn)
        nwn, nlay, nprof, nmol = wn.shape[0], int(prof["nlay"]), int(prof["nprof"]), int(prof["nmol"])
        p, t, clw, wbrodl = (_f(prof[k], (nlay, nprof)) for k in ("p", "t", "clw", "wbrodl"))
        tz = _f(prof["tz"], (nlay + 1, nprof))
        wkl = _f(prof["wkl"], (MXMOL, nlay, nprof))
        scor = None if scor is None else _f(scor, (42, 9, nlay, nprof))
        ts = _f(np.broadcast_to(np.asarray(tmpsfc, dtype=np.float64), (nprof,)).copy())
        emiss, reflc = _f(emiss, (nwn,)), _f(reflc, (nwn,))
        if out is not None:
            outs = {}
            for k in ("rad", "tb", "tmr", "trtot", "rup", "rdn"):
                v = out[k]
                if v.dtype != np.float64 or v.size != nwn * nprof or not (v.flags.f_contiguous or v.flags.c_contiguous and min(v.shape) == 1):
                    raise ValueError("out[%r] must be a contiguous float64 array of nwn*nprof elements" % k)
                outs[k] = v
        else:
            outs = {k: np.zeros((nwn, nprof), order="F") for k in ("rad", "tb", "tmr", "trtot", "rup", "rdn")}
        o = np.zeros((nwn, nlay, nprof), order="F") if want_o else None
        otot = np.zeros((MXMOL, nwn, nprof), order="F") if want_otot_by_mol else None
        sel = None
        if selection:
            sel = (np.zeros((nwn, nlay, nprof), np.int64, order="F"), np.zeros((nwn, nlay, nprof), np.uint64, order="F"))
        gr = global_range or (None, None, 0)
        opts = self._opts(gr[0], gr[1], gr[2], sel, line_mode)
        c7 = _f(np.array(cntnm, dtype=np.float64), (7,))
        self._check(self.lib.mrtm_profiles(
            self.h, nprof, nwn, _ptr(wn), float(dvset), nlay, _ptr(p), _ptr(t), _ptr(tz), _ptr(clw), nmol,
            _ptr(wkl), _ptr(wbrodl), _ptr(scor), float(sclcpl), float(sclhw), float(y0res), _ptr(c7),
            int(ibrd), int(irt), int(iout), int(idu), _ptr(ts), _ptr(emiss), _ptr(reflc),
            _ptr(outs["rad"]), _ptr(outs["tb"]), _ptr(outs["tmr"]), _ptr(outs["trtot"]), _ptr(outs["rup"]),
            _ptr(outs["rdn"]), _ptr(o), _ptr(otot), C.byref(opts)))
        outs["tmpsfc"] = ts
        if want_o:
            outs["o"] = o
        if want_otot_by_mol:
            outs["otot_by_mol"] = otot
        if sel:
            outs["sel_count"], outs["sel_hash"] = sel
        return outs

    def profiles_dev(self, nprof, nwn, nlay, nmol, dvset, ptrs, v1, v2, iw0, irt, cntnm=CNTNM_ALL_ONE, iout=1,
                     idu=1, sclcpl=1.0, sclhw=1.0, y0res=0.0, ibrd=0, stream=None, line_mode=0):
        """Device-resident path.  ptrs: dict of integer device addresses (e.g. torch tensor.data_ptr()):
        wn,p,t,tz,clw,wkl,wbrodl,scor(or 0),tmpsfc,emiss,reflc,rad,tb,tmr,trtot,rup,rdn,o(or 0)."""
        opts = self._opts(v1, v2, iw0, line_mode=line_mode)
        if stream is not None:
            opts.stream = C.c_void_p(int(stream))
        c7 = _f(np.array(cntnm, dtype=np.float64), (7,))

        def a(k):
            v = ptrs.get(k, 0)
            return C.c_void_p(int(v)) if v else None

        self._check(self.lib.mrtm_profiles_dev(
            self.h, int(nprof), int(nwn), a("wn"), float(dvset), int(nlay), a("p"), a("t"), a("tz"), a("clw"),
            int(nmol), a("wkl"), a("wbrodl"), a("scor"), float(sclcpl), float(sclhw), float(y0res), _ptr(c7),
            int(ibrd), int(irt), int(iout), int(idu), a("tmpsfc"), a("emiss"), a("reflc"), a("rad"), a("tb"),
            a("tmr"), a("trtot"), a("rup"), a("rdn"), a("o"), C.byref(opts)))
