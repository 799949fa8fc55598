"""ref_exec_xsec.py -- run the reference's MONORTM_XSEC_SUB + convolve (src/monortm_sub.F90:1540-1834) on the CPU through the
mechanical translator tools/f90fn.py.  TEST INFRASTRUCTURE, used in THIS container only (it reads /root/reference/src) by
tools/gen_ref_goldens.py to produce tests/golden/ref_xsec_*.npz.

The translator does not execute Fortran I/O (READ / OPEN are skipped), and the routine reads its tables from files in the
middle of the computation (loop 3000, :1656-1671).  The text handed to the translator is therefore the reference's own lines
with exactly two statements replaced -- the header READ (:1662-1665) and the data READ (:1671) -- by calls to Python
callables that deliver what the READ would have delivered from a cross-section file in the reference's format
(monortm_b200/xsfile.py reads such files; FORMAT 910 and list-directed data).  Every arithmetic statement -- the
temperature interpolation, the removal of the radiation term, the pressure convolution with its hand-written GOTO loops,
the sums over regions, molecules and layers and the final RADFN -- is the reference's text, executed as written.
`convolve` is translated without any change.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import f90fn  # noqa: E402
import f90rt as rt  # noqa: E402

REF = "/root/reference/src/"
F = rt.FArr
_ns = None
_files = {}          # (ixtemp, ixsr, ixmol) -> dict(v1x, v2x, npts, t, pres, torr, data)

HDR_READ = "READ (ifile,910) AMOL,V1x,V2x,NPTSx,"
DAT_READ = "read(ifile,*)(xsdat(i,ixtemp),i=1,nptsx)"


def _patched_text():
    lines = open(REF + "monortm_sub.F90", errors="replace").read().split("\n")
    sub = lines[1539:1834]                     # SUBROUTINE MONORTM_XSEC_SUB ... end of convolve
    assert sub[0].strip().upper().startswith("SUBROUTINE MONORTM_XSEC_SUB") and sub[-1].strip() == "end"
    out, i, n_hdr, n_dat = [], 0, 0, 0
    while i < len(sub):
        s = sub[i]
        if s.strip().startswith(HDR_READ):
            # READ (ifile,910) AMOL,V1x,V2x,NPTSx, & TX(ixtemp,ixsr,ixmol),PRES, & SMAX,SOURCE   (three lines)
            assert sub[i + 1].strip().startswith("TX(ixtemp,ixsr,ixmol),PRES,") and sub[i + 2].strip() == "SMAX,SOURCE"
            out += ["      call xs_hdr(ixtemp,ixsr,ixmol,xshdr)",
                    "      V1x = xshdr(1)", "      V2x = xshdr(2)", "      NPTSx = xshdr(3)",
                    "      TX(ixtemp,ixsr,ixmol) = xshdr(4)", "      PRES = xshdr(5)",
                    "      if (xshdr(6) .gt. 0.5) source(3) = ctorr", "      if (xshdr(6) .le. 0.5) source(3) = '          '"]
            i += 3
            n_hdr += 1
            continue
        if s.strip() == DAT_READ:
            out.append("      call xs_dat(ixtemp,ixsr,ixmol,xsdat)")
            i += 1
            n_dat += 1
            continue
        if s.strip().startswith("dimension xsdat(150000,6),xspd(150000)"):
            out.append(s)
            out.append("      dimension xshdr(8)")
            i += 1
            continue
        out.append(s)
        i += 1
    assert n_hdr == 1 and n_dat == 1
    return "\n".join(out) + "\n"


def _xs_hdr(ixtemp, ixsr, ixmol, xshdr):
    f = _files[(int(ixtemp), int(ixsr), int(ixmol))]
    xshdr.a[:6] = [f["v1x"], f["v2x"], f["npts"], f["t"], f["pres_raw"], 1.0 if f["torr"] else 0.0]   # PRES as the file has it
    return (None,)


def _xs_dat(ixtemp, ixsr, ixmol, xsdat):
    f = _files[(int(ixtemp), int(ixsr), int(ixmol))]
    xsdat.a[:f["npts"], int(ixtemp) - 1] = f["data"]
    return (None,)


def load():
    global _ns
    if _ns is None:
        tmp = os.path.join("/tmp", "monortm_xsec_sub_patched_%d.f90" % os.getpid())
        with open(tmp, "w") as fh:
            fh.write(_patched_text())
        try:
            _ns = f90fn.load([REF + "PhysConstants.f90", REF + "lblparams.f90", REF + "lblrtm_sub.f90", REF + "RTMmono.f90", tmp],
                             externals={"xs_hdr": _xs_hdr, "xs_dat": _xs_dat})
        finally:
            os.unlink(tmp)
    return _ns


def _common(ns, blk, name):
    """(offset, shape) of a member of a COMMON block"""
    tr = ns["__translator__"]
    u, members = tr.commons[blk][0]
    for v in members:
        if v.name == name:
            return v.common_off
    raise KeyError(name)


def xsec_sub(wn, p, t, regions, xamnt):
    """regions: list of dicts (monortm_b200.xsfile.read_regions): ixmol (0-based), v1fx, v2fx, xdoplr, files=[dict(v1x, v2x,
    npts, t, pres, torr, data)] in FSCDXS order; xamnt (mx_xs, nlay).  Returns odxsec (nwn, nlay)."""
    ns = load()
    mx_xs, mxlay = ns["M_lblparams"].mx_xs, ns["M_lblparams"].mxlay
    nwnmx = ns["M_rtmmono"].nwnmx
    nwn, nlay = len(wn), len(p)
    px, xr = ns["C_pathx"].s, ns["C_xsectr"].s
    px[:] = 0.0
    xr[:] = 0.0
    _files.clear()
    nmol = 1 + max(r["ixmol"] for r in regions)
    px[_common(ns, "pathx", "ixmols")] = nmol
    o_xamnt = _common(ns, "pathx", "xamnt")
    for i in range(nmol):
        for il in range(nlay):
            px[o_xamnt + i + il * mx_xs] = float(xamnt[i, il])
    off = {k: _common(ns, "xsectr", k) for k in ("v1fx", "v2fx", "ntempf", "nspecr", "xdoplr")}
    count = {}
    for r in regions:
        m = r["ixmol"]
        k = count.get(m, 0)
        count[m] = k + 1
        xr[off["v1fx"] + k + 5 * m] = r["v1fx"]
        xr[off["v2fx"] + k + 5 * m] = r["v2fx"]
        xr[off["ntempf"] + k + 5 * m] = len(r["files"])
        xr[off["xdoplr"] + k + 5 * m] = r["xdoplr"]
        for j, f in enumerate(r["files"]):
            _files[(j + 1, k + 1, m + 1)] = f
    for m, k in count.items():
        xr[off["nspecr"] + m] = k
    for m in range(nmol):
        xr[off["nspecr"] + m] = int(xr[off["nspecr"] + m])
    wnp = np.zeros(nwnmx)
    wnp[:nwn] = wn
    pp, tp = np.zeros(mxlay), np.zeros(mxlay)
    pp[:nlay], tp[:nlay] = p, t
    od = F.zeros((nwnmx, mxlay))
    ns["monortm_xsec_sub"](F(wnp), nwn, F(pp), F(tp), nlay, od)
    return od.a[:nwn, :nlay].copy()


def convolve(xspd, v1x, v2x, delvx, pd, hwdop, tave, pave, wn):
    ns = load()
    nwnmx = ns["M_rtmmono"].nwnmx
    a = np.zeros(150000)
    a[:len(xspd)] = xspd
    wnp = np.zeros(nwnmx)
    wnp[:len(wn)] = wn
    out = F.zeros(nwnmx)
    ns["convolve"](F(a), float(v1x), float(v2x), float(delvx), float(pd), float(hwdop), float(tave), float(pave), 0.0, F(wnp), len(wn), out)
    return out.a[:len(wn)].copy()
