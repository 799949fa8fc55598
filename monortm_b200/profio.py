"""Reader for MONORTM_PROF.IN (IATM=0 layer input, TAPE7-like), following the READ statements and
formats of src/monortm.f90:380-408, 599, 608-612 and the mixing-ratio -> column-density conversion
of :423-483.  Harness side; in production the unchanged Fortran host reads the file.
"""
import numpy as np

MXMOL = 39


def _fld(s, a, b, conv, default=0):
    t = s[a:b]
    return conv(t) if t.strip() else default


def read_prof_in(path):
    """Returns a list of profile dicts (one per profile in the file) with Fortran-ordered arrays and a
    trailing profile dimension of 1, compatible with Session.profiles()."""
    with open(path) as f:
        lines = f.read().split("\n")
    out, i = [], 0
    while i < len(lines) and lines[i].strip():
        h = lines[i].ljust(80)
        i += 1
        # FORMAT 925 (1X,I1,I3,I5,F10.6,2A8,4X,F8.2,4X,F8.2,5X,F8.3,5X,I2)
        iform, nlay, nmol = _fld(h, 1, 2, int), _fld(h, 2, 5, int), _fld(h, 5, 10, int)
        secnt0 = _fld(h, 10, 20, float, 0.0)
        h1, h2, angle = _fld(h, 40, 48, float, 0.0), _fld(h, 52, 60, float, 0.0), _fld(h, 65, 73, float, 0.0)
        p, t, clw = np.zeros(nlay), np.zeros(nlay), np.zeros(nlay)
        tz, pz, altz = np.zeros(nlay + 1), np.zeros(nlay + 1), np.zeros(nlay + 1)
        wkl = np.zeros((MXMOL, nlay), order="F")
        wbrodl = np.zeros(nlay)
        for il in range(nlay):
            s = lines[i].ljust(120)
            i += 1
            if iform == 1:   # 975 / 9752: e15.7, 2f10.4, 3x, i2, 1x, ...
                p[il], t[il] = float(s[0:15]), float(s[15:25])
                rest = s[41:]
            else:            # 974 / 9742: f10.4, f10.4 (TAPE7 IFORM=0: 1PE15.7 replaced by F10.4)
                p[il], t[il] = float(s[0:10]), float(s[10:20])
                rest = s[36:]
            if il == 0:      # both boundaries: 2(f7.2,f8.3,f7.2), then f7.3
                altz[0], pz[0], tz[0] = float(rest[0:7]), float(rest[7:15]), float(rest[15:22])
                altz[1], pz[1], tz[1] = float(rest[22:29]), float(rest[29:37]), float(rest[37:44])
                c = rest[44:51]
            else:            # 23x, (f7.2,f8.3,f7.2), f7.3
                altz[il + 1], pz[il + 1], tz[il + 1] = float(rest[22:29]), float(rest[29:37]), float(rest[37:44])
                c = rest[44:51]
            clw[il] = float(c) if c.strip() else 0.0
            v = _read_8e15(lines[i]); i += 1                 # 978: 8E15.7 -> WKL(1:7), WBRODL
            wkl[0:7, il] = v[0:7]
            wbrodl[il] = v[7]
            k = 7
            while k < nmol:                                  # WKL(8:NMOL), 8 per line
                v = _read_8e15(lines[i]); i += 1
                m = min(8, nmol - k)
                wkl[k:k + m, il] = v[:m]
                k += m
            # mixing ratio input (monortm.f90:423-483)
            wdnsty, wmxrat = wbrodl[il], 0.0
            for m in range(1, nmol):
                if wkl[m, il] > 1:
                    wdnsty += wkl[m, il]
                else:
                    wmxrat += wkl[m, il]
            if wbrodl[il] < 1.0 and wbrodl[il] != 0.0:
                raise ValueError("STOP: WBRODL must be a column density")
            if wdnsty == 0.0 and wmxrat != 0.0:
                raise ValueError("STOP 'WMXRAT AND/OR WDNSTY NOT PROPERLY SPECIFIED IN PATH'")
            if not wmxrat < 1.0:
                raise ValueError("STOP 'WMXRAT EXCEEDS 1.0'")
            wdrair = wdnsty / (1.0 - wmxrat)
            for m in range(nmol):
                if wkl[m, il] < 1.0:
                    wkl[m, il] = wkl[m, il] * wdrair
        irt = 1 if angle > 90.0 else (3 if angle < 90.0 else 2)      # monortm.f90:383-385
        F = dict(order="F")
        out.append(dict(nlay=nlay, nprof=1, nmol=nmol, irt=irt, angle=angle, h1=h1, h2=h2, secnt0=secnt0,
                        p=np.asarray(p.reshape(nlay, 1), **F), t=np.asarray(t.reshape(nlay, 1), **F),
                        clw=np.asarray(clw.reshape(nlay, 1), **F), wbrodl=np.asarray(wbrodl.reshape(nlay, 1), **F),
                        tz=np.asarray(tz.reshape(nlay + 1, 1), **F), pz=pz, altz=altz,
                        wkl=np.asarray(wkl.reshape(MXMOL, nlay, 1), **F)))
    return out


def _read_8e15(s):
    s = s.rstrip("\n")
    vals = []
    for k in range(8):
        t = s[15 * k:15 * (k + 1)]
        vals.append(float(t) if t.strip() else 0.0)
    return vals
