"""Cross-section files of the reference: the master file FSCDXS and the per-temperature tables it names.

Host-side mirror of XSREAD (src/monortm_sub.F90:1246-1420) and of the READs inside MONORTM_XSEC_SUB (:1656-1671), plus a
writer for seeded synthetic files in the same formats (the tests cannot depend on /root/reference/cross-sections at run
time).  The product's reader is the C++ mrtm_host_xsread; this module is what the tests compare it with and what
tools/ref_exec_xsec.py feeds the executed reference text from.  Nothing here computes optical depths.
"""
import os
import re

import numpy as np

_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[Ee][-+]?\d+)?")
MX_XS = 38                                       # lblparams.f90:29
# ALIAS(1:4, 1:15) and XSMASS(1:15), BLOCK DATA BXSECT (src/monortm_sub.F90:1445-1474); ' ZZZZZZZZ ' = no alias
XS_ALIAS = [
    ("CLONO2", "CLNO3", None, None), ("HNO4", None, None, None), ("CHCL2F", "CFC21", "CFC21", "F21"),
    ("CCL4", None, None, None), ("CCL3F", "CFCL3", "CFC11", "F11"), ("CCL2F2", "CF2CL2", "CFC12", "F12"),
    ("C2CL2F4", "C2F4CL2", "CFC114", "F114"), ("C2CL3F3", "C2F3CL3", "CFC113", "F113"), ("N2O5", None, None, None),
    ("HNO3", None, None, None), ("CF4", None, "CFC14", "F14"), ("CHCLF2", "CHF2CL", "CFC22", "F22"),
    ("CCLF3", None, "CFC13", "F13"), ("C2CLF5", None, "CFC115", "F115"), ("NO2", None, None, None),
]
XS_MASS = [97.46, 79.01, 102.92, 153.82, 137.37, 120.91, 170.92, 187.38, 108.01, 63.01, 88.00, 86.47, 104.46, 154.47, 45.99]


def molecule_index(name):
    """IXINDX of a (left-justified, upper-case) molecule name: XSREAD :1305-1322; raises like its STOP."""
    nm = name.strip()
    for j, al in enumerate(XS_ALIAS):
        if nm in [a for a in al if a]:
            return j
    raise ValueError("THE NAME: %s IS NOT ONE OF THE CROSS SECTION MOLECULES" % name)


# ------------------------------------------------------------------------------------------ writers (synthetic data)
def write_xs_file(path, amol, v1x, v2x, data, temp, pres, torr=False):
    """One temperature file: header FORMAT 910 (A10,2F10.4,I10,3G10.3,3A10), then list-directed values, 10 per line."""
    data = np.asarray(data, dtype=np.float64)
    src3 = "      TORR" if torr else "     synth"
    with open(path, "w") as f:
        f.write("%10s%10.4f%10.4f%10d%10.3g%10.3g%10.3g%10s%10s%10s\n" %
                (amol[:10].rjust(10), v1x, v2x, len(data), temp, pres, float(np.max(data)), "  monortm_", "b200 test ", src3))
        for i in range(0, len(data), 10):
            f.write("".join("%10.3E" % v for v in data[i:i + 10]) + "\n")


def write_fscdxs(path, records):
    """records: list of (name, v1, v2, dv, [file names, ascending temperature]).  FORMAT 915 (A10,2F10.4,F10.8,I5,5X,I5,A1,4X,6A10)
    after the two records FORMAT 905 skips."""
    with open(path, "w") as f:
        f.write("MOLECULE       V1        V2        DV     NTEMP    FORMAT    FILE1     FILE2     FILE3     FILE4     FILE5     FILE6\n")
        f.write("   \n")
        for name, v1, v2, dv, files in records:
            assert len(files) <= 6 and all(len(x) <= 10 for x in files)
            f.write("%-10s%10.4f%10.4f%10.8f%5d%5s%5s%1s%4s%s\n" %
                    (name[:10], v1, v2, dv, len(files), "", "", "N", "", "".join("%-10s" % x for x in files)))
        f.write("%\n")


def synthetic_set(directory, seed=20261018):
    """A small seeded cross-section data set in the reference's formats: two molecules, three spectral regions, 1-3
    temperature files, one of them with its pressure in TORR.  Every table starts above 0 cm-1 (the radiation term that
    MONORTM_XSEC_SUB divides out vanishes at 0).  Returns the FSCDXS record list."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(directory, "xs"), exist_ok=True)
    recs = []

    def region(name, tag, v1, v2, npts, temps, pres, torr):
        v = np.linspace(v1, v2, npts)
        files = []
        for k, (tt, pp) in enumerate(zip(temps, pres)):
            bumps = sum(a * np.exp(-0.5 * ((v - c) / w) ** 2) for a, c, w in
                        zip(rng.uniform(0.2, 1.0, 5), rng.uniform(v1, v2, 5), rng.uniform(0.02, 0.3, 5) * (v2 - v1) / 10))
            dat = 1e-20 * (0.05 + bumps) * (1.0 + 0.002 * (tt - 250.0)) * v / (v + 1.0)
            dat = np.array([float("%10.3E" % x) for x in dat])
            fn = "xs/%s%sT%d" % (name[:4], tag, k + 1)
            write_xs_file(os.path.join(directory, fn), name, v1, v2, dat, tt, pp, torr)
            files.append(fn)
        recs.append((name, v1, v2, (v2 - v1) / (npts - 1), files))

    region("HNO3", "A", 2.0, 12.0, 1001, (203., 233., 273., 296.), (250., 250., 250., 250.), False)
    region("HNO3", "B", 40.0, 44.0, 801, (220., 296.), (100., 300.), False)
    region("F11", "A", 18.0, 21.0, 601, (250.,), (150.,), True)
    region("F11", "B", 300.0, 310.0, 101, (250.,), (150.,), False)          # outside a microwave run: XSREAD must skip it
    write_fscdxs(os.path.join(directory, "FSCDXS"), recs)
    return recs


# ------------------------------------------------------------------------------------------ reader (mirror of XSREAD + READs)
def _read_xs_file(path):
    with open(path) as f:
        hdr = f.readline().rstrip("\n").ljust(100)
        # values written with 1PE10.3 touch when negative ("0.000E+00-1.025E-26" in the shipped HNO3 files): split on the
        # number syntax, not on blanks
        vals = np.array(_NUM.findall(f.read().replace("D", "E").replace("d", "E")), dtype=np.float64)
    v1x, v2x = float(hdr[10:20]), float(hdr[20:30])
    npts = int(hdr[30:40])
    temp, pres = float(hdr[40:50]), float(hdr[50:60])
    torr = hdr[90:100] == "      TORR"
    if len(vals) < npts:
        raise ValueError("%s: %d values, header says %d" % (path, len(vals), npts))
    return dict(v1x=v1x, v2x=v2x, npts=npts, t=temp, pres=pres * (1013. / 760 if torr else 1.0), torr=torr, data=vals[:npts],
                pres_raw=pres)


def read_regions(directory, names, xv1, xv2):
    """XSREAD for the requested molecule names over [xv1, xv2] followed by the file READs of MONORTM_XSEC_SUB.  Returns the
    regions ordered by requested molecule, then by their order in FSCDXS: dict(ixmol (index into `names`), index (IXINDX),
    v1fx, v2fx, xdoplr, files=[...])."""
    idx = [molecule_index(n) for n in names]
    out = [[] for _ in names]
    found = [False] * len(names)
    with open(os.path.join(directory, "FSCDXS")) as f:
        f.readline()
        f.readline()
        for rec in f:
            rec = rec.rstrip("\n").ljust(120)
            if rec[0] == "*":
                continue
            if rec[0] == "%":
                break
            xname = rec[0:10].strip()
            # blanks inside a numeric field are ignored by a Fortran formatted READ (the shipped FSCDXS has a record with an
            # 11-character name that runs into the V1 field)
            v1x, v2x = float(rec[10:20].replace(" ", "") or 0), float(rec[20:30].replace(" ", "") or 0)
            ntemp = int(rec[40:45].replace(" ", "") or 0)
            files = [rec[60 + 10 * k:70 + 10 * k].strip() for k in range(ntemp)]
            for i, j in enumerate(idx):
                if xname in [a for a in XS_ALIAS[j] if a]:
                    found[i] = True
                    if v2x > xv1 and v1x < xv2:
                        if len(out[i]) >= 5:                 # the arrays hold 5 regions (the reference tests > 6 and overruns)
                            raise ValueError("XSREAD - NSPECR .GT. 5")
                        xdoplr = 3.58115E-07 * (0.5 * (v1x + v2x)) * np.sqrt(296.0 / XS_MASS[j])
                        out[i].append(dict(ixmol=i, index=j, v1fx=v1x, v2fx=v2x, xdoplr=float(xdoplr),
                                           files=[_read_xs_file(os.path.join(directory, fn)) for fn in files]))
    if not all(found):
        raise ValueError("MOLECULE SELECTED - %s - IS NOT FOUND ON FILE FSCDXS" % names[found.index(False)])
    return [r for rs in out for r in rs]
