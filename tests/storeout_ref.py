"""Independent (Python) statement of STOREOUT's record formats (src/monortm_sub.F90:611-667, FORMATs 11/21
:780-782) used to check the C++ writer; Fortran edit descriptors are expressed with Python % formatting."""
import numpy as np

CLIGHT = 2.99792458E+10
HMOLC = ['  H2O   ', '  CO2   ', '   O3   ', '  N2O   ', '   CO   ', '  CH4   ', '   O2   ', '   NO   ',
         '  SO2   ', '  NO2   ', '  NH3   ', ' HNO3   ', '   OH   ', '   HF   ', '  HCL   ', '  HBR   ',
         '   HI   ', '  CLO   ', '  OCS   ', ' H2CO   ', ' HOCL   ', '   N2   ', '  HCN   ', ' CH3CL  ',
         ' H2O2   ', ' C2H2   ', ' C2H6   ', '  PH3   ', ' COF2   ', '  SF6   ', '  H2S   ', ' HCOOH  ',
         '  HO2   ', '   O+   ', ' ClONO2 ', '   NO+  ', '  HOBr  ', ' C2H4   ', ' CH3OH  ']


def A(w, s):
    return s[:w] if len(s) >= w else s.rjust(w)


def F(w, d, v):
    s = "%.*f" % (d, v)
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return "*" * w if len(s) > w else s.rjust(w)


def E1P(w, d, v):
    s = "%.*E" % (d, v)
    m, e = s.split("E")
    if len(e) > 3:                      # three-digit exponent: the letter is dropped
        s = m + e
    return "*" * w if len(s) > w else s.rjust(w)


def header(nwn, wn, id_mol):
    giga = wn[0] < 100
    units = "FREQ(GHz)   " if giga else "FREQ(cm-1)  "
    out = ["MONORTM RESULTS:", "----------------", "NWN :" + "%8d" % nwn + " " * 101 + A(42, "Molecular Optical Depths -->")]
    h = A(5, "PROF ") + A(10, units) + A(11, "BT(K) ") + A(11, "TMR(K)") + A(22, "  RAD(W/cm2_ster_cm-1)") + A(8, "TRANS") + \
        A(8, "PWV") + A(8, "CLW") + A(8, "TBOUND") + A(8, "EMIS") + A(8, "REFL") + A(9, "ANGLE") + A(12, "TOTAL_OD")
    for im in id_mol:
        h += A(12, HMOLC[im - 1])
    h += A(12, "XSEC_OD")
    return out + [h]


def rows(npr, wn, tb, tmr, rad, trtot, wvcolmn, clwcolmn, tmpsfc, emiss, reflc, angle, otot, otot_by_mol, odxtot, id_mol):
    giga = wn[0] < 100
    out = []
    for i in range(len(wn)):
        freq = wn[i] * CLIGHT / 1.E9 if giga else wn[i]
        r = "%5d" % npr + F(10, 3, freq) + F(11, 5, tb[i]) + F(11, 5, tmr[i]) + E1P(21, 9, rad[i]) + F(9, 5, trtot[i]) + \
            F(8, 4, wvcolmn) + F(8, 4, clwcolmn) + F(8, 2, tmpsfc) + F(8, 2, emiss[i]) + F(8, 2, reflc[i]) + F(9, 3, angle) + \
            E1P(12, 4, otot[i])
        for im in id_mol:
            r += E1P(12, 4, otot_by_mol[im - 1, i])
        r += E1P(12, 4, odxtot[i])
        out.append(r)
    return out


def layer_sums(o, o_by_mol, oc, odxsec=None):
    """OTOT, OTOT_BY_MOL(39,nwn), ODXTOT with STOREOUT's summation order (:643-656)."""
    nwn, nlay = o.shape
    otot, odx = np.zeros(nwn), np.zeros(nwn)
    obm = np.zeros((39, nwn))
    for j in range(nlay):
        otot = otot + o[:, j]
        if odxsec is not None:
            odx = odx + odxsec[:, j]
        obm = (obm + o_by_mol[:, :, j].T) + oc[:, :, j].T
    return otot, obm, odx


def id_mols(wkl, wbrodl, nmol):
    w = np.array(wkl, copy=True)
    if nmol < 22:
        w[21, :] = wbrodl
    tot = np.zeros(39)
    for l in range(w.shape[1]):
        tot = tot + w[:, l]
    return [im + 1 for im in range(39) if tot[im] > 0]
