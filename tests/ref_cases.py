"""The small seeded cases behind tests/golden/ref_case_*.npz (outputs of the reference's own Fortran text, see
tools/gen_ref_goldens.py).  Shared by the generator and by the tests that regenerate the inputs."""
import hashlib

import numpy as np

import harness

LC = (-1.0, -3.0, -5.0)
KW = dict(n_co2=8, n_sdep=6, n_generic_lc=6, brd_fraction=0.25)


def _zone(centres, mult):
    offs = np.array(sorted({s * m for m in mult for s in (1.0, -1.0)})) * 1e-6
    c = np.asarray(centres, dtype=np.float64)
    return np.unique((c[:, None] * (1.0 + offs[None, :])).ravel())


def special_centres(n_filler, kw):
    ls = harness.synthetic_store(n_filler, v1=0.0, v2=55.0, **kw)
    sd, co2, glc, o2lc = [], [], [], []
    for m in range(39):
        n, j = int(ls.nblm[m]), 0
        while j < n:
            xg = float(ls.xg[m, j])
            v = float(ls.xnu0[m, j])
            if abs(ls.sdep[m, j]) > 1e-4:
                sd.append(v)
            elif m == 1:
                co2.append(v)
            elif m != 6 and xg in LC:
                glc.append(v)
            elif m == 6 and xg == -1.0:
                o2lc.append(v)
            j += 2 if xg in LC else 1
    return sd, co2, glc, o2lc


# name -> parameters of harness.make_case (+ "wn" recipe)
CASES = {
    "c1_channels_dn": dict(n_filler=96, nlay=12, wn="c1", irt=3),
    "voigt_zone_cov": dict(n_filler=48, nlay=10, wn="zone", irt=1, line_kw=KW, emis=0.7),
    "voigt_zone_ibrd": dict(n_filler=48, nlay=8, wn="zone_small", irt=1, ibrd=1, line_kw=KW),
    "cloud_limb_scaled": dict(n_filler=64, nlay=8, wn="sweep", dvset=0.025, irt=2, clw=True,
                              cntnm=(0.7, 1.2, 1.0, 1.0, 1.0, 0.0, 1.0), sclcpl=0.87, sclhw=1.1, y0res=0.01),
    "nmol7_up_wide": dict(n_filler=64, nlay=6, wn="wide", irt=1, nmol=7, emis=0.9, tmpsfc=285.0),
}


def build_case(spec):
    spec = dict(spec)
    recipe = spec.pop("wn")
    kw = spec.get("line_kw") or {}
    if recipe == "c1":
        wn = np.array([0.789344, 0.79828, 1.043027, 1.051763, 0.7417, 2.0, 3.96, 6.1, 18.58, 25.0])
    elif recipe in ("zone", "zone_small"):
        sd, co2, glc, o2lc = special_centres(spec["n_filler"], kw)
        if recipe == "zone":
            wn = _zone(sd[:4] + co2[:4] + glc[:4] + o2lc[:2], (0.0, 0.4, 30.0))
        else:
            wn = _zone(sd[:2] + co2[:2] + glc[:2] + o2lc[:1], (0.0, 3.0))
        wn = wn[(wn > 0.4) & (wn < 55.0)]
    elif recipe == "sweep":
        wn = 0.2 + spec["dvset"] * np.arange(33)
    elif recipe == "wide":
        wn = np.array([0.1, 0.74, 1.9, 2.05, 3.96, 6.11, 10.8, 18.6, 25.1, 33.0, 47.0, 54.9])
    else:
        raise ValueError(recipe)
    return harness.make_case(wn=wn, **spec)


def inputs_digest(case):
    """sha256 over everything the case feeds the hot path"""
    h = hashlib.sha256()
    ls, pr = case["ls"], case["prof"]
    n = int(max(ls.nblm))
    for name in ("nblm",):
        h.update(np.ascontiguousarray(getattr(ls, name)).tobytes())
    for name in ("iso", "xnu0", "deltnu", "e", "alps", "alpf", "x", "xg", "s0", "rmol", "sdep"):
        h.update(np.ascontiguousarray(getattr(ls, name)[:, :n]).tobytes())
    for name in ("brd_mol_flg", "brd_mol_tmp", "brd_mol_hw", "brd_mol_shft"):
        h.update(np.ascontiguousarray(getattr(ls, name)[:, :, :n]).tobytes())
    for name in ("p", "t", "tz", "clw", "wkl", "wbrodl"):
        h.update(np.ascontiguousarray(pr[name]).tobytes())
    h.update(np.ascontiguousarray(case["wn"]).tobytes())
    h.update(np.ascontiguousarray(case["emiss"]).tobytes())
    return h.hexdigest()
