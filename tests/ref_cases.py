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


def _ir(wn, nlay=3, **kw):
    """a case above the microwave: a few random lines inside the range + every continuum branch it touches"""
    return dict(n_filler=16, nlay=nlay, wn=list(wn), irt=1, emis=0.8,
                line_kw=dict(vmin=float(wn[0]), vmax=float(wn[-1]), with_physical=False), **kw)


# the continuum branches of contnm.f90 above 820 cm-1, each range chosen to cross the gates and table ends of the branches
CASES.update({
    # H2O foreign > 600 (:436-447), CO2 XFACCO2 (:510-513), O2 fundamental (:657), N2 fundamental (:963), Rayleigh (:1107)
    "ir_a_620_3050": _ir([620., 700.5, 850., 1000., 1345., 1500.3, 1700., 1849., 1995., 2002., 2100.7, 2385., 2500., 2899.,
                          2990., 3050.]),
    # N2 overtone (:1025), O2 1.27 um (:709), first Chappuis points (:536), O2 9100-11000 (:745)
    "ir_b_4335_9300": _ir([4335., 4400.2, 4600., 4905., 4915., 6000., 7530., 7540.5, 8000., 8495., 8505., 8915., 8925., 9050.,
                           9105., 9300.4], cntnm=(1., 1., 1., 0.9, 1.1, 1.2, 0.8)),
    # O2 9100-11000, A band (:773), Chappuis
    "ir_c_9095_13300": _ir([9095., 9200., 9375., 9439., 10000.5, 10995., 11005., 12000., 12960., 12962., 13100.3, 13221.,
                            13225., 13300.]),
    # O2 visible (:807), Chappuis
    "vis_d_14995_19900": _ir([14995., 15005., 15140., 15145., 16000.5, 17000., 18000., 19900.], nlay=2),
    # end of Chappuis (24665), start of Hartley-Huggins (27370, :555)
    "vis_e_24000_28900": _ir([24000., 24560., 24660., 24670., 25000., 27365., 27375., 28000.5, 28900.], nlay=2),
    # Herzberg (:834), Hartley-Huggins with the save/restore beyond 40800 (:573-599), first UV points (:603)
    "uv_f_35990_40900": _ir([35990., 36005., 38000., 39999., 40001., 40790., 40805., 40820., 40900.], nlay=2),
    # UV Hartley-Huggins with the save/restore below 40800 (:619-640)
    "uv_g_40700_45000": _ir([40700., 40795., 40805., 41000., 43000.5, 45000.], nlay=2),
    # end of the O3 UV table (54000), O2 far UV (:857), Herzberg
    "uv_h_53000_57900": _ir([53000., 53990., 54010., 56000., 56735., 56745., 57000.5, 57900.], nlay=2),
})


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
    elif isinstance(recipe, (list, tuple)):
        wn = np.array(recipe, dtype=np.float64)
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
