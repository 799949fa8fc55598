"""Seeded synthetic inputs: TAPE3-synth line lists, profiles-synth and the
frequency grids of the five BASELINE configs.  Needed because the reference ships neither a
line file nor example outputs.  Everything is deterministic (numpy PCG64 with fixed seeds).
"""
import numpy as np

from .linefile import REC_DTYPE, coupling_record

RADCT = 6.62606876E-27 * 2.99792458E+10 / 1.3806503E-16   # modm.f90:874
T0 = 296.0
CLIGHT_GHZ = 29.9792458                                    # GHz per cm-1


def lblrtm_strength(s_hitran, vnu):
    """TAPE3 stores S/(v(1-exp(-hcv/kT0))) (src/modm.f90:371-372)."""
    return s_hitran / (vnu * (1.0 - np.exp(-RADCT * vnu / T0)))


def _line(vnu, s_hitran, alfa, hwhm, epp, tmpalf, pshift, mol, iso=1, iflg=0, sdep=0.0, sp_raw=None):
    r = np.zeros((), REC_DTYPE)
    r["vnu"] = vnu
    r["sp"] = np.float32(sp_raw if sp_raw is not None else lblrtm_strength(s_hitran, vnu))
    r["alfa"], r["hwhm"], r["epp"] = alfa, hwhm, epp
    r["tmpalf"], r["pshift"] = tmpalf, pshift
    r["mol"] = mol + 100 * iso
    r["iflg"] = iflg
    r["sdep"] = sdep
    return r


def physical_seed_lines():
    """A handful of real microwave features (approximate parameters): H2O 22/183/325/380/448/557/752
    GHz, the O2 60 GHz complex with first-order coupling records (IFLG=1), O2 118.75 GHz, O2 submm
    lines and the O2 non-resonant (zero-frequency) line with IFLG=3.  Returns a list of groups
    (a line followed by its coefficient records)."""
    groups = []
    h2o = [  # vnu, S(HITRAN), air hw, self hw, E'', n, shift
        (0.741691, 4.39e-25, 0.0900, 0.4699, 446.5107, 0.755, -0.0003),
        (6.114567, 7.76e-23, 0.0992, 0.5170, 136.1639, 0.770, -0.0028),
        (10.845940, 9.00e-23, 0.0929, 0.4600, 315.7795, 0.690, -0.0020),
        (12.682023, 8.10e-22, 0.0907, 0.4520, 212.1564, 0.710, -0.0015),
        (14.943707, 8.60e-22, 0.0850, 0.4300, 285.4186, 0.680, 0.0010),
        (18.577385, 5.20e-20, 0.1030, 0.4800, 23.7944, 0.750, 0.0020),
        (25.085124, 3.50e-20, 0.0990, 0.4700, 70.0908, 0.740, -0.0010),
    ]
    for v, s, a, hw, e, n, d in h2o:
        groups.append([_line(v, s, a, hw, e, n, d, mol=1)])
    # O2 60 GHz complex (N-, N+ for odd N), GHz
    o2_ghz = {1: (118.7503, 56.2648), 3: (62.4863, 58.4466), 5: (60.3061, 59.5910), 7: (59.1642, 60.4348),
              9: (58.3239, 61.1506), 11: (57.6125, 61.8002), 13: (56.9682, 62.4112), 15: (56.3634, 62.9980),
              17: (55.7838, 63.5685), 19: (55.2214, 64.1278), 21: (54.6712, 64.6789), 23: (54.1300, 65.2241),
              25: (53.5957, 65.7648), 27: (53.0669, 66.3021), 29: (52.5424, 66.8368), 31: (52.0214, 67.3696),
              33: (51.5034, 67.9009)}
    for n_q, (fm, fp) in o2_ghz.items():
        e_low = 1.4377 * n_q * (n_q + 1)
        s = 3.5e-26 * (2 * n_q + 1) * np.exp(-RADCT * e_low / T0)
        hw = 0.055 - 0.0006 * n_q
        for sign, ghz in ((-1.0, fm), (1.0, fp)):
            v = ghz / CLIGHT_GHZ
            coupled = not (n_q == 1 and sign < 0)     # 118.75 GHz line: no coupling record
            ln = _line(v, s, hw, hw * 1.03, e_low, 0.8, 0.0, mol=7, iflg=1 if coupled else 0)
            grp = [ln]
            if coupled:
                y296 = sign * (0.045 - 0.0035 * n_q)
                y = [y296 * f for f in (1.35, 1.15, 1.0, 0.88)]
                g = [1.5e-2 * f * (1 if n_q < 15 else -1) for f in (1.6, 1.25, 1.0, 0.8)]
                grp.append(coupling_record(y, g, 1))
            groups.append(grp)
    for ghz, s, e_low in ((368.4983, 2.3e-26, 3.96), (424.7631, 8.9e-26, 2.08), (487.2494, 1.2e-26, 3.96)):
        groups.append([_line(ghz / CLIGHT_GHZ, s, 0.049, 0.050, e_low, 0.8, 0.0, mol=7)])
    # non-resonant O2 "line" at (nearly) zero frequency, IFLG=3: pressure dependence of its width
    nr = _line(1.0e-6, 0.0, 0.0493, 0.0500, 0.0, 0.8, 0.0, mol=7, iflg=3, sp_raw=4.0e-25)
    groups.append([nr, coupling_record([0.02, 0.02, 0.02, 0.02], [0.004, 0.004, 0.004, 0.004], 3)])
    return groups


def synthetic_records(n_filler=4096, seed=20260101, vmax=80.0, with_physical=True, n_co2=0,
                      n_sdep=0, n_generic_lc=0, brd_fraction=0.0, vmin=0.02):
    """Build a TAPE3-synth record array in file order.

    filler: molecule in {1,3,4,5,7} with weights {.15,.6,.1,.05,.1}; S log-uniform 1e-30..1e-22;
    air width U(0.03,0.11); self width U(0.05,0.5); E'' U(0,3000); n U(0.5,0.8); shift U(-5e-3,5e-3);
    isotopologue 1-3.  Options add coverage lines: CO2 lines (IFLG 0/1/5 groups), lines
    with speed dependence, coupled lines of other molecules, species-broadening data (IBRD=1).
    """
    rng = np.random.default_rng(seed)
    groups = physical_seed_lines() if with_physical else []
    mols = np.array([1, 3, 4, 5, 7])
    pm = np.array([.15, .6, .1, .05, .1])
    mol = rng.choice(mols, size=n_filler, p=pm)
    vnu = rng.uniform(vmin, vmax, n_filler)      # vmin > 0.02: a list above the microwave (continuum branches of SURVEY 8f-2)
    s = 10.0 ** rng.uniform(-30, -22, n_filler)
    alfa = rng.uniform(0.03, 0.11, n_filler)
    hwhm = rng.uniform(0.05, 0.5, n_filler)
    epp = rng.uniform(0, 3000, n_filler)
    tmpalf = rng.uniform(0.5, 0.8, n_filler)
    pshift = rng.uniform(-5e-3, 5e-3, n_filler)
    iso = rng.integers(1, 4, n_filler)
    for i in range(n_filler):
        groups.append([_line(vnu[i], s[i], alfa[i], hwhm[i], epp[i], tmpalf[i], pshift[i], int(mol[i]), int(iso[i]))])
    # a few H2O lines with zero self width exercise the 5x fix-up (modm.f90:841)
    for i in range(min(8, n_filler // 64)):
        v = rng.uniform(max(1.0, vmin), vmax)
        groups.append([_line(v, 10.0 ** rng.uniform(-26, -23), 0.08, 0.0, 200.0, 0.7, 1e-3, 1)])
    for i in range(n_sdep):
        v = rng.uniform(0.5, min(vmax, 40.0))
        groups.append([_line(v, 10.0 ** rng.uniform(-25, -22), rng.uniform(0.05, 0.1), rng.uniform(0.2, 0.5),
                             rng.uniform(0, 500), 0.7, 1e-3, int(rng.choice([1, 3])), sdep=rng.uniform(0.06, 0.14))])
    for i in range(n_generic_lc):
        v = rng.uniform(0.5, min(vmax, 60.0))
        fl = int(rng.choice([1, 3]))
        ln = _line(v, 10.0 ** rng.uniform(-25, -22), rng.uniform(0.05, 0.1), rng.uniform(0.2, 0.5),
                   rng.uniform(0, 500), 0.7, 1e-3, int(rng.choice([3, 4])), iflg=fl)
        y = list(rng.uniform(-0.02, 0.02, 4))
        g = list(rng.uniform(-0.01, 0.01, 4))
        groups.append([ln, coupling_record(y, g, fl)])
    # CO2 lines: uncoupled and IFLG=1 groups.  IFLG=5 (foreign+self records) is left out on purpose:
    # the reference's own record walk mis-parses such groups (modm.f90:339 looks behind at XG(I,J-1),
    # so the first line of a run is never blended and its self record is then visited as a line).
    for i in range(n_co2):
        v0 = rng.uniform(5.0, min(vmax, 70.0))
        if i % 2 == 0:
            groups.append([_line(v0, 10.0 ** rng.uniform(-27, -24), 0.07, 0.09, rng.uniform(0, 800), 0.7, -1e-3, 2)])
        else:
            ln = _line(v0, 10.0 ** rng.uniform(-27, -24), 0.07, 0.09, rng.uniform(0, 800), 0.7, -1e-3, 2, iflg=1)
            groups.append([ln, coupling_record(list(rng.uniform(-0.01, 0.01, 4)), list(rng.uniform(-5e-3, 5e-3, 4)), 1)])
    if brd_fraction > 0:
        for grp in groups:
            ln = grp[0]
            if int(ln["mol"]) % 100 <= 7 and rng.uniform() < brd_fraction:
                flg = (rng.uniform(size=7) < 0.4).astype(np.int32)
                dat = np.zeros(21, np.float32)
                dat[0::3] = rng.uniform(0.04, 0.12, 7)
                dat[1::3] = rng.uniform(0.5, 0.8, 7)
                dat[2::3] = rng.uniform(-4e-3, 4e-3, 7)
                ln["brd_flg"] = flg
                ln["brd_dat"] = dat
    groups.sort(key=lambda g: float(g[0]["vnu"]))
    recs = np.zeros(sum(len(g) for g in groups), REC_DTYPE)
    k = 0
    for g in groups:
        for r in g:
            recs[k] = r
            k += 1
    return recs


# --------------------------------------------------------------------------- profiles
def synthetic_profiles(nprof, nlay, seed0=1000, clw_layers=False, nmol=22):
    """profiles-synth.  Returns dict of Fortran-ordered arrays with a trailing profile
    dimension: p,t,clw,wbrodl (nlay,nprof); tz (nlay+1,nprof); wkl (39,nlay,nprof)."""
    f = dict(order="F")
    p = np.zeros((nlay, nprof), **f)
    t = np.zeros((nlay, nprof), **f)
    tz = np.zeros((nlay + 1, nprof), **f)
    clw = np.zeros((nlay, nprof), **f)
    wbrodl = np.zeros((nlay, nprof), **f)
    wkl = np.zeros((39, nlay, nprof), **f)
    plev = 1013.25 * (0.1 / 1013.25) ** (np.arange(nlay + 1) / nlay)
    zlev = -7.0 * np.log(plev / 1013.25)   # km, scale-height altitude
    for ip in range(nprof):
        rng = np.random.default_rng(seed0 + ip)
        tlev = np.maximum(288.2 - 6.5 * zlev, 216.65) + rng.normal(0.0, 3.0)
        tz[:, ip] = tlev
        pl = 0.5 * (plev[:-1] + plev[1:])
        p[:, ip] = pl
        t[:, ip] = 0.5 * (tlev[:-1] + tlev[1:])
        dp = plev[:-1] - plev[1:]
        col = dp * 2.12e22                       # molecules/cm2 of air in the layer
        h2o = 0.01 * (pl / 1013.0) ** 3 * rng.lognormal(0.0, 0.3)
        o3 = 1e-8 + 8e-6 * np.exp(-0.5 * ((np.log(pl) - np.log(10.0)) / 0.9) ** 2)
        vmr = {1: h2o, 2: 4e-4 + 0 * pl, 3: o3, 4: 3.2e-7 + 0 * pl, 5: 1.5e-7 + 0 * pl, 6: 1.7e-6 + 0 * pl,
               7: 0.209 + 0 * pl}
        for m, v in vmr.items():
            if m <= nmol:
                wkl[m - 1, :, ip] = v * col
        wbrodl[:, ip] = (1.0 - 0.209 - 4e-4 - h2o) * col
        if clw_layers:
            k0 = int(np.argmin(np.abs(pl - 800.0)))
            for dk, amt in zip((-1, 0, 1), (0.03, 0.04, 0.03)):
                if 0 <= k0 + dk < nlay:
                    clw[k0 + dk, ip] = amt
    return dict(p=p, t=t, tz=tz, clw=clw, wbrodl=wbrodl, wkl=wkl, nlay=nlay, nprof=nprof, nmol=nmol)


# --------------------------------------------------------------------------- frequency grids
def freq_c1_channels():
    """run/in/MONORTM.IN_IATM0_dn / _MDL_ATM_up: 4 MWR channels (cm-1)."""
    return np.array([0.789344, 0.79828, 1.043027, 1.051763])


def freq_c1_sweep():
    """run/in/MONORTM.IN_MDL_ATM_dn: V1=0.2, V2=1.2, DVSET=0.01 -> 101 points (monortm_sub.F90:278-288)."""
    v1, v2, dv = 0.2, 1.2, 0.01
    n = int(round((v2 - v1) / dv + 1))
    return v1 + np.arange(n) * dv, dv


def freq_c2_sounder():
    ghz = [23.8, 31.4, 50.3, 52.8, 53.596, 54.4, 54.94, 55.5, 57.29, 89.0, 150.0, 157.0,
           183.31 - 7, 183.31 - 3, 183.31 - 1, 183.31 + 1, 183.31 + 3, 183.31 + 7, 190.31]
    return np.sort(np.array(ghz) / CLIGHT_GHZ)


def freq_c3_dense(n=1000000):
    return 5.5e-5 * np.arange(1, n + 1)


def freq_c4_channels(n=1000):
    return np.exp(np.linspace(np.log(0.1), np.log(30.0), n))


def freq_c5(n=10000):
    return np.linspace(0.0055, 55.0, n)
