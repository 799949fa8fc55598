/*
 * monortm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, built -O0 -ffp-contract=off -fcx-fortran-rules to
 * mirror the reference's linuxGNUdbl build, build/makefile.common:195-198) of
 * the monoRTM optical-depth + radiance hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (monortm_b200/) never does.
 *
 * PARITY PIN: the reference ships no golden outputs and no TAPE3 line file, and it cannot be COMPILED here (100 %
 * Fortran, no Fortran compiler in the image), so there is no oracle/_ref binary.  Instead the
 * reference's own source TEXT is executed: tools/f90fn.py translates /root/reference/src/{modm,contnm,RTMmono,
 * lblrtm_sub,CloudOptProp,CntnmFactors,tips_2003,PhysConstants,PlanetEarth,lblparams}.f90 + isotope.incl mechanically,
 * statement by statement, into Python (tools/ref_exec.py), and tools/gen_ref_goldens.py writes its outputs to
 * tests/golden/ref_*.npz (scalar routines and whole MODM + CALCTMR + RTM cases, TIPS_2003 and CONTNM inside).
 * tests/test_ref_goldens.py holds this restatement to those vectors: scalar routines to a few ulp, layer optical
 * depths to 1e-13 (measured: 0 .. 4e-15), TB to 1e-9 K.  What that pin does not cover: libgfortran's own intrinsics
 * (exp, log, pow, tanh here come from glibc / CPython) and the binary128 promotion of d0 literals in
 * CloudOptProp.f90 under -fdefault-real-8 (modelled here with __float128; ~1e-14 relative on the cloud term).
 * Further independent pins: from-scratch numpy physics checks and scipy.special.wofz (tests/test_oracle_units.py).
 *
 * All arrays are Fortran column-major, exactly as the reference holds them.
 */
#ifndef MONORTM_ORACLE_H
#define MONORTM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MXMOL 39      /* lblparams.f90:28 */
#define ORC_MXBRDMOL 7    /* struct_types.f90:25 */

/* Line store as lnfl_mod holds it (src/lnfl_mod.f90:5-13): element (mo,ii) of a
 * (39,IIM) array is at [(mo-1) + (ii-1)*39]; brd arrays (7,7,IIM) at
 * [(mo-1) + (k-1)*7 + (ii-1)*49].  iim is the allocated second dimension. */
typedef struct {
    int64_t iim;
    int64_t nblm[ORC_MXMOL];
    int64_t *iso;
    double *xnu0, *deltnu, *e, *alps, *alpf, *x, *xg, *s0, *rmol, *sdep;
    int32_t *brd_mol_flg;
    double *brd_mol_tmp, *brd_mol_hw, *brd_mol_shft;
} orc_lines;

/* allocate / free a zero-initialised line store */
orc_lines *orc_lines_alloc(int64_t iim);
void orc_lines_free(orc_lines *ln);

/* GET_LNFL (src/lnfl_mod.f90:22-133): read TAPE3 for [v1-25, v2+25], block
 * granular.  Returns 0, or >0 on error (message in orc_last_error()). */
int orc_get_lnfl(const char *hfile, double v1, double v2, orc_lines *ln);

/* MODM (src/modm.f90:21-274) with tips_2003 hoisted to the caller:
 * scor is (42,9,nlay).  o,o_clw,odxsec are (nwn,nlay); o_by_mol,oc are
 * (nwn,39,nlay).  odxsec_in may be NULL (ixsect=0) else (nwn,nlay) added as the
 * reference adds monortm_xsec_sub's result.  sel_count/sel_hash (nwn,nlay) may
 * be NULL: number and order-independent hash of the (molecule,record) pairs
 * passing the cutoff test modm.f90:384.  n_voigt (may be NULL) returns the
 * number of shape evaluations that took the Voigt branch (modm.f90:430). */
int orc_modm(int64_t nwn, const double *wn, double dvset, int64_t nlay,
             const double *p, const double *t, const double *clw,
             double *o, double *o_by_mol, double *oc, double *o_clw, double *odxsec,
             int64_t nmol, const double *wkl, const double *wbrodl,
             double sclcpl, double sclhw, double y0res,
             const double cntnm[7], int64_t ixsect, const double *odxsec_in,
             int64_t ibrd, const double *scor, orc_lines *ln,
             int64_t *sel_count, uint64_t *sel_hash, int64_t *n_voigt);

/* CALCTMR (src/RTMmono.f90:239-325). tz is (0:nlay). */
int orc_calctmr(int64_t nlayrs, int64_t nwn, const double *wn, const double *t,
                const double *tz, const double *o, double *tmr);

/* RTM (src/RTMmono.f90:13-155) incl. RAD_UP_DN (157-221).  tmpsfc is in/out
 * (set to 2.75 for irt 2,3: RTMmono.f90:122). */
int orc_rtm(int64_t iout, int64_t irt, int64_t nwn, const double *wn, int64_t nlay,
            const double *t, const double *tz, const double *o, double *tmpsfc,
            double *rup, double *trtot, double *rdn, const double *reflc,
            const double *emiss, double *rad, double *tb, int64_t idu);

/* building blocks exposed for unit tests */
void orc_w4(double x, double y, double *re, double *im);                 /* modm.f90:1100-1130 */
void orc_sd_humlicek(double x1, double y1, double x2, double y2, double *re, double *im); /* :1150-1251 */
double orc_sdvoigt(double deltnu, double alphal, double alphad, double sdep, int *err);   /* :965-1087 */
double orc_radfn(double vi, double xkt);                                 /* lblrtm_sub.f90:36-97 */
double orc_odclw(double wn, double temp, double clw);                    /* CloudOptProp.f90:29-53 */
double orc_bb_fn(double v, double fbeta);                                /* RTMmono.f90:223-237 */
/* CONTNM(JRAD=0) for one species selector im in {1,2,3,7,22,99} (modm.f90:207-215):
 * fills absrb[0..nptabs-1] on the 1 cm-1 grid; wk is the 60-element /FILHDR/ WK. */
int orc_contnm_one(int64_t im, const double cntnm[7], double pave, double tave,
                   const double *wk, double wbroad, int64_t nmol, double v1, double v2,
                   double v1abs, double v2abs, int64_t nptabs, double *absrb);
/* TIPS_2003 (src/tips_2003.f90:2-298): scor(42,9) = Q(296)/Q(T) for molecules 1..mol_max (the Fortran host calls it per
 * layer, modm.f90:250).  Lets the CPU legs of bench.py run without the product library. */
int orc_tips_2003(int64_t mol_max, double temp_lbl, double *scor);
/* test instrumentation of the last orc_modm call on this thread: [0] Voigt-branch evaluations, [1] speed-dependent SDVOIGT
 * calls, [2] CO2 lines on the Voigt branch, [3] of those XF=-1, [4] coupled non-CO2/O2 lines, [5] coupled O2 lines,
 * [6] lines with XG outside {0,-1,-3,-5}, [7] Voigt-branch evaluations with the negative-frequency resonance */
void orc_branch_counts(int64_t out[8]);
uint64_t orc_line_key(int64_t mol, int64_t rec);  /* splitmix64 of (mol<<32 | rec), 1-based */
const char *orc_last_error(void);

/* ---- cross sections ------------------------------------------------------------------------------
 * One spectral region of one cross-section molecule as MONORTM_XSEC_SUB holds it after its READs
 * (src/monortm_sub.F90:1656-1671): the FSCDXS range that decides whether the region is processed (:1647-1652), the
 * header values of the LAST temperature file read (V1x, V2x, NPTSx stay those of the last file, :1662) and the
 * per-temperature tables in ascending temperature order. */
typedef struct {
    int32_t ixmol;              /* 0-based index of the molecule in XAMNT's first dimension (loop 6000) */
    int32_t ntemp;              /* NTEMPF(ixsr,ixmol) <= 6 */
    int64_t npts;               /* NPTSx */
    double v1fx, v2fx;          /* V1FX, V2FX from FSCDXS */
    double v1x, v2x;            /* file header */
    double xdoplr;              /* XDOPLR(ixsr,ixmol), XSREAD :1385-1386 */
    double tx[6], pdx[6];       /* temperature, pressure in mb (TORR converted, :1665-1669) */
    const double *xsdat[6];     /* npts values each */
} orc_xs_region;

/* MONORTM_XSEC_SUB (src/monortm_sub.F90:1540-1749) with convolve (:1751-1834).  regs must be ordered by molecule, then
 * by spectral region (the loop order 6000 / 5000).  xamnt is (ld_xamnt, nlay) column-major; odxsec (nwn,nlay).
 * Returns 52 where the reference would run off xspd_int(0:10000000) (:1755, 1773-1774). */
int orc_xsec_sub(int64_t nwn, const double *wn, int64_t nlay, const double *p, const double *t,
                 int64_t nreg, const orc_xs_region *regs, int64_t ld_xamnt, const double *xamnt, double *odxsec);
/* convolve alone (pins against the executed reference text) */
int orc_convolve(const double *xspd, int64_t nxspd, double v1x, double v2x, double delvx, double pd, double hwdop,
                 double tave, double pave, const double *wn, int64_t nwn, double *xspave);

#ifdef __cplusplus
}
#endif
#endif
