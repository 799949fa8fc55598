/*
 * monortm_oracle.c -- TEST INFRASTRUCTURE ONLY (see monortm_oracle.h).
 *
 * Statement-for-statement CPU restatement of the monoRTM hot path in plain C.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Evaluation order, literal constants and integer
 * truncations follow the reference's linuxGNUdbl build: every REAL is binary64,
 * every d0 literal is binary128 (gfortran -fdefault-real-8 without
 * -fdefault-double-8), no FMA contraction, Fortran complex rules.
 *
 * Build: gcc -O0 -ffp-contract=off -fcx-fortran-rules (oracle/Makefile).
 *
 * PARITY PIN (see monortm_oracle.h): the reference cannot be compiled here (Fortran), so its own source TEXT is executed
 * through the mechanical translator tools/f90fn.py and this file is held to those outputs (tests/golden/ref_*.npz,
 * tests/test_ref_goldens.py): scalar routines to a few ulp, MODM + CALCTMR + RTM cases from 0.1 to 57900 cm-1 to 1e-13
 * (measured 0 .. 5e-16), MONORTM_XSEC_SUB + convolve bit-identical.  Further independent pins:
 * tests/test_oracle_units.py -- from-scratch numpy evaluations of the textbook expressions (isolated Lorentz line,
 * O2 line with first-order mixing, Voigt branch against scipy, layer-exact radiative transfer, the self continuum at
 * its table nodes), analytic identities and scipy.special.wofz for W4.
 *
 * Defined behaviour chosen where the reference has undefined behaviour:
 *  - modm.f90:845 indexes rho_molec(mol) for mol > 7 although the array has 7
 *    elements; we use the natural extension rhorat*wk(mol)/wtot.
 *  - modm.f90:339 reads XG(I,0) when J=1; treated as "not -5".
 */
#include "monortm_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../monortm_b200/csrc/tables/mtckd_tables.inc"
#include "../monortm_b200/csrc/tables/mtckd_ir_tables.inc"
#include "../monortm_b200/csrc/tables/smass_table.inc"

typedef double _Complex cplx;

/* ---- constants: PhysConstants.f90:19-39, PlanetEarth.f90:19-20 ---------- */
static const double PIref = 3.1415926535898;
static const double PLANCKref = 6.62606876E-27;
static const double BOLTZref = 1.3806503E-16;
static const double CLIGHTref = 2.99792458E+10;
static const double AVOGADref = 6.02214199E+23;
static const double RADCN1ref = 1.191042722E-12;
static const double RADCN2ref = 1.4387752;

/* /LAMCHN/ modm.f90:170-171 */
static const double ONEPL = 1.001;
static const double ONEMI = 0.999;

#define N_ABSRB 5050
#define NPTC_MAX 6000

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

#define IX(i, j) ((size_t)((i)-1) + (size_t)((j)-1) * ORC_MXMOL)
#define IB(i, k, j) ((size_t)((i)-1) + (size_t)((k)-1) * 7 + (size_t)((j)-1) * 49)

uint64_t orc_line_key(int64_t mol, int64_t rec)
{
    uint64_t z = ((uint64_t)mol << 32) | (uint64_t)rec;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* ======================================================================= */
/*  line store + TAPE3 reader  (lnfl_mod.f90, struct_types.f90, bufin_sgl)   */
/* ======================================================================= */
orc_lines *orc_lines_alloc(int64_t iim)
{
    orc_lines *ln = (orc_lines *)calloc(1, sizeof *ln);
    size_t n = (size_t)ORC_MXMOL * (size_t)iim, nb = 49u * (size_t)iim;
    ln->iim = iim;
    ln->iso = (int64_t *)calloc(n, sizeof(int64_t));
    ln->xnu0 = (double *)calloc(n, 8);
    ln->deltnu = (double *)calloc(n, 8);
    ln->e = (double *)calloc(n, 8);
    ln->alps = (double *)calloc(n, 8);
    ln->alpf = (double *)calloc(n, 8);
    ln->x = (double *)calloc(n, 8);
    ln->xg = (double *)calloc(n, 8);
    ln->s0 = (double *)calloc(n, 8);
    ln->rmol = (double *)calloc(n, 8);
    ln->sdep = (double *)calloc(n, 8);
    ln->brd_mol_flg = (int32_t *)calloc(nb, 4);
    ln->brd_mol_tmp = (double *)calloc(nb, 8);
    ln->brd_mol_hw = (double *)calloc(nb, 8);
    ln->brd_mol_shft = (double *)calloc(nb, 8);
    return ln;
}

void orc_lines_free(orc_lines *ln)
{
    if (!ln) return;
    free(ln->iso); free(ln->xnu0); free(ln->deltnu); free(ln->e); free(ln->alps);
    free(ln->alpf); free(ln->x); free(ln->xg); free(ln->s0); free(ln->rmol);
    free(ln->sdep); free(ln->brd_mol_flg); free(ln->brd_mol_tmp);
    free(ln->brd_mol_hw); free(ln->brd_mol_shft);
    free(ln);
}

/* one Fortran sequential-unformatted record with 4-byte markers
 * (-frecord-marker=4, build/makefile.common:198).  Returns payload length,
 * -1 on EOF, -2 on a malformed record. */
static long read_record(FILE *f, unsigned char *buf, size_t cap)
{
    int32_t n1, n2;
    if (fread(&n1, 4, 1, f) != 1) return -1;
    if (n1 < 0) return -2;
    size_t take = (size_t)n1 < cap ? (size_t)n1 : cap;
    if (fread(buf, 1, take, f) != take) return -2;
    if ((size_t)n1 > take && fseek(f, (long)((size_t)n1 - take), SEEK_CUR) != 0) return -2;
    if (fread(&n2, 4, 1, f) != 1 || n2 != n1) return -2;
    return (long)n1;
}

#define NLINEREC 250
/* INPUT_BLOCK byte offsets, struct_types.f90:33-43 */
#define OFF_VNU 0
#define OFF_SP 2000
#define OFF_ALFA 3000
#define OFF_EPP 4000
#define OFF_MOL 5000
#define OFF_HWHM 6000
#define OFF_TMPALF 7000
#define OFF_PSHIFT 8000
#define OFF_IFLG 9000
#define OFF_BRDFLG 10000
#define OFF_BRDDAT 17000
#define OFF_SDEP 38000
#define BLOCK_BYTES 39000

static float rd_f32(const unsigned char *b, size_t off) { float v; memcpy(&v, b + off, 4); return v; }
static int32_t rd_i32(const unsigned char *b, size_t off) { int32_t v; memcpy(&v, b + off, 4); return v; }
static double rd_f64(const unsigned char *b, size_t off) { double v; memcpy(&v, b + off, 8); return v; }

int orc_get_lnfl(const char *hfile, double v1, double v2, orc_lines *ln)
{
    FILE *f = fopen(hfile, "rb");
    if (!f) return fail(1, "ERROR OPENING HITRAN FILE");       /* lnfl_mod.f90:130-131 */
    static unsigned char hdr[1 << 16];
    static unsigned char blk[BLOCK_BYTES];
    memset(ln->nblm, 0, sizeof ln->nblm);                        /* :34 */

    /* PRLNHD, lnfl_mod.f90:211-331 */
    long n = read_record(f, hdr, sizeof hdr);
    if (n < 0) { fclose(f); return fail(2, "LAYER; TAPE3 DOES NOT EXIST"); }      /* :267 */
    if (n < 1664) { fclose(f); return fail(3, "TAPE3 header record too short"); }
    if (hdr[6 * 8 + 7] == '^') {                                 /* :258-262 */
        if (read_record(f, hdr + 2048, sizeof hdr - 2048) < 0) { fclose(f); return fail(3, "TAPE3 second header missing"); }
    }
    if (hdr[9 * 8 + 7] != 'I') {                                 /* :295-302 */
        fclose(f);
        return fail(4, " PRLNHD - NO ISOTOPE INFO ON LINFIL ");
    }

    int64_t mo_prev = 0;
    int ieof = 0;
    while (ieof == 0) {                                          /* :43 */
        /* RDLNFL, lnfl_mod.f90:136-209 */
        int64_t ilo = 1, ihi = 0;
        double vlo_adj = fmax(0.0, v1 - 25.0);                   /* :161 */
        int32_t nrec = 0, nwds = 0;
        double last_vnu = 0.0;
        for (;;) {
            unsigned char ph[24];
            long m = read_record(f, ph, sizeof ph);              /* BUFIN_sgl, 6 words */
            if (m < 0) { ieof = 1; break; }                      /* :162, :202-204 */
            if (m < 24) { fclose(f); return fail(5, "TAPE3 panel header too short"); }
            double vmax = rd_f64(ph, 8);
            nrec = rd_i32(ph, 16);
            nwds = rd_i32(ph, 20);
            if (vmax < vlo_adj) {                                /* :163-165 skip the data record */
                unsigned char dum[4];
                if (read_record(f, dum, sizeof dum) < 0) { ieof = 1; break; }
                continue;
            }
            memset(blk, 0, sizeof blk);
            long got = read_record(f, blk, sizeof blk);          /* :167 */
            if (got < 0) { ieof = 1; break; }
            (void)nwds;
            ihi = nrec;                                          /* :198 */
            break;
        }
        if (ieof) break;
        if (ihi < 0 || ihi > NLINEREC) { fclose(f); return fail(6, "TAPE3 block NREC out of range"); }

        for (int64_t ik = ilo; ik <= ihi; ik++) {                /* :45 */
            int64_t iflg = rd_i32(blk, OFF_IFLG + 4 * (ik - 1));
            int64_t mol = rd_i32(blk, OFF_MOL + 4 * (ik - 1));   /* int*4 -> int*8, :177 */
            int64_t mo;
            if (iflg >= 0 && iflg <= 100) {                      /* :47-48 */
                mo = mol % 100;
            } else if (iflg >= -3 && iflg <= -1) {               /* :49-50 */
                if (ik == 1) { fclose(f); return fail(7, "coupling record first in block"); }
                mo = (int64_t)rd_i32(blk, OFF_MOL + 4 * (ik - 2)) % 100;
            } else if (iflg == -5) {                             /* :51-60 */
                if (ik == 1) { fclose(f); return fail(7, "coupling record first in block"); }
                if (rd_i32(blk, OFF_IFLG + 4 * (ik - 2)) >= 0) {
                    mo = (int64_t)rd_i32(blk, OFF_MOL + 4 * (ik - 2)) % 100;
                    mo_prev = mo;
                } else {
                    mo = mo_prev;
                }
            } else {                                             /* :61-63 */
                fclose(f);
                return fail(8, "LC flag not recongnized. Must be 1, 3 or 5.");
            }
            if (mo < 1 || mo > ORC_MXMOL) { fclose(f); return fail(9, "molecule number out of range"); }
            ln->nblm[mo - 1] += 1;                               /* :65 */
            int64_t ii = ln->nblm[mo - 1];
            if (ii > ln->iim) { fclose(f); return fail(10, "line store IIM exceeded"); }
            ln->iso[IX(mo, ii)] = (mol % 1000) / 100;            /* :67 */
            ln->xnu0[IX(mo, ii)] = rd_f64(blk, OFF_VNU + 8 * (ik - 1));
            ln->s0[IX(mo, ii)] = rd_f32(blk, OFF_SP + 4 * (ik - 1));
            ln->alpf[IX(mo, ii)] = rd_f32(blk, OFF_ALFA + 4 * (ik - 1));
            ln->alps[IX(mo, ii)] = rd_f32(blk, OFF_HWHM + 4 * (ik - 1));
            ln->e[IX(mo, ii)] = rd_f32(blk, OFF_EPP + 4 * (ik - 1));
            ln->x[IX(mo, ii)] = rd_f32(blk, OFF_TMPALF + 4 * (ik - 1));
            ln->deltnu[IX(mo, ii)] = rd_f32(blk, OFF_PSHIFT + 4 * (ik - 1));
            if (iflg >= 0) ln->xg[IX(mo, ii)] = (double)(-1 * iflg);   /* :75-79 */
            else ln->xg[IX(mo, ii)] = (double)iflg;
            {   /* :80-82: int*4 bits -> real*4 -> real*8 */
                int32_t m4 = (int32_t)mol; float xm; memcpy(&xm, &m4, 4);
                ln->rmol[IX(mo, ii)] = (double)xm;
            }
            if (mo <= ORC_MXBRDMOL) {                            /* :84-90, RDLNFL :183-192 */
                for (int k = 1; k <= 7; k++) {
                    ln->brd_mol_flg[IB(mo, k, ii)] = rd_i32(blk, OFF_BRDFLG + 4 * ((k - 1) + 7 * (ik - 1)));
                    size_t d = OFF_BRDDAT + 4 * (size_t)(3 * (k - 1) + 21 * (ik - 1));
                    ln->brd_mol_hw[IB(mo, k, ii)] = rd_f32(blk, d);
                    ln->brd_mol_tmp[IB(mo, k, ii)] = rd_f32(blk, d + 4);
                    ln->brd_mol_shft[IB(mo, k, ii)] = rd_f32(blk, d + 8);
                }
            }
            ln->sdep[IX(mo, ii)] = rd_f32(blk, OFF_SDEP + 4 * (ik - 1));   /* :92 */

            if (mo == 7 && iflg >= 0) {                          /* :98-104 */
                double rvmr = 0.21;
                ln->alpf[IX(mo, ii)] = (ln->alpf[IX(mo, ii)] - rvmr * ln->alps[IX(mo, ii)]) / (1.0 - rvmr);
                if (ln->brd_mol_flg[IB(mo, mo, ii)] > 0)
                    ln->deltnu[IX(mo, ii)] = (ln->deltnu[IX(mo, ii)] - rvmr * ln->brd_mol_shft[IB(mo, mo, ii)]) / (1.0 - rvmr);
            }
            if (mo == 22 && iflg >= 0) {                         /* :105-113 */
                double rvmr = 0.79;
                ln->alpf[IX(mo, ii)] = (ln->alpf[IX(mo, ii)] - rvmr * ln->alps[IX(mo, ii)]) / (1.0 - rvmr);
            }
            last_vnu = rd_f64(blk, OFF_VNU + 8 * (ik - 1));
        }
        if (ihi >= 1 && last_vnu > (v2 + 25.)) ieof = 1;        /* :116 */
    }
    fclose(f);
    return 0;
}

/* ======================================================================= */
/*  lblrtm_sub.f90                                                          */
/* ======================================================================= */
/* XINT, lblrtm_sub.f90:1-34.  a and r3 are 1-based in the reference. */
static void xint(double v1a, double v2a, double dva, const double *a, double afact,
                 double vft, double dvr3, double *r3, int64_t n1r3, int64_t n2r3)
{
    double recdva = 1. / dva;
    int64_t ilo = (int64_t)((v1a + dva - vft) / dvr3 + 1. + ONEMI);
    if (ilo < n1r3) ilo = n1r3;
    int64_t ihi = (int64_t)((v2a - dva - vft) / dvr3 + ONEMI);
    if (ihi > n2r3) ihi = n2r3;
    for (int64_t i = ilo; i <= ihi; i++) {
        double vi = vft + dvr3 * (double)(i - 1);
        int64_t j = (int64_t)((vi - v1a) * recdva + ONEPL);
        double vj = v1a + dva * (double)(j - 1);
        double p = recdva * (vi - vj);
        double c = (3. - 2. * p) * p * p;
        double b = 0.5 * p * (1. - p);
        double b1 = b * (1. - p);
        double b2 = b * p;
        double conti = -a[j - 2] * b1 + a[j - 1] * (1. - c + b2) + a[j] * (c + b1) - a[j + 1] * b2;
        r3[i - 1] = r3[i - 1] + conti * afact;
    }
}

/* RADFN, lblrtm_sub.f90:36-97 */
double orc_radfn(double vi, double xkt)
{
    double xvi = vi, radfn;
    if (xkt > 0.0) {
        double xviokt = xvi / xkt;
        if (xviokt <= 0.01) {
            radfn = 0.5 * xviokt * xvi;
        } else if (xviokt <= 10.0) {
            double expvkt = exp(-xviokt);
            radfn = xvi * (1. - expvkt) / (1. + expvkt);
        } else {
            radfn = xvi;
        }
    } else {
        radfn = xvi;
    }
    return radfn;
}

/* ======================================================================= */
/*  contnm.f90 (microwave / far-IR subset: V2 < 820 cm-1)                    */
/* ======================================================================= */
/* pre_xint, contnm.f90:1146-1164 */
static void pre_xint(double v1ss, double v2ss, double v1abs, double dvabs, int64_t nptabs,
                     int64_t *ist, int64_t *last)
{
    int64_t nbnd_v1c = (int64_t)(2 + (v1ss - v1abs) / dvabs + 1.e-5);
    *ist = nbnd_v1c > 1 ? nbnd_v1c : 1;
    int64_t nbnd_v2c = (int64_t)(1 + (v2ss - v1abs) / dvabs + 1.e-5);
    *last = nptabs < nbnd_v2c ? nptabs : nbnd_v2c;
}

/* common index set-up of the table accessors SL296/SL260/FRN296/FRNCO2/xn2_r
 * (contnm.f90:1441-1456 and the identical blocks at 1950, 2457, 2981, 4192) */
static void accessor_grid(double v1abs, double v2abs, double v1s, double dvs, int64_t npts,
                          double *v1c, double *v2c, double *dvc, int64_t *nptc, int64_t *i1out)
{
    int64_t i1, i2;
    *dvc = dvs;
    *v1c = v1abs - *dvc;
    *v2c = v2abs + *dvc;
    if (*v1c < v1s) i1 = -1;
    else i1 = (int64_t)((*v1c - v1s) / dvs + 0.01);
    *v1c = v1s + dvs * (double)(i1 - 1);
    i2 = (int64_t)((*v2c - v1s) / dvs + 0.01);
    *nptc = i2 - i1 + 3;
    if (*nptc > npts) *nptc = npts + 4;
    *v2c = *v1c + dvs * (double)(*nptc - 1);
    *i1out = i1;
}

/* the same set-up with the two details some accessors change: the offset added before truncation (O2FUV uses 1.e-5,
 * contnm.f90:9968-9973) and the NPTS cap (absent in O2HERZ, :9833-9836) */
static void accessor_grid2(double v1abs, double v2abs, double v1s, double dvs, int64_t npts, double eps, int cap,
                           double *v1c, double *v2c, double *dvc, int64_t *nptc, int64_t *i1out)
{
    int64_t i1, i2;
    *dvc = dvs;
    *v1c = v1abs - *dvc;
    *v2c = v2abs + *dvc;
    if (*v1c < v1s) i1 = -1;
    else i1 = (int64_t)((*v1c - v1s) / dvs + eps);
    *v1c = v1s + dvs * (double)(i1 - 1);
    i2 = (int64_t)((*v2c - v1s) / dvs + eps);
    *nptc = i2 - i1 + 3;
    if (cap && *nptc > npts) *nptc = npts + 4;
    *v2c = *v1c + dvs * (double)(*nptc - 1);
    *i1out = i1;
}

/* HERTDA + HERPRS, contnm.f90:9856-9948 */
static double herzberg(double v, double t, double p)
{
    double herz = 0.0;
    if (!(v <= 36000.00)) {
        double corr = 0.;
        if (v <= 40000.) corr = ((40000. - v) / 4000.) * 7.917E-07;
        double yratio = v / 48811.0;
        double lg = log(yratio);
        herz = 6.884E-04 * (yratio) * exp(-69.738 * (lg * lg)) - corr;
    }
    const double po = 1013., to = 273.16;
    herz = herz * (1. + .83 * (p / po) * (to / t));
    return herz;
}

/* XFAC_RHU(-1:61), contnm.f90:186-202 */
static double xfac_rhu(int64_t i) { return MTCKD_XFAC_RHU[i + 1]; }

int orc_contnm_one(int64_t im, const double cntnm[7], double pave, double tave,
                   const double *wk, double wbroad, int64_t nmol, double v1, double v2,
                   double v1abs, double v2abs, int64_t nptabs, double *absrb)
{
    static double c[NPTC_MAX], c0[N_ABSRB], c1[N_ABSRB], c2[N_ABSRB], cself[N_ABSRB];
    /* oneMolecCntnm, CntnmFactors.f90:95-139 */
    double xself = 0., xfrgn = 0., xco2c = 0., xo3cn = 0., xo2cn = 0., xn2cn = 0., xrayl = 0.;
    switch (im) {
    case 1: xself = cntnm[0]; xfrgn = cntnm[1]; break;
    case 2: xco2c = cntnm[2]; break;
    case 3: xo3cn = cntnm[3]; break;
    case 7: xo2cn = cntnm[4]; break;
    case 22: xn2cn = cntnm[5]; break;
    case 99: xrayl = cntnm[6]; break;
    default: break;
    }
    if (nptabs > N_ABSRB - 2) return fail(21, "NPTABS too large");

    const double dvabs = 1.0;
    const double P0 = 1013., T0 = 296., XLOSMT = 2.68675E+19;   /* contnm.f90:86-87 */
    double rhoave = (pave / P0) * (T0 / tave);                    /* :222 */
    double amagat = (pave / P0) * (273. / tave);                  /* :229 */
    double wtot = wbroad;                                         /* :231-234 */
    for (int64_t m = 1; m <= nmol; m++) wtot = wtot + wk[m - 1];
    double x_vmr_h2o = wk[0] / wtot;                              /* :236-240 */
    double x_vmr_o2 = wk[6] / wtot;
    double x_vmr_n2 = 1. - x_vmr_h2o - x_vmr_o2;
    double wn2 = x_vmr_n2 * wtot;

    double h2o_fac = wk[0] / wtot;                                /* :300-302 */
    double rself = h2o_fac * rhoave * 1.e-20 * xself;
    double rfrgn = (1. - h2o_fac) * rhoave * 1.e-20 * xfrgn;

    double v1c, v2c, dvc;
    int64_t nptc, i1, ist, last;

    /* ---- H2O self, contnm.f90:325-371 ---- */
    if ((v2 > -20.0) && (v1 < 20000.) && xself > 0.) {
        double *sh2ot0 = c0, *sh2ot1 = c1;
        memset(c0, 0, sizeof c0);
        memset(c1, 0, sizeof c1);
        accessor_grid(v1abs, v2abs, -20.0, 10.0, 2003, &v1c, &v2c, &dvc, &nptc, &i1);   /* SL296 :1432-1469 */
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            sh2ot0[j - 1] = 0.;
            if (i >= 1 && i <= 2003) sh2ot0[j - 1] = MTCKD_SH2O_296[i - 1];
        }
        accessor_grid(v1abs, v2abs, -20.0, 10.0, 2003, &v1c, &v2c, &dvc, &nptc, &i1);   /* SL260 :1940-1977 */
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            sh2ot1[j - 1] = 0.;
            if (i >= 1 && i <= 2003) sh2ot1[j - 1] = MTCKD_SH2O_260[i - 1];
        }
        double tfac = (tave - T0) / (260. - T0);                  /* :334 */
        for (int64_t j = 1; j <= nptc; j++) {                     /* :339-363 */
            double sh2o = 0.;
            if (sh2ot0[j - 1] > 0.) sh2o = sh2ot0[j - 1] * pow(sh2ot1[j - 1] / sh2ot0[j - 1], tfac);
            cself[j - 1] = wk[0] * (sh2o * rself);
        }
        pre_xint(-20.0, 20000.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, cself, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- H2O foreign, contnm.f90:380-474 ---- */
    if ((v2 > -20.0) && (v1 < 20000.) && xfrgn > 0.) {
        double *fh2o = c2;
        memset(c2, 0, sizeof c2);
        double f0 = 0.06, v0f1 = 255.67, hwsq1 = 240. * 240., beta1 = 57.83, c_1 = -0.42;
        double c_2 = 0.3, beta2 = 630.;
        accessor_grid(v1abs, v2abs, -20.0, 10.0, 2003, &v1c, &v2c, &dvc, &nptc, &i1);   /* FRN296 :2448-2485 */
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            fh2o[j - 1] = 0.;
            if (i >= 1 && i <= 2003) fh2o[j - 1] = MTCKD_FH2O[i - 1];
        }
        for (int64_t j = 1; j <= nptc; j++) {                     /* :417-452 */
            double vj = v1c + dvc * (double)(j - 1);
            double fscal;
            if (vj <= 600.) {
                int64_t jfac = (int64_t)((vj + 10.) / 10. + 0.00001);
                fscal = xfac_rhu(jfac);
            } else {
                double vdelsq1 = (vj - v0f1) * (vj - v0f1);
                double vdelmsq1 = (vj + v0f1) * (vj + v0f1);
                double t1 = (vj - v0f1) / beta1, t2 = (vj + v0f1) / beta1, t3 = vj / beta2;
                double vf1 = t1 * t1; vf1 = vf1 * vf1; vf1 = vf1 * vf1;   /* **8 by squaring (pow_r8_i8) */
                double vmf1 = t2 * t2; vmf1 = vmf1 * vmf1; vmf1 = vmf1 * vmf1;
                double vf2 = t3 * t3; vf2 = vf2 * vf2; vf2 = vf2 * vf2;
                fscal = 1. + (f0 + c_1 * ((hwsq1 / (vdelsq1 + hwsq1 + vf1)) + (hwsq1 / (vdelmsq1 + hwsq1 + vmf1)))) /
                                 (1. + c_2 * vf2);
            }
            fh2o[j - 1] = fh2o[j - 1] * fscal;
            double c_f = wk[0] * fh2o[j - 1];
            c[j - 1] = c_f * rfrgn;
        }
        pre_xint(-20.0, 20000.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- CO2, contnm.f90:484-528, FRNCO2 :2958-3014 ---- */
    if ((v2 > -20.0) && (v1 < 10000.) && xco2c > 0) {
        double *fco2 = c0;
        memset(c0, 0, sizeof c0);
        double wco2 = wk[1] * rhoave * 1.0E-20 * xco2c;
        double trat = tave / 246.;
        accessor_grid(v1abs, v2abs, -4.0, 2.0, 5003, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            fco2[j - 1] = 0.;
            if (i >= 1 && i <= 5003) {
                double tcor = 1.;
                if (i >= 1196 && i <= 1220) tcor = pow(trat, MTCKD_CO2_TDEP_BANDHEAD[i - 1196]);
                fco2[j - 1] = tcor * MTCKD_FCO2[i - 1];
            }
        }
        for (int64_t j = 1; j <= nptc; j++) {
            double vj = v1c + dvc * (double)(j - 1);
            double cfac = 1.;
            if (vj >= 2000. && vj <= 2998.) {                        /* :510-513 */
                int64_t jfac = (int64_t)((vj - 1998.) / 2. + 0.00001);
                cfac = MTCKD_XFACCO2[jfac - 1];
            }
            fco2[j - 1] = cfac * fco2[j - 1];
            c[j - 1] = fco2[j - 1] * wco2;
        }
        pre_xint(-4.0, 10000.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- N2 collision-induced pure rotation, contnm.f90:906-943, xn2_r :4160-4230 ---- */
    if ((v2 > -10.0) && (v1 < 350.) && xn2cn > 0.) {
        memset(c0, 0, sizeof c0);
        memset(c1, 0, sizeof c1);
        double a_h2o = 1.;
        double tau_fac = xn2cn * (wn2 / XLOSMT) * amagat;
        double xo2 = 0.21, xn2 = 0.79, t_296 = 296., t_220 = 220.;
        double tfac = (tave - t_296) / (t_220 - t_296);
        accessor_grid(v1abs, v2abs, -10., 5.0, 73, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            c0[j - 1] = 0.;
            if (i < 1 || i > 73) continue;
            c0[j - 1] = MTCKD_N2RT_296[i - 1] * pow(MTCKD_N2RT_220[i - 1] / MTCKD_N2RT_296[i - 1], tfac);
            double sf_t = MTCKD_N2RT_296_SF[i - 1] * pow(MTCKD_N2RT_220_SF[i - 1] / MTCKD_N2RT_296_SF[i - 1], tfac);
            c1[j - 1] = (sf_t - 1.) * (xn2) / (xo2);
        }
        for (int64_t j = 1; j <= nptc; j++)
            c[j - 1] = tau_fac * c0[j - 1] * (x_vmr_n2 + c1[j - 1] * x_vmr_o2 + a_h2o * x_vmr_h2o);
        pre_xint(-10., 350., v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- N2 collision-induced fundamental, contnm.f90:963-1013, n2_ver_1 :4331-4417 ---- */
    if ((v2 > 2001.77) && (v1 < 2897.59) && xn2cn > 0.) {
        double *cn0 = c0, *cn1 = c1, *cn2 = c2;
        memset(c0, 0, sizeof c0);
        memset(c1, 0, sizeof c1);
        memset(c2, 0, sizeof c2);
        double tau_fac = xn2cn * (wn2 / XLOSMT) * amagat;
        const double t_272 = 272., t_228 = 228.;
        double xtfac = ((1. / tave) - (1. / t_272)) / ((1. / t_228) - (1. / t_272));
        double xt_lin = (tave - t_272) / (t_228 - t_272);
        double a_o2 = 1.294 - 0.4545 * tave / 296.;
        accessor_grid(v1abs, v2abs, 1997.784896, 3.981461525, 228, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            cn0[j - 1] = 0.;
            if (i < 1 || i > 228) continue;
            double vj = v1c + dvc * (double)(j - 1);
            double a = MTCKD_N2F_272[i - 1], b = MTCKD_N2F_228[i - 1];
            if ((a > 0.) && (b > 0.)) cn0[j - 1] = a * pow(b / a, xtfac);
            else cn0[j - 1] = a + (b - a) * xt_lin;
            cn0[j - 1] = cn0[j - 1] / vj;
            cn1[j - 1] = a_o2 * cn0[j - 1];
            cn2[j - 1] = (9. / 7.) * MTCKD_N2F_AH2O[i - 1] * cn0[j - 1];
        }
        for (int64_t j = 1; j <= nptc; j++)
            c[j - 1] = tau_fac * (x_vmr_n2 * cn0[j - 1] + x_vmr_o2 * cn1[j - 1] + x_vmr_h2o * cn2[j - 1]);
        pre_xint(1997.784896, 2901.576661, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- N2 collision-induced first overtone, contnm.f90:1025-1068, n2_overtone1 :4579-4631 ---- */
    if ((v2 > 4340.0) && (v1 < 4910.) && xn2cn > 0.) {
        memset(c0, 0, sizeof c0);
        double a_o2 = 1., a_h2o = 1.;
        double tau_fac = xn2cn * (wn2 / XLOSMT) * amagat * (x_vmr_n2 + a_o2 * x_vmr_o2 + a_h2o * x_vmr_h2o);
        accessor_grid(v1abs, v2abs, 4340.0, 3.0, 191, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            c0[j - 1] = 0.;
            if (i < 1 || i > 191) continue;
            double vj = v1c + dvc * (double)(j - 1);
            c0[j - 1] = MTCKD_N2F1[i - 1] / vj;
        }
        for (int64_t j = 1; j <= nptc; j++) c[j - 1] = tau_fac * c0[j - 1];
        pre_xint(4340.0, 4910.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O3 Chappuis / Wulf, contnm.f90:536-553, XO3CHP :4685-4732 ---- */
    if (v2 > 8920.0 && v1 <= 24665.0 && xo3cn > 0.) {
        double wo3 = wk[2] * 1.0E-20 * xo3cn;
        double dt = tave - 273.15;
        accessor_grid(v1abs, v2abs, 8920.0, 5.0, 3150, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double cch0 = 0., cch1 = 0., cch2 = 0.;
            if (!(i < 1 || i > 3150)) {
                double vj = v1c + dvc * (double)(j - 1);
                cch0 = MTCKD_O3CH_X[i - 1] / vj;
                cch1 = MTCKD_O3CH_Y[i - 1] / vj;
                cch2 = MTCKD_O3CH_Z[i - 1] / vj;
            }
            c[j - 1] = (cch0 + (cch1 + cch2 * dt) * dt) * wo3;
        }
        pre_xint(8920.0, 24665.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O3 Hartley-Huggins, contnm.f90:555-599, O3HHT0/1/2 :6850-8216 (the three tables share one grid) ---- */
    if (v2 > 27370. && v1 < 40800. && xo3cn > 0.) {
        static double absbsv[N_ABSRB];
        double wo3 = wk[2] * 1.E-20 * xo3cn;
        double tc = tave - 273.15;
        accessor_grid(v1abs, v2abs, 27370., 5.0, 2687, &v1c, &v2c, &dvc, &nptc, &i1);
        double vj = 0.;
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double s0v = 0., ct1 = 0., ct2 = 0.;
            vj = v1c + dvc * (double)(j - 1);
            if (!(i < 1 || i > 2687)) {
                s0v = MTCKD_O3HH0[i - 1] / vj;
                ct1 = MTCKD_O3HH1[i - 1];
                ct2 = MTCKD_O3HH2[i - 1];
            }
            c[j - 1] = s0v * wo3;
            c[j - 1] = c[j - 1] * (1. + ct1 * tc + ct2 * tc * tc);
        }
        int64_t i_fix = 0;
        int fix = (vj > 40815.) && (v2 > 40800);                  /* :573-578: VJ is the last grid point of the loop */
        if (fix) {
            i_fix = (int64_t)((40800. - v1abs) / dvabs + 1.001);
            for (int64_t i = i_fix; i <= nptabs; i++) absbsv[i - 1] = absrb[i - 1];
        }
        pre_xint(27370., 40800., v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
        if (fix)
            for (int64_t i = i_fix; i <= nptabs; i++) absrb[i - 1] = absbsv[i - 1];
    }

    /* ---- O3 UV Hartley-Huggins, contnm.f90:603-642, O3HHUV :8826-8866 ---- */
    if (v2 > 40800. && v1 < 54000. && xo3cn > 0.) {
        static double absbsv[N_ABSRB];
        double wo3 = wk[2] * xo3cn;
        accessor_grid(v1abs, v2abs, 40800., 100., 133, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 133)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = MTCKD_O3HUV[i - 1] / vj;
            }
            c[j - 1] = c0v * wo3;
        }
        int64_t i_fix = 0;
        if (v1 < 40800) {
            i_fix = (int64_t)((40800. - v1abs) / dvabs + 1.001);
            for (int64_t i = 1; i <= i_fix - 1; i++) absbsv[i - 1] = absrb[i - 1];
        }
        pre_xint(40800., 54000., v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
        if (v1 < 40800)
            for (int64_t i = 1; i <= i_fix - 1; i++) absrb[i - 1] = absbsv[i - 1];
    }

    /* ---- O2 collision-induced fundamental, contnm.f90:657-693, o2_ver_1 :8917-8981 ---- */
    if ((v2 > 1340.0) && (v1 < 1850.) && xo2cn > 0.) {
        double tau_fac = xo2cn * wk[6] * 1.e-20 * amagat;
        double xktfac = (1. / 296.) - (1. / tave);
        double factor = (1.e+20 / 2.68675e+19);
        accessor_grid(v1abs, v2abs, 1340.0, 5.0, 103, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 103)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = factor * MTCKD_O2F[i - 1] * exp(MTCKD_O2F_T[i - 1] * xktfac) / vj;
            }
            c[j - 1] = tau_fac * c0v;
        }
        pre_xint(1340.0, 1850.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 1.27 um (Mate et al.), contnm.f90:709-734, O2INF1 :9047-9100 ---- */
    if ((v2 > 7536.0) && (v1 < 8500.) && xo2cn > 0.) {
        double a_o2 = 1. / 0.446, a_n2 = 0.3 / 0.446, a_h2o = 1.;
        double tau_fac = xo2cn * (wk[6] / XLOSMT) * amagat * (a_o2 * x_vmr_o2 + a_n2 * x_vmr_n2 + a_h2o * x_vmr_h2o);
        accessor_grid(v1abs, v2abs, 7536.0, 2.0, 483, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 483)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = MTCKD_O2INF1[i - 1] / vj;
            }
            c[j - 1] = tau_fac * c0v;
        }
        pre_xint(7536.0, 8500.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 9100-11000 cm-1 (Mlawer et al.), contnm.f90:745-766, O2INF2 :9227-9279 ---- */
    if ((v2 > 9100.0) && (v1 < 11000.) && xo2cn > 0.) {
        const double v1_osc = 9375., hw1 = 58.96, v2_osc = 9439., hw2 = 45.04, s1 = 1.166E-04, s2 = 3.086E-05;
        const double v1s = 9100., v2s = 11000., dvs = 2.;
        double wo2 = xo2cn * (wk[6] * 1.e-20) * rhoave;
        double adjwo2 = (wk[6] / wtot) * (1. / 0.209) * wo2;
        dvc = dvs;
        v1c = v1abs - dvc;
        v2c = v2abs + dvc;
        if (v1c < v1s) v1c = v1s - 2. * dvs;
        if (v2c > v2s) v2c = v2s + 2. * dvs;
        nptc = (int64_t)((v2c - v1c) / dvc + 3.01);
        v2c = v1c + dvc * (double)(nptc - 1);
        for (int64_t j = 1; j <= nptc; j++) {
            double c0v = 0.;
            double vj = v1c + dvc * (double)(j - 1);
            if ((vj > v1s) && (vj < v2s)) {
                double dv1 = vj - v1_osc, dv2 = vj - v2_osc, damp1, damp2;
                if (dv1 < 0.0) damp1 = exp(dv1 / 176.1); else damp1 = 1.0;
                if (dv2 < 0.0) damp2 = exp(dv2 / 176.1); else damp2 = 1.0;
                double q1 = dv1 / hw1, q2 = dv2 / hw2;
                double o2inf = 0.31831 * (((s1 * damp1 / hw1) / (1. + q1 * q1)) + ((s2 * damp2 / hw2) / (1. + q2 * q2))) * 1.054;
                c0v = o2inf / vj;
            }
            c[j - 1] = c0v * adjwo2;
        }
        pre_xint(v1s, v2s, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 A band, contnm.f90:773-792, O2INF3 :9282-9330 ---- */
    if ((v2 > 12961.5) && (v1 < 13221.5) && xo2cn > 0.) {
        double tau_fac = xo2cn * (wk[6] / XLOSMT) * amagat;
        accessor_grid(v1abs, v2abs, 12961.5, 1.0, 261, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 261)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = MTCKD_O2INF3[i - 1] / vj;
            }
            c[j - 1] = tau_fac * c0v;
        }
        pre_xint(12961.5, 13221.5, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 visible (Greenblatt et al.), contnm.f90:807-830, O2_vis :9400-9457 ---- */
    if ((v2 > 15000.0) && (v1 < 29870.) && xo2cn > 0.) {
        double wo2 = wk[6] * 1.e-20 * ((pave / 1013.) * (273. / tave)) * xo2cn;
        double chio2 = wk[6] / wtot;
        double adjwo2 = chio2 * wo2;
        double q = 55. * 273. / 296.;
        double factor = 1. / ((XLOSMT * 1.e-20 * (q * q)) * 89.5);
        accessor_grid(v1abs, v2abs, 15140.0, 10.0, 1474, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 1474)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = factor * MTCKD_O2VIS[i - 1] / vj;
            }
            c[j - 1] = c0v * adjwo2;
        }
        pre_xint(15140.0, 29870.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 Herzberg continuum, contnm.f90:834-850, O2HERZ :9808-9852 ---- */
    if (v2 > 36000.0 && xo2cn > 0.) {
        double wo2 = wk[6] * 1.e-20 * xo2cn;
        accessor_grid2(v1abs, v2abs, 36000., 10., 0, 0.01, 0, &v1c, &v2c, &dvc, &nptc, &i1);
        if (nptc > NPTC_MAX) return fail(23, "O2HERZ: NPTC exceeds C(6000)");
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = herzberg(vj, tave, pave) / vj;
            }
            c[j - 1] = c0v * wo2;
        }
        pre_xint(36000., 99999., v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- O2 far UV (Schumann-Runge continuum), contnm.f90:857-875, O2FUV :9952-9994 ---- */
    if (v2 > 56740.0 && xo2cn > 0.) {
        double wo2 = wk[6] * 1.e-20 * xo2cn;
        accessor_grid2(v1abs, v2abs, 56740.0, 20.0, 1512, 1.e-5, 1, &v1c, &v2c, &dvc, &nptc, &i1);
        for (int64_t j = 1; j <= nptc; j++) {
            int64_t i = i1 + (j - 1);
            double c0v = 0.;
            if (!(i < 1 || i > 1512)) {
                double vj = v1c + dvc * (double)(j - 1);
                c0v = MTCKD_O2FUV[i - 1] / vj;
            }
            c[j - 1] = c0v * wo2;
        }
        pre_xint(56740.0, 86960.0, v1abs, dvabs, nptabs, &ist, &last);
        xint(v1c, v2c, dvc, c, 1.0, v1abs, dvabs, absrb, ist, last);
    }

    /* ---- Rayleigh extinction, contnm.f90:1107-1131 (IAERSL = 0 in monoRTM; JRAD = 0: the radiation term MODM applies
     * later, modm.f90:243, is divided out here) ---- */
    if (v2 >= 820. && xrayl > 0.) {
        double conv_cm2mol = xrayl * 1.E-20 / (2.68675e-1 * 1.e5);
        double xkt = tave / RADCN2ref;
        for (int64_t i = 1; i <= nptabs; i++) {
            double vrayleigh = v1abs + (double)(i - 1) * dvabs;
            double xvrayleigh = vrayleigh / 1.e4;
            double ray_ext = ((xvrayleigh * xvrayleigh * xvrayleigh) / (9.38076E2 - 10.8426 * (xvrayleigh * xvrayleigh))) *
                             (wtot * conv_cm2mol);
            ray_ext = ray_ext * xvrayleigh / orc_radfn(vrayleigh, xkt);
            absrb[i - 1] = absrb[i - 1] + ray_ext;
        }
    }
    return 0;
}

/* ======================================================================= */
/*  CloudOptProp.f90                                                        */
/* ======================================================================= */
/* Forward_TKC, CloudOptProp.f90:79-157.  d0 literals are binary128 in the
 * parity build. */
/* real(16)**real(16) with exponent 2.: libquadmath powq(x,2) is within 1 binary128 ulp of
 * the correctly rounded x*x used here; invisible after rounding to binary64. */
static __float128 powq2(__float128 x) { return x * x; }

static double forward_tkc(double freq, double temp)
{
    const double Hz_per_GHz = 1.e9, cm_per_m = 100.;
    const double a_1 = 8.110808E+01, b_1 = 4.433736E-03, c_1 = 1.301700E-13, d_1 = 6.627126E+02;
    const double a_2 = 2.025164E+00, b_2 = 1.072976E-02, c_2 = 1.011945E-14, d_2 = 6.089168E+02;
    const double t_c = 1.342433E+02;
    double frq = freq * Hz_per_GHz;
    double pi = PIref, cl = CLIGHTref / cm_per_m;
    double eps_s = (double)(87.9144Q - 0.404399Q * (__float128)temp + 9.58726E-4Q * (__float128)pow(temp, 2.) -
                            1.32802E-6Q * (__float128)pow(temp, 3.));
    double delta_1 = a_1 * exp(-b_1 * temp);
    double tau_1 = c_1 * exp(d_1 / (temp + t_c));
    double delta_2 = a_2 * exp(-b_2 * temp);
    double tau_2 = c_2 * exp(d_2 / (temp + t_c));
    double term1_p1 = (double)((__float128)(pow(tau_1, 2.) * delta_1) /
                               (1.Q + powq2(2.Q * (__float128)pi * (__float128)frq * (__float128)tau_1)));
    double term2_p1 = (double)((__float128)(pow(tau_2, 2.) * delta_2) /
                               (1.Q + powq2(2.Q * (__float128)pi * (__float128)frq * (__float128)tau_2)));
    double eps1 = (double)((__float128)eps_s -
                           (powq2(2.Q * (__float128)pi * (__float128)frq)) * (__float128)(term1_p1 + term2_p1));
    term1_p1 = (double)((__float128)(tau_1 * delta_1) /
                        (1.Q + powq2(2.Q * (__float128)pi * (__float128)frq * (__float128)tau_1)));
    term2_p1 = (double)((__float128)(tau_2 * delta_2) /
                        (1.Q + powq2(2.Q * (__float128)pi * (__float128)frq * (__float128)tau_2)));
    double eps2 = (double)(2.Q * (__float128)pi * (__float128)frq * (__float128)(term1_p1 + term2_p1));
    cplx epsl = CMPLX(eps1, eps2);
    cplx re = (epsl - 1.) / (epsl + 2.);
    double alpha = (double)(6.Q * (__float128)pi * (__float128)cimag(re) * (__float128)frq * 1.E-3Q / (__float128)cl);
    return alpha;
}

/* ODCLW_TKC, CloudOptProp.f90:29-53 */
double orc_odclw(double wn, double temp, double clw)
{
    const double K_at_0C = 273.15, Hz_per_GHz = 1.e9;
    double freq = wn * CLIGHTref / Hz_per_GHz;
    double tempc = temp - K_at_0C;
    double absclw = forward_tkc(freq, tempc);
    return absclw * clw;
}

/* ======================================================================= */
/*  modm.f90: line shapes                                                   */
/* ======================================================================= */
/* W4, modm.f90:1100-1130 */
static cplx w4(double x, double y)
{
    cplx t = CMPLX(y, -x), u;
    double s = fabs(x) + y;
    if (!(s < 15.)) return t * .5641896 / (.5 + t * t);                        /* region I */
    if (!(s < 5.5)) {                                                          /* region II */
        u = t * t;
        return t * (1.410474 + u * .5641896) / (.75 + u * (3. + u));
    }
    if (!(y < 0.195 * fabs(x) - 0.176))                                        /* region III */
        return (16.4955 + t * (20.20933 + t * (11.96482 + t * (3.778987 + t * .5642236)))) /
               (16.4955 + t * (38.82363 + t * (39.27121 + t * (21.69274 + t * (6.699398 + t)))));
    u = t * t;                                                                 /* region IV */
    return cexp(u) - t * (36183.31 - u * (3321.9905 - u * (1540.787 - u * (219.0313 - u * (35.76683 - u * (1.320522 - u * .56419)))))) /
                         (32066.6 - u * (24322.84 - u * (9022.228 - u * (2186.181 - u * (364.2191 - u * (61.57037 - u * (1.841439 - u)))))));
}

void orc_w4(double x, double y, double *re, double *im)
{
    cplx w = w4(x, y);
    *re = creal(w);
    *im = cimag(w);
}

static cplx hum3(cplx t)
{
    return (16.4955 + t * (20.20933 + t * (11.96482 + t * (3.778987 + t * .5642236)))) /
           (16.4955 + t * (38.82363 + t * (39.27121 + t * (21.69274 + t * (6.699398 + t)))));
}
static cplx hum4(cplx t, cplx u)
{
    return cexp(u) - t * (36183.31 - u * (3321.9905 - u * (1540.787 - u * (219.0313 - u * (35.76683 - u * (1.320522 - u * .56419)))))) /
                         (32066.6 - u * (24322.84 - u * (9022.228 - u * (2186.181 - u * (364.2191 - u * (61.57037 - u * (1.841439 - u)))))));
}

/* SD_Humlicek, modm.f90:1150-1251 */
static cplx sd_humlicek(double x1, double y1, double x2, double y2)
{
    cplx t1 = CMPLX(y1, -x1), t2 = CMPLX(y2, -x2), u1, u2, w1, w2;
    double s1 = fabs(x1) + y1, s2 = fabs(x2) + y2;
    int region1, region2, region;
    if (s1 >= 15.0) region1 = 1;
    else if (s1 >= 6.0 && s1 < 15.0) region1 = 2;
    else { region1 = 3; if (y1 < 0.195 * fabs(x1) - 0.176) region1 = 4; }
    if (s2 >= 15.0) region2 = 1;
    else if (s2 >= 6.0 && s2 < 15.0) region2 = 2;
    else { region2 = 3; if (y2 < 0.195 * fabs(x2) - 0.176) region2 = 4; }
    region = region1 > region2 ? region1 : region2;
    if (!(region > 1)) {
        w1 = t1 * .5641896 / (.5 + t1 * t1);
        w2 = t2 * .5641896 / (.5 + t2 * t2);
        return w1 - w2;
    }
    if (!(region > 2)) {
        u1 = t1 * t1;
        u2 = t2 * t2;
        w1 = t1 * (1.410474 + u1 * .5641896) / (.75 + u1 * (3. + u1));
        w2 = t2 * (1.410474 + u2 * .5641896) / (.75 + u2 * (3. + u2));
        return w1 - w2;
    }
    if (!(region > 3)) {
        w1 = hum3(t1);
        w2 = hum3(t2);
        return w1 - w2;
    }
    u1 = t1 * t1;
    u2 = t2 * t2;
    if (region1 == 4) w1 = hum4(t1, u1); else w1 = hum3(t1);
    if (region2 == 4) w2 = hum4(t2, u2); else w2 = hum3(t2);
    return w1 - w2;
}

void orc_sd_humlicek(double x1, double y1, double x2, double y2, double *re, double *im)
{
    cplx w = sd_humlicek(x1, y1, x2, y2);
    *re = creal(w);
    *im = cimag(w);
}

static __thread int g_sdv_stop;   /* set when modm.f90:1062 would STOP */
static __thread int64_t g_nvoigt;
/* test instrumentation: which Voigt-branch cases a run reached.  [0] Voigt-branch evaluations (= g_nvoigt), [1] SDVOIGT calls
 * that took the speed-dependent branch (:1022-1066), [2] CO2 lines, [3] CO2 lines with XF=-1, [4] coupled lines of molecules
 * other than CO2/O2 (:588-615), [5] coupled O2 lines, [6] lines with an XG outside {0,-1,-3,-5}, [7] negative-frequency resonance
 * evaluated on the Voigt branch (DIFF <= 0) */
static __thread int64_t g_branch[8];
void orc_branch_counts(int64_t out[8]) { for (int i = 0; i < 8; i++) out[i] = g_branch[i]; }

/* SDVOIGT, modm.f90:965-1087 (the AVC interpolation :1004-1009 feeds nothing) */
static double sdvoigt(double deltnu, double alphal, double alphad, double sdep)
{
    double pi = PIref, tiny = 1.0e-4;
    double zeta = alphal / (alphal + alphad);
    double al = 0., dnu = 0., anorm1;
    cplx v;
    if (zeta < 1.00) {
        al = alphal / alphad;
        dnu = deltnu / alphad;
    }
    if (zeta == 1.00 && fabs(sdep) < tiny)
        return (alphal / (pi * (alphal * alphal + (deltnu) * (deltnu))));
    if (fabs(sdep) > tiny) {
        g_branch[1]++;
        double gamma2 = alphal * sdep;
        double alfa = (alphal / gamma2) - 1.5;
        double beta = (deltnu / gamma2);
        double delta = (1.0 / 4.0 / log(2.)) * (alphad * alphad / gamma2 / gamma2);
        double alfadelta = alfa + delta;
        double temp = sqrt(alfadelta * alfadelta + beta * beta);
        double x1 = (1.0 / sqrt(2.0)) * sqrt(temp + alfadelta) - sqrt(delta);
        double x2 = x1 + 2.0 * sqrt(delta);
        double sign;
        if (beta > 0.0) sign = 1; else if (beta == 0.0) sign = 0; else sign = -1;
        double y1 = sign * sqrt((temp - delta - alfa) / 2.0);
        double y2 = y1;
        v = sd_humlicek(y1, x1, y2, x2);
        if (creal(v) < 0.0) g_sdv_stop = 1;                                     /* :1062 STOP */
        anorm1 = sqrt(log(2.) / pi) / alphad;
    } else {
        double x = sqrt(log(2.)) * (dnu);
        double y = 1000.;
        if (zeta < 1.000) y = sqrt(log(2.)) * al;
        v = w4(x, y);
        anorm1 = sqrt(log(2.) / pi) / alphad;
    }
    anorm1 = sqrt(log(2.) / pi) / alphad;
    return creal(v) * anorm1;
}

double orc_sdvoigt(double deltnu, double alphal, double alphad, double sdep, int *err)
{
    g_sdv_stop = 0;
    double r = sdvoigt(deltnu, alphal, alphad, sdep);
    if (err) *err = g_sdv_stop;
    return r;
}

/* XLORENTZ, modm.f90:888-895 */
static double xlorentz(double z)
{
    double pi = 3.1415926535898;
    return 1. / (pi * (1. + (z * z)));
}

#define IS_LC(xf) (((xf) == -1) || ((xf) == -3) || ((xf) == -5))

/* LSF_SDVOIGT, modm.f90:567-704 (chi_fn always returns 1, :1286) */
static double lsf_sdvoigt(double xf, double rp, double rp2, double aip, double bip, double hwhm,
                          double wn, double xnu, double ad, int64_t mol, double sdep)
{
    const int64_t MOL_CO2 = 2, MOL_O2 = 7;
    double deltnuc = 25., diff = (wn + xnu) - deltnuc, sls = 0., chi = 1.;
    double deltxnu, xl1, xl2, xl3, y1, y1p, y2, y2p, xp4, yp1;
    if ((mol != MOL_O2) && (mol != MOL_CO2)) {
        if (IS_LC(xf)) {
            deltxnu = (wn - xnu);
            xl1 = sdvoigt(deltxnu, hwhm, ad, sdep);
            xl3 = sdvoigt(deltnuc, hwhm, ad, sdep);
            y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
            y1p = (1. + (aip * (1 / hwhm) * rp * (deltnuc)) + (bip * rp2));
            if (diff <= 0.) {
                deltxnu = (wn + xnu);
                xl2 = sdvoigt(deltxnu, hwhm, ad, sdep);
                y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                y2p = (1. - (aip * (1 / hwhm) * rp * (deltnuc)) + (bip * rp2));
                sls = (y1 * (xl1) - y1p * (xl3) + y2 * (xl2) - y2p * (xl3));
            } else {
                sls = y1 * (xl1) - y1p * (xl3);
            }
        } else {
            deltxnu = (wn - xnu);
            xl1 = sdvoigt(deltxnu, hwhm, ad, sdep);
            xl3 = sdvoigt(deltnuc, hwhm, ad, sdep);
            if (diff <= 0.) {
                deltxnu = (wn + xnu);
                xl2 = sdvoigt(deltxnu, hwhm, ad, sdep);
                sls = (xl1 + xl2 - (2 * xl3));
            } else {
                sls = (xl1 - xl3);
            }
        }
    } else {
        if ((fabs(wn - xnu) <= deltnuc) && (xf != -1) && (xf != -3) && (xf != -5)) {
            deltxnu = (wn - xnu);
            xl1 = sdvoigt(deltxnu, hwhm, ad, sdep);
            if (mol == MOL_O2) {
                if (diff <= 0.) {
                    deltxnu = (wn + xnu);
                    xl2 = sdvoigt(deltxnu, hwhm, ad, sdep);
                    sls = (xl1 + xl2);
                } else {
                    sls = (xl1);
                }
            } else {
                deltxnu = (wn - xnu);
                xl3 = sdvoigt(deltnuc, hwhm, ad, sdep);
                xl3 = xl3 * (2. - ((deltxnu * deltxnu) / (deltnuc * deltnuc)));
                sls = chi * (xl1 - xl3);
            }
        } else {
            if (mol == MOL_O2) {
                if (IS_LC(xf)) {
                    deltxnu = (wn - xnu);
                    xl1 = sdvoigt(deltxnu, hwhm, ad, sdep);
                    deltxnu = (wn + xnu);
                    xl2 = sdvoigt(deltxnu, hwhm, ad, sdep);
                    if (xf == -1) {
                        y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
                        y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                        sls = (xl1 * (y1) + xl2 * (y2));
                    } else {
                        sls = (xl1 + xl2);
                    }
                }
            } else {
                if ((xf == -1) || (xf == -3) || (xf != -5)) {        /* sic, modm.f90:659 */
                    deltxnu = (wn - xnu);
                    xl1 = sdvoigt(deltxnu, hwhm, ad, sdep);
                    xl3 = sdvoigt(deltnuc, hwhm, ad, sdep);
                    if (xf == -1 || xf == -5) {
                        y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
                        xp4 = xl3 * (2. - ((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc));
                        yp1 = (y1 - 1.) * (2. - ((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc));
                        sls = chi * (xl1 * (y1)-xp4 - xl3 * (yp1));
                    } else {
                        xp4 = xl3 * (2. - ((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc));
                        sls = chi * (xl1 - xp4);
                    }
                }
            }
        }
    }
    return sls;
}

/* LSF_LORTZ, modm.f90:706-831 */
static double lsf_lortz(double xf, double rp, double rp2, double aip, double bip, double hwhm,
                        double wn, double xnu, int64_t mol)
{
    const int64_t MOL_CO2 = 2, MOL_O2 = 7;
    double deltnuc = 25., diff = (wn + xnu) - deltnuc, sls = 0., chi = 1.;
    double deltxnu, xl1, xl2, xl3, y1, y1p, y2, y2p, xp4, yp1;
    if ((mol != MOL_O2) && (mol != MOL_CO2)) {
        if (IS_LC(xf)) {
            deltxnu = (wn - xnu);
            xl1 = xlorentz((deltxnu) / hwhm);
            xl3 = xlorentz((deltnuc) / hwhm);
            y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
            y1p = (1. + (aip * (1 / hwhm) * rp * (deltnuc)) + (bip * rp2));
            if (diff <= 0.) {
                deltxnu = (wn + xnu);
                xl2 = xlorentz((deltxnu) / hwhm);
                y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                y2p = (1. - (aip * (1 / hwhm) * rp * (deltnuc)) + (bip * rp2));
                sls = (y1 * (xl1)-y1p * (xl3) + y2 * (xl2)-y2p * (xl3)) / hwhm;
            } else {
                sls = (y1 * (xl1)-y1p * (xl3)) / hwhm;
            }
        } else {
            deltxnu = (wn - xnu);
            xl1 = xlorentz((deltxnu) / hwhm);
            xl3 = xlorentz((deltnuc) / hwhm);
            if (diff <= 0.) {
                deltxnu = (wn + xnu);
                xl2 = xlorentz((deltxnu) / hwhm);
                sls = (xl1 + xl2 - (2 * xl3)) / hwhm;
            } else {
                sls = (xl1 - xl3) / hwhm;
            }
        }
    } else {
        if ((fabs(wn - xnu) <= deltnuc) && (xf != -1) && (xf != -3) && (xf != -5)) {
            deltxnu = (wn - xnu);
            xl1 = xlorentz((deltxnu) / hwhm);
            if (mol == MOL_O2) {
                if (diff <= 0.) {
                    deltxnu = (wn + xnu);
                    xl2 = xlorentz((deltxnu) / hwhm);
                    sls = (xl1 + xl2) / hwhm;
                } else {
                    sls = (xl1) / hwhm;
                }
            } else {
                deltxnu = (wn - xnu);
                xl3 = xlorentz((deltnuc) / hwhm);
                xl3 = xl3 * (2. - ((deltxnu * deltxnu) / (deltnuc * deltnuc)));
                sls = chi * (xl1 - xl3) / hwhm;
            }
        } else {
            if (mol == MOL_O2) {
                if (IS_LC(xf)) {
                    deltxnu = (wn - xnu);
                    xl1 = xlorentz((deltxnu) / hwhm);
                    deltxnu = (wn + xnu);
                    xl2 = xlorentz((deltxnu) / hwhm);
                    if (xf == -1) {
                        y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
                        y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                        sls = (xl1 * (y1) + xl2 * (y2)) / hwhm;
                    } else {
                        sls = (xl1 + xl2) / hwhm;
                    }
                }
            } else {
                if (IS_LC(xf)) {
                    deltxnu = (wn - xnu);
                    xl1 = xlorentz((deltxnu) / hwhm);
                    xl3 = xlorentz((deltnuc) / hwhm);
                    if ((xf == -1) || (xf == -5)) {
                        y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
                        xp4 = xl3 * (2. - ((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc));
                        yp1 = (y1 - 1.) * (2. - ((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc));
                        sls = chi * (xl1 * (y1)-xp4 - xl3 * (yp1)) / hwhm;
                    } else {
                        xp4 = xl3 * (2. - (((wn - xnu) * (wn - xnu)) / (deltnuc * deltnuc)));
                        sls = chi * (xl1 - xp4) / hwhm;
                    }
                }
            }
        }
    }
    return sls;
}

/* HALFWHM_C, modm.f90:833-857.  as_ is in/out (H2O fix-up written back, :841). */
static double halfwhm_c(double af, double *as_, double rt, double xtild, double rhorat, int64_t mol,
                        const double *rho_molec, double rho_self, const int32_t *brd_flg,
                        const double *brd_hw, const double *brd_tmp)
{
    if ((mol == 1) && (*as_ == 0.)) *as_ = 5 * af;
    double alfa0i = af * pow(rt, xtild);
    double hwhmsi = *as_ * pow(rt, xtild);
    double h = alfa0i * (rhorat - rho_self) + hwhmsi * rho_self;
    int64_t sflg = 0;
    for (int k = 0; k < 7; k++) sflg += brd_flg[k];
    if (sflg > 0) {
        double alfsum = 0., sflgrho = 0.;
        for (int k = 0; k < 7; k++) {
            double tmpcor = pow(rt, brd_tmp[k]);
            double alfa_tmp = brd_hw[k] * tmpcor;
            alfsum = alfsum + rho_molec[k] * (double)brd_flg[k] * alfa_tmp;
        }
        for (int k = 0; k < 7; k++) sflgrho = sflgrho + rho_molec[k] * (double)brd_flg[k];
        h = (rhorat - sflgrho) * alfa0i + alfsum;
        if (mol <= 7 ? (brd_flg[mol - 1] == 0) : 1) h = h + rho_self * (hwhmsi - alfa0i);
    }
    return h;
}

/* HALFWHM_D, modm.f90:442-454 */
static double halfwhm_d(int64_t mol, int64_t iso, double xnu, double t)
{
    double m = ISO_SMASS[(mol - 1) * 9 + (iso - 1)];
    return (xnu / CLIGHTref) * sqrt(2. * log(2.) * ((BOLTZref * t) / (m / AVOGADref)));
}

/* INTENS, modm.f90:860-865 */
static double intens(double t, double s0s, double es, double radct, double t0, double xnus, double xipsf)
{
    double s = s0s * (exp(-radct * es / t) / exp(-radct * es / t0)) * xipsf;
    return s * ((1 + exp(-(radct * xnus / t))) / (xnus * (1 - exp(-(radct * xnus / t0)))));
}

/* LINES, modm.f90:277-440 */
static int lines(double xn, double wn, double t, int64_t nmol, const double *wk, double wbrod,
                 double radct, double t0, double *o_by_mol, int64_t ld_mol, double xn0, double rft,
                 double p, double p0, double sclcpl, double sclhw, double y0res, const double *scor,
                 int64_t ibrd, orc_lines *ln, int64_t *sel_count, uint64_t *sel_hash)
{
    const double templc[4] = {200.0, 250.0, 296.0, 340.0};
    double a[4], b[4], rho_molec[7];
    int32_t brd_flg[7];
    double brd_tmp[7], brd_hw[7];
    double aip = 0., bip = 0.;
    double deltnuc = 25.;
    double wtot = 0.;
    for (int64_t i = 0; i < nmol; i++) wtot = wtot + wk[i];
    wtot = wtot + wbrod;
    double rp = p / p0, rp2 = rp * rp;
    int ilc = 1;
    for (int il = 1; il <= 3; il++) {
        ilc = il;
        if (t < templc[ilc]) break;
    }
    double rectlc = 1.0 / (templc[ilc] - templc[ilc - 1]);
    double tmpdif = t - templc[ilc - 1];
    double rt = t / t0;
    double rhorat = (xn / xn0);
    for (int k = 0; k < 7; k++) rho_molec[k] = rhorat * wk[k] / wtot;   /* wk(1:7): caller guarantees 7 readable */
    int64_t cnt = 0;
    uint64_t hash = 0;

    for (int64_t i = 1; i <= nmol; i++) {
        double w_species = wk[i - 1], ol;
        if (w_species == 0.) {
            ol = 0.;
            o_by_mol[(i - 1) * ld_mol] = ol;
            continue;
        }
        double rho_self = (i <= 7) ? rho_molec[i - 1] : rhorat * wk[i - 1] / wtot;   /* UB in reference for i>7 */
        double sf = 0.;
        int64_t j = 0;
        while (j < ln->nblm[i - 1]) {
            j = j + 1;
            int64_t jj = j;
            double xgj = ln->xg[IX(i, j)];
            if (IS_LC(xgj)) {
                jj = j + 1;
                if (jj + 1 > ln->iim) return fail(30, "coupling record beyond line store");
                a[0] = ln->xnu0[IX(i, jj)];  b[0] = ln->s0[IX(i, jj)];
                a[1] = ln->alpf[IX(i, jj)];  b[1] = ln->e[IX(i, jj)];
                a[2] = ln->rmol[IX(i, jj)];  b[2] = ln->alps[IX(i, jj)];
                a[3] = ln->x[IX(i, jj)];     b[3] = ln->deltnu[IX(i, jj)];
                if ((xgj == -5) && (j > 1) && (ln->xg[IX(i, j - 1)] == -5)) {
                    jj = jj + 1;
                    double rho_for = (rhorat - rho_self) / rhorat;
                    double rho_sel = rho_self / rhorat;
                    a[0] = rho_for * a[0] + rho_sel * ln->xnu0[IX(i, jj)];
                    b[0] = rho_for * b[0] + rho_sel * ln->s0[IX(i, jj)];
                    a[1] = rho_for * a[1] + rho_sel * ln->alpf[IX(i, jj)];
                    b[1] = rho_for * b[1] + rho_sel * ln->e[IX(i, jj)];
                    a[2] = rho_for * a[2] + rho_sel * ln->rmol[IX(i, jj)];
                    b[2] = rho_for * b[2] + rho_sel * ln->alps[IX(i, jj)];
                    a[3] = rho_for * a[3] + rho_sel * ln->x[IX(i, jj)];
                    b[3] = rho_for * b[3] + rho_sel * ln->deltnu[IX(i, jj)];
                }
                aip = a[ilc - 1] + ((a[ilc] - a[ilc - 1]) * rectlc) * tmpdif;
                bip = b[ilc - 1] + ((b[ilc] - b[ilc - 1]) * rectlc) * tmpdif;
            }
            if (xgj == -1) {
                aip = aip * sclcpl + y0res;
                bip = bip * sclcpl + y0res;
            }
            if (xgj == -3) {
                aip = aip * sclhw;
                bip = bip * sclhw;
            }
            double xnu0 = ln->xnu0[IX(i, j)];
            double s0_adj = ln->s0[IX(i, j)] * (xnu0 * (1.0 - exp(-(radct * xnu0 / t0))));
            double xnu = xnu0 + (ln->deltnu[IX(i, j)] * (xn / xn0));
            if (i <= 7 && ibrd != 0) {
                double s = 0.;
                for (int k = 1; k <= 7; k++)
                    s = s + rho_molec[k - 1] * (double)ln->brd_mol_flg[IB(i, k, j)] *
                                (ln->brd_mol_shft[IB(i, k, j)] - ln->deltnu[IX(i, j)]);
                xnu = xnu + s;
            }
            if ((fabs(wn - xnu) > deltnuc) && (i != 7)) { j = jj; continue; }     /* :384 */
            cnt++;
            hash += orc_line_key(i, j);

            int64_t iso = ln->iso[IX(i, j)];
            if (iso < 1 || iso > 9) return fail(31, "isotope index out of 1..9");
            double xipsf = scor[(i - 1) + (iso - 1) * 42];
            double stild = intens(t, s0_adj, ln->e[IX(i, j)], radct, t0, xnu, xipsf);
            double xtild = ln->x[IX(i, j)];
            for (int k = 0; k < 7; k++) { brd_flg[k] = 0; brd_hw[k] = 0.0; brd_tmp[k] = 0.0; }
            if (i <= 7 && ibrd != 0) {
                for (int k = 1; k <= 7; k++) {
                    brd_flg[k - 1] = ln->brd_mol_flg[IB(i, k, j)];
                    brd_hw[k - 1] = ln->brd_mol_hw[IB(i, k, j)];
                    brd_tmp[k - 1] = ln->brd_mol_tmp[IB(i, k, j)];
                }
            }
            double hwhm_c = halfwhm_c(ln->alpf[IX(i, j)], &ln->alps[IX(i, j)], rt, xtild, rhorat, i,
                                      rho_molec, rho_self, brd_flg, brd_hw, brd_tmp);
            double hwhm_d = halfwhm_d(i, iso, xnu, t);
            if (xgj == -3.) hwhm_c = hwhm_c * (1 - (aip * (rp)) - (bip * (rp2)));
            double zeta = hwhm_c / (hwhm_c + hwhm_d);
            int ilshp = 1;
            if ((fabs(wn - xnu) > (100. * hwhm_d)) || (zeta > 0.99)) ilshp = 0;
            double sls;
            if (ilshp == 0) {
                sls = lsf_lortz(xgj, rp, rp2, aip, bip, hwhm_c, wn, xnu, i);
            } else {
                sls = lsf_sdvoigt(xgj, rp, rp2, aip, bip, hwhm_c, wn, xnu, hwhm_d, i, ln->sdep[IX(i, j)]);
                g_nvoigt++;
                g_branch[0]++;
                if (i == 2) g_branch[2]++;
                if (i == 2 && xgj == -1) g_branch[3]++;
                if (i != 2 && i != 7 && IS_LC(xgj)) g_branch[4]++;
                if (i == 7 && IS_LC(xgj)) g_branch[5]++;
                if (!IS_LC(xgj) && xgj != 0) g_branch[6]++;
                if ((wn + xnu) - 25. <= 0.) g_branch[7]++;
            }
            sf = sf + (stild * sls);
            j = jj;
        }
        double spsd = w_species * sf;
        ol = rft * spsd;
        o_by_mol[(i - 1) * ld_mol] = ol;
    }
    if (sel_count) *sel_count = cnt;
    if (sel_hash) *sel_hash = hash;
    return 0;
}

/* ======================================================================= */
/*  MODM, modm.f90:21-274                                                   */
/* ======================================================================= */
int orc_modm(int64_t nwn, const double *wn, double dvset, int64_t nlay,
             const double *p, const double *t, const double *clw,
             double *o, double *o_by_mol, double *oc, double *o_clw, double *odxsec,
             int64_t nmol, const double *wkl, const double *wbrodl,
             double sclcpl, double sclhw, double y0res,
             const double cntnm[7], int64_t ixsect, const double *odxsec_in,
             int64_t ibrd, const double *scor, orc_lines *ln,
             int64_t *sel_count, uint64_t *sel_hash, int64_t *n_voigt)
{
    static const int64_t index_cont[6] = {1, 2, 3, 7, 22, 99};     /* :166 */
    const int ncont = 6;
    static double absrb[N_ABSRB];
    double wkc[60];
    memset(wkc, 0, sizeof wkc);
    g_sdv_stop = 0;
    g_nvoigt = 0;
    for (int ib_ = 0; ib_ < 8; ib_++) g_branch[ib_] = 0;
    if (nwn < 1 || nlay < 1 || nmol < 1 || nmol > ORC_MXMOL) return fail(40, "bad dimensions");

    double radcn2 = RADCN2ref;
    double v1 = wn[0], v2 = wn[nwn - 1];                             /* :180-185 */
    double dvabs = 1.0;
    double v1abs = (double)((int64_t)v1) - 3. * dvabs;
    double v2abs = (double)((int64_t)(v2 + 3. * dvabs + 0.5));
    int64_t nptabs = (int64_t)((v2abs - v1abs) / dvabs + 1.5);

    const size_t LD2 = (size_t)nwn * ORC_MXMOL;
    double *oc_rayl = (double *)calloc((size_t)nwn * (size_t)nlay, 8);
    for (int64_t k = 0; k < nlay; k++) {                            /* :192-195 */
        memset(oc + (size_t)k * LD2, 0, LD2 * 8);
        memset(o_by_mol + (size_t)k * LD2, 0, LD2 * 8);              /* (allocate leaves it undefined; we zero) */
        memset(odxsec + (size_t)k * nwn, 0, (size_t)nwn * 8);
        memset(o + (size_t)k * nwn, 0, (size_t)nwn * 8);
    }
    if (ixsect == 1 && odxsec_in)                                    /* :197-198 result of monortm_xsec_sub */
        memcpy(odxsec, odxsec_in, (size_t)nwn * (size_t)nlay * 8);

    int rc = 0;
    for (int64_t k = 1; k <= nlay && rc == 0; k++) {                 /* :200 */
        double pave = p[k - 1], tave = t[k - 1], wbroad = wbrodl[k - 1];
        double xkt = tave / radcn2;
        const double *wklk = wkl + (size_t)(k - 1) * ORC_MXMOL;
        for (int64_t m = 0; m < nmol; m++) wkc[m] = wklk[m];         /* :208-209 */
        if (nmol < 22) wkc[21] = wbroad;
        for (int icount = 1; icount <= ncont; icount++) {            /* :210-247 */
            int64_t im = index_cont[icount - 1];
            memset(absrb, 0, sizeof absrb);
            rc = orc_contnm_one(im, cntnm, pave, tave, wkc, wbroad, nmol, v1, v2, v1abs, v2abs, nptabs, absrb + 1);
            if (rc) break;
            /* absrb + 1: xint reads A(J-1) with J>=1 never below index 1 in practice; keep one guard cell */
            double *dst = (icount < ncont) ? oc + (size_t)(k - 1) * LD2 + (size_t)(im - 1) * nwn
                                           : oc_rayl + (size_t)(k - 1) * nwn;
            if (dvset != 0) xint(v1abs, v2abs, dvabs, absrb + 1, 1.0, v1, dvset, dst, 1, nwn);
            if (dvset == 0) {
                for (int64_t iw = 1; iw <= nwn; iw++)
                    xint(v1abs, v2abs, dvabs, absrb + 1, 1.0, wn[iw - 1], 1.0, dst + (iw - 1), 1, 1);
            }
            if (icount < ncont) {
                for (int64_t iw = 1; iw <= nwn; iw++) dst[iw - 1] = dst[iw - 1] * orc_radfn(wn[iw - 1], xkt);
            } else {
                for (int64_t iw = 1; iw <= nwn; iw++) dst[iw - 1] = dst[iw - 1] * wn[iw - 1] / 1.0e4;
            }
        }
        if (rc) break;
        const double *scor_k = scor + (size_t)(k - 1) * 42 * 9;      /* tips_2003 result, :250 */

        for (int64_t m = 1; m <= nwn; m++) {                         /* :253 */
            /* INITI, modm.f90:868-883 */
            double radct = PLANCKref * CLIGHTref / BOLTZref;
            double t0 = 296., p0 = 1013.25;
            double xn0 = (p0 / (BOLTZref * t0)) * 1.E+3;
            double xn = (p[k - 1] / (BOLTZref * t[k - 1])) * 1.E+3;
            double rft = wn[m - 1] * tanh((radct * wn[m - 1]) / (2 * t[k - 1]));   /* :257 */
            size_t fl = (size_t)(m - 1) + (size_t)(k - 1) * nwn;
            rc = lines(xn, wn[m - 1], t[k - 1], nmol, wklk, wbrodl[k - 1], radct, t0,
                       o_by_mol + (size_t)(k - 1) * LD2 + (size_t)(m - 1), nwn, xn0, rft, p[k - 1], p0,
                       sclcpl, sclhw, y0res, scor_k, ibrd, ln,
                       sel_count ? sel_count + fl : NULL, sel_hash ? sel_hash + fl : NULL);
            if (rc) break;
            o_clw[fl] = orc_odclw(wn[m - 1], t[k - 1], clw[k - 1]);              /* :264 */
            for (int64_t imol = 1; imol <= nmol; imol++)                          /* :265-267 */
                o[fl] = o[fl] + o_by_mol[(size_t)(k - 1) * LD2 + (size_t)(imol - 1) * nwn + (size_t)(m - 1)];
            double soc = 0.;                                                      /* sum(oc(m,1:22,k)) */
            for (int64_t im = 1; im <= index_cont[4]; im++)
                soc = soc + oc[(size_t)(k - 1) * LD2 + (size_t)(im - 1) * nwn + (size_t)(m - 1)];
            o[fl] = o[fl] + odxsec[fl] + oc_rayl[fl] + soc + o_clw[fl];           /* :268-269 */
        }
    }
    free(oc_rayl);
    if (n_voigt) *n_voigt = g_nvoigt;
    if (rc) return rc;
    if (g_sdv_stop) return fail(41, "SDVOIGT: REAL(v) < 0 (modm.f90:1062 STOP)");
    return 0;
}

/* ======================================================================= */
/*  RTMmono.f90                                                             */
/* ======================================================================= */
/* bb_fn, RTMmono.f90:223-237 */
double orc_bb_fn(double v, double fbeta)
{
    return RADCN1ref * (v * v * v) / (exp(v * fbeta) - 1.);
}

/* RAD_UP_DN, RTMmono.f90:157-221.  o is (nwn,nlayer), tz is (0:nlayer). */
static int rad_up_dn(const double *t, int64_t nlayer, const double *tz, const double *wn, double *rup,
                     double *trtot, const double *o, double *rdn, int64_t nwn, int64_t idu, int64_t irt)
{
    if (idu != 1) return fail(50, "ERROR IN IDU. OPTION NOT SUPPORTED YET");
    double *bbvec = (double *)malloc((size_t)(nlayer + 1) * 8);
    double *bbavec = (double *)malloc((size_t)(nlayer + 1) * 8);
    double radcn2 = RADCN2ref;
    for (int64_t i = 1; i <= nwn; i++) {
        double vv = wn[i - 1];
        rup[i - 1] = 0.;
        rdn[i - 1] = 0.;
        trtot[i - 1] = 1.;
        double odtot = 0.;
        for (int64_t layer = 1; layer <= nlayer; layer++) {
            odtot = odtot + o[(i - 1) + (size_t)(layer - 1) * nwn];
            double beta = radcn2 / t[layer - 1];
            double beta_a = radcn2 / tz[layer];
            bbvec[layer] = orc_bb_fn(vv, beta);
            bbavec[layer] = orc_bb_fn(vv, beta_a);
            beta_a = radcn2 / tz[layer - 1];
            bbavec[layer - 1] = orc_bb_fn(vv, beta_a);
        }
        if (irt != 3) {
            double odt = odtot;
            for (int64_t layer = 1; layer <= nlayer; layer++) {
                double bb = bbvec[layer], bba = bbavec[layer];
                double odvi = o[(i - 1) + (size_t)(layer - 1) * nwn];
                double tri = exp(-odvi);
                odt = odt - odvi;
                trtot[i - 1] = exp(-odt);
                double pade = 0.193 * odvi + 0.013 * (odvi * odvi);
                rup[i - 1] = rup[i - 1] + trtot[i - 1] * (1. - tri) * (bb + pade * bba) / (1. + pade);
            }
        }
        double odt = odtot;
        for (int64_t layer = nlayer; layer >= 1; layer--) {
            double bb = bbvec[layer], bba = bbavec[layer - 1];
            double odvi = o[(i - 1) + (size_t)(layer - 1) * nwn];
            odt = odt - odvi;
            double tri = exp(-odvi);
            trtot[i - 1] = exp(-odt);
            double pade = 0.193 * odvi + 0.013 * (odvi * odvi);
            rdn[i - 1] = rdn[i - 1] + trtot[i - 1] * (1. - tri) * (bb + pade * bba) / (1. + pade);
        }
        trtot[i - 1] = exp(-odtot);
    }
    free(bbvec);
    free(bbavec);
    return 0;
}

/* RTM, RTMmono.f90:13-155 */
int orc_rtm(int64_t iout, int64_t irt, int64_t nwn, const double *wn, int64_t nlay,
            const double *t, const double *tz, const double *o, double *tmpsfc,
            double *rup, double *trtot, double *rdn, const double *reflc,
            const double *emiss, double *rad, double *tb, int64_t idu)
{
    double radcn1 = RADCN1ref, radcn2 = RADCN2ref;
    int rc = rad_up_dn(t, nlay, tz, wn, rup, trtot, o, rdn, nwn, idu, irt);
    if (rc) return rc;
    double tsky = 2.75;
    if (irt == 3 || irt == 2) *tmpsfc = tsky;
    double alph = radcn2 / tsky;
    double beta = radcn2 / *tmpsfc;
    for (int64_t i = 1; i <= nwn; i++) {
        double vv = wn[i - 1];
        double surfrad = orc_bb_fn(vv, beta);
        double cosmos = orc_bb_fn(vv, alph);
        double esfc = emiss[i - 1], rsfc = reflc[i - 1];
        if (irt == 1)
            rad[i - 1] = rup[i - 1] + trtot[i - 1] * (esfc * surfrad + rsfc * (rdn[i - 1] + trtot[i - 1] * cosmos));
        if (irt == 2) rad[i - 1] = rup[i - 1] + trtot[i - 1] * (rdn[i - 1] + trtot[i - 1] * cosmos);
        if (irt == 3) rad[i - 1] = rdn[i - 1] + (trtot[i - 1] * cosmos);
        if (iout == 1) {
            double x = radcn1 * (wn[i - 1] * wn[i - 1] * wn[i - 1]) / rad[i - 1] + 1.;
            tb[i - 1] = radcn2 * wn[i - 1] / log(x);
        }
    }
    return 0;
}

/* calctmr, RTMmono.f90:239-325 */
int orc_calctmr(int64_t nlayrs, int64_t nwn, const double *wn, const double *t,
                const double *tz, const double *o, double *tmr)
{
    double radcn1 = RADCN1ref, radcn2 = RADCN2ref;
    double *bbvec = (double *)malloc((size_t)(nlayrs + 1) * 8);
    double *bbavec = (double *)malloc((size_t)(nlayrs + 1) * 8);
    for (int64_t ifr = 1; ifr <= nwn; ifr++) {
        double sumexp = 0.;
        double vv = wn[ifr - 1];
        double trtot = 1.;
        double odtot = 0.;
        for (int64_t ilay = 1; ilay <= nlayrs; ilay++) {
            odtot = odtot + o[(ifr - 1) + (size_t)(ilay - 1) * nwn];
            double beta = radcn2 / t[ilay - 1];
            double beta_a = radcn2 / tz[ilay];
            bbvec[ilay] = orc_bb_fn(vv, beta);
            bbavec[ilay] = orc_bb_fn(vv, beta_a);
            beta_a = radcn2 / tz[ilay - 1];
            bbavec[ilay - 1] = orc_bb_fn(vv, beta_a);
        }
        double odt = odtot;
        for (int64_t ilay = nlayrs; ilay >= 1; ilay--) {
            double bb = bbvec[ilay], bba = bbavec[ilay - 1];
            double odvi = o[(ifr - 1) + (size_t)(ilay - 1) * nwn];
            odt = odt - odvi;
            double tri = exp(-odvi);
            trtot = exp(-odt);
            double pade = 0.193 * odvi + 0.013 * (odvi * odvi);
            double beff = (bb + pade * bba) / (1. + pade);
            sumexp = sumexp + beff * trtot * (1 - tri);
        }
        double radtmr = sumexp / (1. - exp(-1 * odtot));
        double x = radcn1 * (wn[ifr - 1] * wn[ifr - 1] * wn[ifr - 1]) / radtmr + 1.;
        tmr[ifr - 1] = radcn2 * wn[ifr - 1] / log(x);
    }
    free(bbvec);
    free(bbavec);
    return 0;
}


/* ======================================================================= */
/*  monortm_sub.F90: cross sections                                          */
/* ======================================================================= */
#define XSPD_INT_MAX 10000000      /* xspd_int(0:10000000), monortm_sub.F90:1755 */

/* convolve, monortm_sub.F90:1751-1834.  xspd is 1-based in the reference (xspd[i-1] = xspd(i)); reads one element past
 * the table at the upper edge (xspd(ind+2) with ind+2 = nptsx+1, :1779) -- the caller's array is xspd(150000), so that
 * element exists and holds whatever an earlier region left there; it is multiplied by coef = 0 up to rounding.  Here the
 * table is passed with its length and elements beyond it read as 0. */
static int convolve(const double *xspd, int64_t nxspd, double v1x, double v2x, double delvx, double pd, double hwdop,
                    double tave, double pave, const double *wn, int64_t nwn, double *xspave, double *xspd_int)
{
    const double p0 = 1013.;
#define XSPD(i) (((i) >= 1 && (i) <= nxspd) ? xspd[(i) - 1] : 0.)
    double hwpave = 0.1 * (pave / p0) * (273.15 / tave);
    double hwd = 0.1 * (pd / p0) * (273.15 / tave);
    hwd = hwd > hwdop ? hwd : hwdop;
    if (hwd > hwpave) hwpave = 1.001 * hwd;
    double hwb = hwpave - hwd;
    double ratio = 0.25;
    double step = ratio * hwb;
    if (step > delvx) step = delvx;
    double q = (v2x - v1x) / step;
    if (!(q < (double)XSPD_INT_MAX + 1.)) return fail(52, "convolve: NPTS exceeds xspd_int(0:10000000) (monortm_sub.F90:1755)");
    int64_t npts = (int64_t)q;
    step = (v2x - v1x) / (double)npts;
    ratio = step / hwb;
    for (int64_t i = 0; i <= npts; i++) {
        double vv = v1x + (double)i * step;
        double delvv = vv - v1x;
        int64_t ind = (int64_t)(delvv / delvx);
        double coef = (delvv - (double)ind * delvx) / delvx;
        xspd_int[i] = (1. - coef) * XSPD(ind + 1) + coef * XSPD(ind + 2);
    }
    double hwb2 = hwb * hwb;
    for (int64_t iwn = 1; iwn <= nwn; iwn++) {
        double w = wn[iwn - 1];
        if (w < v1x || w > v2x) { xspave[iwn - 1] = 0.; continue; }
        if (hwb / hwd > 0.1) {
            double wn_v1x = w - v1x;
            int64_t ind = (int64_t)(wn_v1x / step);
            double dvlo = w - (v1x + (double)ind * step);
            double dvhi = w - (v1x + (double)(ind + 1) * step);
            /* xspd_int(ind+1) with ind = npts at w = v2x is one past what loop 500 filled: stale in the reference, 0 here */
            double xi1 = (ind + 1 <= npts) ? xspd_int[ind + 1] : 0.;
            double answer = (hwb / (hwb2 + dvlo * dvlo)) * xspd_int[ind] + (hwb / (hwb2 + dvhi * dvhi)) * xi1;
            int64_t j = 1;
            for (;;) {
                double contlo, conthi;
                double vlo = v1x + (double)(ind - j) * step;
                if (vlo > v1x) {
                    dvlo = w - vlo;
                    contlo = (hwb / (hwb2 + dvlo * dvlo)) * xspd_int[ind - j];
                } else contlo = 0.;
                double vhi = v1x + (double)(ind + j + 1) * step;
                if (vhi < v2x) {
                    dvhi = w - vhi;
                    conthi = (hwb / (hwb2 + dvhi * dvhi)) * xspd_int[ind + j + 1];
                } else conthi = 0.;
                double xincr = contlo + conthi;
                if ((xincr / answer) < ratio * 1e-6) break;
                answer = answer + xincr;
                j = j + 1;
                /* zero or NaN table around the frequency: 0/0 is never < ratio*1e-6 and the reference loops forever */
                if (j > npts + 2) return fail(53, "convolve: the outward sum does not terminate (monortm_sub.F90:1800-1821)");
            }
            xspave[iwn - 1] = answer * step / 3.14159;
        } else {
            double wn_v1x = w - v1x;
            int64_t ind = (int64_t)(wn_v1x / delvx);
            double coef = (wn_v1x - (double)ind * delvx) / delvx;
            xspave[iwn - 1] = (1. - coef) * XSPD(ind) + coef * XSPD(ind + 1);      /* xspd(0) at w = v1x: 0 here */
        }
    }
#undef XSPD
    return 0;
}

int orc_convolve(const double *xspd, int64_t nxspd, double v1x, double v2x, double delvx, double pd, double hwdop,
                 double tave, double pave, const double *wn, int64_t nwn, double *xspave)
{
    double *xi = (double *)calloc((size_t)XSPD_INT_MAX + 2, sizeof(double));
    if (!xi) return fail(50, "out of memory");
    int rc = convolve(xspd, nxspd, v1x, v2x, delvx, pd, hwdop, tave, pave, wn, nwn, xspave, xi);
    free(xi);
    return rc;
}

/* MONORTM_XSEC_SUB, monortm_sub.F90:1540-1749 */
int orc_xsec_sub(int64_t nwn, const double *wn, int64_t nlay, const double *p, const double *t,
                 int64_t nreg, const orc_xs_region *regs, int64_t ld_xamnt, const double *xamnt, double *odxsec)
{
    const double dvbuf = 1.0;
    size_t nn = (size_t)nwn * (size_t)nlay;
    double *xstot = (double *)calloc(nn ? nn : 1, sizeof(double));
    double *xsmoltot = (double *)calloc(nn ? nn : 1, sizeof(double));
    double *xspave = (double *)calloc((size_t)nwn + 1, sizeof(double));
    double *xi = (double *)calloc((size_t)XSPD_INT_MAX + 2, sizeof(double));
    int64_t maxpts = 1;
    for (int64_t r = 0; r < nreg; r++) if (regs[r].npts > maxpts) maxpts = regs[r].npts;
    double *xspd = (double *)calloc((size_t)maxpts + 2, sizeof(double));
    if (!xstot || !xsmoltot || !xspave || !xi || !xspd) return fail(50, "out of memory");
    int rc = 0;
    int64_t r = 0;
    while (r < nreg && !rc) {
        int32_t ixmol = regs[r].ixmol;
        memset(xsmoltot, 0, nn * sizeof(double));                                  /* :1644 */
        for (; r < nreg && regs[r].ixmol == ixmol && !rc; r++) {                   /* loop 5000 */
            const orc_xs_region *g = &regs[r];
            int need = 0;                                                          /* :1647-1653 */
            for (int64_t ipt = 1; ipt <= nwn; ipt++)
                if (wn[ipt - 1] >= g->v1fx - dvbuf && wn[ipt - 1] <= g->v2fx + dvbuf) { need = 1; break; }
            if (!need) continue;
            for (int64_t il = 1; il <= nlay && !rc; il++) {                        /* loop 4000 */
                double pave = p[il - 1], tave = t[il - 1];
                double coef1 = 1., coef2 = 0.;
                int ind1, ind2 = 1, it = 1;
                if (g->ntemp == 1 || tave <= g->tx[it - 1]) {
                    ind1 = 1;
                } else {
                    for (;;) {
                        it = it + 1;
                        if (it > g->ntemp) { ind1 = g->ntemp; ind2 = g->ntemp; break; }
                        else if (tave <= g->tx[it - 1]) {
                            ind1 = it - 1;
                            ind2 = it;
                            coef1 = (tave - g->tx[it - 1]) / (g->tx[it - 2] - g->tx[it - 1]);
                            coef2 = 1. - coef1;
                            break;
                        }
                    }
                }
                double pd = coef1 * g->pdx[ind1 - 1] + coef2 * g->pdx[ind2 - 1];
                double xkt1 = g->tx[ind1 - 1] / RADCN2ref, xkt2 = g->tx[ind2 - 1] / RADCN2ref;
                double delvx = (g->v2x - g->v1x) / (double)(g->npts - 1);
                for (int64_t i = 1; i <= g->npts; i++) {                           /* loop 3300 */
                    double vv = g->v1x + (double)(i - 1) * delvx;
                    xspd[i - 1] = coef1 * g->xsdat[ind1 - 1][i - 1] / orc_radfn(vv, xkt1) +
                                  coef2 * g->xsdat[ind2 - 1][i - 1] / orc_radfn(vv, xkt2);
                }
                double hwdop = g->xdoplr * sqrt(tave / 296.);
                rc = convolve(xspd, g->npts, g->v1x, g->v2x, delvx, pd, hwdop, tave, pave, wn, nwn, xspave, xi);
                if (rc) break;
                for (int64_t iw = 0; iw < nwn; iw++)
                    xsmoltot[(size_t)iw + (size_t)(il - 1) * nwn] = xsmoltot[(size_t)iw + (size_t)(il - 1) * nwn] + xspave[iw];
            }
        }
        for (int64_t il = 1; il <= nlay; il++)                                     /* loop 5500 */
            for (int64_t iw = 0; iw < nwn; iw++) {
                size_t k = (size_t)iw + (size_t)(il - 1) * nwn;
                xstot[k] = xstot[k] + xamnt[(size_t)ixmol + (size_t)(il - 1) * ld_xamnt] * xsmoltot[k];
            }
    }
    for (int64_t il = 1; il <= nlay; il++) {                                       /* loop 6500 */
        double xkt = t[il - 1] / RADCN2ref;
        for (int64_t iw = 0; iw < nwn; iw++) {
            size_t k = (size_t)iw + (size_t)(il - 1) * nwn;
            odxsec[k] = xstot[k] * orc_radfn(wn[iw], xkt);
        }
    }
    free(xstot); free(xsmoltot); free(xspave); free(xi); free(xspd);
    return rc;
}

/* ---- building blocks exported for the reference-text pins (tests/test_ref_goldens.py) ---------------------- */
double orc_lsf_lortz(double xf, double rp, double rp2, double aip, double bip, double hwhm, double wn, double xnu, int64_t mol)
{
    return lsf_lortz(xf, rp, rp2, aip, bip, hwhm, wn, xnu, mol);
}
double orc_lsf_sdvoigt(double xf, double rp, double rp2, double aip, double bip, double hwhm, double wn, double xnu,
                       double ad, int64_t mol, double sdep)
{
    return lsf_sdvoigt(xf, rp, rp2, aip, bip, hwhm, wn, xnu, ad, mol, sdep);
}
double orc_halfwhm_d(int64_t mol, int64_t iso, double xnu, double t) { return halfwhm_d(mol, iso, xnu, t); }
double orc_intens(double t, double s0s, double es, double radct, double t0, double xnus, double xipsf)
{
    return intens(t, s0s, es, radct, t0, xnus, xipsf);
}
double orc_xlorentz(double z) { return xlorentz(z); }
/* XINT with 1-based arrays a[0..na-1] = A(1..na), r3[0..] = R3(1..) as the reference passes them */
void orc_xint(double v1a, double v2a, double dva, const double *a, double afact, double vft, double dvr3, double *r3,
              int64_t n1r3, int64_t n2r3)
{
    xint(v1a, v2a, dva, a, afact, vft, dvr3, r3, n1r3, n2r3);
}

/* ======================================================================= */
/*  tips_2003.f90 (harness side of the boundary: scor = Q(296)/Q(T))        */
/* ======================================================================= */
#include "../monortm_b200/csrc/tables/tips_tables.inc"

/* AtoB, tips_2003.f90:4610-4700: 4-point Lagrange (3-point at the ends of the table); a[], b[] are used 1-based */
static double tips_atob(double aa, const double *a0, const double *b0, int npt)
{
    double bb = 0.;
#define a(i) a0[(i) - 1]
#define b(i) b0[(i) - 1]
#define NZ(d) ((d) == 0. ? 0.0001 : (d))
    for (int i = 2; i <= npt; i++) {
        if (a(i) >= aa) {
            if (i < 3 || i == npt) {
                int j = i;
                if (i < 3) j = 3;
                if (i == npt) j = npt;
                double d01 = NZ(a(j - 2) - a(j - 1)), d02 = NZ(a(j - 2) - a(j));
                double d11 = NZ(a(j - 1) - a(j - 2)), d12 = NZ(a(j - 1) - a(j));
                double d21 = NZ(a(j) - a(j - 2)), d22 = NZ(a(j) - a(j - 1));
                double c0 = (aa - a(j - 1)) * (aa - a(j)) / (d01 * d02);
                double c1 = (aa - a(j - 2)) * (aa - a(j)) / (d11 * d12);
                double c2 = (aa - a(j - 2)) * (aa - a(j - 1)) / (d21 * d22);
                bb = c0 * b(j - 2) + c1 * b(j - 1) + c2 * b(j);
            } else {
                int j = i;
                double d01 = NZ(a(j - 2) - a(j - 1)), d02 = NZ(a(j - 2) - a(j)), d03 = NZ(a(j - 2) - a(j + 1));
                double d11 = NZ(a(j - 1) - a(j - 2)), d12 = NZ(a(j - 1) - a(j)), d13 = NZ(a(j - 1) - a(j + 1));
                double d21 = NZ(a(j) - a(j - 2)), d22 = NZ(a(j) - a(j - 1)), d23 = NZ(a(j) - a(j + 1));
                double d31 = NZ(a(j + 1) - a(j - 2)), d32 = NZ(a(j + 1) - a(j - 1)), d33 = NZ(a(j + 1) - a(j));
                double c0 = (aa - a(j - 1)) * (aa - a(j)) * (aa - a(j + 1));
                c0 = c0 / (d01 * d02 * d03);
                double c1 = (aa - a(j - 2)) * (aa - a(j)) * (aa - a(j + 1));
                c1 = c1 / (d11 * d12 * d13);
                double c2 = (aa - a(j - 2)) * (aa - a(j - 1)) * (aa - a(j + 1));
                c2 = c2 / (d21 * d22 * d23);
                double c3 = (aa - a(j - 2)) * (aa - a(j - 1)) * (aa - a(j));
                c3 = c3 / (d31 * d32 * d33);
                bb = c0 * b(j - 2) + c1 * b(j - 1) + c2 * b(j) + c3 * b(j + 1);
            }
            break;                                               /* GO TO 100 */
        }
    }
#undef NZ
#undef a
#undef b
    return bb;
}

/* TIPS_2003, tips_2003.f90:2-298.  scor is (42,9).  Returns 0, or 11 where the reference STOPs
 * ("partition sum less than 0.": T outside 70..3000 K gives Qt = -1 in every QT_* routine). */
int orc_tips_2003(int64_t mol_max, double temp_lbl, double *scor)
{
    if (mol_max < 1 || mol_max > 39) return fail(11, "tips_2003: mol_max out of range");
    double qt = 0., qt_296 = 0., qt_temp = 0.;
    for (int64_t mol = 1; mol <= mol_max; mol++) {
        int niso = TIPS_ISONM[mol - 1] < 9 ? TIPS_ISONM[mol - 1] : 9;         /* min(9,isonm(mol)), :60 */
        for (int iso = 1; iso <= niso; iso++) {
            for (int itemp = 1; itemp <= 2; itemp++) {
                double temp = (itemp == 1) ? 296. : temp_lbl;
                if (mol == 34) {
                    qt = 1.;                                                  /* not applicable to O; set to 1, :233-238 */
                } else if (mol == 39) {
                    /* :260-268: sets qt_296 / qt_temp, which :288-289 then overwrite with the stale QT */
                    if (itemp == 1) qt_296 = 296.;
                    if (itemp == 2) qt_temp = pow(temp / 296., 1.5);
                } else {
                    if (temp < 70. || temp > 3000.) qt = -1.;                 /* every QT_* routine */
                    else qt = tips_atob(temp, TIPS_TDAT, TIPS_QOFT + (size_t)(TIPS_QOFFSET[mol - 1] + iso - 1) * 119, 119);
                }
                if (qt <= 0.) return fail(11, "tips_2003: partition sum less than 0.");
                if (itemp == 1) qt_296 = qt;
                if (itemp == 2) qt_temp = qt;
            }
            scor[(mol - 1) + (size_t)(iso - 1) * 42] = qt_296 / qt_temp;
        }
    }
    return 0;
}
