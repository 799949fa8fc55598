// mrtm_kernels.cuh -- the sm_100a kernels of the hot path.
//   layer_prep_kernel   per-(profile,layer) scalars with the reference's evaluation order
//   continuum_kernel    MT_CKD (V2<820 cm-1 subset) onto the 1 cm-1 ABSRB grid per layer
//   derive_kernel       per-(line,layer) derived parameters (shifted centre, widths, STILD ...)
//   lines_kernel        the line-by-line accumulate + fused continuum/cloud/RFT epilogue
//   colsum_kernel       layer sums of per-molecule optical depths (STOREOUT's columns)
//   rt_kernel           CALCTMR + RAD_UP_DN + RTM
#pragma once
#include <cuda_runtime.h>

#include "mrtm_device.cuh"

namespace mrtm {

// ---------------------------------------------------------------------------------------------
// static line planes (device pointers)
struct LinesDev {
    int32_t n, n_pad;
    const int32_t *mol, *iso, *xf, *cls, *sidx, *lcidx, *brdidx, *segidx;
    const double *xnu0, *s0adj, *e, *alpf, *alps, *x, *deltnu, *sdep, *mass;
    const unsigned long long* key;
    const unsigned long long* keypre;   // [n_pad+1]
    const double* lc;          // [nlc][16]
    const int32_t* lc_self;    // [nlc]
    const double* brd;         // [nbrd][28]
    int32_t nsi;               // compact scor slots
    const int32_t* scor_index; // [nsi] -> (mol-1)+(iso-1)*42
    const int32_t* slot_mol;   // [nslot] molecule of each compact slot (ascending)
};

struct ContTablesDev {
    const double *sh2o_296, *sh2o_260, *fh2o, *fco2, *n2_296, *n2_296_sf, *n2_220, *n2_220_sf;
    const double *xfac_rhu, *co2_tdep;
};

struct TipsDev {
    const double* qoft;      // [rows][119]
    const double* tdat;      // [119]
    const int32_t* row;      // [nsi] row in qoft for compact slot, or -1
};

// =============================================================================================
// layer_prep_kernel: one thread per (profile,layer).  INITI (modm.f90:868-883), the layer part
// of LINES (:302-313) and the scalar part of CONTNM (contnm.f90:222-240,300-302,334,487,919).
// Only + - * / appear, evaluated with non-contracted IEEE operations so that the shift ratio
// Xn/XN0 -- which decides the selected line set -- is bit-identical to the reference's.
// =============================================================================================
struct LayerPrepArgs {
    int64_t nlayers;          // nprof*nlay
    int64_t nlay;
    int32_t nmol, ibrd;
    const double *p, *t, *clw, *wkl, *wbrodl;   // (nlay,nprof), wkl (39,nlay,nprof)
    double cntnm[7];
    double max_abs_deltnu, max_abs_brd_dshift;
    LayerDev* out;
    // scor: either gathered from a full (42,9,L) device array or computed from TIPS tables
    const double* scor_full;   // may be null
    TipsDev tips;
    int32_t nsi;
    const int32_t* scor_index;
    double* scorc;             // [L][nsi]
    int* errflag;              // bit0: TIPS range/partition-sum failure
    unsigned long long* sm_max_bits;   // max over the batch of shift_margin (bits of a non-negative double)
    // upper bound of 100*HWHM_D/|Xnu| per segment over the layers of the batch (bits of a non-negative double): the
    // plans need it before derive_kernel has run (the plan kernels overlap it on a second stream)
    unsigned long long* vtmax;         // [nseg]
    const Segment* seg;
    int32_t nseg, pad3;
};

// AtoB, tips_2003.f90:4610-4700 (4-point Lagrange, 3-point at the table ends)
__device__ inline double tips_atob(double aa, const double* A, const double* B, int npt)
{
    double bb = 0.;
    for (int I = 2; I <= npt; I++) {
        if (A[I - 1] >= aa) {
            if (I < 3 || I == npt) {
                int J = I;
                if (I < 3) J = 3;
                if (I == npt) J = npt;
                double a0d1 = xsub(A[J - 3], A[J - 2]); if (a0d1 == 0.) a0d1 = 0.0001;
                double a0d2 = xsub(A[J - 3], A[J - 1]); if (a0d2 == 0.) a0d2 = 0.0001;
                double a1d1 = xsub(A[J - 2], A[J - 3]); if (a1d1 == 0.) a1d1 = 0.0001;
                double a1d2 = xsub(A[J - 2], A[J - 1]); if (a1d2 == 0.) a1d2 = 0.0001;
                double a2d1 = xsub(A[J - 1], A[J - 3]); if (a2d1 == 0.) a2d1 = 0.0001;
                double a2d2 = xsub(A[J - 1], A[J - 2]); if (a2d2 == 0.) a2d2 = 0.0001;
                double a0 = xdiv(xmul(xsub(aa, A[J - 2]), xsub(aa, A[J - 1])), xmul(a0d1, a0d2));
                double a1 = xdiv(xmul(xsub(aa, A[J - 3]), xsub(aa, A[J - 1])), xmul(a1d1, a1d2));
                double a2 = xdiv(xmul(xsub(aa, A[J - 3]), xsub(aa, A[J - 2])), xmul(a2d1, a2d2));
                bb = xadd(xadd(xmul(a0, B[J - 3]), xmul(a1, B[J - 2])), xmul(a2, B[J - 1]));
            } else {
                int J = I;
                double a0d1 = xsub(A[J - 3], A[J - 2]); if (a0d1 == 0.) a0d1 = 0.0001;
                double a0d2 = xsub(A[J - 3], A[J - 1]); if (a0d2 == 0.) a0d2 = 0.0001;
                double a0d3 = xsub(A[J - 3], A[J]);     if (a0d3 == 0.) a0d3 = 0.0001;
                double a1d1 = xsub(A[J - 2], A[J - 3]); if (a1d1 == 0.) a1d1 = 0.0001;
                double a1d2 = xsub(A[J - 2], A[J - 1]); if (a1d2 == 0.) a1d2 = 0.0001;
                double a1d3 = xsub(A[J - 2], A[J]);     if (a1d3 == 0.) a1d3 = 0.0001;
                double a2d1 = xsub(A[J - 1], A[J - 3]); if (a2d1 == 0.) a2d1 = 0.0001;
                double a2d2 = xsub(A[J - 1], A[J - 2]); if (a2d2 == 0.) a2d2 = 0.0001;
                double a2d3 = xsub(A[J - 1], A[J]);     if (a2d3 == 0.) a2d3 = 0.0001;
                double a3d1 = xsub(A[J], A[J - 3]);     if (a3d1 == 0.) a3d1 = 0.0001;
                double a3d2 = xsub(A[J], A[J - 2]);     if (a3d2 == 0.) a3d2 = 0.0001;
                double a3d3 = xsub(A[J], A[J - 1]);     if (a3d3 == 0.) a3d3 = 0.0001;
                double a0 = xmul(xmul(xsub(aa, A[J - 2]), xsub(aa, A[J - 1])), xsub(aa, A[J]));
                a0 = xdiv(a0, xmul(xmul(a0d1, a0d2), a0d3));
                double a1 = xmul(xmul(xsub(aa, A[J - 3]), xsub(aa, A[J - 1])), xsub(aa, A[J]));
                a1 = xdiv(a1, xmul(xmul(a1d1, a1d2), a1d3));
                double a2 = xmul(xmul(xsub(aa, A[J - 3]), xsub(aa, A[J - 2])), xsub(aa, A[J]));
                a2 = xdiv(a2, xmul(xmul(a2d1, a2d2), a2d3));
                double a3 = xmul(xmul(xsub(aa, A[J - 3]), xsub(aa, A[J - 2])), xsub(aa, A[J - 1]));
                a3 = xdiv(a3, xmul(xmul(a3d1, a3d2), a3d3));
                bb = xadd(xadd(xadd(xmul(a0, B[J - 3]), xmul(a1, B[J - 2])), xmul(a2, B[J - 1])), xmul(a3, B[J]));
            }
            break;
        }
    }
    return bb;
}

__global__ void layer_prep_kernel(LayerPrepArgs a)
{
    int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= a.nlayers) return;
    const double* wk = a.wkl + (size_t)L * MRTM_MXMOL;
    const double pp = a.p[L], tt = a.t[L], wbrod = a.wbrodl[L];
    LayerDev o;
    // INITI
    o.radct = xdiv(xmul(kPLANCK, kCLIGHT), kBOLTZ);
    double xn0 = xmul(xdiv(kP0, xmul(kBOLTZ, kT0)), 1.E+3);
    double xn = xmul(xdiv(pp, xmul(kBOLTZ, tt)), 1.E+3);
    // LINES prologue
    double wtot = 0.;
    for (int m = 0; m < a.nmol; m++) wtot = xadd(wtot, wk[m]);
    wtot = xadd(wtot, wbrod);
    o.wtot = wtot;
    o.t = tt;
    o.p = pp;
    o.rp = xdiv(pp, kP0);
    o.rp2 = xmul(o.rp, o.rp);
    const double templc[4] = {200.0, 250.0, 296.0, 340.0};
    int ilc = 1;
    for (int il = 1; il <= 3; il++) {
        ilc = il;
        if (tt < templc[ilc]) break;
    }
    o.ilc = ilc;
    o.rectlc = xdiv(1.0, xsub(templc[ilc], templc[ilc - 1]));
    o.tmpdif = xsub(tt, templc[ilc - 1]);
    o.rt = xdiv(tt, kT0);
    o.lnrt = log(o.rt);
    o.dinvt = 1. / kT0 - 1. / tt;
    o.rhorat = xdiv(xn, xn0);
    for (int k = 0; k < 7; k++) o.rho_molec[k] = xdiv(xmul(o.rhorat, wk[k]), wtot);
    for (int m = 0; m < MRTM_MXMOL; m++) {
        o.wk[m] = (m < a.nmol) ? wk[m] : 0.;
        // rho_molec(mol) for mol>7 is out of bounds in the reference (modm.f90:845); natural extension
        o.rho_self[m] = (m < 7) ? o.rho_molec[m] : ((m < a.nmol) ? xdiv(xmul(o.rhorat, wk[m]), wtot) : 0.);
    }
    o.xkt = xdiv(tt, kRADCN2);
    o.clw = a.clw[L];
    o.sqrt_t = sqrt(tt);
    double sm = a.max_abs_deltnu * fabs(o.rhorat) * (1. + 1e-9) + 1e-12;
    if (a.ibrd != 0) {
        double sr = 0.;
        for (int k = 0; k < 7; k++) sr += fabs(o.rho_molec[k]);
        sm += sr * a.max_abs_brd_dshift * (1. + 1e-9);
    }
    o.shift_margin = sm;
    if (a.sm_max_bits) atomicMax(a.sm_max_bits, (unsigned long long)__double_as_longlong(sm));   // sm >= 0
    if (a.vtmax)
        for (int s = 0; s < a.nseg; s++) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(a.seg[s].vrate * o.sqrt_t);
            if (bits > *(volatile unsigned long long*)(a.vtmax + s)) atomicMax(a.vtmax + s, bits);
        }
    // CONTNM scalars (P0=1013, T0=296 there: contnm.f90:86)
    {
        const double cp0 = 1013., ct0 = 296., xlosmt = 2.68675E+19;
        double rhoave = xmul(xdiv(pp, cp0), xdiv(ct0, tt));
        double amagat = xmul(xdiv(pp, cp0), xdiv(273., tt));
        double cw = wbrod;
        for (int m = 0; m < a.nmol; m++) cw = xadd(cw, wk[m]);
        double wk1 = wk[0];
        double wk2 = (a.nmol >= 2) ? wk[1] : 0.;
        double wk7 = (a.nmol >= 7) ? wk[6] : 0.;
        double xh2o = xdiv(wk1, cw), xo2 = xdiv(wk7, cw);
        double xn2 = xsub(xsub(1., xh2o), xo2);
        double wn2 = xmul(xn2, cw);
        double h2o_fac = xdiv(wk1, cw);
        o.c_wk1 = wk1;
        o.c_rself = xmul(xmul(xmul(h2o_fac, rhoave), 1.e-20), a.cntnm[0]);
        o.c_rfrgn = xmul(xmul(xmul(xsub(1., h2o_fac), rhoave), 1.e-20), a.cntnm[1]);
        o.c_tfac_h2o = xdiv(xsub(tt, ct0), xsub(260., ct0));
        o.c_wco2 = xmul(xmul(xmul(wk2, rhoave), 1.0E-20), a.cntnm[2]);
        o.c_trat = xdiv(tt, 246.);
        o.c_taufac = xmul(xmul(a.cntnm[5], xdiv(wn2, xlosmt)), amagat);
        o.c_tfac_n2 = xdiv(xsub(tt, 296.), xsub(220., 296.));
        o.c_xn2 = xn2;
        o.c_xo2 = xo2;
        o.c_xh2o = xh2o;
    }
    o.pad = 0;
    a.out[L] = o;

    // scor for the compact (molecule,isotopologue) list
    for (int s = 0; s < a.nsi; s++) {
        double v;
        if (a.scor_full) {
            v = a.scor_full[(size_t)L * (MRTM_NSCOR1 * MRTM_NSCOR2) + a.scor_index[s]];
        } else {
            int row = a.tips.row[s];
            if (row < 0 || tt < 70. || tt > 3000.) {
                atomicOr(a.errflag, 1);
                v = 1.;
            } else {
                double q296 = tips_atob(296., a.tips.tdat, a.tips.qoft + (size_t)row * 119, 119);
                double qt = tips_atob(tt, a.tips.tdat, a.tips.qoft + (size_t)row * 119, 119);
                if (!(qt > 0.) || !(q296 > 0.)) atomicOr(a.errflag, 1);
                v = xdiv(q296, qt);
            }
        }
        a.scorc[(size_t)L * a.nsi + s] = v;
    }
}

// =============================================================================================
// continuum_kernel: one CTA per (profile,layer).  MT_CKD_3.5 branches that fire for V2 < 820:
// H2O self (contnm.f90:325-371), H2O foreign (:380-457), CO2 (:484-528), N2 roto-translational
// CIA (:906-943), each 4-point interpolated (XINT, lblrtm_sub.f90:1-34) onto the 1 cm-1 grid.
// Output planes: absrb[L][3][nptabs_pad] for species selectors im = 1 (H2O), 2 (CO2), 22 (N2).
// =============================================================================================
struct ContArgs {
    int64_t nlayers;
    ContGrid g[4];            // 0 self, 1 foreign, 2 co2, 3 n2
    double v1abs;
    int32_t nptabs, nptabs_pad;
    ContTablesDev tb;
    const LayerDev* lay;
    double* absrb;
};

__device__ __forceinline__ double xint_point(const double* a, double v1a, double dva, double vi)
{
    // body of the XINT loop (lblrtm_sub.f90:20-31); a is 0-based with a[j-1] = A(J)
    const double onepl = 1.001;
    double recdva = 1. / dva;
    int j = (int)((vi - v1a) * recdva + onepl);
    double vj = v1a + dva * (double)(j - 1);
    double p = recdva * (vi - vj);
    double c = (3. - 2. * p) * p * p;
    double b = 0.5 * p * (1. - p);
    double b1 = b * (1. - p);
    double b2 = b * p;
    return -a[j - 2] * b1 + a[j - 1] * (1. - c + b2) + a[j] * (c + b1) - a[j + 1] * b2;
}

__global__ void __launch_bounds__(128) continuum_kernel(ContArgs a)
{
    extern __shared__ double sm[];
    const int64_t L = blockIdx.x;
    const LayerDev& ly = a.lay[L];
    double* s_self = sm;
    double* s_frgn = s_self + a.g[0].nptc;
    double* s_co2 = s_frgn + a.g[1].nptc;
    double* s_n2 = s_co2 + a.g[2].nptc;
    const int tid = threadIdx.x;

    if (a.g[0].active) {
        for (int j = tid; j < a.g[0].nptc; j += blockDim.x) {
            int i = a.g[0].i1 + j;
            double s0 = 0., s1 = 0., sh2o = 0.;
            if (i >= 1 && i <= 2003) { s0 = a.tb.sh2o_296[i - 1]; s1 = a.tb.sh2o_260[i - 1]; }
            if (s0 > 0.) sh2o = s0 * pow(s1 / s0, ly.c_tfac_h2o);
            s_self[j] = ly.c_wk1 * (sh2o * ly.c_rself);
        }
    }
    if (a.g[1].active) {
        const double f0 = 0.06, v0f1 = 255.67, hwsq1 = 240. * 240., beta1 = 57.83, c_1 = -0.42, c_2 = 0.3, beta2 = 630.;
        for (int j = tid; j < a.g[1].nptc; j += blockDim.x) {
            int i = a.g[1].i1 + j;
            double f = (i >= 1 && i <= 2003) ? a.tb.fh2o[i - 1] : 0.;
            double vj = a.g[1].v1c + a.g[1].dvc * (double)j;
            double fscal;
            if (vj <= 600.) {
                int jfac = (int)((vj + 10.) / 10. + 0.00001);
                fscal = a.tb.xfac_rhu[jfac + 1];
            } else {
                double t1 = (vj - v0f1) / beta1, t2 = (vj + v0f1) / beta1, t3 = vj / beta2;
                double vf1 = t1 * t1; vf1 *= vf1; vf1 *= vf1;
                double vmf1 = t2 * t2; vmf1 *= vmf1; vmf1 *= vmf1;
                double vf2 = t3 * t3; vf2 *= vf2; vf2 *= vf2;
                fscal = 1. + (f0 + c_1 * ((hwsq1 / ((vj - v0f1) * (vj - v0f1) + hwsq1 + vf1)) +
                                          (hwsq1 / ((vj + v0f1) * (vj + v0f1) + hwsq1 + vmf1)))) /
                                 (1. + c_2 * vf2);
            }
            f = f * fscal;
            s_frgn[j] = (ly.c_wk1 * f) * ly.c_rfrgn;
        }
    }
    if (a.g[2].active) {
        for (int j = tid; j < a.g[2].nptc; j += blockDim.x) {
            int i = a.g[2].i1 + j;
            double f = 0.;
            if (i >= 1 && i <= 5003) {
                double tcor = 1.;
                if (i >= 1196 && i <= 1220) tcor = pow(ly.c_trat, a.tb.co2_tdep[i - 1196]);
                f = tcor * a.tb.fco2[i - 1];
            }
            s_co2[j] = f * ly.c_wco2;
        }
    }
    if (a.g[3].active) {
        for (int j = tid; j < a.g[3].nptc; j += blockDim.x) {
            int i = a.g[3].i1 + j;
            double c0 = 0., c1 = 0.;
            if (i >= 1 && i <= 73) {
                c0 = a.tb.n2_296[i - 1] * pow(a.tb.n2_220[i - 1] / a.tb.n2_296[i - 1], ly.c_tfac_n2);
                double sf_t = a.tb.n2_296_sf[i - 1] * pow(a.tb.n2_220_sf[i - 1] / a.tb.n2_296_sf[i - 1], ly.c_tfac_n2);
                c1 = (sf_t - 1.) * 0.79 / 0.21;
            }
            s_n2[j] = ly.c_taufac * c0 * (ly.c_xn2 + c1 * ly.c_xo2 + 1. * ly.c_xh2o);
        }
    }
    __syncthreads();
    double* out = a.absrb + (size_t)L * 3 * a.nptabs_pad;
    for (int i = 1 + tid; i <= a.nptabs_pad; i += blockDim.x) {
        double vi = a.v1abs + 1.0 * (double)(i - 1);
        double h = 0., c = 0., n = 0.;
        if (i <= a.nptabs) {
            if (a.g[0].active && i >= a.g[0].ilo && i <= a.g[0].ihi) h = h + xint_point(s_self, a.g[0].v1c, a.g[0].dvc, vi);
            if (a.g[1].active && i >= a.g[1].ilo && i <= a.g[1].ihi) h = h + xint_point(s_frgn, a.g[1].v1c, a.g[1].dvc, vi);
            if (a.g[2].active && i >= a.g[2].ilo && i <= a.g[2].ihi) c = xint_point(s_co2, a.g[2].v1c, a.g[2].dvc, vi);
            if (a.g[3].active && i >= a.g[3].ilo && i <= a.g[3].ihi) n = xint_point(s_n2, a.g[3].v1c, a.g[3].dvc, vi);
        }
        out[i - 1] = h;
        out[a.nptabs_pad + i - 1] = c;
        out[2 * a.nptabs_pad + i - 1] = n;
    }
}

// =============================================================================================
// derive_kernel: one thread per (line, layer).  Everything in LINES that does not depend on the
// frequency: coupling coefficients (modm.f90:328-368), shifted centre (:375-380,
// bit exact), INTENS (:860-865), HALFWHM_C (:833-857), HALFWHM_D (:442-454), zeta (:419).
// =============================================================================================
struct DeriveArgs {
    int64_t nlayers;
    LinesDev ln;
    const LayerDev* lay;
    const double* scorc;      // [L][nsi]
    double sclcpl, sclhw, y0res;
    int32_t ibrd, pad;
    double* planes;           // [L][D_NPLANES][n_pad]
    unsigned long long* layer_voigt;   // [L] bits of the smallest |Xnu| among the lines of the layer that can take the Voigt branch
                                       // (zeta <= 0.99; 100*HWHM_D grows with |Xnu|), all ones = none
    int32_t nseg, pad2;
};

#ifndef MRTM_DERIVE_MINB
#define MRTM_DERIVE_MINB 6
#endif
// one (line, layer): returns the bits of |Xnu| when the line can take the Voigt branch in this layer, else all ones
__device__ __forceinline__ unsigned long long derive_one(const DeriveArgs& a, int q, int64_t L)
{
    const unsigned long long kNone = ~0ull;
    if (q >= a.ln.n_pad) return kNone;
    double* pl = a.planes + (size_t)L * D_NPLANES * a.ln.n_pad;
    if (q >= a.ln.n) {   // padding: far away, zero strength
        pl[(size_t)D_XNU * a.ln.n_pad + q] = 1.0e30;
        pl[(size_t)D_H2 * a.ln.n_pad + q] = 1.0;
        pl[(size_t)D_CN * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_P3 * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_P4 * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_H * a.ln.n_pad + q] = 1.0;
        pl[(size_t)D_AD * a.ln.n_pad + q] = 1.0;
        pl[(size_t)D_VT * a.ln.n_pad + q] = -1.0;
        pl[(size_t)D_STILD * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_AIP * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_BIP * a.ln.n_pad + q] = 0.;
        return kNone;
    }
    const LayerDev& ly = a.lay[L];
    const int mol = a.ln.mol[q], xf = a.ln.xf[q], cls = a.ln.cls[q];
    const double rhorat = ly.rhorat, rho_self = ly.rho_self[mol - 1];
    const double radct = ly.radct, t = ly.t;

    double aip = 0., bip = 0.;
    const int lci = a.ln.lcidx[q];
    if (lci >= 0) {
        const double* c = a.ln.lc + (size_t)lci * 16;
        double A[4] = {c[0], c[1], c[2], c[3]}, B[4] = {c[4], c[5], c[6], c[7]};
        if (a.ln.lc_self[lci]) {
            double rho_for = (rhorat - rho_self) / rhorat;
            double rho_sel = rho_self / rhorat;
            for (int k = 0; k < 4; k++) {
                A[k] = xadd(xmul(rho_for, A[k]), xmul(rho_sel, c[8 + k]));
                B[k] = xadd(xmul(rho_for, B[k]), xmul(rho_sel, c[12 + k]));
            }
        }
        const int ilc = ly.ilc;
        aip = A[ilc - 1] + ((A[ilc] - A[ilc - 1]) * ly.rectlc) * ly.tmpdif;
        bip = B[ilc - 1] + ((B[ilc] - B[ilc - 1]) * ly.rectlc) * ly.tmpdif;
    }
    if (xf == -1) {
        aip = aip * a.sclcpl + a.y0res;
        bip = bip * a.sclcpl + a.y0res;
    }
    if (xf == -3) {
        aip = aip * a.sclhw;
        bip = bip * a.sclhw;
    }

    // shifted line centre: exactly Xnu0 + deltnu*(Xn/XN0) [+ sum(rho*flg*(shft-deltnu))], no FMA
    const double xnu0 = a.ln.xnu0[q], deltnu = a.ln.deltnu[q];
    double xnu = xadd(xnu0, xmul(deltnu, rhorat));
    const int bi = a.ln.brdidx[q];
    const bool use_brd = (mol <= MRTM_MXBRDMOL) && (a.ibrd != 0);
    const double* brd = (bi >= 0) ? a.ln.brd + (size_t)bi * 28 : nullptr;
    if (use_brd) {
        double s = 0.;
        if (brd)
            for (int k = 0; k < 7; k++) s = xadd(s, xmul(xmul(ly.rho_molec[k], brd[k]), xsub(brd[21 + k], deltnu)));
        xnu = xadd(xnu, s);
    }

    // INTENS
    const double xipsf = a.scorc[(size_t)L * a.ln.nsi + a.ln.sidx[q]];
    const double es = a.ln.e[q];
    // exp(-c2 E/T)/exp(-c2 E/T0) as one exponential (modm.f90 INTENS)
    double s = a.ln.s0adj[q] * exp(radct * es * ly.dinvt) * xipsf;
    double stild = s * ((1 + exp(-(radct * xnu / t))) / (xnu * (1 - exp(-(radct * xnu / kT0)))));

    // HALFWHM_C
    const double af = a.ln.alpf[q], as = a.ln.alps[q];
    const double rtx = exp(a.ln.x[q] * ly.lnrt);         // (T/T0)^x with the layer's log(T/T0)
    const double alfa0i = af * rtx, hwhmsi = as * rtx;
    double hwhm_c = alfa0i * (rhorat - rho_self) + hwhmsi * rho_self;
    if (use_brd && brd) {
        double alfsum = 0., sflgrho = 0.;
        for (int k = 0; k < 7; k++) {
            double tmpcor = pow(ly.rt, brd[14 + k]);
            alfsum = alfsum + ly.rho_molec[k] * brd[k] * (brd[7 + k] * tmpcor);
            sflgrho = sflgrho + ly.rho_molec[k] * brd[k];
        }
        hwhm_c = (rhorat - sflgrho) * alfa0i + alfsum;
        if (brd[mol - 1] == 0.) hwhm_c = hwhm_c + rho_self * (hwhmsi - alfa0i);
    }
    // HALFWHM_D
    const double hwhm_d = (xnu / kCLIGHT) * sqrt(2. * log(2.) * ((kBOLTZ * t) / (a.ln.mass[q] / kAVOGAD)));
    if (xf == -3) hwhm_c = hwhm_c * (1 - (aip * ly.rp) - (bip * ly.rp2));
    const double zeta = hwhm_c / (hwhm_c + hwhm_d);

    const double h2 = hwhm_c * hwhm_c;
    const double cn = stild * hwhm_c / kPI;
    double p3 = 0., p4 = 0.;
    if (cls == CLS_PED) p3 = cn / (kDELTNUC * kDELTNUC + h2);
    if (cls == CLS_O2_LC1) {
        p3 = cn * (1. + bip * ly.rp2);
        p4 = cn * (aip * (1 / hwhm_c) * ly.rp);
    }
    const size_t np = a.ln.n_pad;
    pl[(size_t)D_XNU * np + q] = xnu;
    pl[(size_t)D_H2 * np + q] = h2;
    pl[(size_t)D_CN * np + q] = cn;
    pl[(size_t)D_P3 * np + q] = p3;
    pl[(size_t)D_P4 * np + q] = p4;
    pl[(size_t)D_H * np + q] = hwhm_c;
    pl[(size_t)D_AD * np + q] = hwhm_d;
    const double vt = (zeta > 0.99) ? -1.0 : 100. * hwhm_d;
    pl[(size_t)D_VT * np + q] = vt;
    pl[(size_t)D_STILD * np + q] = stild;
    pl[(size_t)D_AIP * np + q] = aip;
    pl[(size_t)D_BIP * np + q] = bip;
    return (vt >= 0.) ? (unsigned long long)__double_as_longlong(fabs(xnu)) : kNone;
}

__global__ void __launch_bounds__(256, MRTM_DERIVE_MINB) derive_kernel(DeriveArgs a)
{
    const int64_t L = blockIdx.y;
    unsigned long long xb = derive_one(a, blockIdx.x * blockDim.x + threadIdx.x, L);
    // smallest |Xnu| of the layer's Voigt-capable lines: warp minimum, block minimum, one global atomic per block at most
    __shared__ unsigned long long s_min[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, xb, off);
        xb = o < xb ? o : xb;
    }
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = xb;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = s_min[0];
        for (int w = 1; w < 8; w++) m = s_min[w] < m ? s_min[w] : m;
        if (m < *(volatile unsigned long long*)(a.layer_voigt + L)) atomicMin(a.layer_voigt + L, m);
    }
}

// =============================================================================================
// The line path (modm.f90:277-440).  plan_kernel classifies every (frequency tile, segment) once per call,
// far_kernel / far_warp_kernel expand the far lines level by level, near2_kernel (near_kernel for oversize
// tiles) evaluates the remaining (line, layer, frequency) triples with the reference's exact tests,
// voigt_kernel adds the Voigt-branch pairs and final_kernel closes the sums (RFT :257, continuum
// interpolation + RADFN :218-230, cloud liquid water :264, total :265-269).  LinesArgs is shared by them.
// =============================================================================================
struct SegWork;
struct TileHdr;
struct NearPiece;
constexpr int kMaxLevels = 4;   // far-field hierarchy: level 0 = the line kernel's own tiles
struct LinesArgs {
    int32_t nwn, nlay;            // frequencies in this call/chunk, layers per profile
    int32_t nseg, n_pad;
    int64_t iw0;                  // global 0-based index of wn[0] (gridded continuum interpolation)
    const double* wn;             // [nwn]
    const Segment* seg;           // [nseg]
    const double* xnu0;           // static centres (sorted inside segments)
    const int32_t *mol_s, *xf_s;  // static per line
    const double* sdep_s;
    const unsigned long long* key;
    const unsigned long long* keypre;   // [n_pad+1] prefix sums of key (selection hash of a whole range)
    double ff_ratio;              // far-field expansion: poles >= ff_ratio tile half-widths away; 0 = direct only
    double ffw_ratio;             // the same ratio for near2_kernel's in-warp expansion about the warp's own block
    unsigned long long* counters; // [2] far-field expansions, direct (line,frequency) evaluations (may be null)
    const double* planes;         // [L][D_NPLANES][n_pad]
    const LayerDev* lay;          // [L]
    // far-field hierarchy: level 0 = this kernel's tiles, level lv tiles are S^lv times wider
    int32_t nlev, S;
    const SegWork* plan[kMaxLevels];    // [ntiles_lv][nseg]
    const TileHdr* hdr[kMaxLevels];     // [ntiles_lv]
    const double* coef[kMaxLevels];     // lv >= 1: [tile][L][slot][kFarK] from far_kernel
    int32_t nslot, pad0;
    const NearPiece* near_pieces;       // [ntiles][kMaxNearPieces] (plan_kernel, level 0); null: near2_kernel not in use
    const unsigned long long* layer_voigt;   // [L] bits of the smallest |Xnu| of the layer's Voigt-capable lines (derive_kernel), all ones = none
    const int32_t* slot_mol;            // [nslot]
    // continuum
    const double* absrb;          // [L][3][nptabs_pad]
    int32_t nptabs, nptabs_pad;
    double v1abs, v2abs, v1, dvset;
    // outputs (any may be null).  Strides in elements.
    double* o;        int64_t o_lds;  int64_t o_prof;     // o[iw + k*o_lds + prof*o_prof]
    double* o_v;                  // zeroed scratch with o's strides: voigt_kernel adds there (it then runs beside the near field), or null
    double* o_by_mol; int64_t obm_ldm; int64_t obm_ldk;   // [iw + (mol-1)*ldm + k*ldk] (+prof*ldk*nlay)
    double* oc;                                           // same strides as o_by_mol
    double* o_clw;                                        // same strides as o
    const double* odxsec;                                 // same strides as o (input, may be null)
    long long* sel_count; unsigned long long* sel_hash;   // same strides as o
    int* errflag;                                         // bit1: SDVOIGT negative real part
};

__device__ __forceinline__ int lower_bound_d(const double* a, int lo, int hi, double v)
{   // first index in [lo,hi) with a[i] >= v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* a, int lo, int hi, double v)
{   // first index in [lo,hi) with a[i] > v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Can a (line, frequency) pair of this layer take the Voigt branch for a frequency <= whi?  A pair needs
// |WN-Xnu| <= 100*HWHM_D <= 2e-3*|Xnu| (T < 3000 K, molecular mass >= 1), so Xnu <= 1.0021*whi; the layer's
// Voigt-capable lines all have |Xnu| >= the recorded minimum.
__device__ __forceinline__ bool voigt_possible(const unsigned long long* layer_voigt, int64_t L, double whi)
{
    const unsigned long long bound = (unsigned long long)__double_as_longlong(fabs(whi) * 1.01 + 1e-3);
    return bound >= layer_voigt[L];
}

// RADFN, lblrtm_sub.f90:36-97
__device__ __forceinline__ double radfn(double vi, double xkt)
{
    if (xkt > 0.0) {
        double x = vi / xkt;
        if (x <= 0.01) return 0.5 * x * vi;
        if (x <= 10.0) {
            double e = exp(-x);
            return vi * (1. - e) / (1. + e);
        }
        return vi;
    }
    return vi;
}

// ODCLW_TKC / Forward_TKC, CloudOptProp.f90:29-157 (binary64 here; the parity build evaluates
// the d0-literal expressions in binary128 and rounds, a ~1e-16 relative difference)
__device__ __noinline__ double odclw_tkc(double wn, double temp, double clw)
{
    const double a_1 = 8.110808E+01, b_1 = 4.433736E-03, c_1 = 1.301700E-13, d_1 = 6.627126E+02;
    const double a_2 = 2.025164E+00, b_2 = 1.072976E-02, c_2 = 1.011945E-14, d_2 = 6.089168E+02;
    const double t_c = 1.342433E+02;
    double freq = wn * kCLIGHT / 1.e9;
    double tc = temp - 273.15;
    double frq = freq * 1.e9;
    double cl = kCLIGHT / 100.;
    double eps_s = 87.9144 - 0.404399 * tc + 9.58726E-4 * (tc * tc) - 1.32802E-6 * (tc * tc * tc);
    double delta_1 = a_1 * exp(-b_1 * tc), tau_1 = c_1 * exp(d_1 / (tc + t_c));
    double delta_2 = a_2 * exp(-b_2 * tc), tau_2 = c_2 * exp(d_2 / (tc + t_c));
    double w = 2. * kPI * frq;
    double den1 = 1. + (w * tau_1) * (w * tau_1), den2 = 1. + (w * tau_2) * (w * tau_2);
    double eps1 = eps_s - (w * w) * ((tau_1 * tau_1 * delta_1) / den1 + (tau_2 * tau_2 * delta_2) / den2);
    double eps2 = w * ((tau_1 * delta_1) / den1 + (tau_2 * delta_2) / den2);
    cplx e = cmk(eps1, eps2);
    cplx re = (cmk(eps1 - 1., eps2)) / (cmk(eps1 + 2., eps2));
    (void)e;
    double alpha = 6. * kPI * re.im * frq * 1.e-3 / cl;
    return alpha * clw;
}

// ---- TMA (bulk async copy) + mbarrier primitives used to stream line-parameter tiles ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#ifndef MRTM_LINES_MINB
#define MRTM_LINES_MINB 4
#endif
#ifndef MRTM_UNROLL_BOTH
#define MRTM_UNROLL_BOTH 2
#endif
#define MRTM_PRAGMA(x) _Pragma(#x)
#define MRTM_UNROLL(n) MRTM_PRAGMA(unroll n)
constexpr int kTile = 128;      // lines per smem tile
constexpr int kStages = 8;      // tile ring
constexpr int kPrefetch = 5;    // TMA jobs in flight ahead of the consumer; a warp may run kStages-kPrefetch tiles ahead of the slowest
#ifndef MRTM_FARK
#define MRTM_FARK 14
#endif
constexpr int kFarK = MRTM_FARK; // Taylor terms of the far-field expansion (degree kFarK-1); <= 16 (reduce_coefs)
static_assert(kFarK >= 4 && kFarK <= 16, "kFarK out of range");
constexpr int kMaxBp = 12;      // break points per segment
constexpr int kMaxRun = 6;      // direct runs per segment

// sub-range mode bits
constexpr int M_EDGE = 1;       // per-(line,frequency) window test |WN-Xnu| > 25 (modm.f90:384)
constexpr int M_NEG = 2;        // per-(line,frequency) test WN+Xnu <= 25 (modm.f90:746)
constexpr int M_VOIGT = 4;      // per-(line,frequency) test |WN-Xnu| <= 100*HWHM_D (modm.f90:427)
constexpr int M_NEAR = 8;       // direct evaluation (a pole of the line is too close to the tile to expand)

// Classification of one (molecule, class) segment against one frequency tile.  It does not depend on
// the layer: the margins are maxima over the layers of the batch (shift margin, 100*HWHM_D), so one
// plan serves every layer and every profile of a call.
struct SegWork {
    // searched fields, in the order of the plan tasks (kept contiguous)
    int q0, q1;        // lines that can be inside the 25 cm-1 window of some frequency of the tile
    int eb, ec;        // [q0,eb) and [ec,q1): window-edge bands
    int n0, n1;        // [n0,n1): band where WN+Xnu<=25 flips; < n0: both resonances for every frequency
    int v0, v1;        // [v0,v1): Voigt zone
    int z0;            // < z0: the negative-frequency pole -Xnu is near the tile
    int f0, f1;        // [f0,f1): the pole +Xnu is near the tile
    // derived
    int nbp;                     // break points bp[0..nbp-1]; sub-range u = [bp[u], bp[u+1])
    int bp[kMaxBp];
    unsigned char mode[kMaxBp];  // mode bits of sub-range u; 0 = far field (Taylor expansion)
    int nrun;                    // maximal runs of consecutive direct (mode != 0) sub-ranges
    int run_lo[kMaxRun], run_hi[kMaxRun], run_t0[kMaxRun], run_nt[kMaxRun];
    int run_off[kMaxRun];        // near_kernel (stage-all mode): offset of the run in the CTA's staging area
    int run_u0[kMaxRun], run_u1[kMaxRun];   // sub-ranges [u0,u1) that make up the run
    int tma;                     // class streams its direct runs through shared memory
    int has_far;
};
constexpr int kSegTasks = 11;
struct TileHdr {
    double wlo, whi;             // frequency extent of the tile
    int total_lines;             // lines (padded to 4 per run) of all direct runs of the streamed classes
    int npieces, nunits;         // far-field work list of the tile: pieces, and work units of the non-mixing pieces
    int nterms;                  // Taylor expansions of the non-mixing pieces (statistics)
    int nnear, pad;              // direct sub-ranges of the streamed classes (near2_kernel's piece list); -1 = too many
};
// One direct sub-range of a streamed (TMA) class in the CTA's staging area (near2_kernel)
struct NearPiece {
    int soff;                    // staged offset of its first line
    int n;                       // lines
    int q0;                      // index (staged order of the line list) of its first line
    int info;                    // segment | mode << 8 | negall << 16 | kind << 17 (0 PED, 1 O2, 2 O2_LC35)
};
constexpr int kMaxNearPieces = 192;
#ifndef MRTM_NEAR2_CAP
#define MRTM_NEAR2_CAP 640
#endif
#ifndef MRTM_NEAR2_MINB
#define MRTM_NEAR2_MINB 5
#endif
#ifndef MRTM_FAR_MINB
#define MRTM_FAR_MINB 6
#endif
constexpr int kNearCap = MRTM_NEAR2_CAP;        // staged lines per CTA of near2_kernel (tiles with more go to near_kernel)
static_assert(kNearCap <= kStages * kTile && kNearCap % 32 == 0, "kNearCap");
// One contiguous range of lines that far_kernel expands for a tile: far at this level and not at the parent level.
constexpr int kPiecePerSeg = 12;
struct FarPiece {
    int lo, n;                   // lines [lo, lo+n)
    int off;                     // first work unit of the piece in the tile's unit numbering (non-mixing pieces); a unit is
                                 // one line with both resonances, or two adjacent single-resonance lines
    int info;                    // segment | both << 16 | mix << 17
};

// the far (mode 0) sub-ranges of a segment at this level, minus the ones the parent level already
// expanded (the parent's far set is a subset of the child's by construction of the margins)
template <class Fn>
__device__ __forceinline__ void for_each_far_piece(const SegWork& wk, const SegWork* pk, Fn fn)
{
    for (int u = 0; u + 1 < wk.nbp; u++) {
        if (wk.mode[u] != 0) continue;
        const int lo = wk.bp[u], hi = wk.bp[u + 1];
        int cur = lo;
        if (pk) {
            for (int v = 0; v + 1 < pk->nbp && cur < hi; v++) {
                if (pk->mode[v] != 0) continue;
                const int pa = pk->bp[v], pb = pk->bp[v + 1];
                if (pb <= cur) continue;
                if (pa >= hi) break;
                if (pa > cur) fn(cur, pa, lo);
                cur = pb > cur ? pb : cur;
            }
        }
        if (cur < hi) fn(cur, hi, lo);
    }
}

// =============================================================================================
// plan_kernel: one CTA per frequency tile of one hierarchy level.  All window / band / near-zone
// searches of all segments run in parallel (one binary search per thread), then one thread per
// segment orders the break points, assigns the sub-range modes and the direct runs.
// =============================================================================================
struct PlanArgs {
    int32_t nwn, tile_freqs, nseg, pad;
    const double* wn;
    const Segment* seg;
    const double* xnu0;
    const unsigned long long* sm_max_bits;    // max shift margin over the layers of the batch (bits of a double)
    const unsigned long long* vtmax_seg;      // [nseg] max 100*HWHM_D/|Xnu| of Voigt-capable lines (bits), 0 = none
    double ff_ratio;
    SegWork* out;                             // [ntiles][nseg]
    TileHdr* hdr;                             // [ntiles]
    const SegWork* pplan;                     // parent level's plan (computed first) or null
    int32_t S, pad2;                          // tiles of this level per parent tile
    FarPiece* pieces;                         // [ntiles][nseg*kPiecePerSeg]
    NearPiece* near_pieces;                   // [ntiles][kMaxNearPieces] (level 0 only, may be null)
};

__global__ void __launch_bounds__(128) plan_kernel(PlanArgs a)
{
    constexpr int NT = 128;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    SegWork* s_work = reinterpret_cast<SegWork*>(s_dyn);
    __shared__ double s_lo[4], s_hi[4];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int i0 = tile * a.tile_freqs;
    const int i1 = min(i0 + a.tile_freqs, a.nwn);
    double wlo = 1e300, whi = -1e300;
    for (int i = i0 + tid; i < i1; i += NT) {
        const double w = a.wn[i];
        wlo = fmin(wlo, w);
        whi = fmax(whi, w);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        wlo = fmin(wlo, __shfl_xor_sync(0xffffffffu, wlo, off));
        whi = fmax(whi, __shfl_xor_sync(0xffffffffu, whi, off));
    }
    if ((tid & 31) == 0) { s_lo[tid >> 5] = wlo; s_hi[tid >> 5] = whi; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) { wlo = fmin(wlo, s_lo[i]); whi = fmax(whi, s_hi[i]); }
    if (tid == 0) { a.hdr[tile].wlo = wlo; a.hdr[tile].whi = whi; }
    const double sm = __longlong_as_double((long long)*a.sm_max_bits);
    const int nseg = a.nseg;
    const double cen = 0.5 * (wlo + whi), hh = 0.5 * (whi - wlo);
    const bool ff = a.ff_ratio > 0.;
    const double Rn = a.ff_ratio * hh;

    for (int task = tid; task < nseg * kSegTasks; task += NT) {
        const int s = task / kSegTasks, w = task - s * kSegTasks;
        const Segment sg = a.seg[s];
        const int cls = sg.cls;
        const bool tma_cls = (cls == CLS_PED) || (cls == CLS_O2) || (cls == CLS_O2_LC35);
        const bool exp_cls = tma_cls || (cls == CLS_O2_LC1);      // classes with a far-field path
        const bool has_win = (cls == CLS_PED) || (cls == CLS_O2) || (cls == CLS_GENERAL && sg.mol != 7);
        int r;
        switch (w) {
        case 0: r = has_win ? lower_bound_d(a.xnu0, sg.begin, sg.end, wlo - kDELTNUC - sm) : sg.begin; break;
        case 1: r = has_win ? upper_bound_d(a.xnu0, sg.begin, sg.end, whi + kDELTNUC + sm) : sg.end; break;
        case 2: r = (has_win && tma_cls) ? upper_bound_d(a.xnu0, sg.begin, sg.end, whi - kDELTNUC + sm) : sg.begin; break;
        case 3: r = (has_win && tma_cls) ? lower_bound_d(a.xnu0, sg.begin, sg.end, wlo + kDELTNUC - sm) : sg.end; break;
        case 4: r = (has_win && tma_cls) ? lower_bound_d(a.xnu0, sg.begin, sg.end, kDELTNUC - whi - sm) : sg.end; break;
        case 5: r = (has_win && tma_cls) ? upper_bound_d(a.xnu0, sg.begin, sg.end, kDELTNUC - wlo + sm + 1e-9) : sg.end; break;
        case 6:
        case 7: {
            // Voigt zone: where a frequency can come within max(100*HWHM_D) of a centre (modm.f90:427); the
            // maximum is over the lines of the segment that are not Lorentz-only (zeta <= 0.99) in some layer
            const unsigned long long vbits = a.vtmax_seg[s];
            if (exp_cls && vbits != 0ull) {
                // 100*HWHM_D <= rate*|Xnu| and a line of the zone has |Xnu| <= max|WN| + 1
                const double vb = __longlong_as_double((long long)vbits) * (fmax(fabs(wlo), fabs(whi)) + 1.0) * (1. + 1e-9) + sm + 1e-9;
                r = (w == 6) ? lower_bound_d(a.xnu0, sg.begin, sg.end, wlo - vb) : upper_bound_d(a.xnu0, sg.begin, sg.end, whi + vb);
            } else {
                r = sg.begin;      // empty zone after clipping
            }
        } break;
        case 8: r = (ff && exp_cls) ? lower_bound_d(a.xnu0, sg.begin, sg.end, Rn - cen + sm) : sg.end; break;
        case 9: r = (ff && exp_cls) ? lower_bound_d(a.xnu0, sg.begin, sg.end, cen - Rn - sm) : sg.begin; break;
        default: r = (ff && exp_cls) ? upper_bound_d(a.xnu0, sg.begin, sg.end, cen + Rn + sm) : sg.end; break;
        }
        (&s_work[s].q0)[w] = r;
    }
    __syncthreads();
    for (int s = tid; s < nseg; s += NT) {
        SegWork& wk = s_work[s];
        const Segment sg = a.seg[s];
        const int cls = sg.cls;
        const bool tma_cls = (cls == CLS_PED) || (cls == CLS_O2) || (cls == CLS_O2_LC35);
        const bool has_win = tma_cls && (cls != CLS_O2_LC35);
        const bool force_both = (cls == CLS_O2_LC35) || (cls == CLS_O2_LC1);
        const int q0 = wk.q0, q1 = wk.q1 > wk.q0 ? wk.q1 : wk.q0;
        wk.q1 = q1;
        int c[9] = {wk.eb, wk.ec, wk.n0, wk.n1, wk.v0, wk.v1, wk.z0, wk.f0, wk.f1};
        for (int i = 0; i < 9; i++) c[i] = c[i] < q0 ? q0 : (c[i] > q1 ? q1 : c[i]);
        wk.eb = c[0]; wk.ec = c[1]; wk.n0 = c[2]; wk.n1 = c[3]; wk.v0 = c[4]; wk.v1 = c[5];
        wk.z0 = c[6]; wk.f0 = c[7]; wk.f1 = c[8];
        for (int i = 1; i < 9; i++) { int v = c[i], j = i - 1; while (j >= 0 && c[j] > v) { c[j + 1] = c[j]; j--; } c[j + 1] = v; }
        int nbp = 0;
        wk.bp[nbp++] = q0;
        for (int i = 0; i < 9; i++) if (c[i] > wk.bp[nbp - 1]) wk.bp[nbp++] = c[i];
        if (q1 > wk.bp[nbp - 1]) wk.bp[nbp++] = q1;
        wk.nbp = nbp;
        wk.tma = tma_cls ? 1 : 0;
        int nrun = 0, has_far = 0;
        bool open = false;
        for (int u = 0; u + 1 < nbp; u++) {
            const int x = wk.bp[u];
            int mode = 0;
            if (has_win && ((x < wk.eb) || (x >= wk.ec))) mode |= M_EDGE;
            if (has_win && (x >= wk.n0) && (x < wk.n1)) mode |= M_NEG;
            if ((x >= wk.v0) && (x < wk.v1)) mode |= M_VOIGT;
            const bool second = force_both || (has_win && x < wk.n1);      // the negative-frequency term can be present
            if (((x >= wk.f0) && (x < wk.f1)) || (second && x < wk.z0)) mode |= M_NEAR;
            if (!(tma_cls || cls == CLS_O2_LC1)) mode |= M_NEAR;            // CLS_GENERAL: always direct
            wk.mode[u] = (unsigned char)mode;
            if (mode != 0) {
                if (open) {
                    wk.run_hi[nrun - 1] = wk.bp[u + 1];
                    wk.run_u1[nrun - 1] = u + 1;
                } else {
                    wk.run_lo[nrun] = x;
                    wk.run_hi[nrun] = wk.bp[u + 1];
                    wk.run_u0[nrun] = u;
                    wk.run_u1[nrun] = u + 1;
                    nrun++;
                    open = true;
                }
            } else {
                has_far = 1;
                open = false;
            }
        }
        for (int r = 0; r < nrun; r++) {
            wk.run_t0[r] = wk.run_lo[r] & ~3;
            wk.run_nt[r] = (wk.run_hi[r] - wk.run_t0[r] + kTile - 1) / kTile;
            wk.run_off[r] = 0;
        }
        wk.nrun = nrun;
        wk.has_far = has_far;
    }
    __syncthreads();
    if (tid == 0) {      // staging offsets of the direct runs (near_kernel, stage-all mode)
        int tot = 0;
        for (int s = 0; s < nseg; s++) {
            SegWork& w = s_work[s];
            if (!w.tma) continue;
            for (int r = 0; r < w.nrun; r++) {
                w.run_off[r] = tot;
                tot += ((w.run_hi[r] - w.run_t0[r]) + 3) & ~3;
            }
        }
        a.hdr[tile].total_lines = tot;
        int nn = 0;
        if (a.near_pieces) {        // the same runs, sub-range by sub-range, in staging coordinates
            NearPiece* np = a.near_pieces + (size_t)tile * kMaxNearPieces;
            for (int s = 0; s < nseg; s++) {
                const SegWork& w = s_work[s];
                if (!w.tma) continue;
                const int cls = a.seg[s].cls;
                const int kind = (cls == CLS_PED) ? 0 : ((cls == CLS_O2) ? 1 : 2);
                for (int r = 0; r < w.nrun; r++)
                    for (int u = w.run_u0[r]; u < w.run_u1[r]; u++) {
                        const int x = w.bp[u], n = w.bp[u + 1] - x;
                        if (n <= 0) continue;
                        if (nn < kMaxNearPieces) {
                            const int negall = ((kind == 2) || (x < w.n0)) ? 1 : 0;
                            NearPiece pc;
                            pc.soff = w.run_off[r] + (x - w.run_t0[r]);
                            pc.n = n;
                            pc.q0 = x;
                            pc.info = s | ((int)w.mode[u] << 8) | (negall << 16) | (kind << 17);
                            np[nn] = pc;
                        }
                        nn++;
                    }
            }
            if (nn > kMaxNearPieces) nn = -1;
        }
        a.hdr[tile].nnear = nn;
        a.hdr[tile].pad = 0;
    }
    // far-field work list: per segment the far sub-ranges minus the parent's, then one term numbering per tile
    __shared__ int s_npc[kMaxSegments];
    FarPiece* s_pc = reinterpret_cast<FarPiece*>(s_work + nseg);      // [nseg][kPiecePerSeg]
    for (int s = tid; s < nseg; s += NT) {
        const SegWork& wk = s_work[s];
        const SegWork* pk = a.pplan ? a.pplan + (size_t)(tile / a.S) * nseg + s : nullptr;
        const int cls = a.seg[s].cls;
        const bool force_both = (cls == CLS_O2_LC35) || (cls == CLS_O2_LC1);
        const int mix = (cls == CLS_O2_LC1) ? 1 : 0;
        int n = 0;
        if (wk.has_far)
            for_each_far_piece(wk, pk, [&](int lo, int hi, int sub_lo) {
                if (n < kPiecePerSeg) {
                    FarPiece fp;
                    fp.lo = lo;
                    fp.n = hi - lo;
                    fp.off = 0;
                    fp.info = s | ((force_both || (sub_lo < wk.n0)) ? (1 << 16) : 0) | (mix << 17);
                    s_pc[s * kPiecePerSeg + n] = fp;
                }
                n++;
            });
        s_npc[s] = n < kPiecePerSeg ? n : kPiecePerSeg;       // cannot overflow: <= 5 far sub-ranges, <= 5 parent cuts
    }
    __syncthreads();
    if (tid == 0) {
        FarPiece* dst = a.pieces + (size_t)tile * nseg * kPiecePerSeg;
        int np = 0, nu = 0, nt = 0;
        for (int pass = 0; pass < 2; pass++)          // non-mixing pieces first (they share the unit numbering)
            for (int s = 0; s < nseg; s++)
                for (int i = 0; i < s_npc[s]; i++) {
                    FarPiece fp = s_pc[s * kPiecePerSeg + i];
                    if (((fp.info >> 17) & 1) != pass) continue;
                    fp.off = nu;
                    if (pass == 0) {
                        const bool both = (fp.info >> 16) & 1;
                        nu += both ? fp.n : (fp.n + 1) / 2;
                        nt += both ? 2 * fp.n : fp.n;
                    }
                    dst[np++] = fp;
                }
        a.hdr[tile].npieces = np;
        a.hdr[tile].nunits = nu;
        a.hdr[tile].nterms = nt;
    }
    __syncthreads();
    {   // plan -> HBM
        const int nw = nseg * (int)(sizeof(SegWork) / 4);
        const int* src = reinterpret_cast<const int*>(s_work);
        int* dst = reinterpret_cast<int*>(a.out + (size_t)tile * nseg);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
    }
}

// ---------------------------------------------------------------------------------------------
// far field
// One far-field term: w/((D+t)^2+h2) (+ optional pedestal) expanded in s = t/h about the tile centre,
//   sum_k b_k s^k,  b_0 = w*u, b_1 = al*b_0, b_k = al*b_{k-1} + be*b_{k-2},  u = 1/(D^2+h2), al = -2*D*h*u, be = -h^2*u.
// The poles of the term sit at distance sqrt(D^2+h2) >= ratio*h from the centre, so the series converges like ratio^-k.
__device__ __forceinline__ void far_accum(double D, double h2, double w, double ped, double m2h, double mhh, double (&A)[kFarK])
{
    const double u = rcp3(fma(D, D, h2));
    const double al = (D * m2h) * u, be = mhh * u;
    double b0 = w * u;
    double b1 = al * b0;
    A[0] += b0 - ped;
    A[1] += b1;
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b2 = fma(al, b1, be * b0);
        A[k] += b2;
        b0 = b1;
        b1 = b2;
    }
}
// two independent terms interleaved (instruction-level parallelism for the two recurrences)
__device__ __forceinline__ void far_accum2(double D1, double h21, double w1, double p1, double D2, double h22, double w2, double p2,
                                           double m2h, double mhh, double (&A)[kFarK])
{
    const double u1 = rcp3(fma(D1, D1, h21)), u2 = rcp3(fma(D2, D2, h22));
    const double al1 = (D1 * m2h) * u1, be1 = mhh * u1, al2 = (D2 * m2h) * u2, be2 = mhh * u2;
    double b01 = w1 * u1, b02 = w2 * u2;
    double b11 = al1 * b01, b12 = al2 * b02;
    A[0] += (b01 - p1) + (b02 - p2);
    A[1] += b11 + b12;
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b21 = fma(al1, b11, be1 * b01), b22 = fma(al2, b12, be2 * b02);
        A[k] += b21 + b22;
        b01 = b11; b11 = b21;
        b02 = b12; b12 = b22;
    }
}
// first-order line mixing (modm.f90:777-786): (g + c*(D+t))/((D+t)^2+h2), c = +-cq
__device__ __forceinline__ void far_accum_mix(double D, double h2, double cg, double cq, double hh, double m2h, double mhh, double (&A)[kFarK])
{
    const double u = rcp3(fma(D, D, h2));
    const double al = (D * m2h) * u, be = mhh * u;
    const double g1 = fma(cq, D, cg), g2 = cq * hh;
    double b0 = u;
    double b1 = al * b0;
    A[0] = fma(g1, b0, A[0]);
    A[1] = fma(g1, b1, fma(g2, b0, A[1]));
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b2 = fma(al, b1, be * b0);
        A[k] = fma(g1, b2, fma(g2, b1, A[k]));
        b0 = b1;
        b1 = b2;
    }
}

// CTA-wide sums of the per-thread coefficients in a fixed order (deterministic); result in s_coef[kFarK]
template <int NT>
__device__ __forceinline__ void reduce_coefs(const double (&A)[kFarK], int tid, double (*s_red)[NT], double (*s_red2)[8], double* s_coef)
{
    constexpr int CH = NT / 8;
#pragma unroll
    for (int i = 0; i < kFarK; i++) s_red[i][tid] = A[i];
    __syncthreads();
    if (tid < kFarK * 8) {
        const int i = tid >> 3, part = tid & 7;
        double t = 0.;
        for (int j = 0; j < CH; j++) t += s_red[i][part * CH + ((j + tid) & (CH - 1))];
        s_red2[i][part] = t;
    }
    __syncthreads();
    if (tid < kFarK) {
        double t = 0.;
#pragma unroll
        for (int j = 0; j < 8; j++) t += s_red2[tid][j];
        s_coef[tid] = t;
    }
    __syncthreads();
}

// binomial coefficients C(j,k), j,k < kFarK (polynomial translation between hierarchy levels)
struct BinomTable {
    double c[kFarK][kFarK];
    constexpr BinomTable() : c{}
    {
        for (int j = 0; j < kFarK; j++)
            for (int k = 0; k < kFarK; k++) {
                double v = 0.;
                if (k <= j) {
                    v = 1.;
                    for (int i = 1; i <= k; i++) v = v * (double)(j - k + i) / (double)i;
                }
                c[j][k] = v;
            }
    }
};
__constant__ BinomTable c_binom = BinomTable();

// =============================================================================================
// far_kernel: one launch per hierarchy level, top level first.  CTA = (level tile, layer, profile).
// Expands the lines that are far at this level but not at the parent level in kFarK Taylor terms about
// the tile centre, adds the parent tile's polynomial re-expanded about this centre (exact polynomial
// translation), and writes one coefficient set per molecule slot.  After the level-0 launch every line
// that is far from a level-0 tile -- at whatever level it was expanded -- is contained in that tile's
// coefficients; final_kernel evaluates them once per frequency.
// =============================================================================================
struct FarArgs {
    int32_t nlay, nseg, n_pad, nslot;
    const Segment* seg;
    const FarPiece* pieces;   // [ntiles][nseg*kPiecePerSeg] work list of this level (plan_kernel)
    const TileHdr* hdr;
    const TileHdr* phdr;      // parent level or null
    const double* pcoef;      // parent's coefficients [ptile][L][slot][kFarK]
    int32_t S;                // tiles of this level per parent tile
    int32_t combined;         // 1: one coefficient set, lines weighted by their molecule's column amount (nslot == 1)
    const double* planes;
    const LayerDev* lay;
    double* coef;             // [tile][L][slot][kFarK]
    unsigned long long* counters;
};

__global__ void __launch_bounds__(128, MRTM_FAR_MINB) far_kernel(FarArgs a)
{
    constexpr int NT = 128;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    FarPiece* s_pc = reinterpret_cast<FarPiece*>(s_dyn);        // the tile's work list
    __shared__ double s_red[kFarK][NT];
    __shared__ double s_red2[kFarK][8];
    __shared__ double s_coef[kFarK];
    __shared__ double s_pcoef[kMaxSlots * kFarK];       // the parent tile's coefficients
    __shared__ double s_w[kMaxSegments];                // column amount of each segment's molecule
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int64_t Ltot = (int64_t)gridDim.y * gridDim.z;
    const int64_t L = (int64_t)blockIdx.z * a.nlay + blockIdx.y;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ pP4 = pl + (size_t)D_P4 * a.n_pad;
    const int nseg = a.nseg;
    const int ptile = tile / a.S;
    const TileHdr th = a.hdr[tile];
    const int np = th.npieces;
    {
        const int nw = np * (int)(sizeof(FarPiece) / 4);
        const int* src = reinterpret_cast<const int*>(a.pieces + (size_t)tile * nseg * kPiecePerSeg);
        int* dst = reinterpret_cast<int*>(s_pc);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        if (a.pcoef) {
            const double* pin = a.pcoef + ((size_t)ptile * Ltot + L) * a.nslot * kFarK;
            for (int i = tid; i < a.nslot * kFarK; i += NT) s_pcoef[i] = pin[i];
        }
        for (int s = tid; s < nseg; s += NT) s_w[s] = ly.wk[a.seg[s].mol - 1];
    }
    __syncthreads();
    const double cen = 0.5 * (th.wlo + th.whi), hh = 0.5 * (th.whi - th.wlo);
    const double m2h = -2. * hh, mhh = -hh * hh;
    double alpha = 0., beta = 0.;       // parent variable s_p = alpha + beta*s
    if (a.pcoef) {
        const TileHdr ph = a.phdr[ptile];
        const double pc = 0.5 * (ph.wlo + ph.whi), phh = 0.5 * (ph.whi - ph.wlo);
        if (phh > 0.) { alpha = (cen - pc) / phh; beta = hh / phh; }
    }
    double* out = a.coef + ((size_t)tile * Ltot + L) * a.nslot * kFarK;
    const double* pin = a.pcoef ? s_pcoef : nullptr;
    // coefficient tid of the parent's polynomial p(alpha + beta*s) re-expanded in s:
    // beta^tid * sum_{j>=tid} c_j C(j,tid) alpha^(j-tid)
    auto translated = [&](const double* pcf) -> double {
        double acc = 0.;
        for (int j = kFarK - 1; j >= tid; j--) acc = fma(acc, alpha, pcf[j] * c_binom.c[j][tid]);
        double bk = 1.;
        for (int i = 0; i < tid; i++) bk *= beta;
        return acc * bk;
    };
    // number of leading non-mixing pieces
    int npm = 0;
    while (npm < np && !((s_pc[npm].info >> 17) & 1)) npm++;

    // work units [vbeg,vend) of the non-mixing pieces [pbeg,pend): one unit (two independent recurrences) per thread
    // and step, the next unit's line parameters already loading
    auto accumulate = [&](int pbeg, int pend, int vbeg, int vend, bool weighted, double (&A)[kFarK]) {
        auto fetch = [&](int v, int& pi, double& D1, double& g1, double& w1, double& p1, double& D2, double& g2, double& w2, double& p2) {
            D1 = 1.; g1 = 1.; w1 = 0.; p1 = 0.; D2 = 1.; g2 = 1.; w2 = 0.; p2 = 0.;
            if (v >= vend) return;
            while (pi + 1 < pend && v >= s_pc[pi + 1].off) pi++;
            const FarPiece fp = s_pc[pi];
            const int kk = v - fp.off;
            const double ws = weighted ? s_w[fp.info & 0xffff] : 1.;
            if ((fp.info >> 16) & 1) {                    // both resonances of one line: cen - xnu and cen + xnu
                const int q = fp.lo + kk;
                const double xnu = __ldg(pXNU + q);
                g1 = g2 = __ldg(pH2 + q);
                w1 = w2 = ws * __ldg(pCN + q);
                p1 = p2 = ws * __ldg(pP3 + q);
                D1 = cen - xnu;
                D2 = cen + xnu;
            } else {                                      // two adjacent single-resonance lines
                const int q = fp.lo + 2 * kk;
                D1 = cen - __ldg(pXNU + q);
                g1 = __ldg(pH2 + q);
                w1 = ws * __ldg(pCN + q);
                p1 = ws * __ldg(pP3 + q);
                if (2 * kk + 1 < fp.n) {
                    D2 = cen - __ldg(pXNU + q + 1);
                    g2 = __ldg(pH2 + q + 1);
                    w2 = ws * __ldg(pCN + q + 1);
                    p2 = ws * __ldg(pP3 + q + 1);
                }
            }
        };
        int pi = pbeg;
        int v = vbeg + tid;
        double D1, g1, w1, p1, D2, g2, w2, p2;
        fetch(v, pi, D1, g1, w1, p1, D2, g2, w2, p2);
        while (v < vend) {
            const int vn = v + NT;
            double D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n;
            fetch(vn, pi, D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n);
            far_accum2(D1, g1, w1, p1, D2, g2, w2, p2, m2h, mhh, A);
            D1 = D1n; g1 = g1n; w1 = w1n; p1 = p1n;
            D2 = D2n; g2 = g2n; w2 = w2n; p2 = p2n;
            v = vn;
        }
    };
    // first-order mixing pieces [pbeg,pend): few lines, one line (both resonances) per thread and step
    auto accumulate_mix = [&](int pbeg, int pend, bool weighted, double (&A)[kFarK]) {
        for (int pi = pbeg; pi < pend; pi++) {
            const FarPiece fp = s_pc[pi];
            const double ws = weighted ? s_w[fp.info & 0xffff] : 1.;
            for (int q = fp.lo + tid; q < fp.lo + fp.n; q += NT) {
                const double xnu = __ldg(pXNU + q), h2 = __ldg(pH2 + q), cg = ws * __ldg(pP3 + q), cq = ws * __ldg(pP4 + q);
                far_accum_mix(cen - xnu, h2, cg, cq, hh, m2h, mhh, A);
                far_accum_mix(cen + xnu, h2, cg, -cq, hh, m2h, mhh, A);
            }
        }
    };

    if (a.combined) {
        // one coefficient set for all molecules: every line enters with its molecule's column amount W
        // (o = RFT * sum_mol W_mol*SF_mol, modm.f90:436-438, 265-267)
        double A[kFarK];
#pragma unroll
        for (int i = 0; i < kFarK; i++) A[i] = 0.;
        if (np > 0) {
            accumulate(0, npm, 0, th.nunits, true, A);
            accumulate_mix(npm, np, true, A);
            reduce_coefs<NT>(A, tid, s_red, s_red2, s_coef);
        }
        if (tid < kFarK) {
            double v = (np > 0) ? s_coef[tid] : 0.;
            if (pin) v += translated(pin);
            out[tid] = v;
        }
    } else {
        // one coefficient set per molecule slot; the pieces of a molecule are contiguous in both lists
        int s = 0;
        while (s < nseg) {
            const int mol = a.seg[s].mol, slot = a.seg[s].slot;
            int s_end = s;
            while (s_end < nseg && a.seg[s_end].mol == mol) s_end++;
            int pb = 0, pe, mb = npm, me;
            while (pb < npm && (s_pc[pb].info & 0xffff) < s) pb++;
            pe = pb;
            while (pe < npm && (s_pc[pe].info & 0xffff) < s_end) pe++;
            while (mb < np && (s_pc[mb].info & 0xffff) < s) mb++;
            me = mb;
            while (me < np && (s_pc[me].info & 0xffff) < s_end) me++;
            const bool active = ly.wk[mol - 1] != 0.;
            const bool work = active && (pe > pb || me > mb);
            if (work) {
                double A[kFarK];
#pragma unroll
                for (int i = 0; i < kFarK; i++) A[i] = 0.;
                if (pe > pb) {
                    const FarPiece last = s_pc[pe - 1];
                    accumulate(pb, pe, s_pc[pb].off, last.off + ((((last.info >> 16) & 1)) ? last.n : (last.n + 1) / 2), false, A);
                }
                accumulate_mix(mb, me, false, A);
                reduce_coefs<NT>(A, tid, s_red, s_red2, s_coef);
            }
            if (tid < kFarK) {
                double v = work ? s_coef[tid] : 0.;
                if (pin && active) v += translated(pin + (size_t)slot * kFarK);
                out[(size_t)slot * kFarK + tid] = v;
            }
            __syncthreads();        // s_coef is reused by the next molecule
            s = s_end;
        }
    }
    if (a.counters && tid == 0) {
        long long n_far = th.nterms;
        for (int pi = npm; pi < np; pi++) n_far += 2ll * s_pc[pi].n;
        atomicAdd(a.counters + 0, (unsigned long long)n_far);
    }
}

// =============================================================================================
// far_warp_kernel: far_kernel for the levels with many small tiles (level 0 above all) in the combined
// mode (one coefficient set for all molecules).  A WARP owns one (tile, layer): the four warps of a CTA
// take four consecutive layers of the same tile and share its work list in shared memory; each lane walks
// the tile's work units with stride 32, the coefficients are summed across the warp by shuffles (fixed
// butterfly order: deterministic) -- no CTA barrier after the list is staged, and four times more units
// per lane than far_kernel has per thread, which amortises the set-up, reduction and translation.
// =============================================================================================
#ifndef MRTM_FARW_MINB
#define MRTM_FARW_MINB 5
#endif
constexpr int kFarWarps = 4;
__global__ void __launch_bounds__(32 * kFarWarps, MRTM_FARW_MINB) far_warp_kernel(FarArgs a)
{
    constexpr int NT = 32 * kFarWarps;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    FarPiece* s_pc = reinterpret_cast<FarPiece*>(s_dyn);        // the tile's work list
    __shared__ double s_pcoef[kFarWarps][kFarK];                // the parent tile's coefficients, per layer
    __shared__ double s_w[kFarWarps][kMaxSegments];             // column amount of each segment's molecule, per layer
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const int nseg = a.nseg;
    const TileHdr th = a.hdr[tile];
    const int np = th.npieces;
    {
        const int nw = np * (int)(sizeof(FarPiece) / 4);
        const int* src = reinterpret_cast<const int*>(a.pieces + (size_t)tile * nseg * kPiecePerSeg);
        int* dst = reinterpret_cast<int*>(s_pc);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const int k = blockIdx.y * kFarWarps + wid;
    if (k >= a.nlay) return;                                    // no barrier follows
    const int64_t Ltot = (int64_t)a.nlay * gridDim.z;
    const int64_t L = (int64_t)blockIdx.z * a.nlay + k;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ pP4 = pl + (size_t)D_P4 * a.n_pad;
    const int ptile = tile / a.S;
    double* sw = s_w[wid];
    for (int s = lane; s < nseg; s += 32) sw[s] = ly.wk[a.seg[s].mol - 1];
    if (a.pcoef && lane < kFarK) s_pcoef[wid][lane] = a.pcoef[((size_t)ptile * Ltot + L) * kFarK + lane];
    __syncwarp();
    const double cen = 0.5 * (th.wlo + th.whi), hh = 0.5 * (th.whi - th.wlo);
    const double m2h = -2. * hh, mhh = -hh * hh;
    int npm = 0;                        // number of leading non-mixing pieces
    while (npm < np && !((s_pc[npm].info >> 17) & 1)) npm++;

    double A[kFarK];
#pragma unroll
    for (int i = 0; i < kFarK; i++) A[i] = 0.;
    {
        const int vend = th.nunits;
        auto fetch = [&](int v, int& pi, double& D1, double& g1, double& w1, double& p1, double& D2, double& g2, double& w2, double& p2) {
            D1 = 1.; g1 = 1.; w1 = 0.; p1 = 0.; D2 = 1.; g2 = 1.; w2 = 0.; p2 = 0.;
            if (v >= vend) return;
            while (pi + 1 < npm && v >= s_pc[pi + 1].off) pi++;
            const FarPiece fp = s_pc[pi];
            const int kk = v - fp.off;
            const double ws = sw[fp.info & 0xffff];
            if ((fp.info >> 16) & 1) {                    // both resonances of one line: cen - xnu and cen + xnu
                const int q = fp.lo + kk;
                const double xnu = __ldg(pXNU + q);
                g1 = g2 = __ldg(pH2 + q);
                w1 = w2 = ws * __ldg(pCN + q);
                p1 = p2 = ws * __ldg(pP3 + q);
                D1 = cen - xnu;
                D2 = cen + xnu;
            } else {                                      // two adjacent single-resonance lines
                const int q = fp.lo + 2 * kk;
                D1 = cen - __ldg(pXNU + q);
                g1 = __ldg(pH2 + q);
                w1 = ws * __ldg(pCN + q);
                p1 = ws * __ldg(pP3 + q);
                if (2 * kk + 1 < fp.n) {
                    D2 = cen - __ldg(pXNU + q + 1);
                    g2 = __ldg(pH2 + q + 1);
                    w2 = ws * __ldg(pCN + q + 1);
                    p2 = ws * __ldg(pP3 + q + 1);
                }
            }
        };
        int pi = 0;
        int v = lane;
        double D1, g1, w1, p1, D2, g2, w2, p2;
        if (npm > 0) {
            fetch(v, pi, D1, g1, w1, p1, D2, g2, w2, p2);
            while (v < vend) {
                const int vn = v + 32;
                double D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n;
                fetch(vn, pi, D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n);
                far_accum2(D1, g1, w1, p1, D2, g2, w2, p2, m2h, mhh, A);
                D1 = D1n; g1 = g1n; w1 = w1n; p1 = p1n;
                D2 = D2n; g2 = g2n; w2 = w2n; p2 = p2n;
                v = vn;
            }
        }
        // first-order mixing pieces: few lines, one line (both resonances) per lane and step
        for (int pj = npm; pj < np; pj++) {
            const FarPiece fp = s_pc[pj];
            const double ws = sw[fp.info & 0xffff];
            for (int q = fp.lo + lane; q < fp.lo + fp.n; q += 32) {
                const double xnu = __ldg(pXNU + q), h2 = __ldg(pH2 + q), cg = ws * __ldg(pP3 + q), cq = ws * __ldg(pP4 + q);
                far_accum_mix(cen - xnu, h2, cg, cq, hh, m2h, mhh, A);
                far_accum_mix(cen + xnu, h2, cg, -cq, hh, m2h, mhh, A);
            }
        }
    }
    // warp sums; lane i keeps coefficient i
    double mine = 0.;
#pragma unroll
    for (int i = 0; i < kFarK; i++) {
        double v = A[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == i) mine = v;
    }
    if (lane < kFarK) {
        if (a.pcoef) {
            // coefficient `lane` of the parent's polynomial p(alpha + beta*s) re-expanded in s:
            // beta^lane * sum_{j>=lane} c_j C(j,lane) alpha^(j-lane)
            const TileHdr ph = a.phdr[ptile];
            const double pc = 0.5 * (ph.wlo + ph.whi), phh = 0.5 * (ph.whi - ph.wlo);
            double alpha = 0., beta = 0.;
            if (phh > 0.) { alpha = (cen - pc) / phh; beta = hh / phh; }
            const double* pcf = s_pcoef[wid];
            double acc = 0.;
            for (int j = kFarK - 1; j >= lane; j--) acc = fma(acc, alpha, pcf[j] * c_binom.c[j][lane]);
            double bk = 1.;
            for (int i = 0; i < lane; i++) bk *= beta;
            mine += acc * bk;
        }
        a.coef[((size_t)tile * Ltot + L) * kFarK + lane] = mine;
    }
    if (a.counters && lane == 0) {
        long long n_far = th.nterms;
        for (int pj = npm; pj < np; pj++) n_far += 2ll * s_pc[pj].n;
        atomicAdd(a.counters + 0, (unsigned long long)n_far);
    }
}

// =============================================================================================
// near_kernel: the per-(line,layer,frequency) evaluations that remain after the far field is taken
// out.  CTA = (NT*F frequencies, one layer, one profile); each thread owns F (frequency, layer)
// accumulators.  The tile's plan (plan_kernel) is copied from HBM: no searches here.
//  * line-parameter tiles (XNU, H2, CN, P3) of the direct runs stream through shared memory with TMA bulk
//    copies on an 8-stage mbarrier ring; all threads read the same line -> smem broadcast
//  * interior ranges run branch-free (4 lines share one reciprocal); the narrow bands (window edges, the
//    WN+Xnu<=25 boundary, the Voigt zone) run loops with the reference's exact per-(line,frequency) tests
//    (modm.f90:384, 427, 746); (line,frequency) pairs on the Voigt branch are left to voigt_kernel
//  * writes sum_mol W_mol*SF_mol (direct part, without RFT) to O and, when per-molecule outputs are
//    requested, W_mol*SF_mol to O_BY_MOL; final_kernel completes them
// =============================================================================================
template <int F, bool SEL, int NT>
__global__ void __launch_bounds__(NT, (MRTM_LINES_MINB * 128) / NT) near_kernel(LinesArgs a)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x;
    if (a.near_pieces) {                          // near2_kernel took the tiles whose direct lines fit its staging area
        const TileHdr th0 = a.hdr[0][blockIdx.x];
        if (th0.total_lines <= kNearCap && th0.nnear >= 0) return;
    }
    const int k = blockIdx.y;                     // layer within profile
    const int prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ pP4 = pl + (size_t)D_P4 * a.n_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;

    __shared__ __align__(8) uint64_t s_bar[kStages];
    __shared__ __align__(8) uint64_t s_all_bar;
    __shared__ double s_ped[2][NW];
    __shared__ unsigned char s_act[kMaxSegments];
    extern __shared__ __align__(128) unsigned char s_dyn[];
    double (*s_tile)[4][kTile] = reinterpret_cast<double (*)[4][kTile]>(s_dyn);      // [kStages][4][kTile]
    SegWork* s_work = reinterpret_cast<SegWork*>(s_dyn + sizeof(double) * kStages * 4 * kTile);

    const int nseg = a.nseg;
    {
        const int nw = nseg * (int)(sizeof(SegWork) / 4);
        const int* src = reinterpret_cast<const int*>(a.plan[0] + (size_t)blockIdx.x * nseg);
        int* dst = reinterpret_cast<int*>(s_work);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        for (int s = tid; s < nseg; s += NT) s_act[s] = (ly.wk[a.seg[s].mol - 1] != 0.) ? 1 : 0;   // W_SPECIES == 0: skipped (:318-321)
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; i++) mbar_init(&s_bar[i], 1);
        mbar_init(&s_all_bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // this thread's frequencies (strided so global accesses coalesce)
    const int base = blockIdx.x * (NT * F);
    double wn[F];
    bool valid[F];
#pragma unroll
    for (int f = 0; f < F; f++) {
        int iw = base + f * NT + tid;
        valid[f] = iw < a.nwn;
        wn[f] = a.wn[valid[f] ? iw : (a.nwn - 1)];
    }
    const double rp = ly.rp, rp2 = ly.rp2;
    __syncthreads();
    // Staging: when all direct runs of the CTA fit the tile memory (the usual case with the far field on) they
    // are staged at once -- one mbarrier wait, no per-tile hand-shake; otherwise tiles stream through the ring.
    constexpr int kCap = kStages * kTile;
    const bool stage_all = a.hdr[0][blockIdx.x].total_lines <= kCap;
    double* s_all = reinterpret_cast<double*>(s_dyn);       // [4][kCap] in stage-all mode

    // ---- TMA tile jobs: (segment, run, tile) in consumption order; thread 0 keeps a cursor ahead of the consumer
    auto advance = [&](int& js, int& jr, int& jt) -> bool {
        jt++;
        while (js < nseg) {
            const SegWork& w = s_work[js];
            if (w.tma && s_act[js] && jr < w.nrun) {
                if (jt < w.run_nt[jr]) return true;
                jr++;
                jt = 0;
                continue;
            }
            js++;
            jr = 0;
            jt = 0;
        }
        return false;
    };
    auto issue = [&](int js, int jr, int jt, int st) {      // one elected thread: TMA one tile into stage st
        const int qs = s_work[js].run_t0[jr] + jt * kTile;
        int n = a.n_pad - qs;
        n = n > kTile ? kTile : n;
        const uint32_t bytes = (uint32_t)n * 8u;
        mbar_expect_tx(&s_bar[st], 4u * bytes);
        tma_load_1d(&s_tile[st][0][0], pXNU + qs, bytes, &s_bar[st]);
        tma_load_1d(&s_tile[st][1][0], pH2 + qs, bytes, &s_bar[st]);
        tma_load_1d(&s_tile[st][2][0], pCN + qs, bytes, &s_bar[st]);
        tma_load_1d(&s_tile[st][3][0], pP3 + qs, bytes, &s_bar[st]);
    };
    int ped_buf = 0;            // alternates per reduction (s_ped double buffer)
    int pjs = 0, pjr = 0, pjt = -1;      // prefetch cursor (thread 0 only)
    bool more = !stage_all;
    if (stage_all) {
        // every lane of warp 0 announces and issues the copies of its own segments (barrier count 32)
        if (tid < 32) {
            uint32_t mybytes = 0;
            for (int s = tid; s < nseg; s += 32) {
                const SegWork& w = s_work[s];
                if (!(w.tma && s_act[s])) continue;
                for (int r = 0; r < w.nrun; r++) mybytes += (uint32_t)(((w.run_hi[r] - w.run_t0[r]) + 3) & ~3) * 32u;
            }
            mbar_expect_tx(&s_all_bar, mybytes);
            for (int s = tid; s < nseg; s += 32) {
                const SegWork& w = s_work[s];
                if (!(w.tma && s_act[s])) continue;
                for (int r = 0; r < w.nrun; r++) {
                    const int qs = w.run_t0[r], off = w.run_off[r];
                    const uint32_t bytes = (uint32_t)(((w.run_hi[r] - qs) + 3) & ~3) * 8u;
                    tma_load_1d(s_all + 0 * kCap + off, pXNU + qs, bytes, &s_all_bar);
                    tma_load_1d(s_all + 1 * kCap + off, pH2 + qs, bytes, &s_all_bar);
                    tma_load_1d(s_all + 2 * kCap + off, pCN + qs, bytes, &s_all_bar);
                    tma_load_1d(s_all + 3 * kCap + off, pP3 + qs, bytes, &s_all_bar);
                }
            }
        }
        mbar_wait(&s_all_bar, 0u);
    } else if (tid == 0) {
        for (int i = 0; i < kStages - 1 && more; i++) {
            more = advance(pjs, pjr, pjt);
            if (more) issue(pjs, pjr, pjt, i);
        }
    }
    double ped_mol = 0., ped_w = 0.;     // stage-all mode: this thread's share of the interior pedestals (molecule / weighted total)
    auto cta_sum = [&](double v) -> double {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) s_ped[ped_buf][tid >> 5] = v;
        __syncthreads();
        double t = 0.;
#pragma unroll
        for (int i = 0; i < NW; i++) t += s_ped[ped_buf][i];
        ped_buf ^= 1;
        return t;
    };

    int gtile = 0;              // global tile counter: stage = gtile % kStages, mbarrier parity = (gtile / kStages) & 1

    double osum[F], sf[F];
    long long cnt[F];
    unsigned long long hsh[F];
#pragma unroll
    for (int f = 0; f < F; f++) { osum[f] = 0.; sf[f] = 0.; cnt[f] = 0; hsh[f] = 0ull; }

    int err = 0;
    long long n_direct = 0;      // work counter (thread 0 reports it)
    int nvalid = 0;
    if (a.counters) {
        const int rem = a.nwn - base;
        nvalid = rem < NT * F ? rem : NT * F;
    }
    // a layer without Voigt-capable lines runs its Voigt zones as plain near-field ranges
    const int vmode_mask = voigt_possible(a.layer_voigt, L, a.hdr[0][blockIdx.x].whi) ? 0xff : (0xff & ~M_VOIGT);
    int cur_mol = 0;
    auto finish_mol = [&](int mol) {
        if (mol <= 0) return;
        const double w = ly.wk[mol - 1];
        if (stage_all) {
            if (a.o_by_mol) {           // per-molecule outputs: close the pedestal sum per molecule
                const double pacc = cta_sum(ped_mol);
#pragma unroll
                for (int f = 0; f < F; f++) sf[f] -= pacc;
            } else {
                ped_w = fma(w, ped_mol, ped_w);
            }
            ped_mol = 0.;
        }
#pragma unroll
        for (int f = 0; f < F; f++) {
            const double ol = (w == 0.) ? 0. : (w * sf[f]);            // W*SF; RFT is applied by final_kernel (modm.f90:436-438)
            osum[f] = osum[f] + ol;                                    // :265-267 (molecule order)
            if (a.o_by_mol && valid[f]) {
                int iw = base + f * NT + tid;
                a.o_by_mol[(size_t)iw + (size_t)(mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] = ol;
            }
            sf[f] = 0.;
        }
    };

    for (int s = 0; s < nseg; s++) {
        const Segment sg = a.seg[s];
        if (sg.mol != cur_mol) {
            finish_mol(cur_mol);
            cur_mol = sg.mol;
        }
        const SegWork& wk = s_work[s];
        if (!s_act[s]) continue;
        const int cls = sg.cls;
        if (SEL) {
            if (sg.mol == 7) {                             // every O2 line passes modm.f90:384
#pragma unroll
                for (int f = 0; f < F; f++) { cnt[f] += sg.count_all; hsh[f] += sg.hash_all; }
            } else if (cls == CLS_PED) {                   // every far line (any level) is inside the window of every frequency
                for (int u = 0; u + 1 < wk.nbp; u++) {
                    if (wk.mode[u] != 0) continue;
                    const int lo = wk.bp[u], hi = wk.bp[u + 1];
                    const unsigned long long hs = a.keypre[hi] - a.keypre[lo];
#pragma unroll
                    for (int f = 0; f < F; f++) { cnt[f] += hi - lo; hsh[f] += hs; }
                }
            }
        }
        if (cls == CLS_PED || cls == CLS_O2 || cls == CLS_O2_LC35) {
            const bool force_both = (cls == CLS_O2_LC35);
            const bool count_sel = SEL && (cls == CLS_PED);
            const int n0 = wk.n0;
            double psum[F];
#pragma unroll
            for (int f = 0; f < F; f++) psum[f] = 0.;
            double pacc = 0.;                              // pedestal total of the interior ranges (uniform)
            for (int r = 0; r < wk.nrun; r++) {
                const int rlo = wk.run_lo[r], rhi = wk.run_hi[r], t0 = wk.run_t0[r];
                const int ntile = stage_all ? 1 : wk.run_nt[r];
                for (int t = 0; t < ntile; t++) {
                    const double *tX, *tH, *tC, *tP;
                    int tb, thi;
                    if (stage_all) {
                        const int off = wk.run_off[r];
                        tX = s_all + off; tH = s_all + kCap + off; tC = s_all + 2 * kCap + off; tP = s_all + 3 * kCap + off;
                        tb = t0;
                        thi = rhi;
                    } else {
                        const int st = gtile % kStages;
                        if (tid == 0 && more) {        // refill the stage the previous tile released
                            more = advance(pjs, pjr, pjt);
                            if (more) issue(pjs, pjr, pjt, (gtile + kStages - 1) % kStages);
                        }
                        mbar_wait(&s_bar[st], (uint32_t)(gtile / kStages) & 1u);
                        tX = s_tile[st][0]; tH = s_tile[st][1]; tC = s_tile[st][2]; tP = s_tile[st][3];
                        tb = t0 + t * kTile;
                        thi = (tb + kTile) < rhi ? (tb + kTile) : rhi;
                        gtile++;
                    }
                    const int tlo = tb > rlo ? tb : rlo;
                    double pmine = 0.;                          // this thread's share of the tile's interior pedestals
                    for (int u = wk.run_u0[r]; u < wk.run_u1[r]; u++) {
                        const int x = wk.bp[u];
                        const int mode = wk.mode[u] & vmode_mask;
                        int lo = x > tlo ? x : tlo;
                        int hi = wk.bp[u + 1] < thi ? wk.bp[u + 1] : thi;
                        if (lo >= hi || wk.mode[u] == 0) continue;
                        const bool negall = force_both || (x < n0);
                        if (a.counters) n_direct += (long long)(hi - lo) * nvalid;
                        if ((mode & 7) != 0) {
                            // ---- band loops: the reference's exact per-(line,frequency) tests
                            const bool edge = (mode & M_EDGE) != 0, negtest = (mode & M_NEG) != 0, vz = (mode & M_VOIGT) != 0;
                            if (negall || negtest) {
                                for (int q = lo; q < hi; q++) {
                                    const int j = q - tb;
                                    const double xnu = tX[j], h2 = tH[j], cn = tC[j], ped = tP[j];
                                    const double vt = vz ? __ldg(pVT + q) : -1.0;
#pragma unroll
                                    for (int f = 0; f < F; f++) {
                                        const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                                        const bool inwin = edge ? !(fabs(dm) > kDELTNUC) : true;
                                        if (count_sel && inwin) { cnt[f]++; hsh[f] += a.key[q]; }
                                        const bool take = inwin && !(fabs(dm) <= vt);      // Voigt-branch pairs: voigt_kernel
                                        const bool neg = negall || (sp <= kDELTNUC);
                                        const double r1 = rcp3(fma(dm, dm, h2)), r2 = rcp3(fma(sp, sp, h2));
                                        const double val = cn * (r1 + (neg ? r2 : 0.)) - (neg ? 2. * ped : ped);
                                        sf[f] += take ? val : 0.;
                                    }
                                }
                            } else {
                                for (int q = lo; q < hi; q++) {
                                    const int j = q - tb;
                                    const double xnu = tX[j], h2 = tH[j], cn = tC[j], ped = tP[j];
                                    const double vt = vz ? __ldg(pVT + q) : -1.0;
#pragma unroll
                                    for (int f = 0; f < F; f++) {
                                        const double dm = wn[f] - xnu;
                                        const bool inwin = edge ? !(fabs(dm) > kDELTNUC) : true;
                                        if (count_sel && inwin) { cnt[f]++; hsh[f] += a.key[q]; }
                                        const bool take = inwin && !(fabs(dm) <= vt);
                                        const double val = fma(cn, rcp3(fma(dm, dm, h2)), -ped);
                                        sf[f] += take ? val : 0.;
                                    }
                                }
                            }
                            continue;
                        }
                        if (count_sel) {
                            const unsigned long long hs = a.keypre[hi] - a.keypre[lo];
#pragma unroll
                            for (int f = 0; f < F; f++) { cnt[f] += hi - lo; hsh[f] += hs; }
                        }
                        // this thread's share of the interior pedestals of the tile (reduced across the CTA below)
                        {
                            const double w = negall ? 2. : 1.;          // pedestal counted for both resonances (:749)
                            for (int q = lo + tid; q < hi; q += NT) pmine = fma(w, tP[q - tb], pmine);
                        }
                        if (negall) {
                            // ---- interior, both resonances: cn*(1/a+1/b) = cn*(a+b)/(a*b), one reciprocal
MRTM_UNROLL(MRTM_UNROLL_BOTH)
                            for (int q = lo; q < hi; q++) {
                                const int j = q - tb;
                                const double xnu = tX[j], h2 = tH[j], cn = tC[j];
#pragma unroll
                                for (int f = 0; f < F; f++) {
                                    const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                                    const double aa = fma(dm, dm, h2), bb = fma(sp, sp, h2);
                                    const double r = rcp3(aa * bb);
                                    psum[f] = fma(cn * (aa + bb), r, psum[f]);
                                }
                            }
                        } else {
                            // ---- interior, single resonance (modm.f90:751): four lines share one reciprocal,
                            // sum c_i/a_i = N/(a1 a2 a3 a4), N = (c1 a2 + c2 a1)(a3 a4) + (c3 a4 + c4 a3)(a1 a2);
                            // 21 FP64 ops + 1 MUFU per 4 evaluations
                            int q = lo;
                            for (; q + 4 <= hi; q += 4) {
                                const int j = q - tb;
                                const double x1 = tX[j], x2 = tX[j + 1], x3 = tX[j + 2], x4 = tX[j + 3];
                                const double g1 = tH[j], g2 = tH[j + 1], g3 = tH[j + 2], g4 = tH[j + 3];
                                const double c1 = tC[j], c2 = tC[j + 1], c3 = tC[j + 2], c4 = tC[j + 3];
#pragma unroll
                                for (int f = 0; f < F; f++) {
                                    const double d1 = wn[f] - x1, d2 = wn[f] - x2, d3 = wn[f] - x3, d4 = wn[f] - x4;
                                    const double a1 = fma(d1, d1, g1), a2 = fma(d2, d2, g2);
                                    const double a3 = fma(d3, d3, g3), a4 = fma(d4, d4, g4);
                                    const double p12 = a1 * a2, p34 = a3 * a4;
                                    const double n12 = fma(c1, a2, c2 * a1), n34 = fma(c3, a4, c4 * a3);
                                    const double r = rcp3(p12 * p34);
                                    psum[f] = fma(fma(n12, p34, n34 * p12), r, psum[f]);
                                }
                            }
                            for (; q < hi; q++) {
                                const int j = q - tb;
                                const double xnu = tX[j], h2 = tH[j], cn = tC[j];
#pragma unroll
                                for (int f = 0; f < F; f++) {
                                    const double dm = wn[f] - xnu;
                                    psum[f] = fma(cn, rcp3(fma(dm, dm, h2)), psum[f]);
                                }
                            }
                        }
                    }
                    if (stage_all) {
                        ped_mol += pmine;
                    } else {
                        // CTA-wide sum of the interior pedestals of this tile (uniform result); the barrier inside also
                        // orders all reads of this stage before it is refilled
                        pacc += cta_sum(pmine);
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < F; f++) sf[f] += psum[f] - pacc;
        } else if (cls == CLS_O2_LC1) {
            for (int r = 0; r < wk.nrun; r++) {
                if (a.counters) n_direct += (long long)(wk.run_hi[r] - wk.run_lo[r]) * nvalid;
                for (int q = wk.run_lo[r]; q < wk.run_hi[r]; q++) {
                    const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q], cq = pP4[q], vt = pVT[q];
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                        const double r1 = rcp3(fma(dm, dm, h2));
                        const double r2 = rcp3(fma(sp, sp, h2));
                        const double val = fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;
                        sf[f] += (fabs(dm) <= vt) ? 0. : val;           // Voigt-branch pairs: voigt_kernel
                    }
                }
            }
        } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
            if (a.counters) n_direct += (long long)(wk.q1 - wk.q0) * nvalid;
            for (int q = wk.q0; q < wk.q1; q++) {
                const double xnu = pXNU[q], vt = pVT[q];
                const double hw = pl[(size_t)D_H * a.n_pad + q], ad = pl[(size_t)D_AD * a.n_pad + q];
                const double st = pl[(size_t)D_STILD * a.n_pad + q];
                const double aip = pl[(size_t)D_AIP * a.n_pad + q], bip = pl[(size_t)D_BIP * a.n_pad + q];
                const int xf = a.xf_s[q];
#pragma unroll
                for (int f = 0; f < F; f++) {
                    const double dm = wn[f] - xnu;
                    if ((fabs(dm) > kDELTNUC) && (sg.mol != 7)) continue;          // modm.f90:384
                    if (SEL && sg.mol != 7) { cnt[f]++; hsh[f] += a.key[q]; }
                    const bool voigt = fabs(dm) <= vt;
                    sf[f] += st * lsf_general(sg.mol, xf, rp, rp2, aip, bip, hw, wn[f], xnu, ad, a.sdep_s[q], voigt, &err);
                }
            }
        }
    }
    finish_mol(cur_mol);
    if (stage_all && !a.o_by_mol) {       // one CTA reduction for the W-weighted interior pedestals of all molecules
        const double pacc = cta_sum(ped_w);
#pragma unroll
        for (int f = 0; f < F; f++) osum[f] -= pacc;
    }
    if (err) atomicOr(a.errflag, 2);
    if (a.counters && tid == 0) atomicAdd(a.counters + 1, (unsigned long long)n_direct);
#pragma unroll
    for (int f = 0; f < F; f++) {
        if (!valid[f]) continue;
        const int iw = base + f * NT + tid;
        const size_t fl = (size_t)iw + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
        a.o[fl] = osum[f];
        if (SEL) {
            if (a.sel_count) a.sel_count[fl] = cnt[f];
            if (a.sel_hash) a.sel_hash[fl] = hsh[f];
        }
    }
}

// =============================================================================================
// near2_kernel: the near field of the tiles whose direct lines all fit the staging area (the usual case
// with the far field on).  Same staging as near_kernel (TMA bulk copies of the XNU, H2, CN, P3 runs), but
// each WARP owns a contiguous block of 32*F frequencies and re-plans the staged lines for its own block:
//   * with the warp's lowest and highest frequency it decides per line -- exactly, the floating-point
//     differences are monotone in the frequency -- whether the window test (modm.f90:384), the
//     WN+Xnu<=25 test (:746) or the Voigt test (:427) can come out differently inside the block; only those
//     lines run the per-(line,frequency) tests (lists T1/T2)
//   * a line whose poles Xnu +- i*HWHM (and -Xnu +- i*HWHM) are at least ff_ratio block half-widths from
//     the block centre is expanded about that centre (kFarK Taylor terms, one line per lane, coefficients
//     summed across the warp by shuffles and evaluated once per frequency)
//   * the rest is evaluated per frequency from per-warp index lists (D1 single resonance, four lines share
//     one reciprocal; D2 both resonances)
// One group of lists serves all molecules (line strengths pre-multiplied by the column amounts) unless
// per-molecule outputs are requested.  No CTA barrier after the staging phase.
// =============================================================================================
template <int F, bool SEL, int NT>
__global__ void __launch_bounds__(NT, MRTM_NEAR2_MINB) near2_kernel(LinesArgs a)
{
    constexpr int NW = NT / 32;
    constexpr int kCap = kNearCap;
    constexpr int kDummy = kCap;                 // neutral staged slot (CN = 0) that pads the D1 list to groups of four
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const TileHdr th = a.hdr[0][blockIdx.x];
    if (th.total_lines > kCap || th.nnear < 0) return;       // near_kernel streams this tile
    const int k = blockIdx.y;                     // layer within profile
    const int prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ pP4 = pl + (size_t)D_P4 * a.n_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;

    __shared__ __align__(8) uint64_t s_all_bar;
    __shared__ unsigned char s_act[kMaxSegments];
    __shared__ double s_wseg[kMaxSegments];
    extern __shared__ __align__(128) unsigned char s_dyn[];
    constexpr int kPlane = kCap + 8;              // + the neutral slot, 64-byte multiple
    double* tX = reinterpret_cast<double*>(s_dyn);
    double* tH = tX + kPlane;
    double* tC = tH + kPlane;
    double* tP = tC + kPlane;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(tP + kPlane);          // [NW][2][kCap + 8]
    constexpr int kListLen = kCap + 8;
    unsigned char* s_pid = reinterpret_cast<unsigned char*>(s_list + NW * 2 * kListLen);   // [kCap]
    NearPiece* s_np = reinterpret_cast<NearPiece*>(s_pid + kCap);                     // [kMaxNearPieces]
    SegWork* s_work = reinterpret_cast<SegWork*>(s_np + kMaxNearPieces);              // [nseg]

    const int nseg = a.nseg;
    const int npc = th.nnear;
    const int total = th.total_lines;
    const bool by_mol = a.o_by_mol != nullptr;
    {
        const int nw = nseg * (int)(sizeof(SegWork) / 4);
        const int* src = reinterpret_cast<const int*>(a.plan[0] + (size_t)blockIdx.x * nseg);
        int* dst = reinterpret_cast<int*>(s_work);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        const int np4 = npc * (int)(sizeof(NearPiece) / 4);
        const int* psrc = reinterpret_cast<const int*>(a.near_pieces + (size_t)blockIdx.x * kMaxNearPieces);
        int* pdst = reinterpret_cast<int*>(s_np);
        for (int i = tid; i < np4; i += NT) pdst[i] = psrc[i];
        for (int s = tid; s < nseg; s += NT) {
            const double w = ly.wk[a.seg[s].mol - 1];
            s_act[s] = (w != 0.) ? 1 : 0;         // W_SPECIES == 0: skipped (modm.f90:318-321)
            s_wseg[s] = w;
        }
        for (int i = tid; i < (total + 3) / 4; i += NT) reinterpret_cast<uint32_t*>(s_pid)[i] = 0xffffffffu;
        if (tid == 0) {
            mbar_init(&s_all_bar, 32);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            tX[kDummy] = th.wlo; tH[kDummy] = 1.0; tC[kDummy] = 0.; tP[kDummy] = 0.;
        }
    }
    __syncthreads();
    // ---- staging: every lane of warp 0 announces and issues the copies of its own segments (barrier count 32)
    if (tid < 32) {
        uint32_t mybytes = 0;
        for (int s = tid; s < nseg; s += 32) {
            const SegWork& w = s_work[s];
            if (!(w.tma && s_act[s])) continue;
            for (int r = 0; r < w.nrun; r++) mybytes += (uint32_t)(((w.run_hi[r] - w.run_t0[r]) + 3) & ~3) * 32u;
        }
        mbar_expect_tx(&s_all_bar, mybytes);
        for (int s = tid; s < nseg; s += 32) {
            const SegWork& w = s_work[s];
            if (!(w.tma && s_act[s])) continue;
            for (int r = 0; r < w.nrun; r++) {
                const int qs = w.run_t0[r], off = w.run_off[r];
                const uint32_t bytes = (uint32_t)(((w.run_hi[r] - qs) + 3) & ~3) * 8u;
                tma_load_1d(tX + off, pXNU + qs, bytes, &s_all_bar);
                tma_load_1d(tH + off, pH2 + qs, bytes, &s_all_bar);
                tma_load_1d(tC + off, pCN + qs, bytes, &s_all_bar);
                tma_load_1d(tP + off, pP3 + qs, bytes, &s_all_bar);
            }
        }
    }
    // piece id of every staged line (0xff: alignment padding or a molecule with zero column amount)
    for (int p = wid; p < npc; p += NW) {
        const NearPiece pc = s_np[p];
        if (!s_act[pc.info & 0xff]) continue;
        for (int i = lane; i < pc.n; i += 32) s_pid[pc.soff + i] = (unsigned char)p;
    }
    mbar_wait(&s_all_bar, 0u);
    __syncthreads();
    if (!by_mol) {
        // one sum over all molecules: strengths and pedestals carry the column amount W (o = RFT*sum_mol W_mol*SF_mol)
        for (int j = tid; j < total; j += NT) {
            const int pid = s_pid[j];
            if (pid == 0xff) continue;
            const double w = s_wseg[s_np[pid].info & 0xff];
            tC[j] *= w;
            tP[j] *= w;
        }
        __syncthreads();
    }

    // ---- this warp's frequencies: a contiguous block of 32*F
    const int base = blockIdx.x * (NT * F) + wid * (32 * F);
    double wn[F];
    bool valid[F];
    double wA = 1e300, wB = -1e300;
#pragma unroll
    for (int f = 0; f < F; f++) {
        const int iw = base + f * 32 + lane;
        valid[f] = iw < a.nwn;
        wn[f] = a.wn[valid[f] ? iw : (a.nwn - 1)];
        if (valid[f]) { wA = fmin(wA, wn[f]); wB = fmax(wB, wn[f]); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        wA = fmin(wA, __shfl_xor_sync(0xffffffffu, wA, off));
        wB = fmax(wB, __shfl_xor_sync(0xffffffffu, wB, off));
    }
    if (wB < wA) return;                           // no frequency in this warp's block (tail tile); no barrier follows
    int nvalid_w = 0;
    if (a.counters) {
        const int rem = a.nwn - base;
        nvalid_w = rem < 32 * F ? rem : 32 * F;
    }
    const double cen = 0.5 * (wA + wB), hh = 0.5 * (wB - wA);
    const double hinv = hh > 0. ? 1. / hh : 0.;
    const double Rn = a.ffw_ratio * hh, R2 = Rn * Rn;
    const double m2h = -2. * hh, mhh = -hh * hh;
    const double rp = ly.rp, rp2 = ly.rp2;
    const int vmode_mask = voigt_possible(a.layer_voigt, L, th.whi) ? 0xff : (0xff & ~M_VOIGT);
    // a line is in at most one list: D1 and T1 grow from the front of their array, D2 and T2 from the back
    unsigned short* lD1 = s_list + (size_t)(wid * 2 + 0) * kListLen;
    unsigned short* lD2 = lD1 + (kListLen - 1);
    unsigned short* lT1 = s_list + (size_t)(wid * 2 + 1) * kListLen;
    unsigned short* lT2 = lT1 + (kListLen - 1);
    const unsigned lt_mask = (1u << lane) - 1u;

    double osum[F];
    long long cnt[F];
    unsigned long long hsh[F];
#pragma unroll
    for (int f = 0; f < F; f++) { osum[f] = 0.; cnt[f] = 0; hsh[f] = 0ull; }
    int err = 0;
    long long n_direct = 0, n_far = 0;

    // selection bookkeeping that does not depend on the staged lines
    if (SEL) {
        for (int s = 0; s < nseg; s++) {
            if (!s_act[s]) continue;
            const Segment sg = a.seg[s];
            if (sg.mol == 7) {                             // every O2 line passes modm.f90:384
#pragma unroll
                for (int f = 0; f < F; f++) { cnt[f] += sg.count_all; hsh[f] += sg.hash_all; }
            } else if (sg.cls == CLS_PED) {                // every far line (any level) is inside the window of every frequency
                const SegWork& wk = s_work[s];
                for (int u = 0; u + 1 < wk.nbp; u++) {
                    if (wk.mode[u] != 0) continue;
                    const int lo = wk.bp[u], hi = wk.bp[u + 1];
                    const unsigned long long hs = a.keypre[hi] - a.keypre[lo];
#pragma unroll
                    for (int f = 0; f < F; f++) { cnt[f] += hi - lo; hsh[f] += hs; }
                }
            }
        }
    }

    // ---- groups: all molecules at once, or one molecule at a time when per-molecule outputs are requested
    int s_beg = 0, p_beg = 0;
    while (s_beg < nseg) {
        int s_end = nseg, p_end = npc;
        const int mol = a.seg[s_beg].mol;
        if (by_mol) {
            s_end = s_beg;
            while (s_end < nseg && a.seg[s_end].mol == mol) s_end++;
            p_end = p_beg;
            while (p_end < npc && (s_np[p_end].info & 0xff) < s_end) p_end++;
        }
        double sf[F];
#pragma unroll
        for (int f = 0; f < F; f++) sf[f] = 0.;

        if (p_end > p_beg) {
            const int j0 = s_np[p_beg].soff & ~31;
            const NearPiece plast = s_np[p_end - 1];
            const int j1 = plast.soff + plast.n;
            int nD1 = 0, nD2 = 0, nT1 = 0, nT2 = 0;
            double ped_acc = 0.;
            long long cnt_u = 0;
            unsigned long long hsh_u = 0ull;
            double A[kFarK];
#pragma unroll
            for (int i = 0; i < kFarK; i++) A[i] = 0.;
            bool any_far = false;
            for (int jb = j0; jb < j1; jb += 32) {
                const int j = jb + lane;
                int pid = (j < j1) ? (int)s_pid[j] : 0xff;
                if (pid != 0xff && (pid < p_beg || pid >= p_end)) pid = 0xff;
                const bool ok = pid != 0xff;
                const NearPiece pc = s_np[ok ? pid : p_beg];
                const int mode = ((pc.info >> 8) & 0xff) & vmode_mask;
                const int kind = (pc.info >> 17) & 3;
                bool both = ((pc.info >> 16) & 1) != 0;
                const double x = ok ? tX[j] : 0., h2 = ok ? tH[j] : 1., c = ok ? tC[j] : 0., pd = ok ? tP[j] : 0.;
                const double dA = wA - x, dB = wB - x;
                bool skip = !ok, test = false, neg_possible = both;
                if (ok && (mode & M_EDGE)) {                                 // kind 0/1: the window test can fail
                    const bool all_out = (dA > kDELTNUC) || (dB < -kDELTNUC);
                    const bool all_in = !(fabs(dA) > kDELTNUC) && !(fabs(dB) > kDELTNUC);
                    if (all_out) skip = true;
                    else if (!all_in) test = true;
                }
                if (ok && (mode & M_NEG)) {
                    if ((wB + x) <= kDELTNUC) { both = true; neg_possible = true; }
                    else if ((wA + x) > kDELTNUC) { both = false; neg_possible = false; }
                    else { test = true; neg_possible = true; }
                }
                if (ok && !skip && (mode & M_VOIGT)) {
                    const double vt = __ldg(pVT + pc.q0 + (j - pc.soff));
                    if (vt >= 0.) {
                        const double mind = (dA <= 0. && dB >= 0.) ? 0. : fmin(fabs(dA), fabs(dB));
                        if (mind <= vt) test = true;
                    }
                }
                const bool plain = ok && !skip && !test;
                const double Dm = cen - x, Dp = cen + x;
                const bool far = plain && (fma(Dm, Dm, h2) >= R2) && (!both || (fma(Dp, Dp, h2) >= R2));
                // in-window for every frequency of the block: pedestal and selection bookkeeping once per line
                if (plain) {
                    ped_acc += both ? 2. * pd : pd;
                    if (SEL && kind == 0) { cnt_u++; hsh_u += a.key[pc.q0 + (j - pc.soff)]; }
                }
                const unsigned mfar = __ballot_sync(0xffffffffu, far);
                if (mfar) {
                    any_far = true;
                    const unsigned mboth = __ballot_sync(0xffffffffu, far && both);
                    const double wf = far ? c : 0.;
                    if (mboth) far_accum2(far ? Dm : 1., far ? h2 : 1., wf, 0., (far && both) ? Dp : 1., (far && both) ? h2 : 1., (far && both) ? c : 0., 0., m2h, mhh, A);
                    else far_accum(far ? Dm : 1., far ? h2 : 1., wf, 0., m2h, mhh, A);
                    if (a.counters) n_far += __popc(mfar) + __popc(mboth);
                }
                const bool d1 = plain && !far && !both, d2 = plain && !far && both;
                const bool t1 = ok && !skip && test && !neg_possible, t2 = ok && !skip && test && neg_possible;
                const unsigned m1 = __ballot_sync(0xffffffffu, d1), m2 = __ballot_sync(0xffffffffu, d2);
                const unsigned m3 = __ballot_sync(0xffffffffu, t1), m4 = __ballot_sync(0xffffffffu, t2);
                if (d1) lD1[nD1 + __popc(m1 & lt_mask)] = (unsigned short)j;
                if (d2) lD2[-(nD2 + __popc(m2 & lt_mask))] = (unsigned short)j;
                if (t1) lT1[nT1 + __popc(m3 & lt_mask)] = (unsigned short)j;
                if (t2) lT2[-(nT2 + __popc(m4 & lt_mask))] = (unsigned short)j;
                nD1 += __popc(m1); nD2 += __popc(m2); nT1 += __popc(m3); nT2 += __popc(m4);
            }
            if (lane < 3) lD1[nD1 + lane] = (unsigned short)kDummy;       // pad to a group of four
            // warp sums: pedestals, selection, far-field coefficients
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ped_acc += __shfl_xor_sync(0xffffffffu, ped_acc, off);
            if (SEL) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    cnt_u += __shfl_xor_sync(0xffffffffu, cnt_u, off);
                    hsh_u += __shfl_xor_sync(0xffffffffu, hsh_u, off);
                }
#pragma unroll
                for (int f = 0; f < F; f++) { cnt[f] += cnt_u; hsh[f] += hsh_u; }
            }
            if (any_far) {
#pragma unroll
                for (int i = 0; i < kFarK; i++) {
                    double v = A[i];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    A[i] = v;
                }
#pragma unroll
                for (int f = 0; f < F; f++) {
                    const double sv = (wn[f] - cen) * hinv;
                    double p = A[kFarK - 1];
#pragma unroll
                    for (int i = kFarK - 2; i >= 0; i--) p = fma(p, sv, A[i]);
                    sf[f] = p;
                }
            }
#pragma unroll
            for (int f = 0; f < F; f++) sf[f] -= ped_acc;
            __syncwarp();
            if (a.counters) n_direct += (long long)(nD1 + nD2 + nT1 + nT2) * nvalid_w;

            // ---- D1: single resonance (modm.f90:751); four lines share one reciprocal,
            // sum c_i/a_i = N/(a1 a2 a3 a4), N = (c1 a2 + c2 a1)(a3 a4) + (c3 a4 + c4 a3)(a1 a2)
            {
                double psum[F];
#pragma unroll
                for (int f = 0; f < F; f++) psum[f] = 0.;
                for (int g = 0; g < nD1; g += 4) {
                    const uint2 iv = *reinterpret_cast<const uint2*>(lD1 + g);
                    const int i1 = iv.x & 0xffff, i2 = iv.x >> 16, i3 = iv.y & 0xffff, i4 = iv.y >> 16;
                    const double x1 = tX[i1], x2 = tX[i2], x3 = tX[i3], x4 = tX[i4];
                    const double g1 = tH[i1], g2 = tH[i2], g3 = tH[i3], g4 = tH[i4];
                    const double c1 = tC[i1], c2 = tC[i2], c3 = tC[i3], c4 = tC[i4];
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const double d1v = wn[f] - x1, d2v = wn[f] - x2, d3v = wn[f] - x3, d4v = wn[f] - x4;
                        const double a1 = fma(d1v, d1v, g1), a2 = fma(d2v, d2v, g2);
                        const double a3 = fma(d3v, d3v, g3), a4 = fma(d4v, d4v, g4);
                        const double p12 = a1 * a2, p34 = a3 * a4;
                        const double n12 = fma(c1, a2, c2 * a1), n34 = fma(c3, a4, c4 * a3);
                        const double r = rcp3(p12 * p34);
                        psum[f] = fma(fma(n12, p34, n34 * p12), r, psum[f]);
                    }
                }
                // ---- D2: both resonances: cn*(1/a+1/b) = cn*(a+b)/(a*b), one reciprocal
MRTM_UNROLL(MRTM_UNROLL_BOTH)
                for (int g = 0; g < nD2; g++) {
                    const int j = lD2[-g];
                    const double xnu = tX[j], h2 = tH[j], cn = tC[j];
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                        const double aa = fma(dm, dm, h2), bb = fma(sp, sp, h2);
                        const double r = rcp3(aa * bb);
                        psum[f] = fma(cn * (aa + bb), r, psum[f]);
                    }
                }
#pragma unroll
                for (int f = 0; f < F; f++) sf[f] += psum[f];
            }
            // ---- T1 / T2: the reference's exact per-(line,frequency) tests (modm.f90:384, 427, 746)
            for (int g = 0; g < nT1; g++) {
                const int j = lT1[g];
                const NearPiece pc = s_np[s_pid[j]];
                const int q = pc.q0 + (j - pc.soff);
                const int kind = (pc.info >> 17) & 3;
                const bool has_win = kind != 2, count_sel = SEL && (kind == 0);
                const double xnu = tX[j], h2 = tH[j], cn = tC[j], ped = tP[j];
                const double vt = (((pc.info >> 8) & vmode_mask) & M_VOIGT) ? __ldg(pVT + q) : -1.0;
#pragma unroll
                for (int f = 0; f < F; f++) {
                    const double dm = wn[f] - xnu;
                    const bool inwin = has_win ? !(fabs(dm) > kDELTNUC) : true;
                    if (count_sel && inwin) { cnt[f]++; hsh[f] += a.key[q]; }
                    const bool take = inwin && !(fabs(dm) <= vt);          // Voigt-branch pairs: voigt_kernel
                    const double val = fma(cn, rcp3(fma(dm, dm, h2)), -ped);
                    sf[f] += take ? val : 0.;
                }
            }
            for (int g = 0; g < nT2; g++) {
                const int j = lT2[-g];
                const NearPiece pc = s_np[s_pid[j]];
                const int q = pc.q0 + (j - pc.soff);
                const int kind = (pc.info >> 17) & 3;
                const bool has_win = kind != 2, count_sel = SEL && (kind == 0);
                const bool negall = kind == 2;
                const double xnu = tX[j], h2 = tH[j], cn = tC[j], ped = tP[j];
                const double vt = (((pc.info >> 8) & vmode_mask) & M_VOIGT) ? __ldg(pVT + q) : -1.0;
#pragma unroll
                for (int f = 0; f < F; f++) {
                    const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                    const bool inwin = has_win ? !(fabs(dm) > kDELTNUC) : true;
                    if (count_sel && inwin) { cnt[f]++; hsh[f] += a.key[q]; }
                    const bool take = inwin && !(fabs(dm) <= vt);
                    const bool neg = negall || (sp <= kDELTNUC);
                    const double r1 = rcp3(fma(dm, dm, h2)), r2 = rcp3(fma(sp, sp, h2));
                    const double val = cn * (r1 + (neg ? r2 : 0.)) - (neg ? 2. * ped : ped);
                    sf[f] += take ? val : 0.;
                }
            }
            __syncwarp();          // the lists are rebuilt by the next group
        }

        // ---- classes that are not staged: first-order O2 mixing (few lines) and the general case tree
        for (int s = s_beg; s < s_end; s++) {
            if (!s_act[s]) continue;
            const Segment sg = a.seg[s];
            const int cls = sg.cls;
            if (cls != CLS_O2_LC1 && cls != CLS_GENERAL) continue;
            const SegWork& wk = s_work[s];
            double ss[F];
#pragma unroll
            for (int f = 0; f < F; f++) ss[f] = 0.;
            if (cls == CLS_O2_LC1) {
                for (int r = 0; r < wk.nrun; r++) {
                    if (a.counters) n_direct += (long long)(wk.run_hi[r] - wk.run_lo[r]) * nvalid_w;
                    for (int q = wk.run_lo[r]; q < wk.run_hi[r]; q++) {
                        const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q], cq = pP4[q], vt = pVT[q];
#pragma unroll
                        for (int f = 0; f < F; f++) {
                            const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                            const double r1 = rcp3(fma(dm, dm, h2));
                            const double r2 = rcp3(fma(sp, sp, h2));
                            const double val = fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;
                            ss[f] += (fabs(dm) <= vt) ? 0. : val;           // Voigt-branch pairs: voigt_kernel
                        }
                    }
                }
            } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
                if (a.counters) n_direct += (long long)(wk.q1 - wk.q0) * nvalid_w;
                for (int q = wk.q0; q < wk.q1; q++) {
                    const double xnu = pXNU[q], vt = pVT[q];
                    const double hw = pl[(size_t)D_H * a.n_pad + q], ad = pl[(size_t)D_AD * a.n_pad + q];
                    const double st = pl[(size_t)D_STILD * a.n_pad + q];
                    const double aip = pl[(size_t)D_AIP * a.n_pad + q], bip = pl[(size_t)D_BIP * a.n_pad + q];
                    const int xf = a.xf_s[q];
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const double dm = wn[f] - xnu;
                        if ((fabs(dm) > kDELTNUC) && (sg.mol != 7)) continue;          // modm.f90:384
                        if (SEL && sg.mol != 7) { cnt[f]++; hsh[f] += a.key[q]; }
                        const bool voigt = fabs(dm) <= vt;
                        ss[f] += st * lsf_general(sg.mol, xf, rp, rp2, aip, bip, hw, wn[f], xnu, ad, a.sdep_s[q], voigt, &err);
                    }
                }
            }
            const double w = by_mol ? 1. : s_wseg[s];
#pragma unroll
            for (int f = 0; f < F; f++) sf[f] = fma(w, ss[f], sf[f]);
        }

        // ---- close the group: W*SF; RFT is applied by final_kernel (modm.f90:436-438, 265-267)
        if (by_mol) {
            const double w = ly.wk[mol - 1];
#pragma unroll
            for (int f = 0; f < F; f++) {
                const double ol = (w == 0.) ? 0. : (w * sf[f]);
                osum[f] = osum[f] + ol;
                if (valid[f]) {
                    const int iw = base + f * 32 + lane;
                    a.o_by_mol[(size_t)iw + (size_t)(mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] = ol;
                }
            }
        } else {
#pragma unroll
            for (int f = 0; f < F; f++) osum[f] += sf[f];
        }
        s_beg = s_end;
        p_beg = p_end;
    }
    if (err) atomicOr(a.errflag, 2);
    if (a.counters && lane == 0) {
        atomicAdd(a.counters + 0, (unsigned long long)n_far);
        atomicAdd(a.counters + 1, (unsigned long long)n_direct);
    }
#pragma unroll
    for (int f = 0; f < F; f++) {
        if (!valid[f]) continue;
        const int iw = base + f * 32 + lane;
        const size_t fl = (size_t)iw + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
        a.o[fl] = osum[f];
        if (SEL) {
            if (a.sel_count) a.sel_count[fl] = cnt[f];
            if (a.sel_hash) a.sel_hash[fl] = hsh[f];
        }
    }
}

// =============================================================================================
// voigt_kernel: the Voigt branch (modm.f90:427-431).  CTA = (frequency tile, layer, profile); it leaves
// at once when the layer has no Voigt-capable line.  For the lines of the plan's Voigt zones it applies
// the reference's test |WN-Xnu| <= 100*HWHM_D per (line, frequency) and adds W*STILD*SLS of the pairs that
// pass (the near kernels skipped exactly those) to O [and O_BY_MOL].
// Each warp owns a contiguous block of 32*F frequencies and walks it in F sub-blocks of 32.  Per sub-block
// the lanes first cull the staged zone lines against the sub-block's frequency extent (one line per lane,
// exact: the rounded difference WN-Xnu is monotone in WN) into a compact list, so the per-(line,frequency)
// loop only visits lines whose zone reaches the sub-block.
// =============================================================================================
#ifndef MRTM_VOIGT_MINB
#define MRTM_VOIGT_MINB 8
#endif
template <int F, int NT>
__global__ void __launch_bounds__(NT, MRTM_VOIGT_MINB) voigt_kernel(LinesArgs a)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.y, prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    if (!voigt_possible(a.layer_voigt, L, a.hdr[0][blockIdx.x].whi)) return;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;
    const double* __restrict__ pH = pl + (size_t)D_H * a.n_pad;
    const double* __restrict__ pAD = pl + (size_t)D_AD * a.n_pad;
    const double* __restrict__ pST = pl + (size_t)D_STILD * a.n_pad;
    const double* __restrict__ pAIP = pl + (size_t)D_AIP * a.n_pad;
    const double* __restrict__ pBIP = pl + (size_t)D_BIP * a.n_pad;
    const SegWork* plan = a.plan[0] + (size_t)blockIdx.x * a.nseg;
    // The zone lines of all segments are staged together (one barrier pair per CTA in the usual case), and per zone
    // line, once per CTA (amortised over the NT*F frequencies), everything of LSF_SDVOIGT/SDVOIGT that does not depend
    // on the frequency: 1/alphaD, y = sqrt(ln2)*alphaL/alphaD, STILD*sqrt(ln2/pi)/alphaD, the Voigt pedestal at
    // 25 cm-1 (modm.f90:590) and the mixing factors (:595-596).
    constexpr int kVCap = 192;
    __shared__ double s_vt[kVCap], s_x[kVCap], s_inv[kVCap], s_y[kVCap], s_c[kVCap], s_pd[kVCap], s_g[kVCap], s_b[kVCap];
    __shared__ int s_q[kVCap];
    __shared__ unsigned char s_kind[kVCap], s_mol[kVCap];
    // fast path (generic uncoupled line, single resonance, Humlicek region I): Re w = y*(a+q)/(q*(q+b)+a*a), q = x*x,
    // a = .5+y*y, b = 2*y*y-1 -- the reference's t*.5641896/(.5+t*t) (modm.f90:1105) multiplied out
    __shared__ double s_fa[kVCap], s_fb[kVCap], s_fa2[kVCap], s_fcy[kVCap], s_fcpd[kVCap];
    __shared__ unsigned short s_list[NW][kVCap];
    __shared__ int s_zlo[kMaxSegments], s_zoff[kMaxSegments + 1];
    const int base = blockIdx.x * (NT * F) + wid * (32 * F);
    const double rp = ly.rp, rp2 = ly.rp2;
    const double sl2 = 0.8325546111576977;         // sqrt(log(2))
    const bool by_mol = a.o_by_mol != nullptr;
    const unsigned lt_mask = (1u << lane) - 1u;
    // zone directory: entry e of the CTA's zone list lies in segment s with s_zoff[s] <= e < s_zoff[s+1]
    // (one thread per segment fetches its zone from the plan, then one thread sums the counts in shared memory)
    for (int s = tid; s < a.nseg; s += NT) {
        const Segment sg = a.seg[s];
        const bool use = (sg.cls != CLS_GENERAL) && (ly.wk[sg.mol - 1] != 0.);
        const int v0 = plan[s].v0, v1 = plan[s].v1;
        s_zlo[s] = v0;
        s_zoff[s + 1] = (use && v1 > v0) ? (v1 - v0) : 0;
    }
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int s = 0; s < a.nseg; s++) {
            const int c = s_zoff[s + 1];
            s_zoff[s] = tot;
            tot += c;
        }
        s_zoff[a.nseg] = tot;
    }
    __syncthreads();
    const int total = s_zoff[a.nseg];
    if (total == 0) return;
    int err = 0;
    unsigned short* lst = s_list[wid];
    double* const odst = (a.o_v ? a.o_v : a.o) + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
    for (int e0 = 0; e0 < total; e0 += kVCap) {
        const int n = min(kVCap, total - e0);
        if (e0 > 0) __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const int e = e0 + i;
            int sg = 0;
            while (s_zoff[sg + 1] <= e) sg++;
            const int q = s_zlo[sg] + (e - s_zoff[sg]);
            const int cls = a.seg[sg].cls, mol = a.seg[sg].mol;
            const int kind = (cls == CLS_PED) ? 0 : ((cls == CLS_O2) ? 1 : ((cls == CLS_O2_LC35) ? 2 : 3));
            const double vt = __ldg(pVT + q);
            s_vt[i] = vt;
            s_x[i] = __ldg(pXNU + q);
            s_q[i] = q;
            s_kind[i] = (unsigned char)kind;
            s_mol[i] = (unsigned char)mol;
            if (vt >= 0.) {
                const double hw = __ldg(pH + q), ad = __ldg(pAD + q);
                const double zeta = hw / (hw + ad);
                if (fabs(__ldg(a.sdep_s + q)) > 1.0e-4 || !(zeta < 1.0)) {
                    s_inv[i] = -1.;        // speed dependence / degenerate Doppler width: the general routine per pair
                    s_c[i] = by_mol ? 1. : ly.wk[mol - 1];
                } else {
                    const double inv = 1. / ad;
                    const double y = sl2 * (hw * inv);
                    const double wgt = by_mol ? 1. : ly.wk[mol - 1];           // one sum over all molecules: weight folded in
                    s_inv[i] = inv;
                    s_y[i] = y;
                    s_c[i] = wgt * (__ldg(pST + q) * (0.46971863934982516 * inv));   // sqrt(log(2)/PI), 13-digit PI
                    s_pd[i] = (kind == 0) ? w4_re_fast(sl2 * (kDELTNUC * inv), y) : 0.;
                    s_g[i] = (kind == 3) ? (__ldg(pAIP + q) * (1 / hw) * rp) : 0.;
                    s_b[i] = (kind == 3) ? (__ldg(pBIP + q) * rp2) : 0.;
                    if (kind == 0 && vt <= kDELTNUC) {      // inside the zone the window test cannot fail
                        const double y2 = y * y, aa = .5 + y2;
                        s_fa[i] = aa;
                        s_fb[i] = 2. * y2 - 1.;
                        s_fa2[i] = aa * aa;
                        s_fcy[i] = s_c[i] * (.5641896 * y);
                        s_fcpd[i] = s_c[i] * s_pd[i];
                        s_kind[i] = (unsigned char)(kind | 0x80);
                    }
                }
            }
        }
        __syncthreads();
        for (int f = 0; f < F; f++) {              // not unrolled: four copies of the loop body run slower
            const int iw = base + f * 32 + lane;
            const bool valid = iw < a.nwn;
            const double wn = a.wn[valid ? iw : (a.nwn - 1)];
            double wA = valid ? wn : 1e300, wB = valid ? wn : -1e300;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                wA = fmin(wA, __shfl_xor_sync(0xffffffffu, wA, off));
                wB = fmax(wB, __shfl_xor_sync(0xffffffffu, wB, off));
            }
            if (wB < wA) break;                    // no frequency in this sub-block (nor in the following ones)
            // cull: lines whose zone cannot reach [wA, wB] fail the test for every lane
            int nl = 0;
            for (int ib = 0; ib < n; ib += 32) {
                const int i = ib + lane;
                bool hit = false;
                int ent = i;
                if (i < n) {
                    const double vt = s_vt[i];
                    if (vt >= 0.) {
                        const double x = s_x[i];
                        hit = !((wB - x) < -vt) && !((wA - x) > vt);
                        // fast entry: plain line and no frequency of the sub-block has the second resonance
                        // (WN+Xnu-25 <= 0, modm.f90:746; the rounded sum is monotone in WN)
                        if ((s_kind[i] & 0x80) && ((wA + x) - kDELTNUC) > 0.) ent |= 0x100;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) lst[nl + __popc(m & lt_mask)] = (unsigned short)ent;
                nl += __popc(m);
            }
            __syncwarp();
            if (nl == 0) continue;                 // warp-uniform: no zone reaches this sub-block
            // the optical depth this sub-block adds to: loaded now, needed after the evaluation loop
            const double oprev = valid ? odst[iw] : 0.;
            double vsum = 0., msum = 0.;
            int cur_mol = -1;
            bool many = false;
            auto flush_mol = [&]() {      // per-molecule outputs: close the molecule's sum
                if (cur_mol > 0 && many) {
                    const double ol = ly.wk[cur_mol - 1] * msum;
                    vsum += ol;
                    if (valid && msum != 0.)
                        a.o_by_mol[(size_t)iw + (size_t)(cur_mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] += ol;
                    msum = 0.;
                }
                many = false;
            };
            auto one = [&](const int ent) {
                const int i = ent & 0xff;
                if (by_mol && (int)s_mol[i] != cur_mol) {
                    flush_mol();
                    cur_mol = s_mol[i];
                }
                const double xnu = s_x[i];
                const double dm = wn - xnu;
                if (ent & 0x100) {
                    if (fabs(dm) <= s_vt[i]) {
                        const double y = s_y[i];
                        const double x = sl2 * (dm * s_inv[i]);
                        if (!(fabs(x) + y < 15.)) {
                            const double q = x * x;
                            const double den = fma(q, q + s_fb[i], s_fa2[i]);
                            msum += fma(s_fcy[i] * (s_fa[i] + q), rcp3(den), -s_fcpd[i]);
                        } else {
                            msum = fma(s_c[i], w4_re_near(x, y), msum) - s_fcpd[i];
                        }
                        many = true;
                    }
                    return;
                }
                const int kind = s_kind[i] & 0x7f;
                const bool inwin = (kind <= 1) ? !(fabs(dm) > kDELTNUC) : true;
                if (inwin && fabs(dm) <= s_vt[i]) {
                    const double inv = s_inv[i];
                    if (inv < 0.) {
                        msum = fma(s_c[i], voigt_lines_term(kind, wn, xnu, pl, a.n_pad, s_q[i], a.sdep_s[s_q[i]], rp, rp2, &err), msum);
                    } else {
                        const double y = s_y[i], sp = wn + xnu;
                        const bool second = (kind >= 2) || ((sp - kDELTNUC) <= 0.);
                        double sls = w4_re_fast(sl2 * (dm * inv), y);
                        if (kind == 3) sls *= (1. + (s_g[i] * dm) + s_b[i]);
                        if (second) {
                            double v2 = w4_re_fast(sl2 * (sp * inv), y);
                            if (kind == 3) v2 *= (1. - (s_g[i] * sp) + s_b[i]);
                            sls += v2;
                        }
                        if (kind == 0) sls -= (second ? 2. : 1.) * s_pd[i];
                        msum = fma(s_c[i], sls, msum);
                    }
                    many = true;
                }
            };
            // two fast entries per step when possible: the region-I arithmetic of both is independent (it is done for
            // every lane and selected afterwards), which hides the shared-memory and FP64 latencies of the serial walk
            int g = 0;
            if (!by_mol) {
                for (; g + 1 < nl; g += 2) {
                    const int e0 = lst[g], e1 = lst[g + 1];
                    if (!(e0 & e1 & 0x100)) { one(e0); one(e1); continue; }
                    const int i0 = e0 & 0xff, i1 = e1 & 0xff;
                    const double dm0 = wn - s_x[i0], dm1 = wn - s_x[i1];
                    const bool in0 = fabs(dm0) <= s_vt[i0], in1 = fabs(dm1) <= s_vt[i1];
                    const double y0 = s_y[i0], y1 = s_y[i1];
                    const double x0 = sl2 * (dm0 * s_inv[i0]), x1 = sl2 * (dm1 * s_inv[i1]);
                    const bool r0 = !(fabs(x0) + y0 < 15.), r1 = !(fabs(x1) + y1 < 15.);
                    const double q0 = x0 * x0, q1 = x1 * x1;
                    const double den0 = fma(q0, q0 + s_fb[i0], s_fa2[i0]), den1 = fma(q1, q1 + s_fb[i1], s_fa2[i1]);
                    const double v0 = fma(s_fcy[i0] * (s_fa[i0] + q0), rcp3(den0), -s_fcpd[i0]);
                    const double v1 = fma(s_fcy[i1] * (s_fa[i1] + q1), rcp3(den1), -s_fcpd[i1]);
                    if (in0) msum = r0 ? (msum + v0) : (fma(s_c[i0], w4_re_near(x0, y0), msum) - s_fcpd[i0]);
                    if (in1) msum = r1 ? (msum + v1) : (fma(s_c[i1], w4_re_near(x1, y1), msum) - s_fcpd[i1]);
                    many = many || in0 || in1;
                }
            }
            for (; g < nl; g++) one(lst[g]);
            if (by_mol) flush_mol(); else vsum = msum;
            if (valid && vsum != 0.) odst[iw] = oprev + vsum;
            __syncwarp();                          // the list is rebuilt by the next sub-block
        }
    }
    if (err) atomicOr(a.errflag, 2);
}

// =============================================================================================
// final_kernel: per (frequency, layer): the far-field polynomial of the level-0 tile, RFT (modm.f90:257),
// the continuum interpolation + RADFN (:218-230), cloud liquid water (:264) and the total (:265-269).
// =============================================================================================
#ifndef MRTM_FINAL_MINB
#define MRTM_FINAL_MINB 8
#endif
template <int F, int NT>
__global__ void __launch_bounds__(NT, MRTM_FINAL_MINB) final_kernel(LinesArgs a)
{
    const int tid = threadIdx.x;
    const int k = blockIdx.y, prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    const int64_t Ltot = (int64_t)gridDim.y * gridDim.z;
    const LayerDev& ly = a.lay[L];
    __shared__ double s_coef[kFarK];
    const int base = blockIdx.x * (NT * F);
    double wn[F], sv[F];
    bool valid[F];
    const bool have_far = a.coef[0] != nullptr;
    double cen = 0., hinv = 0.;
    if (have_far) {
        const TileHdr th = a.hdr[0][blockIdx.x];
        const double hh = 0.5 * (th.whi - th.wlo);
        cen = 0.5 * (th.wlo + th.whi);
        hinv = hh > 0. ? 1. / hh : 0.;
    }
#pragma unroll
    for (int f = 0; f < F; f++) {
        int iw = base + f * NT + tid;
        valid[f] = iw < a.nwn;
        wn[f] = a.wn[valid[f] ? iw : (a.nwn - 1)];
        sv[f] = (wn[f] - cen) * hinv;
    }
    const double* cf = have_far ? a.coef[0] + ((size_t)blockIdx.x * Ltot + L) * a.nslot * kFarK : nullptr;
    double osum[F];
    if (!a.o_by_mol) {
        // one polynomial: sum over molecules of W_mol * coefficients (molecule order)
        if (have_far && tid < kFarK) s_coef[tid] = cf[tid];        // far_kernel ran in combined mode (nslot == 1)
        __syncthreads();
#pragma unroll
        for (int f = 0; f < F; f++) {
            const int iw = base + f * NT + tid;
            const size_t fl = (size_t)iw + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
            double v = valid[f] ? a.o[fl] : 0.;
            if (a.o_v && valid[f]) v += a.o_v[fl];
            if (have_far) {
                double p = s_coef[kFarK - 1];
#pragma unroll
                for (int i = kFarK - 2; i >= 0; i--) p = fma(p, sv[f], s_coef[i]);
                v += p;
            }
            const double rft = wn[f] * tanh((ly.radct * wn[f]) / (2 * ly.t));   // modm.f90:257
            osum[f] = rft * v;
        }
    } else {
#pragma unroll
        for (int f = 0; f < F; f++) osum[f] = 0.;
        for (int sl = 0; sl < a.nslot; sl++) {
            const int mol = a.slot_mol[sl];
            const double w = ly.wk[mol - 1];
            double c[kFarK];
#pragma unroll
            for (int i = 0; i < kFarK; i++) c[i] = have_far ? __ldg(cf + (size_t)sl * kFarK + i) : 0.;
#pragma unroll
            for (int f = 0; f < F; f++) {
                if (!valid[f]) continue;
                const int iw = base + f * NT + tid;
                const size_t idx = (size_t)iw + (size_t)(mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk;
                double p = c[kFarK - 1];
#pragma unroll
                for (int i = kFarK - 2; i >= 0; i--) p = fma(p, sv[f], c[i]);
                const double rft = wn[f] * tanh((ly.radct * wn[f]) / (2 * ly.t));
                const double ol = (w == 0.) ? 0. : rft * (a.o_by_mol[idx] + w * p);     // modm.f90:436-438
                a.o_by_mol[idx] = ol;
                osum[f] = osum[f] + ol;                                              // :265-267 (molecule order)
            }
        }
    }

    // ---- epilogue: continuum, cloud, totals ---------------------------------------------------
    const double* ab = a.absrb + (size_t)L * 3 * a.nptabs_pad;
    const int cont_mol[3] = {1, 2, 22};
#pragma unroll
    for (int f = 0; f < F; f++) {
        if (!valid[f]) continue;
        const int iw = base + f * NT + tid;
        const size_t fl = (size_t)iw + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
        double soc = 0.;
        // gridded mode interpolates at V1+DVSET*(I-1) inside [ILO,IHI] (modm.f90:218-219), list mode at WN
        double vi = wn[f];
        bool in_rng = true;
        if (a.dvset != 0.) {
            const long long I = a.iw0 + iw + 1;
            vi = a.v1 + a.dvset * (double)(I - 1);
            long long ilo = (long long)((a.v1abs + 1.0 - a.v1) / a.dvset + 1. + 0.999);
            long long ihi = (long long)((a.v2abs - 1.0 - a.v1) / a.dvset + 0.999);
            in_rng = (I >= (ilo > 1 ? ilo : 1)) && (I <= ihi);
        } else {
            long long ilo = (long long)((a.v1abs + 1.0 - vi) / 1.0 + 1. + 0.999);
            long long ihi = (long long)((a.v2abs - 1.0 - vi) / 1.0 + 0.999);
            in_rng = (ilo <= 1) && (ihi >= 1);
        }
        const double rf = radfn(wn[f], ly.xkt);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double v = 0.;
            if (in_rng) v = 0. + xint_point(ab + (size_t)c * a.nptabs_pad, a.v1abs, 1.0, vi) * 1.0;
            v = v * rf;
            soc = soc + v;                                             // sum(oc(m,1:22,k)) in index order
            if (a.oc) a.oc[(size_t)iw + (size_t)(cont_mol[c] - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] = v;
        }
        double oclw = (ly.clw == 0.) ? 0. : odclw_tkc(wn[f], ly.t, ly.clw);   // modm.f90:264
        double odx = a.odxsec ? a.odxsec[fl] : 0.;
        double tot = osum[f] + odx + 0. + soc + oclw;                  // :268-269 (oc_rayl = 0 for V2 < 820)
        a.o[fl] = tot;
        if (a.o_clw) a.o_clw[fl] = oclw;
    }
}

// =============================================================================================
// colsum_kernel: otot_by_mol(im, iw) = sum over layers of o_by_mol(iw,im,k)+oc(iw,im,k)
// (STOREOUT, src/monortm_sub.F90:649-656), layers added in index order.
// =============================================================================================
__global__ void colsum_kernel(int nwn, int nlay, const double* o_by_mol, const double* oc,
                              int64_t ldm, int64_t ldk, double* otot_by_mol /* (39,nwn) */)
{
    int iw = blockIdx.x * blockDim.x + threadIdx.x;
    int im = blockIdx.y;
    if (iw >= nwn) return;
    double s = 0.;
    for (int k = 0; k < nlay; k++) {
        size_t idx = (size_t)iw + (size_t)im * ldm + (size_t)k * ldk;
        s = s + o_by_mol[idx] + oc[idx];
    }
    otot_by_mol[(size_t)im + (size_t)iw * MRTM_MXMOL] = s;
}

// =============================================================================================
// rt_kernel: one thread per (frequency, profile).  CALCTMR (RTMmono.f90:239-325), RAD_UP_DN
// (:157-221) and RTM (:13-155) in one pass structure; O(iw,layer) is read with iw fastest so a
// warp reads 256 contiguous bytes per layer.  ODT is formed by successive subtraction from the
// layer total exactly as the reference does (:196,:212).
// =============================================================================================
struct RtArgs {
    int32_t nwn, nlay, nprof;
    int32_t irt, iout, do_tmr, do_rtm;
    const double* wn;
    const double* o;  int64_t o_lds, o_prof;
    const double *t, *tz;          // (nlay,nprof), (nlay+1,nprof)
    const double *fb, *fbz;        // RADCN2/T (nlay,nprof), RADCN2/TZ (nlay+1,nprof) from rt_prep_kernel
    double* tmpsfc;                // (nprof) device, in/out
    const double *emiss, *reflc;   // (nwn)
    double *rad, *tb, *tmr, *trtot, *rup, *rdn;   // (nwn,nprof), any may be null
};

__device__ __forceinline__ double bb_fn(double v, double fbeta)
{
    return kRADCN1 * (v * v * v) / (exp(v * fbeta) - 1.);
}

// fbeta = RADCN2/T of every layer and level (RTMmono.f90:183-185,199,215): frequency independent
__global__ void rt_prep_kernel(int n_t, const double* t, double* fb, int n_tz, const double* tz, double* fbz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_t) fb[i] = kRADCN2 / t[i];
    if (i < n_tz) fbz[i] = kRADCN2 / tz[i];
}

// One pass from the top layer down serves the three reference loops: the downwelling sum (RTMmono.f90:207-219)
// and CALCTMR's (:300-317) run in their own order, ODT by successive subtraction from the total exactly as written
// there; the upwelling sum (:192-204) needs, for layer l, the optical depth above it, which the same pass carries
// as a running sum from the top (the reference subtracts from the total going up: same value up to rounding).
// Per (frequency, layer): Planck at the layer temperature and at one new level (the lower boundary becomes the
// next layer's upper boundary) and exp(-tau); the two path transmittances follow by recurrence: 3 exp instead of 8.
// kRtParts threads share a frequency: each walks a contiguous block of layers (part 0 the uppermost; a part's
// starting transmittances come from the optical depth of the parts above it) and the partial sums meet in shared
// memory.  kRtParts times the warps for the same arithmetic -- the layer loop is a latency-bound chain.
#ifndef MRTM_RT_PARTS
#define MRTM_RT_PARTS 4
#endif
constexpr int kRtParts = MRTM_RT_PARTS;
constexpr int kRtFreqs = 128 / kRtParts;     // frequencies per CTA (128 threads)
__global__ void __launch_bounds__(kRtParts * kRtFreqs) rt_kernel(RtArgs a)
{
    __shared__ double s_sum[kRtParts][kRtFreqs], s_rdn[kRtParts][kRtFreqs], s_rup[kRtParts][kRtFreqs];
    const int fi = threadIdx.x % kRtFreqs, part = threadIdx.x / kRtFreqs;
    const int iw_raw = blockIdx.x * kRtFreqs + fi;
    const bool live = iw_raw < a.nwn;
    const int iw = live ? iw_raw : (a.nwn - 1);
    const int prof = blockIdx.y;
    const double vv = a.wn[iw];
    const double* o = a.o + (size_t)prof * a.o_prof + iw;
    const double* __restrict__ fb = a.fb + (size_t)prof * a.nlay;
    const double* __restrict__ fbz = a.fbz + (size_t)prof * (a.nlay + 1);
    const size_t out = (size_t)iw + (size_t)prof * a.nwn;
    // part p owns layers (lo_p, hi_p], cut points at multiples of nlay/kRtParts counted from the top
    const int l_hi = a.nlay - (int)(((long long)a.nlay * part) / kRtParts);
    const int l_lo = a.nlay - (int)(((long long)a.nlay * (part + 1)) / kRtParts) + 1;

    double psum = 0.;
    for (int l = l_lo; l <= l_hi; l++) psum = psum + o[(size_t)(l - 1) * a.o_lds];
    s_sum[part][fi] = psum;
    __syncthreads();
    double od_above = 0., odtot = 0.;
#pragma unroll
    for (int p = kRtParts - 1; p >= 0; p--) odtot = odtot + s_sum[p][fi];       // lowest layers first (RTMmono.f90:177-181)
#pragma unroll
    for (int p = 0; p < kRtParts; p++) od_above += (p < part) ? s_sum[p][fi] : 0.;

    const bool up = a.do_rtm && a.irt != 3;
    const double c1v3 = kRADCN1 * (vv * vv * vv);
    double rup = 0., rdn = 0.;
    // optical depth below the current layer after the subtraction (down loops); a lower part starts below the parts above it
    double odt = odtot - od_above;
    double bb_top = c1v3 / (exp(vv * __ldg(fbz + l_hi)) - 1.);
    // Path transmittances by recurrence instead of one exponential each per layer: above the layer
    // tra = prod(tri of the layers above) (underflow to 0 is the right limit); below it trt(l) = trt(l+1)/tri(l),
    // re-anchored with exp(-odt) while either factor is too small to divide by (opaque columns).  The relative
    // error grows by ~1.5 ulp per layer (<= 1e-13 over 300 layers; bar: 1e-5 K).  3 exp per layer instead of 5.
    double trt = exp(-odt);
    double tra = part == 0 ? 1. : exp(-od_above);
    for (int l = l_hi; l >= l_lo; l--) {
        const double odvi = o[(size_t)(l - 1) * a.o_lds];
        const double bb = c1v3 * rcp3(exp(vv * __ldg(fb + l - 1)) - 1.);
        const double bb_bot = c1v3 * rcp3(exp(vv * __ldg(fbz + l - 1)) - 1.);
        odt = odt - odvi;
        const double tri = exp(-odvi);
        trt = (trt > 1e-250 && tri > 1e-50) ? trt * rcp3(tri) : exp(-odt);
        const double pade = 0.193 * odvi + 0.013 * (odvi * odvi);
        const double rden = rcp3(1. + pade);
        const double emis = 1. - tri;
        rdn = rdn + trt * emis * ((bb + pade * bb_bot) * rden);
        if (up) {
            rup = rup + tra * emis * ((bb + pade * bb_top) * rden);
            tra = tra * tri;
        }
        bb_top = bb_bot;
    }
    s_rdn[part][fi] = rdn;
    s_rup[part][fi] = rup;
    __syncthreads();
    if (part != 0 || !live) return;
#pragma unroll
    for (int p = 1; p < kRtParts; p++) {                // upper layers first, as the reference's top-down loops add them
        rdn = rdn + s_rdn[p][fi];
        rup = rup + s_rup[p][fi];
    }
    const double trtot = exp(-odtot);
    if (a.do_tmr && a.tmr) {
        double radtmr = rdn / (1. - exp(-1 * odtot));
        double x = kRADCN1 * (vv * vv * vv) / radtmr + 1.;
        a.tmr[out] = kRADCN2 * vv / log(x);
    }
    if (a.do_rtm) {
        if (a.rup) a.rup[out] = rup;
        if (a.rdn) a.rdn[out] = rdn;
        if (a.trtot) a.trtot[out] = trtot;
        const double tsky = 2.75;
        // RTMmono.f90:113-123: for downwelling / limb runs the boundary is reset to the cosmic value
        const double tsfc = (a.irt == 3 || a.irt == 2) ? tsky : a.tmpsfc[prof];
        if ((a.irt == 3 || a.irt == 2) && iw == 0) a.tmpsfc[prof] = tsky;
        const double alph = kRADCN2 / tsky, beta = kRADCN2 / tsfc;
        const double surfrad = bb_fn(vv, beta), cosmos = bb_fn(vv, alph);
        const double esfc = a.emiss[iw], rsfc = a.reflc[iw];
        double rad = 0.;
        if (a.irt == 1) rad = rup + trtot * (esfc * surfrad + rsfc * (rdn + trtot * cosmos));
        if (a.irt == 2) rad = rup + trtot * (rdn + trtot * cosmos);
        if (a.irt == 3) rad = rdn + (trtot * cosmos);
        if (a.rad) a.rad[out] = rad;
        if (a.iout == 1 && a.tb) {
            double x = kRADCN1 * (vv * vv * vv) / rad + 1.;
            a.tb[out] = kRADCN2 * vv / log(x);
        }
    }
}

// =============================================================================================
// FP64 FMA throughput probe (roofline denominator for the line-shape kernel)
// =============================================================================================
__global__ void fp64_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace mrtm
