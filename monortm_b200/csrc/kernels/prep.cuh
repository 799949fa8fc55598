// prep.cuh -- layer_prep_kernel, continuum_kernel, derive_kernel: per-(profile,layer) and per-(line,layer) preparation.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
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
    o.inv_t = 1. / tt;
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
        o.c_amagat = amagat;
        o.c_rhoave = rhoave;
        o.c_wk3 = (a.nmol >= 3) ? wk[2] : 0.;
        o.c_wk7 = wk7;
        o.c_cw = cw;
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
            if (row == -2 && !(tt < 70. || tt > 3000.)) {
                v = 1.;                 // molecules 34 and 39: scor == 1 (tips_2003.f90:233-238, 260-292)
            } else if (row < 0 || tt < 70. || tt > 3000.) {
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
// continuum_kernel: one CTA per (profile,layer).  Every MT_CKD_3.5 component of CONTNM (contnm.f90:325-1131): H2O self and
// foreign, CO2, O3 (Chappuis / Hartley-Huggins / UV), O2 (fundamental, 1.27 um, 9100-11000, A band, visible, Herzberg, far
// UV), N2 (roto-translational, fundamental, first overtone) and Rayleigh.  A component is evaluated on its own coefficient
// grid and 4-point interpolated (XINT, lblrtm_sub.f90:1-34) onto the 1 cm-1 grid of its species' plane:
// absrb[L][CP_COUNT][nptabs_pad], planes H2O, CO2, O3, O2, N2 (species selectors im = 1, 2, 3, 7, 22 of modm.f90:165) and
// Rayleigh.  Which components fire, their table offsets and XINT bounds are layer independent (host, make_grid).
// =============================================================================================
struct ContArgs {
    int64_t nlayers;
    ContGrid g[CB_COUNT];
    double v1abs;
    int32_t nptabs, nptabs_pad;
    int32_t rayl_active, pad;
    double cntnm[7];
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

__device__ __forceinline__ int cont_plane(int b)
{
    switch (b) {
    case CB_H2O_SELF: case CB_H2O_FRGN: return CP_H2O;
    case CB_CO2: return CP_CO2;
    case CB_N2_ROT: case CB_N2_FUND: case CB_N2_OVER: return CP_N2;
    case CB_O3_CHAP: case CB_O3_HH: case CB_O3_UV: return CP_O3;
    default: return CP_O2;
    }
}

// RADFN, lblrtm_sub.f90:36-97 (the Rayleigh term divides it out again, contnm.f90:1124)
__device__ __forceinline__ double cont_radfn(double vi, double xkt)
{
    if (xkt > 0.0) {
        const double x = vi / xkt;
        if (x <= 0.01) return 0.5 * x * vi;
        if (x <= 10.0) { const double e = exp(-x); return vi * (1. - e) / (1. + e); }
        return vi;
    }
    return vi;
}

// value of component b at point j (0-based) of its coefficient grid: the accessor routine (table look-up, radiation term
// removed) times the layer factor of the calling block in CONTNM
__device__ double cont_value(int b, const ContGrid& g, int j, const LayerDev& ly, const ContArgs& a)
{
    const int i = g.i1 + j;                                   // 1-based table index
    const double vj = g.v1c + g.dvc * (double)j;
    const double xlosmt = 2.68675E+19;
    switch (b) {
    case CB_H2O_SELF: {                                       // contnm.f90:325-363, SL296/SL260
        double s0 = 0., s1 = 0., sh2o = 0.;
        if (i >= 1 && i <= 2003) { s0 = a.tb.sh2o_296[i - 1]; s1 = a.tb.sh2o_260[i - 1]; }
        if (s0 > 0.) sh2o = s0 * pow(s1 / s0, ly.c_tfac_h2o);
        return ly.c_wk1 * (sh2o * ly.c_rself);
    }
    case CB_H2O_FRGN: {                                       // :380-457, FRN296
        const double f0 = 0.06, v0f1 = 255.67, hwsq1 = 240. * 240., beta1 = 57.83, c_1 = -0.42, c_2 = 0.3, beta2 = 630.;
        double f = (i >= 1 && i <= 2003) ? a.tb.fh2o[i - 1] : 0.;
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
        return (ly.c_wk1 * f) * ly.c_rfrgn;
    }
    case CB_CO2: {                                            // :484-528, FRNCO2
        double f = 0.;
        if (i >= 1 && i <= 5003) {
            double tcor = 1.;
            if (i >= 1196 && i <= 1220) tcor = pow(ly.c_trat, a.tb.co2_tdep[i - 1196]);
            f = tcor * a.tb.fco2[i - 1];
        }
        if (vj >= 2000. && vj <= 2998.) {                     // :510-513
            const int jfac = (int)((vj - 1998.) / 2. + 0.00001);
            f = a.tb.xfacco2[jfac - 1] * f;
        }
        return f * ly.c_wco2;
    }
    case CB_N2_ROT: {                                         // :906-943, xn2_r
        double c0 = 0., c1 = 0.;
        if (i >= 1 && i <= 73) {
            c0 = a.tb.n2_296[i - 1] * pow(a.tb.n2_220[i - 1] / a.tb.n2_296[i - 1], ly.c_tfac_n2);
            double sf_t = a.tb.n2_296_sf[i - 1] * pow(a.tb.n2_220_sf[i - 1] / a.tb.n2_296_sf[i - 1], ly.c_tfac_n2);
            c1 = (sf_t - 1.) * 0.79 / 0.21;
        }
        return ly.c_taufac * c0 * (ly.c_xn2 + c1 * ly.c_xo2 + 1. * ly.c_xh2o);
    }
    case CB_N2_FUND: {                                        // :963-1013, n2_ver_1 :4331-4417
        if (i < 1 || i > 228) return 0.;
        const double t_272 = 272., t_228 = 228.;
        const double xtfac = ((1. / ly.t) - (1. / t_272)) / ((1. / t_228) - (1. / t_272));
        const double xt_lin = (ly.t - t_272) / (t_228 - t_272);
        const double a_o2 = 1.294 - 0.4545 * ly.t / 296.;
        const double x2 = a.tb.n2f_272[i - 1], x8 = a.tb.n2f_228[i - 1];
        double c = ((x2 > 0.) && (x8 > 0.)) ? x2 * pow(x8 / x2, xtfac) : x2 + (x8 - x2) * xt_lin;
        c = c / vj;
        const double c1 = a_o2 * c, c2 = (9. / 7.) * a.tb.n2f_ah2o[i - 1] * c;
        return ly.c_taufac * (ly.c_xn2 * c + ly.c_xo2 * c1 + ly.c_xh2o * c2);
    }
    case CB_N2_OVER: {                                        // :1025-1068, n2_overtone1
        if (i < 1 || i > 191) return 0.;
        return (ly.c_taufac * (ly.c_xn2 + 1. * ly.c_xo2 + 1. * ly.c_xh2o)) * (a.tb.n2f1[i - 1] / vj);
    }
    case CB_O3_CHAP: {                                        // :536-553, XO3CHP
        if (i < 1 || i > 3150) return 0.;
        const double wo3 = ly.c_wk3 * 1.0E-20 * a.cntnm[3], dt = ly.t - 273.15;
        const double c0 = a.tb.o3ch_x[i - 1] / vj, c1 = a.tb.o3ch_y[i - 1] / vj, c2 = a.tb.o3ch_z[i - 1] / vj;
        return (c0 + (c1 + c2 * dt) * dt) * wo3;
    }
    case CB_O3_HH: {                                          // :555-599, O3HHT0/1/2
        if (i < 1 || i > 2687) return 0.;
        const double wo3 = ly.c_wk3 * 1.E-20 * a.cntnm[3], tc = ly.t - 273.15;
        double c = (a.tb.o3hh0[i - 1] / vj) * wo3;
        return c * (1. + a.tb.o3hh1[i - 1] * tc + a.tb.o3hh2[i - 1] * tc * tc);
    }
    case CB_O3_UV: {                                          // :603-642, O3HHUV
        if (i < 1 || i > 133) return 0.;
        return (a.tb.o3huv[i - 1] / vj) * (ly.c_wk3 * a.cntnm[3]);
    }
    case CB_O2_FUND: {                                        // :657-693, o2_ver_1
        if (i < 1 || i > 103) return 0.;
        const double tau_fac = a.cntnm[4] * ly.c_wk7 * 1.e-20 * ly.c_amagat;
        const double xktfac = (1. / 296.) - (1. / ly.t);
        return tau_fac * ((1.e+20 / 2.68675e+19) * a.tb.o2f[i - 1] * exp(a.tb.o2f_t[i - 1] * xktfac) / vj);
    }
    case CB_O2_INF1: {                                        // :709-734, O2INF1
        if (i < 1 || i > 483) return 0.;
        const double tau_fac = a.cntnm[4] * (ly.c_wk7 / xlosmt) * ly.c_amagat *
                               ((1. / 0.446) * ly.c_xo2 + (0.3 / 0.446) * ly.c_xn2 + 1. * ly.c_xh2o);
        return tau_fac * (a.tb.o2inf1[i - 1] / vj);
    }
    case CB_O2_INF2: {                                        // :745-766, O2INF2 (analytic, its own grid set-up)
        if (!((vj > 9100.) && (vj < 11000.))) return 0.;
        const double v1_osc = 9375., hw1 = 58.96, v2_osc = 9439., hw2 = 45.04, s1 = 1.166E-04, s2 = 3.086E-05;
        const double dv1 = vj - v1_osc, dv2 = vj - v2_osc;
        const double damp1 = (dv1 < 0.0) ? exp(dv1 / 176.1) : 1.0, damp2 = (dv2 < 0.0) ? exp(dv2 / 176.1) : 1.0;
        const double q1 = dv1 / hw1, q2 = dv2 / hw2;
        const double o2inf = 0.31831 * (((s1 * damp1 / hw1) / (1. + q1 * q1)) + ((s2 * damp2 / hw2) / (1. + q2 * q2))) * 1.054;
        const double wo2 = a.cntnm[4] * (ly.c_wk7 * 1.e-20) * ly.c_rhoave;
        return (o2inf / vj) * ((ly.c_wk7 / ly.c_cw) * (1. / 0.209) * wo2);
    }
    case CB_O2_INF3: {                                        // :773-792, O2INF3
        if (i < 1 || i > 261) return 0.;
        return (a.cntnm[4] * (ly.c_wk7 / xlosmt) * ly.c_amagat) * (a.tb.o2inf3[i - 1] / vj);
    }
    case CB_O2_VIS: {                                         // :807-830, O2_vis
        if (i < 1 || i > 1474) return 0.;
        const double wo2 = ly.c_wk7 * 1.e-20 * ((ly.p / 1013.) * (273. / ly.t)) * a.cntnm[4];
        const double q = 55. * 273. / 296.;
        const double factor = 1. / ((xlosmt * 1.e-20 * (q * q)) * 89.5);
        return (factor * a.tb.o2vis[i - 1] / vj) * ((ly.c_wk7 / ly.c_cw) * wo2);
    }
    case CB_O2_HERZ: {                                        // :834-850, O2HERZ, HERTDA, HERPRS
        if (i < 1) return 0.;
        double herz = 0.0;
        if (!(vj <= 36000.00)) {
            double corr = 0.;
            if (vj <= 40000.) corr = ((40000. - vj) / 4000.) * 7.917E-07;
            const double yratio = vj / 48811.0, lg = log(yratio);
            herz = 6.884E-04 * (yratio) * exp(-69.738 * (lg * lg)) - corr;
        }
        herz = herz * (1. + .83 * (ly.p / 1013.) * (273.16 / ly.t));
        return (herz / vj) * (ly.c_wk7 * 1.e-20 * a.cntnm[4]);
    }
    case CB_O2_FUV: {                                         // :857-875, O2FUV
        if (i < 1 || i > 1512) return 0.;
        return (a.tb.o2fuv[i - 1] / vj) * (ly.c_wk7 * 1.e-20 * a.cntnm[4]);
    }
    default: return 0.;
    }
}

__global__ void __launch_bounds__(128) continuum_kernel(ContArgs a)
{
    extern __shared__ double s_c[];
    const int64_t L = blockIdx.x;
    const LayerDev& ly = a.lay[L];
    const int tid = threadIdx.x;
    double* out = a.absrb + (size_t)L * CP_COUNT * a.nptabs_pad;
    for (int i = tid; i < CP_COUNT * a.nptabs_pad; i += blockDim.x) out[i] = 0.;
    if (a.rayl_active) {                                      // contnm.f90:1107-1131 with JRAD = 0, IAERSL = 0
        const double conv_cm2mol = a.cntnm[6] * 1.E-20 / (2.68675e-1 * 1.e5);
        double* pr = out + (size_t)CP_RAYL * a.nptabs_pad;
        for (int i = 1 + tid; i <= a.nptabs; i += blockDim.x) {
            const double vr = a.v1abs + (double)(i - 1) * 1.0;
            const double x = vr / 1.e4;
            double ray = ((x * x * x) / (9.38076E2 - 10.8426 * (x * x))) * (ly.c_cw * conv_cm2mol);
            pr[i - 1] = ray * x / cont_radfn(vr, ly.xkt);
        }
    }
    __syncthreads();
    for (int b = 0; b < CB_COUNT; b++) {
        const ContGrid g = a.g[b];
        if (!g.active) continue;
        for (int j = tid; j < g.nptc; j += blockDim.x) s_c[j] = cont_value(b, g, j, ly, a);
        __syncthreads();
        double* pl = out + (size_t)cont_plane(b) * a.nptabs_pad;
        for (int i = g.ilo + tid; i <= g.ihi; i += blockDim.x)
            pl[i - 1] = pl[i - 1] + xint_point(s_c, g.v1c, g.dvc, a.v1abs + 1.0 * (double)(i - 1));
        __syncthreads();
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
    double* lcplanes;         // [L][LCP_NPLANES][nlc_pad]
    int32_t nlc_pad, pad3;
    unsigned long long* layer_voigt;   // [L] bits of the smallest |Xnu| among the lines of the layer that can take the Voigt branch
                                       // (zeta <= 0.99; 100*HWHM_D grows with |Xnu|), all ones = none
    int32_t nseg, pad2;
};

#ifndef MRTM_DERIVE_MINB
#define MRTM_DERIVE_MINB 4
#endif
// the static parameters of one line: read once per thread, used for kDeriveLayers layers
struct LineStatic {
    int32_t mol, xf, cls, lci, bi, sidx;
    double xnu0, deltnu, es, s0adj, af, as, x, dopf;
};
__device__ __forceinline__ LineStatic load_line_static(const LinesDev& ln, int q)
{
    LineStatic t;
    t.mol = ln.mol[q]; t.xf = ln.xf[q]; t.cls = ln.cls[q]; t.lci = ln.lcidx[q]; t.bi = ln.brdidx[q]; t.sidx = ln.sidx[q];
    t.xnu0 = ln.xnu0[q]; t.deltnu = ln.deltnu[q]; t.es = ln.e[q]; t.s0adj = ln.s0adj[q];
    t.af = ln.alpf[q]; t.as = ln.alps[q]; t.x = ln.x[q]; t.dopf = ln.dopf[q];
    return t;
}
// one (line, layer): returns the bits of |Xnu| when the line can take the Voigt branch in this layer, else all ones
__device__ __forceinline__ unsigned long long derive_one(const DeriveArgs& a, int q, int64_t L, const LineStatic& ls)
{
    const unsigned long long kNone = ~0ull;
    if (q >= a.ln.n_pad) return kNone;
    double* pl = a.planes + (size_t)L * D_NPLANES * a.ln.n_pad;
    if (q >= a.ln.n) {   // padding: far away, zero strength
        pl[(size_t)D_XNU * a.ln.n_pad + q] = 1.0e30;
        pl[(size_t)D_H2 * a.ln.n_pad + q] = 1.0;
        pl[(size_t)D_CN * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_P3 * a.ln.n_pad + q] = 0.;
        pl[(size_t)D_VT * a.ln.n_pad + q] = -1.0;
        return kNone;
    }
    const LayerDev& ly = a.lay[L];
    const int mol = ls.mol, xf = ls.xf, cls = ls.cls;
    const double rhorat = ly.rhorat, rho_self = ly.rho_self[mol - 1];
    const double radct = ly.radct;

    double aip = 0., bip = 0.;
    const int lci = ls.lci;
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
    const double xnu0 = ls.xnu0, deltnu = ls.deltnu;
    double xnu = xadd(xnu0, xmul(deltnu, rhorat));
    const int bi = ls.bi;
    const bool use_brd = (mol <= MRTM_MXBRDMOL) && (a.ibrd != 0);
    const double* brd = (bi >= 0) ? a.ln.brd + (size_t)bi * 28 : nullptr;
    if (use_brd) {
        double s = 0.;
        if (brd)
            for (int k = 0; k < 7; k++) s = xadd(s, xmul(xmul(ly.rho_molec[k], brd[k]), xsub(brd[21 + k], deltnu)));
        xnu = xadd(xnu, s);
    }

    // INTENS.  The divisions of the reference (:860-865, :453, :419) are regrouped into multiplications by per-layer and
    // per-line reciprocals and a Newton reciprocal (~1 ulp each; the bar on optical depths is 1e-9).
    const double xipsf = a.scorc[(size_t)L * a.ln.nsi + ls.sidx];
    const double es = ls.es;
    // exp(-c2 E/T)/exp(-c2 E/T0) as one exponential (modm.f90 INTENS)
    double s = ls.s0adj * exp(radct * es * ly.dinvt) * xipsf;
    const double rx = radct * xnu;
    const double stim_n = 1 + exp(-(rx * ly.inv_t)), stim_d = xnu * (1 - exp(-(rx * (1. / kT0))));
    double stild = (stim_d > 1e-280 && stim_d < 1e280) ? s * (stim_n * rcp3(stim_d)) : s * (stim_n / stim_d);

    // HALFWHM_C
    const double af = ls.af, as = ls.as;
    const double rtx = exp(ls.x * ly.lnrt);         // (T/T0)^x with the layer's log(T/T0)
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
    // HALFWHM_D = (Xnu/c)*sqrt(2 ln2 kT/(M/N_A)): the constants and the mass are folded per line at staging (dopf)
    const double hwhm_d = xnu * (ls.dopf * ly.sqrt_t);
    if (xf == -3) hwhm_c = hwhm_c * (1 - (aip * ly.rp) - (bip * ly.rp2));
    // zeta = HWHM_C/(HWHM_C+HWHM_D) decides Voigt or Lorentz at 0.99 (:419): the exact quotient only where the fast one is
    // too close to the threshold to decide
    const double zsum = hwhm_c + hwhm_d;
    double zeta = (zsum > 1e-280 && zsum < 1e280) ? hwhm_c * rcp3(zsum) : hwhm_c / zsum;
    if (fabs(zeta - 0.99) < 1e-12) zeta = hwhm_c / zsum;

    const double h2 = hwhm_c * hwhm_c;
    const double cn = (stild * hwhm_c) * (1. / kPI);
    double p3 = 0.;
    if (cls == CLS_PED) p3 = cn * rcp3(kDELTNUC * kDELTNUC + h2);
    if (cls == CLS_O2_LC1) p3 = cn * (1. + bip * ly.rp2);
    const size_t np = a.ln.n_pad;
    pl[(size_t)D_XNU * np + q] = xnu;
    pl[(size_t)D_H2 * np + q] = h2;
    pl[(size_t)D_CN * np + q] = cn;
    pl[(size_t)D_P3 * np + q] = p3;
    const double vt = (zeta > 0.99) ? -1.0 : 100. * hwhm_d;
    pl[(size_t)D_VT * np + q] = vt;
    if (lci >= 0) {
        double* lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
        lcp[(size_t)LCP_AIP * a.nlc_pad + lci] = aip;
        lcp[(size_t)LCP_BIP * a.nlc_pad + lci] = bip;
    }
    return (vt >= 0.) ? (unsigned long long)__double_as_longlong(fabs(xnu)) : kNone;
}

#ifndef MRTM_DERIVE_LAYERS
#define MRTM_DERIVE_LAYERS 8        // measured (profiles/r02_sweeps.md): 1 layer/thread 171 us, 4 at 64 registers 138, 8: 133
#endif
constexpr int kDeriveLayers = MRTM_DERIVE_LAYERS;      // layers per thread: the line's static parameters are read once for all of them
__global__ void __launch_bounds__(256, MRTM_DERIVE_MINB) derive_kernel(DeriveArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    LineStatic ls;
    if (q < a.ln.n) ls = load_line_static(a.ln, q);
    __shared__ unsigned long long s_min[8];
    for (int i = 0; i < kDeriveLayers; i++) {
        const int64_t L = (int64_t)blockIdx.y * kDeriveLayers + i;
        if (L >= a.nlayers) break;
        unsigned long long xb = derive_one(a, q, L, ls);
        // smallest |Xnu| of the layer's Voigt-capable lines: warp minimum, block minimum, one global atomic per block at most
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, xb, off);
            xb = o < xb ? o : xb;
        }
        if (i > 0) __syncthreads();                        // s_min of the previous layer has been read
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = xb;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long m = s_min[0];
            for (int w = 1; w < 8; w++) m = s_min[w] < m ? s_min[w] : m;
            if (m < *(volatile unsigned long long*)(a.layer_voigt + L)) atomicMin(a.layer_voigt + L, m);
        }
    }
}
