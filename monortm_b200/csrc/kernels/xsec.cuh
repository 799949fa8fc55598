// xsec.cuh -- cross-section optical depths: MONORTM_XSEC_SUB (src/monortm_sub.F90:1540-1749) and convolve (:1751-1834).
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// The reference walks molecule -> spectral region -> layer, interpolates the tabulated cross sections of the region to the
// layer temperature with the radiation term divided out (loop 3300), resamples that table on a grid of about four points
// per Lorentz half width (loop 500, up to 1e7 points per (region, layer)), and sums, for every frequency, Lorentzians of
// width HWB outward from the frequency until a term drops below RATIO*1e-6 of the running sum (the GOTO loop 1000).
// Here:
//   xs_layer_kernel  one thread per (region, layer): temperature bracket, table pressure, the half widths, step, NPTS
//   xs_table_kernel  one thread per (table point, layer, region): the temperature-interpolated table (loop 3300) in HBM
//   xs_need_kernel   which regions have a frequency inside [V1FX-1, V2FX+1] (:1647-1653)
//   xs_conv_kernel   one thread per (frequency, layer): the outward sum with the resampled table evaluated on the fly
//                    (the 1e7-point intermediate of loop 500 is never stored); regions and molecules in the reference's order
// The outward sum stops on a floating-point comparison of the running sum, so its arithmetic is written with
// non-contracted IEEE operations in the reference's association: the number of terms and the sum itself are those of
// the oracle, up to the last-bit differences of exp() inside RADFN.
// =============================================================================================
constexpr long long kXsIntMax = 10000000;            // xspd_int(0:10000000), monortm_sub.F90:1755

struct XsRegionDev {
    int32_t ixmol, ntemp;
    int64_t npts;
    double v1fx, v2fx, v1x, v2x, xdoplr;
    double tx[6], pdx[6];
    int64_t dat_off[6];        // offsets of the temperature tables in the staged data array
    int64_t tab_off;           // offset of this region's per-layer interpolated tables: tab[tab_off + il*(npts+3) + i], i = 0..npts+2
};

struct XsLayerDev {            // per (region, layer)
    double coef1, coef2, pd, hwb, hwd, hwb2, step, ratio, delvx;
    long long npts;            // points of the resampled grid (loop 500)
    int32_t ind1, ind2;
    int32_t lorentz;           // 1: Lorentz convolution, 0: interpolation shortcut (:1791)
    int32_t pad;
};

struct XsArgs {
    int32_t nreg, nlay, nwn, ld_xamnt;
    const XsRegionDev* reg;
    const double* dat;         // staged tables
    XsLayerDev* lay;           // [nreg][nlay]
    double* tab;               // interpolated tables
    int* need;                 // [nreg]
    const double *wn, *p, *t, *xamnt;
    double* odxsec;            // (nwn, nlay)
    int* errflag;              // bit 4: resampled grid exceeds the reference's array; bit 5: the outward sum does not terminate
};

__global__ void xs_layer_kernel(XsArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.nreg * a.nlay) return;
    const int r = idx / a.nlay, il = idx % a.nlay;
    const XsRegionDev& g = a.reg[r];
    const double pave = a.p[il], tave = a.t[il];
    XsLayerDev o;
    double coef1 = 1., coef2 = 0.;
    int ind1, ind2 = 1, it = 1;
    if (g.ntemp == 1 || tave <= g.tx[it - 1]) {                // :1680-1699
        ind1 = 1;
    } else {
        for (;;) {
            it = it + 1;
            if (it > g.ntemp) { ind1 = g.ntemp; ind2 = g.ntemp; break; }
            else if (tave <= g.tx[it - 1]) {
                ind1 = it - 1;
                ind2 = it;
                coef1 = xdiv(xsub(tave, g.tx[it - 1]), xsub(g.tx[it - 2], g.tx[it - 1]));
                coef2 = xsub(1., coef1);
                break;
            }
        }
    }
    o.coef1 = coef1; o.coef2 = coef2; o.ind1 = ind1; o.ind2 = ind2;
    o.pd = xadd(xmul(coef1, g.pdx[ind1 - 1]), xmul(coef2, g.pdx[ind2 - 1]));
    o.delvx = xdiv(xsub(g.v2x, g.v1x), (double)(g.npts - 1));
    const double hwdop = xmul(g.xdoplr, sqrt(xdiv(tave, 296.)));
    // convolve :1760-1776
    const double p0 = 1013.;
    double hwpave = xmul(xmul(0.1, xdiv(pave, p0)), xdiv(273.15, tave));
    double hwd = xmul(xmul(0.1, xdiv(o.pd, p0)), xdiv(273.15, tave));
    hwd = hwd > hwdop ? hwd : hwdop;
    if (hwd > hwpave) hwpave = xmul(1.001, hwd);
    const double hwb = xsub(hwpave, hwd);
    double step = xmul(0.25, hwb);
    if (step > o.delvx) step = o.delvx;
    const double q = xdiv(xsub(g.v2x, g.v1x), step);
    long long npts = 0;
    if (!(q < (double)kXsIntMax + 1.)) atomicOr(a.errflag, 16);
    else npts = (long long)q;
    step = xdiv(xsub(g.v2x, g.v1x), (double)npts);
    o.npts = npts;
    o.step = step;
    o.ratio = xdiv(step, hwb);
    o.hwb = hwb;
    o.hwd = hwd;
    o.hwb2 = xmul(hwb, hwb);
    o.lorentz = (xdiv(hwb, hwd) > 0.1) ? 1 : 0;
    o.pad = 0;
    a.lay[idx] = o;
}

// loop 3300: xspd(i) = coef1*xsdat(i,ind1)/radfn(vv,xkt1) + coef2*xsdat(i,ind2)/radfn(vv,xkt2); stored with one guard
// element below (xspd(0), read by the shortcut at w = v1x) and two above (xspd(nptsx+1), xspd(nptsx+2)): all zero
__global__ void xs_table_kernel(XsArgs a)
{
    const int r = blockIdx.z, il = blockIdx.y;
    const XsRegionDev& g = a.reg[r];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. npts+2
    if (i > g.npts + 2) return;
    const XsLayerDev& ly = a.lay[r * a.nlay + il];
    double v = 0.;
    if (i >= 1 && i <= g.npts) {
        const double kRADCN2x = 1.4387752;
        const double xkt1 = xdiv(g.tx[ly.ind1 - 1], kRADCN2x), xkt2 = xdiv(g.tx[ly.ind2 - 1], kRADCN2x);
        const double vv = xadd(g.v1x, xmul((double)(i - 1), ly.delvx));
        const double d1 = a.dat[g.dat_off[ly.ind1 - 1] + (i - 1)], d2 = a.dat[g.dat_off[ly.ind2 - 1] + (i - 1)];
        v = xadd(xdiv(xmul(ly.coef1, d1), radfn(vv, xkt1)), xdiv(xmul(ly.coef2, d2), radfn(vv, xkt2)));
    }
    a.tab[g.tab_off + (long long)il * (g.npts + 3) + i] = v;
}

__global__ void xs_need_kernel(XsArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nwn) return;
    const double w = a.wn[i];
    for (int r = 0; r < a.nreg; r++)
        if (w >= a.reg[r].v1fx - 1.0 && w <= a.reg[r].v2fx + 1.0 && !a.need[r]) atomicOr(a.need + r, 1);
}

// xspd_int(i) of loop 500, evaluated where it is needed
__device__ __forceinline__ double xs_resampled(const double* __restrict__ tab, const double v1x, const double step,
                                               const double delvx, const long long i)
{
    const double vv = xadd(v1x, xmul((double)i, step));
    const double delvv = xsub(vv, v1x);
    const long long ind = (long long)xdiv(delvv, delvx);
    const double coef = xdiv(xsub(delvv, xmul((double)ind, delvx)), delvx);
    return xadd(xmul(xsub(1., coef), tab[ind + 1]), xmul(coef, tab[ind + 2]));     // tab[k] = xspd(k)
}

__global__ void __launch_bounds__(128) xs_conv_kernel(XsArgs a)
{
    const int iw = blockIdx.x * blockDim.x + threadIdx.x;
    const int il = blockIdx.y;
    if (iw >= a.nwn) return;
    const double w = a.wn[iw];
    double xstot = 0.;
    int r = 0, err = 0;
    while (r < a.nreg) {
        const int ixmol = a.reg[r].ixmol;
        double xsmol = 0.;
        for (; r < a.nreg && a.reg[r].ixmol == ixmol; r++) {
            if (!a.need[r]) continue;
            const XsRegionDev& g = a.reg[r];
            const XsLayerDev& ly = a.lay[r * a.nlay + il];
            const double* __restrict__ tab = a.tab + g.tab_off + (long long)il * (g.npts + 3);
            const double v1x = g.v1x, v2x = g.v2x;
            double xspave;
            if (w < v1x || w > v2x) {
                xspave = 0.;
            } else if (ly.lorentz) {
                const double hwb = ly.hwb, hwb2 = ly.hwb2, step = ly.step, delvx = ly.delvx;
                const double wn_v1x = xsub(w, v1x);
                const long long ind = (long long)xdiv(wn_v1x, step);
                double dvlo = xsub(w, xadd(v1x, xmul((double)ind, step)));
                double dvhi = xsub(w, xadd(v1x, xmul((double)(ind + 1), step)));
                const double xi1 = (ind + 1 <= ly.npts) ? xs_resampled(tab, v1x, step, delvx, ind + 1) : 0.;
                double answer = xadd(xmul(xdiv(hwb, xadd(hwb2, xmul(dvlo, dvlo))), xs_resampled(tab, v1x, step, delvx, ind)),
                                     xmul(xdiv(hwb, xadd(hwb2, xmul(dvhi, dvhi))), xi1));
                const double thr = xmul(ly.ratio, 1e-6);
                long long j = 1;
                for (;;) {
                    double contlo = 0., conthi = 0.;
                    const double vlo = xadd(v1x, xmul((double)(ind - j), step));
                    if (vlo > v1x) {
                        dvlo = xsub(w, vlo);
                        contlo = xmul(xdiv(hwb, xadd(hwb2, xmul(dvlo, dvlo))), xs_resampled(tab, v1x, step, delvx, ind - j));
                    }
                    const double vhi = xadd(v1x, xmul((double)(ind + j + 1), step));
                    if (vhi < v2x) {
                        dvhi = xsub(w, vhi);
                        conthi = xmul(xdiv(hwb, xadd(hwb2, xmul(dvhi, dvhi))), xs_resampled(tab, v1x, step, delvx, ind + j + 1));
                    }
                    const double xincr = xadd(contlo, conthi);
                    if (xdiv(xincr, answer) < thr) break;
                    answer = xadd(answer, xincr);
                    j = j + 1;
                    if (j > ly.npts + 2) { err = 32; break; }          // the reference would never leave the loop (0/0 or NaN)
                }
                xspave = xdiv(xmul(answer, step), 3.14159);
            } else {
                const double wn_v1x = xsub(w, v1x);
                const long long ind = (long long)xdiv(wn_v1x, ly.delvx);
                const double coef = xdiv(xsub(wn_v1x, xmul((double)ind, ly.delvx)), ly.delvx);
                xspave = xadd(xmul(xsub(1., coef), tab[ind]), xmul(coef, tab[ind + 1]));
            }
            xsmol = xadd(xsmol, xspave);
        }
        xstot = xadd(xstot, xmul(a.xamnt[(size_t)ixmol + (size_t)il * a.ld_xamnt], xsmol));
    }
    if (err) atomicOr(a.errflag, err);
    a.odxsec[(size_t)iw + (size_t)il * a.nwn] = xmul(xstot, radfn(w, xdiv(a.t[il], 1.4387752)));
}
