// mrtm_device.cuh -- device-side math shared by the kernels: exact (non-contracted) arithmetic
// helpers, complex helpers, the Humlicek / speed-dependent Voigt routines and the general
// line-shape case tree.  Reference: src/modm.f90 (line numbers cited per function).
#pragma once
#include <cuda_runtime.h>

#include "mrtm_internal.h"

namespace mrtm {

// ---- exact IEEE arithmetic: never contracted into FMA by the compiler -----------------------
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// reciprocal for strictly positive, normal x: MUFU seed (2^-23) + two Newton steps (~1 ulp)
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// reciprocal of a strictly positive normal double: MUFU seed (rel. err <= 2^-20) and one
// third-order step r0*(1+e+e^2), e = 1-x*r0: error e^3 <= 2^-60, i.e. ~1 ulp after rounding.
__device__ __forceinline__ double rcp3(double x)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    double t = fma(e, e, e);
    return fma(r0, t, r0);
}

// ---- complex helpers ------------------------------------------------------------------------
struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmk(double re, double im) { return cplx{re, im}; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx operator+(double a, cplx b) { return cplx{a + b.re, b.im}; }
__device__ __forceinline__ cplx operator-(double a, cplx b) { return cplx{a - b.re, -b.im}; }
__device__ __forceinline__ cplx operator*(cplx a, cplx b)
{
    return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cplx operator*(cplx a, double b) { return cplx{a.re * b, a.im * b}; }
__device__ __forceinline__ cplx operator/(cplx a, cplx b)
{
    // Smith's algorithm (range reduction, as Fortran complex division rules do)
    double r, d;
    if (fabs(b.re) >= fabs(b.im)) {
        r = b.im / b.re;
        d = b.re + b.im * r;
        return cplx{(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    }
    r = b.re / b.im;
    d = b.im + b.re * r;
    return cplx{(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}
__device__ __forceinline__ cplx cexpd(cplx u)
{
    double s, c;
    sincos(u.im, &s, &c);
    double e = exp(u.re);
    return cplx{e * c, e * s};
}

// Humlicek regions (modm.f90:1105-1128); t = CMPLX(y,-x)
__device__ __forceinline__ cplx hum1(cplx t) { return t * .5641896 / (.5 + t * t); }
__device__ __forceinline__ cplx hum2(cplx t)
{
    cplx u = t * t;
    return t * (1.410474 + u * .5641896) / (.75 + u * (3. + u));
}
__device__ __forceinline__ cplx hum3(cplx t)
{
    return (16.4955 + t * (20.20933 + t * (11.96482 + t * (3.778987 + t * .5642236)))) /
           (16.4955 + t * (38.82363 + t * (39.27121 + t * (21.69274 + t * (6.699398 + t)))));
}
__device__ __forceinline__ cplx hum4(cplx t)
{
    cplx u = t * t;
    return cexpd(u) - t * (36183.31 - u * (3321.9905 - u * (1540.787 - u * (219.0313 - u * (35.76683 - u * (1.320522 - u * .56419)))))) /
                          (32066.6 - u * (24322.84 - u * (9022.228 - u * (2186.181 - u * (364.2191 - u * (61.57037 - u * (1.841439 - u)))))));
}

// W4, modm.f90:1100-1130
__device__ __noinline__ cplx w4(double x, double y)
{
    cplx t = cmk(y, -x);
    double s = fabs(x) + y;
    if (!(s < 15.)) return hum1(t);
    if (!(s < 5.5)) return hum2(t);
    if (!(y < 0.195 * fabs(x) - 0.176)) return hum3(t);
    return hum4(t);
}

// SD_Humlicek, modm.f90:1150-1251
__device__ __noinline__ cplx sd_humlicek(double x1, double y1, double x2, double y2)
{
    cplx t1 = cmk(y1, -x1), t2 = cmk(y2, -x2);
    double s1 = fabs(x1) + y1, s2 = fabs(x2) + y2;
    int r1, r2;
    if (s1 >= 15.0) r1 = 1;
    else if (s1 >= 6.0 && s1 < 15.0) r1 = 2;
    else { r1 = 3; if (y1 < 0.195 * fabs(x1) - 0.176) r1 = 4; }
    if (s2 >= 15.0) r2 = 1;
    else if (s2 >= 6.0 && s2 < 15.0) r2 = 2;
    else { r2 = 3; if (y2 < 0.195 * fabs(x2) - 0.176) r2 = 4; }
    int region = r1 > r2 ? r1 : r2;
    if (!(region > 1)) return hum1(t1) - hum1(t2);
    if (!(region > 2)) return hum2(t1) - hum2(t2);
    if (!(region > 3)) return hum3(t1) - hum3(t2);
    cplx w1 = (r1 == 4) ? hum4(t1) : hum3(t1);
    cplx w2 = (r2 == 4) ? hum4(t2) : hum3(t2);
    return w1 - w2;
}

// SDVOIGT, modm.f90:965-1087.  *err is set when REAL(v) < 0 (:1062 STOP).
__device__ __noinline__ double sdvoigt(double deltnu, double alphal, double alphad, double sdep, int* err)
{
    const double tiny = 1.0e-4;
    double zeta = alphal / (alphal + alphad);
    double al = 0., dnu = 0.;
    cplx v;
    if (zeta < 1.00) {
        al = alphal / alphad;
        dnu = deltnu / alphad;
    }
    if (zeta == 1.00 && fabs(sdep) < tiny) return (alphal / (kPI * (alphal * alphal + deltnu * deltnu)));
    if (fabs(sdep) > tiny) {   // Boone et al. 2011 speed dependence, :1022-1066
        double gamma2 = alphal * sdep;
        double alfa = (alphal / gamma2) - 1.5;
        double beta = (deltnu / gamma2);
        double delta = (1.0 / 4.0 / log(2.)) * (alphad * alphad / gamma2 / gamma2);
        double alfadelta = alfa + delta;
        double temp = sqrt(alfadelta * alfadelta + beta * beta);
        double x1 = (1.0 / sqrt(2.0)) * sqrt(temp + alfadelta) - sqrt(delta);
        double x2 = x1 + 2.0 * sqrt(delta);
        double sign = (beta > 0.0) ? 1. : ((beta == 0.0) ? 0. : -1.);
        double y1 = sign * sqrt((temp - delta - alfa) / 2.0);
        double y2 = y1;
        v = sd_humlicek(y1, x1, y2, x2);
        if (v.re < 0.0) *err = 1;
    } else {
        double x = sqrt(log(2.)) * dnu;
        double y = 1000.;
        if (zeta < 1.000) y = sqrt(log(2.)) * al;
        v = w4(x, y);
    }
    double anorm1 = sqrt(log(2.) / kPI) / alphad;
    return v.re * anorm1;
}

// Real part of Humlicek region I (modm.f90:1105-1106, s = |x|+y >= 15): Re[t*.5641896/(.5+t*t)], t = (y,-x)
__device__ __forceinline__ double hum1_re(double x, double y)
{
    const double dre = .5 + (y * y - x * x), dim = -2. * (x * y);
    return (.5641896 * (y * dre - x * dim)) / (dre * dre + dim * dim);
}
__device__ __forceinline__ double w4_re(double x, double y)
{
    if (!(fabs(x) + y < 15.)) return hum1_re(x, y);
    return w4(x, y).re;
}

// Real part of W4 for voigt_kernel's staged lines: the same Humlicek regions, boundaries and constants as
// W4 (modm.f90:1100-1130); the complex quotients are formed as a*conj(b)/|b|^2 with the ~1 ulp reciprocal
// instead of IEEE divisions (the denominators are far from over/underflow for s < 15).
__device__ __forceinline__ double cdiv_re(cplx a, cplx b)
{
    return (a.re * b.re + a.im * b.im) * rcp3(b.re * b.re + b.im * b.im);
}
// the three near regions one by one ...
__device__ __forceinline__ double w4_re_region2(double x, double y)
{
    const cplx t = cmk(y, -x);
    const cplx u = t * t;
    return cdiv_re(t * (1.410474 + u * .5641896), .75 + u * (3. + u));
}
__device__ __forceinline__ double w4_re_region3(double x, double y)
{
    const cplx t = cmk(y, -x);
    return cdiv_re(16.4955 + t * (20.20933 + t * (11.96482 + t * (3.778987 + t * .5642236))),
                   16.4955 + t * (38.82363 + t * (39.27121 + t * (21.69274 + t * (6.699398 + t)))));
}
__device__ __forceinline__ double w4_re_region4(double x, double y)
{
    const cplx t = cmk(y, -x);
    const cplx u = t * t;
    const cplx num = t * (36183.31 - u * (3321.9905 - u * (1540.787 - u * (219.0313 - u * (35.76683 - u * (1.320522 - u * .56419))))));
    const cplx den = 32066.6 - u * (24322.84 - u * (9022.228 - u * (2186.181 - u * (364.2191 - u * (61.57037 - u * (1.841439 - u))))));
    return cexpd(u).re - cdiv_re(num, den);
}
// 0: region II (5.5 <= s < 15), 1: region III, 2: region IV -- the boundaries of modm.f90:1111, 1117 for s = |x|+y < 15
__device__ __forceinline__ int w4_near_region(double x, double y)
{
    if (!(fabs(x) + y < 5.5)) return 0;
    return !(y < 0.195 * fabs(x) - 0.176) ? 1 : 2;
}
// ... and behind one entry point
__device__ __noinline__ double w4_re_near(double x, double y)
{
    const int r = w4_near_region(x, y);
    if (r == 0) return w4_re_region2(x, y);
    if (r == 1) return w4_re_region3(x, y);
    return w4_re_region4(x, y);
}
__device__ __forceinline__ double w4_re_fast(double x, double y)
{
    if (!(fabs(x) + y < 15.)) {          // region I: Re[t*.5641896/(.5+t*t)]
        const double dre = .5 + (y * y - x * x), dim = -2. * (x * y);
        return (.5641896 * (y * dre - x * dim)) * rcp3(dre * dre + dim * dim);
    }
    return w4_re_near(x, y);
}

// The Voigt branch (modm.f90:427-431 -> LSF_SDVOIGT :567-704) for the line classes the line kernel
// streams: kind 0 generic molecule without coupling (Voigt pedestal at 25 cm-1, :588-599), 1 O2 without
// coupling (:618-626), 2 O2 XF=-3/-5 (:650-653), 3 O2 XF=-1 first-order mixing (:642-649).
// Returns STILD*SLS.  Same Humlicek regions and constants as SDVOIGT/W4; the three profile values share
// one reciprocal of the Doppler width, and the far ones (s >= 15) take region I inline.
struct ColdLine;
template <class CL>
__device__ __noinline__ double voigt_lines_term(int kind, double wn, double xnu, const CL cl, double sdep, double rp, double rp2, int* err)
{
    const double hw = cl.hw, ad = cl.ad, stild = cl.stild;
    const double dm = wn - xnu, sp = wn + xnu;
    const bool second = (kind >= 2) || ((sp - kDELTNUC) <= 0.);
    double y1 = 1., y2 = 1.;
    if (kind == 3) {
        const double aip = cl.aip, bip = cl.bip;
        y1 = (1. + (aip * (1 / hw) * rp * dm) + (bip * rp2));
        y2 = (1. - (aip * (1 / hw) * rp * sp) + (bip * rp2));
    }
    const double nped = (kind == 0) ? (second ? 2. : 1.) : 0.;
    const double zeta = hw / (hw + ad);
    if (fabs(sdep) > 1.0e-4 || !(zeta < 1.0)) {      // speed dependence / degenerate Doppler width: the general routine
        double sls = y1 * sdvoigt(dm, hw, ad, sdep, err);
        if (second) sls += y2 * sdvoigt(sp, hw, ad, sdep, err);
        if (nped != 0.) sls -= nped * sdvoigt(kDELTNUC, hw, ad, sdep, err);
        return stild * sls;
    }
    const double inv = 1. / ad;
    const double sl2 = 0.8325546111576977;         // sqrt(log(2))
    const double y = sl2 * (hw * inv);
    const double anorm = 0.46971863934982516 * inv;       // sqrt(log(2)/PI) with the reference's 13-digit PI
    double sls = y1 * w4_re(sl2 * (dm * inv), y);
    if (second) sls += y2 * w4_re(sl2 * (sp * inv), y);
    if (nped != 0.) sls -= nped * w4_re(sl2 * (kDELTNUC * inv), y);
    return stild * (sls * anorm);
}

// XLORENTZ(z)/HWHM: the Lorentz profile with the reference's truncated PI (modm.f90:888-895)
__device__ __forceinline__ double lorentz_profile(double d, double hwhm)
{
    double z = d / hwhm;
    return (1. / (kPI * (1. + (z * z)))) / hwhm;
}

// General line-shape case tree: LSF_LORTZ (modm.f90:706-831) when !voigt, LSF_SDVOIGT
// (:567-704) when voigt.  Returns SLS (already divided by HWHM in the Lorentz case).
// The CO2 typo XF.NE.-5 (:659) is kept: a CO2 line with XF=-5 contributes nothing on the
// Voigt branch.  chi_fn is identically 1 (:1286).
__device__ __noinline__ double lsf_general(int mol, int xf, double rp, double rp2, double aip, double bip,
                                           double hwhm, double wn, double xnu, double ad, double sdep,
                                           bool voigt, int* err)
{
    const double dc = kDELTNUC;
    auto shape = [&](double d) -> double {
        return voigt ? sdvoigt(d, hwhm, ad, sdep, err) : lorentz_profile(d, hwhm);
    };
    const bool lc = (xf == -1) || (xf == -3) || (xf == -5);
    const double diff = (wn + xnu) - dc;
    double sls = 0.;
    if (mol != 7 && mol != 2) {
        double xl1 = shape(wn - xnu), xl3 = shape(dc);
        if (lc) {
            double y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
            double y1p = (1. + (aip * (1 / hwhm) * rp * dc) + (bip * rp2));
            if (diff <= 0.) {
                double xl2 = shape(wn + xnu);
                double y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                double y2p = (1. - (aip * (1 / hwhm) * rp * dc) + (bip * rp2));
                sls = (y1 * xl1 - y1p * xl3 + y2 * xl2 - y2p * xl3);
            } else {
                sls = y1 * xl1 - y1p * xl3;
            }
        } else {
            if (diff <= 0.) sls = (xl1 + shape(wn + xnu) - (2 * xl3));
            else sls = (xl1 - xl3);
        }
    } else if ((fabs(wn - xnu) <= dc) && !lc) {
        double xl1 = shape(wn - xnu);
        if (mol == 7) {
            sls = (diff <= 0.) ? (xl1 + shape(wn + xnu)) : xl1;
        } else {
            double d = wn - xnu;
            double xl3 = shape(dc) * (2. - ((d * d) / (dc * dc)));
            sls = xl1 - xl3;
        }
    } else if (mol == 7) {
        if (lc) {
            double xl1 = shape(wn - xnu), xl2 = shape(wn + xnu);
            if (xf == -1) {
                double y1 = (1. + (aip * (1 / hwhm) * rp * (wn - xnu)) + (bip * rp2));
                double y2 = (1. - (aip * (1 / hwhm) * rp * (wn + xnu)) + (bip * rp2));
                sls = (xl1 * y1 + xl2 * y2);
            } else {
                sls = (xl1 + xl2);
            }
        }
    } else {   // CO2 with coupling flags
        bool take = voigt ? ((xf == -1) || (xf == -3) || (xf != -5)) : lc;
        if (take) {
            double d = wn - xnu;
            double xl1 = shape(d), xl3 = shape(dc);
            double q = (2. - (d * d) / (dc * dc));
            if (xf == -1 || xf == -5) {
                double y1 = (1. + (aip * (1 / hwhm) * rp * d) + (bip * rp2));
                double xp4 = xl3 * q;
                double yp1 = (y1 - 1.) * q;
                sls = (xl1 * y1 - xp4 - xl3 * yp1);
            } else {
                sls = (xl1 - xl3 * q);
            }
        }
    }
    return sls;
}

}  // namespace mrtm
