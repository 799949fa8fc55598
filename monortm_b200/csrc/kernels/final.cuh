// final.cuh -- final_kernel: far-field polynomial, RFT, continuum, cloud, totals.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// final_kernel: per (frequency, layer): the far-field polynomial of the level-0 tile, RFT (modm.f90:257),
// the continuum interpolation + RADFN (:218-230), cloud liquid water (:264) and the total (:265-269).
// =============================================================================================
#ifndef MRTM_FINAL_MINB
#define MRTM_FINAL_MINB 10
#endif
// kMwOnly: the call lies below 820 cm-1 with no O3 / O2 / Rayleigh component (the usual microwave case): their planes are
// not touched
template <int F, int NT, bool kMwOnly>
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
    const double hinv_2t = 0.5 * ly.inv_t;
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
            const double rft = wn[f] * tanh((ly.radct * wn[f]) * hinv_2t);   // modm.f90:257 (the division by 2T as a product)
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
                const double rft = wn[f] * tanh((ly.radct * wn[f]) * hinv_2t);
                const double ol = (w == 0.) ? 0. : rft * (a.o_by_mol[idx] + w * p);     // modm.f90:436-438
                a.o_by_mol[idx] = ol;
                osum[f] = osum[f] + ol;                                              // :265-267 (molecule order)
            }
        }
    }

    // ---- epilogue: continuum, cloud, totals ---------------------------------------------------
    const double* ab = a.absrb + (size_t)L * CP_COUNT * a.nptabs_pad;
    const int cont_mol[5] = {1, 2, 3, 7, 22};                          // index_cont, modm.f90:165-166 (planes in CP_* order)
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
        const double rf = radfn_r(wn[f], ly.xkt);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            if (kMwOnly && (c == CP_O3 || c == CP_O2)) continue;       // microwave call: O3 and O2 have no component (compile time)
            if (!((a.cont_mask >> c) & 1)) continue;                   // no component of this species fires: its oc stays 0
            double v = 0.;
            if (in_rng) v = 0. + xint_point(ab + (size_t)c * a.nptabs_pad, a.v1abs, 1.0, vi) * 1.0;
            v = v * rf;
            soc = soc + v;                                             // sum(oc(m,1:22,k)) in index order
            if (a.oc) a.oc[(size_t)iw + (size_t)(cont_mol[c] - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] = v;
        }
        double oray = 0.;                                              // modm.f90:231-244 (zero below 820 cm-1)
        if (!kMwOnly && ((a.cont_mask >> CP_RAYL) & 1)) {
            if (in_rng) oray = 0. + xint_point(ab + (size_t)CP_RAYL * a.nptabs_pad, a.v1abs, 1.0, vi) * 1.0;
            oray = oray * wn[f] / 1.0e4;
        }
        double oclw = (ly.clw == 0.) ? 0. : odclw_tkc(wn[f], ly.t, ly.clw);   // modm.f90:264
        double odx = a.odxsec ? a.odxsec[fl] : 0.;
        double tot = osum[f] + odx + oray + soc + oclw;                // :268-269
        a.o[fl] = tot;
        if (a.o_clw) a.o_clw[fl] = oclw;
    }
}
