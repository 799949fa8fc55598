// voigt_t.cuh -- voigtT_kernel: the Voigt branch on dense frequency tiles, lines over warps.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// Same job and same arithmetic per (line, frequency) pair as voigt_kernel (modm.f90:427-431 -> LSF_SDVOIGT :567-704): for
// the pairs with |WN-Xnu| <= 100*HWHM_D it adds W*STILD*(SLS_Voigt - SLS_Lorentz) to O [and O_BY_MOL].  The mapping is
// transposed: on a dense grid the zone of a line covers a contiguous run of some 10 ... 300 frequencies of the tile, so
//   * the CTA (frequency tile, layer, profile) stages its zone lines once, two threads per line: one forms the
//     frequency-independent terms, the other brackets the line's run [lo, hi) of tile frequencies by two binary searches
//     in shared memory (exact: the rounded difference WN-Xnu is monotone in WN, so the reference's test is a monotone
//     predicate on an ascending tile; a tile that is not ascending keeps the whole tile as the run and relies on the
//     per-pair test),
//   * each warp takes every NW-th line, holds the line's terms in registers and strides its lanes over the run; the
//     per-pair test of the reference stays in the loop, so the selected pairs are unchanged,
//   * sums go to a per-warp accumulator tile in shared memory (no atomics: the order of the additions is fixed), the
//     warps' tiles are added in a fixed order and O makes one read-modify-write trip per frequency.
// voigt_kernel culls the lines per 32-frequency sub-block and walks the survivors serially with the line terms re-read
// from shared memory per pair; here a pair costs its arithmetic only.  Coarse tiles (vplan_kernel's candidate lists,
// a handful of frequencies per zone) stay with voigt_kernel.
// =============================================================================================
#ifndef MRTM_VOIGTT_MINB
#define MRTM_VOIGTT_MINB 7
#endif
#ifndef MRTM_VOIGTT_CAP
#define MRTM_VOIGTT_CAP 64
#endif
template <int F, int NT>
__global__ void __launch_bounds__(NT, MRTM_VOIGTT_MINB) voigtT_kernel(LinesArgs a)
{
    constexpr int NW = NT / 32, TF = NT * F;
    constexpr int kVCap = MRTM_VOIGTT_CAP;
    static_assert(2 * kVCap <= NT, "two staging threads per line of a chunk");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.y, prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    if (!voigt_possible(a.layer_voigt, L, a.hdr[0][blockIdx.x].whi)) return;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const SegWork* plan = a.plan[0] + (size_t)blockIdx.x * a.nseg;
    __shared__ double s_vt[kVCap], s_x[kVCap], s_inv[kVCap], s_y[kVCap], s_c[kVCap], s_pd[kVCap], s_g[kVCap], s_b[kVCap];
    __shared__ double s_h2[kVCap], s_cn[kVCap], s_p3[kVCap], s_p4[kVCap];
    __shared__ int s_q[kVCap];
    __shared__ short s_lo[kVCap], s_hi[kVCap];
    __shared__ unsigned char s_kind[kVCap], s_mol[kVCap], s_single[kVCap];
    __shared__ int s_zlo[kMaxSegments], s_zoff[kMaxSegments + 1];
    __shared__ double s_wn[TF];
    __shared__ double s_acc[NW][TF];
    const int i0 = blockIdx.x * TF;
    const int nf = min(TF, a.nwn - i0);
    const double rp = ly.rp, rp2 = ly.rp2;
    const double sl2 = 0.8325546111576977;         // sqrt(log(2))
    const bool by_mol = a.o_by_mol != nullptr;
    for (int s = tid; s < a.nseg; s += NT) {
        const Segment sg = a.seg[s];
        const bool use = (sg.cls != CLS_GENERAL) && (ly.wk[sg.mol - 1] != 0.);
        const int v0 = plan[s].v0, v1 = plan[s].v1;
        s_zlo[s] = v0;
        s_zoff[s + 1] = (use && v1 > v0) ? (v1 - v0) : 0;
    }
#pragma unroll
    for (int m = 0; m < F; m++) {
        const int j = tid + m * NT;
        s_wn[j] = a.wn[min(i0 + j, a.nwn - 1)];
#pragma unroll
        for (int w = 0; w < NW; w++) s_acc[w][j] = 0.;
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
    bool asc = true;
#pragma unroll
    for (int m = 0; m < F; m++) {
        const int j = tid + m * NT;
        if (j + 1 < nf && s_wn[j + 1] < s_wn[j]) asc = false;
    }
    const bool sorted = __syncthreads_and(asc) != 0;
    const int total = s_zoff[a.nseg];
    if (total == 0) return;
    const double wmin_tile = a.hdr[0][blockIdx.x].wlo;
    int err = 0;
    double* const acc = s_acc[wid];
    double tot_o[F];
#pragma unroll
    for (int m = 0; m < F; m++) tot_o[m] = 0.;
    double* const odst = (a.o_v ? a.o_v : a.o) + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;

    for (int e0 = 0; e0 < total; e0 += kVCap) {
        const int n = min(kVCap, total - e0);
        if (e0 > 0) __syncthreads();
        // two threads per line: one forms the frequency-independent terms, the other brackets the line's run of frequencies
        {
            const int i = tid & (kVCap - 1);
            const bool searcher = tid >= kVCap;
            if (i < n && tid < 2 * kVCap) {
                const int e = e0 + i;
                int sg = 0;
                while (s_zoff[sg + 1] <= e) sg++;
                const int q = s_zlo[sg] + (e - s_zoff[sg]);
                const double vt = __ldg(pVT + q);
                const double xq = __ldg(pXNU + q);
                if (searcher) {
                    int lo = 0, hi = 0;
                    if (vt >= 0.) {
                        // the run of tile frequencies that can pass |WN-Xnu| <= vt (the test itself is repeated per pair)
                        hi = nf;
                        if (sorted) {
                            int b = 0, e2 = nf;
                            while (b < e2) { const int mid = (b + e2) >> 1; if ((s_wn[mid] - xq) < -vt) b = mid + 1; else e2 = mid; }
                            lo = b;
                            e2 = nf;
                            while (b < e2) { const int mid = (b + e2) >> 1; if ((s_wn[mid] - xq) > vt) e2 = mid; else b = mid + 1; }
                            hi = b;
                        }
                    }
                    // no frequency of the run has the second resonance (WN+Xnu-25 <= 0, modm.f90:746; the rounded sum is
                    // monotone in WN)
                    const double wfirst = (sorted && lo < nf) ? s_wn[lo] : wmin_tile;
                    s_single[i] = (((wfirst + xq) - kDELTNUC) > 0.) ? 1 : 0;
                    s_lo[i] = (short)lo;
                    s_hi[i] = (short)hi;
                } else {
                    const int cls = a.seg[sg].cls, mol = a.seg[sg].mol;
                    const int kind = (cls == CLS_PED) ? 0 : ((cls == CLS_O2) ? 1 : ((cls == CLS_O2_LC35) ? 2 : 3));
                    s_vt[i] = vt;
                    s_x[i] = xq;
                    s_q[i] = q;
                    s_mol[i] = (unsigned char)mol;
                    int kflag = kind;
                    if (vt >= 0.) {
                        const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                        const double hw = cl.hw, ad = cl.ad;
                        // vt >= 0 says zeta <= 0.99 (derive_kernel): the degenerate case zeta >= 1 of SDVOIGT (modm.f90:1075)
                        // can only be a vanishing Doppler width
                        const bool degenerate = !(ad > 0.) || !(hw < 1e300);
                        const double wl = by_mol ? 1. : ly.wk[mol - 1];
                        const double h2 = __ldg(pH2 + q), cn = __ldg(pCN + q);
                        s_h2[i] = h2;
                        s_cn[i] = wl * cn;
                        s_p3[i] = wl * __ldg(pP3 + q);
                        s_p4[i] = (kind == 3) ? wl * lc1_slope(cn, h2, cl.aip, rp) : 0.;
                        if (fabs(__ldg(a.sdep_s + q)) > 1.0e-4 || degenerate) {
                            s_inv[i] = -1.;        // speed dependence / degenerate Doppler width: the general routine per pair
                            s_c[i] = wl;
                        } else {
                            const double inv = 1. / ad;
                            const double y = sl2 * (hw * inv);
                            s_inv[i] = inv;
                            s_y[i] = y;
                            s_c[i] = wl * (cl.stild * (0.46971863934982516 * inv));   // sqrt(log(2)/PI), 13-digit PI
                            s_pd[i] = (kind == 0) ? w4_re_fast(sl2 * (kDELTNUC * inv), y) : 0.;
                            s_g[i] = (kind == 3) ? (cl.aip * (1 / hw) * rp) : 0.;
                            s_b[i] = (kind == 3) ? (cl.bip * rp2) : 0.;
                            // fast form: plain line and the window test cannot fail inside the zone (the second resonance
                            // is the searcher's s_single)
                            if (kind == 0 && vt <= kDELTNUC) kflag |= 0x80;
                        }
                    }
                    s_kind[i] = (unsigned char)kflag;
                }
            }
        }
        __syncthreads();
        // groups of staged lines that share one sum: everything (the molecule's amount is folded into the line terms) or,
        // with per-molecule outputs, the runs of one molecule (the staged order is molecule-major)
        int g0 = 0;
        while (g0 < n) {
            int g1 = n;
            if (by_mol) {
                g1 = g0 + 1;
                while (g1 < n && s_mol[g1] == s_mol[g0]) g1++;
            }
            for (int i = g0 + wid; i < g1; i += NW) {          // warp-uniform
                const int lo = s_lo[i], hi = s_hi[i];
                if (lo >= hi) continue;
                const double xnu = s_x[i], vt = s_vt[i], inv = s_inv[i];
                const double h2 = s_h2[i], cn = s_cn[i], p3 = s_p3[i], c = s_c[i];
                const int kraw = s_kind[i];
                if ((kraw & 0x80) && s_single[i]) {
                    // Region I: Re w = y*(a+q)/(q*(q+b)+a*a), q = x*x, a = .5+y*y, b = 2*y*y-1 -- the reference's
                    // t*.5641896/(.5+t*t) (modm.f90:1105) multiplied out.  Every lane forms it; the lanes next to the centre
                    // replace it by their own region's value.
                    const double y = s_y[i], y2 = y * y, fa = .5 + y2, fb = 2. * y2 - 1., fa2 = fa * fa;
                    const double fcy = c * (.5641896 * y), fcpd = c * s_pd[i];
                    for (int jb = lo; jb < hi; jb += 32) {
                        const int j = jb + lane;
                        const bool inr = j < hi;
                        const double dm = s_wn[inr ? j : lo] - xnu;
                        const bool take = inr && (fabs(dm) <= vt);
                        const double x = sl2 * (dm * inv);
                        const double lor = fma(cn, rcp3(fma(dm, dm, h2)), -p3);
                        const double q = x * x;
                        double v = fcy * (fa + q) * rcp3(fma(q, q + fb, fa2));
                        if (take && (fabs(x) + y < 15.)) v = c * w4_re_near(x, y);     // regions II-IV next to the centre
                        if (take) acc[j] += (v - fcpd) - lor;
                    }
                    continue;
                }
                const int kind = kraw & 0x7f;
                const double p4 = s_p4[i];
                if (inv < 0.) {
                    const ColdLine cl = cold_line(pl, a.n_pad, s_q[i], a.lcidx_s, lcp, a.nlc_pad);
                    const double sdep = a.sdep_s[s_q[i]];
                    for (int j = lo + lane; j < hi; j += 32) {
                        const double wn = s_wn[j];
                        const double dm = wn - xnu;
                        const bool inwin = (kind <= 1) ? !(fabs(dm) > kDELTNUC) : true;
                        if (inwin && fabs(dm) <= vt) {
                            const double sp = wn + xnu;
                            const bool second = (kind >= 2) || ((sp - kDELTNUC) <= 0.);
                            const double r1 = rcp3(fma(dm, dm, h2));
                            double lor;
                            if (kind == 3) {
                                lor = fma(p4, dm, p3) * r1 + fma(-p4, sp, p3) * rcp3(fma(sp, sp, h2));
                            } else {
                                lor = cn * r1;
                                if (second) lor = fma(cn, rcp3(fma(sp, sp, h2)), lor);
                                if (kind == 0) lor -= (second ? 2. : 1.) * p3;
                            }
                            acc[j] += fma(c, voigt_lines_term(kind, wn, xnu, cl, sdep, rp, rp2, &err), -lor);
                        }
                    }
                    continue;
                }
                const double y = s_y[i], pd = s_pd[i], gg = s_g[i], bb = s_b[i];
                for (int j = lo + lane; j < hi; j += 32) {
                    const double wn = s_wn[j];
                    const double dm = wn - xnu;
                    const bool inwin = (kind <= 1) ? !(fabs(dm) > kDELTNUC) : true;
                    if (inwin && fabs(dm) <= vt) {
                        const double sp = wn + xnu;
                        const bool second = (kind >= 2) || ((sp - kDELTNUC) <= 0.);
                        // what the near / far kernels added for this pair (their regrouped Lorentz forms, modm.f90:742-791)
                        const double r1 = rcp3(fma(dm, dm, h2));
                        double lor;
                        if (kind == 3) {
                            lor = fma(p4, dm, p3) * r1 + fma(-p4, sp, p3) * rcp3(fma(sp, sp, h2));
                        } else {
                            lor = cn * r1;
                            if (second) lor = fma(cn, rcp3(fma(sp, sp, h2)), lor);
                            if (kind == 0) lor -= (second ? 2. : 1.) * p3;
                        }
                        double sls = w4_re_fast(sl2 * (dm * inv), y);
                        if (kind == 3) sls *= (1. + (gg * dm) + bb);
                        if (second) {
                            double v2 = w4_re_fast(sl2 * (sp * inv), y);
                            if (kind == 3) v2 *= (1. - (gg * sp) + bb);
                            sls += v2;
                        }
                        if (kind == 0) sls -= (second ? 2. : 1.) * pd;
                        acc[j] += fma(c, sls, -lor);
                    }
                }
            }
            if (by_mol) {           // close the molecule's sum: O_BY_MOL = W * sum (modm.f90:436-438), then start the next one
                const int mol = s_mol[g0];
                __syncthreads();
#pragma unroll
                for (int m = 0; m < F; m++) {
                    const int j = tid + m * NT;
                    double msum = 0.;
#pragma unroll
                    for (int w = 0; w < NW; w++) { msum += s_acc[w][j]; s_acc[w][j] = 0.; }
                    if (msum != 0. && j < nf) {
                        const double ol = ly.wk[mol - 1] * msum;
                        tot_o[m] += ol;
                        a.o_by_mol[(size_t)(i0 + j) + (size_t)(mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] += ol;
                    }
                }
                __syncthreads();
            }
            g0 = g1;
        }
    }
    if (!by_mol) {
        __syncthreads();
#pragma unroll
        for (int m = 0; m < F; m++) {
            const int j = tid + m * NT;
#pragma unroll
            for (int w = 0; w < NW; w++) tot_o[m] += s_acc[w][j];
        }
    }
#pragma unroll
    for (int m = 0; m < F; m++) {
        const int j = tid + m * NT;
        if (j < nf && tot_o[m] != 0.) odst[i0 + j] += tot_o[m];
    }
    if (err) atomicOr(a.errflag, 2);
}

