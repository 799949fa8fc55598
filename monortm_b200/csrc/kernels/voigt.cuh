// voigt.cuh -- voigt_kernel: the Voigt branch.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// voigt_kernel: the Voigt branch (modm.f90:427-431).  CTA = (frequency tile, layer, profile); it leaves
// at once when the layer has no Voigt-capable line.  For the lines of the plan's Voigt zones it applies
// the reference's test |WN-Xnu| <= 100*HWHM_D per (line, frequency) and, for the pairs that pass, adds
// W*STILD*(SLS_Voigt - SLS_Lorentz) to O [and O_BY_MOL]: the near- and far-field kernels evaluate every pair with the
// Lorentz form (they never look at the Doppler width), this kernel replaces it where the reference takes the Voigt
// branch.  Same result as evaluating those pairs once with the Voigt form, to rounding.
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
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const SegWork* plan = a.plan[0] + (size_t)blockIdx.x * a.nseg;
    // The zone lines of all segments are staged together (one barrier pair per CTA in the usual case), and per zone
    // line, once per CTA (amortised over the NT*F frequencies), everything of LSF_SDVOIGT/SDVOIGT that does not depend
    // on the frequency: 1/alphaD, y = sqrt(ln2)*alphaL/alphaD, STILD*sqrt(ln2/pi)/alphaD, the Voigt pedestal at
    // 25 cm-1 (modm.f90:590) and the mixing factors (:595-596).
#ifndef MRTM_VOIGT_CAP
#define MRTM_VOIGT_CAP 64
#endif
    constexpr int kVCap = MRTM_VOIGT_CAP;
    __shared__ double s_vt[kVCap], s_x[kVCap], s_inv[kVCap], s_y[kVCap], s_c[kVCap], s_pd[kVCap], s_g[kVCap], s_b[kVCap];
    __shared__ int s_q[kVCap];
    __shared__ unsigned char s_kind[kVCap], s_mol[kVCap];
    // fast path (generic uncoupled line, single resonance, Humlicek region I): Re w = y*(a+q)/(q*(q+b)+a*a), q = x*x,
    // a = .5+y*y, b = 2*y*y-1 -- the reference's t*.5641896/(.5+t*t) (modm.f90:1105) multiplied out
    __shared__ double s_fa[kVCap], s_fb[kVCap], s_fa2[kVCap], s_fcy[kVCap], s_fcpd[kVCap];
    // the Lorentz form the other kernels added for the same pair: H2, W*CN, W*P3 (pedestal or mixing numerator), W*P4
    __shared__ double s_h2[kVCap], s_cn[kVCap], s_p3[kVCap], s_p4[kVCap];
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
    // coarse tiles: only the zone lines that have a frequency of the tile within reach (vplan_kernel)
    const int ncand = a.vcand_count ? a.vcand_count[blockIdx.x] : -1;
    const int* cand = a.vcand ? a.vcand + (size_t)blockIdx.x * kVCandMax : nullptr;
    const unsigned char* cand_seg = a.vcand_seg ? a.vcand_seg + (size_t)blockIdx.x * kVCandMax : nullptr;
    const int total = (ncand >= 0) ? ncand : s_zoff[a.nseg];
    if (total == 0) return;
    int err = 0;
    unsigned short* lst = s_list[wid];
    double* const odst = (a.o_v ? a.o_v : a.o) + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
    for (int e0 = 0; e0 < total; e0 += kVCap) {
        const int n = min(kVCap, total - e0);
        if (e0 > 0) __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const int e = e0 + i;
            int sg = 0, q;
            if (ncand >= 0) {
                sg = cand_seg[e];
                q = cand[e];
            } else {
                while (s_zoff[sg + 1] <= e) sg++;
                q = s_zlo[sg] + (e - s_zoff[sg]);
            }
            const int cls = a.seg[sg].cls, mol = a.seg[sg].mol;
            const int kind = (cls == CLS_PED) ? 0 : ((cls == CLS_O2) ? 1 : ((cls == CLS_O2_LC35) ? 2 : 3));
            // a candidate of a molecule without amount in this layer is skipped like its whole segment (:318-321)
            const double vt = (ncand >= 0 && ly.wk[mol - 1] == 0.) ? -1. : __ldg(pVT + q);
            s_vt[i] = vt;
            s_x[i] = __ldg(pXNU + q);
            s_q[i] = q;
            s_kind[i] = (unsigned char)kind;
            s_mol[i] = (unsigned char)mol;
            if (vt >= 0.) {
                const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                const double hw = cl.hw, ad = cl.ad;
                const double zeta = hw / (hw + ad);
                {
                    const double wl = by_mol ? 1. : ly.wk[mol - 1];
                    s_h2[i] = __ldg(pH2 + q);
                    s_cn[i] = wl * __ldg(pCN + q);
                    s_p3[i] = wl * __ldg(pP3 + q);
                    s_p4[i] = (kind == 3) ? wl * lc1_slope(__ldg(pCN + q), s_h2[i], cl.aip, rp) : 0.;
                }
                if (fabs(__ldg(a.sdep_s + q)) > 1.0e-4 || !(zeta < 1.0)) {
                    s_inv[i] = -1.;        // speed dependence / degenerate Doppler width: the general routine per pair
                    s_c[i] = by_mol ? 1. : ly.wk[mol - 1];
                } else {
                    const double inv = 1. / ad;
                    const double y = sl2 * (hw * inv);
                    const double wgt = by_mol ? 1. : ly.wk[mol - 1];           // one sum over all molecules: weight folded in
                    s_inv[i] = inv;
                    s_y[i] = y;
                    s_c[i] = wgt * (cl.stild * (0.46971863934982516 * inv));   // sqrt(log(2)/PI), 13-digit PI
                    s_pd[i] = (kind == 0) ? w4_re_fast(sl2 * (kDELTNUC * inv), y) : 0.;
                    s_g[i] = (kind == 3) ? (cl.aip * (1 / hw) * rp) : 0.;
                    s_b[i] = (kind == 3) ? (cl.bip * rp2) : 0.;
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
            // what the near / far kernels added for this pair (their regrouped Lorentz forms, modm.f90:742-791)
            auto lorentz_of = [&](const int i, const int kind, const double dm, const double sp, const bool second) -> double {
                const double h2 = s_h2[i], cn = s_cn[i];
                const double r1 = rcp3(fma(dm, dm, h2));
                if (kind == 3) return fma(s_p4[i], dm, s_p3[i]) * r1 + fma(-s_p4[i], sp, s_p3[i]) * rcp3(fma(sp, sp, h2));
                double v = cn * r1;
                if (second) v = fma(cn, rcp3(fma(sp, sp, h2)), v);
                if (kind == 0) v -= (second ? 2. : 1.) * s_p3[i];
                return v;
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
                        const double lor = fma(s_cn[i], rcp3(fma(dm, dm, s_h2[i])), -s_p3[i]);
                        if (!(fabs(x) + y < 15.)) {
                            const double q = x * x;
                            const double den = fma(q, q + s_fb[i], s_fa2[i]);
                            msum += fma(s_fcy[i] * (s_fa[i] + q), rcp3(den), -s_fcpd[i]) - lor;
                        } else {
                            msum = (fma(s_c[i], w4_re_near(x, y), msum) - s_fcpd[i]) - lor;
                        }
                        many = true;
                    }
                    return;
                }
                const int kind = s_kind[i] & 0x7f;
                const bool inwin = (kind <= 1) ? !(fabs(dm) > kDELTNUC) : true;
                if (inwin && fabs(dm) <= s_vt[i]) {
                    const double inv = s_inv[i];
                    {
                        const double spl = wn + xnu;
                        msum -= lorentz_of(i, kind, dm, spl, (kind >= 2) || ((spl - kDELTNUC) <= 0.));
                    }
                    if (inv < 0.) {
                        msum = fma(s_c[i], voigt_lines_term(kind, wn, xnu, cold_line(pl, a.n_pad, s_q[i], a.lcidx_s, lcp, a.nlc_pad), a.sdep_s[s_q[i]], rp, rp2, &err), msum);
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
                    const double l0 = fma(s_cn[i0], rcp3(fma(dm0, dm0, s_h2[i0])), -s_p3[i0]);
                    const double l1 = fma(s_cn[i1], rcp3(fma(dm1, dm1, s_h2[i1])), -s_p3[i1]);
                    if (in0) msum = (r0 ? (msum + v0) : (fma(s_c[i0], w4_re_near(x0, y0), msum) - s_fcpd[i0])) - l0;
                    if (in1) msum = (r1 ? (msum + v1) : (fma(s_c[i1], w4_re_near(x1, y1), msum) - s_fcpd[i1])) - l1;
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
