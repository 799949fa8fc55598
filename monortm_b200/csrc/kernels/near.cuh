// near.cuh -- near_kernel (streamed direct loops) and near2_kernel (per-warp re-planning of the staged near field).
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
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
__global__ void __launch_bounds__(NT, NT <= 32 ? MRTM_LINES_MINB_32 : (MRTM_LINES_MINB * 128) / NT) near_kernel(LinesArgs a)
{
    constexpr int NW = NT / 32;
    constexpr int kSt = near_stages<NT>();        // stages of the tile ring
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
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;

    __shared__ __align__(8) uint64_t s_bar[kSt];
    __shared__ __align__(8) uint64_t s_all_bar;
    __shared__ double s_ped[2][NW];
    __shared__ unsigned char s_act[kMaxSegments];
    extern __shared__ __align__(128) unsigned char s_dyn[];
    double (*s_tile)[4][kTile] = reinterpret_cast<double (*)[4][kTile]>(s_dyn);      // [kSt][4][kTile]
    SegWork* s_work = reinterpret_cast<SegWork*>(s_dyn + sizeof(double) * kSt * 4 * kTile);

    const int nseg = a.nseg;
    {
        const int nw = nseg * (int)(sizeof(SegWork) / 4);
        const int* src = reinterpret_cast<const int*>(a.plan[0] + (size_t)blockIdx.x * nseg);
        int* dst = reinterpret_cast<int*>(s_work);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        for (int s = tid; s < nseg; s += NT) s_act[s] = (ly.wk[a.seg[s].mol - 1] != 0.) ? 1 : 0;   // W_SPECIES == 0: skipped (:318-321)
    }
    if (tid == 0) {
        for (int i = 0; i < kSt; i++) mbar_init(&s_bar[i], 1);
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
    constexpr int kCap = kSt * kTile;
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
        for (int i = 0; i < kSt - 1 && more; i++) {
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

    int gtile = 0;              // global tile counter: stage = gtile % kSt, mbarrier parity = (gtile / kSt) & 1

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
    // every pair is evaluated with the Lorentz form here; voigt_kernel replaces it by the Voigt form where the reference
    // takes that branch (modm.f90:427), so the plan's Voigt zones run as plain near-field ranges
    const int vmode_mask = 0xff & ~M_VOIGT;
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
                        const int st = gtile % kSt;
                        if (tid == 0 && more) {        // refill the stage the previous tile released
                            more = advance(pjs, pjr, pjt);
                            if (more) issue(pjs, pjr, pjt, (gtile + kSt - 1) % kSt);
                        }
                        mbar_wait(&s_bar[st], (uint32_t)(gtile / kSt) & 1u);
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
                    const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q];
                    const double cq = lc1_slope(pCN[q], h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx_s[q]], rp);
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                        const double r1 = rcp3(fma(dm, dm, h2));
                        const double r2 = rcp3(fma(sp, sp, h2));
                        const double val = fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;
                        sf[f] += val;                                   // Voigt-branch pairs: corrected by voigt_kernel
                    }
                }
            }
        } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
            if (a.counters) n_direct += (long long)(wk.q1 - wk.q0) * nvalid;
            for (int q = wk.q0; q < wk.q1; q++) {
                const double xnu = pXNU[q], vt = pVT[q];
                const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                const double hw = cl.hw, ad = cl.ad, st = cl.stild, aip = cl.aip, bip = cl.bip;
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
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
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
    const int vmode_mask = 0xff & ~M_VOIGT;       // Voigt-branch pairs: Lorentz here, corrected by voigt_kernel
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
                        const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q];
                    const double cq = lc1_slope(pCN[q], h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx_s[q]], rp);
#pragma unroll
                        for (int f = 0; f < F; f++) {
                            const double dm = wn[f] - xnu, sp = wn[f] + xnu;
                            const double r1 = rcp3(fma(dm, dm, h2));
                            const double r2 = rcp3(fma(sp, sp, h2));
                            const double val = fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;
                            ss[f] += val;                                   // Voigt-branch pairs: corrected by voigt_kernel
                        }
                    }
                }
            } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
                if (a.counters) n_direct += (long long)(wk.q1 - wk.q0) * nvalid_w;
                for (int q = wk.q0; q < wk.q1; q++) {
                    const double xnu = pXNU[q], vt = pVT[q];
                    const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                    const double hw = cl.hw, ad = cl.ad, st = cl.stild, aip = cl.aip, bip = cl.bip;
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
