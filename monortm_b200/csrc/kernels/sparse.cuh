// sparse.cuh -- nearT_kernel: the direct (line, layer, frequency) evaluations of coarse frequency lists (channel sets,
// log-spaced ensembles: BASELINE configs 1, 2, 4), and vplan_kernel: the Voigt-zone candidates of such tiles.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// near_kernel gives every thread a frequency and walks the lines; the reference's per-(line, frequency) tests (window
// modm.f90:384, second resonance :746) then cost a band of slow, fully tested iterations as wide as the TILE's spectral
// extent -- on a channel list a 128-frequency tile spans several cm-1 and 40 % of its in-window lines fall in such bands
// (measured: 31 % of the FP64 peak on the 1000-channel ensemble).  nearT_kernel transposes the mapping:
//   * CTA = (tile of 128 frequencies, layer, profile), 8 warps; a warp owns 16 frequencies of the tile and keeps them in
//     registers (the same values in every lane); the lanes own LINES, four each, so the tests are per-lane predicates on
//     numerators (2-3 extra instructions) instead of separate loops, and the four-lines-one-reciprocal form
//     sum c_i/a_i = N/(a1 a2 a3 a4) works inside a lane
//     (measured on B200 against near_kernel: 27 % faster on 19 sounder channels, about even on the 1000-channel log-spaced
//     ensemble, 24 % slower on the 5.5e-3 cm-1 grid of the 300-layer case: it is selected for calls with fewer than 64
//     frequencies, mrtm_api.cu; MRTM_NEART=2 selects it for every coarse-tile call)
//   * the tile's direct runs (plan_kernel: everything the far field does not take) stream through shared memory with the
//     same TMA ring as near_kernel: one 128-line tile feeds 32 lanes x 4 lines, every warp reads it once for 16 frequencies
//   * sub-ranges without tests take two predicate-free forms (single resonance: 23 FP64 instructions per 4 pairs; both
//     resonances: 44 per 4 pairs, i.e. 5.5 per Lorentzian); sub-ranges with a test take the general form (selects on
//     numerator and denominator, still one reciprocal per 4 lines)
//   * the column amount W is folded into the strengths as they are loaded; per-frequency sums are reduced across the warp
//     once per (tile, layer) -- or once per molecule when per-molecule outputs are requested
//   * the classes that are not streamed (first-order O2 mixing, the general case tree) are few lines: they run afterwards
//     with one thread per frequency exactly as in near_kernel
// The selected line set is unchanged: every pair runs the literal tests on the bit-exact shifted centre.
// =============================================================================================
#ifndef MRTM_NT_C
#define MRTM_NT_C 16
#endif
#ifndef MRTM_NT_MINB
#define MRTM_NT_MINB 2
#endif
#ifndef MRTM_NT_UNROLL
#define MRTM_NT_UNROLL 4
#endif
#ifndef MRTM_NT_VOL
#define MRTM_NT_VOL
#endif
constexpr int kNTC = MRTM_NT_C;                // frequencies per warp
constexpr int kNTW = 128 / kNTC;               // warps per CTA (kNTW * kNTC = 128 frequencies: the F = 1 tile)

template <bool SEL>
__global__ void __launch_bounds__(32 * kNTW, MRTM_NT_MINB) nearT_kernel(LinesArgs a)
{
    constexpr int NT = 32 * kNTW, C = kNTC;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.y, prof = blockIdx.z;
    const int64_t L = (int64_t)prof * a.nlay + k;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;

    __shared__ __align__(8) uint64_t s_bar[kStages];
    __shared__ unsigned char s_act[kMaxSegments];
    __shared__ double s_sum[128], s_wn[128];
    __shared__ long long s_cnt[SEL ? 128 : 1];
    __shared__ unsigned long long s_hsh[SEL ? 128 : 1];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // this warp's frequencies (every lane holds all of them)
    const int base = blockIdx.x * 128;
        // (kept in shared memory and read as broadcasts where they are used: 32 registers less per thread)
    if (tid < 128) s_wn[tid] = a.wn[(base + tid) < a.nwn ? (base + tid) : (a.nwn - 1)];
    const MRTM_NT_VOL double* wn = s_wn + wid * C;
    __syncthreads();
    double wA = 1e300, wB = -1e300;                // extent of this warp's frequencies
#pragma unroll
    for (int c = 0; c < C; c++) { wA = fmin(wA, wn[c]); wB = fmax(wB, wn[c]); }
    // the thread-per-frequency view used for the unstreamed classes and the final store
    const int iw_t = base + tid;
    const bool valid_t = tid < 128 && iw_t < a.nwn;
    const double wn_t = a.wn[(tid < 128 && iw_t < a.nwn) ? iw_t : (a.nwn - 1)];
    const double rp = ly.rp, rp2 = ly.rp2;
    __syncthreads();

    // ---- TMA tile jobs: (segment, run, tile) in consumption order; thread 0 keeps a cursor ahead of the consumers
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
    auto issue = [&](int js, int jr, int jt, int st) {
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
    int pjs = 0, pjr = 0, pjt = -1;
    bool more = true;
    if (tid == 0) {
        for (int i = 0; i < kStages - 1 && more; i++) {
            more = advance(pjs, pjr, pjt);
            if (more) issue(pjs, pjr, pjt, i);
        }
    }
    int gtile = 0;

    double acc[C];               // this lane's share of W*SF of the current molecule (or of all molecules), per frequency
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.;
    double ped_lane = 0.;        // pedestals of the untested sub-ranges (the same for every frequency)
    long long cnt[SEL ? C : 1];
    unsigned long long hsh[SEL ? C : 1];
    long long cnt_u = 0;         // selection counts that are the same for every frequency (lane 0's copy is used)
    unsigned long long hsh_u = 0ull;
    long long cnt_l = 0;         // this lane's lines that were selected for every frequency of the warp
    unsigned long long hsh_l = 0ull;
    if (SEL) {
#pragma unroll
        for (int c = 0; c < C; c++) { cnt[c] = 0; hsh[c] = 0ull; }
    }
    double osum_t = 0., tail_t = 0.;          // thread-per-frequency: closed molecules, unstreamed classes of the current molecule
    long long cnt_t = 0;
    unsigned long long hsh_t = 0ull;
    int err = 0;
    long long n_direct = 0;
    int nvalid = 0;
    if (a.counters) {
        const int rem = a.nwn - base;
        nvalid = rem < 128 ? rem : 128;
    }
    const bool by_mol = a.o_by_mol != nullptr;

    // warp sums of the per-frequency accumulators -> s_sum[frequency of the tile]
    auto reduce_to_smem = [&]() {
#pragma unroll
        for (int c = 0; c < C; c++) {
            double v = acc[c] - ped_lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) s_sum[wid * C + c] = v;
            acc[c] = 0.;
        }
        ped_lane = 0.;
    };
    int cur_mol = 0;
    auto finish_mol = [&](int mol) {
        if (mol <= 0 || !by_mol) return;
        reduce_to_smem();
        __syncthreads();
        if (tid < 128) {
            const double ol = s_sum[tid] + tail_t;                   // W*SF; RFT is applied by final_kernel (modm.f90:436-438)
            osum_t = osum_t + ol;                                     // :265-267 (molecule order)
            if (valid_t) a.o_by_mol[(size_t)iw_t + (size_t)(mol - 1) * a.obm_ldm + (size_t)L * a.obm_ldk] = ol;
            tail_t = 0.;
        }
        __syncthreads();
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
        const double w = ly.wk[sg.mol - 1];
        if (SEL) {
            if (sg.mol == 7) {                             // every O2 line passes modm.f90:384
                cnt_u += sg.count_all; hsh_u += sg.hash_all;
            } else if (cls == CLS_PED) {                   // every far line (any level) is inside the window of every frequency
                for (int u = 0; u + 1 < wk.nbp; u++) {
                    if (wk.mode[u] != 0) continue;
                    cnt_u += wk.bp[u + 1] - wk.bp[u];
                    hsh_u += a.keypre[wk.bp[u + 1]] - a.keypre[wk.bp[u]];
                }
            }
        }
        if (cls == CLS_PED || cls == CLS_O2 || cls == CLS_O2_LC35) {
            const bool force_both = (cls == CLS_O2_LC35);
            const bool has_win = !force_both;                  // the first-order O2 pairs (-3) are not window tested (:384 I.NE.7 ...)
            const bool count_sel = SEL && (cls == CLS_PED);
            for (int r = 0; r < wk.nrun; r++) {
                const int rlo = wk.run_lo[r], rhi = wk.run_hi[r], t0 = wk.run_t0[r];
                for (int t = 0; t < wk.run_nt[r]; t++) {
                    const int st = gtile % kStages;
                    if (tid == 0 && more) {            // refill the stage the previous tile released (barrier at the end of the loop body)
                        more = advance(pjs, pjr, pjt);
                        if (more) issue(pjs, pjr, pjt, (gtile + kStages - 1) % kStages);
                    }
                    mbar_wait(&s_bar[st], (uint32_t)(gtile / kStages) & 1u);
                    const int tb = t0 + t * kTile;
                    const int thi = (tb + kTile) < rhi ? (tb + kTile) : rhi;
                    const int tlo = tb > rlo ? tb : rlo;
                    gtile++;
                    // this lane's four lines of the tile
                    const int j = 4 * lane, q = tb + j;
                    const double2 xa = *reinterpret_cast<const double2*>(&s_tile[st][0][j]), xb = *reinterpret_cast<const double2*>(&s_tile[st][0][j + 2]);
                    const double2 ga = *reinterpret_cast<const double2*>(&s_tile[st][1][j]), gb = *reinterpret_cast<const double2*>(&s_tile[st][1][j + 2]);
                    const double2 ca = *reinterpret_cast<const double2*>(&s_tile[st][2][j]), cb = *reinterpret_cast<const double2*>(&s_tile[st][2][j + 2]);
                    const double2 pa = *reinterpret_cast<const double2*>(&s_tile[st][3][j]), pb = *reinterpret_cast<const double2*>(&s_tile[st][3][j + 2]);
                    // one pass per tile.  Whether a test can fail is decided by the warp for ITS frequencies and the layer's actual
                    // centres: with wA, wB the extent of the warp's frequencies, a line is inside the window for all of them when it is
                    // for wA and wB (the rounded difference is monotone in the frequency), has the second resonance for all when
                    // wB + Xnu <= 25 and for none when wA + Xnu > 25 (modm.f90:384, 746).  If that holds for every line of the
                    // tile a predicate-free form runs, otherwise the general form with the reference's tests as predicates.
                    {
                        const int lo = tlo, hi = thi;
                        if (lo >= hi) { __syncthreads(); continue; }
                        if (a.counters) n_direct += (long long)(hi - lo) * nvalid;
                        // lines of the tile outside the sub-range (or beyond the run) contribute nothing: zero strength, safe widths
                        const bool m1 = (q >= lo) && (q < hi), m2 = (q + 1 >= lo) && (q + 1 < hi);
                        const bool m3 = (q + 2 >= lo) && (q + 2 < hi), m4 = (q + 3 >= lo) && (q + 3 < hi);
                        const double x1 = m1 ? xa.x : 0., x2 = m2 ? xa.y : 0., x3 = m3 ? xb.x : 0., x4 = m4 ? xb.y : 0.;   // (stale smem may hold anything)
                        const double g1 = m1 ? ga.x : 1., g2 = m2 ? ga.y : 1., g3 = m3 ? gb.x : 1., g4 = m4 ? gb.y : 1.;
                        const double c1 = m1 ? w * ca.x : 0., c2 = m2 ? w * ca.y : 0., c3 = m3 ? w * cb.x : 0., c4 = m4 ? w * cb.y : 0.;
                        const double p1 = m1 ? w * pa.x : 0., p2 = m2 ? w * pa.y : 0., p3 = m3 ? w * pb.x : 0., p4 = m4 ? w * pb.y : 0.;
                        bool in_all = true, both_all = true, none_all = true;
                        {
                            const double xs[4] = {x1, x2, x3, x4};
                            const bool ms[4] = {m1, m2, m3, m4};
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                if (!ms[i]) continue;
                                if (has_win) in_all = in_all && !(fabs(wA - xs[i]) > kDELTNUC) && !(fabs(wB - xs[i]) > kDELTNUC);
                                both_all = both_all && ((wB + xs[i]) <= kDELTNUC);
                                none_all = none_all && !((wA + xs[i]) <= kDELTNUC);
                            }
                        }
                        const bool w_in = __all_sync(0xffffffffu, in_all);
                        const bool w_both = force_both || __all_sync(0xffffffffu, both_all);
                        const bool w_none = !force_both && __all_sync(0xffffffffu, none_all);
                        unsigned long long k1 = 0, k2 = 0, k3 = 0, k4 = 0;
                        if (count_sel) {
                            k1 = m1 ? a.key[q] : 0ull; k2 = m2 ? a.key[q + 1] : 0ull;
                            k3 = m3 ? a.key[q + 2] : 0ull; k4 = m4 ? a.key[q + 3] : 0ull;
                        }
                        if (w_in && w_none) {
                            // ---- no test can fail, single resonance (modm.f90:751): N/(a1 a2 a3 a4)
                            if (count_sel) { cnt_l += (int)m1 + (int)m2 + (int)m3 + (int)m4; hsh_l += (k1 + k2) + (k3 + k4); }
                            ped_lane += (p1 + p2) + (p3 + p4);
MRTM_UNROLL(MRTM_NT_UNROLL)
                            for (int c = 0; c < C; c++) {
                                const double d1 = wn[c] - x1, d2 = wn[c] - x2, d3 = wn[c] - x3, d4 = wn[c] - x4;
                                const double a1 = fma(d1, d1, g1), a2 = fma(d2, d2, g2), a3 = fma(d3, d3, g3), a4 = fma(d4, d4, g4);
                                const double p12 = a1 * a2, p34 = a3 * a4;
                                const double n12 = fma(c1, a2, c2 * a1), n34 = fma(c3, a4, c4 * a3);
                                acc[c] = fma(fma(n12, p34, n34 * p12), rcp3(p12 * p34), acc[c]);
                            }
                        } else if (w_in && w_both) {
                            // ---- no test can fail, both resonances: c*(aa+bb)/(aa*bb) per line, four lines share the reciprocal
                            if (count_sel) { cnt_l += (int)m1 + (int)m2 + (int)m3 + (int)m4; hsh_l += (k1 + k2) + (k3 + k4); }
                            ped_lane += 2. * ((p1 + p2) + (p3 + p4));
MRTM_UNROLL(MRTM_NT_UNROLL)
                            for (int c = 0; c < C; c++) {
                                const double wv = wn[c];
                                const double d1 = wv - x1, d2 = wv - x2, d3 = wv - x3, d4 = wv - x4;
                                const double s1 = wv + x1, s2 = wv + x2, s3 = wv + x3, s4 = wv + x4;
                                const double a1 = fma(d1, d1, g1), a2 = fma(d2, d2, g2), a3 = fma(d3, d3, g3), a4 = fma(d4, d4, g4);
                                const double b1 = fma(s1, s1, g1), b2 = fma(s2, s2, g2), b3 = fma(s3, s3, g3), b4 = fma(s4, s4, g4);
                                const double D1 = a1 * b1, D2 = a2 * b2, D3 = a3 * b3, D4 = a4 * b4;
                                const double N1 = c1 * (a1 + b1), N2 = c2 * (a2 + b2), N3 = c3 * (a3 + b3), N4 = c4 * (a4 + b4);
                                const double D12 = D1 * D2, D34 = D3 * D4;
                                const double N12 = fma(N1, D2, N2 * D1), N34 = fma(N3, D4, N4 * D3);
                                acc[c] = fma(fma(N12, D34, N34 * D12), rcp3(D12 * D34), acc[c]);
                            }
                        } else {
                            // ---- a test can fail for some (line, frequency): the reference's tests as predicates
                            const bool edge = has_win;
MRTM_UNROLL(MRTM_NT_UNROLL)
                            for (int c = 0; c < C; c++) {
                                const double wv = wn[c];
                                const double d1 = wv - x1, d2 = wv - x2, d3 = wv - x3, d4 = wv - x4;
                                const double s1 = wv + x1, s2 = wv + x2, s3 = wv + x3, s4 = wv + x4;
                                const bool t1 = edge ? !(fabs(d1) > kDELTNUC) : true, t2 = edge ? !(fabs(d2) > kDELTNUC) : true;
                                const bool t3 = edge ? !(fabs(d3) > kDELTNUC) : true, t4 = edge ? !(fabs(d4) > kDELTNUC) : true;
                                const bool e1 = force_both || (s1 <= kDELTNUC), e2 = force_both || (s2 <= kDELTNUC);
                                const bool e3 = force_both || (s3 <= kDELTNUC), e4 = force_both || (s4 <= kDELTNUC);
                                if (count_sel) {
                                    cnt[c] += (int)(t1 && m1) + (int)(t2 && m2) + (int)(t3 && m3) + (int)(t4 && m4);
                                    hsh[c] += (t1 ? k1 : 0ull) + (t2 ? k2 : 0ull) + (t3 ? k3 : 0ull) + (t4 ? k4 : 0ull);
                                }
                                const double a1 = fma(d1, d1, g1), a2 = fma(d2, d2, g2), a3 = fma(d3, d3, g3), a4 = fma(d4, d4, g4);
                                const double b1 = fma(s1, s1, g1), b2 = fma(s2, s2, g2), b3 = fma(s3, s3, g3), b4 = fma(s4, s4, g4);
                                const double cc1 = t1 ? c1 : 0., cc2 = t2 ? c2 : 0., cc3 = t3 ? c3 : 0., cc4 = t4 ? c4 : 0.;
                                const double D1 = e1 ? a1 * b1 : a1, D2 = e2 ? a2 * b2 : a2, D3 = e3 ? a3 * b3 : a3, D4 = e4 ? a4 * b4 : a4;
                                const double N1 = e1 ? cc1 * (a1 + b1) : cc1, N2 = e2 ? cc2 * (a2 + b2) : cc2;
                                const double N3 = e3 ? cc3 * (a3 + b3) : cc3, N4 = e4 ? cc4 * (a4 + b4) : cc4;
                                const double D12 = D1 * D2, D34 = D3 * D4;
                                const double N12 = fma(N1, D2, N2 * D1), N34 = fma(N3, D4, N4 * D3);
                                double v = fma(N12, D34, N34 * D12) * rcp3(D12 * D34);
                                v -= (t1 ? (e1 ? 2. * p1 : p1) : 0.) + (t2 ? (e2 ? 2. * p2 : p2) : 0.);
                                v -= (t3 ? (e3 ? 2. * p3 : p3) : 0.) + (t4 ? (e4 ? 2. * p4 : p4) : 0.);
                                acc[c] += v;
                            }
                        }
                    }
                    __syncthreads();          // every warp is done with this stage before it is refilled
                }
            }
        } else if (cls == CLS_O2_LC1) {
            if (tid < 128) {
                double sf = 0.;
                for (int r = 0; r < wk.nrun; r++) {
                    if (a.counters) n_direct += (tid == 0) ? (long long)(wk.run_hi[r] - wk.run_lo[r]) * nvalid : 0;
                    for (int q = wk.run_lo[r]; q < wk.run_hi[r]; q++) {
                        const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q];
                        const double cq = lc1_slope(pCN[q], h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx_s[q]], rp);
                        const double dm = wn_t - xnu, sp = wn_t + xnu;
                        const double r1 = rcp3(fma(dm, dm, h2));
                        const double r2 = rcp3(fma(sp, sp, h2));
                        sf += fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;          // Voigt-branch pairs: corrected by voigt_kernel
                    }
                }
                tail_t = fma(w, sf, tail_t);
            }
        } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
            if (tid < 128) {
                double sf = 0.;
                if (a.counters) n_direct += (tid == 0) ? (long long)(wk.q1 - wk.q0) * nvalid : 0;
                for (int q = wk.q0; q < wk.q1; q++) {
                    const double xnu = pXNU[q], vt = pVT[q];
                    const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                    const int xf = a.xf_s[q];
                    const double dm = wn_t - xnu;
                    if ((fabs(dm) > kDELTNUC) && (sg.mol != 7)) continue;          // modm.f90:384
                    if (SEL && sg.mol != 7) { cnt_t++; hsh_t += a.key[q]; }
                    const bool voigt = fabs(dm) <= vt;
                    sf += cl.stild * lsf_general(sg.mol, xf, rp, rp2, cl.aip, cl.bip, cl.hw, wn_t, xnu, cl.ad, a.sdep_s[q], voigt, &err);
                }
                tail_t = fma(w, sf, tail_t);
            }
        }
    }
    finish_mol(cur_mol);
    if (!by_mol) {
        reduce_to_smem();
        __syncthreads();
        if (tid < 128) osum_t = s_sum[tid] + tail_t;
    }
    if (SEL) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < C; c++) {
            long long v = cnt[c] + cnt_l;
            unsigned long long h = hsh[c] + hsh_l;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, off); h += __shfl_xor_sync(0xffffffffu, h, off); }
            if (lane == 0) { s_cnt[wid * C + c] = v + cnt_u; s_hsh[wid * C + c] = h + hsh_u; }
        }
        __syncthreads();
    }
    if (err) atomicOr(a.errflag, 2);
    if (a.counters && (tid & 31) == 0 && (wid == 0 || n_direct != 0)) {
        // every warp walked the same streamed sub-ranges: warp 0 reports them; the unstreamed classes were counted by thread 0
        if (wid == 0) atomicAdd(a.counters + 1, (unsigned long long)n_direct);
    }
    if (valid_t) {
        const size_t fl = (size_t)iw_t + (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
        a.o[fl] = osum_t;
        if (SEL) {
            if (a.sel_count) a.sel_count[fl] = s_cnt[tid] + cnt_t;
            if (a.sel_hash) a.sel_hash[fl] = s_hsh[tid] + hsh_t;
        }
    }
}

// =============================================================================================
// vplan_kernel: Voigt-zone candidates of a tile.  voigt_kernel stages, per (tile, layer), every line of the plan's Voigt
// zone [wlo - vb, whi + vb]; on a channel list that zone spans cm-1 (thousands of lines) although a line matters only if
// some frequency of the tile lies within 100*HWHM_D (~2e-4 |Xnu|) of its centre.  This kernel keeps, once per call and
// independent of the layer (the bound uses the segment's largest 100*HWHM_D rate and the largest line shift of the
// call), the zone lines that have a frequency that close: voigt_kernel then stages only those.
// out: cand[tile][kVCandMax] line indices in plan order, segment of each, count per tile (-1: list too long or the
// tile's frequencies are not ascending: voigt_kernel walks the whole zone as before)
// =============================================================================================
constexpr int kVCandMax = 4096;
struct VPlanArgs {
    int32_t nwn, nseg, tile_freqs, pad;
    const double* wn;
    const Segment* seg;
    const double* xnu0;
    const SegWork* plan;                      // level 0
    const unsigned long long* sm_max_bits;
    const unsigned long long* vtmax_seg;
    const int* replan;
    int* cand;                                // [ntiles][kVCandMax]
    unsigned char* cand_seg;                  // [ntiles][kVCandMax]
    int* count;                               // [ntiles]
};

__global__ void __launch_bounds__(128) vplan_kernel(VPlanArgs a)
{
    constexpr int NT = 128;
    if (a.replan && *a.replan == 0) return;
    __shared__ double s_wn[512];
    __shared__ int s_wcount[NT / 32], s_total, s_sorted;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const int i0 = tile * a.tile_freqs;
    const int nf = min(a.tile_freqs, a.nwn - i0);
    if (tid == 0) { s_total = 0; s_sorted = 1; }
    __syncthreads();
    for (int i = tid; i < nf && i < 512; i += NT) s_wn[i] = a.wn[i0 + i];
    __syncthreads();
    for (int i = tid; i + 1 < nf; i += NT) if (s_wn[i + 1] < s_wn[i]) s_sorted = 0;
    __syncthreads();
    if (!s_sorted || nf > 512) { if (tid == 0) a.count[tile] = -1; return; }
    const double sm = __longlong_as_double((long long)*a.sm_max_bits);
    const SegWork* plan = a.plan + (size_t)tile * a.nseg;
    int* cand = a.cand + (size_t)tile * kVCandMax;
    unsigned char* cseg = a.cand_seg + (size_t)tile * kVCandMax;
    for (int s = 0; s < a.nseg; s++) {
        const Segment sg = a.seg[s];
        const int v0 = plan[s].v0, v1 = plan[s].v1;
        if (sg.cls == CLS_GENERAL || v1 <= v0) continue;
        const double rate = __longlong_as_double((long long)a.vtmax_seg[s]);
        for (int b = v0; b < v1; b += NT) {
            const int q = b + tid;
            bool hit = false;
            if (q < v1) {
                const double x0 = a.xnu0[q];
                // |WN - Xnu| <= 100*HWHM_D <= rate*|Xnu| with |Xnu - Xnu0| <= sm
                const double bound = rate * (fabs(x0) + sm) * (1. + 1e-9) + sm + 1e-12;
                int lo = 0, hi = nf;                       // first frequency >= x0
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_wn[mid] < x0) lo = mid + 1; else hi = mid; }
                if (lo < nf && s_wn[lo] - x0 <= bound) hit = true;
                if (lo > 0 && x0 - s_wn[lo - 1] <= bound) hit = true;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[wid] = __popc(m);
            __syncthreads();
            int off = s_total;
            for (int i = 0; i < wid; i++) off += s_wcount[i];
            if (hit) {
                const int pos = off + __popc(m & ((1u << lane) - 1u));
                if (pos < kVCandMax) { cand[pos] = q; cseg[pos] = (unsigned char)s; }
            }
            __syncthreads();
            if (tid == 0) { int t = s_total; for (int i = 0; i < NT / 32; i++) t += s_wcount[i]; s_total = t; }
            __syncthreads();
        }
    }
    if (tid == 0) a.count[tile] = (s_total <= kVCandMax) ? s_total : -1;
}
