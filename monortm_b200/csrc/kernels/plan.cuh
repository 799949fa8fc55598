// plan.cuh -- plan_kernel: layer-independent classification of every (frequency tile, segment).
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
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
    const int* replan;                        // plan cache: 0 = the plans of the previous call are still valid, leave them
};

// ---- plan cache ------------------------------------------------------------------------------------------------
// The plans depend on the frequencies, the staged centres and two kinds of margins (largest line shift, largest Voigt-zone
// rate per segment, both maxima over the layers of the call).  A repeated call on the same frequency list (the steps of a
// sweep, the batches of an ensemble) reuses them: wn_hash_kernel hashes the list, plan_check_kernel compares the hash and
// checks that the margins the cached plans were built with still cover this call's; the plan kernels return at once when
// they do.  Everything stays on the device: no host synchronisation.
__global__ void wn_hash_kernel(const double* __restrict__ wn, int n, unsigned long long* acc)
{
    unsigned long long h = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)__double_as_longlong(wn[i]) + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        h += z ^ (z >> 31);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) h += __shfl_xor_sync(0xffffffffu, h, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, h);
}
struct PlanCheckArgs {
    const unsigned long long* need;           // [1 + nseg] this call's margins (bits of non-negative doubles): shift, Voigt rates
    unsigned long long* have;                 // [1 + nseg] the margins the cached plans were built with (the plan kernels read these)
    unsigned long long* hash;                 // [2] this call's frequency hash (cleared here for the next call), the cached one
    int nseg, force;
    int* replan;
};
__global__ void plan_check_kernel(PlanCheckArgs a)
{
    __shared__ int s_re;
    if (threadIdx.x == 0) s_re = (a.force || a.hash[0] != a.hash[1]) ? 1 : 0;
    __syncthreads();
    for (int i = threadIdx.x; i <= a.nseg; i += blockDim.x)
        if (a.need[i] > a.have[i]) s_re = 1;              // non-negative doubles order like their bit patterns
    __syncthreads();
    if (s_re) {
        // rebuild with some head room so that the neighbouring profiles of an ensemble still fit
        for (int i = threadIdx.x; i <= a.nseg; i += blockDim.x) {
            const double v = __longlong_as_double((long long)a.need[i]);
            a.have[i] = (unsigned long long)__double_as_longlong(v * (i == 0 ? 1.25 : 1.08));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *a.replan = s_re;
        a.hash[1] = a.hash[0];
        a.hash[0] = 0ull;
    }
}

__global__ void __launch_bounds__(128) plan_kernel(PlanArgs a)
{
    constexpr int NT = 128;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    SegWork* s_work = reinterpret_cast<SegWork*>(s_dyn);
    __shared__ double s_lo[4], s_hi[4];
    if (a.replan && *a.replan == 0) return;           // cached plans are valid
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
