// near3.cuh -- plan3_kernel + near3_kernel: the production near-field kernel (one sum over all molecules, no selection
// instrumentation).  Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// near2_kernel re-plans the staged lines of a tile in every warp and for every layer (about a third of its instructions)
// and evaluates directly every line that is near the 128 frequencies a warp owns.  near3_kernel repeats nothing that is
// layer independent and halves the direct work:
//   * plan3_kernel (once per call, layer independent -- margins are maxima over the layers) sorts the staged lines of a
//     512-frequency tile, for each of its eight 64-frequency LEAVES, into index lists: R (far from the leaf and the
//     reference's tests modm.f90:384, 427, 746 cannot come out differently inside it: expanded about the leaf centre),
//     D1 / D2 (near: evaluated per frequency, single / both resonances) and C (CANDIDATES: the outcome of a test, or near
//     versus far, depends on the layer's shifted centre).  The Voigt test (:427) plays no role here: every pair is
//     evaluated with the Lorentz form and voigt_kernel replaces it by the Voigt form where the reference takes that branch
//   * near3_kernel: CTA = (tile, group of layers), 8 warps = 8 leaves, each lane owns two frequencies.  Per layer four
//     planes of the staged lines (XNU, H2, CN, P3) arrive by TMA (the next layer's copy runs under the current layer's
//     arithmetic) and are regrouped into one 32-byte record per line with the column amount folded in; a warp decides its few candidates exactly (the floating-point differences are
//     monotone in the frequency), expands the ring (14 Taylor terms, one line per lane, coefficients reduced across the warp
//     by a transposing butterfly: 16 adds instead of 70) and evaluates the direct lists (four lines share one reciprocal)
// The selected line set is unchanged: a line is expanded only where the reference's window test passes for every frequency
// of the leaf in every layer; every other pair runs the literal test.
// =============================================================================================
constexpr int kN3Leaves = 8;                      // 64-frequency leaves per tile (one warp each)
constexpr int kN3Lists = 4 * kN3Leaves;           // per leaf: R, D1, D2, C
constexpr int kN3Pool = 8 * kNearCap + 4 * kN3Lists + 32 * kN3Leaves;   // list entries (u16) per tile: every staged line is in at most one list per leaf
constexpr int kN3CMax = 96;                       // candidates per leaf (more: the tile goes to near_kernel)
constexpr int kN3Runs = 48;                       // staged runs per tile
constexpr int kN3Tail = 16;                       // segments of the unstaged classes (first-order O2 mixing, general case tree)
constexpr int kN3Idx = 0x3ff;                     // entry: staged index
constexpr int kN3Both = 0x8000;                   // R entry: both resonances for every frequency of the leaf
// C entry: index | mode (EDGE, NEG) << 10 | both resonances everywhere in the tile << 13 | kind << 14
// T entry: index | kind << 10
static_assert(kNearCap + 8 <= kN3Idx, "staged index must fit the list entry");

struct Run3 { int qs, off, n4; };                 // TMA copy: n4 lines (multiple of four) from plane index qs to staged offset off
struct Tail3 { int seg, lo, hi; };                // lines [lo, hi) of an unstaged segment that are direct for this tile
struct Tile3 {
    double wA[kN3Leaves], wB[kN3Leaves];          // frequency extent of the leaf (wB < wA: empty)
    unsigned short off[kN3Lists], cnt[kN3Lists];  // list l = pool[off[l] .. off[l] + cnt[l]); starts are multiples of four
    int npool, nrun, ntail, ok;                   // ok = 1: near3_kernel handles the tile
    Run3 run[kN3Runs];
    Tail3 tail[kN3Tail];
};

struct Plan3Args {
    int32_t nwn, nseg;
    const double* wn;
    const Segment* seg;
    const double* xnu0;
    const double* deltnu;                         // static pressure-shift coefficients (per-line shift margin)
    const int32_t* brdidx;                        // >= 0: the line carries species-broadening shifts (full margin)
    double max_abs_deltnu;
    const unsigned long long* sm_max_bits;
    const unsigned long long* vtmax_seg;
    TileHdr* hdr;                                 // level 0 (nnear is set to -1 when the lists do not fit)
    const SegWork* plan;                          // level 0
    const NearPiece* near_pieces;
    double ffw_ratio;
    Tile3* out;
    unsigned short* pool;                         // [ntiles][kN3Pool]
    unsigned char* segof;                         // [ntiles][kNearCap] segment of every staged line (0xff: padding)
    const int* replan;                            // plan cache: 0 = keep the lists of the previous call
};

__global__ void __launch_bounds__(128) plan3_kernel(Plan3Args a)
{
    constexpr int NT = 128;
    if (a.replan && *a.replan == 0) return;
    __shared__ NearPiece s_np[kMaxNearPieces];
    __shared__ unsigned char s_code[kNearCap][kN3Leaves];   // 0 none, 1 ring, 2 ring both, 3 D1, 4 D2, 5 candidate
    __shared__ unsigned short s_cent[kNearCap];             // candidate entry bits of the line (mode, both, kind)
    __shared__ double s_wA[kN3Leaves], s_wB[kN3Leaves];
    __shared__ int s_cnt[kN3Lists], s_off[kN3Lists];
    __shared__ int s_ok;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const TileHdr th = a.hdr[tile];
    Tile3* out = a.out + tile;
    if (th.total_lines > kNearCap || th.nnear < 0) {
        if (tid == 0) out->ok = 0;
        return;
    }
    const int npc = th.nnear, total = th.total_lines;
    // leaf extents (the frequencies need not be sorted)
    for (int l = 2 * wid; l < 2 * wid + 2; l++) {
        double wA = 1e300, wB = -1e300;
        for (int f = 0; f < 2; f++) {
            const int iw = tile * 512 + 64 * l + 32 * f + lane;
            if (iw < a.nwn) {
                const double w = a.wn[iw];
                wA = fmin(wA, w);
                wB = fmax(wB, w);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            wA = fmin(wA, __shfl_xor_sync(0xffffffffu, wA, off));
            wB = fmax(wB, __shfl_xor_sync(0xffffffffu, wB, off));
        }
        if (lane == 0) { s_wA[l] = wA; s_wB[l] = wB; }
    }
    {
        const int np4 = npc * (int)(sizeof(NearPiece) / 4);
        const int* psrc = reinterpret_cast<const int*>(a.near_pieces + (size_t)tile * kMaxNearPieces);
        int* pdst = reinterpret_cast<int*>(s_np);
        for (int i = tid; i < np4; i += NT) pdst[i] = psrc[i];
        for (int i = tid; i < total * kN3Leaves; i += NT) (&s_code[0][0])[i] = 0;
        unsigned char* so = a.segof + (size_t)tile * kNearCap;
        for (int i = tid; i < kNearCap; i += NT) so[i] = 0xff;
    }
    __syncthreads();
    const double sm_all = __longlong_as_double((long long)*a.sm_max_bits);
    const double eps = 1e-9;                                   // slack of the "cannot come out differently" decisions
    for (int p = wid; p < npc; p += NT / 32) {
        const NearPiece pc = s_np[p];
        const int s = pc.info & 0xff, mode = (pc.info >> 8) & 0xff, kind = (pc.info >> 17) & 3;
        const bool negall = ((pc.info >> 16) & 1) != 0, has_win = kind != 2;
        unsigned char* so = a.segof + (size_t)tile * kNearCap;
        for (int i = lane; i < pc.n; i += 32) {
            const int j = pc.soff + i;
            const double x0 = a.xnu0[pc.q0 + i];
            // |Xnu - Xnu0| <= |deltnu| * max(Xn/XN0) for a line without species-broadening shifts (modm.f90:375); the call's
            // margin sm_all is that bound for the largest |deltnu| of the list
            double sm = sm_all;
            if (a.brdidx[pc.q0 + i] < 0 && a.max_abs_deltnu > 0.)
                sm = fmin(sm_all, sm_all * (fabs(a.deltnu[pc.q0 + i]) / a.max_abs_deltnu) * (1. + 1e-9) + 1e-12);
            so[j] = (unsigned char)s;
            s_cent[j] = (unsigned short)(((mode & 7) << 10) | (negall ? (1 << 13) : 0) | (kind << 14));
            for (int l = 0; l < kN3Leaves; l++) {
                const double gA = s_wA[l], gB = s_wB[l];
                if (gB < gA) continue;
                bool unc = false, out_ = false;
                if (has_win && (mode & M_EDGE)) {
                    if ((gA - x0 - sm > kDELTNUC + eps) || (gB - x0 + sm < -kDELTNUC - eps)) out_ = true;
                    else if (!((fabs(gA - x0) + sm <= kDELTNUC - eps) && (fabs(gB - x0) + sm <= kDELTNUC - eps))) unc = true;
                }
                if (out_) continue;                            // outside the window for every frequency of the leaf in every layer
                bool both = negall;
                if (!negall && (mode & M_NEG)) {
                    if (gB + x0 + sm <= kDELTNUC - eps) both = true;
                    else if (gA + x0 - sm > kDELTNUC + eps) both = false;
                    else unc = true;
                }
                int code;
                if (unc) {
                    code = 5;
                } else {
                    // far for every layer / near for every layer / in between: the layer's shifted centre decides
                    const double cen = 0.5 * (gA + gB), R = a.ffw_ratio * 0.5 * (gB - gA);
                    const double dm = fabs(cen - x0), dp = fabs(cen + x0);
                    const bool far = (dm - sm >= R) && (!both || (dp - sm >= R));
                    const bool near_ = (dm + sm < R) || (both && (dp + sm < R));
                    code = far ? (both ? 2 : 1) : (near_ ? (both ? 4 : 3) : 5);
                }
                s_code[j][l] = (unsigned char)code;
            }
        }
    }
    __syncthreads();
    // list l = 4*leaf + {R, D1, D2, C} in staged order (deterministic): one thread per list counts, then writes
    auto member = [&](int l, int j, int& ent) -> bool {
        const int leaf = l >> 2, k = l & 3;
        const int c = s_code[j][leaf];
        if (k == 0) { ent = j | (c == 2 ? kN3Both : 0); return c == 1 || c == 2; }
        if (k == 3) { ent = j | s_cent[j]; return c == 5; }
        ent = j;
        return (k == 1) ? (c == 3) : (c == 4);
    };
    if (tid < kN3Lists) {
        int n = 0, ent;
        for (int j = 0; j < total; j++) n += member(tid, j, ent) ? 1 : 0;
        s_cnt[tid] = n;
    }
    __syncthreads();
    if (tid == 0) {
        int tot = 0, ok = 1;
        for (int l = 0; l < kN3Lists; l++) {
            const int c = s_cnt[l];
            if ((l & 3) == 3 && c > kN3CMax) ok = 0;
            s_off[l] = tot;
            tot += ((l & 3) == 0) ? ((c + 31) & ~31) : ((c + 3) & ~3);
        }
        if (tot > kN3Pool) ok = 0;
        // the TMA copies of the staged runs and the direct ranges of the unstaged classes
        int nrun = 0, ntail = 0;
        const SegWork* plan = a.plan + (size_t)tile * a.nseg;
        for (int s = 0; s < a.nseg && ok; s++) {
            const SegWork& w = plan[s];
            const int cls = a.seg[s].cls;
            if (w.tma) {
                for (int r = 0; r < w.nrun; r++) {
                    if (nrun >= kN3Runs) { ok = 0; break; }
                    out->run[nrun].qs = w.run_t0[r];
                    out->run[nrun].off = w.run_off[r];
                    out->run[nrun].n4 = ((w.run_hi[r] - w.run_t0[r]) + 3) & ~3;
                    nrun++;
                }
            } else if (cls == CLS_O2_LC1) {
                for (int r = 0; r < w.nrun; r++) {
                    if (ntail >= kN3Tail) { ok = 0; break; }
                    out->tail[ntail].seg = s; out->tail[ntail].lo = w.run_lo[r]; out->tail[ntail].hi = w.run_hi[r];
                    ntail++;
                }
            } else if (cls == CLS_GENERAL && w.q1 > w.q0) {
                if (ntail >= kN3Tail) { ok = 0; break; }
                out->tail[ntail].seg = s; out->tail[ntail].lo = w.q0; out->tail[ntail].hi = w.q1;
                ntail++;
            }
        }
        out->npool = tot;
        out->nrun = nrun;
        out->ntail = ntail;
        out->ok = ok;
        if (!ok) a.hdr[tile].nnear = -1;                       // near_kernel streams this tile
        s_ok = ok;
        for (int l = 0; l < kN3Leaves; l++) { out->wA[l] = s_wA[l]; out->wB[l] = s_wB[l]; }
    }
    __syncthreads();
    if (!s_ok) return;
    if (tid < kN3Lists) {
        out->off[tid] = (unsigned short)s_off[tid];
        out->cnt[tid] = (unsigned short)s_cnt[tid];
        unsigned short* dst = a.pool + (size_t)tile * kN3Pool + s_off[tid];
        int n = 0, ent;
        for (int j = 0; j < total; j++)
            if (member(tid, j, ent)) dst[n++] = (unsigned short)ent;
        const int end = ((tid & 3) == 0) ? ((n + 31) & ~31) : ((n + 3) & ~3);   // ring lists are read a warp at a time
        for (; n < end; n++) dst[n] = (unsigned short)kNearCap;        // the neutral staged slot (zero strength) pads the groups
    }
}

// transposing butterfly: every lane contributes v[0..15]; afterwards lane l holds the warp-wide sum of v[(l >> 1) & 15]
// (16 adds and 16 64-bit shuffles instead of 16 x 5)
__device__ __forceinline__ double warp_transpose_sum16(double (&v)[16], int lane)
{
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double send = hi ? v[i] : v[8 + i];
            const double keep = hi ? v[8 + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double send = hi ? v[i] : v[4 + i];
            const double keep = hi ? v[4 + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double send = hi ? v[i] : v[2 + i];
            const double keep = hi ? v[2 + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool hi = (lane & 2) != 0;
        const double send = hi ? v[0] : v[1];
        const double keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

#ifndef MRTM_NEAR3_MINB
#define MRTM_NEAR3_MINB 3
#endif
static_assert(kFarK <= 15, "near3_kernel packs the coefficients and the pedestal sum into 16 slots");

struct Near3Args {
    const Tile3* t3;
    const unsigned short* pool;
    const unsigned char* segof;
    int32_t lb, pad;                  // layers per CTA
};
constexpr int kN3Dyn = 128;                       // dynamic list capacity: kN3CMax entries + padding to whole warps
constexpr size_t kN3PlaneLen = kNearCap + 8;
// TMA buffer (4 planes, structure of arrays), work buffer (4 doubles per line, weights applied), lists
constexpr size_t kN3Smem = sizeof(double) * (4 + 4) * kN3PlaneLen + sizeof(unsigned short) * (kN3Pool + kN3Leaves * 5 * kN3Dyn) + kNearCap;

__global__ void __launch_bounds__(256, MRTM_NEAR3_MINB) near3_kernel(LinesArgs a, Near3Args n3)
{
    constexpr int NT = 256, NW = kN3Leaves;
    constexpr int kCap = kNearCap, kDummy = kCap, kPlane = (int)kN3PlaneLen;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const TileHdr th = a.hdr[0][blockIdx.x];
    if (th.total_lines > kCap || th.nnear < 0) return;       // near_kernel streams this tile
    const int prof = blockIdx.z;
    const int k_beg = blockIdx.y * n3.lb, k_end = min(k_beg + n3.lb, a.nlay);

    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_wseg[kMaxSegments];
    __shared__ Tile3 s_t3;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    double* s_tma = reinterpret_cast<double*>(s_dyn);                                          // [4][kPlane] as the planes lie in HBM
    double4* s_ln = reinterpret_cast<double4*>(s_tma + 4 * kPlane);                            // [kPlane] {XNU, H2, W*CN, W*P3}
    unsigned short* s_pool = reinterpret_cast<unsigned short*>(s_ln + kPlane);                 // [kN3Pool]
    unsigned short* s_dynl = s_pool + kN3Pool;                                                 // [NW][5][kN3Dyn]
    unsigned char* s_seg = reinterpret_cast<unsigned char*>(s_dynl + NW * 5 * kN3Dyn);         // [kCap]

    const int total = th.total_lines;
    {
        const int* tsrc = reinterpret_cast<const int*>(n3.t3 + blockIdx.x);
        int* tdst = reinterpret_cast<int*>(&s_t3);
        for (int i = tid; i < (int)(sizeof(Tile3) / 4); i += NT) tdst[i] = tsrc[i];
        const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(n3.segof + (size_t)blockIdx.x * kCap);
        for (int i = tid; i < kCap / 4; i += NT) reinterpret_cast<uint32_t*>(s_seg)[i] = ssrc[i];
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            s_ln[kDummy] = make_double4(th.wlo, 1.0, 0., 0.);     // the neutral slot that pads the lists
        }
    }
    __syncthreads();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(n3.pool + (size_t)blockIdx.x * kN3Pool);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_pool);
        for (int i = tid; i < (s_t3.npool + 1) / 2; i += NT) dst[i] = src[i];
    }
    // one elected thread issues the copies of a layer: four planes of every staged run
    auto issue = [&](int k) {
        const int64_t L = (int64_t)prof * a.nlay + k;
        const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
        uint32_t bytes = 0;
        for (int r = 0; r < s_t3.nrun; r++) bytes += (uint32_t)s_t3.run[r].n4 * 32u;
        mbar_expect_tx(&s_bar, bytes);
        for (int r = 0; r < s_t3.nrun; r++) {
            const Run3 rn = s_t3.run[r];
            const uint32_t b = (uint32_t)rn.n4 * 8u;
            tma_load_1d(s_tma + 0 * kPlane + rn.off, pl + (size_t)D_XNU * a.n_pad + rn.qs, b, &s_bar);
            tma_load_1d(s_tma + 1 * kPlane + rn.off, pl + (size_t)D_H2 * a.n_pad + rn.qs, b, &s_bar);
            tma_load_1d(s_tma + 2 * kPlane + rn.off, pl + (size_t)D_CN * a.n_pad + rn.qs, b, &s_bar);
            tma_load_1d(s_tma + 3 * kPlane + rn.off, pl + (size_t)D_P3 * a.n_pad + rn.qs, b, &s_bar);
        }
    };
    if (tid == 0 && k_beg < k_end) issue(k_beg);

    // this warp's leaf: lane owns frequencies base + lane and base + 32 + lane
    const int base = blockIdx.x * 512 + wid * 64;
    const int iw0 = base + lane, iw1 = base + 32 + lane;
    const bool valid0 = iw0 < a.nwn, valid1 = iw1 < a.nwn;
    const double wn0 = a.wn[valid0 ? iw0 : (a.nwn - 1)], wn1 = a.wn[valid1 ? iw1 : (a.nwn - 1)];
    const double wA = s_t3.wA[wid], wB = s_t3.wB[wid];
    const bool leaf_empty = wB < wA;
    const double cen = 0.5 * (wA + wB), hh = 0.5 * (wB - wA);
    const double hinv = hh > 0. ? 1. / hh : 0.;
    const double Rn = a.ffw_ratio * hh, R2 = Rn * Rn;
    const double m2h = -2. * hh, mhh = -hh * hh;
    const double sv0 = (wn0 - cen) * hinv, sv1 = (wn1 - cen) * hinv;
    const unsigned short* lR = s_pool + s_t3.off[4 * wid];
    const unsigned short* lD1 = s_pool + s_t3.off[4 * wid + 1];
    const unsigned short* lD2 = s_pool + s_t3.off[4 * wid + 2];
    const unsigned short* lC = s_pool + s_t3.off[4 * wid + 3];
    const int nR = s_t3.cnt[4 * wid], nD1 = (s_t3.cnt[4 * wid + 1] + 3) & ~3, nD2 = s_t3.cnt[4 * wid + 2], nC = s_t3.cnt[4 * wid + 3];
    unsigned short* dR = s_dynl + (size_t)wid * 5 * kN3Dyn;     // dynamic lists of the current layer
    unsigned short* dD1 = dR + kN3Dyn;
    unsigned short* dD2 = dD1 + kN3Dyn;
    unsigned short* dT1 = dD2 + kN3Dyn;
    unsigned short* dT2 = dT1 + kN3Dyn;
    const unsigned lt_mask = (1u << lane) - 1u;
    int nvalid = 0;
    if (a.counters) {
        const int rem = a.nwn - base;
        nvalid = rem < 0 ? 0 : (rem < 64 ? rem : 64);
    }

    for (int k = k_beg; k < k_end; k++) {
        const int it = k - k_beg;
        const int64_t L = (int64_t)prof * a.nlay + k;
        const LayerDev& ly = a.lay[L];
        for (int s = tid; s < a.nseg; s += NT) s_wseg[s] = ly.wk[a.seg[s].mol - 1];
        mbar_wait(&s_bar, (uint32_t)it & 1u);
        __syncthreads();                  // the previous layer is finished everywhere (its work buffer may be overwritten)
        // planes -> one 32-byte record per line, with the column amount W folded in (o = RFT*sum_mol W_mol*SF_mol); a molecule
        // with zero amount contributes nothing (modm.f90:318-321)
        for (int j = tid; j < total; j += NT) {
            const int sg = s_seg[j];
            const double w = (sg == 0xff) ? 0. : s_wseg[sg];
            double4 rec = make_double4(s_tma[j], s_tma[kPlane + j], s_tma[2 * kPlane + j] * w, s_tma[3 * kPlane + j] * w);
            if (w == 0.) rec = make_double4(rec.x, 1.0, 0., 0.);
            s_ln[j] = rec;
        }
        __syncthreads();
        // the next layer's lines travel while this layer is evaluated
        if (tid == 0 && k + 1 < k_end) issue(k + 1);

        double lsum0 = 0., lsum1 = 0.;
        int err = 0;
        long long n_direct = 0, n_far = 0;
        if (!leaf_empty) {
            double A[16];
#pragma unroll
            for (int i = 0; i < 16; i++) A[i] = 0.;
            double (&A14)[kFarK] = *reinterpret_cast<double (*)[kFarK]>(&A[0]);
            // -- candidates: the reference's tests decided exactly for this layer and this leaf
            int ndR = 0, ndD1 = 0, ndD2 = 0, nT1 = 0, nT2 = 0;
            for (int e0 = 0; e0 < nC; e0 += 32) {
                const int e = e0 + lane;
                const bool ok = e < nC;
                const int ent = ok ? (int)lC[e] : kDummy;
                const int j = ent & kN3Idx;
                const int mode = (ent >> 10) & 3, kind = (ent >> 14) & 3;
                bool both = ((ent >> 13) & 1) != 0;
                const double4 ln = s_ln[j];
                const double x = ln.x, h2 = ln.y, pd = ln.w;
                const double dA = wA - x, dB = wB - x;
                bool skip = !ok, test = false, neg_possible = both;
                if (ok && (mode & M_EDGE)) {
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
                const bool plain = ok && !skip && !test;
                const double Dm = cen - x, Dp = cen + x;
                const bool far = plain && (fma(Dm, Dm, h2) >= R2) && (!both || (fma(Dp, Dp, h2) >= R2));
                if (plain) A[15] += both ? 2. * pd : pd;
                const bool d1 = plain && !far && !both, d2 = plain && !far && both;
                const bool t1 = ok && !skip && test && !neg_possible, t2 = ok && !skip && test && neg_possible;
                const unsigned m0 = __ballot_sync(0xffffffffu, far), m1 = __ballot_sync(0xffffffffu, d1);
                const unsigned m2 = __ballot_sync(0xffffffffu, d2), m3 = __ballot_sync(0xffffffffu, t1);
                const unsigned m4 = __ballot_sync(0xffffffffu, t2);
                const unsigned short tent = (unsigned short)(j | (kind << 10));
                if (far) dR[ndR + __popc(m0 & lt_mask)] = (unsigned short)(j | (both ? kN3Both : 0));
                if (d1) dD1[ndD1 + __popc(m1 & lt_mask)] = (unsigned short)j;
                if (d2) dD2[ndD2 + __popc(m2 & lt_mask)] = (unsigned short)j;
                if (t1) dT1[nT1 + __popc(m3 & lt_mask)] = tent;
                if (t2) dT2[nT2 + __popc(m4 & lt_mask)] = tent;
                ndR += __popc(m0); ndD1 += __popc(m1); ndD2 += __popc(m2); nT1 += __popc(m3); nT2 += __popc(m4);
            }
            if (lane < 3) dD1[ndD1 + lane] = (unsigned short)kDummy;       // pad to a group of four
            dR[ndR + lane] = (unsigned short)kDummy;                       // pad the dynamic ring to whole warps
            __syncwarp();
            // -- the ring: static list, then the candidates that turned out plain and far
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const unsigned short* lst = pass ? dR : lR;
                const int nl = pass ? ndR : nR;
                const double pw = pass ? 0. : 1.;               // the candidates' pedestals are already counted
#pragma unroll 1
                for (int e0 = 0; e0 < nl; e0 += 32) {
                    const int ent = lst[e0 + lane];                 // padded with the neutral slot: no bounds test
                    const int j = ent & kN3Idx;
                    const bool both = (ent & kN3Both) != 0;
                    const double4 ln = s_ln[j];
                    A[15] = fma(both ? 2. * pw : pw, ln.w, A[15]);
                    far_accum(cen - ln.x, ln.y, ln.z, 0., m2h, mhh, A14);
                    const unsigned mboth = __ballot_sync(0xffffffffu, both);
                    if (mboth) far_accum(cen + ln.x, ln.y, both ? ln.z : 0., 0., m2h, mhh, A14);
                    if (a.counters) n_far += __popc(__ballot_sync(0xffffffffu, j != kDummy)) + __popc(mboth);
                }
            }
            // static direct lists: inside the window for every frequency of the leaf -> their pedestals once per line
            for (int e = lane; e < nD1; e += 32) A[15] += s_ln[lD1[e]].w;
            for (int e = lane; e < nD2; e += 32) A[15] += 2. * s_ln[lD2[e]].w;
            {
                const double mine = warp_transpose_sum16(A, lane);
                const double ped = __shfl_sync(0xffffffffu, mine, 30);
                lsum0 = lsum1 = __shfl_sync(0xffffffffu, mine, 2 * (kFarK - 1));
#pragma unroll
                for (int i = kFarK - 2; i >= 0; i--) {
                    const double ci = __shfl_sync(0xffffffffu, mine, 2 * i);
                    lsum0 = fma(lsum0, sv0, ci);
                    lsum1 = fma(lsum1, sv1, ci);
                }
                lsum0 -= ped;
                lsum1 -= ped;
            }
            if (a.counters) n_direct += (long long)(s_t3.cnt[4 * wid + 1] + nD2 + ndD1 + ndD2 + nT1 + nT2) * nvalid;
            // -- D1: single resonance (modm.f90:751); four lines share one reciprocal,
            // sum c_i/a_i = N/(a1 a2 a3 a4), N = (c1 a2 + c2 a1)(a3 a4) + (c3 a4 + c4 a3)(a1 a2)
            double p0 = 0., p1 = 0.;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const unsigned short* lst = pass ? dD1 : lD1;
                const int nl = pass ? ndD1 : nD1;
#pragma unroll 1
                for (int g = 0; g < nl; g += 4) {
                    const uint2 iv = *reinterpret_cast<const uint2*>(lst + g);
                    const int i1 = iv.x & 0xffff, i2 = iv.x >> 16, i3 = iv.y & 0xffff, i4 = iv.y >> 16;
                    const double2 xg1 = *reinterpret_cast<const double2*>(&s_ln[i1]), xg2 = *reinterpret_cast<const double2*>(&s_ln[i2]);
                    const double2 xg3 = *reinterpret_cast<const double2*>(&s_ln[i3]), xg4 = *reinterpret_cast<const double2*>(&s_ln[i4]);
                    const double x1 = xg1.x, x2 = xg2.x, x3 = xg3.x, x4 = xg4.x;
                    const double g1 = xg1.y, g2 = xg2.y, g3 = xg3.y, g4 = xg4.y;
                    const double c1 = s_ln[i1].z, c2 = s_ln[i2].z, c3 = s_ln[i3].z, c4 = s_ln[i4].z;
                    {
                        const double d1v = wn0 - x1, d2v = wn0 - x2, d3v = wn0 - x3, d4v = wn0 - x4;
                        const double a1 = fma(d1v, d1v, g1), a2 = fma(d2v, d2v, g2);
                        const double a3 = fma(d3v, d3v, g3), a4 = fma(d4v, d4v, g4);
                        const double p12 = a1 * a2, p34 = a3 * a4;
                        const double n12 = fma(c1, a2, c2 * a1), n34 = fma(c3, a4, c4 * a3);
                        p0 = fma(fma(n12, p34, n34 * p12), rcp3(p12 * p34), p0);
                    }
                    {
                        const double d1v = wn1 - x1, d2v = wn1 - x2, d3v = wn1 - x3, d4v = wn1 - x4;
                        const double a1 = fma(d1v, d1v, g1), a2 = fma(d2v, d2v, g2);
                        const double a3 = fma(d3v, d3v, g3), a4 = fma(d4v, d4v, g4);
                        const double p12 = a1 * a2, p34 = a3 * a4;
                        const double n12 = fma(c1, a2, c2 * a1), n34 = fma(c3, a4, c4 * a3);
                        p1 = fma(fma(n12, p34, n34 * p12), rcp3(p12 * p34), p1);
                    }
                }
            }
            // -- D2: both resonances: cn*(1/a+1/b) = cn*(a+b)/(a*b), one reciprocal
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const unsigned short* lst = pass ? dD2 : lD2;
                const int nl = pass ? ndD2 : nD2;
#pragma unroll 2
                for (int g = 0; g < nl; g++) {
                    const int j = lst[g];
                    const double4 ln = s_ln[j];
                    const double xnu = ln.x, h2 = ln.y, cn = ln.z;
                    {
                        const double dm = wn0 - xnu, sp = wn0 + xnu;
                        const double aa = fma(dm, dm, h2), bb = fma(sp, sp, h2);
                        p0 = fma(cn * (aa + bb), rcp3(aa * bb), p0);
                    }
                    {
                        const double dm = wn1 - xnu, sp = wn1 + xnu;
                        const double aa = fma(dm, dm, h2), bb = fma(sp, sp, h2);
                        p1 = fma(cn * (aa + bb), rcp3(aa * bb), p1);
                    }
                }
            }
            lsum0 += p0;
            lsum1 += p1;
            // -- T1 / T2: the reference's exact per-(line,frequency) tests (modm.f90:384, 427, 746)
#pragma unroll 1
            for (int g = 0; g < nT1; g++) {
                const int ent = dT1[g], j = ent & kN3Idx;
                const bool has_win = ((ent >> 10) & 3) != 2;
                const double4 ln = s_ln[j];
                const double xnu = ln.x, h2 = ln.y, cn = ln.z, ped = ln.w;
                const double dm0 = wn0 - xnu, dm1 = wn1 - xnu;
                const bool take0 = has_win ? !(fabs(dm0) > kDELTNUC) : true;
                const bool take1 = has_win ? !(fabs(dm1) > kDELTNUC) : true;
                const double v0 = fma(cn, rcp3(fma(dm0, dm0, h2)), -ped), v1 = fma(cn, rcp3(fma(dm1, dm1, h2)), -ped);
                lsum0 += take0 ? v0 : 0.;
                lsum1 += take1 ? v1 : 0.;
            }
#pragma unroll 1
            for (int g = 0; g < nT2; g++) {
                const int ent = dT2[g], j = ent & kN3Idx;
                const int kind = (ent >> 10) & 3;
                const bool has_win = kind != 2, negall = kind == 2;
                const double4 ln = s_ln[j];
                const double xnu = ln.x, h2 = ln.y, cn = ln.z, ped = ln.w;
#pragma unroll
                for (int ff = 0; ff < 2; ff++) {
                    const double w = ff ? wn1 : wn0;
                    const double dm = w - xnu, sp = w + xnu;
                    const bool take = has_win ? !(fabs(dm) > kDELTNUC) : true;
                    const bool neg = negall || (sp <= kDELTNUC);
                    const double r1 = rcp3(fma(dm, dm, h2)), r2 = rcp3(fma(sp, sp, h2));
                    const double val = cn * (r1 + (neg ? r2 : 0.)) - (neg ? 2. * ped : ped);
                    if (ff) lsum1 += take ? val : 0.; else lsum0 += take ? val : 0.;
                }
            }
            // ---- classes that are not staged: first-order O2 mixing (few lines) and the general case tree
            if (s_t3.ntail > 0) {
                const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
                const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
                const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
                const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
                const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
                const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
                const double* __restrict__ pVT = pl + (size_t)D_VT * a.n_pad;
                const double rp = ly.rp, rp2 = ly.rp2;
#pragma unroll 1
                for (int ti = 0; ti < s_t3.ntail; ti++) {
                    const Tail3 tl = s_t3.tail[ti];
                    const double w = s_wseg[tl.seg];
                    if (w == 0.) continue;
                    const Segment sg = a.seg[tl.seg];
                    double ss0 = 0., ss1 = 0.;
                    if (a.counters) n_direct += (long long)(tl.hi - tl.lo) * nvalid;
                    if (sg.cls == CLS_O2_LC1) {
                        for (int q = tl.lo; q < tl.hi; q++) {
                            const double xnu = pXNU[q], h2 = pH2[q], cg = pP3[q];
                            const double cq = lc1_slope(pCN[q], h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx_s[q]], rp);
#pragma unroll
                            for (int ff = 0; ff < 2; ff++) {
                                const double wv = ff ? wn1 : wn0;
                                const double dm = wv - xnu, sp = wv + xnu;
                                const double r1 = rcp3(fma(dm, dm, h2));
                                const double r2 = rcp3(fma(sp, sp, h2));
                                const double val = fma(cq, dm, cg) * r1 + fma(-cq, sp, cg) * r2;     // Voigt-branch pairs: corrected by voigt_kernel
                                if (ff) ss1 += val; else ss0 += val;
                            }
                        }
                    } else {   // CLS_GENERAL: faithful case tree per (line, frequency)
                        for (int q = tl.lo; q < tl.hi; q++) {
                            const double xnu = pXNU[q], vt = pVT[q];
                            const ColdLine cl = cold_line(pl, a.n_pad, q, a.lcidx_s, lcp, a.nlc_pad);
                            const double hw = cl.hw, ad = cl.ad, st = cl.stild, aip = cl.aip, bip = cl.bip;
                            const int xf = a.xf_s[q];
#pragma unroll 1
                            for (int ff = 0; ff < 2; ff++) {
                                const double wv = ff ? wn1 : wn0;
                                const double dm = wv - xnu;
                                if ((fabs(dm) > kDELTNUC) && (sg.mol != 7)) continue;          // modm.f90:384
                                const bool voigt = fabs(dm) <= vt;
                                const double add = st * lsf_general(sg.mol, xf, rp, rp2, aip, bip, hw, wv, xnu, ad, a.sdep_s[q], voigt, &err);
                                if (ff) ss1 += add; else ss0 += add;
                            }
                        }
                    }
                    lsum0 = fma(w, ss0, lsum0);
                    lsum1 = fma(w, ss1, lsum1);
                }
            }
            if (err) atomicOr(a.errflag, 2);
            if (a.counters && lane == 0) {
                atomicAdd(a.counters + 0, (unsigned long long)n_far);
                atomicAdd(a.counters + 1, (unsigned long long)n_direct);
            }
            const size_t ob = (size_t)k * a.o_lds + (size_t)prof * a.o_prof;
            if (valid0) a.o[ob + iw0] = lsum0;
            if (valid1) a.o[ob + iw1] = lsum1;
        }
    }
}
