// far.cuh -- far-field expansion: far_accum*, far_kernel, far_warp_kernel.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// ---------------------------------------------------------------------------------------------
// far field
// One far-field term: w/((D+t)^2+h2) (+ optional pedestal) expanded in s = t/h about the tile centre,
//   sum_k b_k s^k,  b_0 = w*u, b_1 = al*b_0, b_k = al*b_{k-1} + be*b_{k-2},  u = 1/(D^2+h2), al = -2*D*h*u, be = -h^2*u.
// The poles of the term sit at distance sqrt(D^2+h2) >= ratio*h from the centre, so the series converges like ratio^-k.
__device__ __forceinline__ void far_accum(double D, double h2, double w, double ped, double m2h, double mhh, double (&A)[kFarK])
{
    const double u = rcp3(fma(D, D, h2));
    const double al = (D * m2h) * u, be = mhh * u;
    double b0 = w * u;
    double b1 = al * b0;
    A[0] += b0 - ped;
    A[1] += b1;
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b2 = fma(al, b1, be * b0);
        A[k] += b2;
        b0 = b1;
        b1 = b2;
    }
}
// two independent terms interleaved (instruction-level parallelism for the two recurrences)
__device__ __forceinline__ void far_accum2(double D1, double h21, double w1, double p1, double D2, double h22, double w2, double p2,
                                           double m2h, double mhh, double (&A)[kFarK])
{
    const double u1 = rcp3(fma(D1, D1, h21)), u2 = rcp3(fma(D2, D2, h22));
    const double al1 = (D1 * m2h) * u1, be1 = mhh * u1, al2 = (D2 * m2h) * u2, be2 = mhh * u2;
    double b01 = w1 * u1, b02 = w2 * u2;
    double b11 = al1 * b01, b12 = al2 * b02;
    A[0] += (b01 - p1) + (b02 - p2);
    A[1] += b11 + b12;
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b21 = fma(al1, b11, be1 * b01), b22 = fma(al2, b12, be2 * b02);
        A[k] += b21 + b22;
        b01 = b11; b11 = b21;
        b02 = b12; b12 = b22;
    }
}
// first-order line mixing (modm.f90:777-786): (g + c*(D+t))/((D+t)^2+h2), c = +-cq
__device__ __forceinline__ void far_accum_mix(double D, double h2, double cg, double cq, double hh, double m2h, double mhh, double (&A)[kFarK])
{
    const double u = rcp3(fma(D, D, h2));
    const double al = (D * m2h) * u, be = mhh * u;
    const double g1 = fma(cq, D, cg), g2 = cq * hh;
    double b0 = u;
    double b1 = al * b0;
    A[0] = fma(g1, b0, A[0]);
    A[1] = fma(g1, b1, fma(g2, b0, A[1]));
#pragma unroll
    for (int k = 2; k < kFarK; k++) {
        const double b2 = fma(al, b1, be * b0);
        A[k] = fma(g1, b2, fma(g2, b1, A[k]));
        b0 = b1;
        b1 = b2;
    }
}

// CTA-wide sums of the per-thread coefficients in a fixed order (deterministic); result in s_coef[kFarK]
template <int NT>
__device__ __forceinline__ void reduce_coefs(const double (&A)[kFarK], int tid, double (*s_red)[NT], double (*s_red2)[8], double* s_coef)
{
    constexpr int CH = NT / 8;
#pragma unroll
    for (int i = 0; i < kFarK; i++) s_red[i][tid] = A[i];
    __syncthreads();
    if (tid < kFarK * 8) {
        const int i = tid >> 3, part = tid & 7;
        double t = 0.;
        for (int j = 0; j < CH; j++) t += s_red[i][part * CH + ((j + tid) & (CH - 1))];
        s_red2[i][part] = t;
    }
    __syncthreads();
    if (tid < kFarK) {
        double t = 0.;
#pragma unroll
        for (int j = 0; j < 8; j++) t += s_red2[tid][j];
        s_coef[tid] = t;
    }
    __syncthreads();
}

// binomial coefficients C(j,k), j,k < kFarK (polynomial translation between hierarchy levels)
struct BinomTable {
    double c[kFarK][kFarK];
    constexpr BinomTable() : c{}
    {
        for (int j = 0; j < kFarK; j++)
            for (int k = 0; k < kFarK; k++) {
                double v = 0.;
                if (k <= j) {
                    v = 1.;
                    for (int i = 1; i <= k; i++) v = v * (double)(j - k + i) / (double)i;
                }
                c[j][k] = v;
            }
    }
};
__constant__ BinomTable c_binom = BinomTable();

// =============================================================================================
// far_kernel: one launch per hierarchy level, top level first.  CTA = (level tile, layer, profile).
// Expands the lines that are far at this level but not at the parent level in kFarK Taylor terms about
// the tile centre, adds the parent tile's polynomial re-expanded about this centre (exact polynomial
// translation), and writes one coefficient set per molecule slot.  After the level-0 launch every line
// that is far from a level-0 tile -- at whatever level it was expanded -- is contained in that tile's
// coefficients; final_kernel evaluates them once per frequency.
// =============================================================================================
struct FarArgs {
    int32_t nlay, nseg, n_pad, nslot;
    const Segment* seg;
    const FarPiece* pieces;   // [ntiles][nseg*kPiecePerSeg] work list of this level (plan_kernel)
    const TileHdr* hdr;
    const TileHdr* phdr;      // parent level or null
    const double* pcoef;      // parent's coefficients [ptile][L][slot][kFarK]
    int32_t S;                // tiles of this level per parent tile
    int32_t combined;         // 1: one coefficient set, lines weighted by their molecule's column amount (nslot == 1)
    const double* planes;
    const double* lcplanes;   // [L][LCP_NPLANES][nlc_pad]
    const int32_t* lcidx;     // static per line
    int32_t nlc_pad, pad1;
    const LayerDev* lay;
    double* coef;             // [tile][L][slot][kFarK]
    unsigned long long* counters;
};

__global__ void __launch_bounds__(128, MRTM_FAR_MINB) far_kernel(FarArgs a)
{
    constexpr int NT = 128;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    FarPiece* s_pc = reinterpret_cast<FarPiece*>(s_dyn);        // the tile's work list
    __shared__ double s_red[kFarK][NT];
    __shared__ double s_red2[kFarK][8];
    __shared__ double s_coef[kFarK];
    __shared__ double s_pcoef[kMaxSlots * kFarK];       // the parent tile's coefficients
    __shared__ double s_w[kMaxSegments];                // column amount of each segment's molecule
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int64_t Ltot = (int64_t)gridDim.y * gridDim.z;
    const int64_t L = (int64_t)blockIdx.z * a.nlay + blockIdx.y;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const int nseg = a.nseg;
    const int ptile = tile / a.S;
    const TileHdr th = a.hdr[tile];
    const int np = th.npieces;
    {
        const int nw = np * (int)(sizeof(FarPiece) / 4);
        const int* src = reinterpret_cast<const int*>(a.pieces + (size_t)tile * nseg * kPiecePerSeg);
        int* dst = reinterpret_cast<int*>(s_pc);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
        if (a.pcoef) {
            const double* pin = a.pcoef + ((size_t)ptile * Ltot + L) * a.nslot * kFarK;
            for (int i = tid; i < a.nslot * kFarK; i += NT) s_pcoef[i] = pin[i];
        }
        for (int s = tid; s < nseg; s += NT) s_w[s] = ly.wk[a.seg[s].mol - 1];
    }
    __syncthreads();
    const double cen = 0.5 * (th.wlo + th.whi), hh = 0.5 * (th.whi - th.wlo);
    const double m2h = -2. * hh, mhh = -hh * hh;
    double alpha = 0., beta = 0.;       // parent variable s_p = alpha + beta*s
    if (a.pcoef) {
        const TileHdr ph = a.phdr[ptile];
        const double pc = 0.5 * (ph.wlo + ph.whi), phh = 0.5 * (ph.whi - ph.wlo);
        if (phh > 0.) { alpha = (cen - pc) / phh; beta = hh / phh; }
    }
    double* out = a.coef + ((size_t)tile * Ltot + L) * a.nslot * kFarK;
    const double* pin = a.pcoef ? s_pcoef : nullptr;
    // coefficient tid of the parent's polynomial p(alpha + beta*s) re-expanded in s:
    // beta^tid * sum_{j>=tid} c_j C(j,tid) alpha^(j-tid)
    auto translated = [&](const double* pcf) -> double {
        double acc = 0.;
        for (int j = kFarK - 1; j >= tid; j--) acc = fma(acc, alpha, pcf[j] * c_binom.c[j][tid]);
        double bk = 1.;
        for (int i = 0; i < tid; i++) bk *= beta;
        return acc * bk;
    };
    // number of leading non-mixing pieces
    int npm = 0;
    while (npm < np && !((s_pc[npm].info >> 17) & 1)) npm++;

    // work units [vbeg,vend) of the non-mixing pieces [pbeg,pend): one unit (two independent recurrences) per thread
    // and step, the next unit's line parameters already loading
    auto accumulate = [&](int pbeg, int pend, int vbeg, int vend, bool weighted, double (&A)[kFarK]) {
        auto fetch = [&](int v, int& pi, double& D1, double& g1, double& w1, double& p1, double& D2, double& g2, double& w2, double& p2) {
            D1 = 1.; g1 = 1.; w1 = 0.; p1 = 0.; D2 = 1.; g2 = 1.; w2 = 0.; p2 = 0.;
            if (v >= vend) return;
            while (pi + 1 < pend && v >= s_pc[pi + 1].off) pi++;
            const FarPiece fp = s_pc[pi];
            const int kk = v - fp.off;
            const double ws = weighted ? s_w[fp.info & 0xffff] : 1.;
            if ((fp.info >> 16) & 1) {                    // both resonances of one line: cen - xnu and cen + xnu
                const int q = fp.lo + kk;
                const double xnu = __ldg(pXNU + q);
                g1 = g2 = __ldg(pH2 + q);
                w1 = w2 = ws * __ldg(pCN + q);
                p1 = p2 = ws * __ldg(pP3 + q);
                D1 = cen - xnu;
                D2 = cen + xnu;
            } else {                                      // two adjacent single-resonance lines
                const int q = fp.lo + 2 * kk;
                D1 = cen - __ldg(pXNU + q);
                g1 = __ldg(pH2 + q);
                w1 = ws * __ldg(pCN + q);
                p1 = ws * __ldg(pP3 + q);
                if (2 * kk + 1 < fp.n) {
                    D2 = cen - __ldg(pXNU + q + 1);
                    g2 = __ldg(pH2 + q + 1);
                    w2 = ws * __ldg(pCN + q + 1);
                    p2 = ws * __ldg(pP3 + q + 1);
                }
            }
        };
        int pi = pbeg;
        int v = vbeg + tid;
        double D1, g1, w1, p1, D2, g2, w2, p2;
        fetch(v, pi, D1, g1, w1, p1, D2, g2, w2, p2);
        while (v < vend) {
            const int vn = v + NT;
            double D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n;
            fetch(vn, pi, D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n);
            far_accum2(D1, g1, w1, p1, D2, g2, w2, p2, m2h, mhh, A);
            D1 = D1n; g1 = g1n; w1 = w1n; p1 = p1n;
            D2 = D2n; g2 = g2n; w2 = w2n; p2 = p2n;
            v = vn;
        }
    };
    // first-order mixing pieces [pbeg,pend): few lines, one line (both resonances) per thread and step
    auto accumulate_mix = [&](int pbeg, int pend, bool weighted, double (&A)[kFarK]) {
        for (int pi = pbeg; pi < pend; pi++) {
            const FarPiece fp = s_pc[pi];
            const double ws = weighted ? s_w[fp.info & 0xffff] : 1.;
            for (int q = fp.lo + tid; q < fp.lo + fp.n; q += NT) {
                const double xnu = __ldg(pXNU + q), h2 = __ldg(pH2 + q), cg = ws * __ldg(pP3 + q);
                const double cq = ws * lc1_slope(__ldg(pCN + q), h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx[q]], ly.rp);
                far_accum_mix(cen - xnu, h2, cg, cq, hh, m2h, mhh, A);
                far_accum_mix(cen + xnu, h2, cg, -cq, hh, m2h, mhh, A);
            }
        }
    };

    if (a.combined) {
        // one coefficient set for all molecules: every line enters with its molecule's column amount W
        // (o = RFT * sum_mol W_mol*SF_mol, modm.f90:436-438, 265-267)
        double A[kFarK];
#pragma unroll
        for (int i = 0; i < kFarK; i++) A[i] = 0.;
        if (np > 0) {
            accumulate(0, npm, 0, th.nunits, true, A);
            accumulate_mix(npm, np, true, A);
            reduce_coefs<NT>(A, tid, s_red, s_red2, s_coef);
        }
        if (tid < kFarK) {
            double v = (np > 0) ? s_coef[tid] : 0.;
            if (pin) v += translated(pin);
            out[tid] = v;
        }
    } else {
        // one coefficient set per molecule slot; the pieces of a molecule are contiguous in both lists
        int s = 0;
        while (s < nseg) {
            const int mol = a.seg[s].mol, slot = a.seg[s].slot;
            int s_end = s;
            while (s_end < nseg && a.seg[s_end].mol == mol) s_end++;
            int pb = 0, pe, mb = npm, me;
            while (pb < npm && (s_pc[pb].info & 0xffff) < s) pb++;
            pe = pb;
            while (pe < npm && (s_pc[pe].info & 0xffff) < s_end) pe++;
            while (mb < np && (s_pc[mb].info & 0xffff) < s) mb++;
            me = mb;
            while (me < np && (s_pc[me].info & 0xffff) < s_end) me++;
            const bool active = ly.wk[mol - 1] != 0.;
            const bool work = active && (pe > pb || me > mb);
            if (work) {
                double A[kFarK];
#pragma unroll
                for (int i = 0; i < kFarK; i++) A[i] = 0.;
                if (pe > pb) {
                    const FarPiece last = s_pc[pe - 1];
                    accumulate(pb, pe, s_pc[pb].off, last.off + ((((last.info >> 16) & 1)) ? last.n : (last.n + 1) / 2), false, A);
                }
                accumulate_mix(mb, me, false, A);
                reduce_coefs<NT>(A, tid, s_red, s_red2, s_coef);
            }
            if (tid < kFarK) {
                double v = work ? s_coef[tid] : 0.;
                if (pin && active) v += translated(pin + (size_t)slot * kFarK);
                out[(size_t)slot * kFarK + tid] = v;
            }
            __syncthreads();        // s_coef is reused by the next molecule
            s = s_end;
        }
    }
    if (a.counters && tid == 0) {
        long long n_far = th.nterms;
        for (int pi = npm; pi < np; pi++) n_far += 2ll * s_pc[pi].n;
        atomicAdd(a.counters + 0, (unsigned long long)n_far);
    }
}

// =============================================================================================
// far_warp_kernel: far_kernel for the levels with many small tiles (level 0 above all) in the combined
// mode (one coefficient set for all molecules).  A WARP owns one (tile, layer): the four warps of a CTA
// take four consecutive layers of the same tile and share its work list in shared memory; each lane walks
// the tile's work units with stride 32, the coefficients are summed across the warp by shuffles (fixed
// butterfly order: deterministic) -- no CTA barrier after the list is staged, and four times more units
// per lane than far_kernel has per thread, which amortises the set-up, reduction and translation.
// =============================================================================================
#ifndef MRTM_FARW_MINB
#define MRTM_FARW_MINB 7
#endif
constexpr int kFarWarps = 4;
__global__ void __launch_bounds__(32 * kFarWarps, MRTM_FARW_MINB) far_warp_kernel(FarArgs a)
{
    constexpr int NT = 32 * kFarWarps;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    FarPiece* s_pc = reinterpret_cast<FarPiece*>(s_dyn);        // the tile's work list
    __shared__ double s_pcoef[kFarWarps][kFarK];                // the parent tile's coefficients, per layer
    __shared__ double s_w[kFarWarps][kMaxSegments];             // column amount of each segment's molecule, per layer
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const int nseg = a.nseg;
    const TileHdr th = a.hdr[tile];
    const int np = th.npieces;
    {
        const int nw = np * (int)(sizeof(FarPiece) / 4);
        const int* src = reinterpret_cast<const int*>(a.pieces + (size_t)tile * nseg * kPiecePerSeg);
        int* dst = reinterpret_cast<int*>(s_pc);
        for (int i = tid; i < nw; i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const int k = blockIdx.y * kFarWarps + wid;
    if (k >= a.nlay) return;                                    // no barrier follows
    const int64_t Ltot = (int64_t)a.nlay * gridDim.z;
    const int64_t L = (int64_t)blockIdx.z * a.nlay + k;
    const LayerDev& ly = a.lay[L];
    const double* pl = a.planes + (size_t)L * D_NPLANES * a.n_pad;
    const double* __restrict__ pXNU = pl + (size_t)D_XNU * a.n_pad;
    const double* __restrict__ pH2 = pl + (size_t)D_H2 * a.n_pad;
    const double* __restrict__ pCN = pl + (size_t)D_CN * a.n_pad;
    const double* __restrict__ pP3 = pl + (size_t)D_P3 * a.n_pad;
    const double* __restrict__ lcp = a.lcplanes + (size_t)L * LCP_NPLANES * a.nlc_pad;
    const int ptile = tile / a.S;
    double* sw = s_w[wid];
    for (int s = lane; s < nseg; s += 32) sw[s] = ly.wk[a.seg[s].mol - 1];
    if (a.pcoef && lane < kFarK) s_pcoef[wid][lane] = a.pcoef[((size_t)ptile * Ltot + L) * kFarK + lane];
    __syncwarp();
    const double cen = 0.5 * (th.wlo + th.whi), hh = 0.5 * (th.whi - th.wlo);
    const double m2h = -2. * hh, mhh = -hh * hh;
    int npm = 0;                        // number of leading non-mixing pieces
    while (npm < np && !((s_pc[npm].info >> 17) & 1)) npm++;

    double A[kFarK];
#pragma unroll
    for (int i = 0; i < kFarK; i++) A[i] = 0.;
    {
        const int vend = th.nunits;
        auto fetch = [&](int v, int& pi, double& D1, double& g1, double& w1, double& p1, double& D2, double& g2, double& w2, double& p2) {
            D1 = 1.; g1 = 1.; w1 = 0.; p1 = 0.; D2 = 1.; g2 = 1.; w2 = 0.; p2 = 0.;
            if (v >= vend) return;
            while (pi + 1 < npm && v >= s_pc[pi + 1].off) pi++;
            const FarPiece fp = s_pc[pi];
            const int kk = v - fp.off;
            const double ws = sw[fp.info & 0xffff];
            if ((fp.info >> 16) & 1) {                    // both resonances of one line: cen - xnu and cen + xnu
                const int q = fp.lo + kk;
                const double xnu = __ldg(pXNU + q);
                g1 = g2 = __ldg(pH2 + q);
                w1 = w2 = ws * __ldg(pCN + q);
                p1 = p2 = ws * __ldg(pP3 + q);
                D1 = cen - xnu;
                D2 = cen + xnu;
            } else {                                      // two adjacent single-resonance lines
                const int q = fp.lo + 2 * kk;
                D1 = cen - __ldg(pXNU + q);
                g1 = __ldg(pH2 + q);
                w1 = ws * __ldg(pCN + q);
                p1 = ws * __ldg(pP3 + q);
                if (2 * kk + 1 < fp.n) {
                    D2 = cen - __ldg(pXNU + q + 1);
                    g2 = __ldg(pH2 + q + 1);
                    w2 = ws * __ldg(pCN + q + 1);
                    p2 = ws * __ldg(pP3 + q + 1);
                }
            }
        };
        int pi = 0;
        int v = lane;
        double D1, g1, w1, p1, D2, g2, w2, p2;
        if (npm > 0) {
            fetch(v, pi, D1, g1, w1, p1, D2, g2, w2, p2);
            while (v < vend) {
                const int vn = v + 32;
                double D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n;
                fetch(vn, pi, D1n, g1n, w1n, p1n, D2n, g2n, w2n, p2n);
                far_accum2(D1, g1, w1, p1, D2, g2, w2, p2, m2h, mhh, A);
                D1 = D1n; g1 = g1n; w1 = w1n; p1 = p1n;
                D2 = D2n; g2 = g2n; w2 = w2n; p2 = p2n;
                v = vn;
            }
        }
        // first-order mixing pieces: few lines, one line (both resonances) per lane and step
        for (int pj = npm; pj < np; pj++) {
            const FarPiece fp = s_pc[pj];
            const double ws = sw[fp.info & 0xffff];
            for (int q = fp.lo + lane; q < fp.lo + fp.n; q += 32) {
                const double xnu = __ldg(pXNU + q), h2 = __ldg(pH2 + q), cg = ws * __ldg(pP3 + q);
                const double cq = ws * lc1_slope(__ldg(pCN + q), h2, lcp[(size_t)LCP_AIP * a.nlc_pad + a.lcidx[q]], ly.rp);
                far_accum_mix(cen - xnu, h2, cg, cq, hh, m2h, mhh, A);
                far_accum_mix(cen + xnu, h2, cg, -cq, hh, m2h, mhh, A);
            }
        }
    }
    // warp sums; lane i keeps coefficient i
    double mine = 0.;
#pragma unroll
    for (int i = 0; i < kFarK; i++) {
        double v = A[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == i) mine = v;
    }
    if (lane < kFarK) {
        if (a.pcoef) {
            // coefficient `lane` of the parent's polynomial p(alpha + beta*s) re-expanded in s:
            // beta^lane * sum_{j>=lane} c_j C(j,lane) alpha^(j-lane)
            const TileHdr ph = a.phdr[ptile];
            const double pc = 0.5 * (ph.wlo + ph.whi), phh = 0.5 * (ph.whi - ph.wlo);
            double alpha = 0., beta = 0.;
            if (phh > 0.) { alpha = (cen - pc) / phh; beta = hh / phh; }
            const double* pcf = s_pcoef[wid];
            double acc = 0.;
            for (int j = kFarK - 1; j >= lane; j--) acc = fma(acc, alpha, pcf[j] * c_binom.c[j][lane]);
            double bk = 1.;
            for (int i = 0; i < lane; i++) bk *= beta;
            mine += acc * bk;
        }
        a.coef[((size_t)tile * Ltot + L) * kFarK + lane] = mine;
    }
    if (a.counters && lane == 0) {
        long long n_far = th.nterms;
        for (int pj = npm; pj < np; pj++) n_far += 2ll * s_pc[pj].n;
        atomicAdd(a.counters + 0, (unsigned long long)n_far);
    }
}
