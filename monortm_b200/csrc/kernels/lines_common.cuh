// lines_common.cuh -- LinesArgs, tile plans, TMA / mbarrier primitives and small device helpers shared by the line-path kernels.
// Part of mrtm_kernels.cuh (included from there, inside namespace mrtm).
// =============================================================================================
// The line path (modm.f90:277-440).  plan_kernel classifies every (frequency tile, segment) once per call,
// far_kernel / far_warp_kernel expand the far lines level by level, near2_kernel (near_kernel for oversize
// tiles) evaluates the remaining (line, layer, frequency) triples with the reference's exact tests,
// voigt_kernel adds the Voigt-branch pairs and final_kernel closes the sums (RFT :257, continuum
// interpolation + RADFN :218-230, cloud liquid water :264, total :265-269).  LinesArgs is shared by them.
// =============================================================================================
struct SegWork;
struct TileHdr;
struct NearPiece;
constexpr int kMaxLevels = 8;   // far-field hierarchy: level 0 = the line kernel's own tiles
struct LinesArgs {
    int32_t nwn, nlay;            // frequencies in this call/chunk, layers per profile
    int32_t nseg, n_pad;
    int64_t iw0;                  // global 0-based index of wn[0] (gridded continuum interpolation)
    const double* wn;             // [nwn]
    const Segment* seg;           // [nseg]
    const double* xnu0;           // static centres (sorted inside segments)
    const int32_t *mol_s, *xf_s;  // static per line
    const double* sdep_s;
    const unsigned long long* key;
    const unsigned long long* keypre;   // [n_pad+1] prefix sums of key (selection hash of a whole range)
    double ff_ratio;              // far-field expansion: poles >= ff_ratio tile half-widths away; 0 = direct only
    double ffw_ratio;             // the same ratio for near2_kernel's in-warp expansion about the warp's own block
    unsigned long long* counters; // [2] far-field expansions, direct (line,frequency) evaluations (may be null)
    const double* planes;         // [L][D_NPLANES][n_pad]
    const double* lcplanes;       // [L][LCP_NPLANES][nlc_pad] coupling coefficients of the coupled lines (compact)
    const int32_t* lcidx_s;       // static per line: index into the compact coupling planes or -1
    int32_t nlc_pad, pad1;
    const LayerDev* lay;          // [L]
    // far-field hierarchy: level 0 = this kernel's tiles, level lv tiles are S^lv times wider
    int32_t nlev, S;
    const SegWork* plan[kMaxLevels];    // [ntiles_lv][nseg]
    const TileHdr* hdr[kMaxLevels];     // [ntiles_lv]
    const double* coef[kMaxLevels];     // lv >= 1: [tile][L][slot][kFarK] from far_kernel
    int32_t nslot, pad0;
    const NearPiece* near_pieces;       // [ntiles][kMaxNearPieces] (plan_kernel, level 0); null: near2_kernel not in use
    // Voigt-zone candidates per tile (vplan_kernel) or null: line indices, their segments, count (-1: walk the whole zone)
    const int* vcand; const unsigned char* vcand_seg; const int* vcand_count;
    const unsigned long long* layer_voigt;   // [L] bits of the smallest |Xnu| of the layer's Voigt-capable lines (derive_kernel), all ones = none
    const int32_t* slot_mol;            // [nslot]
    // continuum
    const double* absrb;          // [L][CP_COUNT][nptabs_pad]
    int32_t nptabs, nptabs_pad;
    int32_t cont_mask, pad_cm;    // bit c: plane c of absrb has an active component (bits of ContPlane)
    double v1abs, v2abs, v1, dvset;
    // outputs (any may be null).  Strides in elements.
    double* o;        int64_t o_lds;  int64_t o_prof;     // o[iw + k*o_lds + prof*o_prof]
    double* o_v;                  // zeroed scratch with o's strides: voigt_kernel adds there (it then runs beside the near field), or null
    double* o_by_mol; int64_t obm_ldm; int64_t obm_ldk;   // [iw + (mol-1)*ldm + k*ldk] (+prof*ldk*nlay)
    double* oc;                                           // same strides as o_by_mol
    double* o_clw;                                        // same strides as o
    const double* odxsec;                                 // same strides as o (input, may be null)
    long long* sel_count; unsigned long long* sel_hash;   // same strides as o
    int* errflag;                                         // bit1: SDVOIGT negative real part
};

// the per-(line,layer) quantities that are not stored (see DPlane)
struct ColdLine { double hw, ad, stild, aip, bip; };
__device__ __forceinline__ ColdLine cold_line(const double* __restrict__ pl, int n_pad, int q, const int32_t* __restrict__ lcidx,
                                              const double* __restrict__ lcp, int nlc_pad)
{
    ColdLine c;
    c.hw = sqrt(pl[(size_t)D_H2 * n_pad + q]);
    const double vt = pl[(size_t)D_VT * n_pad + q];
    c.ad = vt >= 0. ? vt * 0.01 : 1.0;
    c.stild = (pl[(size_t)D_CN * n_pad + q] * kPI) / c.hw;
    const int lci = lcidx[q];
    c.aip = lci >= 0 ? lcp[(size_t)LCP_AIP * nlc_pad + lci] : 0.;
    c.bip = lci >= 0 ? lcp[(size_t)LCP_BIP * nlc_pad + lci] : 0.;
    return c;
}
// first-order mixing slope of a CLS_O2_LC1 line: CN*AIP*(1/HWHM_C)*RP (modm.f90:781-782 regrouped), as derive_kernel would form it
__device__ __forceinline__ double lc1_slope(double cn, double h2, double aip, double rp) { return cn * (aip * (1 / sqrt(h2)) * rp); }

__device__ __forceinline__ int lower_bound_d(const double* a, int lo, int hi, double v)
{   // first index in [lo,hi) with a[i] >= v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* a, int lo, int hi, double v)
{   // first index in [lo,hi) with a[i] > v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Can a (line, frequency) pair of this layer take the Voigt branch for a frequency <= whi?  A pair needs
// |WN-Xnu| <= 100*HWHM_D <= 2e-3*|Xnu| (T < 3000 K, molecular mass >= 1), so Xnu <= 1.0021*whi; the layer's
// Voigt-capable lines all have |Xnu| >= the recorded minimum.
__device__ __forceinline__ bool voigt_possible(const unsigned long long* layer_voigt, int64_t L, double whi)
{
    const unsigned long long bound = (unsigned long long)__double_as_longlong(fabs(whi) * 1.01 + 1e-3);
    return bound >= layer_voigt[L];
}

// RADFN, lblrtm_sub.f90:36-97
__device__ __forceinline__ double radfn(double vi, double xkt)
{
    if (xkt > 0.0) {
        double x = vi / xkt;
        if (x <= 0.01) return 0.5 * x * vi;
        if (x <= 10.0) {
            double e = exp(-x);
            return vi * (1. - e) / (1. + e);
        }
        return vi;
    }
    return vi;
}

// the same with the quotient (1-e)/(1+e) through a Newton reciprocal (~1 ulp); the branch variable x stays the exact quotient
// (x as vi*(1/xkt) with an exact fallback next to the thresholds was measured: final_kernel 197 -> 207 us, not kept)
__device__ __forceinline__ double radfn_r(double vi, double xkt)
{
    if (xkt > 0.0) {
        double x = vi / xkt;
        if (x <= 0.01) return 0.5 * x * vi;
        if (x <= 10.0) {
            double e = exp(-x);
            return (vi * (1. - e)) * rcp3(1. + e);
        }
        return vi;
    }
    return vi;
}

// ODCLW_TKC / Forward_TKC, CloudOptProp.f90:29-157 (binary64 here; the parity build evaluates
// the d0-literal expressions in binary128 and rounds, a ~1e-16 relative difference)
__device__ __noinline__ double odclw_tkc(double wn, double temp, double clw)
{
    const double a_1 = 8.110808E+01, b_1 = 4.433736E-03, c_1 = 1.301700E-13, d_1 = 6.627126E+02;
    const double a_2 = 2.025164E+00, b_2 = 1.072976E-02, c_2 = 1.011945E-14, d_2 = 6.089168E+02;
    const double t_c = 1.342433E+02;
    double freq = wn * kCLIGHT / 1.e9;
    double tc = temp - 273.15;
    double frq = freq * 1.e9;
    double cl = kCLIGHT / 100.;
    double eps_s = 87.9144 - 0.404399 * tc + 9.58726E-4 * (tc * tc) - 1.32802E-6 * (tc * tc * tc);
    double delta_1 = a_1 * exp(-b_1 * tc), tau_1 = c_1 * exp(d_1 / (tc + t_c));
    double delta_2 = a_2 * exp(-b_2 * tc), tau_2 = c_2 * exp(d_2 / (tc + t_c));
    double w = 2. * kPI * frq;
    double den1 = 1. + (w * tau_1) * (w * tau_1), den2 = 1. + (w * tau_2) * (w * tau_2);
    double eps1 = eps_s - (w * w) * ((tau_1 * tau_1 * delta_1) / den1 + (tau_2 * tau_2 * delta_2) / den2);
    double eps2 = w * ((tau_1 * delta_1) / den1 + (tau_2 * delta_2) / den2);
    cplx e = cmk(eps1, eps2);
    cplx re = (cmk(eps1 - 1., eps2)) / (cmk(eps1 + 2., eps2));
    (void)e;
    double alpha = 6. * kPI * re.im * frq * 1.e-3 / cl;
    return alpha * clw;
}

// ---- TMA (bulk async copy) + mbarrier primitives used to stream line-parameter tiles ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#ifndef MRTM_LINES_MINB_32
#define MRTM_LINES_MINB_32 20       // near_kernel on one-warp CTAs: resident CTAs per SM the register budget is sized for
#endif
#ifndef MRTM_LINES_MINB
#define MRTM_LINES_MINB 4
#endif
#ifndef MRTM_UNROLL_BOTH
#define MRTM_UNROLL_BOTH 2
#endif
#define MRTM_PRAGMA(x) _Pragma(#x)
#define MRTM_UNROLL(n) MRTM_PRAGMA(unroll n)
#ifndef MRTM_NEAR_STAGES_32
#define MRTM_NEAR_STAGES_32 2
#endif
#ifndef MRTM_NEAR_STAGES_64
#define MRTM_NEAR_STAGES_64 4
#endif
constexpr int kTile = 128;      // lines per smem tile
constexpr int kStages = 8;      // tile ring
// near_kernel's ring per CTA size: a one-warp CTA (32-channel tiles of channel lists) consumes a 128-line stage in some 5000 cycles,
// two stages cover the copy latency and leave room for 20 CTAs per SM instead of 6 (36 KB of ring at eight stages; measured 3: 64.9 ms, 2 with a 96-register budget: 62.9 ms per 64 profiles of the ensemble)
template <int NT> __host__ __device__ constexpr int near_stages() { return NT <= 32 ? MRTM_NEAR_STAGES_32 : (NT <= 64 ? MRTM_NEAR_STAGES_64 : kStages); }
constexpr int kPrefetch = 5;    // TMA jobs in flight ahead of the consumer; a warp may run kStages-kPrefetch tiles ahead of the slowest
#ifndef MRTM_FARK
#define MRTM_FARK 14
#endif
constexpr int kFarK = MRTM_FARK; // Taylor terms of the far-field expansion (degree kFarK-1); <= 16 (reduce_coefs)
static_assert(kFarK >= 4 && kFarK <= 16, "kFarK out of range");
constexpr int kMaxBp = 12;      // break points per segment
constexpr int kMaxRun = 6;      // direct runs per segment

// sub-range mode bits
constexpr int M_EDGE = 1;       // per-(line,frequency) window test |WN-Xnu| > 25 (modm.f90:384)
constexpr int M_NEG = 2;        // per-(line,frequency) test WN+Xnu <= 25 (modm.f90:746)
constexpr int M_VOIGT = 4;      // per-(line,frequency) test |WN-Xnu| <= 100*HWHM_D (modm.f90:427)
constexpr int M_NEAR = 8;       // direct evaluation (a pole of the line is too close to the tile to expand)

// Classification of one (molecule, class) segment against one frequency tile.  It does not depend on
// the layer: the margins are maxima over the layers of the batch (shift margin, 100*HWHM_D), so one
// plan serves every layer and every profile of a call.
struct SegWork {
    // searched fields, in the order of the plan tasks (kept contiguous)
    int q0, q1;        // lines that can be inside the 25 cm-1 window of some frequency of the tile
    int eb, ec;        // [q0,eb) and [ec,q1): window-edge bands
    int n0, n1;        // [n0,n1): band where WN+Xnu<=25 flips; < n0: both resonances for every frequency
    int v0, v1;        // [v0,v1): Voigt zone
    int z0;            // < z0: the negative-frequency pole -Xnu is near the tile
    int f0, f1;        // [f0,f1): the pole +Xnu is near the tile
    // derived
    int nbp;                     // break points bp[0..nbp-1]; sub-range u = [bp[u], bp[u+1])
    int bp[kMaxBp];
    unsigned char mode[kMaxBp];  // mode bits of sub-range u; 0 = far field (Taylor expansion)
    int nrun;                    // maximal runs of consecutive direct (mode != 0) sub-ranges
    int run_lo[kMaxRun], run_hi[kMaxRun], run_t0[kMaxRun], run_nt[kMaxRun];
    int run_off[kMaxRun];        // near_kernel (stage-all mode): offset of the run in the CTA's staging area
    int run_u0[kMaxRun], run_u1[kMaxRun];   // sub-ranges [u0,u1) that make up the run
    int tma;                     // class streams its direct runs through shared memory
    int has_far;
};
constexpr int kSegTasks = 11;
struct TileHdr {
    double wlo, whi;             // frequency extent of the tile
    int total_lines;             // lines (padded to 4 per run) of all direct runs of the streamed classes
    int npieces, nunits;         // far-field work list of the tile: pieces, and work units of the non-mixing pieces
    int nterms;                  // Taylor expansions of the non-mixing pieces (statistics)
    int nnear, pad;              // direct sub-ranges of the streamed classes (near2_kernel's piece list); -1 = too many
};
// One direct sub-range of a streamed (TMA) class in the CTA's staging area (near2_kernel)
struct NearPiece {
    int soff;                    // staged offset of its first line
    int n;                       // lines
    int q0;                      // index (staged order of the line list) of its first line
    int info;                    // segment | mode << 8 | negall << 16 | kind << 17 (0 PED, 1 O2, 2 O2_LC35)
};
constexpr int kMaxNearPieces = 192;
#ifndef MRTM_NEAR2_CAP
#define MRTM_NEAR2_CAP 512
#endif
#ifndef MRTM_NEAR2_MINB
#define MRTM_NEAR2_MINB 5
#endif
#ifndef MRTM_FAR_MINB
#define MRTM_FAR_MINB 6
#endif
constexpr int kNearCap = MRTM_NEAR2_CAP;        // staged lines per CTA of near2_kernel (tiles with more go to near_kernel)
static_assert(kNearCap <= kStages * kTile && kNearCap % 32 == 0, "kNearCap");
// One contiguous range of lines that far_kernel expands for a tile: far at this level and not at the parent level.
constexpr int kPiecePerSeg = 12;
struct FarPiece {
    int lo, n;                   // lines [lo, lo+n)
    int off;                     // first work unit of the piece in the tile's unit numbering (non-mixing pieces); a unit is
                                 // one line with both resonances, or two adjacent single-resonance lines
    int info;                    // segment | both << 16 | mix << 17
};

// the far (mode 0) sub-ranges of a segment at this level, minus the ones the parent level already
// expanded (the parent's far set is a subset of the child's by construction of the margins)
template <class Fn>
__device__ __forceinline__ void for_each_far_piece(const SegWork& wk, const SegWork* pk, Fn fn)
{
    for (int u = 0; u + 1 < wk.nbp; u++) {
        if (wk.mode[u] != 0) continue;
        const int lo = wk.bp[u], hi = wk.bp[u + 1];
        int cur = lo;
        if (pk) {
            for (int v = 0; v + 1 < pk->nbp && cur < hi; v++) {
                if (pk->mode[v] != 0) continue;
                const int pa = pk->bp[v], pb = pk->bp[v + 1];
                if (pb <= cur) continue;
                if (pa >= hi) break;
                if (pa > cur) fn(cur, pa, lo);
                cur = pb > cur ? pb : cur;
            }
        }
        if (cur < hi) fn(cur, hi, lo);
    }
}
