// mrtm_api.cu -- context, staging upload, orchestration and the extern "C" entry points of
// libmonortm_b200.so (include/monortm_b200.h).  No CPU fallback exists: every compute entry
// point needs a live sm_100 device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "mrtm_kernels.cuh"
#include "mrtm_stage.h"
#include "tables/mtckd_tables.inc"
#include "tables/mtckd_ir_tables.inc"

namespace mrtm {
const double* tips_qoft();
const double* tips_tdat();
int tips_rows();
int tips_row(int mol, int iso);
}  // namespace mrtm

using namespace mrtm;

// ---------------------------------------------------------------------------------------------
struct DevBuf {   // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
};

struct mrtm_ctx {
    // mrtm_init_multi: a context without device resources of its own that drives one single-device context per GPU
    std::vector<mrtm_ctx*> peers;
    std::vector<int64_t> mg_bounds;               // frequency partition of the last frequency-split call (ndev + 1 entries)
    std::vector<double> mg_ms;                    // host-measured time of each device's share of that call
    int64_t mg_nwn = 0;
    int mg_mode = 0;                              // last split: 0 none, 1 by profile, 2 by frequency
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8];
    std::string err;
    bool have_lines = false;
    HostLines hl;
    LinesDev ld;
    std::vector<void*> line_allocs;
    Segment* seg_dev = nullptr;
    ContTablesDev tb;
    TipsDev tips;
    std::vector<void*> table_allocs;
    int32_t* tips_row_dev = nullptr;
    int* errflag_dev = nullptr;
    unsigned long long* counters_dev = nullptr;   // [2] far expansions, direct evaluations
    double ffw_ratio = 6.0;                       // in-warp expansion ratio of near2_kernel (MRTM_FFW_RATIO)
    double ff_ratio = 6.0;                        // far-field pole-distance ratio (MRTM_FF_RATIO; 0 = direct only): the 14-term series
                                                  // truncates at (1/6)^14 = 1.3e-11 of an expanded line's own contribution (bar on layer optical
                                                  // depths: 1e-9; measured 8e-12 on the bench workload; 8: 7e-14 and 5 % slower)
    DevBuf b_vtmax;                               // [0] sm_max bits, [1..nseg] vtmax per segment
    DevBuf b_lvoigt;                              // [L] per-layer "has Voigt-capable lines" flag
    DevBuf b_plan[kMaxLevels], b_hdr[kMaxLevels], b_coef[kMaxLevels], b_pieces[kMaxLevels], b_npieces;
    int64_t farw_min = 1400;                      // (tile, layer) pairs of a level from which far_warp_kernel takes it (MRTM_FARW_MIN)
    int ff_levels = 4, ff_S = 4;                  // far-field hierarchy (MRTM_FF_LEVELS 1..4, MRTM_FF_S)
    cudaStream_t side = nullptr;                  // high-priority stream: plans and far-field levels overlap derive / near field
    cudaEvent_t evf[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in = nullptr;                  // emiss / reflc arrived (host-buffer path)
    double tile_width = 0.1;                      // cm-1, MRTM_TILE_WIDTH (see the tile-size choice in run_device)
    const void* span_ptr = nullptr; int64_t span_nwn = 0; double span_val = -1.;   // cached span of a device-resident wn
    int use_voigt_t = 1;                          // dense tiles: 1 voigtT_kernel, 0 voigt_kernel (MRTM_VOIGT_T, comparison arm)
    int voigt_side = 0;                           // MRTM_VOIGT_SIDE=1: Voigt branch on the side stream into a scratch plane (measured: no gain, off by default)
    DevBuf b_ov;
    int use_side = 1;                             // MRTM_SIDE_STREAM=0: everything on one stream
    int use_near2 = 1;                            // per-warp re-planning near-field kernel (MRTM_NEAR2=0 disables)
    int coarse_tile = 32;                         // MRTM_COARSE_TILE: channels per tile on channel lists (128, 64, 32)
    int coarse_f = 2;                             // MRTM_COARSE_F: frequencies per thread on 128-frequency tiles (CTA of 128/F threads)
    int use_neart = 1;                            // transposed direct kernel for coarse frequency lists (MRTM_NEART=0: near_kernel)
    DevBuf b_vcand, b_vcseg, b_vccount;
    int use_near3 = 0;                            // MRTM_NEAR3=1: plan-driven near-field kernel (813 us alone against near2_kernel's 890 us,
                                                  // 26 % fewer instructions; inside the two-stream step near2_kernel is 1 % faster: default)
    int force_f = 0;                              // MRTM_LINES_F: frequencies per thread of the line kernels (1, 2, 4; 0 = chosen per call)
    int near3_lb = 0;                             // layers per CTA of near3_kernel (MRTM_NEAR3_LB; 0 = chosen per call)
    DevBuf b_t3, b_pool3, b_segof3;
    // plan cache (see kernels/plan.cuh): margins the cached plans were built with, frequency hashes, replan flag
    DevBuf b_pcache;
    int use_plan_cache = 1;                       // MRTM_PLAN_CACHE=0: rebuild the plans on every call
    int64_t stage_gen = 0;                        // bumped by mrtm_stage_lines
    struct PlanKey { int64_t nwn, stage_gen, batch_layers; int F, nlev, S, nseg, near2, near3; double ff_ratio, ffw_ratio; const void* bufs[4 * kMaxLevels + 4]; };
    PlanKey plan_key;
    bool plan_key_valid = false;
    DevBuf b_layer, b_scorc, b_absrb, b_planes, b_lcplanes, b_o, b_obm, b_oc, b_in[16], b_out[16], b_sel[2], b_tmps, b_fbeta;
    // cross sections (mrtm_stage_xsec): region descriptors, staged tables, per-call scratch
    std::vector<XsRegionDev> xs_regs;
    int64_t xs_tab_per_layer = 0;                 // sum over regions of (npts + 3)
    int xs_nmol = 0;
    DevBuf b_xsreg, b_xsdat, b_xslay, b_xstab, b_xsneed, b_xsod, b_xsin[4];
    mrtm_stats st;
    // asynchronous device-resident calls (mrtm_profiles_dev): timing events and device flags are read back lazily
    bool pending = false, pending_lines = false, pending_rt = false;
    int deferred_rc = MRTM_OK;
    size_t planes_budget = (size_t)24 << 30;   // of the 180 GB (ensemble: 24 GB batches are 5 % faster than 8 GB)
};

static thread_local std::string g_err_noctx;

static int set_err(mrtm_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_err_noctx = msg;
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return set_err(ctx, MRTM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static int ensure(mrtm_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return MRTM_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return set_err(ctx, MRTM_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    b.cap = want;
    return MRTM_OK;
}

template <class T>
static int upload(mrtm_ctx* ctx, const std::vector<T>& v, std::vector<void*>& owner, const T** out)
{
    void* d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    owner.push_back(d);
    if (!v.empty()) CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return MRTM_OK;
}

static int upload_arr(mrtm_ctx* ctx, const double* src, size_t n, std::vector<void*>& owner, const double** out)
{
    std::vector<double> v(src, src + n);
    return upload(ctx, v, owner, out);
}

// ---------------------------------------------------------------------------------------------
extern "C" const char* mrtm_version(void) { return "monortm_b200 0.1 (hot path of MonoRTM v5.6, MT_CKD_3.5)"; }

extern "C" const char* mrtm_strerror(int code)
{
    switch (code) {
    case MRTM_OK: return "ok";
    case MRTM_ENODEV: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    case MRTM_ECUDA: return "CUDA runtime error";
    case MRTM_EARG: return "bad argument";
    case MRTM_ENOLINES: return "line list not staged (call mrtm_stage_lines first)";
    case MRTM_ELINEFILE: return "malformed line store";
    case MRTM_ERANGE: return "spectral range out of bounds";
    case MRTM_ESDVOIGT: return "SDVOIGT: REAL(v) < 0 (reference STOPs, modm.f90:1062)";
    case MRTM_EIDU: return "ERROR IN IDU. OPTION NOT SUPPORTED YET";
    case MRTM_ENOMEM: return "out of (device) memory";
    case MRTM_EIO: return "file missing or malformed";
    case MRTM_EXSEC: return "cross sections: not staged, the resampled grid exceeds the reference's xspd_int(0:10000000), or the convolution does not terminate";
    case MRTM_ETIPS: return "TIPS: temperature outside 70-3000 K or partition sum <= 0";
    default: return "unknown error";
    }
}

extern "C" const char* mrtm_last_error(mrtm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err_noctx.c_str(); }

extern "C" int mrtm_init(int device, mrtm_ctx** out)
{
    mrtm_ctx* ctx = nullptr;
    if (!out) return MRTM_EARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return set_err(nullptr, MRTM_ENODEV, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return set_err(nullptr, MRTM_EARG, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return set_err(nullptr, MRTM_ENODEV, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return set_err(nullptr, MRTM_ENODEV, "device is not sm_100 (this library ships sm_100a code only)");
    ctx = new mrtm_ctx();
    ctx->device = device;
    std::memset(&ctx->st, 0, sizeof ctx->st);
    std::memset(&ctx->ld, 0, sizeof ctx->ld);
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return set_err(nullptr, MRTM_ECUDA, "cudaSetDevice failed"); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return set_err(nullptr, MRTM_ECUDA, "cudaStreamCreate failed"); }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, hi) != cudaSuccess) ctx->side = nullptr;
        for (auto& ev : ctx->evf) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming);
    }
    if (const char* s = std::getenv("MRTM_SIDE_STREAM")) ctx->use_side = std::atoi(s) != 0;
    if (const char* s = std::getenv("MRTM_TILE_WIDTH")) ctx->tile_width = std::atof(s);
    if (const char* s = std::getenv("MRTM_VOIGT_SIDE")) ctx->voigt_side = std::atoi(s) != 0;
    if (const char* s = std::getenv("MRTM_VOIGT_T")) ctx->use_voigt_t = std::atoi(s);
    if (const char* s = std::getenv("MRTM_PLANES_GB")) ctx->planes_budget = (size_t)(std::atof(s) * (double)(1ull << 30));
    if (const char* s = std::getenv("MRTM_FF_LEVELS")) ctx->ff_levels = std::min(std::max(std::atoi(s), 1), kMaxLevels);
    if (const char* s = std::getenv("MRTM_NEAR2")) ctx->use_near2 = std::atoi(s) != 0;
    if (const char* s = std::getenv("MRTM_PLAN_CACHE")) ctx->use_plan_cache = std::atoi(s) != 0;
    if (const char* s = std::getenv("MRTM_NEAR3")) ctx->use_near3 = std::atoi(s) != 0;
    if (const char* s = std::getenv("MRTM_COARSE_TILE")) ctx->coarse_tile = std::atoi(s);
    if (const char* s = std::getenv("MRTM_COARSE_F")) { const int v = std::atoi(s); ctx->coarse_f = (v == 2 || v == 4) ? v : 1; }
    if (const char* s = std::getenv("MRTM_NEART")) ctx->use_neart = std::atoi(s);      // 0 off, 1 by tile width, 2 every F = 1 call
    if (const char* s = std::getenv("MRTM_LINES_F")) ctx->force_f = std::atoi(s);
    if (const char* s = std::getenv("MRTM_NEAR3_LB")) ctx->near3_lb = std::max(std::atoi(s), 0);
    if (const char* s = std::getenv("MRTM_FFW_RATIO")) ctx->ffw_ratio = std::max(std::atof(s), 4.0);
    if (const char* s = std::getenv("MRTM_FARW_MIN")) ctx->farw_min = std::atoll(s);
    if (const char* s = std::getenv("MRTM_FF_S")) ctx->ff_S = std::min(std::max(std::atoi(s), 2), 64);
    if (const char* s = std::getenv("MRTM_FF_RATIO")) {
        double v = std::atof(s);
        ctx->ff_ratio = (v <= 0.) ? 0. : std::max(v, 4.0);
    }
    // continuum + TIPS tables -> HBM (about 350 KB)
    int rc;
#define UP(NAME, FIELD) if ((rc = upload_arr(ctx, NAME, sizeof(NAME) / sizeof(double), ctx->table_allocs, &ctx->tb.FIELD))) { *out = ctx; return rc; }
    UP(MTCKD_SH2O_296, sh2o_296) UP(MTCKD_SH2O_260, sh2o_260) UP(MTCKD_FH2O, fh2o) UP(MTCKD_FCO2, fco2)
    UP(MTCKD_N2RT_296, n2_296) UP(MTCKD_N2RT_296_SF, n2_296_sf) UP(MTCKD_N2RT_220, n2_220) UP(MTCKD_N2RT_220_SF, n2_220_sf)
    UP(MTCKD_XFAC_RHU, xfac_rhu) UP(MTCKD_CO2_TDEP_BANDHEAD, co2_tdep)
    UP(MTCKD_XFACCO2, xfacco2) UP(MTCKD_N2F_272, n2f_272) UP(MTCKD_N2F_228, n2f_228) UP(MTCKD_N2F_AH2O, n2f_ah2o) UP(MTCKD_N2F1, n2f1)
    UP(MTCKD_O3CH_X, o3ch_x) UP(MTCKD_O3CH_Y, o3ch_y) UP(MTCKD_O3CH_Z, o3ch_z) UP(MTCKD_O3HH0, o3hh0) UP(MTCKD_O3HH1, o3hh1)
    UP(MTCKD_O3HH2, o3hh2) UP(MTCKD_O3HUV, o3huv) UP(MTCKD_O2F, o2f) UP(MTCKD_O2F_T, o2f_t) UP(MTCKD_O2INF1, o2inf1)
    UP(MTCKD_O2INF3, o2inf3) UP(MTCKD_O2VIS, o2vis) UP(MTCKD_O2FUV, o2fuv)
#undef UP
    if ((rc = upload_arr(ctx, tips_qoft(), (size_t)tips_rows() * 119, ctx->table_allocs, &ctx->tips.qoft))) { *out = ctx; return rc; }
    if ((rc = upload_arr(ctx, tips_tdat(), 119, ctx->table_allocs, &ctx->tips.tdat))) { *out = ctx; return rc; }
    ctx->tips.row = nullptr;
    if (cudaMalloc(&ctx->errflag_dev, sizeof(int)) != cudaSuccess) { *out = ctx; return set_err(ctx, MRTM_ENOMEM, "cudaMalloc errflag"); }
    cudaMemset(ctx->errflag_dev, 0, sizeof(int));
    if (cudaMalloc(&ctx->counters_dev, 2 * sizeof(unsigned long long)) != cudaSuccess) { *out = ctx; return set_err(ctx, MRTM_ENOMEM, "cudaMalloc counters"); }
    cudaMemset(ctx->counters_dev, 0, 2 * sizeof(unsigned long long));
    *out = ctx;
    return MRTM_OK;
}

static void free_lines(mrtm_ctx* ctx)
{
    for (void* p : ctx->line_allocs) cudaFree(p);
    ctx->line_allocs.clear();
    ctx->have_lines = false;
}

extern "C" int mrtm_free(mrtm_ctx* ctx)
{
    if (!ctx) return MRTM_OK;
    if (!ctx->peers.empty()) {
        for (mrtm_ctx* c : ctx->peers) mrtm_free(c);
        delete ctx;
        return MRTM_OK;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_lines(ctx);
    for (void* p : ctx->table_allocs) cudaFree(p);
    DevBuf* bufs[] = {&ctx->b_vcand, &ctx->b_vcseg, &ctx->b_vccount, &ctx->b_xsreg, &ctx->b_xsdat, &ctx->b_xslay, &ctx->b_xstab, &ctx->b_xsneed, &ctx->b_xsod, &ctx->b_xsin[0], &ctx->b_xsin[1],
                      &ctx->b_xsin[2], &ctx->b_xsin[3], &ctx->b_pcache, &ctx->b_t3, &ctx->b_pool3, &ctx->b_segof3, &ctx->b_ov, &ctx->b_lvoigt, &ctx->b_vtmax, &ctx->b_layer, &ctx->b_scorc, &ctx->b_absrb, &ctx->b_planes, &ctx->b_lcplanes, &ctx->b_o, &ctx->b_obm, &ctx->b_oc, &ctx->b_sel[0], &ctx->b_sel[1], &ctx->b_tmps, &ctx->b_fbeta, &ctx->b_npieces};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (int i = 0; i < kMaxLevels; i++) {
        if (ctx->b_plan[i].p) cudaFree(ctx->b_plan[i].p);
        if (ctx->b_hdr[i].p) cudaFree(ctx->b_hdr[i].p);
        if (ctx->b_coef[i].p) cudaFree(ctx->b_coef[i].p);
        if (ctx->b_pieces[i].p) cudaFree(ctx->b_pieces[i].p);
    }
    for (auto& b : ctx->b_in) if (b.p) cudaFree(b.p);
    for (auto& b : ctx->b_out) if (b.p) cudaFree(b.p);
    if (ctx->errflag_dev) cudaFree(ctx->errflag_dev);
    if (ctx->counters_dev) cudaFree(ctx->counters_dev);
    for (auto& ev : ctx->ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->evf) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MRTM_OK;
}

extern "C" int64_t mrtm_num_lines(mrtm_ctx* ctx)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];
    return (ctx && ctx->have_lines) ? ctx->hl.n : 0;
}

static int stage_lines_single(mrtm_ctx* ctx, const int64_t nblm[MRTM_MXMOL], int64_t iim,
                              const int64_t* iso, const double* xnu0, const double* deltnu,
                              const double* e, const double* alps, const double* alpf,
                              const double* x, const double* xg, const double* s0,
                              const double* rmol, const double* sdep,
                              const int32_t* brd_mol_flg, const double* brd_mol_tmp,
                              const double* brd_mol_hw, const double* brd_mol_shft);

// run fn(i, peer i) on one host thread per device; returns the first error and copies its text to the parent
template <class Fn>
static int for_each_peer(mrtm_ctx* ctx, Fn fn)
{
    const size_t n = ctx->peers.size();
    std::vector<int> rcs(n, MRTM_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; i++) th.emplace_back([&, i]() { rcs[i] = fn((int)i, ctx->peers[i]); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (rcs[i]) return set_err(ctx, rcs[i], "device " + std::to_string(ctx->peers[i]->device) + ": " + ctx->peers[i]->err);
    return MRTM_OK;
}

extern "C" int mrtm_stage_lines(mrtm_ctx* ctx, const int64_t nblm[MRTM_MXMOL], int64_t iim,
                                const int64_t* iso, const double* xnu0, const double* deltnu,
                                const double* e, const double* alps, const double* alpf,
                                const double* x, const double* xg, const double* s0,
                                const double* rmol, const double* sdep,
                                const int32_t* brd_mol_flg, const double* brd_mol_tmp,
                                const double* brd_mol_hw, const double* brd_mol_shft)
{
    if (ctx && !ctx->peers.empty())                  // replicated on every device
        return for_each_peer(ctx, [&](int, mrtm_ctx* c) {
            return stage_lines_single(c, nblm, iim, iso, xnu0, deltnu, e, alps, alpf, x, xg, s0, rmol, sdep, brd_mol_flg, brd_mol_tmp,
                                      brd_mol_hw, brd_mol_shft);
        });
    return stage_lines_single(ctx, nblm, iim, iso, xnu0, deltnu, e, alps, alpf, x, xg, s0, rmol, sdep, brd_mol_flg, brd_mol_tmp,
                              brd_mol_hw, brd_mol_shft);
}

static int stage_lines_single(mrtm_ctx* ctx, const int64_t nblm[MRTM_MXMOL], int64_t iim,
                              const int64_t* iso, const double* xnu0, const double* deltnu,
                              const double* e, const double* alps, const double* alpf,
                              const double* x, const double* xg, const double* s0,
                              const double* rmol, const double* sdep,
                              const int32_t* brd_mol_flg, const double* brd_mol_tmp,
                              const double* brd_mol_hw, const double* brd_mol_shft)
{
    if (!ctx) return MRTM_EARG;
    if (!nblm || !iso || !xnu0 || !deltnu || !e || !alps || !alpf || !x || !xg || !s0 || !rmol || !sdep)
        return set_err(ctx, MRTM_EARG, "null line array");
    CU(cudaSetDevice(ctx->device));
    free_lines(ctx);
    int rc = stage_lines_host(nblm, iim, iso, xnu0, deltnu, e, alps, alpf, x, xg, s0, rmol, sdep,
                              brd_mol_flg, brd_mol_tmp, brd_mol_hw, brd_mol_shft, ctx->hl);
    if (rc) return set_err(ctx, rc, ctx->hl.error);
    HostLines& h = ctx->hl;
    LinesDev& d = ctx->ld;
    std::memset(&d, 0, sizeof d);
    d.n = (int32_t)h.n;
    d.n_pad = (int32_t)h.n_pad;
    auto& own = ctx->line_allocs;
#define UPV(F) if ((rc = upload(ctx, h.F, own, &d.F))) return rc;
    UPV(mol) UPV(iso) UPV(xf) UPV(cls) UPV(sidx) UPV(lcidx) UPV(brdidx) UPV(segidx)
    UPV(xnu0) UPV(s0adj) UPV(e) UPV(alpf) UPV(alps) UPV(x) UPV(deltnu) UPV(sdep) UPV(mass) UPV(dopf)
    UPV(lc) UPV(lc_self) UPV(brd) UPV(scor_index) UPV(slot_mol)
#undef UPV
    {
        std::vector<unsigned long long> k(h.key.begin(), h.key.end());
        if ((rc = upload(ctx, k, own, &d.key))) return rc;
        std::vector<unsigned long long> kp(k.size() + 1, 0ull);
        for (size_t i = 0; i < k.size(); i++) kp[i + 1] = kp[i] + k[i];
        if ((rc = upload(ctx, kp, own, &d.keypre))) return rc;
    }
    d.nsi = (int32_t)h.scor_index.size();
    {
        const Segment* sd = nullptr;
        if ((rc = upload(ctx, h.segments, own, &sd))) return rc;
        ctx->seg_dev = const_cast<Segment*>(sd);
        std::vector<int32_t> rows(h.scor_index.size());
        for (size_t s = 0; s < rows.size(); s++) {
            int mol = h.scor_index[s] % MRTM_NSCOR1 + 1, iso_ = h.scor_index[s] / MRTM_NSCOR1 + 1;
            rows[s] = tips_row(mol, iso_);
        }
        const int32_t* rd = nullptr;
        if ((rc = upload(ctx, rows, own, &rd))) return rc;
        ctx->tips.row = rd;
    }
    ctx->have_lines = true;
    ctx->stage_gen++;
    ctx->plan_key_valid = false;
    ctx->st.lines_staged = h.n;
    return MRTM_OK;
}

// ---------------------------------------------------------------------------------------------
// continuum index set-up (layer independent): table accessors (contnm.f90:1441-1456 and twins),
// pre_xint (:1146-1164) and the XINT loop bounds (lblrtm_sub.f90:13-17), evaluated on the host.
// ---------------------------------------------------------------------------------------------
static ContGrid make_grid(double v1abs, double v2abs, int nptabs, double v1s, double v2s, double dvs, int npts, bool active,
                          double eps = 0.01, bool cap = true)
{
    ContGrid g;
    std::memset(&g, 0, sizeof g);
    const double dvabs = 1.0, onemi = 0.999;
    double dvc = dvs;
    double v1c = v1abs - dvc, v2c = v2abs + dvc;
    long long i1;
    if (v1c < v1s) i1 = -1;
    else i1 = (long long)((v1c - v1s) / dvs + eps);
    v1c = v1s + dvs * (double)(i1 - 1);
    long long i2 = (long long)((v2c - v1s) / dvs + eps);
    long long nptc = i2 - i1 + 3;
    if (cap && nptc > npts) nptc = npts + 4;
    v2c = v1c + dvs * (double)(nptc - 1);
    long long nb1 = (long long)(2 + (v1s - v1abs) / dvabs + 1.e-5);
    long long ist = std::max<long long>(1, nb1);
    long long nb2 = (long long)(1 + (v2s - v1abs) / dvabs + 1.e-5);
    long long last = std::min<long long>(nptabs, nb2);
    long long ilo = (long long)((v1c + dvc - v1abs) / dvabs + 1. + onemi);
    ilo = std::max(ilo, ist);
    long long ihi = (long long)((v2c - dvc - v1abs) / dvabs + onemi);
    ihi = std::min(ihi, last);
    g.v1c = v1c;
    g.dvc = dvc;
    g.nptc = (int32_t)nptc;
    g.i1 = (int32_t)i1;
    g.ilo = (int32_t)ilo;
    g.ihi = (int32_t)ihi;
    g.active = (active && nptc > 0) ? 1 : 0;
    return g;
}

// O2INF2 (contnm.f90:9227-9279) sets up its own grid: no table, the coefficient range clamps V1C, V2C
static ContGrid make_grid_o2inf2(double v1abs, double v2abs, int nptabs, bool active)
{
    ContGrid g;
    std::memset(&g, 0, sizeof g);
    const double v1s = 9100., v2s = 11000., dvs = 2., dvabs = 1.0, onemi = 0.999;
    double dvc = dvs, v1c = v1abs - dvc, v2c = v2abs + dvc;
    if (v1c < v1s) v1c = v1s - 2. * dvs;
    if (v2c > v2s) v2c = v2s + 2. * dvs;
    long long nptc = (long long)((v2c - v1c) / dvc + 3.01);
    v2c = v1c + dvc * (double)(nptc - 1);
    long long ist = std::max<long long>(1, (long long)(2 + (v1s - v1abs) / dvabs + 1.e-5));
    long long last = std::min<long long>(nptabs, (long long)(1 + (v2s - v1abs) / dvabs + 1.e-5));
    long long ilo = std::max((long long)((v1c + dvc - v1abs) / dvabs + 1. + onemi), ist);
    long long ihi = std::min((long long)((v2c - dvc - v1abs) / dvabs + onemi), last);
    g.v1c = v1c;
    g.dvc = dvc;
    g.nptc = (int32_t)nptc;
    g.i1 = 0;
    g.ilo = (int32_t)ilo;
    g.ihi = (int32_t)ihi;
    g.active = (active && nptc > 0) ? 1 : 0;
    return g;
}

// every component of CONTNM with its gate (contnm.f90:325, 387, 484, 536, 555, 603, 657, 709, 745, 773, 807, 834, 857, 906,
// 963, 1025, 1107); returns the ContPlane mask of the species that have an active component
static int make_cont_grids(ContGrid* g, double v1, double v2, double v1abs, double v2abs, int nptabs, const double cn[7], int* rayl)
{
    const double xself = cn[0], xfrgn = cn[1], xco2c = cn[2], xo3cn = cn[3], xo2cn = cn[4], xn2cn = cn[5], xrayl = cn[6];
    const bool gate_h2o = (v2 > -20.0) && (v1 < 20000.);
    g[CB_H2O_SELF] = make_grid(v1abs, v2abs, nptabs, -20.0, 20000.0, 10.0, 2003, gate_h2o && xself > 0.);
    g[CB_H2O_FRGN] = make_grid(v1abs, v2abs, nptabs, -20.0, 20000.0, 10.0, 2003, gate_h2o && xfrgn > 0.);
    g[CB_CO2] = make_grid(v1abs, v2abs, nptabs, -4.0, 10000.0, 2.0, 5003, (v2 > -20.0) && (v1 < 10000.) && xco2c > 0);
    g[CB_N2_ROT] = make_grid(v1abs, v2abs, nptabs, -10., 350., 5.0, 73, (v2 > -10.0) && (v1 < 350.) && xn2cn > 0.);
    g[CB_N2_FUND] = make_grid(v1abs, v2abs, nptabs, 1997.784896, 2901.576661, 3.981461525, 228, (v2 > 2001.77) && (v1 < 2897.59) && xn2cn > 0.);
    g[CB_N2_OVER] = make_grid(v1abs, v2abs, nptabs, 4340.0, 4910.0, 3.0, 191, (v2 > 4340.0) && (v1 < 4910.) && xn2cn > 0.);
    g[CB_O3_CHAP] = make_grid(v1abs, v2abs, nptabs, 8920.0, 24665.0, 5.0, 3150, v2 > 8920.0 && v1 <= 24665.0 && xo3cn > 0.);
    g[CB_O3_HH] = make_grid(v1abs, v2abs, nptabs, 27370., 40800., 5.0, 2687, v2 > 27370. && v1 < 40800. && xo3cn > 0.);
    g[CB_O3_UV] = make_grid(v1abs, v2abs, nptabs, 40800., 54000., 100., 133, v2 > 40800. && v1 < 54000. && xo3cn > 0.);
    {
        // Hartley-Huggins / UV hand-over at 40800 cm-1: each branch restores the other's side of the 1 cm-1 grid after its
        // interpolation (ABSBSV, contnm.f90:573-599 and :619-640), i.e. it contributes only on its own side of I_FIX
        const long long i_fix = (long long)((40800. - v1abs) / 1.0 + 1.001);
        ContGrid& hh = g[CB_O3_HH];
        const double vj_last = hh.v1c + hh.dvc * (double)(hh.nptc - 1);
        if (hh.active && (vj_last > 40815.) && (v2 > 40800)) hh.ihi = (int32_t)std::min<long long>(hh.ihi, i_fix - 1);
        ContGrid& uv = g[CB_O3_UV];
        if (uv.active && (v1 < 40800)) uv.ilo = (int32_t)std::max<long long>(uv.ilo, i_fix);
    }
    g[CB_O2_FUND] = make_grid(v1abs, v2abs, nptabs, 1340.0, 1850.0, 5.0, 103, (v2 > 1340.0) && (v1 < 1850.) && xo2cn > 0.);
    g[CB_O2_INF1] = make_grid(v1abs, v2abs, nptabs, 7536.0, 8500.0, 2.0, 483, (v2 > 7536.0) && (v1 < 8500.) && xo2cn > 0.);
    g[CB_O2_INF2] = make_grid_o2inf2(v1abs, v2abs, nptabs, (v2 > 9100.0) && (v1 < 11000.) && xo2cn > 0.);
    g[CB_O2_INF3] = make_grid(v1abs, v2abs, nptabs, 12961.5, 13221.5, 1.0, 261, (v2 > 12961.5) && (v1 < 13221.5) && xo2cn > 0.);
    g[CB_O2_VIS] = make_grid(v1abs, v2abs, nptabs, 15140.0, 29870.0, 10.0, 1474, (v2 > 15000.0) && (v1 < 29870.) && xo2cn > 0.);
    g[CB_O2_HERZ] = make_grid(v1abs, v2abs, nptabs, 36000., 99999., 10., 0, v2 > 36000.0 && xo2cn > 0., 0.01, false);
    g[CB_O2_FUV] = make_grid(v1abs, v2abs, nptabs, 56740.0, 86960.0, 20.0, 1512, v2 > 56740.0 && xo2cn > 0., 1.e-5, true);
    *rayl = (v2 >= 820. && xrayl > 0.) ? 1 : 0;
    int mask = 0;
    if (g[CB_H2O_SELF].active || g[CB_H2O_FRGN].active) mask |= 1 << CP_H2O;
    if (g[CB_CO2].active) mask |= 1 << CP_CO2;
    if (g[CB_O3_CHAP].active || g[CB_O3_HH].active || g[CB_O3_UV].active) mask |= 1 << CP_O3;
    for (int b = CB_O2_FUND; b <= CB_O2_FUV; b++) if (g[b].active) mask |= 1 << CP_O2;
    if (g[CB_N2_ROT].active || g[CB_N2_FUND].active || g[CB_N2_OVER].active) mask |= 1 << CP_N2;
    if (*rayl) mask |= 1 << CP_RAYL;
    return mask;
}

// everything below works on device pointers
struct RunDesc {
    int64_t nprof, nwn, nlay, nmol;
    const double* wn;
    double dvset;
    const double *p, *t, *tz, *clw, *wkl, *wbrodl, *scor;
    double sclcpl, sclhw, y0res, cntnm[7];
    int64_t ibrd, irt, iout, idu;
    double* tmpsfc;                 // device (nprof)
    const double *emiss, *reflc;
    double *rad, *tb, *tmr, *trtot, *rup, *rdn;
    double* o;                      // (nwn,nlay,nprof) or null
    double *o_by_mol, *oc, *o_clw;  // modm-style outputs, nprof==1 only
    const double* odxsec;
    const double* xamnt;            // device (ld_xamnt, nlay, nprof) or null: cross-section amounts, odxsec computed here
    int64_t ld_xamnt;
    long long* sel_count;
    unsigned long long* sel_hash;
    bool do_lines, do_tmr, do_rtm;
    bool async_ok;                  // the caller does not need timings / device flags before returning (mrtm_profiles_dev)
    double v1, v2;
    int64_t iw0;
    int line_mode;                  // mrtm_opts.line_mode
    double wn_span;                 // |wn[nwn-1] - wn[0]| of this call's frequencies, < 0: unknown (tile-size heuristic only)
    double* h_spec[6];              // host destinations of rad, tb, tmr, trtot, rup, rdn (or null): copied per batch right behind rt_kernel
    cudaEvent_t rt_ready;           // recorded when emiss / reflc have arrived on another stream (or null)
};

template <int F, int NT>
static void launch_lines(const LinesArgs& la, dim3 grid, bool sel, cudaStream_t s, cudaEvent_t far_done, cudaStream_t sv,
                         const Near3Args* n3, bool neart, int voigt_t)
{
    // near field (direct), Voigt branch, then polynomial + continuum + totals
    const size_t dyn = sizeof(double) * near_stages<NT>() * 4 * kTile + (size_t)std::max(la.nseg, 1) * sizeof(SegWork);
    const size_t dynT = sizeof(double) * kStages * 4 * kTile + (size_t)std::max(la.nseg, 1) * sizeof(SegWork);     // nearT_kernel's ring
    if (neart && F == 1) {                     // coarse frequency lists: lanes own lines, warps own frequencies
        if (sel) {
            cudaFuncSetAttribute(nearT_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynT);
            nearT_kernel<true><<<grid, 32 * kNTW, dynT, s>>>(la);
        } else {
            cudaFuncSetAttribute(nearT_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynT);
            nearT_kernel<false><<<grid, 32 * kNTW, dynT, s>>>(la);
        }
    } else if (la.near_pieces && n3 && F == 4) {      // production path: plan-driven lists, a group of layers per CTA
        cudaFuncSetAttribute(near3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kN3Smem);
        near3_kernel<<<dim3(grid.x, (grid.y + n3->lb - 1) / n3->lb, grid.z), 256, kN3Smem, s>>>(la, *n3);
    } else if (la.near_pieces) {       // tiles whose direct lines fit the staging area: per-warp re-planning kernel
        const size_t dyn2 = sizeof(double) * 4 * (kNearCap + 8) + sizeof(unsigned short) * (NT / 32) * 2 * (kNearCap + 8) + kNearCap +
                            sizeof(NearPiece) * kMaxNearPieces + (size_t)std::max(la.nseg, 1) * sizeof(SegWork);
        if (sel) {
            cudaFuncSetAttribute(near2_kernel<F, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2);
            near2_kernel<F, true, NT><<<grid, NT, dyn2, s>>>(la);
        } else {
            cudaFuncSetAttribute(near2_kernel<F, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2);
            near2_kernel<F, false, NT><<<grid, NT, dyn2, s>>>(la);
        }
    }
    if (neart && F == 1) {
    } else if (sel) {
        cudaFuncSetAttribute(near_kernel<F, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        near_kernel<F, true, NT><<<grid, NT, dyn, s>>>(la);
    } else {
        cudaFuncSetAttribute(near_kernel<F, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        near_kernel<F, false, NT><<<grid, NT, dyn, s>>>(la);
    }
    // with a scratch plane for its sums the Voigt branch runs on the side stream (after the far field, beside the near field)
    // dense tiles (no candidate lists from vplan_kernel): lines over warps, lanes over each line's run of frequencies
    constexpr bool kHasT = (NT == 128 && F >= 2);
    const bool vt_dense = kHasT && voigt_t != 0 && !la.vcand_count;
    cudaStream_t svk = (la.o_v && sv) ? sv : s;
    if constexpr (kHasT) {
        if (vt_dense) voigtT_kernel<F, NT><<<grid, NT, 0, svk>>>(la);
        else voigt_kernel<F, NT><<<grid, NT, 0, svk>>>(la);
    } else {
        voigt_kernel<F, NT><<<grid, NT, 0, svk>>>(la);
    }
    if (la.o_v && sv) cudaEventRecord(far_done, sv);
    if (far_done) cudaStreamWaitEvent(s, far_done, 0);     // the far-field coefficients come from the side stream
    if ((la.cont_mask & ((1 << CP_O3) | (1 << CP_O2) | (1 << CP_RAYL))) == 0) final_kernel<F, NT, true><<<grid, NT, 0, s>>>(la);
    else final_kernel<F, NT, false><<<grid, NT, 0, s>>>(la);
}

static int run_xsec_device(mrtm_ctx* ctx, int64_t nwn, const double* wn, int64_t nlay, const double* p, const double* t,
                           int64_t ld_xamnt, const double* xamnt, double* od, cudaStream_t s);

// what an asynchronous call left undone: kernel times from the recorded events, device counters and error flags
static int finalize_pending(mrtm_ctx* ctx)
{
    if (!ctx->pending) return ctx->deferred_rc;
    ctx->pending = false;
    mrtm_stats& st = ctx->st;
    CU(cudaEventSynchronize(ctx->pending_rt ? ctx->ev[5] : ctx->ev[3]));
    float ms = 0.f;
    if (ctx->pending_lines) {
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); st.last_derive_kernel_ms += ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); st.last_lines_kernel_ms += ms;
    }
    if (ctx->pending_rt) { cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); st.last_rt_kernel_ms += ms; }
    if (ctx->pending_lines) {
        int flag = 0;
        unsigned long long cnt[2] = {0ull, 0ull};
        CU(cudaMemcpy(&flag, ctx->errflag_dev, sizeof(int), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(cnt, ctx->counters_dev, sizeof cnt, cudaMemcpyDeviceToHost));
        st.far_expansions = (double)cnt[0];
        st.direct_evals = (double)cnt[1];
        if (flag & 1) ctx->deferred_rc = set_err(ctx, MRTM_ETIPS, mrtm_strerror(MRTM_ETIPS));
        else if (flag & 2) ctx->deferred_rc = set_err(ctx, MRTM_ESDVOIGT, mrtm_strerror(MRTM_ESDVOIGT));
        else if (flag & (16 | 32)) ctx->deferred_rc = set_err(ctx, MRTM_EXSEC, mrtm_strerror(MRTM_EXSEC));
    }
    return ctx->deferred_rc;
}

static int run_device(mrtm_ctx* ctx, const RunDesc& r, cudaStream_t s)
{
    if (r.nprof < 1 || r.nwn < 1 || r.nlay < 1) return set_err(ctx, MRTM_EARG, "nprof, nwn, nlay must be >= 1");
    if (r.nwn > 0x7fffffff / 64 || r.nlay > 65535) return set_err(ctx, MRTM_EARG, "nwn/nlay too large for one call");
    finalize_pending(ctx);          // the previous asynchronous call's events and flags (its error, if any, was for mrtm_sync)
    ctx->deferred_rc = MRTM_OK;
    mrtm_stats& st = ctx->st;
    st.last_lines_kernel_ms = st.last_rt_kernel_ms = st.last_derive_kernel_ms = 0.;
    const int64_t nwn = r.nwn, nlay = r.nlay;
    DevBuf& bo = ctx->b_o;

    if (r.do_lines) {
        if (!ctx->have_lines) return set_err(ctx, MRTM_ENOLINES, mrtm_strerror(MRTM_ENOLINES));
        if (r.nmol < 1 || r.nmol > MRTM_MXMOL) return set_err(ctx, MRTM_EARG, "nmol out of 1..39");
        if ((r.o_by_mol || r.oc || r.o_clw) && r.nprof != 1) return set_err(ctx, MRTM_EARG, "per-molecule outputs need nprof == 1");
    }
    if ((r.do_rtm || r.do_tmr) && r.do_rtm && r.idu != 1) return set_err(ctx, MRTM_EIDU, mrtm_strerror(MRTM_EIDU));

    // spectral set-up (modm.f90:180-185)
    const double v1 = r.v1, v2 = r.v2, dvabs = 1.0;
    const double v1abs = (double)((long long)v1) - 3. * dvabs;
    const double v2abs = (double)((long long)(v2 + 3. * dvabs + 0.5));
    const int nptabs = (int)((v2abs - v1abs) / dvabs + 1.5);
    const int nptabs_pad = ((nptabs + 4 + 3) / 4) * 4;

    const HostLines& h = ctx->hl;
    const int n_pad = r.do_lines ? (int)h.n_pad : 0;
    const int nlc_pad = r.do_lines ? (int)(((h.lc.size() / 16) + 8) & ~(size_t)7) : 8;
    // profiles per batch
    int64_t B = r.nprof;
    if (r.do_lines) {
        size_t per_prof = (size_t)nlay * D_NPLANES * (size_t)n_pad * 8;
        int64_t bmem = (int64_t)std::max<size_t>(1, ctx->planes_budget / std::max<size_t>(per_prof, 1));
        B = std::min<int64_t>(B, bmem);
        B = std::min<int64_t>(B, std::max<int64_t>(1, 65535 / nlay));
    }
    B = std::min<int64_t>(B, 65535);
    const bool own_o = (r.o == nullptr);
    int rc;
    if (own_o && (rc = ensure(ctx, bo, (size_t)nwn * nlay * B * 8))) return rc;

    ContArgs ca;
    int cont_mask = 0;
    if (r.do_lines) {
        std::memset(&ca, 0, sizeof ca);
        cont_mask = make_cont_grids(ca.g, v1, v2, v1abs, v2abs, nptabs, r.cntnm, &ca.rayl_active);
        for (int i = 0; i < 7; i++) ca.cntnm[i] = r.cntnm[i];
        ca.v1abs = v1abs;
        ca.nptabs = nptabs;
        ca.nptabs_pad = nptabs_pad;
        ca.tb = ctx->tb;
        if ((rc = ensure(ctx, ctx->b_layer, (size_t)B * nlay * sizeof(LayerDev)))) return rc;
        if ((rc = ensure(ctx, ctx->b_scorc, (size_t)B * nlay * std::max(1, (int)ctx->ld.nsi) * 8))) return rc;
        if ((rc = ensure(ctx, ctx->b_absrb, (size_t)B * nlay * CP_COUNT * nptabs_pad * 8))) return rc;
        if ((rc = ensure(ctx, ctx->b_planes, (size_t)B * nlay * D_NPLANES * (size_t)n_pad * 8))) return rc;
        if ((rc = ensure(ctx, ctx->b_lcplanes, (size_t)B * nlay * LCP_NPLANES * (size_t)nlc_pad * 8))) return rc;
        if ((rc = ensure(ctx, ctx->b_vtmax, (1 + std::max<size_t>(1, h.segments.size())) * 8))) return rc;
        CU(cudaMemsetAsync(ctx->errflag_dev, 0, sizeof(int), s));
        CU(cudaMemsetAsync(ctx->counters_dev, 0, 2 * sizeof(unsigned long long), s));
        st.nominal_evals = (double)h.n * (double)nlay * (double)nwn * (double)r.nprof;
        st.inwindow_evals = -1.;
    }

    for (int64_t b0 = 0; b0 < r.nprof; b0 += B) {
        const int64_t nb = std::min<int64_t>(B, r.nprof - b0);
        const int64_t Lb = nb * nlay;
        double* o_batch = own_o ? (double*)bo.p : r.o + (size_t)b0 * nwn * nlay;
        if (r.do_lines) {
            LayerPrepArgs pa;
            std::memset(&pa, 0, sizeof pa);
            pa.nlayers = Lb;
            pa.nlay = nlay;
            pa.nmol = (int32_t)r.nmol;
            pa.ibrd = (int32_t)r.ibrd;
            pa.p = r.p + (size_t)b0 * nlay;
            pa.t = r.t + (size_t)b0 * nlay;
            pa.clw = r.clw + (size_t)b0 * nlay;
            pa.wkl = r.wkl + (size_t)b0 * nlay * MRTM_MXMOL;
            pa.wbrodl = r.wbrodl + (size_t)b0 * nlay;
            for (int i = 0; i < 7; i++) pa.cntnm[i] = r.cntnm[i];
            pa.max_abs_deltnu = h.max_abs_deltnu;
            pa.max_abs_brd_dshift = h.max_abs_brd_dshift;
            pa.out = (LayerDev*)ctx->b_layer.p;
            pa.scor_full = r.scor ? r.scor + (size_t)b0 * nlay * MRTM_NSCOR1 * MRTM_NSCOR2 : nullptr;
            pa.tips = ctx->tips;
            pa.nsi = ctx->ld.nsi;
            pa.scor_index = ctx->ld.scor_index;
            pa.scorc = (double*)ctx->b_scorc.p;
            pa.errflag = ctx->errflag_dev;
            pa.sm_max_bits = (unsigned long long*)ctx->b_vtmax.p;
            pa.vtmax = (unsigned long long*)ctx->b_vtmax.p + 1;
            pa.seg = ctx->seg_dev;
            pa.nseg = (int32_t)h.segments.size();
            CU(cudaMemsetAsync(ctx->b_vtmax.p, 0, (1 + std::max<size_t>(1, h.segments.size())) * 8, s));
            layer_prep_kernel<<<(unsigned)((Lb + 127) / 128), 128, 0, s>>>(pa);
            st.kernel_launches++;
            // plans (they need only the layer tables) and the far-field levels run on a second, high-priority stream:
            // side: [layer_prep] plan ... | [derive] far ...      main: continuum, derive | [plan] near, voigt | [far] final
            const bool side_on = ctx->use_side && ctx->side != nullptr;
            cudaStream_t sp = side_on ? ctx->side : s;
            if (side_on) {
                CU(cudaEventRecord(ctx->evf[0], s));
                CU(cudaStreamWaitEvent(sp, ctx->evf[0], 0));
            }

            ca.nlayers = Lb;
            ca.lay = (const LayerDev*)ctx->b_layer.p;
            ca.absrb = (double*)ctx->b_absrb.p;
            int nptc_max = 0;
            for (int b = 0; b < CB_COUNT; b++) if (ca.g[b].active) nptc_max = std::max(nptc_max, (int)ca.g[b].nptc);
            const size_t smem = (size_t)(nptc_max + 8) * 8;
            if (smem > 200 * 1024) return set_err(ctx, MRTM_EARG, "spectral range too wide for one call (continuum coefficient grid)");
            if (smem > 48 * 1024) CU(cudaFuncSetAttribute(continuum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            continuum_kernel<<<(unsigned)Lb, 128, smem, s>>>(ca);
            st.kernel_launches++;

            DeriveArgs da;
            std::memset(&da, 0, sizeof da);
            da.nlayers = Lb;
            da.ln = ctx->ld;
            da.lay = (const LayerDev*)ctx->b_layer.p;
            da.scorc = (const double*)ctx->b_scorc.p;
            da.sclcpl = r.sclcpl;
            da.sclhw = r.sclhw;
            da.y0res = r.y0res;
            da.ibrd = (int32_t)r.ibrd;
            da.planes = (double*)ctx->b_planes.p;
            da.lcplanes = (double*)ctx->b_lcplanes.p;
            da.nlc_pad = nlc_pad;
            da.nseg = (int32_t)h.segments.size();
            if ((rc = ensure(ctx, ctx->b_lvoigt, (size_t)Lb * sizeof(unsigned long long)))) return rc;
            CU(cudaMemsetAsync(ctx->b_lvoigt.p, 0xff, (size_t)Lb * sizeof(unsigned long long), s));
            da.layer_voigt = (unsigned long long*)ctx->b_lvoigt.p;
            CU(cudaEventRecord(ctx->ev[0], s));
            derive_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)((Lb + kDeriveLayers - 1) / kDeriveLayers)), 256, 0, s>>>(da);
            CU(cudaEventRecord(ctx->ev[1], s));
            if (side_on) CU(cudaEventRecord(ctx->evf[1], s));
            st.kernel_launches++;

            LinesArgs la;
            std::memset(&la, 0, sizeof la);
            la.nwn = (int32_t)nwn;
            la.nlay = (int32_t)nlay;
            la.nseg = (int32_t)h.segments.size();
            la.n_pad = n_pad;
            la.iw0 = r.iw0;
            la.wn = r.wn;
            la.seg = ctx->seg_dev;
            la.xnu0 = ctx->ld.xnu0;
            la.mol_s = ctx->ld.mol;
            la.xf_s = ctx->ld.xf;
            la.sdep_s = ctx->ld.sdep;
            la.key = ctx->ld.key;
            la.keypre = ctx->ld.keypre;
            la.ff_ratio = (r.line_mode == 1) ? 0. : ctx->ff_ratio;
            la.ffw_ratio = ctx->ffw_ratio;
            la.counters = ctx->counters_dev;
            la.layer_voigt = (const unsigned long long*)ctx->b_lvoigt.p;
            la.planes = (const double*)ctx->b_planes.p;
            la.lcplanes = (const double*)ctx->b_lcplanes.p;
            la.lcidx_s = ctx->ld.lcidx;
            la.nlc_pad = nlc_pad;
            la.lay = (const LayerDev*)ctx->b_layer.p;
            la.absrb = (const double*)ctx->b_absrb.p;
            la.nptabs = nptabs;
            la.nptabs_pad = nptabs_pad;
            la.cont_mask = cont_mask;
            la.v1abs = v1abs;
            la.v2abs = v2abs;
            la.v1 = v1;
            la.dvset = r.dvset;
            la.o = o_batch;
            la.o_lds = nwn;
            la.o_prof = nwn * nlay;
            la.o_by_mol = r.o_by_mol;
            la.oc = r.oc;
            la.obm_ldm = nwn;
            la.obm_ldk = nwn * MRTM_MXMOL;
            la.o_clw = r.o_clw;
            la.odxsec = r.odxsec;
            if (r.xamnt) {                  // cross sections of the batch's profiles, laid out like O
                if ((rc = ensure(ctx, ctx->b_xsod, (size_t)nb * nwn * nlay * 8))) return rc;
                for (int64_t ip = 0; ip < nb; ip++) {
                    const int64_t gp = b0 + ip;
                    if ((rc = run_xsec_device(ctx, nwn, r.wn, nlay, r.p + (size_t)gp * nlay, r.t + (size_t)gp * nlay, r.ld_xamnt,
                                              r.xamnt + (size_t)gp * r.ld_xamnt * nlay, (double*)ctx->b_xsod.p + (size_t)ip * nwn * nlay, s)))
                        return rc;
                }
                la.odxsec = (const double*)ctx->b_xsod.p;
            }
            la.sel_count = r.sel_count ? r.sel_count + (size_t)b0 * nwn * nlay : nullptr;
            la.sel_hash = r.sel_hash ? r.sel_hash + (size_t)b0 * nwn * nlay : nullptr;
            la.errflag = ctx->errflag_dev;
            const bool sel = (r.sel_count != nullptr) || (r.sel_hash != nullptr);
            // frequencies per CTA: 128 threads x F (512 on dense grids, smaller tiles for short channel lists)
            const int NTsel = 128;                         // 256-thread CTAs measured 4% slower on the dense sweep
            const int force_f = ctx->force_f;
            // The far field pays when a tile is narrow next to the line spacing: take the largest tile that the
            // frequencies fill and whose mean spectral width stays under tile_width (measured: 0.028 cm-1 tiles are
            // best on the 5.5e-5 cm-1 sweep; on the 5.5e-3 cm-1 grid of the 300-layer case 128-frequency tiles are 2.2x
            // faster than 512, on 1000 log-spaced channels 1.3x).
            int Fh = (nwn >= 2048) ? 4 : ((nwn >= 512) ? 2 : 1);
            if (r.wn_span >= 0. && nwn > 1)
                while (Fh > 1 && 128. * Fh * (r.wn_span / (double)(nwn - 1)) > ctx->tile_width) Fh >>= 1;
            const int F = (force_f == 1 || force_f == 2 || force_f == 4) ? force_f : Fh;
            // Channel lists (F == 1): a 128-channel tile is spectrally wide (cm-1), so hardly a line is far from it; with
            // tiles of coarse_tile channels (MRTM_COARSE_TILE: 32 by default, one warp per (tile, layer)) the level-0
            // expansions take most of the window.  Measured (profiles/r02_sweeps.md): 1000 log-spaced channels x 64 profiles
            // 98.5 ms (128) / 97.7 (64) / 83.9 (32); 10000 frequencies x 300 layers 27.9 / 25.2 / 25.9 ms.
            const int Tc = (F == 1 && nwn >= 64 && (ctx->coarse_tile == 64 || ctx->coarse_tile == 32)) ? ctx->coarse_tile : 0;
            const int T0 = Tc ? Tc : NTsel * F;
            const int Fc = (F == 1 && nwn >= 64 && !Tc) ? ctx->coarse_f : 1;       // 128-frequency tiles: Fc frequencies per thread, 128/Fc threads
            dim3 grid((unsigned)((nwn + T0 - 1) / T0), (unsigned)nlay, (unsigned)nb);
            CU(cudaEventRecord(ctx->ev[2], s));
            // ---- plans (layer independent) and the upper levels of the far-field hierarchy
            const int nseg_i = (int)h.segments.size();
            // per-molecule outputs need one far-field coefficient set per molecule; otherwise one set serves all
            const bool combined = (r.o_by_mol == nullptr);
            const int nslot = combined ? 1 : (int)h.slot_mol.size();
            int nlev = 1;
            int64_t ntiles[kMaxLevels], tfreq[kMaxLevels];
            ntiles[0] = grid.x;
            tfreq[0] = T0;
            if (la.ff_ratio > 0.)
                while (nlev < ctx->ff_levels && ntiles[nlev - 1] > 1) {
                    tfreq[nlev] = tfreq[nlev - 1] * ctx->ff_S;
                    ntiles[nlev] = (nwn + tfreq[nlev] - 1) / tfreq[nlev];
                    nlev++;
                }
            la.nlev = nlev;
            la.S = ctx->ff_S;
            la.nslot = nslot;
            const bool use3 = ctx->use_near3 && ctx->use_near2 && la.ff_ratio > 0. && !sel && combined && F == 4;
            // a handful of channels (sounder sets): lanes own lines, warps own frequencies -- thread-per-frequency leaves most
            // lanes idle there (measured on B200: 27 % faster on 19 channels; about even on 1000 log-spaced channels, 24 %
            // slower on a 5.5e-3 cm-1 grid, where near_kernel stays)
            const bool neart = ctx->use_neart && F == 1 && NTsel == 128 && !Tc && (ctx->use_neart > 1 || nwn < 64);
            const bool vplan = F == 1 && NTsel == 128;      // Voigt-zone candidates: every coarse-tile call
            // ---- plan cache: buffers first (a reallocation invalidates the cached plans), then the device-side check
            const size_t npc_words = 1 + std::max<size_t>(1, h.segments.size());
            if ((rc = ensure(ctx, ctx->b_pcache, (npc_words + 4) * 8))) return rc;
            for (int lv = 0; lv < nlev; lv++) {
                if ((rc = ensure(ctx, ctx->b_pieces[lv], (size_t)ntiles[lv] * std::max(nseg_i, 1) * kPiecePerSeg * sizeof(FarPiece)))) return rc;
                if ((rc = ensure(ctx, ctx->b_plan[lv], (size_t)ntiles[lv] * std::max(nseg_i, 1) * sizeof(SegWork)))) return rc;
                if ((rc = ensure(ctx, ctx->b_hdr[lv], (size_t)ntiles[lv] * sizeof(TileHdr)))) return rc;
            }
            if (la.ff_ratio > 0. && ctx->use_near2 && (rc = ensure(ctx, ctx->b_npieces, (size_t)ntiles[0] * kMaxNearPieces * sizeof(NearPiece)))) return rc;
            if (use3) {
                if ((rc = ensure(ctx, ctx->b_t3, (size_t)ntiles[0] * sizeof(Tile3)))) return rc;
                if ((rc = ensure(ctx, ctx->b_pool3, (size_t)ntiles[0] * kN3Pool * sizeof(unsigned short)))) return rc;
                if ((rc = ensure(ctx, ctx->b_segof3, (size_t)ntiles[0] * kNearCap))) return rc;
            }
            if (vplan) {
                if ((rc = ensure(ctx, ctx->b_vcand, (size_t)ntiles[0] * kVCandMax * sizeof(int)))) return rc;
                if ((rc = ensure(ctx, ctx->b_vcseg, (size_t)ntiles[0] * kVCandMax))) return rc;
                if ((rc = ensure(ctx, ctx->b_vccount, (size_t)ntiles[0] * sizeof(int)))) return rc;
            }
            unsigned long long* pc_have = (unsigned long long*)ctx->b_pcache.p;          // [npc_words] cached margins
            unsigned long long* pc_hash = pc_have + npc_words;                            // [2]
            int* pc_replan = (int*)(pc_hash + 2);
            {
                mrtm_ctx::PlanKey key;
                std::memset(&key, 0, sizeof key);
                key.nwn = nwn; key.stage_gen = ctx->stage_gen; key.F = F; key.nlev = nlev; key.S = ctx->ff_S; key.nseg = nseg_i;
                key.near2 = ctx->use_near2; key.near3 = (use3 ? 1 : 0) | (neart ? 2 : 0); key.ff_ratio = la.ff_ratio; key.ffw_ratio = ctx->ffw_ratio;
                int nb_ = 0;
                for (int lv = 0; lv < nlev; lv++) { key.bufs[nb_++] = ctx->b_pieces[lv].p; key.bufs[nb_++] = ctx->b_plan[lv].p; key.bufs[nb_++] = ctx->b_hdr[lv].p; }
                key.bufs[nb_++] = ctx->b_vcand.p; key.bufs[nb_++] = ctx->b_vccount.p;
                key.bufs[nb_++] = ctx->b_npieces.p; key.bufs[nb_++] = ctx->b_t3.p; key.bufs[nb_++] = ctx->b_pool3.p; key.bufs[nb_++] = ctx->b_pcache.p;
                const bool same = ctx->use_plan_cache && ctx->plan_key_valid && std::memcmp(&key, &ctx->plan_key, sizeof key) == 0;
                if (!same) CU(cudaMemsetAsync(ctx->b_pcache.p, 0, (npc_words + 4) * 8, sp));
                ctx->plan_key = key;
                ctx->plan_key_valid = true;
                wn_hash_kernel<<<(unsigned)std::min<int64_t>(64, (nwn + 255) / 256), 256, 0, sp>>>(r.wn, (int)nwn, pc_hash);
                PlanCheckArgs pk;
                pk.need = (const unsigned long long*)ctx->b_vtmax.p;
                pk.have = pc_have;
                pk.hash = pc_hash;
                pk.nseg = (int)std::max<size_t>(1, h.segments.size());
                pk.force = same ? 0 : 1;
                pk.replan = pc_replan;
                plan_check_kernel<<<1, 128, 0, sp>>>(pk);
                st.kernel_launches += 2;
            }
            for (int lv = nlev - 1; lv >= 0; lv--) {       // top level first: a level's far work list excludes its parent's
                if ((rc = ensure(ctx, ctx->b_pieces[lv], (size_t)ntiles[lv] * std::max(nseg_i, 1) * kPiecePerSeg * sizeof(FarPiece)))) return rc;
                if ((rc = ensure(ctx, ctx->b_plan[lv], (size_t)ntiles[lv] * std::max(nseg_i, 1) * sizeof(SegWork)))) return rc;
                if ((rc = ensure(ctx, ctx->b_hdr[lv], (size_t)ntiles[lv] * sizeof(TileHdr)))) return rc;
                if (la.ff_ratio > 0. && (rc = ensure(ctx, ctx->b_coef[lv], (size_t)ntiles[lv] * Lb * std::max(nslot, 1) * kFarK * 8))) return rc;
                PlanArgs pl;
                std::memset(&pl, 0, sizeof pl);
                pl.nwn = (int32_t)nwn;
                pl.tile_freqs = (int32_t)tfreq[lv];
                pl.nseg = nseg_i;
                pl.wn = r.wn;
                pl.seg = ctx->seg_dev;
                pl.xnu0 = ctx->ld.xnu0;
                pl.sm_max_bits = pc_have;                 // the (cached) margins, see the plan cache
                pl.vtmax_seg = pc_have + 1;
                pl.replan = pc_replan;
                pl.ff_ratio = la.ff_ratio;
                pl.out = (SegWork*)ctx->b_plan[lv].p;
                pl.hdr = (TileHdr*)ctx->b_hdr[lv].p;
                pl.pplan = (lv + 1 < nlev) ? (const SegWork*)ctx->b_plan[lv + 1].p : nullptr;
                pl.S = ctx->ff_S;
                pl.pieces = (FarPiece*)ctx->b_pieces[lv].p;
                if (lv == 0 && la.ff_ratio > 0. && ctx->use_near2 && !neart) {
                    if ((rc = ensure(ctx, ctx->b_npieces, (size_t)ntiles[0] * kMaxNearPieces * sizeof(NearPiece)))) return rc;
                    pl.near_pieces = (NearPiece*)ctx->b_npieces.p;
                    la.near_pieces = pl.near_pieces;
                }
                plan_kernel<<<(unsigned)ntiles[lv], 128, (size_t)std::max(nseg_i, 1) * (sizeof(SegWork) + kPiecePerSeg * sizeof(FarPiece)), sp>>>(pl);
                st.kernel_launches++;
                la.plan[lv] = (const SegWork*)ctx->b_plan[lv].p;
                la.hdr[lv] = (const TileHdr*)ctx->b_hdr[lv].p;
                la.coef[lv] = (la.ff_ratio > 0.) ? (const double*)ctx->b_coef[lv].p : nullptr;
            }
            la.slot_mol = ctx->ld.slot_mol;
            if (vplan) {                        // Voigt-zone candidates of the coarse tiles (layer independent, cached with the plans)
                VPlanArgs vp;
                std::memset(&vp, 0, sizeof vp);
                vp.nwn = (int32_t)nwn; vp.nseg = nseg_i; vp.tile_freqs = (int32_t)tfreq[0];
                vp.wn = r.wn; vp.seg = ctx->seg_dev; vp.xnu0 = ctx->ld.xnu0;
                vp.plan = (const SegWork*)ctx->b_plan[0].p;
                vp.sm_max_bits = pc_have; vp.vtmax_seg = pc_have + 1; vp.replan = pc_replan;
                vp.cand = (int*)ctx->b_vcand.p; vp.cand_seg = (unsigned char*)ctx->b_vcseg.p; vp.count = (int*)ctx->b_vccount.p;
                vplan_kernel<<<(unsigned)ntiles[0], 128, 0, sp>>>(vp);
                st.kernel_launches++;
                la.vcand = vp.cand; la.vcand_seg = vp.cand_seg; la.vcand_count = vp.count;
            }
            // production path (one sum over all molecules, no selection instrumentation, 512-frequency tiles): plan-driven lists
            Near3Args n3;
            std::memset(&n3, 0, sizeof n3);
            if (use3) {
                Plan3Args p3;
                std::memset(&p3, 0, sizeof p3);
                p3.nwn = (int32_t)nwn;
                p3.nseg = nseg_i;
                p3.wn = r.wn;
                p3.seg = ctx->seg_dev;
                p3.xnu0 = ctx->ld.xnu0;
                p3.deltnu = ctx->ld.deltnu;
                p3.brdidx = ctx->ld.brdidx;
                p3.max_abs_deltnu = h.max_abs_deltnu;
                p3.sm_max_bits = pc_have;
                p3.vtmax_seg = pc_have + 1;
                p3.replan = pc_replan;
                p3.hdr = (TileHdr*)ctx->b_hdr[0].p;
                p3.near_pieces = la.near_pieces;
                p3.plan = (const SegWork*)ctx->b_plan[0].p;
                p3.ffw_ratio = ctx->ffw_ratio;
                p3.out = (Tile3*)ctx->b_t3.p;
                p3.pool = (unsigned short*)ctx->b_pool3.p;
                p3.segof = (unsigned char*)ctx->b_segof3.p;
                plan3_kernel<<<(unsigned)ntiles[0], 128, 0, sp>>>(p3);
                st.kernel_launches++;
                n3.t3 = p3.out;
                n3.pool = p3.pool;
                n3.segof = p3.segof;
                const int64_t pairs = ntiles[0] * nlay * nb;
                n3.lb = ctx->near3_lb > 0 ? ctx->near3_lb : (int)std::min<int64_t>(16, std::max<int64_t>(1, pairs / 4440));
            }
            if (side_on) {
                CU(cudaEventRecord(ctx->evf[3], sp));            // plans done: the near field may start
                CU(cudaStreamWaitEvent(s, ctx->evf[3], 0));
                CU(cudaStreamWaitEvent(sp, ctx->evf[1], 0));     // derived planes ready: the far field may start
            }
            if (la.ff_ratio > 0.)
                for (int lv = nlev - 1; lv >= 0; lv--) {      // top level first: each level folds its parent's polynomial in
                    FarArgs fa;
                    std::memset(&fa, 0, sizeof fa);
                    fa.nlay = (int32_t)nlay;
                    fa.nseg = nseg_i;
                    fa.n_pad = n_pad;
                    fa.nslot = nslot;
                    fa.seg = ctx->seg_dev;
                    fa.pieces = (const FarPiece*)ctx->b_pieces[lv].p;
                    fa.hdr = la.hdr[lv];
                    if (lv + 1 < nlev) {
                        fa.phdr = la.hdr[lv + 1];
                        fa.pcoef = la.coef[lv + 1];
                    }
                    fa.S = ctx->ff_S;
                    fa.combined = combined ? 1 : 0;
                    fa.planes = la.planes;
                    fa.lcplanes = la.lcplanes;
                    fa.lcidx = la.lcidx_s;
                    fa.nlc_pad = la.nlc_pad;
                    fa.lay = la.lay;
                    fa.coef = (double*)ctx->b_coef[lv].p;
                    fa.counters = ctx->counters_dev;
                    // many small tiles (level 0 above all): one warp per (tile, layer); few large tiles: one CTA
                    const size_t far_dyn = (size_t)std::max(nseg_i, 1) * kPiecePerSeg * sizeof(FarPiece);
                    if (combined && ntiles[lv] * nlay * nb >= ctx->farw_min)
                        far_warp_kernel<<<dim3((unsigned)ntiles[lv], (unsigned)((nlay + kFarWarps - 1) / kFarWarps), (unsigned)nb), 32 * kFarWarps, far_dyn, sp>>>(fa);
                    else
                        far_kernel<<<dim3((unsigned)ntiles[lv], (unsigned)nlay, (unsigned)nb), 128, far_dyn, sp>>>(fa);
                    st.kernel_launches++;
                }
            st.kernel_launches += 2;       // voigt_kernel, final_kernel (near_kernel is counted below)
            cudaEvent_t far_done = nullptr;
            if (side_on) {
                CU(cudaEventRecord(ctx->evf[2], sp));
                far_done = ctx->evf[2];
            }
            cudaStream_t sv = nullptr;
            if (side_on && ctx->voigt_side && !r.o_by_mol) {
                const size_t ob = (size_t)nwn * nlay * nb * 8;
                if ((rc = ensure(ctx, ctx->b_ov, ob))) return rc;
                CU(cudaMemsetAsync(ctx->b_ov.p, 0, ob, sp));
                la.o_v = (double*)ctx->b_ov.p;
                sv = sp;
            }
            if (Tc == 64) launch_lines<2, 32>(la, grid, sel, s, far_done, sv, nullptr, false, 0);
            else if (Tc == 32) launch_lines<1, 32>(la, grid, sel, s, far_done, sv, nullptr, false, 0);
            else if (F == 4) launch_lines<4, 128>(la, grid, sel, s, far_done, sv, use3 ? &n3 : nullptr, false, ctx->use_voigt_t);
            else if (F == 2) launch_lines<2, 128>(la, grid, sel, s, far_done, sv, nullptr, false, ctx->use_voigt_t);
            else if (Fc == 2) launch_lines<2, 64>(la, grid, sel, s, far_done, sv, nullptr, false, 0);
            else if (Fc == 4) launch_lines<4, 32>(la, grid, sel, s, far_done, sv, nullptr, false, 0);
            else launch_lines<1, 128>(la, grid, sel, s, far_done, sv, nullptr, neart, 0);
            CU(cudaEventRecord(ctx->ev[3], s));
            st.kernel_launches++;
            CU(cudaGetLastError());
        }
        if (r.do_tmr || r.do_rtm) {
            RtArgs ra;
            std::memset(&ra, 0, sizeof ra);
            ra.nwn = (int32_t)nwn;
            ra.nlay = (int32_t)nlay;
            ra.nprof = (int32_t)nb;
            ra.irt = (int32_t)r.irt;
            ra.iout = (int32_t)r.iout;
            ra.do_tmr = r.do_tmr;
            ra.do_rtm = r.do_rtm;
            ra.wn = r.wn;
            ra.o = o_batch;
            ra.o_lds = nwn;
            ra.o_prof = nwn * nlay;
            ra.t = r.t + (size_t)b0 * nlay;
            ra.tz = r.tz + (size_t)b0 * (nlay + 1);
            ra.tmpsfc = r.tmpsfc ? r.tmpsfc + b0 : nullptr;
            ra.emiss = r.emiss;
            ra.reflc = r.reflc;
            auto off = [&](double* q) { return q ? q + (size_t)b0 * nwn : nullptr; };
            ra.rad = off(r.rad); ra.tb = off(r.tb); ra.tmr = off(r.tmr);
            ra.trtot = off(r.trtot); ra.rup = off(r.rup); ra.rdn = off(r.rdn);
            if (r.rt_ready) CU(cudaStreamWaitEvent(s, r.rt_ready, 0));
            CU(cudaEventRecord(ctx->ev[4], s));
            {   // RADCN2/T per layer and level, once per batch
                const int n_t = (int)(nb * nlay), n_tz = (int)(nb * (nlay + 1));
                if ((rc = ensure(ctx, ctx->b_fbeta, (size_t)(n_t + n_tz) * 8))) return rc;
                double* fb = (double*)ctx->b_fbeta.p;
                rt_prep_kernel<<<(unsigned)((n_tz + 127) / 128), 128, 0, s>>>(n_t, ra.t, fb, n_tz, ra.tz, fb + n_t);
                ra.fb = fb;
                ra.fbz = fb + n_t;
                st.kernel_launches++;
            }
            rt_kernel<<<dim3((unsigned)((nwn + kRtFreqs - 1) / kRtFreqs), (unsigned)nb), kRtParts * kRtFreqs, 0, s>>>(ra);
            CU(cudaEventRecord(ctx->ev[5], s));
            st.kernel_launches++;
            CU(cudaGetLastError());
            // the spectra of this batch go home at once (the host is still waiting on the events below)
            double* dev_spec[6] = {r.rad, r.tb, r.tmr, r.trtot, r.rup, r.rdn};
            for (int i = 0; i < 6; i++)
                if (r.h_spec[i] && dev_spec[i])
                    CU(cudaMemcpyAsync(r.h_spec[i] + (size_t)b0 * nwn, dev_spec[i] + (size_t)b0 * nwn, (size_t)nb * nwn * 8, cudaMemcpyDeviceToHost, s));
        }
        // per-batch kernel times (needs the batch to finish; cheap next to the kernels themselves).  A single-batch
        // asynchronous call returns here without touching the host again: mrtm_sync / mrtm_get_stats / the next call finish it
        const bool defer = r.async_ok && r.nprof <= B;
        if (defer) {
            ctx->pending = true;
            ctx->pending_lines = r.do_lines;
            ctx->pending_rt = r.do_tmr || r.do_rtm;
            return MRTM_OK;
        }
        {
            CU(cudaEventSynchronize(r.do_tmr || r.do_rtm ? ctx->ev[5] : ctx->ev[3]));
            float ms = 0.f;
            if (r.do_lines) {
                cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); st.last_derive_kernel_ms += ms;
                cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); st.last_lines_kernel_ms += ms;
            }
            if (r.do_tmr || r.do_rtm) { cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); st.last_rt_kernel_ms += ms; }
        }
    }
    if (r.do_lines) {
        int flag = 0;
        unsigned long long cnt[2] = {0ull, 0ull};
        CU(cudaMemcpyAsync(&flag, ctx->errflag_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(cnt, ctx->counters_dev, sizeof cnt, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        st.far_expansions = (double)cnt[0];
        st.direct_evals = (double)cnt[1];
        if (flag & 1) return set_err(ctx, MRTM_ETIPS, mrtm_strerror(MRTM_ETIPS));
        if (flag & 2) return set_err(ctx, MRTM_ESDVOIGT, mrtm_strerror(MRTM_ESDVOIGT));
        if (flag & (16 | 32)) return set_err(ctx, MRTM_EXSEC, mrtm_strerror(MRTM_EXSEC));
    }
    return MRTM_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer wrappers
// ---------------------------------------------------------------------------------------------
static int h2d(mrtm_ctx* ctx, DevBuf& b, const void* src, size_t bytes, cudaStream_t s, const double** out)
{
    int rc = ensure(ctx, b, std::max<size_t>(bytes, 8));
    if (rc) return rc;
    if (src && bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    *out = (const double*)b.p;
    return MRTM_OK;
}


// ---- cross sections ---------------------------------------------------------------------------
extern "C" int mrtm_stage_xsec(mrtm_ctx* ctx, int64_t nreg, const mrtm_xs_region* regs)
{
    if (!ctx || nreg < 0 || (nreg > 0 && !regs)) return MRTM_EARG;
    if (!ctx->peers.empty()) return for_each_peer(ctx, [&](int, mrtm_ctx* c) { return mrtm_stage_xsec(c, nreg, regs); });
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->xs_regs.clear();
    ctx->xs_tab_per_layer = 0;
    ctx->xs_nmol = 0;
    std::vector<double> dat;
    for (int64_t i = 0; i < nreg; i++) {
        const mrtm_xs_region& g = regs[i];
        if (g.ntemp < 1 || g.ntemp > 6 || g.npts < 2 || g.ixmol < 0 || g.ixmol >= 38 || !(g.v2x > g.v1x))
            return set_err(ctx, MRTM_EARG, "mrtm_stage_xsec: bad region (ntemp 1..6, npts >= 2, ixmol 0..37, v2x > v1x)");
        if (i > 0 && g.ixmol < regs[i - 1].ixmol) return set_err(ctx, MRTM_EARG, "mrtm_stage_xsec: regions must be ordered by molecule");
        XsRegionDev d;
        std::memset(&d, 0, sizeof d);
        d.ixmol = g.ixmol; d.ntemp = g.ntemp; d.npts = g.npts;
        d.v1fx = g.v1fx; d.v2fx = g.v2fx; d.v1x = g.v1x; d.v2x = g.v2x; d.xdoplr = g.xdoplr;
        for (int k = 0; k < g.ntemp; k++) {
            if (!g.xsdat[k]) return set_err(ctx, MRTM_EARG, "mrtm_stage_xsec: null table");
            d.tx[k] = g.tx[k]; d.pdx[k] = g.pdx[k];
            d.dat_off[k] = (int64_t)dat.size();
            dat.insert(dat.end(), g.xsdat[k], g.xsdat[k] + g.npts);
        }
        d.tab_off = 0;                                    // set per call (depends on nlay)
        ctx->xs_regs.push_back(d);
        ctx->xs_tab_per_layer += g.npts + 3;
        ctx->xs_nmol = std::max(ctx->xs_nmol, (int)g.ixmol + 1);
    }
    int rc;
    if ((rc = ensure(ctx, ctx->b_xsdat, std::max<size_t>(dat.size(), 1) * 8))) return rc;
    if (!dat.empty()) CU(cudaMemcpy(ctx->b_xsdat.p, dat.data(), dat.size() * 8, cudaMemcpyHostToDevice));
    return MRTM_OK;
}

// MONORTM_XSEC_SUB for one profile; all pointers on the device; od (nwn, nlay).  Asynchronous on s; the error flag is
// checked by the caller (bits 4, 5 of errflag_dev)
static int run_xsec_device(mrtm_ctx* ctx, int64_t nwn, const double* wn, int64_t nlay, const double* p, const double* t,
                           int64_t ld_xamnt, const double* xamnt, double* od, cudaStream_t s)
{
    const int nreg = (int)ctx->xs_regs.size();
    if (ld_xamnt < ctx->xs_nmol) return set_err(ctx, MRTM_EARG, "xamnt: leading dimension smaller than the number of staged cross-section molecules");
    if (nreg == 0) { CU(cudaMemsetAsync(od, 0, (size_t)nwn * nlay * 8, s)); return MRTM_OK; }
    int rc;
    std::vector<XsRegionDev> regs = ctx->xs_regs;
    int64_t off = 0, maxpts = 0;
    for (auto& g : regs) { g.tab_off = off; off += (g.npts + 3) * nlay; maxpts = std::max<int64_t>(maxpts, g.npts + 3); }
    if ((rc = ensure(ctx, ctx->b_xsreg, regs.size() * sizeof(XsRegionDev)))) return rc;
    if ((rc = ensure(ctx, ctx->b_xslay, (size_t)nreg * nlay * sizeof(XsLayerDev)))) return rc;
    if ((rc = ensure(ctx, ctx->b_xstab, (size_t)off * 8))) return rc;
    if ((rc = ensure(ctx, ctx->b_xsneed, (size_t)nreg * sizeof(int)))) return rc;
    CU(cudaMemcpyAsync(ctx->b_xsreg.p, regs.data(), regs.size() * sizeof(XsRegionDev), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));                         // regs is a stack copy
    CU(cudaMemsetAsync(ctx->b_xsneed.p, 0, (size_t)nreg * sizeof(int), s));
    XsArgs a;
    std::memset(&a, 0, sizeof a);
    a.nreg = nreg; a.nlay = (int32_t)nlay; a.nwn = (int32_t)nwn; a.ld_xamnt = (int32_t)ld_xamnt;
    a.reg = (const XsRegionDev*)ctx->b_xsreg.p;
    a.dat = (const double*)ctx->b_xsdat.p;
    a.lay = (XsLayerDev*)ctx->b_xslay.p;
    a.tab = (double*)ctx->b_xstab.p;
    a.need = (int*)ctx->b_xsneed.p;
    a.wn = wn; a.p = p; a.t = t; a.xamnt = xamnt; a.odxsec = od;
    a.errflag = ctx->errflag_dev;
    xs_layer_kernel<<<(unsigned)((nreg * nlay + 127) / 128), 128, 0, s>>>(a);
    xs_table_kernel<<<dim3((unsigned)((maxpts + 255) / 256), (unsigned)nlay, (unsigned)nreg), 256, 0, s>>>(a);
    xs_need_kernel<<<(unsigned)((nwn + 255) / 256), 256, 0, s>>>(a);
    xs_conv_kernel<<<dim3((unsigned)((nwn + 127) / 128), (unsigned)nlay), 128, 0, s>>>(a);
    ctx->st.kernel_launches += 4;
    CU(cudaGetLastError());
    return MRTM_OK;
}

static int check_xsec_flags(mrtm_ctx* ctx, cudaStream_t s)
{
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, ctx->errflag_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (flag & 16) return set_err(ctx, MRTM_EXSEC, "convolve: NPTS exceeds xspd_int(0:10000000) (monortm_sub.F90:1755): layer pressure too close to / below the table pressure for this region width");
    if (flag & 32) return set_err(ctx, MRTM_EXSEC, "convolve: the outward sum does not terminate (zero or NaN table around a frequency; the reference loops forever, monortm_sub.F90:1800-1821)");
    return MRTM_OK;
}

extern "C" int mrtm_xsec(mrtm_ctx* ctx, int64_t nwn, const double* wn, int64_t nlay, const double* p, const double* t,
                         int64_t ld_xamnt, const double* xamnt, double* odxsec)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    if (!ctx) return MRTM_EARG;
    if (!wn || !p || !t || !xamnt || !odxsec || nwn < 1 || nlay < 1 || ld_xamnt < 1) return set_err(ctx, MRTM_EARG, "mrtm_xsec: null input or bad dimension");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int rc;
    const double *dwn, *dp, *dt, *dx;
    if ((rc = h2d(ctx, ctx->b_xsin[0], wn, nwn * 8, s, &dwn))) return rc;
    if ((rc = h2d(ctx, ctx->b_xsin[1], p, nlay * 8, s, &dp))) return rc;
    if ((rc = h2d(ctx, ctx->b_xsin[2], t, nlay * 8, s, &dt))) return rc;
    if ((rc = h2d(ctx, ctx->b_xsin[3], xamnt, (size_t)ld_xamnt * nlay * 8, s, &dx))) return rc;
    if ((rc = ensure(ctx, ctx->b_xsod, (size_t)nwn * nlay * 8))) return rc;
    CU(cudaMemsetAsync(ctx->errflag_dev, 0, sizeof(int), s));
    if ((rc = run_xsec_device(ctx, nwn, dwn, nlay, dp, dt, ld_xamnt, dx, (double*)ctx->b_xsod.p, s))) return rc;
    if ((rc = check_xsec_flags(ctx, s))) return rc;
    CU(cudaMemcpyAsync(odxsec, ctx->b_xsod.p, (size_t)nwn * nlay * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MRTM_OK;
}

static void fill_range(RunDesc& r, const double* wn_host, const mrtm_opts* opts)
{
    r.v1 = wn_host[0];
    r.v2 = wn_host[r.nwn - 1];
    r.wn_span = std::fabs(wn_host[r.nwn - 1] - wn_host[0]);
    r.iw0 = 0;
    if (opts && opts->use_global_range) {
        r.v1 = opts->v1_global;
        r.v2 = opts->v2_global;
        r.iw0 = opts->iw0;
    }
}

extern "C" int mrtm_modm(mrtm_ctx* ctx, int64_t nwn, const double* wn, double dvset, int64_t nlay,
                         const double* p, const double* t, const double* clw,
                         double* o, double* o_by_mol, double* oc, double* o_clw, double* odxsec,
                         int64_t nmol, const double* wkl, const double* wbrodl,
                         double sclcpl, double sclhw, double y0res, const double cntnm[7],
                         int64_t ixsect, int64_t ibrd, const double* scor, const mrtm_opts* opts)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    if (!ctx) return MRTM_EARG;
    if (!wn || !p || !t || !clw || !wkl || !wbrodl || !cntnm || nwn < 1 || nlay < 1)
        return set_err(ctx, MRTM_EARG, "mrtm_modm: null input or bad dimension");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    RunDesc r;
    std::memset(&r, 0, sizeof r);
    r.nprof = 1; r.nwn = nwn; r.nlay = nlay; r.nmol = nmol; r.dvset = dvset;
    r.sclcpl = sclcpl; r.sclhw = sclhw; r.y0res = y0res; r.ibrd = ibrd;
    for (int i = 0; i < 7; i++) r.cntnm[i] = cntnm[i];
    r.do_lines = true;
    r.line_mode = opts ? opts->line_mode : 0;
    fill_range(r, wn, opts);
    int rc;
    const size_t fl = (size_t)nwn * nlay * 8, fml = (size_t)nwn * MRTM_MXMOL * nlay * 8;
    if ((rc = h2d(ctx, ctx->b_in[0], wn, nwn * 8, s, &r.wn))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[1], p, nlay * 8, s, &r.p))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[2], t, nlay * 8, s, &r.t))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[3], clw, nlay * 8, s, &r.clw))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[4], wkl, nlay * MRTM_MXMOL * 8, s, &r.wkl))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[5], wbrodl, nlay * 8, s, &r.wbrodl))) return rc;
    if (scor) { if ((rc = h2d(ctx, ctx->b_in[6], scor, (size_t)nlay * MRTM_NSCOR1 * MRTM_NSCOR2 * 8, s, &r.scor))) return rc; }
    const bool own_xs = (ixsect == 1) && opts && opts->xamnt;     // CALL MONORTM_XSEC_SUB inside MODM (modm.f90:197-198)
    if (own_xs) {
        const double* dx;
        if ((rc = h2d(ctx, ctx->b_xsin[3], opts->xamnt, (size_t)opts->ld_xamnt * nlay * 8, s, &dx))) return rc;
        if ((rc = ensure(ctx, ctx->b_in[7], fl))) return rc;
        CU(cudaMemsetAsync(ctx->errflag_dev, 0, sizeof(int), s));
        if ((rc = run_xsec_device(ctx, nwn, r.wn, nlay, r.p, r.t, opts->ld_xamnt, dx, (double*)ctx->b_in[7].p, s))) return rc;
        if ((rc = check_xsec_flags(ctx, s))) return rc;
        r.odxsec = (const double*)ctx->b_in[7].p;
    } else if (ixsect == 1 && odxsec) { if ((rc = h2d(ctx, ctx->b_in[7], odxsec, fl, s, &r.odxsec))) return rc; }
    if ((rc = ensure(ctx, ctx->b_out[0], fl))) return rc;
    r.o = (double*)ctx->b_out[0].p;
    if (o_by_mol) { if ((rc = ensure(ctx, ctx->b_obm, fml))) return rc; r.o_by_mol = (double*)ctx->b_obm.p; CU(cudaMemsetAsync(r.o_by_mol, 0, fml, s)); }
    if (oc) { if ((rc = ensure(ctx, ctx->b_oc, fml))) return rc; r.oc = (double*)ctx->b_oc.p; CU(cudaMemsetAsync(r.oc, 0, fml, s)); }
    if (o_clw) { if ((rc = ensure(ctx, ctx->b_out[1], fl))) return rc; r.o_clw = (double*)ctx->b_out[1].p; }
    if (opts && opts->sel_count) { if ((rc = ensure(ctx, ctx->b_sel[0], fl))) return rc; r.sel_count = (long long*)ctx->b_sel[0].p; }
    if (opts && opts->sel_hash) { if ((rc = ensure(ctx, ctx->b_sel[1], fl))) return rc; r.sel_hash = (unsigned long long*)ctx->b_sel[1].p; }
    if ((rc = run_device(ctx, r, s))) return rc;
    if (o) CU(cudaMemcpyAsync(o, r.o, fl, cudaMemcpyDeviceToHost, s));
    if (o_by_mol) CU(cudaMemcpyAsync(o_by_mol, r.o_by_mol, fml, cudaMemcpyDeviceToHost, s));
    if (oc) CU(cudaMemcpyAsync(oc, r.oc, fml, cudaMemcpyDeviceToHost, s));
    if (o_clw) CU(cudaMemcpyAsync(o_clw, r.o_clw, fl, cudaMemcpyDeviceToHost, s));
    if (own_xs && odxsec) CU(cudaMemcpyAsync(odxsec, r.odxsec, fl, cudaMemcpyDeviceToHost, s));
    if (r.sel_count) CU(cudaMemcpyAsync(opts->sel_count, r.sel_count, fl, cudaMemcpyDeviceToHost, s));
    if (r.sel_hash) CU(cudaMemcpyAsync(opts->sel_hash, r.sel_hash, fl, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (odxsec && ixsect != 1) std::memset(odxsec, 0, fl);    // modm.f90:193
    if (r.sel_count) {
        double tot = 0.;
        for (size_t i = 0; i < (size_t)nwn * nlay; i++) tot += (double)opts->sel_count[i];
        ctx->st.inwindow_evals = tot;
    }
    return MRTM_OK;
}

static int run_rt_host(mrtm_ctx* ctx, bool do_tmr, bool do_rtm, int64_t iout, int64_t irt, int64_t nwn,
                       const double* wn, int64_t nlay, const double* t, const double* tz, const double* o,
                       double* tmpsfc, double* rup, double* trtot, double* rdn, const double* reflc,
                       const double* emiss, double* rad, double* tb, double* tmr, int64_t idu)
{
    if (!ctx) return MRTM_EARG;
    if (!wn || !t || !tz || !o || nwn < 1 || nlay < 1) return set_err(ctx, MRTM_EARG, "rt: null input or bad dimension");
    if (do_rtm && idu != 1) return set_err(ctx, MRTM_EIDU, mrtm_strerror(MRTM_EIDU));
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    RunDesc r;
    std::memset(&r, 0, sizeof r);
    r.nprof = 1; r.nwn = nwn; r.nlay = nlay; r.irt = irt; r.iout = iout; r.idu = idu;
    r.do_tmr = do_tmr; r.do_rtm = do_rtm;
    int rc;
    const double* od = nullptr;
    if ((rc = h2d(ctx, ctx->b_in[0], wn, nwn * 8, s, &r.wn))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[2], t, nlay * 8, s, &r.t))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[8], tz, (nlay + 1) * 8, s, &r.tz))) return rc;
    if ((rc = h2d(ctx, ctx->b_out[0], o, (size_t)nwn * nlay * 8, s, &od))) return rc;
    r.o = (double*)od;
    double tsfc = 0.;
    if (do_rtm) {
        if (!tmpsfc || !reflc || !emiss) return set_err(ctx, MRTM_EARG, "rtm: null tmpsfc/reflc/emiss");
        if (irt == 3 || irt == 2) *tmpsfc = 2.75;             // RTMmono.f90:113-123
        tsfc = *tmpsfc;
        const double* q;
        if ((rc = h2d(ctx, ctx->b_tmps, &tsfc, 8, s, &q))) return rc;
        r.tmpsfc = (double*)q;
        if ((rc = h2d(ctx, ctx->b_in[9], emiss, nwn * 8, s, &r.emiss))) return rc;
        if ((rc = h2d(ctx, ctx->b_in[10], reflc, nwn * 8, s, &r.reflc))) return rc;
    }
    double** outs[6] = {&r.rad, &r.tb, &r.tmr, &r.trtot, &r.rup, &r.rdn};
    double* hosts[6] = {rad, tb, tmr, trtot, rup, rdn};
    for (int i = 0; i < 6; i++) {
        if ((rc = ensure(ctx, ctx->b_out[2 + i], nwn * 8))) return rc;
        *outs[i] = (double*)ctx->b_out[2 + i].p;
    }
    if ((rc = run_device(ctx, r, s))) return rc;
    for (int i = 0; i < 6; i++)
        if (hosts[i]) CU(cudaMemcpyAsync(hosts[i], *outs[i], nwn * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MRTM_OK;
}

extern "C" int mrtm_calctmr(mrtm_ctx* ctx, int64_t nlayrs, int64_t nwn, const double* wn,
                            const double* t, const double* tz, const double* o, double* tmr)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    return run_rt_host(ctx, true, false, 0, 3, nwn, wn, nlayrs, t, tz, o, nullptr, nullptr, nullptr, nullptr,
                       nullptr, nullptr, nullptr, nullptr, tmr, 1);
}

extern "C" int mrtm_rtm(mrtm_ctx* ctx, int64_t iout, int64_t irt, int64_t nwn, const double* wn,
                        int64_t nlay, const double* t, const double* tz, const double* o,
                        double* tmpsfc, double* rup, double* trtot, double* rdn,
                        const double* reflc, const double* emiss, double* rad, double* tb, int64_t idu)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    return run_rt_host(ctx, false, true, iout, irt, nwn, wn, nlay, t, tz, o, tmpsfc, rup, trtot, rdn, reflc,
                       emiss, rad, tb, nullptr, idu);
}

extern "C" int mrtm_profiles_dev(mrtm_ctx* ctx, int64_t nprof, int64_t nwn, const double* wn_dev, double dvset,
                                 int64_t nlay, const double* p, const double* t, const double* tz,
                                 const double* clw, int64_t nmol, const double* wkl, const double* wbrodl,
                                 const double* scor, double sclcpl, double sclhw, double y0res,
                                 const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                                 double* tmpsfc, const double* emiss_dev, const double* reflc_dev,
                                 double* rad_dev, double* tb_dev, double* tmr_dev, double* trtot_dev,
                                 double* rup_dev, double* rdn_dev, double* o_dev, const mrtm_opts* opts)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    if (!ctx) return MRTM_EARG;
    if (!wn_dev || !p || !t || !tz || !clw || !wkl || !wbrodl || !cntnm || !tmpsfc || !emiss_dev || !reflc_dev)
        return set_err(ctx, MRTM_EARG, "mrtm_profiles_dev: null input");
    if (!opts || !opts->use_global_range)
        return set_err(ctx, MRTM_EARG, "mrtm_profiles_dev: opts->use_global_range with v1_global/v2_global is required (wn lives on the device)");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = opts->stream ? (cudaStream_t)opts->stream : ctx->stream;
    RunDesc r;
    std::memset(&r, 0, sizeof r);
    r.nprof = nprof; r.nwn = nwn; r.nlay = nlay; r.nmol = nmol; r.dvset = dvset; r.wn = wn_dev;
    // span of the device-resident frequencies (tile-size heuristic only): read back once per (pointer, count)
    if (ctx->span_ptr != (const void*)wn_dev || ctx->span_nwn != nwn) {
        double ends[2] = {0., 0.};
        CU(cudaMemcpyAsync(&ends[0], wn_dev, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&ends[1], wn_dev + (nwn - 1), 8, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        ctx->span_ptr = wn_dev; ctx->span_nwn = nwn; ctx->span_val = std::fabs(ends[1] - ends[0]);
    }
    r.wn_span = ctx->span_val;
    r.p = p; r.t = t; r.tz = tz; r.clw = clw; r.wkl = wkl; r.wbrodl = wbrodl; r.scor = scor;
    r.sclcpl = sclcpl; r.sclhw = sclhw; r.y0res = y0res; r.ibrd = ibrd; r.irt = irt; r.iout = iout; r.idu = idu;
    for (int i = 0; i < 7; i++) r.cntnm[i] = cntnm[i];
    r.tmpsfc = tmpsfc; r.emiss = emiss_dev; r.reflc = reflc_dev;
    r.rad = rad_dev; r.tb = tb_dev; r.tmr = tmr_dev; r.trtot = trtot_dev; r.rup = rup_dev; r.rdn = rdn_dev;
    r.o = o_dev;
    r.do_lines = true; r.do_tmr = true; r.do_rtm = true;
    r.line_mode = opts->line_mode;
    r.v1 = opts->v1_global; r.v2 = opts->v2_global; r.iw0 = opts->iw0;
    r.async_ok = true;
    return run_device(ctx, r, s);
}

extern "C" int mrtm_sync(mrtm_ctx* ctx)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    if (!ctx) return MRTM_EARG;
    CU(cudaSetDevice(ctx->device));
    const int rc = finalize_pending(ctx);
    ctx->deferred_rc = MRTM_OK;                     // reported once
    CU(cudaStreamSynchronize(ctx->stream));
    return rc;
}

static int profiles_single(mrtm_ctx* ctx, int64_t nprof, int64_t nwn, const double* wn, double dvset,
                             int64_t nlay, const double* p, const double* t, const double* tz,
                             const double* clw, int64_t nmol, const double* wkl, const double* wbrodl,
                             const double* scor, double sclcpl, double sclhw, double y0res,
                             const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                             double* tmpsfc, const double* emiss, const double* reflc,
                             double* rad, double* tb, double* tmr, double* trtot, double* rup, double* rdn,
                             double* o, double* otot_by_mol, const mrtm_opts* opts)
{
    if (!ctx) return MRTM_EARG;
    if (!wn || !p || !t || !tz || !clw || !wkl || !wbrodl || !cntnm || !tmpsfc || !emiss || !reflc || nprof < 1 || nwn < 1 || nlay < 1)
        return set_err(ctx, MRTM_EARG, "mrtm_profiles: null input or bad dimension");
    if (idu != 1) return set_err(ctx, MRTM_EIDU, mrtm_strerror(MRTM_EIDU));
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    RunDesc r;
    std::memset(&r, 0, sizeof r);
    r.nprof = nprof; r.nwn = nwn; r.nlay = nlay; r.nmol = nmol; r.dvset = dvset;
    r.sclcpl = sclcpl; r.sclhw = sclhw; r.y0res = y0res; r.ibrd = ibrd; r.irt = irt; r.iout = iout; r.idu = idu;
    for (int i = 0; i < 7; i++) r.cntnm[i] = cntnm[i];
    r.do_lines = true; r.do_tmr = true; r.do_rtm = true;
    r.line_mode = opts ? opts->line_mode : 0;
    fill_range(r, wn, opts);
    int rc;
    const size_t L = (size_t)nprof * nlay;
    if ((rc = h2d(ctx, ctx->b_in[0], wn, nwn * 8, s, &r.wn))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[1], p, L * 8, s, &r.p))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[2], t, L * 8, s, &r.t))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[3], clw, L * 8, s, &r.clw))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[4], wkl, L * MRTM_MXMOL * 8, s, &r.wkl))) return rc;
    if ((rc = h2d(ctx, ctx->b_in[5], wbrodl, L * 8, s, &r.wbrodl))) return rc;
    if (scor) { if ((rc = h2d(ctx, ctx->b_in[6], scor, L * MRTM_NSCOR1 * MRTM_NSCOR2 * 8, s, &r.scor))) return rc; }
    if ((rc = h2d(ctx, ctx->b_in[8], tz, (size_t)nprof * (nlay + 1) * 8, s, &r.tz))) return rc;
    if (opts && opts->xamnt) {            // IXSECT=1: MONORTM_XSEC_SUB per profile inside (modm.f90:197-198)
        if (opts->ld_xamnt < 1) return set_err(ctx, MRTM_EARG, "mrtm_opts.ld_xamnt must be >= 1");
        if ((rc = h2d(ctx, ctx->b_xsin[3], opts->xamnt, (size_t)opts->ld_xamnt * L * 8, s, &r.xamnt))) return rc;
        r.ld_xamnt = opts->ld_xamnt;
    }
    {   // emissivity / reflectivity are only read by rt_kernel: their copies ride the second stream beside the kernels
        cudaStream_t sc = (ctx->use_side && ctx->side && ctx->ev_in) ? ctx->side : s;
        if ((rc = h2d(ctx, ctx->b_in[9], emiss, nwn * 8, sc, &r.emiss))) return rc;
        if ((rc = h2d(ctx, ctx->b_in[10], reflc, nwn * 8, sc, &r.reflc))) return rc;
        if (sc != s) { CU(cudaEventRecord(ctx->ev_in, sc)); r.rt_ready = ctx->ev_in; }
    }
    if (irt == 3 || irt == 2) for (int64_t i = 0; i < nprof; i++) tmpsfc[i] = 2.75;   // RTMmono.f90:113-123
    { const double* q; if ((rc = h2d(ctx, ctx->b_tmps, tmpsfc, nprof * 8, s, &q))) return rc; r.tmpsfc = (double*)q; }
    double** outs[6] = {&r.rad, &r.tb, &r.tmr, &r.trtot, &r.rup, &r.rdn};
    double* hosts[6] = {rad, tb, tmr, trtot, rup, rdn};
    const size_t fo = (size_t)nwn * nprof * 8;
    for (int i = 0; i < 6; i++) {
        if ((rc = ensure(ctx, ctx->b_out[2 + i], fo))) return rc;
        *outs[i] = (double*)ctx->b_out[2 + i].p;
    }
    const size_t fl = (size_t)nwn * nlay * nprof * 8;
    if (o) { if ((rc = ensure(ctx, ctx->b_out[0], fl))) return rc; r.o = (double*)ctx->b_out[0].p; }
    if (opts && opts->sel_count) { if ((rc = ensure(ctx, ctx->b_sel[0], fl))) return rc; r.sel_count = (long long*)ctx->b_sel[0].p; }
    if (opts && opts->sel_hash) { if ((rc = ensure(ctx, ctx->b_sel[1], fl))) return rc; r.sel_hash = (unsigned long long*)ctx->b_sel[1].p; }

    if (!otot_by_mol) {
        for (int i = 0; i < 6; i++) r.h_spec[i] = hosts[i];
        if ((rc = run_device(ctx, r, s))) return rc;
        for (int i = 0; i < 6; i++) hosts[i] = nullptr;          // already on their way
    } else {
        // per-molecule column sums need the (nwn,39,nlay) planes: one profile at a time
        const size_t fml = (size_t)nwn * MRTM_MXMOL * nlay * 8;
        if ((rc = ensure(ctx, ctx->b_obm, fml))) return rc;
        if ((rc = ensure(ctx, ctx->b_oc, fml))) return rc;
        if ((rc = ensure(ctx, ctx->b_out[1], (size_t)MRTM_MXMOL * nwn * nprof * 8))) return rc;
        double lines_ms = 0., rt_ms = 0., der_ms = 0.;
        for (int64_t ip = 0; ip < nprof; ip++) {
            RunDesc q = r;
            q.nprof = 1;
            q.p += ip * nlay; q.t += ip * nlay; q.clw += ip * nlay; q.wbrodl += ip * nlay;
            q.wkl += ip * nlay * MRTM_MXMOL; q.tz += ip * (nlay + 1);
            if (q.scor) q.scor += (size_t)ip * nlay * MRTM_NSCOR1 * MRTM_NSCOR2;
            q.tmpsfc += ip;
            q.rad += ip * nwn; q.tb += ip * nwn; q.tmr += ip * nwn; q.trtot += ip * nwn; q.rup += ip * nwn; q.rdn += ip * nwn;
            if (q.o) q.o += (size_t)ip * nwn * nlay;
            if (q.sel_count) q.sel_count += (size_t)ip * nwn * nlay;
            if (q.sel_hash) q.sel_hash += (size_t)ip * nwn * nlay;
            q.o_by_mol = (double*)ctx->b_obm.p;
            q.oc = (double*)ctx->b_oc.p;
            CU(cudaMemsetAsync(q.o_by_mol, 0, fml, s));
            CU(cudaMemsetAsync(q.oc, 0, fml, s));
            if ((rc = run_device(ctx, q, s))) return rc;
            lines_ms += ctx->st.last_lines_kernel_ms; rt_ms += ctx->st.last_rt_kernel_ms; der_ms += ctx->st.last_derive_kernel_ms;
            colsum_kernel<<<dim3((unsigned)((nwn + 127) / 128), MRTM_MXMOL), 128, 0, s>>>(
                (int)nwn, (int)nlay, q.o_by_mol, q.oc, nwn, nwn * MRTM_MXMOL,
                (double*)ctx->b_out[1].p + (size_t)ip * MRTM_MXMOL * nwn);
            ctx->st.kernel_launches++;
        }
        ctx->st.last_lines_kernel_ms = lines_ms; ctx->st.last_rt_kernel_ms = rt_ms; ctx->st.last_derive_kernel_ms = der_ms;
        ctx->st.nominal_evals = (double)ctx->hl.n * (double)nlay * (double)nwn * (double)nprof;
        CU(cudaMemcpyAsync(otot_by_mol, ctx->b_out[1].p, (size_t)MRTM_MXMOL * nwn * nprof * 8, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 6; i++)
        if (hosts[i]) CU(cudaMemcpyAsync(hosts[i], *outs[i], fo, cudaMemcpyDeviceToHost, s));
    if (o) CU(cudaMemcpyAsync(o, r.o, fl, cudaMemcpyDeviceToHost, s));
    if (r.sel_count) CU(cudaMemcpyAsync(opts->sel_count, r.sel_count, fl, cudaMemcpyDeviceToHost, s));
    if (r.sel_hash) CU(cudaMemcpyAsync(opts->sel_hash, r.sel_hash, fl, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (r.sel_count) {
        double tot = 0.;
        for (size_t i = 0; i < (size_t)nwn * nlay * nprof; i++) tot += (double)opts->sel_count[i];
        ctx->st.inwindow_evals = tot;
    }
    return MRTM_OK;
}

// ---------------------------------------------------------------------------------------------
// several GPUs behind one context: one host thread per device, no data-path collective
// ---------------------------------------------------------------------------------------------
extern "C" int mrtm_init_multi(uint64_t device_mask, mrtm_ctx** out)
{
    if (!out) return MRTM_EARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return set_err(nullptr, MRTM_ENODEV, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    if (device_mask == 0) device_mask = (ndev >= 64) ? ~0ull : ((1ull << ndev) - 1ull);
    mrtm_ctx* ctx = new mrtm_ctx();
    ctx->device = -1;
    std::memset(&ctx->st, 0, sizeof ctx->st);
    // MRTM_MULTI_CONTEXTS_PER_DEVICE=n: n contexts on every selected device (exercises the partitioning on a single-GPU box)
    int dup = 1;
    if (const char* sdup = std::getenv("MRTM_MULTI_CONTEXTS_PER_DEVICE")) dup = std::min(std::max(std::atoi(sdup), 1), 8);
    for (int dd = 0; dd < 64 * dup; dd++) {
        const int d = dd / dup;
        if (!((device_mask >> d) & 1ull)) continue;
        mrtm_ctx* c = nullptr;
        int rc = (d < ndev) ? mrtm_init(d, &c) : set_err(nullptr, MRTM_EARG, "device_mask names a device that does not exist");
        if (rc) {
            if (c) mrtm_free(c);
            for (mrtm_ctx* q : ctx->peers) mrtm_free(q);
            delete ctx;
            return rc;
        }
        ctx->peers.push_back(c);
    }
    *out = ctx;
    return MRTM_OK;
}

extern "C" int mrtm_num_devices(mrtm_ctx* ctx) { return !ctx ? 0 : (ctx->peers.empty() ? 1 : (int)ctx->peers.size()); }

// frequency partition for the next call: equal counts the first time a grid is seen, afterwards each device's share is
// rescaled by the time it needed for its last share (the cost per frequency depends on the spectral position: more
// expansions where negative-frequency resonances and window edges fall inside the coarse tiles), blocks in multiples of
// 512 frequencies (the tile of the dense-grid kernels)
static void mg_partition(mrtm_ctx* ctx, int64_t nwn, int ndev)
{
    const int64_t q = 512;
    std::vector<double> share((size_t)ndev, 1.0 / ndev);
    const bool have = ctx->mg_nwn == nwn && (int)ctx->mg_bounds.size() == ndev + 1 && (int)ctx->mg_ms.size() == ndev;
    if (have) {
        double tmin = 1e300, tmax = 0., sum = 0.;
        for (int i = 0; i < ndev; i++) { tmin = std::min(tmin, ctx->mg_ms[(size_t)i]); tmax = std::max(tmax, ctx->mg_ms[(size_t)i]); }
        if (tmin > 0. && tmax / tmin < 1.02) return;                   // balanced: keep the partition (and the cached plans)
        for (int i = 0; i < ndev; i++) {
            const double cnt = (double)(ctx->mg_bounds[(size_t)i + 1] - ctx->mg_bounds[(size_t)i]);
            share[(size_t)i] = (ctx->mg_ms[(size_t)i] > 0. && cnt > 0.) ? cnt / ctx->mg_ms[(size_t)i] : 1.0;   // frequencies per ms
            sum += share[(size_t)i];
        }
        for (auto& v : share) v /= sum;
    }
    if (have) {                                     // damped: half way to the shares the last timings suggest
        for (int i = 0; i < ndev; i++) {
            const double old_share = (double)(ctx->mg_bounds[(size_t)i + 1] - ctx->mg_bounds[(size_t)i]) / (double)nwn;
            share[(size_t)i] = 0.5 * (share[(size_t)i] + old_share);
        }
    }
    const int64_t min_blk = (nwn >= (int64_t)ndev * 4 * q) ? 2 * q : 0;       // every device keeps a share
    ctx->mg_bounds.assign((size_t)ndev + 1, 0);
    double acc = 0.;
    for (int i = 1; i < ndev; i++) {
        acc += share[(size_t)i - 1];
        int64_t b = (int64_t)std::llround(acc * (double)nwn / (double)q) * q;
        b = std::max(b, ctx->mg_bounds[(size_t)i - 1] + min_blk);
        b = std::min(b, nwn - (int64_t)(ndev - i) * min_blk);
        ctx->mg_bounds[(size_t)i] = std::min(std::max<int64_t>(b, 0), nwn);
    }
    ctx->mg_bounds[(size_t)ndev] = nwn;
    ctx->mg_nwn = nwn;
}

static int profiles_multi(mrtm_ctx* ctx, int64_t nprof, int64_t nwn, const double* wn, double dvset,
                          int64_t nlay, const double* p, const double* t, const double* tz,
                          const double* clw, int64_t nmol, const double* wkl, const double* wbrodl,
                          const double* scor, double sclcpl, double sclhw, double y0res,
                          const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                          double* tmpsfc, const double* emiss, const double* reflc,
                          double* rad, double* tb, double* tmr, double* trtot, double* rup, double* rdn,
                          double* o, double* otot_by_mol, const mrtm_opts* opts)
{
    if (!wn || !p || !t || !tz || !clw || !wkl || !wbrodl || !cntnm || !tmpsfc || !emiss || !reflc || nprof < 1 || nwn < 1 || nlay < 1)
        return set_err(ctx, MRTM_EARG, "mrtm_profiles: null input or bad dimension");
    const int ndev = (int)ctx->peers.size();
    mrtm_opts base;
    std::memset(&base, 0, sizeof base);
    if (opts) base = *opts;
    base.stream = nullptr;
    const size_t NL = (size_t)nlay, NW = (size_t)nwn;
    auto offd = [](double* q, size_t n) { return q ? q + n : nullptr; };
    auto offc = [](const double* q, size_t n) { return q ? q + n : nullptr; };
    if (nprof >= ndev) {
        // ---- profiles are independent (monortm.f90:357): contiguous blocks of profiles, nothing to reassemble
        ctx->mg_mode = 1;
        return for_each_peer(ctx, [&](int i, mrtm_ctx* c) {
            const int64_t b = nprof / ndev, rem = nprof % ndev;
            const int64_t s0 = (int64_t)i * b + std::min<int64_t>(i, rem), cnt = b + (i < rem ? 1 : 0);
            if (cnt == 0) return (int)MRTM_OK;
            const size_t S = (size_t)s0;
            mrtm_opts oi = base;
            if (oi.sel_count) oi.sel_count += S * NW * NL;
            if (oi.sel_hash) oi.sel_hash += S * NW * NL;
            if (oi.xamnt) oi.xamnt += S * (size_t)oi.ld_xamnt * NL;
            return profiles_single(c, cnt, nwn, wn, dvset, nlay, p + S * NL, t + S * NL, tz + S * (NL + 1), clw + S * NL, nmol,
                                   wkl + S * NL * MRTM_MXMOL, wbrodl + S * NL, offc(scor, S * NL * MRTM_NSCOR1 * MRTM_NSCOR2), sclcpl, sclhw,
                                   y0res, cntnm, ibrd, irt, iout, idu, tmpsfc + S, emiss, reflc, offd(rad, S * NW), offd(tb, S * NW),
                                   offd(tmr, S * NW), offd(trtot, S * NW), offd(rup, S * NW), offd(rdn, S * NW), offd(o, S * NW * NL),
                                   offd(otot_by_mol, S * NW * MRTM_MXMOL), &oi);
        });
    }
    // ---- fewer profiles than devices: every frequency is independent (modm.f90:253, RTMmono.f90:177,286): contiguous
    // blocks of the frequency list per device; v1, v2 and the grid origin stay those of the whole list
    ctx->mg_mode = 2;
    const int use = (int)std::max<int64_t>(1, std::min<int64_t>(ndev, nwn / 2048));       // tiny lists: fewer devices
    mg_partition(ctx, nwn, use);
    const double v1g = base.use_global_range ? base.v1_global : wn[0], v2g = base.use_global_range ? base.v2_global : wn[nwn - 1];
    const int64_t iw0g = base.use_global_range ? base.iw0 : 0;
    std::vector<double> ms((size_t)use, 0.);
    for (int64_t ip = 0; ip < nprof; ip++) {
        const size_t P = (size_t)ip;
        std::vector<double> ts((size_t)use, tmpsfc[ip]);
        int rc = for_each_peer(ctx, [&](int i, mrtm_ctx* c) {
            if (i >= use) return (int)MRTM_OK;
            const int64_t f0 = ctx->mg_bounds[(size_t)i], cnt = ctx->mg_bounds[(size_t)i + 1] - f0;
            if (cnt <= 0) return (int)MRTM_OK;
            const auto t_beg = std::chrono::steady_clock::now();
            const size_t F0 = (size_t)f0, C = (size_t)cnt;
            mrtm_opts oi = base;
            oi.use_global_range = 1;
            oi.v1_global = v1g; oi.v2_global = v2g; oi.iw0 = iw0g + f0;
            // (frequency, layer) arrays are strided by the full list: computed into a block of their own, scattered below
            std::vector<double> o_blk(o ? C * NL : 0);
            std::vector<int64_t> sc_blk(base.sel_count ? C * NL : 0);
            std::vector<uint64_t> sh_blk(base.sel_hash ? C * NL : 0);
            oi.sel_count = base.sel_count ? sc_blk.data() : nullptr;
            oi.sel_hash = base.sel_hash ? sh_blk.data() : nullptr;
            if (oi.xamnt) oi.xamnt += P * (size_t)oi.ld_xamnt * NL;
            int r1 = profiles_single(c, 1, cnt, wn + F0, dvset, nlay, p + P * NL, t + P * NL, tz + P * (NL + 1), clw + P * NL, nmol,
                                     wkl + P * NL * MRTM_MXMOL, wbrodl + P * NL, offc(scor, P * NL * MRTM_NSCOR1 * MRTM_NSCOR2), sclcpl, sclhw,
                                     y0res, cntnm, ibrd, irt, iout, idu, &ts[(size_t)i], emiss + F0, reflc + F0, offd(rad, P * NW + F0),
                                     offd(tb, P * NW + F0), offd(tmr, P * NW + F0), offd(trtot, P * NW + F0), offd(rup, P * NW + F0),
                                     offd(rdn, P * NW + F0), o ? o_blk.data() : nullptr,
                                     offd(otot_by_mol, P * NW * MRTM_MXMOL + F0 * MRTM_MXMOL), &oi);
            if (r1) return r1;
            for (size_t k = 0; k < NL; k++) {
                if (o) std::memcpy(o + P * NW * NL + k * NW + F0, o_blk.data() + k * C, C * 8);
                if (base.sel_count) std::memcpy(base.sel_count + P * NW * NL + k * NW + F0, sc_blk.data() + k * C, C * 8);
                if (base.sel_hash) std::memcpy(base.sel_hash + P * NW * NL + k * NW + F0, sh_blk.data() + k * C, C * 8);
            }
            ms[(size_t)i] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_beg).count();
            return (int)MRTM_OK;
        });
        if (rc) return rc;
        for (int i = 0; i < use; i++)                                  // RTM leaves 2.75 behind for IRT 2, 3 (RTMmono.f90:122)
            if (ctx->mg_bounds[(size_t)i + 1] > ctx->mg_bounds[(size_t)i]) { tmpsfc[ip] = ts[(size_t)i]; break; }
    }
    ctx->mg_ms = ms;
    return MRTM_OK;
}

extern "C" int mrtm_profiles(mrtm_ctx* ctx, int64_t nprof, int64_t nwn, const double* wn, double dvset,
                             int64_t nlay, const double* p, const double* t, const double* tz,
                             const double* clw, int64_t nmol, const double* wkl, const double* wbrodl,
                             const double* scor, double sclcpl, double sclhw, double y0res,
                             const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                             double* tmpsfc, const double* emiss, const double* reflc,
                             double* rad, double* tb, double* tmr, double* trtot, double* rup, double* rdn,
                             double* o, double* otot_by_mol, const mrtm_opts* opts)
{
    if (!ctx) return MRTM_EARG;
    if (idu != 1) return set_err(ctx, MRTM_EIDU, mrtm_strerror(MRTM_EIDU));
    if (!ctx->peers.empty())
        return profiles_multi(ctx, nprof, nwn, wn, dvset, nlay, p, t, tz, clw, nmol, wkl, wbrodl, scor, sclcpl, sclhw, y0res, cntnm, ibrd,
                              irt, iout, idu, tmpsfc, emiss, reflc, rad, tb, tmr, trtot, rup, rdn, o, otot_by_mol, opts);
    return profiles_single(ctx, nprof, nwn, wn, dvset, nlay, p, t, tz, clw, nmol, wkl, wbrodl, scor, sclcpl, sclhw, y0res, cntnm, ibrd,
                           irt, iout, idu, tmpsfc, emiss, reflc, rad, tb, tmr, trtot, rup, rdn, o, otot_by_mol, opts);
}

extern "C" int mrtm_get_stats(mrtm_ctx* ctx, mrtm_stats* st)
{
    if (!ctx || !st) return MRTM_EARG;
    if (!ctx->peers.empty()) {          // counts add up, times are the slowest device's
        std::memset(st, 0, sizeof *st);
        int rc = MRTM_OK;
        for (mrtm_ctx* c : ctx->peers) {
            mrtm_stats q;
            const int r1 = mrtm_get_stats(c, &q);
            if (r1 && !rc) rc = r1;
            st->kernel_launches += q.kernel_launches;
            st->lines_staged = q.lines_staged;
            st->nominal_evals += q.nominal_evals;
            st->far_expansions += q.far_expansions;
            st->direct_evals += q.direct_evals;
            st->inwindow_evals = (q.inwindow_evals < 0 || st->inwindow_evals < 0) ? -1. : st->inwindow_evals + q.inwindow_evals;
            st->last_lines_kernel_ms = std::max(st->last_lines_kernel_ms, q.last_lines_kernel_ms);
            st->last_rt_kernel_ms = std::max(st->last_rt_kernel_ms, q.last_rt_kernel_ms);
            st->last_derive_kernel_ms = std::max(st->last_derive_kernel_ms, q.last_derive_kernel_ms);
            st->last_prep_ms = std::max(st->last_prep_ms, q.last_prep_ms);
        }
        return rc;
    }
    cudaSetDevice(ctx->device);
    const int rc = finalize_pending(ctx);
    ctx->deferred_rc = MRTM_OK;                     // reported once
    *st = ctx->st;
    return rc;
}

extern "C" int mrtm_reset_stats(mrtm_ctx* ctx)
{
    if (!ctx) return MRTM_EARG;
    if (!ctx->peers.empty()) { for (mrtm_ctx* c : ctx->peers) mrtm_reset_stats(c); return MRTM_OK; }
    int64_t keep = ctx->st.lines_staged;
    std::memset(&ctx->st, 0, sizeof ctx->st);
    ctx->st.lines_staged = keep;
    return MRTM_OK;
}

extern "C" int mrtm_fp64_peak(mrtm_ctx* ctx, double* tflops)
{
    if (ctx && !ctx->peers.empty()) ctx = ctx->peers[0];      // single-device entry point: the first GPU of a multi-GPU context
    if (!ctx || !tflops) return MRTM_EARG;
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    DevBuf& b = ctx->b_out[15];
    int rc = ensure(ctx, b, (size_t)blocks * threads * 8);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    fp64_peak_kernel<<<blocks, threads, 0, s>>>((double*)b.p, 1000);   // warm-up
    double best = 0.;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(ctx->ev[6], s));
        fp64_peak_kernel<<<blocks, threads, 0, s>>>((double*)b.p, iters);
        CU(cudaEventRecord(ctx->ev[7], s));
        CU(cudaEventSynchronize(ctx->ev[7]));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
        double fl = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
        ctx->st.kernel_launches++;
    }
    *tflops = best;
    return MRTM_OK;
}
