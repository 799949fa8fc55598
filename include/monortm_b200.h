/*
 * monortm_b200.h -- C ABI of libmonortm_b200.so, the B200 (sm_100a) replacement
 * for monoRTM's monochromatic optical-depth + radiance hot path.
 *
 * Drop-in boundary (all citations relative to /root/reference):
 *   mrtm_stage_lines  <- the module data GET_LNFL fills (src/lnfl_mod.f90:9-18, 22-133)
 *   mrtm_modm         <- SUBROUTINE MODM    (src/modm.f90:21-25, call src/monortm.f90:557-561)
 *   mrtm_calctmr      <- SUBROUTINE CALCTMR (src/RTMmono.f90:239,  call src/monortm.f90:567)
 *   mrtm_rtm          <- SUBROUTINE RTM     (src/RTMmono.f90:13-14, call src/monortm.f90:573-574)
 *   mrtm_profile(s)   <- the three calls fused, optical depths kept in HBM
 *
 * Conventions: every array is Fortran column-major (first index fastest),
 * REAL == double, INTEGER == int64_t (the parity build linuxGNUdbl uses
 * -fdefault-real-8 -fdefault-integer-8, build/makefile.common:195-198), except
 * brd_mol_flg which the reference itself declares integer*4.  Pointers are
 * caller-owned host memory unless the name ends in _dev; the library never
 * frees or retains them after the call returns.  Every function returns 0 on
 * success or an MRTM_E* code (the reference STOPs; the Fortran shim turns a
 * non-zero code into STOP, monortm_b200/shim/monortm_gpu_shim.f90).  One ctx
 * per process/GPU; calls on a ctx are not re-entrant.  There is no CPU
 * fallback: without a usable sm_100 device mrtm_init fails with MRTM_ENODEV.
 */
#ifndef MONORTM_B200_H
#define MONORTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRTM_MXMOL 39     /* src/lblparams.f90:28 */
#define MRTM_MXBRDMOL 7   /* src/struct_types.f90:25 */
#define MRTM_NSCOR1 42    /* scor(42,9), src/modm.f90:142 */
#define MRTM_NSCOR2 9

enum {
    MRTM_OK = 0,
    MRTM_ENODEV = 1,       /* no CUDA device / not sm_100 */
    MRTM_ECUDA = 2,        /* CUDA runtime error (see mrtm_last_error) */
    MRTM_EARG = 3,         /* bad argument / dimension */
    MRTM_ENOLINES = 4,     /* mrtm_stage_lines has not been called */
    MRTM_ELINEFILE = 5,    /* malformed line store (LC flag not 1/3/5: lnfl_mod.f90:61-63; bad isotope) */
    MRTM_ERANGE = 6,       /* reserved (round 1: V2 >= 820 cm-1 was refused; every MT_CKD branch is built now) */
    MRTM_ESDVOIGT = 7,     /* REAL(v) < 0 in SDVOIGT: modm.f90:1062 STOP */
    MRTM_EIDU = 8,         /* IDU != 1: RTMmono.f90:173 STOP */
    MRTM_ENOMEM = 9,
    MRTM_EIO = 10,         /* host helpers: file missing / malformed */
    MRTM_ETIPS = 11,       /* partition sum <= 0 or T outside 70..3000 K: tips_2003.f90:271-277 */
    MRTM_EXSEC = 12        /* cross sections: resampled grid beyond xspd_int(0:10000000) (monortm_sub.F90:1755) or a
                              convolution that cannot terminate (the reference overruns the array / loops forever) */
};

typedef struct mrtm_ctx mrtm_ctx;

/* Per-call options that have no counterpart in the reference's argument lists.
 * Zero-initialise for reference behaviour. */
typedef struct mrtm_opts {
    /* Frequency sharding: the reference derives the line-load range,
     * the continuum grid and the gridded-interpolation origin from wn(1), wn(nwn)
     * (modm.f90:180-185, 218-219).  A rank that holds only wn(iw0+1 : iw0+nwn) of a
     * global grid passes the global v1, v2 and its 0-based offset iw0 here.
     * use_global_range == 0 -> v1 = wn[0], v2 = wn[nwn-1], iw0 = 0. */
    int32_t use_global_range;
    /* Line-shape evaluation mode.  0 (default): lines whose poles are far from a frequency tile are
     * summed through a 14-term Taylor expansion about the tile centre (truncation <= ~1.3e-11 of the
     * line's own contribution), everything else -- near lines, window edges, the Voigt zone -- is
     * evaluated per (line, frequency) like the reference does.  1: every in-window triple is
     * evaluated directly (no expansion); same results to rounding, used for verification. */
    int32_t line_mode;
    double v1_global, v2_global;
    int64_t iw0;
    /* Verification: when non-NULL (host, (nwn,nlay)) the line kernel also returns, per
     * (frequency, layer), the number of (molecule,record) pairs that pass the cutoff
     * test modm.f90:384 and an order-independent 64-bit hash of them. */
    int64_t *sel_count;
    uint64_t *sel_hash;
    /* Asynchronous device-resident mode: CUDA stream (cudaStream_t) to run on; NULL =
     * the context's own stream. */
    void *stream;
    /* Cross sections (IXSECT=1): XAMNT(ld_xamnt, nlay) of COMMON /PATHX/ (src/monortm.f90:233), molecules in the order the
     * regions were staged with (mrtm_stage_xsec).  With ixsect == 1 and xamnt non-NULL mrtm_modm evaluates
     * MONORTM_XSEC_SUB itself, as MODM does (modm.f90:197-198), and returns it in odxsec; with xamnt == NULL the caller's
     * odxsec is the input (round-1 behaviour). */
    const double *xamnt;
    int64_t ld_xamnt;
} mrtm_opts;

/* ---- context ------------------------------------------------------------------------- */
int mrtm_init(int device, mrtm_ctx **ctx);
/* One context that drives several GPUs of the node from ONE host process (the reference is a single process,
 * src/monortm.f90; SURVEY 8b proposed mrtm_init(device_mask, ...)).  Bit d of device_mask selects CUDA device d; 0 = every
 * visible device.  mrtm_stage_lines / mrtm_stage_xsec replicate the staged data on every GPU; mrtm_profiles splits its work
 * (SURVEY 8e): nprof >= number of GPUs -> contiguous blocks of profiles (monortm.f90:357), otherwise contiguous blocks of
 * the frequency list (modm.f90:253; v1, v2 and the grid origin stay those of the whole list), with block sizes that follow
 * the time each GPU needed for its share of the previous call on the same list.  One host thread per GPU, no collective:
 * every result lands directly in the caller's arrays.  The other entry points (mrtm_modm, mrtm_calctmr, mrtm_rtm,
 * mrtm_xsec, mrtm_profiles_dev) run on the first GPU of the context. */
int mrtm_init_multi(uint64_t device_mask, mrtm_ctx **ctx);
int mrtm_num_devices(mrtm_ctx *ctx);
int mrtm_free(mrtm_ctx *ctx);
const char *mrtm_strerror(int code);
const char *mrtm_last_error(mrtm_ctx *ctx);   /* detail of the last failure on ctx (may be NULL ctx) */
const char *mrtm_version(void);

/* ---- line store ---------------------------------------------------------------------- */
/* Stage the line list once as a structure-of-arrays in HBM.  Arguments are the module
 * arrays of src/lnfl_mod.f90:9-13 exactly as GET_LNFL leaves them: nblm(39); iso and the
 * REAL arrays dimensioned (39, iim) [element (mo,ii) at (mo-1)+(ii-1)*39]; brd_mol_*
 * dimensioned (7,7,iim).  The H2O self-width fix-up of modm.f90:841 is applied to the
 * staged copy (the caller's alps is not modified). */
int mrtm_stage_lines(mrtm_ctx *ctx, const int64_t nblm[MRTM_MXMOL], int64_t iim,
                     const int64_t *iso, const double *xnu0, const double *deltnu,
                     const double *e, const double *alps, const double *alpf,
                     const double *x, const double *xg, const double *s0,
                     const double *rmol, const double *sdep,
                     const int32_t *brd_mol_flg, const double *brd_mol_tmp,
                     const double *brd_mol_hw, const double *brd_mol_shft);
/* number of logical lines staged (coupling-coefficient records are not lines) */
int64_t mrtm_num_lines(mrtm_ctx *ctx);

/* ---- MODM ---------------------------------------------------------------------------- */
/* Replaces CALL MODM(IPR,ICP,NWN,WN,dvset,NLAY,P,T,CLW,O,O_BY_MOL,OC,O_CLW,ODXSEC,NMOL,
 * WKL,WBRODL,SCLCPL,SCLHW,Y0RES,HFILE,cntnmScaleFac,ixsect,IBRD)  (src/modm.f90:21-25).
 * Differences: IPR/ICP/HFILE are dropped (ICP is never consulted; the line
 * file is staged by mrtm_stage_lines); tips_2003 (modm.f90:250) stays on the Fortran side
 * and its result crosses as scor(42,9,nlay); odxsec is IN/OUT: with ixsect==1 the caller
 * passes monortm_xsec_sub's result in it (modm.f90:197-198), otherwise it is zeroed.
 *   wn(nwn) ascending; p,t,clw,wbrodl(nlay); wkl(39,nlay) molec/cm2;
 *   o,o_clw,odxsec (nwn,nlay); o_by_mol,oc (nwn,39,nlay).  Any output may be NULL.
 *   cntnm[7] = xself,xfrgn,xco2c,xo3cn,xo2cn,xn2cn,xrayl (src/CntnmFactors.f90:17-19). */
int mrtm_modm(mrtm_ctx *ctx, int64_t nwn, const double *wn, double dvset, int64_t nlay,
              const double *p, const double *t, const double *clw,
              double *o, double *o_by_mol, double *oc, double *o_clw, double *odxsec,
              int64_t nmol, const double *wkl, const double *wbrodl,
              double sclcpl, double sclhw, double y0res, const double cntnm[7],
              int64_t ixsect, int64_t ibrd, const double *scor, const mrtm_opts *opts);

/* ---- cross sections --------------------------------------------------------- */
/* One spectral region of one cross-section molecule, as MONORTM_XSEC_SUB holds it after the READs of loop 3000
 * (src/monortm_sub.F90:1656-1671): FSCDXS range (XSREAD, :1380-1386), header of the LAST temperature file read, the
 * per-temperature tables in ascending temperature order (pressures in mb: TORR already converted, :1665-1669). */
typedef struct mrtm_xs_region {
    int32_t ixmol;            /* 0-based row of XAMNT (the IXMOL of loop 6000) */
    int32_t ntemp;            /* NTEMPF(ixsr,ixmol), 1..6 */
    int64_t npts;             /* NPTSx */
    double v1fx, v2fx;        /* V1FX, V2FX: decide whether the region is processed (:1647-1652) */
    double v1x, v2x;          /* file header */
    double xdoplr;            /* XDOPLR(ixsr,ixmol) */
    double tx[6], pdx[6];
    const double *xsdat[6];   /* npts values per temperature */
} mrtm_xs_region;
/* Stage the tables once (the reference re-reads the files for every profile).  Regions ordered by molecule, then by
 * spectral region.  nreg == 0 clears the store. */
int mrtm_stage_xsec(mrtm_ctx *ctx, int64_t nreg, const mrtm_xs_region *regs);
/* Replaces CALL MONORTM_XSEC_SUB(wn,nwn,p,t,nlay,odxsec)  (src/monortm_sub.F90:1540, call src/modm.f90:198) with its
 * COMMON /PATHX/ input XAMNT passed explicitly: xamnt (ld_xamnt, nlay); odxsec (nwn, nlay). */
int mrtm_xsec(mrtm_ctx *ctx, int64_t nwn, const double *wn, int64_t nlay, const double *p, const double *t,
              int64_t ld_xamnt, const double *xamnt, double *odxsec);

/* ---- CALCTMR / RTM --------------------------------------------------------------------- */
/* CALL CALCTMR(NLAYRS,NWN,WN,T,TZ,O,TMR)  (src/RTMmono.f90:239).  tz is (0:nlayrs). */
int mrtm_calctmr(mrtm_ctx *ctx, int64_t nlayrs, int64_t nwn, const double *wn,
                 const double *t, const double *tz, const double *o, double *tmr);

/* CALL RTM(IOUT,IRT,NWN,WN,NLAY,T,TZ,O,TMPSFC,RUP,TRTOT,RDN,REFLC,EMISS,RAD,TB,IDU)
 * (src/RTMmono.f90:13-14).  tmpsfc is in/out: set to 2.75 for irt 2,3 (RTMmono.f90:122). */
int mrtm_rtm(mrtm_ctx *ctx, int64_t iout, int64_t irt, int64_t nwn, const double *wn,
             int64_t nlay, const double *t, const double *tz, const double *o,
             double *tmpsfc, double *rup, double *trtot, double *rdn,
             const double *reflc, const double *emiss, double *rad, double *tb, int64_t idu);

/* ---- fused path ------------------------------------------------------------------------ */
/* MODM + CALCTMR + RTM for nprof profiles in one call; optical depths stay in HBM.
 * Profile arrays carry a trailing profile dimension: p,t,clw,wbrodl (nlay,nprof);
 * tz (0:nlay, nprof) i.e. (nlay+1,nprof); wkl (39,nlay,nprof); scor (42,9,nlay,nprof);
 * tmpsfc (nprof) in/out; emiss,reflc (nwn) shared by all profiles.
 * Outputs (nwn,nprof): rad,tb,tmr,trtot,rup,rdn.  Optional (may be NULL):
 * o (nwn,nlay,nprof) layer optical depths; otot_by_mol (39,nwn,nprof) = layer sums of
 * o_by_mol+oc per molecule (what STOREOUT prints, src/monortm_sub.F90:649-656). */
int mrtm_profiles(mrtm_ctx *ctx, int64_t nprof, int64_t nwn, const double *wn, double dvset,
                  int64_t nlay, const double *p, const double *t, const double *tz,
                  const double *clw, int64_t nmol, const double *wkl, const double *wbrodl,
                  const double *scor, double sclcpl, double sclhw, double y0res,
                  const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                  double *tmpsfc, const double *emiss, const double *reflc,
                  double *rad, double *tb, double *tmr, double *trtot, double *rup, double *rdn,
                  double *o, double *otot_by_mol, const mrtm_opts *opts);

/* Same computation with every array already resident in device memory (layouts as above),
 * asynchronous on opts->stream; nothing is copied to the host.  Used when the caller keeps
 * inputs/outputs in HBM (ensembles, multi-GPU shards gathered with NCCL).  When the profiles fit one
 * batch of derived planes (MRTM_PLANES_GB) the call returns without waiting for the device: kernel times, counters and
 * device-detected errors (TIPS range, SDVOIGT, cross sections) are collected by the next mrtm_sync / mrtm_get_stats (which
 * return the deferred error code) or at the start of the next call. */
int mrtm_profiles_dev(mrtm_ctx *ctx, int64_t nprof, int64_t nwn, const double *wn_dev, double dvset,
                      int64_t nlay, const double *p, const double *t, const double *tz,
                      const double *clw, int64_t nmol, const double *wkl, const double *wbrodl,
                      const double *scor, double sclcpl, double sclhw, double y0res,
                      const double cntnm[7], int64_t ibrd, int64_t irt, int64_t iout, int64_t idu,
                      double *tmpsfc, const double *emiss_dev, const double *reflc_dev,
                      double *rad_dev, double *tb_dev, double *tmr_dev, double *trtot_dev,
                      double *rup_dev, double *rdn_dev, double *o_dev, const mrtm_opts *opts);

/* Wait for the last asynchronous call on ctx; returns its deferred error code (0 = none). */
int mrtm_sync(mrtm_ctx *ctx);

/* ---- instrumentation ------------------------------------------------------------------- */
typedef struct mrtm_stats {
    int64_t kernel_launches;     /* kernels launched by this ctx since the last reset */
    int64_t lines_staged;
    double last_lines_kernel_ms; /* CUDA-event time of the last line-shape kernel launch(es) of a call */
    double last_rt_kernel_ms;
    double last_derive_kernel_ms;
    double nominal_evals;        /* (logical line, layer, frequency) triples of the last call */
    double inwindow_evals;       /* triples that pass modm.f90:384 (needs opts->sel_count) else -1 */
    double far_expansions;       /* far-field Taylor expansions (line x resonance x frequency tile x layer) of the last call */
    double direct_evals;         /* (line, layer, frequency) triples evaluated directly by the last call */
    double last_prep_ms;         /* layer_prep + continuum kernels of the last call */
} mrtm_stats;
int mrtm_get_stats(mrtm_ctx *ctx, mrtm_stats *st);
int mrtm_reset_stats(mrtm_ctx *ctx);
/* FP64 FMA micro-benchmark on the ctx device: returns achieved TFLOP/s (2 flop per DFMA). */
int mrtm_fp64_peak(mrtm_ctx *ctx, double *tflops);

/* ---- host helpers (the harness side of SURVEY 8f-1; no GPU needed) ---------------------- */
/* GET_LNFL (src/lnfl_mod.f90:22-133): read a TAPE3 into caller-allocated lnfl_mod-layout
 * arrays of second dimension iim. */
int mrtm_host_get_lnfl(const char *hfile, double v1, double v2, int64_t iim, int64_t nblm[MRTM_MXMOL],
                       int64_t *iso, double *xnu0, double *deltnu, double *e, double *alps,
                       double *alpf, double *x, double *xg, double *s0, double *rmol, double *sdep,
                       int32_t *brd_mol_flg, double *brd_mol_tmp, double *brd_mol_hw,
                       double *brd_mol_shft);
/* TIPS_2003 (src/tips_2003.f90:2-298): scor(42,9) = Q(296)/Q(T) for molecules 1..mol_max. */
int mrtm_host_tips_2003(int64_t mol_max, double temp, double *scor);
/* XSREAD (src/monortm_sub.F90:1246-1420) for `ixmols` molecule names (10 characters each, left-justified or not) over
 * [xv1, xv2], followed by the file READs of MONORTM_XSEC_SUB (:1656-1671): parses <dir>/FSCDXS and the tables it names
 * (paths relative to dir).  *regs is allocated by the library (tables included); release with mrtm_host_xs_free. */
int mrtm_host_xsread(const char *dir, int64_t ixmols, const char *names, double xv1, double xv2,
                     int64_t *nreg, mrtm_xs_region **regs);
void mrtm_host_xs_free(mrtm_xs_region *regs, int64_t nreg);

/* ---- host driver (PROGRAM MONORTM around the hot path, layer input IATM=0; no
 * Fortran compiler needed).  The parsing / formatting entry points need no GPU. ------------------- */
/* What RDLBLINP (src/monortm_sub.F90:33-423) returns through its arguments and COMMON blocks
 * /MANE/ (DVSET), /BNDPRP/ (BNDEMI,BNDRFL), /profil_scal/, /CNTSCL/.  wn is malloc'ed by
 * mrtm_host_read_control and released by mrtm_host_free_control. */
typedef struct mrtm_control {
    int64_t ihirac, icntnm, iemit, iplot, iatm, iod, ixsect, ispd, ibrd;   /* record 1.2 */
    double cntnm[7];             /* applyCntnmCombo(ICNTNM) or record 1.2a (ICNTNM=6) */
    double v1, v2, dvset;        /* record 1.3 (dvset = 0 in list mode) */
    int64_t nwn;
    double *wn;                  /* records 1.3.1/1.3.2 or V1 + (j-1)*DVSET */
    double tmpbnd, bndemi[3], bndrfl[3];                                    /* record 1.4 */
    int64_t nmol_scal;
    char hmol_scal[64];
    double xmol_scal[64];
} mrtm_control;
/* detail of the last failing mrtm_host_* call of this thread (the reference's STOP message) */
const char *mrtm_host_last_error(void);
/* RDLBLINP for records 1.1-1.4.  nwnmx <= 0 -> the reference's NWNMX = 80000 (src/RTMmono.f90:10). */
int mrtm_host_read_control(const char *filein, int64_t nwnmx, mrtm_control *out);
void mrtm_host_free_control(mrtm_control *c);
/* GETPROFNUMBER, IATM=0 branch (src/monortm_sub.F90:894-907). */
int mrtm_host_count_profiles(const char *fileprof, int64_t ixsect, int64_t *nprof);
/* The MONORTM_PROF.IN block of PROGRAM MONORTM (src/monortm.f90:380-488) for the index-th
 * (0-based) profile of the file, mixing ratios converted to column amounts (:423-483).
 * p,t,clw,wbrodl (maxlay); altz,pz,tz (0:maxlay); wkl (39,maxlay). */
int mrtm_host_read_profile(const char *fileprof, int64_t index, int64_t maxlay, int64_t *iform, int64_t *nlay,
                           int64_t *nmol, int64_t *irt, double *secnt0, double *h1, double *h2, double *angle,
                           double *p, double *t, double *clw, double *wbrodl, double *altz, double *pz,
                           double *tz, double *wkl);
/* EMISS_REFLEC (src/monortm_sub.F90:506-516) with EMISFN/REFLFN (:426-493); tables are read from
 * <dir>in/EMISSION, <dir>in/REFLECTION when the first coefficient is negative (:319-336). */
int mrtm_host_emiss_reflec(const mrtm_control *c, const char *dir, int64_t nwn, const double *wn,
                           double *emiss, double *reflc);
/* STOREOUT (src/monortm_sub.F90:519-787; formats 11/21/31 :780-783; IOD=1 writes the ODmono_*
 * files next to fileout; the netCDF branch is not built).  Arguments as the reference's, arrays
 * dimensioned o,odxsec (nwn,nlay), o_by_mol,oc (nwn,39,nlay), wkl (39,nlay) IN/OUT (:600).
 * append == 0 truncates fileout (first profile), else appends. */
int mrtm_host_storeout(const char *fileout, int append, int64_t nwn, const double *wn, double *wkl,
                       const double *wbrodl, const double *rad, const double *tb, const double *trtot,
                       int64_t npr, const double *o, const double *o_by_mol, const double *oc,
                       const double *odxsec, const double *tmr, double wvcolmn, double clwcolmn,
                       double tmpsfc, const double *reflc, const double *emiss, int64_t nlay, int64_t nmol,
                       double angle, int64_t iod);
/* PROGRAM MONORTM for IATM=0 (src/monortm.f90:283-588): reads MONORTM.IN, MONORTM_PROF.IN and
 * TAPE3 in workdir, runs every profile through mrtm_profiles on `device`, writes MONORTM.OUT
 * (+ MONORTM.LOG, ODmono_* with IOD=1).  Needs a GPU: there is no CPU path. */
int mrtm_host_run_monortm(const char *workdir, int device, int64_t nwnmx, int verbose);

#ifdef __cplusplus
}
#endif
#endif
