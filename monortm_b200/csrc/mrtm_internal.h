// mrtm_internal.h -- shared host/device definitions of libmonortm_b200 (not part of the ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/monortm_b200.h"

namespace mrtm {

// ---- literal constants, digit for digit: src/PhysConstants.f90:19-39 --------------------
constexpr double kPI = 3.1415926535898;
constexpr double kPLANCK = 6.62606876E-27;
constexpr double kBOLTZ = 1.3806503E-16;
constexpr double kCLIGHT = 2.99792458E+10;
constexpr double kAVOGAD = 6.02214199E+23;
constexpr double kRADCN1 = 1.191042722E-12;
constexpr double kRADCN2 = 1.4387752;
constexpr double kT0 = 296.;       // modm.f90:875
constexpr double kP0 = 1013.25;    // modm.f90:876
constexpr double kDELTNUC = 25.;   // modm.f90:301

// ---- line classes (warp-uniform dispatch in the line kernel) ------------------------------
enum LineClass : int32_t {
    CLS_PED = 0,      // molecule not CO2/O2, no coupling: Lorentz pair with 25 cm-1 pedestal (modm.f90:742-752)
    CLS_O2 = 1,       // O2, no coupling: window test inside the LSF routine, no pedestal (:755-766)
    CLS_O2_LC1 = 2,   // O2, XF=-1: first-order mixing, both resonances, no window (:777-786)
    CLS_O2_LC35 = 3,  // O2, XF=-3/-5: both resonances, no window, Y=1 (:787-791)
    CLS_GENERAL = 4,  // CO2 (any XF) and coupled lines of other molecules: full case tree
    CLS_COUNT = 5
};

struct Segment {          // contiguous run of staged lines of one (molecule, class), ascending xnu0
    int32_t mol;          // 1..39
    int32_t cls;
    int32_t begin, end;   // [begin,end) in staged order
    int32_t count_all;    // = end-begin
    int32_t slot;         // compact index of the molecule among the molecules that own lines
    uint64_t hash_all;    // sum of line keys (for O2 every line passes modm.f90:384)
    double vfac;          // 100*HWHM_D upper bound factor: vthr <= vfac*sqrt(T)
    double vrate;         // the same bound per unit |Xnu|: 100*HWHM_D/|Xnu| <= vrate*sqrt(T)
};

constexpr int kMaxSegments = 96;
constexpr int kMaxSlots = MRTM_MXMOL;   // molecules that own lines, compacted

// derived per-(layer,line) parameter planes written by derive_kernel
// Five dense planes, 40 B per (line, layer).  Everything else a kernel may need follows from them and from two compact
// planes that exist only for coupled lines (LCP_AIP, LCP_BIP, indexed by lcidx): HWHM_C = sqrt(H2) (exact: H2 is the
// rounded square of HWHM_C), HWHM_D = VT/100 where the line can take the Voigt branch (VT >= 0; it is not used otherwise),
// STILD = CN*PI/HWHM_C (2 ulp), the first-order mixing slope CN*AIP*(1/HWHM_C)*RP.
enum DPlane : int {
    D_XNU = 0,    // shifted centre, bit-exact modm.f90:375-380
    D_H2,         // HWHM_C**2
    D_CN,         // STILD*HWHM_C/PI
    D_P3,         // pedestal CN/(625+H2) (CLS_PED) | CN*(1+BIP*RP2) (CLS_O2_LC1) | 0
    D_VT,         // 100*HWHM_D, or -1 when zeta>0.99 (always Lorentz, modm.f90:427)
    D_NPLANES
};
enum LcPlane : int { LCP_AIP = 0, LCP_BIP, LCP_NPLANES };

// per (profile,layer) scalars prepared on the host with reference evaluation order
struct LayerDev {
    double t, p, rp, rp2;
    double rt;          // T/T0
    double rhorat;      // Xn/XN0 (also the shift ratio, modm.f90:312,375)
    double rectlc, tmpdif;
    double radct;
    double xkt;         // T/RADCN2
    double clw;
    double wtot;
    double shift_margin;   // >= max |Xnu-Xnu0| over staged lines for this layer
    double sqrt_t;
    double lnrt;        // log(rt): rt^x = exp(x*lnrt) in derive_kernel
    double inv_t;       // 1/T
    double dinvt;       // 1/T0 - 1/T: Boltzmann ratio exp(-c2 E/T)/exp(-c2 E/T0) = exp(c2*E*dinvt)
    double rho_molec[7];
    double wk[MRTM_MXMOL];     // column amounts (W_species), zero beyond nmol
    double rho_self[MRTM_MXMOL];
    // continuum per-layer scalars (contnm.f90:222-240,300-302,487,919)
    double c_wk1, c_rself, c_rfrgn, c_tfac_h2o, c_wco2, c_trat, c_taufac, c_tfac_n2;
    double c_xn2, c_xo2, c_xh2o;
    // ... and of the branches above the microwave (contnm.f90:229, 541, 665, 718, 756, 820, 838)
    double c_amagat, c_rhoave, c_wk3, c_wk7, c_cw;
    int32_t ilc;       // 1..3
    int32_t pad;
};

// continuum components of CONTNM (contnm.f90:325-1131) and the per-species planes they accumulate into (MODM calls CONTNM
// once per species with the other scale factors zeroed, modm.f90:210-214)
enum ContBranch : int {
    CB_H2O_SELF = 0, CB_H2O_FRGN, CB_CO2, CB_N2_ROT, CB_N2_FUND, CB_N2_OVER, CB_O3_CHAP, CB_O3_HH, CB_O3_UV,
    CB_O2_FUND, CB_O2_INF1, CB_O2_INF2, CB_O2_INF3, CB_O2_VIS, CB_O2_HERZ, CB_O2_FUV, CB_COUNT
};
enum ContPlane : int { CP_H2O = 0, CP_CO2, CP_O3, CP_O2, CP_N2, CP_RAYL, CP_COUNT };

struct ContGrid {     // accessor/XINT index set-up for one continuum component (layer independent)
    double v1c, dvc;
    int32_t nptc, i1;      // source points, table offset (contnm.f90:1447-1456)
    int32_t ilo, ihi;      // XINT loop bounds on the 1 cm-1 grid (1-based, inclusive)
    int32_t active, pad;
};

struct HostLines;   // staging result (mrtm_stage.cpp)

}  // namespace mrtm
