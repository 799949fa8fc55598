// mrtm_kernels.cuh -- the sm_100a kernels of the hot path.
//   layer_prep_kernel   per-(profile,layer) scalars with the reference's evaluation order
//   continuum_kernel    MT_CKD (V2<820 cm-1 subset) onto the 1 cm-1 ABSRB grid per layer
//   derive_kernel       per-(line,layer) derived parameters (shifted centre, widths, STILD ...)
//   plan / far / far_warp / near / near2 / voigt / final   the line path (kernels/*.cuh)
//   colsum_kernel       layer sums of per-molecule optical depths (STOREOUT's columns)
//   rt_kernel           CALCTMR + RAD_UP_DN + RTM
#pragma once
#include <cuda_runtime.h>

#include "mrtm_device.cuh"

namespace mrtm {

// ---------------------------------------------------------------------------------------------
// static line planes (device pointers)
struct LinesDev {
    int32_t n, n_pad;
    const int32_t *mol, *iso, *xf, *cls, *sidx, *lcidx, *brdidx, *segidx;
    const double *xnu0, *s0adj, *e, *alpf, *alps, *x, *deltnu, *sdep, *mass, *dopf;
    const unsigned long long* key;
    const unsigned long long* keypre;   // [n_pad+1]
    const double* lc;          // [nlc][16]
    const int32_t* lc_self;    // [nlc]
    const double* brd;         // [nbrd][28]
    int32_t nsi;               // compact scor slots
    const int32_t* scor_index; // [nsi] -> (mol-1)+(iso-1)*42
    const int32_t* slot_mol;   // [nslot] molecule of each compact slot (ascending)
};

struct ContTablesDev {
    const double *sh2o_296, *sh2o_260, *fh2o, *fco2, *n2_296, *n2_296_sf, *n2_220, *n2_220_sf;
    const double *xfac_rhu, *co2_tdep;
    // branches above the microwave (tables/mtckd_ir_tables.inc)
    const double *xfacco2, *n2f_272, *n2f_228, *n2f_ah2o, *n2f1, *o3ch_x, *o3ch_y, *o3ch_z, *o3hh0, *o3hh1, *o3hh2, *o3huv;
    const double *o2f, *o2f_t, *o2inf1, *o2inf3, *o2vis, *o2fuv;
};

struct TipsDev {
    const double* qoft;      // [rows][119]
    const double* tdat;      // [119]
    const int32_t* row;      // [nsi] row in qoft for compact slot, or -1
};

#include "kernels/prep.cuh"
#include "kernels/lines_common.cuh"
#include "kernels/plan.cuh"
#include "kernels/far.cuh"
#include "kernels/near.cuh"
#include "kernels/near3.cuh"
#include "kernels/sparse.cuh"
#include "kernels/voigt.cuh"
#include "kernels/voigt_t.cuh"
#include "kernels/final.cuh"
#include "kernels/rt.cuh"
#include "kernels/xsec.cuh"

}  // namespace mrtm
