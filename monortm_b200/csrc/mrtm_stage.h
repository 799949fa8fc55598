// mrtm_stage.h -- host-side staging of the lnfl_mod line store into a sorted structure-of-arrays.
#pragma once
#include "mrtm_internal.h"

namespace mrtm {

struct HostLines {
    int64_t n = 0;       // logical lines
    int64_t n_pad = 0;   // padded plane length (multiple of 8, >= n+8)
    // static per-line planes, staged order = (molecule asc, class asc, xnu0 asc [stable])
    std::vector<int32_t> mol, iso, xf, cls, sidx, lcidx, brdidx, rec, segidx;
    std::vector<double> xnu0, s0adj, e, alpf, alps, x, deltnu, sdep, mass;
    std::vector<double> dopf;      // sqrt(2 ln2 k N_A / M)/c per line: HWHM_D = Xnu*dopf*sqrt(T) (modm.f90:453 with the constants folded)
    std::vector<uint64_t> key;
    // compact line-coupling table: 16 doubles per coupled line: A[4],B[4] foreign, A2[4],B2[4] self
    std::vector<double> lc;        // size nlc*16
    std::vector<int32_t> lc_self;  // 1 if the self set applies (modm.f90:339)
    // compact species-broadening table: per entry flg[7], hw[7], tmp[7], shft[7] (28 doubles)
    std::vector<double> brd;       // size nbrd*28
    // compact (molecule,isotopologue) pairs that staged lines use -> index into scor
    std::vector<int32_t> scor_index;   // (mol-1) + (iso-1)*42 for each compact slot
    std::vector<Segment> segments;
    std::vector<int32_t> slot_mol;     // molecules owning lines, ascending (compact "slots")
    int32_t mol_slot[MRTM_MXMOL];      // molecule-1 -> slot or -1
    double max_abs_deltnu = 0.0;
    double max_abs_brd_dshift = 0.0;   // max |shft_k - deltnu| over broadening entries
    std::string error;
};

// Walk the lnfl_mod arrays exactly as LINES does (src/modm.f90:316-435) and build HostLines.
// Returns MRTM_OK or an error code (detail in out.error).
int stage_lines_host(const int64_t nblm[MRTM_MXMOL], int64_t iim, const int64_t* iso,
                     const double* xnu0, const double* deltnu, const double* e, const double* alps,
                     const double* alpf, const double* x, const double* xg, const double* s0,
                     const double* rmol, const double* sdep, const int32_t* brd_mol_flg,
                     const double* brd_mol_tmp, const double* brd_mol_hw, const double* brd_mol_shft,
                     HostLines& out);

uint64_t line_key(int64_t mol, int64_t rec);
double smass(int mol, int iso);   // isotope.incl SMASS(mol,iso), 1-based

}  // namespace mrtm
