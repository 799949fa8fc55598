// mrtm_host.cpp -- host-side helpers above the C ABI (no GPU needed): the TAPE3 reader that
// fills lnfl_mod-layout arrays (GET_LNFL, src/lnfl_mod.f90:22-133; record layouts
// src/struct_types.f90:27-43; BUFIN_sgl src/bufin_sgl.f90) and TIPS_2003
// (src/tips_2003.f90:2-298, AtoB :4610-4700).  In production the unchanged Fortran host does
// both; these exist so the C++/Python harness can run without a Fortran compiler.
// Compile with -ffp-contract=off.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mrtm_internal.h"
#include "tables/tips_tables.inc"

namespace {

inline size_t ix(int64_t i, int64_t j) { return (size_t)(i - 1) + (size_t)(j - 1) * MRTM_MXMOL; }
inline size_t ib(int64_t i, int64_t k, int64_t j) { return (size_t)(i - 1) + (size_t)(k - 1) * 7 + (size_t)(j - 1) * 49; }

// Fortran sequential unformatted record, 4-byte markers (-frecord-marker=4)
struct RecReader {
    FILE* f;
    // returns payload size, -1 at EOF, -2 if malformed
    long next(std::vector<unsigned char>& buf, size_t want)
    {
        int32_t n1 = 0, n2 = 0;
        if (std::fread(&n1, 4, 1, f) != 1) return -1;
        if (n1 < 0) return -2;
        size_t take = std::min((size_t)n1, want);
        buf.assign(want, 0);
        if (take && std::fread(buf.data(), 1, take, f) != take) return -2;
        if ((size_t)n1 > take && std::fseek(f, (long)((size_t)n1 - take), SEEK_CUR) != 0) return -2;
        if (std::fread(&n2, 4, 1, f) != 1 || n1 != n2) return -2;
        return n1;
    }
};

template <class T>
T rd(const std::vector<unsigned char>& b, size_t off)
{
    T v;
    std::memcpy(&v, b.data() + off, sizeof(T));
    return v;
}

constexpr int kRec = 250;   // NLINEREC, struct_types.f90:25
// INPUT_BLOCK member offsets in bytes (struct_types.f90:33-43)
constexpr size_t oVNU = 0, oSP = 2000, oALFA = 3000, oEPP = 4000, oMOL = 5000, oHWHM = 6000,
                 oTMPALF = 7000, oPSHIFT = 8000, oIFLG = 9000, oBFLG = 10000, oBDAT = 17000,
                 oSDEP = 38000, kBlockBytes = 39000;

// AtoB, tips_2003.f90:4610-4700
double atob(double aa, const double* A, const double* B, int npt)
{
    double bb = 0.;
    for (int I = 2; I <= npt; I++) {
        if (A[I - 1] >= aa) {
            auto nz = [](double d) { return d == 0. ? 0.0001 : d; };
            if (I < 3 || I == npt) {
                int J = I;
                if (I < 3) J = 3;
                if (I == npt) J = npt;
                double a0d1 = nz(A[J - 3] - A[J - 2]), a0d2 = nz(A[J - 3] - A[J - 1]);
                double a1d1 = nz(A[J - 2] - A[J - 3]), a1d2 = nz(A[J - 2] - A[J - 1]);
                double a2d1 = nz(A[J - 1] - A[J - 3]), a2d2 = nz(A[J - 1] - A[J - 2]);
                double a0 = (aa - A[J - 2]) * (aa - A[J - 1]) / (a0d1 * a0d2);
                double a1 = (aa - A[J - 3]) * (aa - A[J - 1]) / (a1d1 * a1d2);
                double a2 = (aa - A[J - 3]) * (aa - A[J - 2]) / (a2d1 * a2d2);
                bb = a0 * B[J - 3] + a1 * B[J - 2] + a2 * B[J - 1];
            } else {
                int J = I;
                double a0d1 = nz(A[J - 3] - A[J - 2]), a0d2 = nz(A[J - 3] - A[J - 1]), a0d3 = nz(A[J - 3] - A[J]);
                double a1d1 = nz(A[J - 2] - A[J - 3]), a1d2 = nz(A[J - 2] - A[J - 1]), a1d3 = nz(A[J - 2] - A[J]);
                double a2d1 = nz(A[J - 1] - A[J - 3]), a2d2 = nz(A[J - 1] - A[J - 2]), a2d3 = nz(A[J - 1] - A[J]);
                double a3d1 = nz(A[J] - A[J - 3]), a3d2 = nz(A[J] - A[J - 2]), a3d3 = nz(A[J] - A[J - 1]);
                double a0 = (aa - A[J - 2]) * (aa - A[J - 1]) * (aa - A[J]);
                a0 = a0 / (a0d1 * a0d2 * a0d3);
                double a1 = (aa - A[J - 3]) * (aa - A[J - 1]) * (aa - A[J]);
                a1 = a1 / (a1d1 * a1d2 * a1d3);
                double a2 = (aa - A[J - 3]) * (aa - A[J - 2]) * (aa - A[J]);
                a2 = a2 / (a2d1 * a2d2 * a2d3);
                double a3 = (aa - A[J - 3]) * (aa - A[J - 2]) * (aa - A[J - 1]);
                a3 = a3 / (a3d1 * a3d2 * a3d3);
                bb = a0 * B[J - 3] + a1 * B[J - 2] + a2 * B[J - 1] + a3 * B[J];
            }
            break;
        }
    }
    return bb;
}

}  // namespace

namespace mrtm {
// used by the device-table upload in mrtm_api.cu
const double* tips_qoft() { return TIPS_QOFT; }
const double* tips_tdat() { return TIPS_TDAT; }
int tips_rows() { return TIPS_QROWS; }
// row of (mol, iso) in TIPS_QOFT; -2: the reference's dispatch yields scor == 1 (molecule 34, atomic oxygen: gi=1, QT=1,
// tips_2003.f90:233-238; molecule 39, CH3OH: both qt_296 and qt_temp end up as the stale QT of the previous call,
// :260-268, 288-292); -1: no such (molecule, isotopologue)
int tips_row(int mol, int iso)
{
    if (mol < 1 || mol > 39) return -1;
    if (mol == 34 || mol == 39) return iso == 1 ? -2 : -1;
    int nuse = TIPS_ISONM[mol - 1] < 9 ? TIPS_ISONM[mol - 1] : 9;   // min(9,isonm(mol)), tips_2003.f90:60
    if (iso < 1 || iso > nuse || iso > TIPS_QNISO[mol - 1]) return -1;
    return TIPS_QOFFSET[mol - 1] + (iso - 1);
}
}  // namespace mrtm

extern "C" int mrtm_host_tips_2003(int64_t mol_max, double temp, double* scor)
{
    if (mol_max < 1 || mol_max > 39) return MRTM_ETIPS;
    if (temp < 70. || temp > 3000.) return MRTM_ETIPS;     // Qt=-1 -> STOP
    for (int64_t mol = 1; mol <= mol_max; mol++) {
        if (mol == 34 || mol == 39) {                      // QT := 1 (O, tips_2003.f90:233-238); CH3OH: stale QT / stale QT (:260-292)
            scor[(mol - 1)] = 1.0;
            continue;
        }
        int nuse = TIPS_ISONM[mol - 1] < 9 ? TIPS_ISONM[mol - 1] : 9;
        for (int iso = 1; iso <= nuse; iso++) {
            const double* q = TIPS_QOFT + (size_t)(TIPS_QOFFSET[mol - 1] + iso - 1) * 119;
            double qt_296 = atob(296., TIPS_TDAT, q, 119);
            double qt_temp = atob(temp, TIPS_TDAT, q, 119);
            if (qt_296 <= 0. || qt_temp <= 0.) return MRTM_ETIPS;
            scor[(mol - 1) + (iso - 1) * MRTM_NSCOR1] = qt_296 / qt_temp;
        }
    }
    return MRTM_OK;
}

extern "C" int mrtm_host_get_lnfl(const char* hfile, double v1, double v2, int64_t iim, int64_t nblm[MRTM_MXMOL],
                                  int64_t* iso, double* xnu0, double* deltnu, double* e, double* alps,
                                  double* alpf, double* x, double* xg, double* s0, double* rmol, double* sdep,
                                  int32_t* brd_mol_flg, double* brd_mol_tmp, double* brd_mol_hw,
                                  double* brd_mol_shft)
{
    FILE* f = std::fopen(hfile, "rb");
    if (!f) return MRTM_EIO;                                  // ERROR OPENING HITRAN FILE
    RecReader rr{f};
    std::vector<unsigned char> hdr, blk, ph;
    for (int m = 0; m < MRTM_MXMOL; m++) nblm[m] = 0;
    auto bail = [&](int code) { std::fclose(f); return code; };

    // PRLNHD: file header, optional negative-EPP header, isotope flag (lnfl_mod.f90:250-302)
    long n = rr.next(hdr, 1664);
    if (n < 1664) return bail(MRTM_EIO);
    if (hdr[6 * 8 + 7] == '^') {
        std::vector<unsigned char> h2;
        if (rr.next(h2, 16) < 0) return bail(MRTM_EIO);
    }
    if (hdr[9 * 8 + 7] != 'I') return bail(MRTM_ELINEFILE);  // NO ISOTOPE INFO ON LINFIL

    const double vlo_adj = std::fmax(0.0, v1 - 25.0);         // RDLNFL :161
    int64_t mo_prev = 0;
    bool eof = false;
    while (!eof) {
        // RDLNFL: skip panels wholly below the range, then read one 250-record block
        int32_t nrec = 0;
        for (;;) {
            long m = rr.next(ph, 24);
            if (m < 0) { eof = true; break; }
            if (m < 24) return bail(MRTM_EIO);
            double vmax = rd<double>(ph, 8);
            nrec = rd<int32_t>(ph, 16);
            if (vmax < vlo_adj) {
                std::vector<unsigned char> dum;
                if (rr.next(dum, 4) < 0) { eof = true; }
                if (eof) break;
                continue;
            }
            if (rr.next(blk, kBlockBytes) < 0) { eof = true; }
            break;
        }
        if (eof) break;
        if (nrec < 0 || nrec > kRec) return bail(MRTM_EIO);
        double last_vnu = 0.;
        for (int ik = 1; ik <= nrec; ik++) {
            const int64_t iflg = rd<int32_t>(blk, oIFLG + 4 * (ik - 1));
            const int64_t mol = rd<int32_t>(blk, oMOL + 4 * (ik - 1));
            int64_t mo;
            if (iflg >= 0 && iflg <= 100) {
                mo = mol % 100;
            } else if (iflg >= -3 && iflg <= -1) {
                if (ik == 1) return bail(MRTM_ELINEFILE);
                mo = (int64_t)rd<int32_t>(blk, oMOL + 4 * (ik - 2)) % 100;
            } else if (iflg == -5) {
                if (ik == 1) return bail(MRTM_ELINEFILE);
                if (rd<int32_t>(blk, oIFLG + 4 * (ik - 2)) >= 0) {
                    mo = (int64_t)rd<int32_t>(blk, oMOL + 4 * (ik - 2)) % 100;
                    mo_prev = mo;
                } else {
                    mo = mo_prev;
                }
            } else {
                return bail(MRTM_ELINEFILE);                  // LC flag not recognized
            }
            if (mo < 1 || mo > MRTM_MXMOL) return bail(MRTM_ELINEFILE);
            const int64_t ii = ++nblm[mo - 1];
            if (ii > iim) return bail(MRTM_ENOMEM);
            iso[ix(mo, ii)] = (mol % 1000) / 100;
            xnu0[ix(mo, ii)] = rd<double>(blk, oVNU + 8 * (ik - 1));
            s0[ix(mo, ii)] = rd<float>(blk, oSP + 4 * (ik - 1));
            alpf[ix(mo, ii)] = rd<float>(blk, oALFA + 4 * (ik - 1));
            alps[ix(mo, ii)] = rd<float>(blk, oHWHM + 4 * (ik - 1));
            e[ix(mo, ii)] = rd<float>(blk, oEPP + 4 * (ik - 1));
            x[ix(mo, ii)] = rd<float>(blk, oTMPALF + 4 * (ik - 1));
            deltnu[ix(mo, ii)] = rd<float>(blk, oPSHIFT + 4 * (ik - 1));
            xg[ix(mo, ii)] = (iflg >= 0) ? (double)(-iflg) : (double)iflg;
            rmol[ix(mo, ii)] = (double)rd<float>(blk, oMOL + 4 * (ik - 1));   // int*4 bits viewed as real*4
            if (mo <= MRTM_MXBRDMOL) {
                for (int k = 1; k <= 7; k++) {
                    brd_mol_flg[ib(mo, k, ii)] = rd<int32_t>(blk, oBFLG + 4 * ((k - 1) + 7 * (ik - 1)));
                    const size_t d = oBDAT + 4 * (size_t)(3 * (k - 1) + 21 * (ik - 1));
                    brd_mol_hw[ib(mo, k, ii)] = rd<float>(blk, d);
                    brd_mol_tmp[ib(mo, k, ii)] = rd<float>(blk, d + 4);
                    brd_mol_shft[ib(mo, k, ii)] = rd<float>(blk, d + 8);
                }
            }
            sdep[ix(mo, ii)] = rd<float>(blk, oSDEP + 4 * (ik - 1));
            // air -> foreign width (and O2 shift) correction, lnfl_mod.f90:98-113
            if (mo == 7 && iflg >= 0) {
                const double rvmr = 0.21;
                alpf[ix(mo, ii)] = (alpf[ix(mo, ii)] - rvmr * alps[ix(mo, ii)]) / (1.0 - rvmr);
                if (brd_mol_flg[ib(mo, mo, ii)] > 0)
                    deltnu[ix(mo, ii)] = (deltnu[ix(mo, ii)] - rvmr * brd_mol_shft[ib(mo, mo, ii)]) / (1.0 - rvmr);
            }
            if (mo == 22 && iflg >= 0) {
                const double rvmr = 0.79;
                alpf[ix(mo, ii)] = (alpf[ix(mo, ii)] - rvmr * alps[ix(mo, ii)]) / (1.0 - rvmr);
            }
            last_vnu = rd<double>(blk, oVNU + 8 * (ik - 1));
        }
        if (nrec >= 1 && last_vnu > (v2 + 25.)) eof = true;    // :116
    }
    std::fclose(f);
    return MRTM_OK;
}
