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

// ---------------------------------------------------------------------------------------------
// Cross sections: XSREAD (src/monortm_sub.F90:1246-1420) + the table READs of MONORTM_XSEC_SUB
// (:1656-1671).  ALIAS / XSMASS: BLOCK DATA BXSECT (:1445-1474).
// ---------------------------------------------------------------------------------------------
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

namespace {

const char* const kXsAlias[15][4] = {
    {"CLONO2", "CLNO3", nullptr, nullptr}, {"HNO4", nullptr, nullptr, nullptr}, {"CHCL2F", "CFC21", "CFC21", "F21"},
    {"CCL4", nullptr, nullptr, nullptr}, {"CCL3F", "CFCL3", "CFC11", "F11"}, {"CCL2F2", "CF2CL2", "CFC12", "F12"},
    {"C2CL2F4", "C2F4CL2", "CFC114", "F114"}, {"C2CL3F3", "C2F3CL3", "CFC113", "F113"}, {"N2O5", nullptr, nullptr, nullptr},
    {"HNO3", nullptr, nullptr, nullptr}, {"CF4", nullptr, "CFC14", "F14"}, {"CHCLF2", "CHF2CL", "CFC22", "F22"},
    {"CCLF3", nullptr, "CFC13", "F13"}, {"C2CLF5", nullptr, "CFC115", "F115"}, {"NO2", nullptr, nullptr, nullptr}};
const double kXsMass[15] = {97.46, 79.01, 102.92, 153.82, 137.37, 120.91, 170.92, 187.38, 108.01, 63.01, 88.00, 86.47, 104.46, 154.47, 45.99};

std::string xs_trim(const std::string& s)
{
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(" \r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string xs_field(const std::string& rec, size_t off, size_t len) { return off < rec.size() ? rec.substr(off, len) : std::string(); }
double xs_num(const std::string& f)
{
    std::string t;                                            // blanks inside a numeric field are ignored by a formatted READ
    for (char c : f) if (c != ' ' && c != '\r' && c != '\n') t += (c == 'D' || c == 'd') ? 'E' : c;
    return t.empty() ? 0. : std::strtod(t.c_str(), nullptr);
}
bool xs_matches(const std::string& name, int j)
{
    for (int k = 0; k < 4; k++) if (kXsAlias[j][k] && name == kXsAlias[j][k]) return true;
    return false;
}

static int xs_fail(int code, const std::string& msg);

}  // namespace

extern "C" const char* mrtm_host_last_error(void);
namespace mrtm_hostdrv { int set_host_error(int code, const std::string& msg); }
namespace { int xs_fail(int code, const std::string& msg) { return mrtm_hostdrv::set_host_error(code, msg); } }

extern "C" void mrtm_host_xs_free(mrtm_xs_region* regs, int64_t nreg)
{
    if (!regs) return;
    for (int64_t i = 0; i < nreg; i++)
        for (int k = 0; k < 6; k++) delete[] const_cast<double*>(regs[i].xsdat[k]);
    delete[] regs;
}

extern "C" int mrtm_host_xsread(const char* dir, int64_t ixmols, const char* names, double xv1, double xv2,
                                int64_t* nreg, mrtm_xs_region** regs_out)
{
    if (!dir || !names || !nreg || !regs_out || ixmols < 1 || ixmols > 38) return xs_fail(MRTM_EARG, "mrtm_host_xsread: bad argument");
    *nreg = 0;
    *regs_out = nullptr;
    std::string d(dir);
    if (!d.empty() && d.back() != '/') d += '/';
    std::vector<int> ixindx((size_t)ixmols);
    std::vector<std::string> xsname((size_t)ixmols);
    for (int64_t i = 0; i < ixmols; i++) {                    // :1296-1330: left-justify, match against ALIAS
        xsname[(size_t)i] = xs_trim(std::string(names + 10 * i, 10));
        int found = -1;
        for (int j = 0; j < 15 && found < 0; j++) if (xs_matches(xsname[(size_t)i], j)) found = j;
        if (found < 0) return xs_fail(MRTM_EIO, "  THE NAME: " + xsname[(size_t)i] + " IS NOT ONE OF THE CROSS SECTION MOLECULES. CHECK THE SPELLING.");
        ixindx[(size_t)i] = found;
    }
    std::ifstream f(d + "FSCDXS");
    if (!f) return xs_fail(MRTM_EIO, "FSCDXS does not exist - XSREAD");
    std::string rec;
    std::getline(f, rec);                                     // FORMAT 905 (/): two records skipped
    std::getline(f, rec);
    std::vector<std::vector<mrtm_xs_region>> per((size_t)ixmols);
    std::vector<char> flg((size_t)ixmols, 0);
    auto cleanup = [&]() {
        for (auto& v : per) for (auto& g : v) for (int k = 0; k < 6; k++) delete[] const_cast<double*>(g.xsdat[k]);
    };
    while (std::getline(f, rec)) {
        if (rec.empty()) rec = " ";
        if (rec[0] == '*') continue;
        if (rec[0] == '%') break;
        rec.resize(120, ' ');                                 // FORMAT 915 (A10,2F10.4,F10.8,I5,5X,I5,A1,4X,6A10)
        const std::string xname = xs_trim(xs_field(rec, 0, 10));
        const double v1x = xs_num(xs_field(rec, 10, 10)), v2x = xs_num(xs_field(rec, 20, 10));
        const int ntemp = (int)xs_num(xs_field(rec, 40, 5));
        for (int64_t i = 0; i < ixmols; i++) {
            if (!xs_matches(xname, ixindx[(size_t)i])) continue;
            flg[(size_t)i] = 1;
            if (!(v2x > xv1 && v1x < xv2)) continue;
            if (per[(size_t)i].size() >= 5) { cleanup(); return xs_fail(MRTM_EIO, " XSREAD - NSPECR .GT. 5 (the arrays hold five regions per molecule)"); }
            if (ntemp < 1 || ntemp > 6) { cleanup(); return xs_fail(MRTM_EIO, "FSCDXS: NTEMP out of 1..6 for " + xname); }
            mrtm_xs_region g;
            std::memset(&g, 0, sizeof g);
            g.ixmol = (int32_t)i;
            g.ntemp = ntemp;
            g.v1fx = v1x;
            g.v2fx = v2x;
            g.xdoplr = 3.58115E-07 * (0.5 * (v1x + v2x)) * std::sqrt(296.0 / kXsMass[ixindx[(size_t)i]]);   // :1385-1386
            for (int k = 0; k < ntemp; k++) {                 // MONORTM_XSEC_SUB :1656-1671
                const std::string fn = xs_trim(xs_field(rec, 60 + 10 * (size_t)k, 10));
                std::ifstream x(d + fn);
                std::string hdr;
                if (!x || !std::getline(x, hdr)) { cleanup(); return xs_fail(MRTM_EIO, "cross-section file missing or empty: " + d + fn); }
                hdr.resize(100, ' ');                         // FORMAT 910 (A10,2F10.4,I10,3G10.3,3A10)
                g.v1x = xs_num(xs_field(hdr, 10, 10));
                g.v2x = xs_num(xs_field(hdr, 20, 10));
                g.npts = (int64_t)xs_num(xs_field(hdr, 30, 10));
                g.tx[k] = xs_num(xs_field(hdr, 40, 10));
                const double pres = xs_num(xs_field(hdr, 50, 10));
                g.pdx[k] = (xs_field(hdr, 90, 10) == "      TORR") ? pres * (1013. / 760) : pres;
                if (g.npts < 2 || g.npts > 150000) { cleanup(); return xs_fail(MRTM_EIO, "cross-section file: NPTS out of 2..150000 (xsdat(150000,6)): " + fn); }
                double* dat = new double[(size_t)g.npts];
                g.xsdat[k] = dat;
                // values written with 1PE10.3 touch when negative ("0.000E+00-1.025E-26" in the shipped HNO3 files): walk the
                // number syntax instead of splitting on blanks
                int64_t n = 0;
                std::string line;
                while (n < g.npts && std::getline(x, line)) {
                    for (auto& c : line) if (c == 'D' || c == 'd') c = 'E';
                    const char* q = line.c_str();
                    while (n < g.npts && *q) {
                        while (*q == ' ' || *q == ',' || *q == '\t' || *q == '\r') q++;
                        if (!*q) break;
                        char* end = nullptr;
                        const double v = std::strtod(q, &end);
                        if (end == q) break;
                        dat[n++] = v;
                        q = end;
                    }
                }
                if (n < g.npts) { per[(size_t)i].push_back(g); cleanup(); return xs_fail(MRTM_EIO, "cross-section file shorter than its header says: " + fn); }
            }
            per[(size_t)i].push_back(g);
        }
    }
    for (int64_t i = 0; i < ixmols; i++)
        if (!flg[(size_t)i]) { cleanup(); return xs_fail(MRTM_EIO, "******* MOLECULE SELECTED -" + xsname[(size_t)i] + "- IS NOT FOUND ON FILE FSCDXS *******"); }
    size_t tot = 0;
    for (auto& v : per) tot += v.size();
    mrtm_xs_region* out = new mrtm_xs_region[std::max<size_t>(tot, 1)];
    size_t k = 0;
    for (auto& v : per) for (auto& g : v) out[k++] = g;
    *nreg = (int64_t)tot;
    *regs_out = out;
    return MRTM_OK;
}
