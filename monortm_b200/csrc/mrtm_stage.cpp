// mrtm_stage.cpp -- stage the lnfl_mod line store (src/lnfl_mod.f90:9-13) as a sorted
// structure-of-arrays.  The record walk reproduces the control flow of LINES
// (src/modm.f90:316-435): coupling-coefficient records follow their parent line and are
// skipped with J=JJ (:434); the self-coupling blend is enabled only when XG(I,J)==-5 and
// XG(I,J-1)==-5 (:339; XG(I,0) is treated as "not -5").
// Compile with -ffp-contract=off: S0_adj (:372) is evaluated here once per line.
#include "mrtm_stage.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

#include "tables/smass_table.inc"

namespace mrtm {

uint64_t line_key(int64_t mol, int64_t rec)
{
    uint64_t z = ((uint64_t)mol << 32) | (uint64_t)rec;   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

double smass(int mol, int iso) { return ISO_SMASS[(mol - 1) * 9 + (iso - 1)]; }

namespace {
inline size_t ix(int64_t i, int64_t j) { return (size_t)(i - 1) + (size_t)(j - 1) * MRTM_MXMOL; }
inline size_t ib(int64_t i, int64_t k, int64_t j) { return (size_t)(i - 1) + (size_t)(k - 1) * 7 + (size_t)(j - 1) * 49; }
inline bool is_lc(double xf) { return xf == -1 || xf == -3 || xf == -5; }

struct RawLine {
    int32_t mol, iso, xf, cls, rec, has_self, has_lc, has_brd;
    double xnu0, s0adj, e, alpf, alps, x, deltnu, sdep;
    double lc[16];
    double brd[28];
};
}  // namespace

int stage_lines_host(const int64_t nblm[MRTM_MXMOL], int64_t iim, const int64_t* iso,
                     const double* xnu0, const double* deltnu, const double* e, const double* alps,
                     const double* alpf, const double* x, const double* xg, const double* s0,
                     const double* rmol, const double* sdep, const int32_t* brd_mol_flg,
                     const double* brd_mol_tmp, const double* brd_mol_hw, const double* brd_mol_shft,
                     HostLines& out)
{
    out = HostLines();
    const double radct = kPLANCK * kCLIGHT / kBOLTZ;   // modm.f90:874
    std::vector<RawLine> raw;
    for (int64_t i = 1; i <= MRTM_MXMOL; i++) {
        if (nblm[i - 1] < 0 || nblm[i - 1] > iim) {
            out.error = "nblm out of range";
            return MRTM_EARG;
        }
        int64_t j = 0;
        while (j < nblm[i - 1]) {                        // modm.f90:324
            j = j + 1;
            int64_t jj = j;
            RawLine r;
            std::memset(&r, 0, sizeof r);
            double xgj = xg[ix(i, j)];
            // GET_LNFL stores XG = -IFLG for IFLG in 0..100 (lnfl_mod.f90:44-45, 73-77); LINES and the LSF routines compare XG
            // with -1, -3, -5 only, so every other flag value is an uncoupled line.  Anything else cannot come out of GET_LNFL.
            if (!(xgj <= 0. && xgj >= -100. && xgj == std::floor(xgj))) {
                out.error = "XG outside -(0..100): not a value GET_LNFL stores (lnfl_mod.f90:44-63, 73-77)";
                return MRTM_ELINEFILE;
            }
            if (is_lc(xgj)) {                            // :328-351
                jj = j + 1;
                if (jj + 1 > iim) {
                    out.error = "coupling record beyond line store";
                    return MRTM_ELINEFILE;
                }
                r.has_lc = 1;
                r.lc[0] = xnu0[ix(i, jj)];  r.lc[4] = s0[ix(i, jj)];
                r.lc[1] = alpf[ix(i, jj)];  r.lc[5] = e[ix(i, jj)];
                r.lc[2] = rmol[ix(i, jj)];  r.lc[6] = alps[ix(i, jj)];
                r.lc[3] = x[ix(i, jj)];     r.lc[7] = deltnu[ix(i, jj)];
                if (xgj == -5 && j > 1 && xg[ix(i, j - 1)] == -5) {
                    jj = jj + 1;
                    r.has_self = 1;
                    r.lc[8] = xnu0[ix(i, jj)];   r.lc[12] = s0[ix(i, jj)];
                    r.lc[9] = alpf[ix(i, jj)];   r.lc[13] = e[ix(i, jj)];
                    r.lc[10] = rmol[ix(i, jj)];  r.lc[14] = alps[ix(i, jj)];
                    r.lc[11] = x[ix(i, jj)];     r.lc[15] = deltnu[ix(i, jj)];
                }
            }
            r.mol = (int32_t)i;
            r.rec = (int32_t)j;
            r.xf = (int32_t)xgj;
            r.iso = (int32_t)iso[ix(i, j)];
            if (r.iso < 1 || r.iso > 9) {
                out.error = "isotopologue index outside 1..9 (scor(42,9), SMASS(39,9))";
                return MRTM_ELINEFILE;
            }
            r.xnu0 = xnu0[ix(i, j)];
            // S0_adj, modm.f90:372 (frequency and layer independent)
            r.s0adj = s0[ix(i, j)] * (r.xnu0 * (1.0 - std::exp(-(radct * r.xnu0 / kT0))));
            r.e = e[ix(i, j)];
            r.alpf = alpf[ix(i, j)];
            r.alps = alps[ix(i, j)];
            if (i == 1 && r.alps == 0.) r.alps = 5 * r.alpf;   // HALFWHM_C fix-up, modm.f90:841
            r.x = x[ix(i, j)];
            r.deltnu = deltnu[ix(i, j)];
            r.sdep = sdep[ix(i, j)];
            if (i <= MRTM_MXBRDMOL && brd_mol_flg) {
                int any = 0;
                for (int k = 1; k <= 7; k++) {
                    r.brd[k - 1] = (double)brd_mol_flg[ib(i, k, j)];
                    r.brd[7 + k - 1] = brd_mol_hw[ib(i, k, j)];
                    r.brd[14 + k - 1] = brd_mol_tmp[ib(i, k, j)];
                    r.brd[21 + k - 1] = brd_mol_shft[ib(i, k, j)];
                    any |= brd_mol_flg[ib(i, k, j)] != 0;
                }
                r.has_brd = any;
            }
            if (i == 7) {
                r.cls = !is_lc(xgj) ? CLS_O2 : (r.xf == -1 ? CLS_O2_LC1 : CLS_O2_LC35);
            } else if (i == 2 || is_lc(xgj)) {
                r.cls = CLS_GENERAL;
            } else {
                r.cls = CLS_PED;
            }
            raw.push_back(r);
            j = jj;                                      // :434
        }
    }

    // staged order: molecule, class, xnu0 (stable keeps file order among equal centres)
    std::vector<int64_t> perm(raw.size());
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
        if (raw[a].mol != raw[b].mol) return raw[a].mol < raw[b].mol;
        if (raw[a].cls != raw[b].cls) return raw[a].cls < raw[b].cls;
        return raw[a].xnu0 < raw[b].xnu0;
    });

    const int64_t n = (int64_t)raw.size();
    out.n = n;
    out.n_pad = ((n + 8 + 7) / 8) * 8;
    auto rs_i = [&](std::vector<int32_t>& v) { v.assign(out.n_pad, 0); };
    auto rs_d = [&](std::vector<double>& v) { v.assign(out.n_pad, 0.0); };
    rs_i(out.mol); rs_i(out.iso); rs_i(out.xf); rs_i(out.cls); rs_i(out.sidx); rs_i(out.lcidx);
    rs_i(out.brdidx); rs_i(out.rec); rs_i(out.segidx);
    rs_d(out.xnu0); rs_d(out.s0adj); rs_d(out.e); rs_d(out.alpf); rs_d(out.alps); rs_d(out.x);
    rs_d(out.deltnu); rs_d(out.sdep); rs_d(out.mass);
    out.key.assign(out.n_pad, 0);
    for (int m = 0; m < MRTM_MXMOL; m++) out.mol_slot[m] = -1;

    std::map<int32_t, int32_t> scor_slot;
    for (int64_t q = 0; q < n; q++) {
        const RawLine& r = raw[perm[q]];
        out.mol[q] = r.mol; out.iso[q] = r.iso; out.xf[q] = r.xf; out.cls[q] = r.cls; out.rec[q] = r.rec;
        out.xnu0[q] = r.xnu0; out.s0adj[q] = r.s0adj; out.e[q] = r.e; out.alpf[q] = r.alpf;
        out.alps[q] = r.alps; out.x[q] = r.x; out.deltnu[q] = r.deltnu; out.sdep[q] = r.sdep;
        out.mass[q] = smass(r.mol, r.iso);
        out.key[q] = line_key(r.mol, r.rec);
        int32_t sc = (r.mol - 1) + (r.iso - 1) * MRTM_NSCOR1;
        auto it = scor_slot.find(sc);
        if (it == scor_slot.end()) {
            it = scor_slot.emplace(sc, (int32_t)out.scor_index.size()).first;
            out.scor_index.push_back(sc);
        }
        out.sidx[q] = it->second;
        out.lcidx[q] = -1;
        if (r.has_lc) {
            out.lcidx[q] = (int32_t)(out.lc.size() / 16);
            out.lc.insert(out.lc.end(), r.lc, r.lc + 16);
            out.lc_self.push_back(r.has_self);
        }
        out.brdidx[q] = -1;
        if (r.has_brd) {
            out.brdidx[q] = (int32_t)(out.brd.size() / 28);
            out.brd.insert(out.brd.end(), r.brd, r.brd + 28);
            for (int k = 0; k < 7; k++)
                if (r.brd[k] != 0.)
                    out.max_abs_brd_dshift = std::max(out.max_abs_brd_dshift, std::fabs(r.brd[21 + k] - r.deltnu));
        }
        out.max_abs_deltnu = std::max(out.max_abs_deltnu, std::fabs(r.deltnu));
    }
    // padding lines sit far outside any window and carry no strength
    out.dopf.assign(out.n_pad, 0.0);
    for (int64_t q = 0; q < n; q++)
        out.dopf[q] = std::sqrt(2. * std::log(2.) * (kBOLTZ / (out.mass[q] / kAVOGAD))) / kCLIGHT;
    for (int64_t q = n; q < out.n_pad; q++) {
        out.xnu0[q] = 1.0e30;
        out.mol[q] = 0;
        out.mass[q] = 1.0;
        out.iso[q] = 1;
        out.lcidx[q] = -1;
        out.brdidx[q] = -1;
    }

    // segments
    int64_t q = 0;
    while (q < n) {
        int64_t q1 = q;
        while (q1 < n && out.mol[q1] == out.mol[q] && out.cls[q1] == out.cls[q]) q1++;
        Segment s;
        std::memset(&s, 0, sizeof s);
        s.mol = out.mol[q];
        s.cls = out.cls[q];
        s.begin = (int32_t)q;
        s.end = (int32_t)q1;
        s.count_all = (int32_t)(q1 - q);
        double xmax = 0., mmin = 1e30;
        for (int64_t u = q; u < q1; u++) {
            s.hash_all += out.key[u];
            xmax = std::max(xmax, std::fabs(out.xnu0[u]));
            mmin = std::min(mmin, out.mass[u]);
        }
        // HALFWHM_D (modm.f90:453) upper bound: 100*AD <= vfac*sqrt(T)
        s.vfac = 100. * ((xmax + 1.0) / kCLIGHT) * std::sqrt(2. * std::log(2.) * (kBOLTZ / (mmin / kAVOGAD))) * (1. + 1e-6);
        s.vrate = 100. * (1.0 / kCLIGHT) * std::sqrt(2. * std::log(2.) * (kBOLTZ / (mmin / kAVOGAD))) * (1. + 1e-6);
        if ((int)out.segments.size() >= kMaxSegments) {
            out.error = "too many (molecule,class) segments";
            return MRTM_EARG;
        }
        for (int64_t u = q; u < q1; u++) out.segidx[u] = (int32_t)out.segments.size();
        if (out.mol_slot[s.mol - 1] < 0) {
            if ((int)out.slot_mol.size() >= kMaxSlots) {
                out.error = "more than kMaxSlots molecules own lines";
                return MRTM_EARG;
            }
            out.mol_slot[s.mol - 1] = (int32_t)out.slot_mol.size();
            out.slot_mol.push_back(s.mol);
        }
        s.slot = out.mol_slot[s.mol - 1];
        out.segments.push_back(s);
        q = q1;
    }
    return MRTM_OK;
}

}  // namespace mrtm
