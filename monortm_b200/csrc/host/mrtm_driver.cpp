// mrtm_driver.cpp -- the harness side of SURVEY 8f-1: everything PROGRAM MONORTM does around the
// hot path for layer input (IATM=0), restated in C++ above the C ABI so that the reference's own
// fixtures (run/in/MONORTM.IN_*, MONORTM_PROF.IN_*) run end to end without a Fortran compiler and
// produce a MONORTM.OUT with the reference's record formats:
//   read_control   <- RDLBLINP            src/monortm_sub.F90:33-423  (records 1.1-1.4; IATM=1 needs LBLATM: out of scope)
//   emisfn/reflfn  <- EMISFN/REFLFN/LINTCO/READEM/READRF/EMISS_REFLEC  :1-31, 426-516
//   count_profiles <- GETPROFNUMBER       :869-920
//   ProfReader     <- the MONORTM_PROF.IN block of PROGRAM MONORTM     src/monortm.f90:380-488 (formats :599-612)
//   profil_scal    <- profil_scal_sub (the scaling itself, not its log tables)   src/monortm_sub.F90:937-1045
//   integr         <- INTEGR              :831-845
//   storeout       <- STOREOUT            :519-787 (formats 11/21/31 :780-783; netCDF branch not built)
//   run_monortm    <- the per-profile loop of PROGRAM MONORTM          src/monortm.f90:283-588
// In production the unchanged Fortran host does all of this and calls the hot path through
// monortm_b200/shim/monortm_gpu_shim.f90; nothing here computes optical depths or radiances --
// run_monortm hands every profile to mrtm_profiles (the GPU) and fails if no device is present.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../mrtm_internal.h"
#include "fortio.h"

using fortio::IoError;
using fortio::Record;

namespace {

constexpr int kMXMOL = MRTM_MXMOL;
constexpr int kNMAXCO = 4040;            // PARAMETER (NMAXCO=4040), monortm_sub.F90:4
constexpr int64_t kNWNMX = 80000;        // RTMmono.f90:10

thread_local std::string g_host_err;

int fail(int code, const std::string& msg)
{
    g_host_err = msg;
    return code;
}

struct Stop : std::runtime_error {       // a reference STOP
    int code;
    Stop(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- line-oriented formatted unit ----------------------------------------------------------------
struct Unit {
    std::vector<std::string> lines;
    size_t next = 0;
    bool open(const std::string& path)
    {
        std::ifstream f(path);
        if (!f) return false;
        std::string l;
        while (std::getline(f, l)) lines.push_back(l);
        return true;
    }
    bool eof() const { return next >= lines.size(); }
    Record read()
    {
        if (eof()) throw Stop(MRTM_EIO, "end of file");
        return Record(lines[next++]);
    }
};

struct CoefTable {                       // COMMON /EMSFIN/ , /RFLTIN/
    double v1 = 0, v2 = 0, dv = 0;
    int64_t nlim = 0;
    std::vector<double> z;
};

struct Control {
    int64_t ihirac = 0, icntnm = 0, iemit = 0, iplot = 0, iatm = 0, iod = 0, ixsect = 0, ispd = 0, ibrd = 0;
    double cntnm[7] = {1, 1, 1, 1, 1, 1, 1};
    double v1 = 0, v2 = 0, dvset = 0;
    std::vector<double> wn;
    double tmpbnd = 0, bndemi[3] = {0, 0, 0}, bndrfl[3] = {0, 0, 0};
    int64_t nmol_scal = 0;
    char hmol_scal[64];
    double xmol_scal[64];
    CoefTable emis, refl;
    std::string xid;
    std::vector<std::string> warnings;   // what the reference prints to stdout
    std::string log;                     // what RDLBLINP writes to IPR (MONORTM.LOG)
};

// applyCntnmCombo, src/CntnmFactors.f90:143-186
void apply_cntnm_combo(int64_t icntnm, double f[7])
{
    for (int i = 0; i < 7; i++) f[i] = 1.;
    switch (icntnm) {
    case 0: for (int i = 0; i < 7; i++) f[i] = 0.; break;
    case 1: break;
    case 2: f[0] = 0.; break;
    case 3: f[1] = 0.; break;
    case 4: f[0] = 0.; f[1] = 0.; break;
    case 5: f[6] = 0.; break;
    default: throw Stop(MRTM_EARG, "applyCntnmCombo: invalid ICNTNM (CntnmFactors.f90:180)");
    }
}

void read_coef_table(const std::string& path, CoefTable& t, const char* who)
{
    Unit u;
    if (!u.open(path)) throw Stop(MRTM_EIO, std::string(" EXIT; ERROR OPENING ") + path);
    try {
        Record r = u.read();                                 // 900 FORMAT (3E10.3,5X,I5)
        t.v1 = r.F(10, 3); t.v2 = r.F(10, 3); t.dv = r.F(10, 3);
        r.X(5);
        t.nlim = r.I(5);
        if (t.nlim < 0 || t.nlim > kNMAXCO) throw IoError("table length");
        t.z.resize((size_t)t.nlim);
        for (int64_t i = 0; i < t.nlim; i++) t.z[(size_t)i] = u.read().F(15, 7);     // 910 FORMAT (E15.7)
    } catch (const std::exception&) {
        throw Stop(MRTM_EIO, std::string("INCONSISTENT DATA OR ERROR OPENING IN ") + who);
    }
}

// list-directed READ of n reals (record 1.2a)
void read_list(Unit& u, int n, double* out)
{
    int got = 0;
    while (got < n) {
        Record r = u.read();
        std::string s = r.text();
        for (char& c : s) if (c == ',') c = ' ';
        std::istringstream is(s);
        std::string tok;
        while (got < n && (is >> tok)) {
            if (tok == "/") return;
            for (char& c : tok) if (c == 'd' || c == 'D') c = 'e';
            char* end = nullptr;
            double v = std::strtod(tok.c_str(), &end);
            if (end == tok.c_str()) throw Stop(MRTM_EIO, " EXIT; ERROR READING : record 1.2a");
            out[got++] = v;
        }
    }
}

std::string sfmt(const char* f, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    std::vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

// RDLBLINP, src/monortm_sub.F90:33-423
void read_control(const std::string& filein, const std::string& dir, int64_t nwnmx, Control& c)
{
    Unit u;
    if (!u.open(filein)) throw Stop(MRTM_EIO, " EXIT; ERROR OPENING :" + filein);
    try {
        // record 1.1: skip to the '$' line (:136-144)
        for (;;) {
            if (u.eof()) throw Stop(MRTM_EIO, " EXIT; EOF ON :" + filein);
            Record r = u.read();
            const std::string& t = r.text();
            if (!t.empty() && t[0] == '%') throw Stop(MRTM_EIO, "-END OF FILE:" + filein);
            if (!t.empty() && t[0] == '$') { c.xid = fortio::fmt_A(80, t.size() > 1 ? t.substr(1, 80) : ""); break; }
        }
        {   // record 1.2, FORMAT 925 (:409)
            Record r = u.read();
            r.X(4); c.ihirac = r.I(1); r.X(9); c.icntnm = r.I(1); r.X(9); c.iemit = r.I(1);
            r.X(14); c.iplot = r.I(1); r.X(9); c.iatm = r.I(1); r.X(14); c.iod = r.I(1);
            r.X(4); c.ixsect = r.I(1); r.X(16); c.ispd = r.I(4); c.ibrd = r.I(4);
        }
        // IPR echo, FORMATs 935/940 (:154-161)
        static const char* idcntl[9] = {" HIRAC", " CNTNM", " EMISS", "  PLOT", "  IATM", "   IOD", " XSECT", "  ISPD", "  IBRD"};
        for (int i = 0; i < 9; i++) c.log += std::string(idcntl[i]) + "   ";
        c.log += "\n " + fortio::fmt_I(4, c.ihirac);
        for (int64_t v : {c.icntnm, c.iemit, c.iplot, c.iatm, c.iod, c.ixsect, c.ispd, c.ibrd}) c.log += fortio::fmt_I(9, v);
        c.log += "\n";
        if (c.iemit != 1) c.warnings.push_back("WARNING: IEMIT IS IGNORED IN MONORTM; IT IS SET INTERNALLY TO ONE");
        if (c.iplot != 1) c.warnings.push_back("WARNING: IPLOT MUST BE SET TO 1 TO OUTPUT TBs");
        if (c.iod == 1) c.warnings.push_back("IOD FLAG SET TO OUTPUT LAYER OPTICAL DEPTHS");

        if (c.icntnm == 6) read_list(u, 7, c.cntnm);         // record 1.2a (:181-188)
        else apply_cntnm_combo(c.icntnm, c.cntnm);
        if (c.iemit == 2) u.read();                          // record 1.2.1 (INFLAG,IOTFLG,JULDAT: unused)
        if (c.iemit == 3) throw Stop(MRTM_EARG, "CURRENTLY MONORTM DOES NOT HANDLE DERIVATIVES");

        int64_t nwn = 0;
        if (c.ihirac + c.iemit + c.iatm > 0) {               // record 1.3, FORMAT 970 (:203)
            Record r = u.read();
            c.v1 = r.F(10, 3); c.v2 = r.F(10, 3);
            double sample = r.F(10, 3);
            c.dvset = r.F(10, 3);
            double alfal0 = r.F(10, 3), avmass = r.F(10, 3), dptmin = r.F(10, 3), dptfac = r.F(10, 3);
            r.X(4);
            int64_t ilnflg = r.I(1);
            r.X(5);
            double dvout = r.F(10, 3);
            c.nmol_scal = r.I(5);
            if (c.nmol_scal > 0) {                           // profile scaling records (:209-215)
                if (c.nmol_scal > 38) throw Stop(MRTM_EARG, " nmol_scal .gt. 38 ");
                Record h = u.read();                         // 9701 FORMAT (64a1)
                for (int m = 0; m < c.nmol_scal; m++) c.hmol_scal[m] = h.A(1)[0];
                // 9702 FORMAT (7e15.7,/,(8e15.7,/)): a '/' that format control passes before it runs out of list
                // items consumes a record even when nothing is read from it
                int m = 0, k = 0;
                Record x = u.read();
                for (; k < 7 && m < c.nmol_scal; k++) c.xmol_scal[m++] = x.F(15, 7);
                if (k == 7 && m == c.nmol_scal && !u.eof()) u.next++;
                while (m < c.nmol_scal) {
                    Record y = u.read();
                    int j = 0;
                    for (; j < 8 && m < c.nmol_scal; j++) c.xmol_scal[m++] = y.F(15, 7);
                    if (j == 8 && !u.eof()) u.next++;
                }
            }
            if (sample > 0) c.warnings.push_back("WARNING: SAMPLE IS IGNORED IN MONORTM");
            if (alfal0 > 0) c.warnings.push_back("WARNING: ALFAL0 IS IGNORED IN MONORTM");
            if (avmass > 0) c.warnings.push_back("WARNING: AVMASS IS IGNORED IN MONORTM");
            if (dptmin > 0) c.warnings.push_back("WARNING: DPTMIN IS IGNORED IN MONORTM");
            if (dptfac > 0) c.warnings.push_back("WARNING: DPTFAC IS IGNORED IN MONORTM");
            if (ilnflg > 0) throw Stop(MRTM_EARG, "STOP: ILNFLG MUST BE 0 FOR MONORTM");
            if (dvout > 0) c.warnings.push_back("WARNING: DVOUT IS IGNORED IN MONORTM");
            if (c.dvset <= 0. && c.v1 != c.v2 && c.v1 > 0. && c.v2 > 0.)
                throw Stop(MRTM_EARG, "MONORTM REQUIRES POSITIVE DVSET, OR (V1=V2 AND DVSET=0) IF YOU WANT ONLY ONE FREQ PROCESSED");
            if (c.v1 < 0. || c.v2 < 0.) {                    // record 1.3.1 / 1.3.2 (:264-277)
                nwn = u.read().I(8);
                if (nwn > nwnmx) throw Stop(MRTM_EARG, sfmt("STOP: NUMBER OF WAVENUMBERS EXCEEDS LIMIT. %lld %lld", (long long)nwn, (long long)nwnmx));
                if (nwn < 0) nwn = 0;
                c.wn.resize((size_t)nwn);
                for (int64_t i = 0; i < nwn; i++) c.wn[(size_t)i] = u.read().F(19, 7);
                c.dvset = 0.;
            } else {
                if (c.dvset != 0.) {
                    nwn = (int64_t)std::llround(((c.v2 - c.v1) / c.dvset) + 1.);     // NINT
                    if (nwn > nwnmx) throw Stop(MRTM_EARG, sfmt("STOP: NUMBER OF WAVENUMBERS EXCEEDS LIMIT. %lld %lld", (long long)nwn, (long long)nwnmx));
                    c.wn.resize((size_t)std::max<int64_t>(nwn, 0));
                    for (int64_t j = 1; j <= nwn; j++) c.wn[(size_t)(j - 1)] = c.v1 + (double)(j - 1) * c.dvset;
                } else {
                    if (c.v1 != c.v2) throw Stop(MRTM_EARG, "AMBIGUITY IN THE WAVENUMBER");
                    c.wn.assign(1, c.v1);
                }
            }
        }
        {   // record 1.4, FORMAT 970 (:305-309)
            Record r = u.read();
            c.tmpbnd = r.F(10, 3);
            for (int i = 0; i < 3; i++) c.bndemi[i] = r.F(10, 3);
            for (int i = 0; i < 3; i++) c.bndrfl[i] = r.F(10, 3);
            // FORMAT 985 (:311-315); the last A1 item has no list element, so output stops before it
            c.log += "\n\n\n\n\n0*********** BOUNDARY PROPERTIES ***********\n";
            c.log += "0 TBOUND   = " + fortio::fmt_F(12, 4, c.tmpbnd) + "     BOUNDARY EMISSIVITY   = ";
            for (int i = 0; i < 3; i++) c.log += fortio::fmt_1PE(11, 3, c.bndemi[i]);
            c.log += "\n0" + std::string(29, ' ') + "BOUNDARY REFLECTIVITY = ";
            for (int i = 0; i < 3; i++) c.log += fortio::fmt_1PE(11, 3, c.bndrfl[i]);
            c.log += "\n0" + std::string(29, ' ') + " SURFACE REFLECTIVITY = \n";
        }
        const double xvmid = (c.v1 + c.v2) / 2.;
        if (c.bndemi[0] < 0) {
            read_coef_table(dir + "in/EMISSION", c.emis, "READEM");
        } else {
            double emitst = c.bndemi[0] + c.bndemi[1] * xvmid + c.bndemi[2] * xvmid * xvmid;
            if (emitst < 0. || emitst > 1.) throw Stop(MRTM_EARG, "BNDEMI OUTSIDE PHYSICAL RANGE");
        }
        if (c.bndrfl[0] < 0) {
            read_coef_table(dir + "in/REFLECTION", c.refl, "READRF");
        } else {
            double reftst = c.bndrfl[0] + c.bndrfl[1] * xvmid + c.bndrfl[2] * xvmid * xvmid;
            if (reftst < 0. || reftst > 1.) throw Stop(MRTM_EARG, "BNDRFL OUTSIDE PHYSICAL RANGE");
        }
    } catch (const IoError& e) {
        throw Stop(MRTM_EIO, " EXIT; ERROR READING :" + filein + " (" + e.what() + ")");
    }
}

// LINTCO + EMISFN / REFLFN, src/monortm_sub.F90:426-503
double coef_fn(double vi, const double abc[3], const CoefTable& t, const char* who)
{
    const double a = abc[0], b = abc[1], cc = abc[2];
    if (a < 0.) {
        const int64_t nel = (int64_t)((vi - t.v1) / t.dv);   // INT(): truncation
        if (nel <= 0 || nel >= t.nlim) throw Stop(MRTM_EARG, std::string("ERROR IN ") + who);
        const double v1a = t.v1 + t.dv * (double)nel, v1b = t.v1 + t.dv * (double)(nel + 1);
        const double z1 = t.z[(size_t)(nel - 1)], z2 = t.z[(size_t)nel];
        const double zdel = (z2 - z1) / (v1b - v1a);
        const double zcept = z1 - zdel * v1a;
        return zdel * vi + zcept;
    }
    if (b == 0. && cc == 0.) return a;
    return a + b * vi + cc * vi * vi;
}

// GETPROFNUMBER for IATM=0 (:894-907): every record that reads cleanly under FORMAT 972 counts
int64_t count_profiles(const std::string& fileprof, int64_t ixsect)
{
    Unit u;
    if (!u.open(fileprof)) throw Stop(MRTM_EIO, "ERROR OPENING OR READING FILE in GETPROFNUMBER");
    int64_t nprof = 0;
    while (!u.eof()) {
        Record r = u.read();
        try {
            r.X(1); r.I(1); r.I(3); r.I(5); r.F(10, 6); r.A(8); r.A(8); r.X(4); r.F(8, 2); r.X(4); r.F(8, 2);
            r.X(5); r.F(8, 3); r.X(5); r.I(2);
            nprof++;
        } catch (const IoError&) {
        }
    }
    if (ixsect == 1) nprof /= 2;
    if (nprof == 0) throw Stop(MRTM_EIO, "NO PROFILE FOUND IN GETPROFNUMBER");
    return nprof;
}

struct Profile {
    int64_t iform = 0, nlay = 0, nmol = 0, irt = 0;
    double secnt0 = 0, h1 = 0, h2 = 0, angle = 0;
    std::vector<double> p, t, clw, wbrodl, altz, pz, tz, wkl;   // wkl (39,nlay) column-major
    // IXSECT >= 1 (monortm.f90:491-527): cross-section molecule names (10 characters each) and XAMNT (38, nlay)
    int64_t ixmols = 0;
    std::string xsnames;
    std::vector<double> xamnt;
};

// the MONORTM_PROF.IN block of PROGRAM MONORTM, src/monortm.f90:380-488
struct ProfReader {
    Unit u;
    int64_t ixsect = 0;
    // returns false on END= (label 110)
    bool next(Profile& pr)
    {
        if (u.eof()) return false;
        try {
            Record h = u.read();                              // 925 FORMAT(1X,I1,I3,I5,F10.6,2A8,4X,F8.2,4X,F8.2,5X,F8.3,5X,I2)
            h.X(1);
            pr.iform = h.I(1); pr.nlay = h.I(3); pr.nmol = h.I(5); pr.secnt0 = h.F(10, 6);
            h.A(8); h.A(8); h.X(4);
            pr.h1 = h.F(8, 2); h.X(4); pr.h2 = h.F(8, 2); h.X(5); pr.angle = h.F(8, 3); h.X(5); h.I(2);
            if (pr.nlay <= 0) return false;                   // a blank trailing record: nothing to process
            if (pr.nlay > 603) throw Stop(MRTM_EARG, "NLAYRS exceeds MXLAY=603 (lblparams.f90:28)");
            if (pr.nmol < 0 || pr.nmol > kMXMOL) throw Stop(MRTM_EARG, "NMOL exceeds MXMOL=39 (lblparams.f90:29)");
            if (pr.angle > 90.) pr.irt = 1;
            if (pr.angle < 90.) pr.irt = 3;
            if (pr.angle == 90.) pr.irt = 2;
            const size_t n = (size_t)pr.nlay;
            pr.p.assign(n, 0.); pr.t.assign(n, 0.); pr.clw.assign(n, 0.); pr.wbrodl.assign(n, 0.);
            pr.altz.assign(n + 1, 0.); pr.pz.assign(n + 1, 0.); pr.tz.assign(n + 1, 0.);
            pr.wkl.assign((size_t)kMXMOL * n, 0.);
            for (size_t il = 0; il < n; il++) {
                Record r = u.read();                          // 974/9742 (IFORM=0) | 975/9752 (IFORM=1)
                if (pr.iform == 0) pr.p[il] = r.F(10, 4); else pr.p[il] = r.F(15, 7);
                pr.t[il] = r.F(10, 4);
                r.F(10, 4);                                   // secnt
                r.X(3); r.I(2); r.X(1);                       // ipath
                if (il == 0) {
                    pr.altz[0] = r.F(7, 2); pr.pz[0] = r.F(8, 3); pr.tz[0] = r.F(7, 2);
                } else {
                    r.X(22);
                }
                pr.altz[il + 1] = r.F(7, 2); pr.pz[il + 1] = r.F(8, 3); pr.tz[il + 1] = r.F(7, 2);
                pr.clw[il] = r.F(7, 3);
                double* w = pr.wkl.data() + (size_t)kMXMOL * il;
                Record a = u.read();                          // 978 FORMAT (8E15.7): WKL(1:7), WBRODL
                for (int k = 0; k < 7; k++) w[k] = a.F(15, 7);
                pr.wbrodl[il] = a.F(15, 7);
                for (int64_t k = 7; k < pr.nmol;) {           // WKL(8:NMOL), 8 per record
                    Record b = u.read();
                    for (int j = 0; j < 8 && k < pr.nmol; j++, k++) w[k] = b.F(15, 7);
                }
                // mixing ratio input (:423-483)
                double wdnsty = pr.wbrodl[il], wmxrat = 0.0, wdrair = 0.0;
                for (int64_t m = 2; m <= pr.nmol; m++) {
                    if (w[m - 1] > 1) wdnsty = wdnsty + w[m - 1];
                    else wmxrat = wmxrat + w[m - 1];
                }
                if (pr.wbrodl[il] < 1.0 && pr.wbrodl[il] != 0.0) throw Stop(MRTM_EARG, "STOP (WBRODL must be a column density, monortm.f90:449)");
                if (wdnsty == 0.0 && wmxrat != 0.0) throw Stop(MRTM_EARG, "WMXRAT AND/OR WDNSTY NOT PROPERLY SPECIFIED IN PATH");
                if (wmxrat < 1.0) wdrair = wdnsty / (1.0 - wmxrat);
                else throw Stop(MRTM_EARG, "WMXRAT EXCEEDS 1.0");
                if (w[0] <= 1.0 && w[0] != 0.0 && wdrair == 0.0) throw Stop(MRTM_EARG, "WMXRAT NOT PROPERLY SPECIFIED IN PATH");
                for (int64_t m = 1; m <= pr.nmol; m++)
                    if (w[m - 1] < 1.) w[m - 1] = w[m - 1] * wdrair;
            }
            if (ixsect >= 1) {                                // monortm.f90:491-527
                Record a = u.read();                          // 930 FORMAT (I5,5X,I5): IXMOLS, IXSBIN
                pr.ixmols = a.I(5);
                if (pr.ixmols < 1 || pr.ixmols > 38) throw Stop(MRTM_EARG, "IXMOLS out of 1..38 (MX_XS, lblparams.f90:29)");
                pr.xsnames.assign((size_t)pr.ixmols * 10, ' ');
                for (int64_t i = 0; i < pr.ixmols;) {         // XSREAD :1296-1301: (7A10), then (8A10)
                    Record nrec = u.read();
                    const int64_t per = (i == 0) ? 7 : 8;
                    for (int64_t j = 0; j < per && i < pr.ixmols; j++, i++) {
                        std::string nm = nrec.A(10);
                        nm.resize(10, ' ');
                        std::memcpy(&pr.xsnames[(size_t)i * 10], nm.data(), 10);
                    }
                }
                Record hx = u.read();                         // 900 FORMAT (1X,I1,I3,I5,F10.2,15A4)
                hx.X(1); hx.I(1);
                const int64_t nlayxs = hx.I(3), ixmol = hx.I(5);
                if (ixmol == 0) throw Stop(MRTM_EARG, " PATH - IXMOL 0 ");
                if (ixmol != pr.ixmols) throw Stop(MRTM_EARG, " PATH - IXMOL .NE. IXMOLS ");
                if (pr.nlay != nlayxs) throw Stop(MRTM_EARG, " PATH - NLAYRS .NE. NLAYXS ");
                pr.xamnt.assign((size_t)38 * n, 0.);
                for (size_t il = 0; il < n; il++) {
                    u.read();                                 // 910 / 915: PAVX, TAVX, ... (not used)
                    Record x1 = u.read();                     // 978: XAMNT(1:7,L), WBRODX
                    double* xa = pr.xamnt.data() + (size_t)38 * il;
                    for (int k = 0; k < 7; k++) { const double v = x1.F(15, 7); if (k < ixmol) xa[k] = v; }
                    for (int64_t k = 7; k < ixmol;) {         // XAMNT(8:IXMOL,L), one READ
                        Record x2 = u.read();
                        for (int j = 0; j < 8 && k < ixmol; j++, k++) xa[k] = x2.F(15, 7);
                    }
                }
            }
        } catch (const IoError& e) {
            throw Stop(MRTM_EIO, std::string(" EXIT; ERROR READING :MONORTM_PROF.IN (") + e.what() + ")");
        } catch (const Stop& s) {
            if (std::string(s.what()) == "end of file") return false;       // END=110
            throw;
        }
        return true;
    }
};

// the scaling step of profil_scal_sub (src/monortm_sub.F90:960-1045); xmol_scal is updated in place
// exactly as the reference does (so a second profile sees the already-converted factors)
void profil_scal(Control& c, Profile& pr)
{
    const int64_t nmol = pr.nmol, nlay = pr.nlay;
    double wmt[64] = {0};
    for (int64_t m = 0; m < nmol; m++)
        for (int64_t l = 0; l < nlay; l++) wmt[m] = wmt[m] + pr.wkl[(size_t)m + (size_t)kMXMOL * (size_t)l];
    double wsum_brod = 0.;
    for (int64_t l = 0; l < nlay; l++) wsum_brod = wsum_brod + pr.wbrodl[(size_t)l];
    double wsum_drair = nmol >= 22 ? 0. : wsum_brod;
    for (int64_t m = 1; m < nmol; m++) wsum_drair = wsum_drair + wmt[m];
    for (int64_t m = 0; m < c.nmol_scal; m++) {
        const double x = c.xmol_scal[m];
        const char h = c.hmol_scal[m];
        if (h == ' ') c.xmol_scal[m] = 1.;
        if (h == '0') c.xmol_scal[m] = 0.;
        if (h == '1') c.xmol_scal[m] = x;
        if (h == 'C' || h == 'c') c.xmol_scal[m] = x / wmt[m];
        if (h == 'M' || h == 'm') {
            if (wsum_drair > 0.) c.xmol_scal[m] = x / (wmt[m] / wsum_drair);
            else throw Stop(MRTM_EARG, "mixing ratio failure: wsum_drair = 0.");
        }
        if (h == 'P' || h == 'p') {
            if (m == 0) c.xmol_scal[0] = (x / 2.99150e-23) / wmt[0];
            else throw Stop(MRTM_EARG, " (hmol_scal(m).eq.\"P\" .and. m.ne.1) ");
        }
        if (h == 'D' || h == 'd') c.xmol_scal[m] = (x * 2.68678e16) / wmt[m];
        wmt[m] = 0.;
        for (int64_t l = 0; l < nlay; l++) {
            double& w = pr.wkl[(size_t)m + (size_t)kMXMOL * (size_t)l];
            w = w * c.xmol_scal[m];
            wmt[m] = wmt[m] + w;
        }
    }
}

// STOREOUT state that the reference SAVEs at the first profile (:597-609)
struct StoreState {
    int kount = 0;
    int id_mol[kMXMOL];
    std::string cmol[kMXMOL];
};

const char* const kHMOLC[kMXMOL] = {
    "  H2O   ", "  CO2   ", "   O3   ", "  N2O   ", "   CO   ", "  CH4   ", "   O2   ", "   NO   ",
    "  SO2   ", "  NO2   ", "  NH3   ", " HNO3   ", "   OH   ", "   HF   ", "  HCL   ", "  HBR   ",
    "   HI   ", "  CLO   ", "  OCS   ", " H2CO   ", " HOCL   ", "   N2   ", "  HCN   ", " CH3CL  ",
    " H2O2   ", " C2H2   ", " C2H6   ", "  PH3   ", " COF2   ", "  SF6   ", "  H2S   ", " HCOOH  ",
    "  HO2   ", "   O+   ", " ClONO2 ", "   NO+  ", "  HOBr  ", " C2H4   ", " CH3OH  "};

// STOREOUT (:519-680, formats 11/21 :780-782) with the layer sums already formed:
// otot(nwn), odxtot(nwn), otot_by_mol(39,nwn) indexed by molecule number.
// wkl (39,nlay) is IN/OUT like the reference's (row 22 <- wbrodl when nmol < 22 on the first profile).
void storeout_rows(FILE* iot, StoreState& st, int64_t nwn, const double* wn, double* wkl, const double* wbrodl,
                   const double* rad, const double* tb, const double* trtot, int64_t npr, const double* otot,
                   const double* otot_by_mol, const double* odxtot, const double* tmr, double wvcolmn, double clwcolmn,
                   double tmpsfc, const double* reflc, const double* emiss, int64_t nlay, int64_t nmol, double angle)
{
    if (npr == 1) {
        if (nmol < 22)
            for (int64_t l = 0; l < nlay; l++) wkl[21 + (size_t)kMXMOL * (size_t)l] = wbrodl[l];
        st.kount = 0;
        for (int im = 0; im < kMXMOL; im++) {
            double tot = 0.;
            for (int64_t l = 0; l < nlay; l++) tot = tot + wkl[(size_t)im + (size_t)kMXMOL * (size_t)l];
            if (tot > 0) {
                st.id_mol[st.kount] = im + 1;
                st.cmol[st.kount] = kHMOLC[im];
                st.kount++;
            }
        }
    }
    if (st.kount + 2 > 36)
        throw Stop(MRTM_EARG, "STOREOUT: more than 34 molecules with non-zero amounts overflow FORMAT 21 (36E12.4): "
                              "the reference fails with a format/data mismatch on record reversion");
    std::string s;
    s += "MONORTM RESULTS:\n----------------\n";
    s += fortio::fmt_A(5, "NWN :") + fortio::fmt_I(8, nwn) + std::string(101, ' ') + fortio::fmt_A(42, "Molecular Optical Depths -->") + "\n";
    const bool giga = nwn > 0 && wn[0] < 100;                 // (:619; otherwise never set -> .false.)
    const std::string wnunits = giga ? "FREQ(GHz)   " : "FREQ(cm-1)  ";      // character*12
    // FORMAT 11 (a5,a10,2a11,a22,a8,2a8,3a8,a9,36a12)
    s += fortio::fmt_A(5, "PROF ") + fortio::fmt_A(10, wnunits) + fortio::fmt_A(11, "BT(K) ") + fortio::fmt_A(11, "TMR(K)") +
         fortio::fmt_A(22, "  RAD(W/cm2_ster_cm-1)") + fortio::fmt_A(8, "TRANS") + fortio::fmt_A(8, "PWV") + fortio::fmt_A(8, "CLW") +
         fortio::fmt_A(8, "TBOUND") + fortio::fmt_A(8, "EMIS") + fortio::fmt_A(8, "REFL") + fortio::fmt_A(9, "ANGLE") +
         fortio::fmt_A(12, "TOTAL_OD");
    for (int ik = 0; ik < st.kount; ik++) s += fortio::fmt_A(12, st.cmol[ik]);
    s += fortio::fmt_A(12, "XSEC_OD") + "\n";
    std::fputs(s.c_str(), iot);
    for (int64_t iw = 0; iw < nwn; iw++) {
        const double freq = giga ? wn[iw] * mrtm::kCLIGHT / 1.E9 : wn[iw];
        // FORMAT 21 (i5,f10.3,2f11.5,1p,E21.9,0p,f9.5,2f8.4,3f8.2,f9.3,1p,36E12.4)
        std::string r = fortio::fmt_I(5, npr) + fortio::fmt_F(10, 3, freq) + fortio::fmt_F(11, 5, tb[iw]) + fortio::fmt_F(11, 5, tmr[iw]) +
                        fortio::fmt_1PE(21, 9, rad[iw]) + fortio::fmt_F(9, 5, trtot[iw]) + fortio::fmt_F(8, 4, wvcolmn) +
                        fortio::fmt_F(8, 4, clwcolmn) + fortio::fmt_F(8, 2, tmpsfc) + fortio::fmt_F(8, 2, emiss[iw]) +
                        fortio::fmt_F(8, 2, reflc[iw]) + fortio::fmt_F(9, 3, angle) + fortio::fmt_1PE(12, 4, otot[iw]);
        for (int ik = 0; ik < st.kount; ik++)
            r += fortio::fmt_1PE(12, 4, otot_by_mol[(size_t)(st.id_mol[ik] - 1) + (size_t)kMXMOL * (size_t)iw]);
        r += fortio::fmt_1PE(12, 4, odxtot[iw]) + "\n";
        std::fputs(r.c_str(), iot);
    }
}

// Ew.d with scale factor 0: 0.ddddE+ee
std::string fmt_E0(int w, int d, double v)
{
    std::string out;
    if (fortio::special(v, w, out)) return out;
    char buf[128];
    std::snprintf(buf, sizeof buf, "%.*E", d - 1, v);        // x.yyyE+ee with d significant digits
    std::string s = buf;
    const bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    const size_t e = s.find('E');
    std::string digits = s.substr(0, 1) + s.substr(2, e - 2);
    int ex = std::atoi(s.c_str() + e + 1);
    if (v != 0.) ex += 1;
    char eb[16];
    if (std::abs(ex) < 100) std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    else std::snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
    std::string body = std::string(neg ? "-" : "") + "0." + digits + eb;
    if ((int)body.size() > w) body.erase(neg ? 1 : 0, 1);    // optional leading zero
    return fortio::rjust(body, w);
}

// the IOD=1 branch of STOREOUT (:669-686): one ODmono_prfNNNN_layNNNN file per layer
void write_odmono(const std::string& dir, int64_t npr, int64_t nwn, const double* wn, int64_t nlay, const double* o /* (nwn,nlay) */)
{
    const bool giga = nwn > 0 && wn[0] < 100;
    const std::string wnunits = giga ? "FREQ(GHz)   " : "FREQ(cm-1)  ";
    for (int64_t j = 1; j <= nlay; j++) {
        const std::string name = dir + sfmt("ODmono_prf%04lld_lay%04lld", (long long)npr, (long long)j);
        FILE* f = std::fopen(name.c_str(), "w");
        if (!f) throw Stop(MRTM_EIO, "ERROR OPENING FILE:" + name);
        std::string s = fortio::fmt_A(5, "NWN :") + fortio::fmt_I(8, nwn) + "\n";
        s += fortio::fmt_A(10, wnunits) + fortio::fmt_A(10, " LAYER_OD") + "\n";
        std::fputs(s.c_str(), f);
        for (int64_t iw = 0; iw < nwn; iw++) {
            const double freq = giga ? wn[iw] * mrtm::kCLIGHT / 1.E9 : wn[iw];
            std::string r = fortio::fmt_F(10, 3, freq) + fmt_E0(12, 4, o[(size_t)iw + (size_t)nwn * (size_t)(j - 1)]) + "\n";
            std::fputs(r.c_str(), f);
        }
        std::fclose(f);
    }
}

std::string dir_of(const std::string& p)
{
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

StoreState g_store_state;      // the SAVEd variables of STOREOUT for the stand-alone ABI entry

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" const char* mrtm_host_last_error(void) { return g_host_err.c_str(); }
namespace mrtm_hostdrv { int set_host_error(int code, const std::string& msg) { return fail(code, msg); } }   // for mrtm_host.cpp

extern "C" int mrtm_host_read_control(const char* filein, int64_t nwnmx, mrtm_control* out)
{
    if (!filein || !out) return fail(MRTM_EARG, "null argument");
    std::memset(out, 0, sizeof *out);
    Control c;
    try {
        read_control(filein, dir_of(filein), nwnmx > 0 ? nwnmx : kNWNMX, c);
    } catch (const Stop& s) {
        return fail(s.code, s.what());
    }
    out->ihirac = c.ihirac; out->icntnm = c.icntnm; out->iemit = c.iemit; out->iplot = c.iplot; out->iatm = c.iatm;
    out->iod = c.iod; out->ixsect = c.ixsect; out->ispd = c.ispd; out->ibrd = c.ibrd;
    for (int i = 0; i < 7; i++) out->cntnm[i] = c.cntnm[i];
    out->v1 = c.v1; out->v2 = c.v2; out->dvset = c.dvset;
    out->nwn = (int64_t)c.wn.size();
    out->wn = (double*)std::malloc(sizeof(double) * std::max<size_t>(c.wn.size(), 1));
    if (!out->wn) return fail(MRTM_ENOMEM, "out of memory");
    std::copy(c.wn.begin(), c.wn.end(), out->wn);
    out->tmpbnd = c.tmpbnd;
    for (int i = 0; i < 3; i++) { out->bndemi[i] = c.bndemi[i]; out->bndrfl[i] = c.bndrfl[i]; }
    out->nmol_scal = c.nmol_scal;
    for (int m = 0; m < c.nmol_scal && m < 64; m++) { out->hmol_scal[m] = c.hmol_scal[m]; out->xmol_scal[m] = c.xmol_scal[m]; }
    return MRTM_OK;
}

extern "C" void mrtm_host_free_control(mrtm_control* c)
{
    if (c && c->wn) { std::free(c->wn); c->wn = nullptr; }
}

extern "C" int mrtm_host_count_profiles(const char* fileprof, int64_t ixsect, int64_t* nprof)
{
    if (!fileprof || !nprof) return fail(MRTM_EARG, "null argument");
    try {
        *nprof = count_profiles(fileprof, ixsect);
    } catch (const Stop& s) {
        return fail(s.code, s.what());
    }
    return MRTM_OK;
}

extern "C" int mrtm_host_read_profile(const char* fileprof, int64_t index, int64_t maxlay, int64_t* iform, int64_t* nlay,
                                      int64_t* nmol, int64_t* irt, double* secnt0, double* h1, double* h2, double* angle,
                                      double* p, double* t, double* clw, double* wbrodl, double* altz, double* pz,
                                      double* tz, double* wkl)
{
    if (!fileprof || index < 0) return fail(MRTM_EARG, "bad argument");
    ProfReader rd;
    if (!rd.u.open(fileprof)) return fail(MRTM_EIO, std::string(" EXIT; ERROR OPENING :") + fileprof);
    Profile pr;
    try {
        for (int64_t i = 0; i <= index; i++)
            if (!rd.next(pr)) return fail(MRTM_EIO, "profile index beyond the end of the file");
    } catch (const Stop& s) {
        return fail(s.code, s.what());
    }
    if (pr.nlay > maxlay) return fail(MRTM_ENOMEM, "maxlay too small");
    *iform = pr.iform; *nlay = pr.nlay; *nmol = pr.nmol; *irt = pr.irt;
    *secnt0 = pr.secnt0; *h1 = pr.h1; *h2 = pr.h2; *angle = pr.angle;
    const size_t n = (size_t)pr.nlay;
    std::copy(pr.p.begin(), pr.p.end(), p); std::copy(pr.t.begin(), pr.t.end(), t);
    std::copy(pr.clw.begin(), pr.clw.end(), clw); std::copy(pr.wbrodl.begin(), pr.wbrodl.end(), wbrodl);
    std::copy(pr.altz.begin(), pr.altz.end(), altz); std::copy(pr.pz.begin(), pr.pz.end(), pz);
    std::copy(pr.tz.begin(), pr.tz.end(), tz);
    std::copy(pr.wkl.begin(), pr.wkl.begin() + (ptrdiff_t)(kMXMOL * n), wkl);
    return MRTM_OK;
}

extern "C" int mrtm_host_emiss_reflec(const mrtm_control* c, const char* dir, int64_t nwn, const double* wn, double* emiss, double* reflc)
{
    if (!c || !wn || !emiss || !reflc) return fail(MRTM_EARG, "null argument");
    try {
        CoefTable te, tr;
        const std::string d = dir ? std::string(dir) : std::string();
        if (c->bndemi[0] < 0) read_coef_table(d + "in/EMISSION", te, "READEM");
        if (c->bndrfl[0] < 0) read_coef_table(d + "in/REFLECTION", tr, "READRF");
        for (int64_t j = 0; j < nwn; j++) {                   // EMISS_REFLEC (:506-516)
            reflc[j] = coef_fn(wn[j], c->bndrfl, tr, "REFLFN");
            emiss[j] = coef_fn(wn[j], c->bndemi, te, "EMISFN");
        }
    } catch (const Stop& s) {
        return fail(s.code, s.what());
    }
    return MRTM_OK;
}

extern "C" int mrtm_host_storeout(const char* fileout, int append, int64_t nwn, const double* wn, double* wkl,
                                  const double* wbrodl, const double* rad, const double* tb, const double* trtot,
                                  int64_t npr, const double* o, const double* o_by_mol, const double* oc,
                                  const double* odxsec, const double* tmr, double wvcolmn, double clwcolmn,
                                  double tmpsfc, const double* reflc, const double* emiss, int64_t nlay, int64_t nmol,
                                  double angle, int64_t iod)
{
    if (!fileout || !wn || !wkl || !wbrodl || !rad || !tb || !trtot || !o || !o_by_mol || !oc || !tmr || !reflc || !emiss)
        return fail(MRTM_EARG, "null argument");
    FILE* f = std::fopen(fileout, append ? "a" : "w");
    if (!f) return fail(MRTM_EIO, std::string("ERROR OPENING FILE:") + fileout);
    try {
        // the layer sums of STOREOUT (:643-656), layers in index order
        std::vector<double> otot((size_t)nwn, 0.), odxtot((size_t)nwn, 0.), obm((size_t)kMXMOL * (size_t)nwn, 0.);
        for (int64_t iw = 0; iw < nwn; iw++)
            for (int64_t j = 0; j < nlay; j++) {
                otot[(size_t)iw] = otot[(size_t)iw] + o[(size_t)iw + (size_t)nwn * (size_t)j];
                if (odxsec) odxtot[(size_t)iw] = odxtot[(size_t)iw] + odxsec[(size_t)iw + (size_t)nwn * (size_t)j];
                for (int im = 0; im < kMXMOL; im++) {
                    const size_t idx = (size_t)iw + (size_t)nwn * ((size_t)im + (size_t)kMXMOL * (size_t)j);
                    double& s = obm[(size_t)im + (size_t)kMXMOL * (size_t)iw];
                    s = s + o_by_mol[idx] + oc[idx];
                }
            }
        storeout_rows(f, g_store_state, nwn, wn, wkl, wbrodl, rad, tb, trtot, npr, otot.data(), obm.data(), odxtot.data(),
                      tmr, wvcolmn, clwcolmn, tmpsfc, reflc, emiss, nlay, nmol, angle);
        if (iod == 1) write_odmono(dir_of(fileout), npr, nwn, wn, nlay, o);
    } catch (const Stop& s) {
        std::fclose(f);
        return fail(s.code, s.what());
    }
    std::fclose(f);
    return MRTM_OK;
}

// PROGRAM MONORTM for IATM=0 (src/monortm.f90:283-588)
extern "C" int mrtm_host_run_monortm(const char* workdir, int device, int64_t nwnmx, int verbose)
{
    std::string dir = workdir ? workdir : "";
    if (!dir.empty() && dir.back() != '/') dir += '/';
    const std::string filein = dir + "MONORTM.IN", fileprof = dir + "MONORTM_PROF.IN", hfile = dir + "TAPE3",
                      fileout = dir + "MONORTM.OUT", filelog = dir + "MONORTM.LOG";
    mrtm_ctx* ctx = nullptr;
    FILE *iot = nullptr, *ipr = nullptr;
    int rc = MRTM_OK;
    auto cleanup = [&]() {
        if (iot) std::fclose(iot);
        if (ipr) std::fclose(ipr);
        if (ctx) mrtm_free(ctx);
    };
    try {
        Control c;
        read_control(filein, dir, nwnmx > 0 ? nwnmx : kNWNMX, c);
        if (c.iatm != 0)
            throw Stop(MRTM_EARG, "IATM=1 needs LBLATM (src/lblatm.f90), which stays on the Fortran host (SURVEY 2 row 12); "
                                  "this driver handles layer input (IATM=0, MONORTM_PROF.IN)");
        if (c.ispd == 1) throw Stop(MRTM_EARG, " The ISPD=1 option is no longer valid.");
        const int64_t nprof = count_profiles(fileprof, c.ixsect);
        const int64_t nwn = (int64_t)c.wn.size();
        if (nwn <= 0) throw Stop(MRTM_EARG, "no wavenumbers requested");
        if (verbose) {
            for (const auto& w : c.warnings) std::printf(" %s\n", w.c_str());
            std::printf(" NUMBER OF PROFILES: %lld\n INPUTS FROM MONORTM_PROF.IN\n", (long long)nprof);
        }
        ipr = std::fopen(filelog.c_str(), "w");
        if (!ipr) throw Stop(MRTM_EIO, " EXIT; ERROR OPENING :" + filelog);
        std::fputs(c.log.c_str(), ipr);
        iot = std::fopen(fileout.c_str(), "w");
        if (!iot) throw Stop(MRTM_EIO, " EXIT; ERROR OPENING :" + fileout);

        // the GPU context and the line store (GET_LNFL with v1=wn(1), v2=wn(nwn): modm.f90:180-190)
        rc = mrtm_init(device, &ctx);
        if (rc) throw Stop(rc, std::string("mrtm_init: ") + (mrtm_last_error(ctx) ? mrtm_last_error(ctx) : mrtm_strerror(rc)));
        {
            int64_t iim = 250000;                             // IIM, lnfl_mod.f90:5
            if (const char* e = std::getenv("MRTM_IIM")) iim = std::max<int64_t>(1, std::atoll(e));
            const size_t n2 = (size_t)kMXMOL * (size_t)iim, n3 = (size_t)49 * (size_t)iim;
            std::vector<int64_t> nblm(kMXMOL), iso(n2);
            std::unique_ptr<double[]> a[10];
            for (auto& q : a) q.reset(new double[n2]());
            std::unique_ptr<int32_t[]> bflg(new int32_t[n3]());
            std::unique_ptr<double[]> btmp(new double[n3]()), bhw(new double[n3]()), bshft(new double[n3]());
            rc = mrtm_host_get_lnfl(hfile.c_str(), c.wn.front(), c.wn.back(), iim, nblm.data(), iso.data(), a[0].get(), a[1].get(),
                                    a[2].get(), a[3].get(), a[4].get(), a[5].get(), a[6].get(), a[7].get(), a[8].get(), a[9].get(),
                                    bflg.get(), btmp.get(), bhw.get(), bshft.get());
            if (rc) throw Stop(rc, "GET_LNFL failed on " + hfile + ": " + mrtm_strerror(rc));
            rc = mrtm_stage_lines(ctx, nblm.data(), iim, iso.data(), a[0].get(), a[1].get(), a[2].get(), a[3].get(), a[4].get(),
                                  a[5].get(), a[6].get(), a[7].get(), a[8].get(), a[9].get(), bflg.get(), btmp.get(), bhw.get(),
                                  bshft.get());
            if (rc) throw Stop(rc, std::string("mrtm_stage_lines: ") + (mrtm_last_error(ctx) ? mrtm_last_error(ctx) : ""));
        }

        ProfReader rd;
        rd.ixsect = c.ixsect;
        std::string xs_staged;                                // names the staged cross sections belong to
        if (!rd.u.open(fileprof)) throw Stop(MRTM_EIO, " EXIT; ERROR OPENING :" + fileprof);
        StoreState st;
        double tmpsfc = c.tmpbnd;
        std::vector<double> emiss((size_t)nwn), reflc((size_t)nwn);
        std::vector<double> rad((size_t)nwn), tb((size_t)nwn), tmr((size_t)nwn), trtot((size_t)nwn), rup((size_t)nwn), rdn((size_t)nwn);
        std::vector<double> otot((size_t)nwn), odxtot((size_t)nwn, 0.), obm((size_t)kMXMOL * (size_t)nwn);
        for (int64_t npr = 1; npr <= nprof; npr++) {
            Profile pr;
            if (!rd.next(pr)) break;                          // END=110
            if (c.nmol_scal > 0) profil_scal(c, pr);
            for (int64_t j = 0; j < nwn; j++) {               // EMISS_REFLEC
                reflc[(size_t)j] = coef_fn(c.wn[(size_t)j], c.bndrfl, c.refl, "REFLFN");
                emiss[(size_t)j] = coef_fn(c.wn[(size_t)j], c.bndemi, c.emis, "EMISFN");
            }
            double wvcolmn = 0., clwcolmn = 0.;               // INTEGR (:831-845)
            for (int64_t i = 0; i < pr.nlay; i++) {
                wvcolmn = wvcolmn + pr.wkl[(size_t)kMXMOL * (size_t)i];
                clwcolmn = clwcolmn + pr.clw[(size_t)i];
            }
            wvcolmn = wvcolmn * (2.99150e-23);
            // tips_2003 per layer (modm.f90:250) stays on the host side of the boundary
            std::vector<double> scor((size_t)MRTM_NSCOR1 * MRTM_NSCOR2 * (size_t)pr.nlay, 0.);
            for (int64_t k = 0; k < pr.nlay; k++) {
                rc = mrtm_host_tips_2003(pr.nmol, pr.t[(size_t)k], scor.data() + (size_t)MRTM_NSCOR1 * MRTM_NSCOR2 * (size_t)k);
                if (rc) throw Stop(rc, "TIPS_2003: temperature outside 70..3000 K or partition sum <= 0 (tips_2003.f90:271-277)");
            }
            std::vector<double> o((size_t)nwn * (size_t)pr.nlay);
            mrtm_opts xo;
            std::memset(&xo, 0, sizeof xo);
            if (c.ixsect >= 1) {                              // XSREAD (monortm.f90:494-497) + the tables, staged when the names change
                if (pr.xsnames != xs_staged) {
                    double xv1 = c.wn[0], xv2 = c.wn[0];
                    for (double w : c.wn) { xv1 = std::min(xv1, w); xv2 = std::max(xv2, w); }
                    int64_t nreg = 0;
                    mrtm_xs_region* regs = nullptr;
                    rc = mrtm_host_xsread(dir.c_str(), pr.ixmols, pr.xsnames.data(), xv1, xv2, &nreg, &regs);
                    if (rc) throw Stop(rc, g_host_err);
                    rc = mrtm_stage_xsec(ctx, nreg, regs);
                    mrtm_host_xs_free(regs, nreg);
                    if (rc) throw Stop(rc, std::string("mrtm_stage_xsec: ") + (mrtm_last_error(ctx) ? mrtm_last_error(ctx) : ""));
                    xs_staged = pr.xsnames;
                }
                xo.xamnt = pr.xamnt.data();
                xo.ld_xamnt = 38;
            }
            // MODM + CALCTMR + RTM (monortm.f90:557-574) on the GPU
            rc = mrtm_profiles(ctx, 1, nwn, c.wn.data(), c.dvset, pr.nlay, pr.p.data(), pr.t.data(), pr.tz.data(), pr.clw.data(),
                               pr.nmol, pr.wkl.data(), pr.wbrodl.data(), scor.data(), 1., 1., 0., c.cntnm, c.ibrd, pr.irt, c.iplot, 1,
                               &tmpsfc, emiss.data(), reflc.data(), rad.data(), tb.data(), tmr.data(), trtot.data(), rup.data(),
                               rdn.data(), o.data(), obm.data(), c.ixsect >= 1 ? &xo : nullptr);
            if (rc) throw Stop(rc, std::string("mrtm_profiles: ") + (mrtm_last_error(ctx) ? mrtm_last_error(ctx) : mrtm_strerror(rc)));
            if (c.ixsect >= 1) {                              // ODXTOT (:645, 650): the cross-section part alone, layer sum
                std::vector<double> odx((size_t)nwn * (size_t)pr.nlay);
                rc = mrtm_xsec(ctx, nwn, c.wn.data(), pr.nlay, pr.p.data(), pr.t.data(), 38, pr.xamnt.data(), odx.data());
                if (rc) throw Stop(rc, std::string("mrtm_xsec: ") + (mrtm_last_error(ctx) ? mrtm_last_error(ctx) : mrtm_strerror(rc)));
                for (int64_t iw = 0; iw < nwn; iw++) {
                    double sx = 0.;
                    for (int64_t j = 0; j < pr.nlay; j++) sx = sx + odx[(size_t)iw + (size_t)nwn * (size_t)j];
                    odxtot[(size_t)iw] = sx;
                }
            }
            for (int64_t iw = 0; iw < nwn; iw++) {            // OTOT (:643-646)
                double s = 0.;
                for (int64_t j = 0; j < pr.nlay; j++) s = s + o[(size_t)iw + (size_t)nwn * (size_t)j];
                otot[(size_t)iw] = s;
            }
            storeout_rows(iot, st, nwn, c.wn.data(), pr.wkl.data(), pr.wbrodl.data(), rad.data(), tb.data(), trtot.data(), npr,
                          otot.data(), obm.data(), odxtot.data(), tmr.data(), wvcolmn, clwcolmn, tmpsfc, reflc.data(), emiss.data(),
                          pr.nlay, pr.nmol, pr.angle);
            if (c.iod == 1) write_odmono(dir, npr, nwn, c.wn.data(), pr.nlay, o.data());
            if (verbose) std::printf("%30s%5lld\n", "PROCESSING PROFILE NUMBER:", (long long)npr);
        }
        std::fputs("\n--------------------------------------\n", ipr);
        std::fprintf(ipr, "Modules and versions used in this calculation:\n\n%s\n", mrtm_version());
    } catch (const Stop& s) {
        cleanup();
        return fail(s.code, s.what());
    } catch (const std::exception& e) {
        cleanup();
        return fail(MRTM_ENOMEM, e.what());
    }
    cleanup();
    return MRTM_OK;
}
