// fortio.h -- the slice of Fortran formatted I/O the monoRTM control/profile/result files use,
// as gfortran executes it (the parity build linuxGNUdbl, build/makefile.common:195-198).
// Input : fixed-column fields of a record with PAD='YES', BLANK='NULL' (blank field = 0, embedded
//         blanks ignored), Fw.d/Ew.d with the implied decimal point when the field has none,
//         exponents written E, D or bare sign.
// Output: Iw, Fw.d, 1P Ew.d, Aw with gfortran's width-overflow asterisks, optional leading zero and
//         3-digit exponent form.
// Used by mrtm_driver.cpp (RDLBLINP / MONORTM_PROF.IN / STOREOUT restated for the harness side of
// SURVEY 8f-1).  Host only.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace fortio {

struct IoError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// One formatted input record, consumed left to right by edit descriptors.
class Record {
public:
    explicit Record(std::string s) : s_(std::move(s))
    {
        while (!s_.empty() && (s_.back() == '\n' || s_.back() == '\r')) s_.pop_back();
    }
    void X(int n) { pos_ += (size_t)n; }
    std::string A(int w)
    {
        std::string f = field(w);
        return f;
    }
    long long I(int w)
    {
        std::string f = squeeze(field(w));
        if (f.empty()) return 0;
        size_t k = 0;
        bool neg = false;
        if (f[k] == '+' || f[k] == '-') neg = f[k++] == '-';
        if (k >= f.size()) throw IoError("bad integer field '" + f + "'");
        long long v = 0;
        for (; k < f.size(); k++) {
            if (f[k] < '0' || f[k] > '9') throw IoError("bad integer field '" + f + "'");
            v = v * 10 + (f[k] - '0');
        }
        return neg ? -v : v;
    }
    // Fw.d, Ew.d, Dw.d, Gw.d on input are the same edit
    double F(int w, int d)
    {
        std::string f = squeeze(field(w));
        if (f.empty()) return 0.0;
        // split mantissa / exponent
        size_t k = 0;
        std::string mant, expo;
        if (f[k] == '+' || f[k] == '-') mant += f[k++];
        bool dot = false, digits = false;
        for (; k < f.size(); k++) {
            char c = f[k];
            if (c >= '0' && c <= '9') { mant += c; digits = true; }
            else if (c == '.' && !dot) { mant += c; dot = true; }
            else break;
        }
        if (!digits) throw IoError("bad real field '" + f + "'");
        bool has_exp = false;
        if (k < f.size()) {
            char c = f[k];
            if (c == 'E' || c == 'e' || c == 'D' || c == 'd' || c == 'Q' || c == 'q') { k++; has_exp = true; }
            else if (c == '+' || c == '-') has_exp = true;
            else throw IoError("bad real field '" + f + "'");
            if (k < f.size() && (f[k] == '+' || f[k] == '-')) expo += f[k++];
            bool ed = false;
            for (; k < f.size(); k++) {
                if (f[k] < '0' || f[k] > '9') throw IoError("bad real field '" + f + "'");
                expo += f[k];
                ed = true;
            }
            if (!ed) {
                if (has_exp && expo.empty()) expo = "0";     // "1.0E" -> exponent 0 (gfortran accepts)
                else throw IoError("bad real field '" + f + "'");
            }
        }
        int e10 = expo.empty() ? 0 : std::atoi(expo.c_str());
        if (!dot) e10 -= d;                                  // implied decimal point
        std::string txt = mant + "e" + std::to_string(e10);
        return std::strtod(txt.c_str(), nullptr);            // correctly rounded, like libgfortran
    }
    bool blank_rest() const
    {
        for (size_t k = pos_; k < s_.size(); k++)
            if (s_[k] != ' ' && s_[k] != '\t') return false;
        return true;
    }
    const std::string& text() const { return s_; }

private:
    std::string field(int w)
    {
        std::string f;
        if (pos_ < s_.size()) f = s_.substr(pos_, (size_t)w);
        f.resize((size_t)w, ' ');                            // PAD='YES'
        pos_ += (size_t)w;
        return f;
    }
    static std::string squeeze(const std::string& f)
    {
        std::string o;
        for (char c : f)
            if (c != ' ' && c != '\t') o += c;               // BLANK='NULL'
        return o;
    }
    std::string s_;
    size_t pos_ = 0;
};

// ---- output edits ----------------------------------------------------------------------------
inline std::string stars(int w) { return std::string((size_t)w, '*'); }

inline std::string rjust(const std::string& s, int w)
{
    if ((int)s.size() > w) return stars(w);
    return std::string((size_t)(w - (int)s.size()), ' ') + s;
}

inline std::string fmt_I(int w, long long v) { return rjust(std::to_string(v), w); }

// Aw: right-justified when the item is shorter than w, leftmost w characters when longer
inline std::string fmt_A(int w, const std::string& s)
{
    if ((int)s.size() >= w) return s.substr(0, (size_t)w);
    return std::string((size_t)(w - (int)s.size()), ' ') + s;
}

inline bool special(double v, int w, std::string& out)
{
    if (std::isnan(v)) { out = rjust("NaN", w); return true; }
    if (std::isinf(v)) {
        std::string s = v < 0 ? "-Infinity" : "Infinity";
        if ((int)s.size() > w) s = v < 0 ? "-Inf" : "Inf";
        out = rjust(s, w);
        return true;
    }
    return false;
}

// Fw.d
inline std::string fmt_F(int w, int d, double v)
{
    std::string out;
    if (special(v, w, out)) return out;
    char buf[512];
    std::snprintf(buf, sizeof buf, "%.*f", d, v);
    std::string s = buf;
    if ((int)s.size() > w) {                                 // the leading zero is optional
        if (s.compare(0, 2, "0.") == 0) s.erase(0, 1);
        else if (s.compare(0, 3, "-0.") == 0) s.erase(1, 1);
    }
    return rjust(s, w);
}

// 1P Ew.d : one digit before the point, d after; 3-digit exponents drop the letter
inline std::string fmt_1PE(int w, int d, double v)
{
    std::string out;
    if (special(v, w, out)) return out;
    char buf[512];
    std::snprintf(buf, sizeof buf, "%.*E", d, v);
    std::string s = buf;
    size_t e = s.find('E');
    if (e != std::string::npos && s.size() - e - 2 >= 3) s.erase(e, 1);     // 1.2345+100
    return rjust(s, w);
}

}  // namespace fortio
