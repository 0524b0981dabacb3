// common.hpp -- host-side helpers shared by the swpc_3d and swpc_psv drivers: the input.inf reader and the pieces of
// src/shared (m_readini, m_fdtool, m_std, m_gk, m_geomap, m_seawater) the setup chains need.  Setup-only CPU code;
// kinds follow the Fortran declarations (default real = float, PI = real(DP), src/shared/m_std.f90:14).  Everything
// lives in an anonymous namespace: each driver translation unit gets its own copy (and its own error string).
#pragma once

#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_herr;
int hfail(const std::string &m) {
    g_herr = m;
    return 1;
}

constexpr double PI_D = 3.14159265358979323846;
constexpr double R_EARTH = 6371.0;
constexpr float EPS_SP = 1.1920929e-07f;   // epsilon(1.0)
constexpr int NBD = 9;

// ------------------------------------------------------------------------------------------------------------
// input.inf reader: src/shared/m_readini.f90:29-100 (+ typed wrappers :103-170), m_system.f90:75-102
class IniFile {
  public:
    bool strict = false;
    static IniFile from_text(const std::string &text) {
        IniFile f;
        std::istringstream is(text);
        std::string l;
        while (std::getline(is, l)) {
            if (!l.empty() && l.back() == '\r') l.pop_back();
            f.lines_.push_back(l);
        }
        return f;
    }
    static bool from_file(const std::string &path, IniFile &out) {
        std::ifstream is(path);
        if (!is) return false;
        std::stringstream ss;
        ss << is.rdbuf();
        out = from_text(ss.str());
        return true;
    }
    // first non-comment line that STARTS with the key and continues with '=' wins (:78-92)
    std::string get(const std::string &key, const std::string &def) const {
        for (const std::string &raw : lines_) {
            size_t p = raw.find_first_not_of(" \t");
            if (p == std::string::npos) continue;
            if (raw[p] == '#' || raw[p] == '!') continue;
            if (raw.compare(p, key.size(), key) != 0) continue;
            size_t q = raw.find_first_not_of(" \t", p + key.size());
            if (q == std::string::npos || raw[q] != '=') continue;
            return expand_env(list_directed(raw.substr(q + 1)));
        }
        if (strict) {
            std::fprintf(stderr, "[swpc3d_b200 readini] key %s is not found. Program terminate ...\n", key.c_str());
            std::exit(1);
        }
        return expand_env(def);
    }
    double get_d(const std::string &k, double def) const {
        char b[64];
        std::snprintf(b, sizeof b, "%.17g", def);
        return std::strtod(real_token(get(k, b)).c_str(), nullptr);
    }
    float get_s(const std::string &k, float def) const {
        char b[64];
        std::snprintf(b, sizeof b, "%.9g", (double)def);
        return std::strtof(real_token(get(k, b)).c_str(), nullptr);
    }
    int get_i(const std::string &k, int def) const { return (int)std::strtol(get(k, std::to_string(def)).c_str(), nullptr, 10); }
    bool get_l(const std::string &k, bool def) const {
        std::string v = get(k, def ? "T" : "F");
        size_t p = v.find_first_not_of(' ');
        if (p == std::string::npos) return false;
        if (v[p] == '.') p++;
        return p < v.size() && (v[p] == 'T' || v[p] == 't');
    }

  private:
    std::vector<std::string> lines_;
    // one item of a Fortran list-directed character read (:89)
    static std::string list_directed(const std::string &s) {
        size_t p = s.find_first_not_of(" \t");
        if (p == std::string::npos) return "";
        std::string out;
        if (s[p] == '\'' || s[p] == '"') {
            const char q = s[p++];
            while (p < s.size()) {
                if (s[p] == q) {
                    if (p + 1 < s.size() && s[p + 1] == q) { out += q; p += 2; continue; }
                    break;
                }
                out += s[p++];
            }
        } else {
            while (p < s.size() && s[p] != ' ' && s[p] != '\t' && s[p] != ',' && s[p] != '/') out += s[p++];
        }
        return out;
    }
    static std::string expand_env(const std::string &s) {
        std::string out;
        size_t p = 0;
        while (p < s.size()) {
            if (s[p] == '$' && p + 1 < s.size() && s[p + 1] == '{') {
                size_t e = s.find('}', p);
                if (e == std::string::npos) break;
                const char *v = std::getenv(s.substr(p + 2, e - p - 2).c_str());
                if (v) out += v;
                p = e + 1;
            } else out += s[p++];
        }
        if (p < s.size()) out += s.substr(p);
        while (!out.empty() && out.back() == ' ') out.pop_back();
        return out;
    }
    static std::string real_token(const std::string &s) {
        std::string t;
        for (char ch : s) {
            if (ch == ' ' || ch == '\t' || ch == ',' || ch == '/') { if (t.empty()) continue; break; }
            t += (ch == 'd' || ch == 'D' || ch == 'q' || ch == 'Q') ? 'e' : ch;
        }
        return t;
    }
};

// ------------------------------------------------------------------------------------------------------------
// src/shared/m_fdtool.f90 / m_std.f90 helpers
inline int x2i(float x, float xbeg, float dx) { return (int)std::ceil((x - xbeg) / dx); }               // :600-609
inline float i2x(int i, float xbeg, float dx) { float h = (float)i - 0.5f; return xbeg + h * dx; }       // :639-648
inline float deg2rad_s(float deg) { return (float)(PI_D / (double)180.0f * (double)deg); }               // m_std.f90:132-139
inline float rad2deg_s(float rad) { return (float)((double)180.0f / PI_D * (double)rad); }               // m_std.f90:152-159

void decomp1d(int n, int nproc, int proc, int &np, int &beg, int &end) {   // m_global.f90:234-247, :275-288
    const int m = n % nproc, q = (n - m) / nproc;
    if (proc <= nproc - m - 1) { np = q; beg = proc * (n - m) / nproc + 1; end = (proc + 1) * (n - m) / nproc; }
    else { np = q + 1; beg = proc * (q + 1) - (nproc - m) + 1; end = (proc + 1) * (q + 1) - (nproc - m); }
}

void relax_times(int nm, float *ts, float fmin, float fmax) {   // visco_set_relaxtime :691-727
    const float wa = (float)(2 * PI_D * (double)fmin), wb = (float)(2 * PI_D * (double)fmax);
    if (nm == 0) return;
    if (nm == 1) { ts[0] = 1.0f / std::sqrt(wa * wb); return; }
    for (int im = 1; im <= nm; im++) {
        const double e = (double)(im - 1) / (double)(nm - 1);
        const float w = (float)((double)wa * std::pow((double)(wb / wa), e));
        ts[im - 1] = 1.0f / w;
    }
}

float constq_zeta(int nm, float fmin, float fmax, const float *ts) {   // visco_constq_zeta :756-812
    if (nm == 0) return 0.0f;
    const float wa = (float)(2 * PI_D * (double)fmin), wb = (float)(2 * PI_D * (double)fmax);
    float i0s = 0.0f, i1s = 0.0f, i2s = 0.0f;
    std::vector<float> i0(nm), i1(nm);
    for (int m = 0; m < nm; m++) {
        const float t = ts[m];
        i0[m] = (std::log(1.0f + (wb * wb) * (t * t)) - std::log(1.0f + (wa * wa) * (t * t))) / (2 * t);
        i1[m] = ((std::atan(wb * t) - wb * t / (1 + (wb * wb) * (t * t))) - (std::atan(wa * t) - wa * t / (1 + (wa * wa) * (t * t)))) / (2 * t);
    }
    for (int m = 0; m < nm; m++) i0s += i0[m];
    for (int m = 0; m < nm; m++) i1s += i1[m];
    for (int m = 0; m < nm - 1; m++)
        for (int q = m + 1; q < nm; q++) {
            const float w1 = std::atan(wb * ts[m]) / ts[m] - std::atan(wb * ts[q]) / ts[q];
            const float w2 = std::atan(wa * ts[m]) / ts[m] - std::atan(wa * ts[q]) / ts[q];
            const float v = ts[m] * ts[q] / (ts[q] * ts[q] - ts[m] * ts[m]) * (w1 - w2);
            i2s = i2s + v;
        }
    return i0s / (i1s + 2 * i2s);
}

float stable_dt(float dx, float dy, float dz, float vmax) {   // fdm_stable_dt :81-96
    const float hh = 1.0f / std::sqrt(1 / (dx * dx) + 1 / (dy * dy) + 1 / (dz * dz));
    const float cc = 6.0f / 7.0f;
    return cc * hh / vmax;
}
float moment_magnitude(float m0) { return m0 < EPS_SP ? -12345.0f : (std::log10(m0) - 9.1f) * 2.0f / 3.0f; }   // :281-293
// source time functions, m_fdtool.f90:339-497 (PI is DP: the trigonometric ones are evaluated in double and rounded)
float momentrate(float t, const std::string &stf, float ts, float tr) {
    if (stf == "boxcar") return (ts <= t && t <= ts + tr) ? 1.0f / tr : 0.0f;
    if (stf == "triangle") {
        if (ts <= t && t <= ts + tr / 2) return 4 * (t - ts) / (tr * tr);
        if (ts + tr / 2 < t && t <= ts + tr) return -4 * (t - ts - tr) / (tr * tr);
        return 0.0f;
    }
    if (stf == "herrmann") {
        const float t1 = ts + tr / 4, t2 = ts + 3 * tr / 4, tr3 = tr * tr * tr;
        if (ts <= t && t < t1) return 16 * ((t - ts) * (t - ts)) / tr3;
        if (t1 <= t && t < t2) return -2 * (8 * (t * t + tr * ts + ts * ts - t * tr - 2 * t * ts) + tr * tr) / tr3;
        if (t2 <= t && t <= ts + tr) return 16 * ((ts + tr - t) * (ts + tr - t)) / tr3;
        return 0.0f;
    }
    if (stf == "cosine") return (ts <= t && t <= ts + tr) ? (float)((1 - std::cos(2 * PI_D * (double)(t - ts) / (double)tr)) / (double)tr) : 0.0f;
    if (stf == "texp") {
        if (!(ts <= t)) return 0.0f;
        const float tt = t - ts;
        return (float)((2 * PI_D) * (2 * PI_D) * (double)tt / (double)(tr * tr) * std::exp(-2 * PI_D * (double)tt / (double)tr));
    }
    if (ts <= t && t <= ts + tr) {   // kupper, also the default branch (:494)
        const double sn = std::sin(PI_D * (double)(t - ts) / (double)tr);
        return (float)(3 * PI_D * (sn * sn * sn) / (double)(4 * tr));
    }
    return 0.0f;
}
float powi_sp(float x, int m) {   // real(SP) ** integer the way gfortran does it (libgcc __powisf2)
    unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
    float y = (n % 2) ? x : 1.0f;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return m < 0 ? 1.0f / y : y;
}
float seismic_moment(float mw) { return std::pow(10.0f, 1.5f * mw + 9.05f); }                                     // :296-304

void sdr2moment(float strike, float dip, float rake, float m[6]) {   // :307-336 ; m = mxx myy mzz myz mxz mxy
    const float sd = std::sin(deg2rad_s(dip)), cd = std::cos(deg2rad_s(dip));
    const float s2d = std::sin(deg2rad_s(2 * dip)), c2d = std::cos(deg2rad_s(2 * dip));
    const float sl = std::sin(deg2rad_s(rake)), cl = std::cos(deg2rad_s(rake));
    const float sf = std::sin(deg2rad_s(strike)), cf = std::cos(deg2rad_s(strike));
    const float s2f = std::sin(deg2rad_s(2 * strike)), c2f = std::cos(deg2rad_s(2 * strike));
    m[0] = -(sd * cl * s2f + s2d * sl * sf * sf);
    m[5] = (sd * cl * c2f + s2d * sl * s2f / 2);
    m[4] = -(cd * cl * cf + c2d * sl * sf);
    m[1] = (sd * cl * s2f - s2d * sl * cf * cf);
    m[3] = -(cd * cl * sf - c2d * sl * cf);
    m[2] = (s2d * sl);
}

float seawater_vel(float z, bool munk) {   // m_seawater.f90:34-47
    const double eps = munk ? 0.00737 : 0.0, zc = 1300.0;
    const double zb = 2 * ((double)z * (double)1000.0f - zc) / zc;
    return (float)(1.5 * (1.0 + eps * (zb - 1.0 + std::exp(-zb))));
}

// ADE-CFS PML profile, m_absorb_p.f90:533-573
void damping_profile(float x, float H, float xb, float xe, int na, float fcut, float dt, float g[4]) {
    const float cp = 6.0f, b0 = 7.0f;
    const float R0 = std::pow(10.0f, -(std::log10((float)na) - 1) / std::log10(2.0f) - 3.0f);
    const float d0 = -((1.0f / (2.0f * H)) * 2.0f * cp * std::log(R0));
    const float a0 = (float)(PI_D * (double)fcut);
    float xx = 0.0f;
    if (x <= xb + H) xx = (xb + H) - x;
    else if (x >= xe - H) xx = x - (xe - H);
    const float q = std::fabs(xx / H);
    const float d = d0 * q, a = a0 * (1.0f - q), b = 1.0f + (b0 - 1.0f) * (q * q);
    const float den = 1.0f + (dt / 2.0f) * (a + d / b);
    g[0] = ((1.0f + (dt / 2.0f) * a) / b) / den;
    g[1] = (-1.0f / b) / den;
    g[2] = (1.0f - (dt / 2.0f) * (a + d / b)) / den;
    g[3] = (d / b) / den;
}

// Gauss-Krueger projection, src/shared/m_gk.f90 (all double)
struct GaussKrueger {
    double al[6], be[6], AA[6], de[7];
    static constexpr double a = 6378137.0, F = 298.257222101, m0 = 0.9999;
    double n;
    GaussKrueger() {
        n = 1.0 / (2.0 * F - 1.0);
        al[1] = (1 / 2. + (-2 / 3. + (5 / 16. + (41 / 180. - 127 / 288. * n) * n) * n) * n) * n;
        al[2] = (13 / 48. + (-3 / 5. + (557 / 1440. + 281 / 630. * n) * n) * n) * (n * n);
        al[3] = (61 / 240. + (-103 / 140. + 15061 / 26880. * n) * n) * (n * n * n);
        al[4] = (49561 / 161280. - 179 / 168. * n) * (n * n * n * n);
        al[5] = 34729 / 80640. * (n * n * n * n * n);
        be[1] = (1 / 2. + (-2 / 3. + (37 / 96. + (-1 / 360. - 81 / 512. * n) * n) * n) * n) * n;
        be[2] = ((1 / 48. + (1 / 15. + (-437 / 1440. + 46 / 105. * n) * n) * n) * n) * n;
        be[3] = (((17 / 480. + (-37 / 840. - 209 / 4480. * n) * n) * n) * n) * n;
        be[4] = ((((4397 / 161280. - 11 / 504. * n) * n) * n) * n) * n;
        be[5] = ((((4583 / 161280. * n) * n) * n) * n) * n;
        de[1] = (2 / 1. + (-2 / 3. + (-2 / 1. + (116 / 45. + (26 / 45. + (-2854 / 675.) * n) * n) * n) * n) * n) * n;
        de[2] = ((7 / 3. + (-8 / 5. + (-227 / 45. + (2704 / 315. + (2323 / 945.) * n) * n) * n) * n) * n) * n;
        de[3] = (((56 / 15. + (-136 / 35. + (-1262 / 105. + (73814 / 2835.) * n) * n) * n) * n) * n) * n;
        de[4] = ((((4279 / 630. + (-332 / 35. + (-399572 / 14175.) * n) * n) * n) * n) * n) * n;
        de[5] = (((((4174 / 315. + (-144838 / 6237.) * n) * n) * n) * n) * n) * n;
        de[6] = ((((((601676 / 22275.) * n) * n) * n) * n) * n) * n;
        const double n2 = n * n, n4 = n2 * n2;
        AA[0] = 1 + (1 / 4. + 1 / 64. * n2) * n2;
        AA[1] = -3 / 2. * (1. - 1 / 8. * n2 - 1 / 64. * n4) * n;
        AA[2] = 15 / 16. * (1. - 1 / 4. * n2) * n2;
        AA[3] = -35 / 48. * (1. - 5 / 16. * n2) * (n2 * n);
        AA[4] = 315 / 512. * n4;
        AA[5] = -693 / 1280. * (n4 * n);
    }
    static double atanh0(double x) { return std::log((1. + x) / (1. - x)) / 2.0; }
    double S_phi0(double p0) const {
        double s = AA[0] * p0;
        for (int j = 1; j <= 5; j++) s = s + AA[j] * std::sin(2 * j * p0);
        return s * (m0 * a / (1 + n));
    }
    void ll2xy(double lon, double lat, double lon0, double lat0, double &x, double &y) const {   // :40-92
        const double d2r = PI_D / 180.0;
        const double lam = d2r * lon, lam0 = d2r * lon0, phi = d2r * lat, phi0 = d2r * lat0;
        const double e2n = 2.0 * std::sqrt(n) / (1.0 + n);
        const double lc = std::cos(lam - lam0), ls = std::sin(lam - lam0);
        const double tchi = std::sinh(atanh0(std::sin(phi)) - e2n * std::atanh(e2n * std::sin(phi)));
        const double cchi = std::sqrt(1 + tchi * tchi);
        const double xi = std::atan(tchi / lc), eta = atanh0(ls / cchi);
        const double Abar = m0 * a / (1 + n) * AA[0];
        double xx = xi, yy = eta;
        for (int j = 1; j <= 5; j++) {
            xx = xx + al[j] * std::sin(2 * j * xi) * std::cosh(2 * j * eta);
            yy = yy + al[j] * std::cos(2 * j * xi) * std::sinh(2 * j * eta);
        }
        x = (Abar * xx - S_phi0(phi0)) / 1000;
        y = (Abar * yy) / 1000;
    }
    void xy2ll(double x, double y, double lon0, double lat0, double &lon, double &lat) const {   // :115-158
        const double d2r = PI_D / 180.0, r2d = 180.0 / PI_D;
        const double lam0 = d2r * lon0, phi0 = d2r * lat0;
        const double Abar = m0 * a / (1 + n) * AA[0];
        const double xi = (x * 1000 + S_phi0(phi0)) / Abar, eta = y * 1000 / Abar;
        double xi2 = xi, eta2 = eta;
        for (int j = 1; j <= 5; j++) {
            xi2 = xi2 - be[j] * std::sin(2 * j * xi) * std::cosh(2 * j * eta);
            eta2 = eta2 - be[j] * std::cos(2 * j * xi) * std::sinh(2 * j * eta);
        }
        const double chi = std::asin(std::sin(xi2) / std::cosh(eta2));
        const double lam = lam0 + std::atan(std::sinh(eta2) / std::cos(xi2));
        double phi = chi;
        for (int j = 1; j <= 6; j++) phi = phi + de[j] * std::sin(2 * j * chi);
        lon = r2d * lam;
        lat = r2d * phi;
    }
};
const GaussKrueger &gk() { static GaussKrueger g; return g; }

void geomap_g2c(float lon, float lat, float lon0, float lat0, float phi, float &x, float &y) {   // m_geomap.f90:18-41
    const float pr = deg2rad_s(phi);
    double xd, yd;
    gk().ll2xy(lon, lat, lon0, lat0, xd, yd);
    const float xx = (float)xd, yy = (float)yd;
    x = std::cos(pr) * xx + std::sin(pr) * yy;
    y = -std::sin(pr) * xx + std::cos(pr) * yy;
}
void geomap_c2g(float x, float y, float lon0, float lat0, float phi, float &lon, float &lat) {   // m_geomap.f90:44-65
    const float pr = deg2rad_s(phi);
    const float xx = std::cos(pr) * x - std::sin(pr) * y, yy = std::sin(pr) * x + std::cos(pr) * y;
    double lo, la;
    gk().xy2ll(xx, yy, lon0, lat0, lo, la);
    lon = (float)lo;
    lat = (float)la;
}

std::vector<float> parse_reals(const std::string &line) {
    std::vector<float> v;
    size_t p = 0;
    while (p < line.size()) {
        while (p < line.size() && (line[p] == ' ' || line[p] == '\t' || line[p] == ',' || line[p] == '\r' || line[p] == '\n')) p++;
        if (p >= line.size()) break;
        std::string t;
        while (p < line.size() && line[p] != ' ' && line[p] != '\t' && line[p] != ',' && line[p] != '\r' && line[p] != '\n') {
            char ch = line[p++];
            t += (ch == 'd' || ch == 'D') ? 'e' : ch;
        }
        char *e = nullptr;
        const float x = std::strtof(t.c_str(), &e);
        if (e == t.c_str()) break;
        v.push_back(x);
    }
    return v;
}
bool blank_or_comment(const std::string &l) {
    size_t p = l.find_first_not_of(" \t\r\n");
    return p == std::string::npos || l[p] == '#';
}
std::string join_path(const std::string &base, const std::string &fn) {
    if (fn.empty() || fn[0] == '/' || base.empty()) return fn;
    return base + "/" + fn;
}

// ------------------------------------------------------------------------------------------------------------
// waveform files: wav__write of m_wav.f90 (swpc_3d :677-763, swpc_psv :260-420) for wav_format = sac | csf | tar_st | tar_node.
// A record is a complete SAC file image (632-byte header + npts samples).  `code` is "3d" or "psv"; tar member names are
// title.stnm.3d.cmp.sac in swpc_3d (m_wav.f90:772) but title.psv.stnm.cmp.sac in swpc_psv (m_wav.f90:428).
struct WavTrace { int prod = 0; std::string cmp; std::vector<unsigned char> rec; };
struct WavStation { std::string stnm; std::vector<WavTrace> tr; };

void make_dirs(const std::string &p) {
    for (size_t q = 1; q <= p.size(); q++)
        if (q == p.size() || p[q] == '/') mkdir(p.substr(0, q).c_str(), 0777);
}

int write_wav_files(const std::string &fmt, const std::string &dir, const std::string &title, const std::string &code, bool stnm_first,
                    int myid, int exedate, int ntw, const std::vector<WavStation> &sts, std::string &err) {
    int count = 0;
    if (fmt == "sac") {
        for (const WavStation &st : sts)
            for (const WavTrace &t : st.tr) {
                const std::string fn = dir + "/" + title + "." + code + "." + st.stnm + "." + t.cmp + ".sac";
                FILE *fp = std::fopen(fn.c_str(), "wb");
                if (!fp) { err = "cannot write " + fn; return -1; }
                std::fwrite(t.rec.data(), 1, t.rec.size(), fp);
                std::fclose(fp);
                count++;
            }
    } else if (fmt == "csf") {
        // export_wav__csf + wcsf_s (m_sac.f90:585-648): 'CSFD', ntrace, npts, then (header, data) per trace, one file per rank.
        // Every enabled product goes to the SAME file name in the reference (the second one stops at an interactive
        // overwrite prompt); here the later product replaces the earlier one.
        char cid[16];
        std::snprintf(cid, sizeof(cid), "%05d", myid);
        const std::string fn = dir + "/" + title + "__" + cid + "__.csf";
        for (int prod = 0; prod < 4; prod++) {
            int32_t ntrace = 0;
            for (const WavStation &st : sts) for (const WavTrace &t : st.tr) if (t.prod == prod) ntrace++;
            if (ntrace == 0) continue;
            FILE *fp = std::fopen(fn.c_str(), "wb");
            if (!fp) { err = "cannot write " + fn; return -1; }
            const int32_t npts = ntw;
            std::fwrite("CSFD", 1, 4, fp); std::fwrite(&ntrace, 4, 1, fp); std::fwrite(&npts, 4, 1, fp);
            for (const WavStation &st : sts)
                for (const WavTrace &t : st.tr) if (t.prod == prod) { std::fwrite(t.rec.data(), 1, t.rec.size(), fp); count++; }
            std::fclose(fp);
        }
    } else if (fmt == "tar_st" || fmt == "tar_node") {
        // sac__wtar m_sac.f90:812-829 + tar__whdr m_tar.f90:165-199 (ustar header, octal fields filling their width, no
        // version digits, checksum over the block with the checksum field blank) + tar__wpad (:212-224, a full null block when
        // the size is a multiple of 512) + tar__wend
        static const char zeros[1024] = {0};
        auto tar_member = [&](FILE *fp, const std::string &name, const std::vector<unsigned char> &body) {
            char hd[513];
            std::memset(hd, 0, sizeof(hd));
            std::memcpy(hd, name.data(), std::min<size_t>(name.size(), 100));
            char num[16];
            std::snprintf(num, sizeof(num), "%08o", 420u); std::memcpy(hd + 100, num, 8);
            std::snprintf(num, sizeof(num), "%08o", 0u); std::memcpy(hd + 108, num, 8); std::memcpy(hd + 116, num, 8);
            std::snprintf(num, sizeof(num), "%012o", (unsigned)body.size()); std::memcpy(hd + 124, num, 12);
            std::snprintf(num, sizeof(num), "%012o", (unsigned)exedate); std::memcpy(hd + 136, num, 12);
            std::memset(hd + 148, ' ', 8);
            hd[156] = '0';
            std::memcpy(hd + 257, "ustar", 5);
            std::memcpy(hd + 265, "root", 4);
            std::memcpy(hd + 297, "root", 4);
            unsigned sum = 0;
            for (int q = 0; q < 512; q++) sum += (unsigned char)hd[q];
            std::snprintf(num, sizeof(num), "%08o", sum); std::memcpy(hd + 148, num, 8);
            std::fwrite(hd, 1, 512, fp);
            std::fwrite(body.data(), 1, body.size(), fp);
            std::fwrite(zeros, 1, 512 - body.size() % 512, fp);
        };
        FILE *fp = nullptr;
        if (fmt == "tar_node") {
            char cid[16];
            std::snprintf(cid, sizeof(cid), "%06d", myid);
            const std::string fn = dir + "/" + title + "." + code + "." + cid + ".sac.tar";
            if (!(fp = std::fopen(fn.c_str(), "wb"))) { err = "cannot write " + fn; return -1; }
        }
        for (const WavStation &st : sts) {
            if (fmt == "tar_st") {
                const std::string fn = dir + "/" + title + "." + code + "." + st.stnm + ".sac.tar";
                if (!(fp = std::fopen(fn.c_str(), "wb"))) { err = "cannot write " + fn; return -1; }
            }
            for (const WavTrace &t : st.tr) {
                tar_member(fp, stnm_first ? title + "." + st.stnm + "." + code + "." + t.cmp + ".sac" : title + "." + code + "." + st.stnm + "." + t.cmp + ".sac", t.rec);
                count++;
            }
            if (fmt == "tar_st") { std::fwrite(zeros, 1, 1024, fp); std::fclose(fp); fp = nullptr; }
        }
        if (fp) { std::fwrite(zeros, 1, 1024, fp); std::fclose(fp); }
    }
    // any other wav_format: wav__write has no branch for it and writes nothing
    return count;
}

// ------------------------------------------------------------------------------------------------------------
// netCDF classic (CDF-1) writer for the snapshot files of m_snap.f90 (the image has no netCDF library): dimensions,
// attributes and variables in definition order, fixed-size variables first, then the record section; the header is
// rewritten in place when an attribute value (actual_range) or the record count changes.
struct NcAtt { std::string name; int type; std::vector<unsigned char> val; int nelems; };
struct NcVar {
    std::string name; std::vector<int> dimids; std::vector<NcAtt> atts; int type = 5; bool rec = false;
    std::vector<float> data; long long vsize = 0, begin = 0;
};
void be32(std::vector<unsigned char> &b, uint32_t v) { for (int q = 3; q >= 0; q--) b.push_back((unsigned char)(v >> (8 * q))); }
void bef(std::vector<unsigned char> &b, float f) { uint32_t u; std::memcpy(&u, &f, 4); be32(b, u); }
NcAtt att_text(const std::string &n, const std::string &v) { NcAtt a{n, 2, {}, (int)v.size()}; a.val.assign(v.begin(), v.end()); while (a.val.size() % 4) a.val.push_back(0); return a; }
NcAtt att_int(const std::string &n, int v) { NcAtt a{n, 4, {}, 1}; be32(a.val, (uint32_t)v); return a; }
NcAtt att_floats(const std::string &n, std::initializer_list<float> v) { NcAtt a{n, 5, {}, (int)v.size()}; for (float f : v) bef(a.val, f); return a; }

class NcFile {
  public:
    std::vector<std::pair<std::string, int>> dims;   // length 0 = record dimension
    std::vector<NcAtt> gatts;
    std::vector<NcVar> vars;
    int numrecs = 0;
    FILE *fp = nullptr;
    long long recsize = 0, rec_begin = 0;
    ~NcFile() { if (fp) std::fclose(fp); }
    int var_index(const std::string &n) const { for (size_t q = 0; q < vars.size(); q++) if (vars[q].name == n) return (int)q; return -1; }
    NcAtt *find_att(NcVar &v, const std::string &n) { for (auto &a : v.atts) if (a.name == n) return &a; return nullptr; }
    static void put_name(std::vector<unsigned char> &b, const std::string &n) { be32(b, (uint32_t)n.size()); for (char c : n) b.push_back((unsigned char)c); while (b.size() % 4) b.push_back(0); }
    static void put_atts(std::vector<unsigned char> &b, const std::vector<NcAtt> &atts) {
        if (atts.empty()) { be32(b, 0); be32(b, 0); return; }
        be32(b, 0x0C); be32(b, (uint32_t)atts.size());
        for (const NcAtt &a : atts) { put_name(b, a.name); be32(b, (uint32_t)a.type); be32(b, (uint32_t)a.nelems); b.insert(b.end(), a.val.begin(), a.val.end()); }
    }
    std::vector<unsigned char> header() const {
        std::vector<unsigned char> b = {'C', 'D', 'F', 1};
        be32(b, (uint32_t)numrecs);
        be32(b, 0x0A); be32(b, (uint32_t)dims.size());
        for (auto &d : dims) { put_name(b, d.first); be32(b, (uint32_t)d.second); }
        put_atts(b, gatts);
        be32(b, 0x0B); be32(b, (uint32_t)vars.size());
        for (const NcVar &v : vars) {
            put_name(b, v.name); be32(b, (uint32_t)v.dimids.size());
            for (int d : v.dimids) be32(b, (uint32_t)d);
            put_atts(b, v.atts); be32(b, (uint32_t)v.type); be32(b, (uint32_t)v.vsize); be32(b, (uint32_t)v.begin);
        }
        return b;
    }
    // lay out: fixed-size variables first (definition order), then the record section
    bool create(const std::string &path) {
        for (NcVar &v : vars) {
            long long n = 1;
            for (int d : v.dimids) if (dims[(size_t)d].second > 0) n *= dims[(size_t)d].second;
            v.vsize = (n * 4 + 3) / 4 * 4;
        }
        const long long hsize = (long long)header().size();
        long long off = hsize;
        for (NcVar &v : vars) if (!v.rec) { v.begin = off; off += v.vsize; }
        rec_begin = off; recsize = 0;
        for (NcVar &v : vars) if (v.rec) { v.begin = off; off += v.vsize; recsize += v.vsize; }
        fp = std::fopen(path.c_str(), "wb+");
        if (!fp) return false;
        flush_header();
        for (NcVar &v : vars) if (!v.rec && !v.data.empty()) put_fixed(v);
        return true;
    }
    void flush_header() { const auto b = header(); std::fseek(fp, 0, SEEK_SET); std::fwrite(b.data(), 1, b.size(), fp); std::fflush(fp); }
    static void write_floats(FILE *f, const float *p, size_t n) { std::vector<unsigned char> b; b.reserve(4 * n); for (size_t q = 0; q < n; q++) bef(b, p[q]); std::fwrite(b.data(), 1, b.size(), f); }
    void put_fixed(const NcVar &v) { std::fseek(fp, (long)v.begin, SEEK_SET); write_floats(fp, v.data.data(), v.data.size()); }
    void put_record(int vidx, int rec, const float *p, size_t n) {
        const NcVar &v = vars[(size_t)vidx];
        std::fseek(fp, (long)(v.begin + (long long)rec * recsize), SEEK_SET);
        write_floats(fp, p, n);
        if (rec + 1 > numrecs) numrecs = rec + 1;
    }
};


}   // namespace
