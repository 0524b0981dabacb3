/*
 * oracle/ora_tools.c -- restatement of the small numerical helpers the hot path depends on
 * (TEST INFRASTRUCTURE, see ora.h).  Every function cites the reference lines it follows; the
 * declared kind of every temporary is kept (Fortran default real == float, PI is real(DP),
 * src/shared/m_std.f90:14).
 */
#include "ora.h"

#include <complex.h>
#include <math.h>
#include <string.h>

static const double ORA_PI = 3.14159265358979323846; /* atan(1.0_DP)*4, m_std.f90:14 */
static const double ORA_R_EARTH = 6371.0;           /* m_std.f90:15 */

/* ------------------------------------------------------------------ index helpers */
/* m_fdtool.f90:600-609: x2i = ceiling((x - xbeg) / dx), all real(SP) */
int ora_x2i(float x, float xbeg, float dx) {
    float q = (x - xbeg) / dx;
    return (int)ceilf(q);
}

/* m_fdtool.f90:639-648: i2x = xbeg + (i - 0.5) * dx */
float ora_i2x(int i, float xbeg, float dx) {
    float h = (float)i - 0.5f;
    return xbeg + h * dx;
}

/* m_global.f90:234-247 (block size) and :275-288 (range): remainder cells go to the LAST ranks */
void ora_decomp1d(int n, int nproc, int proc, int *np, int *beg, int *end) {
    int m = n % nproc;
    if (proc <= nproc - m - 1) {
        *np = (n - m) / nproc;
        *beg = proc * (n - m) / nproc + 1;
        *end = (proc + 1) * (n - m) / nproc;
    } else {
        *np = (n - m) / nproc + 1;
        *beg = proc * ((n - m) / nproc + 1) - (nproc - m) + 1;
        *end = (proc + 1) * ((n - m) / nproc + 1) - (nproc - m);
    }
}

/* ------------------------------------------------------------------ source time functions */
/* m_fdtool.f90:339-353; PI is DP so the expression is evaluated in double and rounded on return */
static float stf_kupper(float t, float ts, float tr) {
    if (ts <= t && t <= ts + tr) {
        double s = sin(ORA_PI * (double)(t - ts) / (double)tr);
        return (float)(3 * ORA_PI * (s * s * s) / (double)(4 * tr));
    }
    return 0.0f;
}
/* m_fdtool.f90:356-376 */
static float stf_texp(float t, float ts, float tr) {
    if (ts <= t) {
        float tt = t - ts;
        double a = (2 * ORA_PI) * (2 * ORA_PI) * (double)tt / (double)(tr * tr);
        return (float)(a * exp(-2 * ORA_PI * (double)tt / (double)tr));
    }
    return 0.0f;
}
/* m_fdtool.f90:379-396 */
static float stf_cosine(float t, float ts, float tr) {
    if (ts <= t && t <= ts + tr)
        return (float)((1 - cos(2 * ORA_PI * (double)(t - ts) / (double)tr)) / (double)tr);
    return 0.0f;
}
/* m_fdtool.f90:399-416 */
static float stf_boxcar(float t, float ts, float tr) {
    if (ts <= t && t <= ts + tr) return 1.0f / tr;
    return 0.0f;
}
/* m_fdtool.f90:419-438 */
static float stf_triangle(float t, float ts, float tr) {
    if (ts <= t && t <= ts + tr / 2) return 4 * (t - ts) / (tr * tr);
    if (ts + tr / 2 < t && t <= ts + tr) return -4 * (t - ts - tr) / (tr * tr);
    return 0.0f;
}
/* m_fdtool.f90:441-468 */
static float stf_herrmann(float t, float ts, float tr) {
    float t1 = ts + tr / 4;
    float t2 = ts + 3 * tr / 4;
    float tr3 = tr * tr * tr;
    if (ts <= t && t < t1) return 16 * ((t - ts) * (t - ts)) / tr3;
    if (t1 <= t && t < t2)
        return -2 * (8 * (t * t + tr * ts + ts * ts - t * tr - 2 * t * ts) + tr * tr) / tr3;
    if (t2 <= t && t <= ts + tr) return 16 * ((ts + tr - t) * (ts + tr - t)) / tr3;
    return 0.0f;
}

/* m_fdtool.f90:472-497; default branch is kupper (:494) */
float ora_momentrate(float t, const char *stftype, const float *srcprm) {
    float tbeg = srcprm[0], trise = srcprm[1];
    if (!strcmp(stftype, "boxcar")) return stf_boxcar(t, tbeg, trise);
    if (!strcmp(stftype, "triangle")) return stf_triangle(t, tbeg, trise);
    if (!strcmp(stftype, "herrmann")) return stf_herrmann(t, tbeg, trise);
    if (!strcmp(stftype, "kupper")) return stf_kupper(t, tbeg, trise);
    if (!strcmp(stftype, "cosine")) return stf_cosine(t, tbeg, trise);
    if (!strcmp(stftype, "texp")) return stf_texp(t, tbeg, trise);
    return stf_kupper(t, tbeg, trise);
}

/* ------------------------------------------------------------------ visco-elastic tau-method */
/* m_fdtool.f90:691-727 */
void ora_visco_set_relaxtime(int nm, float *ts, float fmin, float fmax) {
    float omega_a = (float)(2 * ORA_PI * (double)fmin);
    float omega_b = (float)(2 * ORA_PI * (double)fmax);
    float omega;
    if (nm == 0) return;
    if (nm == 1) {
        omega = sqrtf(omega_a * omega_b);
        ts[0] = 1.0f / omega;
        return;
    }
    for (int im = 1; im <= nm; im++) {
        double e = (double)(im - 1) / (double)(nm - 1);
        omega = (float)((double)omega_a * pow((double)(omega_b / omega_a), e));
        ts[im - 1] = 1.0f / omega;
    }
}

/* m_fdtool.f90:756-812 (Blanch's tau-method integrals); all real(SP) */
float ora_visco_constq_zeta(int nm, float fmin, float fmax, const float *ts) {
    if (nm == 0) return 0.0f;
    float i0[ORA_MAXNM], i1[ORA_MAXNM], i2[ORA_MAXNM][ORA_MAXNM];
    float om_a = (float)(2 * ORA_PI * (double)fmin);
    float om_b = (float)(2 * ORA_PI * (double)fmax);
    for (int im = 0; im < nm; im++) {
        float t = ts[im];
        i0[im] = (logf(1.0f + (om_b * om_b) * (t * t)) - logf(1.0f + (om_a * om_a) * (t * t))) / (2 * t);
        i1[im] = ((atanf(om_b * t) - om_b * t / (1 + (om_b * om_b) * (t * t))) -
                  (atanf(om_a * t) - om_a * t / (1 + (om_a * om_a) * (t * t)))) /
                 (2 * t);
    }
    for (int im = 0; im < nm - 1; im++)
        for (int km = im + 1; km < nm; km++) {
            float wk1 = atanf(om_b * ts[im]) / ts[im] - atanf(om_b * ts[km]) / ts[km];
            float wk2 = atanf(om_a * ts[im]) / ts[im] - atanf(om_a * ts[km]) / ts[km];
            i2[km][im] = ts[im] * ts[km] / (ts[km] * ts[km] - ts[im] * ts[im]) * (wk1 - wk2);
        }
    float i0sum = 0.0f, i1sum = 0.0f, i2sum = 0.0f;
    for (int im = 0; im < nm; im++) i0sum += i0[im];
    for (int im = 0; im < nm; im++) i1sum += i1[im];
    for (int im = 0; im < nm - 1; im++)
        for (int km = im + 1; km < nm; km++) i2sum = i2sum + i2[km][im];
    return i0sum / (i1sum + 2 * i2sum);
}

/* m_fdtool.f90:81-96 */
void ora_fdm_stable_dt(float dx, float dy, float dz, float vmax, float *dt) {
    float hh = 1.0f / sqrtf(1 / (dx * dx) + 1 / (dy * dy) + 1 / (dz * dz));
    float cc = 6.0f / 7.0f;
    *dt = cc * hh / vmax;
}

/* m_fdtool.f90:281-293 */
float ora_moment_magnitude(float m0) {
    if (m0 < 1.1920929e-07f) return -12345.0f;
    return (log10f(m0) - 9.1f) * 2.0f / 3.0f;
}

/* m_fdtool.f90:296-304 */
/* real(SP) ** integer as gfortran evaluates it (libgcc __powisf2: square-and-multiply in float) */
float ora_powi_sp(float x, int m) {
    unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
    float y = (n % 2) ? x : 1.0f;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? 1.0f / y : y;
}
float ora_seismic_moment(float mw) { return powf(10.0f, 1.5f * mw + 9.05f); }

/* m_std.f90:132-139  d2r_s = real(PI / 180.0_SP * deg) */
static float d2r_s(float deg) { return (float)(ORA_PI / (double)180.0f * (double)deg); }
/* m_std.f90:152-159 */
static float r2d_s(float rad) { return (float)((double)180.0f / ORA_PI * (double)rad); }
float ora_rad2deg_s(float rad) { return r2d_s(rad); }

float ora_deg2rad(float deg) { return d2r_s(deg); } /* std__deg2rad, m_std.f90:132-139 */

/* m_fdtool.f90:307-336 */
void ora_sdr2moment(float strike, float dip, float rake, float *mxx, float *myy, float *mzz,
                    float *myz, float *mxz, float *mxy) {
    float sind = sinf(d2r_s(dip)), cosd = cosf(d2r_s(dip));
    float sin2d = sinf(d2r_s(2 * dip)), cos2d = cosf(d2r_s(2 * dip));
    float sinl = sinf(d2r_s(rake)), cosl = cosf(d2r_s(rake));
    float sinf_ = sinf(d2r_s(strike)), cosf_ = cosf(d2r_s(strike));
    float sin2f = sinf(d2r_s(2 * strike)), cos2f = cosf(d2r_s(2 * strike));
    *mxx = -(sind * cosl * sin2f + sin2d * sinl * sinf_ * sinf_);
    *mxy = (sind * cosl * cos2f + sin2d * sinl * sin2f / 2);
    *mxz = -(cosd * cosl * cosf_ + cos2d * sinl * sinf_);
    *myy = (sind * cosl * sin2f - sin2d * sinl * cosf_ * cosf_);
    *myz = -(cosd * cosl * sinf_ - cos2d * sinl * cosf_);
    *mzz = (sin2d * sinl);
}

/* m_seawater.f90:34-47; epsil = 0.00737 with the Munk profile, 0 otherwise (:19-29) */
float ora_seawater_vel(float z, int use_munk) {
    double epsil = use_munk ? 0.00737 : 0.0;
    double zc = 1300.0;
    double zb = 2 * ((double)z * (double)1000.0f - zc) / zc;
    return (float)(1.5 * (1.0 + epsil * (zb - 1.0 + exp(-zb))));
}

/* ------------------------------------------------------------------ PML damping profile */
/* m_absorb_p.f90:533-573.  cp = 6 km/s, pd=1, pa=1, pb=2, b0=7, a0 = pi*fcut; all real(SP). */
void ora_damping_profile(float x, float H, float xbeg0, float xend0, int na, float fcut, float dt,
                         float g[4]) {
    const float cp = 6.0f;
    float R0 = powf(10.0f, -(log10f((float)na) - 1) / log10f(2.0f) - 3.0f);
    float d0 = -((1.0f / (2.0f * H)) * (float)(1 + 1) * cp * logf(R0));
    float b0 = 7.0f;
    float a0 = (float)(ORA_PI * (double)fcut);
    float xx;
    if (x <= xbeg0 + H)
        xx = (xbeg0 + H) - x;
    else if (x >= xend0 - H)
        xx = x - (xend0 - H);
    else
        xx = 0.0f;
    float q = fabsf(xx / H);
    float d = d0 * q;                          /* **pd, pd = 1 */
    float a = a0 * (1.0f - q);                 /* **pa, pa = 1 */
    float b = 1.0f + (b0 - 1.0f) * (q * q);    /* **pb, pb = 2 */
    float den = 1.0f + (dt / 2.0f) * (a + d / b);
    g[0] = ((1.0f + (dt / 2.0f) * a) / b) / den;
    g[1] = (-1.0f / b) / den;
    g[2] = (1.0f - (dt / 2.0f) * (a + d / b)) / den;
    g[3] = (d / b) / den;
}

/* ------------------------------------------------------------------ Gauss-Krueger (m_gk.f90) */
static double gk_alpha[6], gk_beta[6], gk_AA[6], gk_delta[7];
static int gk_first = 1;
static const double GK_a = 6378137.0, GK_F = 298.257222101, GK_m0 = 0.9999;
#define GK_n (1.0 / (2.0 * GK_F - 1.0))

/* m_gk.f90:189-223 */
static void gk_set_coef(void) {
    double n = GK_n;
    gk_alpha[1] = (1 / 2. + (-2 / 3. + (5 / 16. + (41 / 180. - 127 / 288. * n) * n) * n) * n) * n;
    gk_alpha[2] = (13 / 48. + (-3 / 5. + (557 / 1440. + 281 / 630. * n) * n) * n) * (n * n);
    gk_alpha[3] = (61 / 240. + (-103 / 140. + 15061 / 26880. * n) * n) * (n * n * n);
    gk_alpha[4] = (49561 / 161280. - 179 / 168. * n) * (n * n * n * n);
    gk_alpha[5] = 34729 / 80640. * (n * n * n * n * n);
    gk_beta[1] = (1 / 2. + (-2 / 3. + (37 / 96. + (-1 / 360. - 81 / 512. * n) * n) * n) * n) * n;
    gk_beta[2] = ((1 / 48. + (1 / 15. + (-437 / 1440. + 46 / 105. * n) * n) * n) * n) * n;
    gk_beta[3] = (((17 / 480. + (-37 / 840. - 209 / 4480. * n) * n) * n) * n) * n;
    gk_beta[4] = ((((4397 / 161280. - 11 / 504. * n) * n) * n) * n) * n;
    gk_beta[5] = ((((4583 / 161280. * n) * n) * n) * n) * n;
    gk_delta[1] = (2 / 1. + (-2 / 3. + (-2 / 1. + (116 / 45. + (26 / 45. + (-2854 / 675.) * n) * n) * n) * n) * n) * n;
    gk_delta[2] = ((7 / 3. + (-8 / 5. + (-227 / 45. + (2704 / 315. + (2323 / 945.) * n) * n) * n) * n) * n) * n;
    gk_delta[3] = (((56 / 15. + (-136 / 35. + (-1262 / 105. + (73814 / 2835.) * n) * n) * n) * n) * n) * n;
    gk_delta[4] = ((((4279 / 630. + (-332 / 35. + (-399572 / 14175.) * n) * n) * n) * n) * n) * n;
    gk_delta[5] = (((((4174 / 315. + (-144838 / 6237.) * n) * n) * n) * n) * n) * n;
    gk_delta[6] = ((((((601676 / 22275.) * n) * n) * n) * n) * n) * n;
    double n2 = n * n, n4 = n2 * n2;
    gk_AA[0] = 1 + (1 / 4. + 1 / 64. * n2) * n2;
    gk_AA[1] = -3 / 2. * (1. - 1 / 8. * n2 - 1 / 64. * n4) * n;
    gk_AA[2] = 15 / 16. * (1. - 1 / 4. * n2) * n2;
    gk_AA[3] = -35 / 48. * (1. - 5 / 16. * n2) * (n2 * n);
    gk_AA[4] = 315 / 512. * n4;
    gk_AA[5] = -693 / 1280. * (n4 * n);
    gk_first = 0;
}
static double gk_atanh0(double x) { return log((1. + x) / (1. - x)) / 2.0; }
static double gk_S_phi0(double phi0) {
    double s = gk_AA[0] * phi0;
    for (int j = 1; j <= 5; j++) s = s + gk_AA[j] * sin(2 * j * phi0);
    return s * (GK_m0 * GK_a / (1 + GK_n));
}
static double d2r_d(double deg) { return ORA_PI / 180.0 * deg; }
static double r2d_d(double rad) { return 180.0 / ORA_PI * rad; }

/* m_gk.f90:40-92 */
static void gk_lltoxy_d(double lon, double lat, double lon0, double lat0, double *x, double *y) {
    if (gk_first) gk_set_coef();
    double lam = d2r_d(lon), lam0 = d2r_d(lon0), phi = d2r_d(lat), phi0 = d2r_d(lat0);
    double n = GK_n;
    double e2n = 2.0 * sqrt(n) / (1.0 + n);
    double lam_c = cos(lam - lam0), lam_s = sin(lam - lam0);
    double tan_chi = sinh(gk_atanh0(sin(phi)) - e2n * atanh(e2n * sin(phi)));
    double cos_chi = sqrt(1 + tan_chi * tan_chi);
    double xi = atan(tan_chi / lam_c);
    double eta = gk_atanh0(lam_s / cos_chi);
    double Abar = GK_m0 * GK_a / (1 + n) * gk_AA[0];
    double xx = xi, yy = eta;
    for (int j = 1; j <= 5; j++) {
        xx = xx + gk_alpha[j] * sin(2 * j * xi) * cosh(2 * j * eta);
        yy = yy + gk_alpha[j] * cos(2 * j * xi) * sinh(2 * j * eta);
    }
    xx = Abar * xx - gk_S_phi0(phi0);
    yy = Abar * yy;
    *x = xx / 1000;
    *y = yy / 1000;
}

/* m_gk.f90:115-158 */
static void gk_xytoll_d(double x, double y, double lon0, double lat0, double *lon, double *lat) {
    if (gk_first) gk_set_coef();
    double lam0 = d2r_d(lon0), phi0 = d2r_d(lat0);
    double n = GK_n;
    double Abar = GK_m0 * GK_a / (1 + n) * gk_AA[0];
    double xi = (x * 1000 + gk_S_phi0(phi0)) / Abar;
    double eta = y * 1000 / Abar;
    double xi2 = xi, eta2 = eta;
    for (int j = 1; j <= 5; j++) {
        xi2 = xi2 - gk_beta[j] * sin(2 * j * xi) * cosh(2 * j * eta);
        eta2 = eta2 - gk_beta[j] * cos(2 * j * xi) * sinh(2 * j * eta);
    }
    double chi = asin(sin(xi2) / cosh(eta2));
    double lam = lam0 + atan(sinh(eta2) / cos(xi2));
    double phi = chi;
    for (int j = 1; j <= 6; j++) phi = phi + gk_delta[j] * sin(2 * j * chi);
    *lon = r2d_d(lam);
    *lat = r2d_d(phi);
}

/* m_geomap.f90:18-41 */
void ora_geomap_g2c(float lon, float lat, float lon0, float lat0, float phi, float *x, float *y) {
    float phi_r = d2r_s(phi);
    double xd, yd;
    gk_lltoxy_d((double)lon, (double)lat, (double)lon0, (double)lat0, &xd, &yd);
    float xx = (float)xd, yy = (float)yd;
    *x = cosf(phi_r) * xx + sinf(phi_r) * yy;
    *y = -sinf(phi_r) * xx + cosf(phi_r) * yy;
}

/* m_geomap.f90:44-65 */
void ora_geomap_c2g(float x, float y, float lon0, float lat0, float phi, float *lon, float *lat) {
    float phi_r = d2r_s(phi);
    float xx = cosf(phi_r) * x - sinf(phi_r) * y;
    float yy = sinf(phi_r) * x + cosf(phi_r) * y;
    double lo, la;
    gk_xytoll_d((double)xx, (double)yy, (double)lon0, (double)lat0, &lo, &la);
    *lon = (float)lo;
    *lat = (float)la;
}

double ora_r_earth(void) { return ORA_R_EARTH; }
double ora_pi(void) { return ORA_PI; }
