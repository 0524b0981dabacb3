// psv_driver.cpp -- host-side mirror of swpc_psv's setup chain and driver loop (include/swpcpsv_host.h).
//
// Setup-only CPU code: it produces, for ONE rank, the arrays the reference's setup modules hand to the time loop (which
// runs only on the GPU, through swpcpsv_b200.h).  Every routine cites the reference lines whose result it must reproduce
// (paths under /root/reference/src/swpc_psv, OpenSWPC 25.05.2); kinds follow the Fortran declarations.
#include "../../../include/swpcpsv_host.h"

#include "common.hpp"
#include "models.hpp"

struct swpcpsv_host {
    bool benchmark_mode = false;
    std::string title, odir, abc_type, vmodel_type, stftype, stf_format, sdep_fit, wav_format, st_format, fn_stf, fn_stloc, base;
    int nproc_x = 1, nx = 256, nz = 256, nt = 1000, ipad = 0, kpad = 0, na = 20, nm = 3;
    double dx = 0.5, dz = 0.5;
    float dt = 0.01f, xbeg = 0, zbeg = 0, tbeg = 0, xend = 0, zend = 0, clon = 0, clat = 0, phi = 0;
    float fq_min = 0.05f, fq_max = 5.0f, fq_ref = 1.0f, vcut = 0.0f;
    bool pw_mode = false, bf_mode = false, earth_flattening = false;
    int ntdec_w = 10, ntdec_r = 10, ntw = 0;
    int ntdec_w_prg = 0;   // m_wav.f90:71, :309-311: waveform files rewritten every ntdec_w_prg steps while the run goes on
    bool sw[4] = {false, false, false, false};   // v u stress strain
    float vmin = 0, vmax = 0, vmin_local = 0, vmax_local = 0, fmax = 0, fcut = 0, M0 = 0, UC = 1e-12f, zeta = 0, d2 = 0;
    float ts[8] = {}, c1[8] = {}, c2[8] = {}, d1[8] = {};
    float evlo = 0, evla = 0, evdp = 0, mxx0 = 0, mzz0 = 0, mxz0 = 0, fx0 = 0, fz0 = 0, otim = 0, sx0 = 0, sy0 = 0;
    int exedate = 0, tz_minutes = 0, field_bytes = 8;
    int myid = 0, nxp = 0, ibeg = 0, iend = 0, ibeg_m = 0, iend_m = 0, kbeg_m = -2, kend_m = 0, nzm = 0, nxm = 0, ibeg_k = 0, iend_k = 0, kend_k = 0;
    std::vector<float> xc, zc, rho, lam, mu, taup, taus, bddep;
    std::vector<int> kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot, kbeg_a;
    std::vector<float> gxc, gxe, gzc, gze, cgx_c, cgx_b, cgz_c, cgz_b;
    std::vector<int> src_ik, st_ik;
    std::vector<double> mo, m3;
    std::vector<float> srcprm, xst, zst, stlo, stla;
    std::vector<std::string> stnm;
    swpcpsv_handle *dev = nullptr;
    // snapshots (m_snap.f90): products 0 ps (divergence, rotation), 1 v, 2 u over the decimated xz plane
    struct Snap {
        bool sw[3] = {false, false, false}, opened = false, native = true;
        int idec = 1, kdec = 1, ntdec_s = 10, nxs = 0, nzs = 0, is0 = 0, is1 = -1, ks0 = 0, ks1 = -1, ionode[3] = {0, 0, 0};
        std::vector<float> xsnp, zsnp, tmp;
        NcFile *nc[3] = {nullptr, nullptr, nullptr};
        FILE *snp[3] = {nullptr, nullptr, nullptr};
        float vmin[3][2] = {}, vmax[3][2] = {};
        ~Snap() { for (int q = 0; q < 3; q++) { delete nc[q]; if (snp[q]) std::fclose(snp[q]); } }
    } snap;
    void setup_snap(const IniFile &ini);
    int snap_open(const std::string &dir);
    int snap_write(int it);
    int snap_close();
    std::vector<float> wav_all[4];
    double loop_seconds = 0;

    size_t i2(int k, int i) const { return (size_t)(k - kbeg_m) + (size_t)nzm * (size_t)(i - ibeg_m); }
    int setup(const IniFile &ini, int nm_, int myid_, int npx, int nt_o);
    int setup_global(const IniFile &ini, int npx, int nt_o);
    int setup_geometry();
    int setup_medium(const IniFile &ini);
    void surface_detection();
    bool lateral_model = false;   // the medium varies along x (lgm_rmed excepted, all models.hpp builders)
    bool stabilize_pending = false;
    void apply_stabilize() {
        if (!stabilize_pending) return;
        stabilize_pending = false;
        const MediumBox mb{ibeg_m, iend_m, 1, 1, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
        stabilize_absorber(mb, kbeg_a.data(), ibeg, iend, 1, 1, nz, vmax);
    }
    int setup_source(const IniFile &ini);
    int setup_planewave(const IniFile &ini);
    std::vector<double> pw_init[5];   // plane-wave initial fields Vx Vz Sxx Szz Sxz over the memory box
    void setup_absorb();
    int setup_wav(const IniFile &ini);
};

int swpcpsv_host::setup_global(const IniFile &ini, int npx, int nt_o) {   // m_global.f90:116-153, :183-185
    benchmark_mode = ini.get_l("benchmark_mode", false);
    title = ini.get("title", "swpc_psv");
    nproc_x = ini.get_i("nproc_x", 1);
    nx = ini.get_i("nx", 256); nz = ini.get_i("nz", 256); nt = ini.get_i("nt", 1000);
    ipad = ini.get_i("ipad", 0); kpad = ini.get_i("kpad", 0);
    odir = ini.get("odir", "./out");
    if (benchmark_mode) {
        dx = dz = 0.5f; dt = 0.04f; na = 20;
        xbeg = -((float)nx / 2.0f * (float)dx);
        zbeg = -30 * (float)dz;
        tbeg = 0.0f; clon = 139.7604f; clat = 35.7182f; phi = 0.0f; abc_type = "pml";
    } else {
        dx = ini.get_d("dx", 0.5); dz = ini.get_d("dz", 0.5);
        dt = ini.get_s("dt", 0.01f);
        na = ini.get_i("na", 20);
        xbeg = ini.get_s("xbeg", -(float)(nx / 2) * (float)dx);
        zbeg = ini.get_s("zbeg", -30 * (float)dz);
        tbeg = ini.get_s("tbeg", 0.0f);
        clon = ini.get_s("clon", 139.7604f); clat = ini.get_s("clat", 35.7182f); phi = ini.get_s("phi", 0.0f);
        abc_type = ini.get("abc_type", "pml");
    }
    if (npx > 0) nproc_x = npx;
    if (nt_o > 0) nt = nt_o;
    xend = xbeg + nx * (float)dx; zend = zbeg + nz * (float)dz;
    if (myid < 0 || myid >= nproc_x) return hfail("myid outside of nproc_x (assert, m_global.f90:199)");
    if (abc_type != "pml" && abc_type != "cerjan") return hfail("abc_type must be 'pml' or 'cerjan' (assert, m_absorb.f90:37)");
    UC = 1e-12f;   // m_global.f90:28
    const time_t now = time(nullptr);
    struct tm lt;
    localtime_r(&now, &lt);
    exedate = (int)now;
    tz_minutes = (int)(lt.tm_gmtoff / 60);
    return 0;
}

int swpcpsv_host::setup_geometry() {   // m_global.f90:189-310
    const int mx = nx % nproc_x, proc_x = myid;
    nxp = (proc_x <= nproc_x - mx + 1) ? (nx - mx) / nproc_x : (nx - mx) / nproc_x + 1;   // :201-208 (sic: "+ 1")
    if (proc_x <= nproc_x - mx - 1) { ibeg = proc_x * (nx - mx) / nproc_x + 1; iend = (proc_x + 1) * (nx - mx) / nproc_x; }   // :232-238
    else { ibeg = proc_x * ((nx - mx) / nproc_x + 1) - (nproc_x - mx) + 1; iend = (proc_x + 1) * ((nx - mx) / nproc_x + 1) - (nproc_x - mx); }
    ibeg_m = ibeg - 3; iend_m = iend + 3 + ipad; kbeg_m = -2; kend_m = nz + 3 + kpad;
    nzm = kend_m - kbeg_m + 1; nxm = iend_m - ibeg_m + 1;
    if ((ibeg_m <= na && iend_m < na + 1) || (iend_m >= nx - na + 1 && ibeg_m > nx - na))
        return hfail("subdomain narrower than the absorber: the reference's absorber homogenisation (m_medium.f90:118-137) would read out of bounds");
    xc.resize(nxm); zc.resize(nzm);
    for (int i = ibeg_m; i <= iend_m; i++) xc[i - ibeg_m] = i2x(i, xbeg, (float)dx);
    for (int k = kbeg_m; k <= kend_m; k++) zc[k - kbeg_m] = i2x(k, zbeg, (float)dz);
    kbeg_a.assign(nxm, 0);
    for (int i = ibeg_m; i <= iend_m; i++) kbeg_a[i - ibeg_m] = (i <= na || nx - na + 1 <= i) ? 1 : nz - na + 1;
    ibeg_k = ibeg; iend_k = iend; kend_k = nz;
    if (abc_type == "pml") {
        if (iend <= na) ibeg_k = iend + 1; else if (ibeg <= na) ibeg_k = na + 1;
        if (ibeg >= nx - na + 1) iend_k = ibeg - 1; else if (iend >= nx - na + 1) iend_k = nx - na;
        kend_k = nz - na;
    }
    return 0;
}

// m_medium.f90:260-292.  kfs_top is assigned twice and kfs_bot never (:281-282): it keeps its allocation value, taken as 0.
void swpcpsv_host::surface_detection() {
    kfs.assign(nxm, 0); kob.assign(nxm, 0); kfs_top.assign(nxm, 0); kfs_bot.assign(nxm, 0); kob_top.assign(nxm, 0); kob_bot.assign(nxm, 0);
    for (int i = ibeg - 1; i <= iend + 2; i++)
        for (int k = 1; k <= nz - 1; k++) {
            const size_t a = i2(k, i), b = i2(k + 1, i);
            if (std::fabs(mu[a]) < EPS_SP && std::fabs(mu[b]) > EPS_SP) kob[i - ibeg_m] = k;
            if (std::fabs(lam[a]) < EPS_SP && std::fabs(lam[b]) > EPS_SP) kfs[i - ibeg_m] = k;
        }
    for (int i = ibeg; i <= iend; i++) {
        int f0 = 1 << 30, f1 = -(1 << 30), o0 = 1 << 30, o1 = -(1 << 30);
        for (int ii = i - 2; ii <= i + 3; ii++) {
            f0 = std::min(f0, kfs[ii - ibeg_m]); f1 = std::max(f1, kfs[ii - ibeg_m]);
            o0 = std::min(o0, kob[ii - ibeg_m]); o1 = std::max(o1, kob[ii - ibeg_m]);
        }
        const int q = i - ibeg_m;
        kfs_top[q] = std::max(f0 - 2, 1);
        kfs_top[q] = std::min(f1 + 2, nz);   // sic
        kob_top[q] = std::max(o0 - 2, 1);
        kob_bot[q] = std::min(o1 + 2, nz);
    }
}

int swpcpsv_host::setup_medium(const IniFile &ini) {   // m_medium.f90:36-178
    const size_t nc = (size_t)nzm * nxm;
    rho.assign(nc, 0.f); lam.assign(nc, 0.f); mu.assign(nc, 0.f); taup.assign(nc, 0.f); taus.assign(nc, 0.f);
    bddep.assign((size_t)nxm * (NBD + 1), -9999.0f);
    // uni / lhm / benchmark are laterally uniform: one depth profile, broadcast over the columns; the builders of models.hpp
    // (lgm and the random-media family) fill the section itself (`lateral`)
    std::vector<float> p_rho(nzm), p_lam(nzm), p_mu(nzm), p_qp(nzm), p_qs(nzm);
    float bd0 = 0.0f;
    bool lateral = false, grd_bddep = false;   // grd: the builder wrote the boundary depths bd(:, 0:NBD) itself
    if (benchmark_mode) {   // :53-72
        fq_min = 0.05f; fq_max = 5.0f; fq_ref = 1.0f;
        for (int q = 0; q < nzm; q++) {
            if (zc[q] < 0.0f) { p_rho[q] = 0.001f; p_mu[q] = 0.0f; p_lam[q] = 0.0f; }
            else { p_rho[q] = 2.7f; p_mu[q] = 2.7f * 3.5f * 3.5f; p_lam[q] = 2.7f * 3.5f * 3.5f; }
            p_qp[q] = 1e10f; p_qs[q] = 1e10f;
        }
    } else {
        fq_min = ini.get_s("fq_min", 0.05f); fq_max = ini.get_s("fq_max", 5.00f); fq_ref = ini.get_s("fq_ref", 1.00f);
        vmodel_type = ini.get("vmodel_type", "uni");
        vcut = ini.get_s("vcut", 0.0f);
        const bool munk = ini.get_l("munk_profile", false), ef = ini.get_l("earth_flattening", false);
        std::vector<float> zs(nzm), Cv(nzm);
        for (int q = 0; q < nzm; q++) {
            if (ef) { zs[q] = (float)(R_EARTH - R_EARTH * std::exp(-(double)zc[q] / R_EARTH)); Cv[q] = (float)std::exp((double)zc[q] / R_EARTH); }
            else { zs[q] = zc[q]; Cv[q] = 1.0f; }
        }
        if (vmodel_type == "uni") {   // m_vmodel_uni.f90:45-129
            const float vp0 = ini.get_s("vp0", 5.0f);
            const float vs0 = ini.get_s("vs0", vp0 / std::sqrt(3.0f));
            const float rho0 = ini.get_s("rho0", 2.7f), qp0 = ini.get_s("qp0", 1000000.0f), qs0 = ini.get_s("qs0", 1000000.0f);
            bd0 = ini.get_s("topo0", 0.0f);
            for (int q = 0; q < nzm; q++) {
                float vp1, vs1, r1;
                if (zs[q] > bd0) { vp1 = Cv[q] * vp0; vs1 = Cv[q] * vs0; r1 = rho0; p_qp[q] = qp0; p_qs[q] = qs0; }
                else if (zc[q] > 0.0f) { vp1 = Cv[q] * seawater_vel(zs[q], munk); vs1 = 0.0f; r1 = 1.0f; p_qp[q] = 1000000.0f; p_qs[q] = 1000000.0f; }
                else { vp1 = 0.0f; vs1 = 0.0f; r1 = 0.001f; p_qp[q] = 10.0f; p_qs[q] = 10.0f; }
                p_rho[q] = r1; p_mu[q] = r1 * vs1 * vs1; p_lam[q] = r1 * (vp1 * vp1 - 2 * vs1 * vs1);
            }
        } else if (vmodel_type == "lhm") {   // m_vmodel_lhm.f90:52-145
            const std::string fn = join_path(base, ini.get("fn_lhm", ""));
            std::ifstream is(fn);
            if (!is) return hfail("vmodel_lhm: cannot open " + fn + " (assert, m_vmodel_lhm.f90:53-54)");
            std::vector<float> depth, r0, vp0, vs0, qp0, qs0;
            std::string line;
            while (std::getline(is, line)) {
                if (blank_or_comment(line)) continue;
                const std::vector<float> v = parse_reals(line);
                if (v.size() < 6) continue;
                depth.push_back(v[0]); r0.push_back(v[1]); vp0.push_back(v[2]); vs0.push_back(v[3]); qp0.push_back(v[4]); qs0.push_back(v[5]);
            }
            const int nl = (int)depth.size();
            if (nl == 0) return hfail("vmodel_lhm: no layer in " + fn);
            for (int l = nl - 2; l >= 0; l--)   // :85-94
                if ((vp0[l] < vcut || vs0[l] < vcut) && (vp0[l] > 0 && vs0[l] > 0)) {
                    vp0[l] = vp0[l + 1]; vs0[l] = vs0[l + 1]; r0[l] = r0[l + 1]; qp0[l] = qp0[l + 1]; qs0[l] = qs0[l + 1];
                }
            bd0 = depth[0];
            for (int q = 0; q < nzm; q++) {
                float r1 = 0, vp1 = 0, vs1 = 0, a = 0, b = 0;
                if (zs[q] < depth[0]) {
                    if (zs[q] < 0.0f) { r1 = 0.001f; a = 10.0f; b = 10.0f; }
                    else { r1 = 1.0f; vp1 = Cv[q] * seawater_vel(zs[q], munk); a = 1000000.0f; b = 1000000.0f; }   // :111 seawater__vel(zs(k))
                } else {
                    for (int l = 0; l < nl; l++)
                        if (zs[q] >= depth[l]) { r1 = r0[l]; vp1 = Cv[q] * vp0[l]; vs1 = Cv[q] * vs0[l]; a = qp0[l]; b = qs0[l]; }
                }
                p_rho[q] = r1; p_mu[q] = r1 * vs1 * vs1; p_lam[q] = r1 * (vp1 * vp1 - 2 * vs1 * vs1); p_qp[q] = a; p_qs[q] = b;
            }
        } else if (vmodel_type == "grd" || vmodel_type == "grd_rmed") {
            // swpc_psv/m_vmodel_grd.f90, m_vmodel_grd_rmed.f90: GMT grids (netCDF classic) sampled along the section y = 0
            lateral = grd_bddep = true;
            const MediumBox mb{ibeg_m, iend_m, 1, 1, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
            ModelEnv env{&ini, base, vcut, dt, dx, 0.0, dz, munk, ef};
            env.psv = true;
            const GrdGeometry gg{nx, 0, na, xbeg, 0.0f, zbeg, clon, clat, phi, xc.data(), nullptr, bddep.data(), nz};
            if (vmodel_grd(env, mb, gg, vmodel_type == "grd_rmed")) return 1;
        } else if (vmodel_type == "lgm" || vmodel_type == "uni_rmed" || vmodel_type == "lhm_rmed" || vmodel_type == "lgm_rmed") {
            // swpc_psv/m_vmodel_{lgm,uni_rmed,lhm_rmed,lgm_rmed}.f90: the 3-D builders without the y axis (ModelEnv::psv marks
            // the few expressions that differ); Qp / Qs go to taup / taus, as in the reference's call (m_medium.f90:101-114)
            lateral = true;
            const MediumBox mb{ibeg_m, iend_m, 1, 1, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
            ModelEnv env{&ini, base, vcut, dt, dx, 0.0, dz, munk, ef};
            env.psv = true;
            const int rc = vmodel_type == "lgm" ? vmodel_lgm(env, mb, bd0) : vmodel_type == "uni_rmed" ? vmodel_uni_rmed(env, mb, bd0)
                         : vmodel_type == "lhm_rmed" ? vmodel_lhm_rmed(env, mb, bd0) : vmodel_lgm_rmed(env, mb, bd0);
            if (rc) return 1;
        } else {
            return hfail("swpc_psv vmodel_type '" + vmodel_type + "' is outside the scope of this build ('user' is a compile-time plug-in of the reference)");
        }
    }
    if (!grd_bddep) for (int i = 0; i < nxm; i++) bddep[i] = bd0;
    lateral_model = lateral;
    if (lateral) {
        // absorber homogenisation (:118-148): the side absorbers repeat the column next to them (when this rank holds it),
        // the bottom absorber repeats the row k = nz - na
        const auto copy_col = [&](int dst, int src) {
            const size_t a = i2(kbeg_m, dst), b = i2(kbeg_m, src);
            for (std::vector<float> *f : {&rho, &lam, &mu, &taup, &taus}) std::copy(f->begin() + b, f->begin() + b + nzm, f->begin() + a);
        };
        if (na + 1 >= ibeg_m && na + 1 <= iend_m) for (int i = ibeg_m; i <= na; i++) copy_col(i, na + 1);
        if (nx - na >= ibeg_m && nx - na <= iend_m) for (int i = std::max(nx - na + 1, ibeg_m); i <= iend_m; i++) copy_col(i, nx - na);
        for (int i = ibeg_m; i <= iend_m; i++) {
            const size_t s0 = i2(nz - na, i);
            for (int k = nz - na + 1; k <= kend_m; k++) {
                const size_t n = i2(k, i);
                rho[n] = rho[s0]; lam[n] = lam[s0]; mu[n] = mu[s0]; taup[n] = taup[s0]; taus[n] = taus[s0];
            }
        }
        relax_times(nm, ts, fq_min, fq_max);   // :151-162
        zeta = constq_zeta(nm, fq_min, fq_max, ts);
        const long long ncl = (long long)nc;
#pragma omp parallel for schedule(static)
        for (long long n = 0; n < ncl; n++) { taup[n] = nm * zeta / taup[n]; taus[n] = nm * zeta / taus[n]; }
        if (nm > 0) {   // relaxed_medium :181-204 + visco_chi src/shared/m_fdtool.f90:730-753
            const float omega = (float)(2 * PI_D * (double)fq_ref);
            std::complex<float> cc(0.0f, 0.0f);
            for (int m = 0; m < nm; m++) {
                const std::complex<double> w = std::complex<double>(0.0, 1.0) * (double)omega * (double)ts[m];
                const std::complex<double> qd = w / (1.0 - w);
                cc = cc + std::complex<float>((float)qd.real(), (float)qd.imag());
            }
            cc = std::complex<float>(cc.real() / (float)nm, cc.imag() / (float)nm);
#pragma omp parallel for schedule(static)
            for (long long n = 0; n < ncl; n++) {
                const float rb2 = mu[n], ra2 = lam[n] + 2 * mu[n];
                const std::complex<float> zs_ = 1.0f - cc * taus[n], zp_ = 1.0f - cc * taup[n];
                const float chi_mu = 1.0f / (1.0f / std::sqrt(zs_)).real();
                const float chi_lam = 1.0f / (1.0f / std::sqrt(zp_)).real();
                mu[n] = rb2 / (chi_mu * chi_mu);
                lam[n] = ra2 / (chi_lam * chi_lam) - 2 * mu[n];
            }
        }
    } else {
    // absorber homogenisation (:118-148): identity in x for laterally uniform input; in z it repeats the value at k = nz-na
    for (int q = nz - na + 1 - kbeg_m; q < nzm; q++) {
        const int s = nz - na - kbeg_m;
        p_rho[q] = p_rho[s]; p_lam[q] = p_lam[s]; p_mu[q] = p_mu[s]; p_qp[q] = p_qp[s]; p_qs[q] = p_qs[s];
    }
    relax_times(nm, ts, fq_min, fq_max);   // :151-162
    zeta = constq_zeta(nm, fq_min, fq_max, ts);
    if (benchmark_mode) zeta = 0.0f;
    std::vector<float> p_tp(nzm), p_tsx(nzm);
    for (int q = 0; q < nzm; q++) { p_tp[q] = nm * zeta / p_qp[q]; p_tsx[q] = nm * zeta / p_qs[q]; }
    if (nm > 0) {   // relaxed_medium :181-204 + visco_chi src/shared/m_fdtool.f90:730-753
        const float omega = (float)(2 * PI_D * (double)fq_ref);
        std::complex<float> cc(0.0f, 0.0f);
        for (int m = 0; m < nm; m++) {
            const std::complex<double> w = std::complex<double>(0.0, 1.0) * (double)omega * (double)ts[m];
            const std::complex<double> qd = w / (1.0 - w);
            cc = cc + std::complex<float>((float)qd.real(), (float)qd.imag());
        }
        cc = std::complex<float>(cc.real() / (float)nm, cc.imag() / (float)nm);
        for (int q = 0; q < nzm; q++) {
            const float rb2 = p_mu[q], ra2 = p_lam[q] + 2 * p_mu[q];
            const std::complex<float> zs_ = 1.0f - cc * p_tsx[q], zp_ = 1.0f - cc * p_tp[q];
            const float chi_mu = 1.0f / (1.0f / std::sqrt(zs_)).real();
            const float chi_lam = 1.0f / (1.0f / std::sqrt(zp_)).real();
            p_mu[q] = rb2 / (chi_mu * chi_mu);
            p_lam[q] = ra2 / (chi_lam * chi_lam) - 2 * p_mu[q];
        }
    }
    for (int i = 0; i < nxm; i++) {
        const size_t o = (size_t)nzm * i;
        std::copy(p_rho.begin(), p_rho.end(), rho.begin() + o); std::copy(p_lam.begin(), p_lam.end(), lam.begin() + o);
        std::copy(p_mu.begin(), p_mu.end(), mu.begin() + o); std::copy(p_tp.begin(), p_tp.end(), taup.begin() + o);
        std::copy(p_tsx.begin(), p_tsx.end(), taus.begin() + o);
    }
    }   // laterally uniform
    surface_detection();
    float vmx = -1.0f, vmn = 1e30f;   // velocity_minmax :294-318 (local part)
    for (int i = ibeg; i <= iend; i++)
        for (int k = kfs[i - ibeg_m] + 1; k <= nz; k++) {
            const size_t n = i2(k, i);
            const float vp = std::sqrt((lam[n] + 2 * mu[n]) / rho[n]), vs = std::sqrt(mu[n] / rho[n]);
            vmx = std::max(vmx, vp);
            if (vs < EPS_SP) continue;
            vmn = std::min(vmn, vs);
        }
    vmin_local = vmin = vmn;
    vmax_local = vmax = vmx;
    // stabilize_absorber (m_medium.f90:171-174, :309-366) needs the GLOBAL vmax: applied here for a single-rank run, otherwise
    // when the caller hands over the reduced values (swpcpsv_host_set_minmax) or, at the latest, before the upload
    stabilize_pending = ini.get_l("stabilize_pml", false);
    if (stabilize_pending && nproc_x == 1) apply_stabilize();
    d2 = 0.0f;   // kernel__setup m_kernel.f90:57-66 (the device computes its own copy; kept for reporting)
    if (nm > 0) {
        float sum = 0.0f;
        for (int m = 0; m < nm; m++) {
            c1[m] = (2 * ts[m] - dt) / (2 * ts[m] + dt); c2[m] = (2) / (2 * ts[m] + dt) / nm; d1[m] = 2 * ts[m] / (2 * ts[m] - dt);
            sum += dt / (2 * ts[m] - dt);
        }
        d2 = sum / nm;
    }
    return 0;
}

// pw_setup, m_source.f90:656-785: plane P / SV wave as the initial condition of the five fields over the memory box;
// pw_strike and pw_rake are read and then forced to 90 degrees (:684-686)
int swpcpsv_host::setup_planewave(const IniFile &ini) {
    const float pw_ztop = ini.get_s("pw_ztop", 1e30f);
    if (!(pw_ztop < zend)) return hfail("assert: pw_ztop < zend (m_source.f90:672)");
    const float pw_zlen = ini.get_s("pw_zlen", -1.0f);
    if (!(pw_zlen > 0.0f)) return hfail("assert: pw_zlen > 0 (m_source.f90:675)");
    const std::string ps = ini.get("pw_ps", "");
    const bool is_p = ps == "p" || ps == "P", is_s = ps == "s" || ps == "S";
    if (!(is_p || is_s)) return hfail("assert: pw_ps must be 'p' or 's' (m_source.f90:678)");
    const float strike = deg2rad_s(90.0f), dip = deg2rad_s(ini.get_s("pw_dip", 0.0f)), rake = deg2rad_s(90.0f);
    stftype = ini.get("stftype", "kupper");
    const float sd = std::sin(dip), cd = std::cos(dip), sf = std::sin(strike), cf = std::cos(strike), sl = std::sin(rake), cl = std::cos(rake);
    const float c2d = std::cos(2 * dip);
    const size_t nc = (size_t)nzm * nxm;
    for (auto &f : pw_init) f.assign(nc, 0.0);
    auto speed = [&](size_t n) { return is_p ? std::sqrt((lam[n] + 2 * mu[n]) / rho[n]) : std::sqrt(mu[n] / rho[n]); };
    for (int i = ibeg_m; i <= iend_m; i++)
        for (int k = kbeg_m; k <= kend_m; k++) {
            const size_t n = i2(k, i);
            const float la0 = lam[n], mu0 = mu[n], v = speed(n);
            if (v < EPS_SP) continue;
            const float x0 = (float)(xbeg + (i - 0.5f) * dx), z0 = (float)(zbeg + (k - 0.5f) * dz - pw_ztop);
            const float x1 = (float)(x0 + dx / 2.0f), z1 = (float)(z0 + dz / 2.0f);
            const float stf_ii = momentrate(sd * sf * x0 + cd * z0, stftype, 0.0f, pw_zlen);
            const float stf_vx = momentrate(sd * sf * x1 + cd * z0 + dt / 2.0f * v, stftype, 0.0f, pw_zlen);
            const float stf_vz = momentrate(sd * sf * x0 + cd * z1 + dt / 2.0f * v, stftype, 0.0f, pw_zlen);
            const float stf_xz = momentrate(sd * sf * x1 + cd * z1, stftype, 0.0f, pw_zlen);
            float vx, vz, sxx, szz, sxz;
            if (is_p) {
                vx = -sd * sf * stf_vx; vz = -cd * stf_vz;
                sxx = -(la0 + 2 * mu0 * sd * sd * sf * sf) * stf_ii / v;
                szz = -(la0 + 2 * mu0 * cd * cd) * stf_ii / v;
                sxz = -2 * mu0 * sd * cd * sf * stf_xz / v;
            } else {
                vx = (cl * cf + sl * cd * sf) * stf_vx; vz = -sl * sd * stf_vz;
                sxx = 2 * mu0 * sd * sf * (cl * cf + sl * cd * sf) * stf_ii / v;
                szz = -2 * mu0 * cd * sl * sd * stf_ii / v;
                sxz = mu0 * (cl * cd * cf + sl * c2d * sf) * stf_xz / v;
            }
            pw_init[0][n] = vx; pw_init[1][n] = vz; pw_init[2][n] = sxx; pw_init[3][n] = szz; pw_init[4][n] = sxz;
        }
    // wavelength condition :764-783: the velocity at the model's centre column, MPI_MAX over the ranks.  A laterally uniform
    // model lets every rank evaluate it on a column of its own; otherwise only the rank holding that column can
    const int kc = x2i(pw_ztop, zbeg, (float)dz);
    int ic = ibeg;
    if (lateral_model) {
        ic = x2i((xbeg + xend) / 2, xbeg, (float)dx);
        if (!(ibeg <= ic && ic <= iend))
            return hfail("pw_mode with a laterally varying vmodel_type on several ranks: fcut comes from the rank holding the centre column "
                         "(m_source.f90:765-781); run this host on one rank or use a 1-D model");
    }
    fcut = speed(i2(kc, ic)) / pw_zlen;
    fmax = fcut * 2.0f;
    return 0;
}

int swpcpsv_host::setup_source(const IniFile &ini) {   // m_source.f90:41-258
    pw_mode = ini.get_l("pw_mode", false);
    if (pw_mode && !benchmark_mode) {   // :68-76: no regular source grid, fictitious scalar moment for the output scaling
        if (setup_planewave(ini)) return 1;
        M0 = 1.0f / UC;
        return 0;
    }
    bf_mode = ini.get_l("bf_mode", false);
    fn_stf = ini.get("fn_stf", "");
    stftype = ini.get("stftype", "kupper");
    stf_format = ini.get("stf_format", "xym0ij");
    sdep_fit = ini.get("sdep_fit", "asis");
    earth_flattening = ini.get_l("earth_flattening", false);
    if (stftype == "scosine") stftype = "cosine";
    struct Src { float x, z, t0, tr, mo, m[3]; };
    std::vector<Src> g;
    if (benchmark_mode) {   // :99-104, :131-139
        stftype = "kupper"; bf_mode = false;
        Src s{};
        s.x = 0.0f; s.z = 5.0f; s.mo = 1e15f; s.m[0] = 1 / std::sqrt(2.0f); s.m[1] = 1 / std::sqrt(2.0f); s.m[2] = 0.0f; s.t0 = 0.1f; s.tr = 2.0f;
        g.push_back(s);
    } else {
        const std::string fn = join_path(base, fn_stf);
        std::ifstream is(fn);
        if (!is) return hfail("source__setup: cannot open " + fn);
        const bool ll = stf_format.compare(0, 2, "ll") == 0, xy = stf_format.compare(0, 2, "xy") == 0;
        const std::string kind = stf_format.size() >= 6 ? stf_format.substr(2, 4) : "";
        std::string line;
        while (std::getline(is, line)) {
            if (blank_or_comment(line)) continue;
            const std::vector<float> v = parse_reals(line);
            Src s{};
            float sy = 0.0f;
            if (bf_mode) {   // source__grid_bodyforce :513-543  x y z tbeg trise fx fy fz
                if (v.size() < 8 || !(ll || xy)) return hfail("source file: bad body-force record / invalid source type");
                if (xy) s.x = v[0]; else geomap_g2c(v[0], v[1], clon, clat, phi, s.x, sy);
                s.z = v[2]; s.t0 = v[3]; s.tr = v[4]; s.m[0] = v[5]; s.m[1] = v[7];
                if (g.empty()) { geomap_c2g(s.x, 0.0f, clon, clat, phi, evlo, evla); evdp = s.z; otim = s.t0; fx0 = s.m[0]; fz0 = s.m[1]; }
                g.push_back(s);
                continue;
            }
            if (!(ll || xy) || !(kind == "m0ij" || kind == "m0dc" || kind == "mwij" || kind == "mwdc"))
                return hfail("swpc_psv stf_format '" + stf_format + "' is outside the scope of this build");
            const size_t need = kind[2] == 'i' ? 12 : 9;
            if (v.size() < need) return hfail("source file: bad moment record (assert(ierr == 0), m_source.f90:309)");
            if (xy) { s.x = v[0]; sy = v[1]; } else geomap_g2c(v[0], v[1], clon, clat, phi, s.x, sy);
            s.z = v[2]; s.t0 = v[3]; s.tr = v[4];
            s.mo = kind[1] == '0' ? v[5] : seismic_moment(v[5]);
            if (kind[2] == 'i') { s.m[0] = v[6]; s.m[1] = v[8]; s.m[2] = v[10]; }   // mxx myy mzz myz mxz mxy -> mxx mzz mxz (:307-308)
            else { float m6[6]; sdr2moment(v[6] - phi, v[7], v[8], m6); s.m[0] = m6[0]; s.m[1] = m6[2]; s.m[2] = m6[4]; }
            if (g.empty()) {   // :454-466
                geomap_c2g(s.x, sy, clon, clat, phi, evlo, evla);
                sx0 = s.x; sy0 = sy; evdp = s.z; mxx0 = s.m[0]; mzz0 = s.m[1]; mxz0 = s.m[2]; otim = s.t0;
            }
            g.push_back(s);
        }
    }
    if (earth_flattening)
        for (Src &s : g) s.z = -(float)(R_EARTH * std::log((R_EARTH - (double)s.z) / R_EARTH));
    if (bf_mode) {   // :151-157
        float sum = 0.0f;
        for (const Src &s : g) sum += s.m[0] * s.m[0] + s.m[1] * s.m[1];
        M0 = std::sqrt(sum);
        UC = UC * 1000;
    } else {
        float sum = 0.0f;
        for (const Src &s : g) sum += s.mo;
        M0 = sum;
    }
    fcut = 0.0f;
    for (const Src &s : g) fcut = std::max(fcut, 1 / s.tr);
    fmax = 2 * fcut;
    src_ik.clear(); mo.clear(); m3.clear(); srcprm.clear();
    for (const Src &s : g) {
        const int is = x2i(s.x, xbeg, (float)dx);
        int ks = x2i(s.z, zbeg, (float)dz);
        if (!(ibeg - 2 <= is && is <= iend + 3 && 1 - 2 <= ks && ks <= nz + 3)) continue;   // :177-178
        float sz = s.z;
        if (sdep_fit.size() == 3 && sdep_fit[0] == 'b' && sdep_fit[1] == 'd' && std::isdigit((unsigned char)sdep_fit[2])) {   // :219-227
            sz = bddep[(size_t)(sdep_fit[2] - '0') * nxm + (is - ibeg_m)];
            ks = x2i(sz, zbeg, (float)dz);
        }
        if (!(xbeg <= s.x && s.x <= xend && zbeg <= sz && sz <= zend)) return hfail("source__setup: source outside of the model space (assert, m_source.f90:233-236)");
        src_ik.push_back(is); src_ik.push_back(ks);
        srcprm.push_back(s.t0); srcprm.push_back(s.tr);
        if (bf_mode) {
            mo.push_back(0.0);
            for (int q = 0; q < 2; q++) m3.push_back(field_bytes == 8 ? (double)s.m[q] / (double)M0 : (double)(s.m[q] / M0));   // :244-245
            m3.push_back(0.0);
        } else {
            mo.push_back(field_bytes == 8 ? (double)s.mo / (double)M0 : (double)(s.mo / M0));   // :247
            for (int q = 0; q < 3; q++) m3.push_back((double)s.m[q]);
        }
    }
    return 0;
}

void swpcpsv_host::setup_absorb() {   // m_absorb_p.f90:57-101 / m_absorb_c.f90:28-96
    const float fdx = (float)dx, fdz = (float)dz;
    if (abc_type == "pml") {
        const float hx = (float)(na * dx), hz = (float)(na * dz);
        const int nxo = iend - ibeg + 1;
        gxc.assign(4 * (size_t)nxo, 0.f); gxe.assign(4 * (size_t)nxo, 0.f); gzc.assign(4 * (size_t)nz, 0.f); gze.assign(4 * (size_t)nz, 0.f);
        for (int i = ibeg; i <= iend; i++) {
            damping_profile(xc[i - ibeg_m], hx, xbeg, xend, na, fcut, dt, &gxc[4 * (i - ibeg)]);
            damping_profile(xc[i - ibeg_m] + fdx / 2.0f, hx, xbeg, xend, na, fcut, dt, &gxe[4 * (i - ibeg)]);
        }
        for (int k = 1; k <= nz; k++) {
            damping_profile(zc[k - kbeg_m], hz, zbeg, zend, na, fcut, dt, &gzc[4 * (k - 1)]);
            damping_profile(zc[k - kbeg_m] + fdz / 2.0f, hz, zbeg, zend, na, fcut, dt, &gze[4 * (k - 1)]);
        }
    } else {
        const float alpha = 0.09f, Lx = (float)(na * dx), Lz = (float)(na * dz);
        auto sq = [](float v) { return v * v; };
        cgx_c.assign(nxm, 1.0f); cgx_b.assign(nxm, 1.0f); cgz_c.assign(nzm, 1.0f); cgz_b.assign(nzm, 1.0f);
        auto fill = [&](int lo, int hi, int lo_m, int n, float d, float L, std::vector<float> &gc, std::vector<float> &gb, bool top_open) {
            for (int q = lo; q <= hi; q++) {
                if (q <= na) {
                    if (top_open) continue;
                    gc[q - lo_m] = std::exp(-(alpha * sq(1.0f - (i2x(q, 0.0f, d)) / L)));
                    gb[q - lo_m] = std::exp(-(alpha * sq(1.0f - ((i2x(q, 0.0f, d) + d / 2)) / L)));
                } else if (q >= n - na + 1) {
                    gc[q - lo_m] = std::exp(-(alpha * sq(1.0f - (i2x(q, n * d, -d) + d / 2) / L)));
                    gb[q - lo_m] = std::exp(-(alpha * sq(1.0f - ((i2x(q, n * d, -d))) / L)));
                }
            }
        };
        fill(ibeg, iend, ibeg_m, nx, fdx, Lx, cgx_c, cgx_b, false);
        fill(1, nz, kbeg_m, nz, fdz, Lz, cgz_c, cgz_b, true);
    }
}

int swpcpsv_host::setup_wav(const IniFile &ini) {   // m_wav.f90:53-141, set_stinfo :433-571
    ntdec_w = ini.get_i("ntdec_w", 10);
    ntdec_w_prg = ini.get_i("ntdec_w_prg", 0);
    sw[0] = ini.get_l("sw_wav_v", false); sw[1] = ini.get_l("sw_wav_u", false);
    sw[2] = ini.get_l("sw_wav_stress", false); sw[3] = ini.get_l("sw_wav_strain", false);
    wav_format = ini.get("wav_format", "sac");
    st_format = ini.get("st_format", "xy");
    fn_stloc = ini.get("fn_stloc", "");
    ntdec_r = ini.get_i("ntdec_r", 10);   // m_report.f90:46
    if (!(sw[0] || sw[1] || sw[2] || sw[3])) return 0;
    ntw = (int)std::floor((float)(nt - 1) / (float)ntdec_w + 1.0f);
    std::ifstream is(join_path(base, fn_stloc));
    if (!is) return 0;
    const float fdx = (float)dx, fdz = (float)dz;
    std::string line;
    while (std::getline(is, line)) {
        if (blank_or_comment(line)) continue;
        std::istringstream ls(line);
        float a, b, z;
        std::string name, zsw;
        if (!(ls >> a >> b >> z >> name >> zsw)) continue;
        name = name.substr(0, 8);
        zsw = zsw.substr(0, 3);
        float x, lo, la, dum;
        if (st_format == "xy") { x = a; geomap_c2g(x, 0.0f, clon, clat, phi, lo, la); }
        else if (st_format == "ll") { lo = a; la = b; geomap_g2c(lo, la, clon, clat, phi, x, dum); }
        else return hfail("unknown st_format: " + st_format);
        const int is_ = x2i(x, xbeg, fdx);
        int ks = x2i(z, zbeg, fdz);
        if (!(i2x(1, xbeg, fdx) < x && x < i2x(nx, xbeg, fdx) && 1 < ks && ks < nz)) continue;   // :497-498
        if (!(ibeg <= is_ && is_ <= iend)) continue;
        if (zsw == "dep") ks = x2i(z, zbeg, fdz);
        else if (zsw == "fsb") ks = kfs[is_ - ibeg_m] + 1;
        else if (zsw == "obb") ks = kob[is_ - ibeg_m] + 1;
        else if (zsw == "oba") ks = kob[is_ - ibeg_m] - 1;
        else if (zsw.size() == 3 && zsw[0] == 'b' && zsw[1] == 'd' && std::isdigit((unsigned char)zsw[2]))
            ks = x2i(bddep[(size_t)(zsw[2] - '0') * nxm + (is_ - ibeg_m)], zbeg, fdz);
        else ks = x2i(z, zbeg, fdz);
        if (ks > nz) ks = nz - 1;
        if (ks < 1) ks = 1 + 1;
        st_ik.push_back(is_); st_ik.push_back(ks);
        xst.push_back(x); zst.push_back(z); stlo.push_back(lo); stla.push_back(la); stnm.push_back(name);
    }
    return 0;
}

int swpcpsv_host::setup(const IniFile &ini, int nm_, int myid_, int npx, int nt_o) {
    if (nm_ < 0 || nm_ > 3) return hfail("nm must be 0..3");
    nm = nm_;
    myid = myid_;
    if (setup_global(ini, npx, nt_o)) return 1;   // main.f90:64-78 order
    if (setup_geometry()) return 1;
    if (setup_medium(ini)) return 1;
    if (setup_source(ini)) return 1;
    setup_absorb();
    setup_snap(ini);
    if (setup_wav(ini)) return 1;
    return 0;
}

// ============================================================================================================
// snapshots: snap__setup m_snap.f90:78-164 (files are created when a device is attached: swpcpsv_host_snap_open)
static const char *PSV_SNAP_TYPE[3] = {"ps", "v2", "u2"}, *PSV_SNAP_TAG[3] = {"ps", "v", "u"};
static const char *PSV_SNAP_VAR[3][2] = {{"divergence", "rotation"}, {"Vx", "Vz"}, {"Ux", "Uz"}}, *PSV_SNAP_UNIT[3] = {"1/s", "m/s", "m"};

void swpcpsv_host::setup_snap(const IniFile &ini) {
    Snap &S = snap;
    S.sw[0] = ini.get_l("xz_ps%sw", false); S.sw[1] = ini.get_l("xz_v%sw", false); S.sw[2] = ini.get_l("xz_u%sw", false);
    S.idec = ini.get_i("idec", 1); S.kdec = ini.get_i("kdec", 1); S.ntdec_s = ini.get_i("ntdec_s", 10);
    S.native = ini.get("snp_format", "native") == "native";
    S.nxs = (nx + (S.idec / 2)) / S.idec;
    S.nzs = (nz + (S.kdec / 2)) / S.kdec;
    S.xsnp.resize((size_t)S.nxs); S.zsnp.resize((size_t)S.nzs);
    for (int i = 1; i <= S.nxs; i++) S.xsnp[(size_t)i - 1] = i2x(i * S.idec - (S.idec / 2), xbeg, (float)dx);
    for (int k = 1; k <= S.nzs; k++) S.zsnp[(size_t)k - 1] = i2x(k * S.kdec - (S.kdec / 2), zbeg, (float)dz);
    S.is0 = (int)std::ceil((float)(ibeg + S.idec / 2) / (float)S.idec);
    S.is1 = (int)std::floor((float)(iend + S.idec / 2) / (float)S.idec);
    S.ks0 = (int)std::ceil((float)(1 + S.kdec / 2) / (float)S.kdec);
    S.ks1 = (int)std::floor((float)(nz + S.kdec / 2) / (float)S.kdec);
    for (int q = 0; q < 3; q++) S.ionode[q] = (q + 1) % nproc_x;   // :117-119
}

// newfile_xz / newfile_xz_nc (:167-269) + write_snp_header (:272-315) / write_nc_header (:317-393)
int swpcpsv_host::snap_open(const std::string &dir) {
    Snap &S = snap;
    if (!(S.sw[0] || S.sw[1] || S.sw[2]) || S.opened) return 0;
    if (!dev) return hfail("swpcpsv_host_snap_open: no device attached");
    make_dirs(dir);
    const size_t np = (size_t)S.nxs * S.nzs;
    for (int q = 0; q < 3; q++) {
        if (!S.sw[q]) continue;
        std::vector<std::vector<float>> med(3, std::vector<float>(np, 0.0f));
        for (int i = S.is0; i <= S.is1; i++)
            for (int k = S.ks0; k <= S.ks1; k++) {
                const int ii = i * S.idec - S.idec / 2, kk = k * S.kdec - S.kdec / 2;
                const size_t o = (size_t)(i - 1) + (size_t)S.nxs * (size_t)(k - 1), n = i2(kk, ii);
                med[0][o] = rho[n]; med[1][o] = lam[n]; med[2][o] = mu[n];
            }
        for (int m = 0; m < 3; m++)
            if (swpcpsv_reduce_sum(dev, med[(size_t)m].data(), (int64_t)np, S.ionode[q])) return hfail(std::string("device: ") + swpcpsv_last_error());
        if (myid != S.ionode[q]) continue;
        const std::string fname = dir + "/" + title + ".psv.xz." + PSV_SNAP_TAG[q] + (S.native ? ".snp" : ".nc");
        if (S.native) {
            FILE *f = S.snp[q] = std::fopen(fname.c_str(), "wb");
            if (!f) return hfail("cannot create " + fname);
            std::string ttl = title; ttl.resize(80, ' ');
            auto wi = [&](int32_t v) { std::fwrite(&v, 4, 1, f); };
            auto wf = [&](float v) { std::fwrite(&v, 4, 1, f); };
            std::fwrite("STREAMIO", 1, 8, f); std::fwrite("SWPC_PSV", 1, 8, f); wi(6);
            std::fwrite(ttl.data(), 1, 80, f); wi(exedate);
            std::fwrite("xz", 1, 2, f); std::fwrite(PSV_SNAP_TYPE[q], 1, 2, f);
            wi(S.nxs); wi(S.nzs); wf(S.xsnp[0]); wf(S.zsnp[0]);
            wf(S.nxs > 1 ? S.xsnp[1] - S.xsnp[0] : 0.0f); wf(S.nzs > 1 ? S.zsnp[1] - S.zsnp[0] : 0.0f);
            wf(dt * (float)S.ntdec_s); wi(na / S.idec); wi(na / S.kdec); wi(3); wi(2);
            wf(clon); wf(clat); wf(phi); wf(0.0f); wf(0.0f); wf(0.0f);
            for (int m = 0; m < 3; m++) std::fwrite(med[(size_t)m].data(), 4, np, f);
            continue;
        }
        NcFile *nc = S.nc[q] = new NcFile();
        nc->dims = {{"x", S.nxs}, {"z", S.nzs}, {"t", 0}};
        auto mm = [](const std::vector<float> &v) { float lo = v[0], hi = v[0]; for (float f : v) { lo = std::min(lo, f); hi = std::max(hi, f); } return std::make_pair(lo, hi); };
        NcVar vx; vx.name = "x"; vx.dimids = {0}; vx.data = S.xsnp;
        NcVar vz; vz.name = "z"; vz.dimids = {1}; vz.data = S.zsnp;
        NcVar vt; vt.name = "t"; vt.dimids = {2}; vt.rec = true;
        vx.atts = {att_text("long_name", "x"), att_text("units", "km"), att_floats("actual_range", {S.xsnp.front(), S.xsnp.back()})};
        vz.atts = {att_text("long_name", "z"), att_text("units", "km"), att_floats("actual_range", {S.zsnp.front(), S.zsnp.back()})};
        vt.atts = {att_text("long_name", "t"), att_text("units", "s")};
        nc->vars = {vx, vz, vt};
        static const char *mname[3] = {"rho", "lambda", "mu"}, *munit[3] = {"10^3 kg/cm^3", "10^9 Pa", "10^9 Pa"};
        for (int m = 0; m < 3; m++) {
            NcVar v; v.name = mname[m]; v.dimids = {1, 0}; v.data = med[(size_t)m];
            const auto r = mm(med[(size_t)m]);
            v.atts = {att_text("long_name", mname[m]), att_text("units", munit[m]), att_floats("actual_range", {r.first, r.second})};
            nc->vars.push_back(v);
        }
        for (int v = 0; v < 2; v++) {
            NcVar sv; sv.name = PSV_SNAP_VAR[q][v]; sv.dimids = {2, 1, 0}; sv.rec = true;
            sv.atts = {att_text("long_name", PSV_SNAP_VAR[q][v]), att_text("units", PSV_SNAP_UNIT[q]), att_floats("actual_range", {0.0f, 0.0f})};
            nc->vars.push_back(sv);
        }
        nc->gatts = {att_text("generated_by", "SWPC"), att_text("title", title), att_int("exedate", exedate), att_int("hdrver", 6), att_text("codetype", "SWPC_PSV"),
                     att_int("ns1", S.nxs), att_int("ns2", S.nzs), att_floats("beg1", {S.xsnp.front()}), att_floats("beg2", {S.zsnp.front()}),
                     att_int("na1", na / S.idec), att_int("na2", na / S.kdec), att_floats("ds1", {S.idec * (float)dx}), att_floats("ds2", {S.kdec * (float)dz}),
                     att_int("nmed", 3), att_int("nsnp", 2), att_text("coordinate", "xz"), att_text("datatype", PSV_SNAP_TYPE[q]),
                     att_floats("dt", {dt * S.ntdec_s}), att_floats("evlo", {evlo}), att_floats("evla", {evla}), att_floats("evdp", {evdp}),
                     att_floats("evx", {sx0}), att_floats("evy", {sy0}), att_floats("clon", {clon}), att_floats("clat", {clat}), att_floats("phi", {phi})};
        if (!nc->create(fname)) return hfail("cannot create " + fname);
    }
    swpcpsv_snap_cfg c{};
    c.idec = S.idec; c.kdec = S.kdec; c.ntdec_s = S.ntdec_s; c.nxs = S.nxs; c.nzs = S.nzs; c.is0 = S.is0; c.is1 = S.is1; c.ks0 = S.ks0; c.ks1 = S.ks1;
    c.sw_ps = S.sw[0]; c.sw_v = S.sw[1]; c.sw_u = S.sw[2]; c.M0 = M0; c.UC = UC;
    if (swpcpsv_snap_setup(dev, &c)) return hfail(std::string("device: ") + swpcpsv_last_error());
    S.opened = true;
    return 0;
}

// the host part of snap__write(it) at an output step (the device part runs inside swpcpsv_step): reduce + record write
// (wbuf_nc :652-681; written at once instead of one cycle later -- same record index it0/ntdec_s+1, same time, same data)
int swpcpsv_host::snap_write(int it) {
    Snap &S = snap;
    if (!S.opened || !(S.ntdec_s > 0 && (it - 1) % S.ntdec_s == 0)) return 0;
    const size_t np = (size_t)S.nxs * S.nzs;
    for (int q = 0; q < 3; q++) {
        if (!S.sw[q]) continue;
        S.tmp.assign(2 * np, 0.0f);
        if (swpcpsv_snap_fetch(dev, q, S.ionode[q], S.tmp.data())) return hfail(std::string("device: ") + swpcpsv_last_error());
        if (S.snp[q]) { std::fwrite(S.tmp.data(), 4, 2 * np, S.snp[q]); std::fflush(S.snp[q]); continue; }
        NcFile *nc = S.nc[q];
        if (!nc) continue;
        const int rec = it / S.ntdec_s;
        const float tval = it * dt;
        nc->put_record(nc->var_index("t"), rec, &tval, 1);
        for (int v = 0; v < 2; v++) {
            const float *d = S.tmp.data() + np * (size_t)v;
            const int vi = nc->var_index(PSV_SNAP_VAR[q][v]);
            nc->put_record(vi, rec, d, np);
            for (size_t n = 0; n < np; n++) { S.vmax[q][v] = std::max(S.vmax[q][v], d[n]); S.vmin[q][v] = std::min(S.vmin[q][v], d[n]); }
            *nc->find_att(nc->vars[(size_t)vi], "actual_range") = att_floats("actual_range", {S.vmin[q][v], S.vmax[q][v]});
        }
        nc->flush_header();
    }
    return 0;
}

int swpcpsv_host::snap_close() {   // snap__closefiles :683-719
    Snap &S = snap;
    for (int q = 0; q < 3; q++) {
        if (S.nc[q]) { S.nc[q]->flush_header(); delete S.nc[q]; S.nc[q] = nullptr; }
        if (S.snp[q]) { std::fclose(S.snp[q]); S.snp[q] = nullptr; }
    }
    S.opened = false;
    return 0;
}

// ============================================================================================================
template <typename T>
static int put(const std::vector<T> &v, void *out, int64_t cap, int64_t *n) {
    if (n) *n = (int64_t)v.size();
    if (out && cap > 0) std::memcpy(out, v.data(), sizeof(T) * (size_t)std::min<int64_t>(cap, (int64_t)v.size()));
    return 0;
}

extern "C" {

static int host_create(IniFile &ini, const char *base_dir, int nm, int myid, int npx, int nt, int fb, swpcpsv_host **out) {
    if (!out) return hfail("null output pointer");
    *out = nullptr;
    if (fb != 8 && fb != 4) return hfail("field_bytes must be 8 or 4");
    swpcpsv_host *h = new swpcpsv_host();
    h->base = base_dir ? base_dir : "";
    h->field_bytes = fb;
    ini.strict = ini.get_l("strict_mode", false);   // main.f90:61-62
    if (h->setup(ini, nm, myid, npx, nt)) { delete h; return 1; }
    *out = h;
    return 0;
}
int swpcpsv_host_create(const char *inf_path, const char *base_dir, int32_t nm, int32_t myid, int32_t npx, int32_t nt, int32_t fb, swpcpsv_host **out) {
    IniFile ini;
    if (!inf_path || !IniFile::from_file(inf_path, ini)) return hfail(std::string("cannot open parameter file ") + (inf_path ? inf_path : "(null)"));
    return host_create(ini, base_dir, nm, myid, npx, nt, fb, out);
}
int swpcpsv_host_create_from_text(const char *text, const char *base_dir, int32_t nm, int32_t myid, int32_t npx, int32_t nt, int32_t fb, swpcpsv_host **out) {
    if (!text) return hfail("null text");
    IniFile ini = IniFile::from_text(text);
    return host_create(ini, base_dir, nm, myid, npx, nt, fb, out);
}
int swpcpsv_host_destroy(swpcpsv_host *h) {
    if (!h) return 0;
    if (h->dev) swpcpsv_destroy(h->dev);
    delete h;
    return 0;
}
const char *swpcpsv_host_last_error(void) { return g_herr.c_str(); }

int swpcpsv_host_get_int(swpcpsv_host *h, const char *name, int32_t *v) {
    if (!h || !name || !v) return hfail("null argument");
    const std::string n = name;
#define GI(f) if (n == #f) { *v = h->f; return 0; }
    GI(nx) GI(nz) GI(nt) GI(na) GI(nm) GI(nproc_x) GI(myid) GI(ibeg) GI(iend) GI(nxp) GI(ibeg_k) GI(iend_k) GI(kend_k) GI(ntw) GI(ntdec_w) GI(ntdec_r)
    GI(exedate) GI(nxm) GI(nzm)
#undef GI
    if (n == "nsrc") { *v = (int)(h->src_ik.size() / 2); return 0; }
    if (n == "nst") { *v = (int)(h->st_ik.size() / 2); return 0; }
    if (n == "bf_mode") { *v = h->bf_mode ? 1 : 0; return 0; }
    return hfail("unknown int " + n);
}
int swpcpsv_host_get_double(swpcpsv_host *h, const char *name, double *v) {
    if (!h || !name || !v) return hfail("null argument");
    const std::string n = name;
#define GD(f) if (n == #f) { *v = (double)h->f; return 0; }
    GD(dx) GD(dz) GD(dt) GD(xbeg) GD(zbeg) GD(tbeg) GD(vmin) GD(vmax) GD(vmin_local) GD(vmax_local) GD(fmax) GD(fcut) GD(M0) GD(UC) GD(zeta) GD(d2)
    GD(loop_seconds) GD(evlo) GD(evla) GD(evdp)
#undef GD
    if (n == "c") { *v = (double)(h->dt / stable_dt((float)h->dx, 1e10f, (float)h->dz, h->vmax)); return 0; }   // m_report.f90:70
    if (n == "r") { *v = (double)((h->vmin / h->fmax) / std::max(std::max((float)h->dx, -1.0f), (float)h->dz)); return 0; }   // m_report.f90:71
    return hfail("unknown double " + n);
}
int swpcpsv_host_set_minmax(swpcpsv_host *h, float vmin, float vmax) {
    if (!h) return hfail("null handle");
    h->vmin = vmin; h->vmax = vmax;
    h->apply_stabilize();
    return 0;
}
int swpcpsv_host_set_exedate(swpcpsv_host *h, int32_t exedate, int32_t tz) {
    if (!h) return hfail("null handle");
    h->exedate = exedate; h->tz_minutes = tz;
    return 0;
}
int swpcpsv_host_get_array(swpcpsv_host *h, const char *name, void *out, int64_t cap, int64_t *n) {
    if (!h || !name) return hfail("null argument");
    const std::string s = name;
#define GA(f) if (s == #f) return put(h->f, out, cap, n);
    GA(rho) GA(lam) GA(mu) GA(taup) GA(taus) GA(gxc) GA(gxe) GA(gzc) GA(gze) GA(srcprm) GA(kfs) GA(kob) GA(kfs_top) GA(kfs_bot) GA(kob_top) GA(kob_bot)
    GA(kbeg_a) GA(src_ik) GA(st_ik) GA(mo) GA(m3)
#undef GA
    if (s == "gx_c") return put(h->cgx_c, out, cap, n);
    if (s == "gx_b") return put(h->cgx_b, out, cap, n);
    if (s == "gz_c") return put(h->cgz_c, out, cap, n);
    if (s == "gz_b") return put(h->cgz_b, out, cap, n);
    if (s == "ts") return put(std::vector<float>(h->ts, h->ts + h->nm), out, cap, n);
    for (int p = 0; p < 4; p++)
        if (s == std::string("wav") + char('0' + p)) return put(h->wav_all[p], out, cap, n);
    static const char *init[5] = {"init_Vx", "init_Vz", "init_Sxx", "init_Szz", "init_Sxz"};   // plane-wave initial condition (pw_mode)
    for (int q = 0; q < 5; q++)
        if (s == init[q]) return put(h->pw_init[q], out, cap, n);
    return hfail("unknown array " + s);
}
int swpcpsv_host_station_name(swpcpsv_host *h, int32_t i, char *buf9) {
    if (!h || !buf9 || i < 0 || (size_t)i >= h->stnm.size()) return hfail("bad station index");
    std::memset(buf9, 0, 9);
    std::strncpy(buf9, h->stnm[(size_t)i].c_str(), 8);
    return 0;
}

int swpcpsv_host_attach_device(swpcpsv_host *h, int32_t device) {   // main.f90:80-93
    if (!h) return hfail("null handle");
    if (h->dev) { swpcpsv_destroy(h->dev); h->dev = nullptr; }
    h->apply_stabilize();
    swpcpsv_grid g{};
    g.nx = h->nx; g.nz = h->nz; g.nproc_x = h->nproc_x; g.myid = h->myid; g.ibeg = h->ibeg; g.iend = h->iend; g.ipad = h->ipad; g.kpad = h->kpad;
    g.ibeg_k = h->ibeg_k; g.iend_k = h->iend_k; g.kend_k = h->kend_k; g.na = h->na; g.nm = h->nm;
    g.abc_type = h->abc_type == "pml" ? SWPCPSV_ABC_PML : SWPCPSV_ABC_CERJAN;
    g.field_bytes = h->field_bytes; g.device = device; g.dx = h->dx; g.dz = h->dz; g.dt = h->dt;
#define DV(call) if (call) return hfail(std::string("device: ") + swpcpsv_last_error());
    DV(swpcpsv_create(&g, h->ts, &h->dev));
    DV(swpcpsv_upload_medium(h->dev, h->rho.data(), h->lam.data(), h->mu.data(), h->taup.data(), h->taus.data(), h->kfs.data(), h->kob.data(),
                             h->kfs_top.data(), h->kfs_bot.data(), h->kob_top.data(), h->kob_bot.data(), h->kbeg_a.data()));
    if (g.abc_type == SWPCPSV_ABC_PML) { DV(swpcpsv_setup_pml(h->dev, h->gxc.data(), h->gxe.data(), h->gzc.data(), h->gze.data())); }
    else { DV(swpcpsv_setup_cerjan(h->dev, h->cgx_c.data(), h->cgx_b.data(), h->cgz_c.data(), h->cgz_b.data())); }
    const int nsrc = (int)(h->src_ik.size() / 2);
    if (nsrc > 0) {
        std::vector<int> a(nsrc), b(nsrc);
        std::vector<double> m[3];
        for (int q = 0; q < 3; q++) m[q].resize(nsrc);
        for (int i = 0; i < nsrc; i++) {
            a[i] = h->src_ik[2 * i]; b[i] = h->src_ik[2 * i + 1];
            for (int q = 0; q < 3; q++) m[q][i] = h->m3[3 * (size_t)i + q];
        }
        DV(swpcpsv_set_sources(h->dev, nsrc, a.data(), b.data(), h->mo.data(), m[0].data(), m[1].data(), m[2].data(), h->srcprm.data(), h->stftype.c_str(),
                               h->bf_mode ? 1 : 0, h->tbeg));
    }
    if (h->pw_mode && !h->pw_init[0].empty()) {   // `!$acc enter data copyin(Vx..Sxz)` with the plane-wave initial condition
        const void *f[5];
        std::vector<float> f32[5];
        for (int q = 0; q < 5; q++) {
            if (h->field_bytes == 4) { f32[q].assign(h->pw_init[q].begin(), h->pw_init[q].end()); f[q] = f32[q].data(); }
            else f[q] = h->pw_init[q].data();
        }
        DV(swpcpsv_upload_fields(h->dev, f[0], f[1], f[2], f[3], f[4]));
        DV(swpcpsv_set_option(h->dev, "pw_mode", 1));
    }
    const int nst = (int)(h->st_ik.size() / 2);
    if (nst > 0 && (h->sw[0] || h->sw[1] || h->sw[2] || h->sw[3])) {
        std::vector<int> a(nst), b(nst);
        for (int i = 0; i < nst; i++) { a[i] = h->st_ik[2 * i]; b[i] = h->st_ik[2 * i + 1]; }
        DV(swpcpsv_set_stations(h->dev, nst, a.data(), b.data(), h->ntdec_w, h->ntw, h->M0, h->UC, h->sw[0], h->sw[1], h->sw[2], h->sw[3]));
    }
#undef DV
    return 0;
}
swpcpsv_handle *swpcpsv_host_handle(swpcpsv_host *h) { return h ? h->dev : nullptr; }

int swpcpsv_host_banner(swpcpsv_host *h) {   // m_report.f90:52-100
    if (!h) return hfail("null handle");
    double c, r;
    swpcpsv_host_get_double(h, "c", &c);
    swpcpsv_host_get_double(h, "r", &r);
    std::fprintf(stderr, "\n ------------------------------------------------------------------------------\n");
    std::fprintf(stderr, "  SWPC_PSV (swpcpsv_b200, B200-native time loop)%s\n", h->benchmark_mode ? " (benchmark mode)   " : (h->bf_mode ? " (body force mode)  " : ""));
    std::fprintf(stderr, " ------------------------------------------------------------------------------\n\n");
    std::fprintf(stderr, "  Grid Size               : %8d x %6d\n", h->nx, h->nz);
    std::fprintf(stderr, "  MPI Partitioning        : %15d\n", h->nproc_x);
    std::fprintf(stderr, "  Stability  Condition c  : %15.3f  (c<1)\n", c);
    std::fprintf(stderr, "  Wavelength Condition r  : %15.3f  (r>5-10)\n", r);
    std::fprintf(stderr, "  Minimum velocity        : %15.3f  [km/s]\n", (double)h->vmin);
    std::fprintf(stderr, "  Maximum velocity        : %15.3f  [km/s]\n", (double)h->vmax);
    std::fprintf(stderr, "  Maximum frequency       : %15.3f  [Hz]\n\n", (double)h->fmax);
    std::fprintf(stderr, " ------------------------------------------------------------------------------\n\n");
    if (c > 1.0) return hfail("stability condition is violated (assert(c <= 1.0), m_report.f90:96-99)");
    return 0;
}

int swpcpsv_host_run(swpcpsv_host *h, int32_t it0, int32_t it1, int32_t verbose, float *vm, int32_t nvm, int32_t *nrec) {
    if (!h || !h->dev) return hfail("swpcpsv_host_run: no device attached");
    int rec = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = it0; it <= it1; it++) {
        if (h->ntdec_r > 0 && it % h->ntdec_r == 0) {   // report__progress m_report.f90:122-172
            float v[2];
            if (swpcpsv_vmax_global(h->dev, v)) return hfail(std::string("device: ") + swpcpsv_last_error());
            for (int q = 0; q < 2; q++) v[q] = v[q] * h->UC * h->M0;
            if (vm && rec < nvm) { vm[2 * rec] = v[0]; vm[2 * rec + 1] = v[1]; }
            rec++;
            if (verbose && h->myid == 0) {
                const double tt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                double etas = (double)(h->nt - it) / (double)std::max(1, it - it0 + 1) * tt;
                const int eh = (int)(etas / 3600); etas -= eh * 3600.0;
                const int em = (int)(etas / 60); etas -= em * 60.0;
                std::fprintf(stderr, "  it=%07d,%6.3f s/loop, eta %03d:%02d:%02d, (%9.2E %9.2E )\n", it, tt / std::max(1, it - it0 + 1), eh, em, (int)etas,
                             (double)v[0], (double)v[1]);
            }
        }
        // snap__write(it) at the top of the iteration (main.f90:99): slices on the device, then -- on output steps -- the
        // reduction onto the I/O rank and the record; the rest of the iteration follows (swpcpsv_step minus its snap_step)
        if (h->snap.opened) {
            if (swpcpsv_snap_step(h->dev, it) || swpcpsv_wav_store(h->dev, it)) return hfail(std::string("device: ") + swpcpsv_last_error());
            if (h->snap_write(it)) return 1;
            if (swpcpsv_update_stress(h->dev) || swpcpsv_stressglut(h->dev, it) || swpcpsv_comm_stress(h->dev) || swpcpsv_update_vel(h->dev, it) ||
                swpcpsv_comm_vel(h->dev))
                return hfail(std::string("device: ") + swpcpsv_last_error());
        } else if (swpcpsv_step(h->dev, it)) return hfail(std::string("device: ") + swpcpsv_last_error());
        if (h->ntdec_w_prg > 0 && (it - 1) % h->ntdec_w_prg == 0) {   // wav__store's tail, m_wav.f90:309-311
            int32_t nf = 0;
            if (swpcpsv_host_write_wav(h, nullptr, &nf)) return 1;
        }
    }
    if (swpcpsv_sync(h->dev)) return hfail(std::string("device: ") + swpcpsv_last_error());
    h->loop_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (nrec) *nrec = rec;
    return 0;
}

// wav__write m_wav.f90:308-420; SAC header values: initialize_sac_header :625-681, set_sac_header :573-623, sac__whdr
// src/shared/m_sac.f90:314-449
static void put_chars(char *dst, const std::string &s, int n) {
    for (int q = 0; q < n; q++) dst[q] = q < (int)s.size() ? s[q] : ' ';
}
int swpcpsv_host_snap_open(swpcpsv_host *h, const char *odir) {
    if (!h) return hfail("null handle");
    return h->snap_open(odir ? odir : h->odir.c_str());
}
int swpcpsv_host_snap_close(swpcpsv_host *h) {
    if (!h) return hfail("null handle");
    return h->snap_close();
}
int swpcpsv_host_write_wav(swpcpsv_host *h, const char *odir, int32_t *nfiles) {
    if (!h) return hfail("null handle");
    if (nfiles) *nfiles = 0;
    const int nst = (int)(h->st_ik.size() / 2);
    if (!(h->sw[0] || h->sw[1] || h->sw[2] || h->sw[3]) || nst == 0 || h->ntw <= 0) return 0;
    if (!h->dev) return hfail("swpcpsv_host_write_wav: no device attached");
    const std::string dir = std::string(odir ? odir : h->odir.c_str()) + "/wav";
    make_dirs(dir);
    static const char *cmpnm[4][3] = {{"Vx", "Vz", ""}, {"Ux", "Uz", ""}, {"Sxx", "Szz", "Sxz"}, {"Exx", "Ezz", "Exz"}};
    const time_t tt = (time_t)h->exedate + (time_t)h->tz_minutes * 60;
    struct tm g;
    gmtime_r(&tt, &g);
    for (int prod = 0; prod < 4; prod++) {
        if (!h->sw[prod]) continue;
        h->wav_all[prod].assign((size_t)h->ntw * (prod < 2 ? 2 : 3) * nst, 0.0f);
        if (swpcpsv_get_wav(h->dev, prod, h->wav_all[prod].data())) return hfail(std::string("device: ") + swpcpsv_last_error());
    }
    auto sac_record = [&](int s, int prod, int c, std::vector<unsigned char> &out) {
        const int ncmp = prod < 2 ? 2 : 3;
        float f[70];
        int32_t iv[35], lv[5];
        char a[192];
        std::fill(f, f + 70, -12345.0f);
        std::fill(iv, iv + 35, -12345);
        std::fill(lv, lv + 5, 0);
        for (int q = 0; q < 24; q++) put_chars(a + 8 * q, "-12345", 8);
        put_chars(a + 8, "-12345", 16);
        const double delta = (double)(h->ntdec_w * h->dt);
        f[0] = (float)((int)(delta * 1e7)) / 1e7f;
        f[5] = h->tbeg; f[7] = h->otim;
        f[31] = h->stla[s]; f[32] = h->stlo[s]; f[34] = h->zst[s] * 1000;
        f[35] = h->evla; f[36] = h->evlo; f[38] = h->evdp; f[39] = moment_magnitude(h->M0);
        if (h->bf_mode) { f[40] = h->fx0; f[42] = h->fz0; }
        else { f[40] = h->mxx0; f[42] = h->mzz0; f[44] = h->mxz0; }
        f[46] = h->clon; f[47] = h->clat; f[48] = h->phi;
        const float dd = h->sx0 - h->xst[s];
        f[50] = std::sqrt(dd * dd);
        f[51] = rad2deg_s(std::atan2(0.0f, h->xst[s] - h->sx0));
        f[52] = rad2deg_s(std::atan2(0.0f, h->sx0 - h->xst[s]));
        if (prod < 2) { f[58] = 90.0f; f[57] = c == 0 ? 0.0f + h->phi : 0.0f; }
        iv[0] = g.tm_year + 1900; iv[1] = g.tm_yday + 1; iv[2] = g.tm_hour; iv[3] = g.tm_min; iv[4] = g.tm_sec; iv[5] = 0;
        iv[6] = 6; iv[9] = h->ntw; iv[15] = 1; iv[16] = prod == 0 ? 7 : (prod == 1 ? 6 : 5);
        lv[0] = 1; lv[2] = 1;
        put_chars(a, h->stnm[s], 8);
        std::string t = h->title;
        t.erase(0, t.find_first_not_of(' ') == std::string::npos ? t.size() : t.find_first_not_of(' '));
        put_chars(a + 8, t.substr(0, 16), 16);
        put_chars(a + 160, cmpnm[prod][c], 8);
        out.resize(632 + 4 * (size_t)h->ntw);
        std::memcpy(out.data(), f, 280); std::memcpy(out.data() + 280, iv, 140); std::memcpy(out.data() + 420, lv, 20); std::memcpy(out.data() + 440, a, 192);
        std::memcpy(out.data() + 632, h->wav_all[prod].data() + (size_t)h->ntw * ncmp * s + (size_t)h->ntw * c, 4 * (size_t)h->ntw);
    };
    std::vector<WavStation> sts((size_t)nst);
    for (int s = 0; s < nst; s++) {
        sts[(size_t)s].stnm = h->stnm[s];
        for (int prod = 0; prod < 4; prod++) {
            if (!h->sw[prod]) continue;
            for (int c = 0; c < (prod < 2 ? 2 : 3); c++) {
                WavTrace t; t.prod = prod; t.cmp = cmpnm[prod][c];
                sac_record(s, prod, c, t.rec);
                sts[(size_t)s].tr.push_back(std::move(t));
            }
        }
    }
    std::string err;
    const int count = write_wav_files(h->wav_format, dir, h->title, "psv", false, h->myid, h->exedate, h->ntw, sts, err);
    if (count < 0) return hfail(err);
    if (nfiles) *nfiles = count;
    return 0;
}

}   // extern "C"
