// driver.cpp -- host-side mirror of swpc_3d's setup chain and driver loop (include/swpc3d_host.h).
//
// Setup-only CPU code: it produces, for ONE rank, exactly the arrays the reference's setup modules hand to
// the time loop (the loop itself runs only on the GPU through swpc3d_b200.h).  Every routine cites the
// reference lines whose result it must reproduce (paths under /root/reference, OpenSWPC 25.05.2); kinds follow
// the Fortran declarations (default real = float, PI = real(DP), src/shared/m_std.f90:14).
#include "../../../include/swpc3d_host.h"

#include "common.hpp"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include "models.hpp"

// ============================================================================================================
struct swpc3d_host {
    // ---- global parameters (m_global.f90:124-216)
    bool benchmark_mode = false;
    std::string title, odir, abc_type, vmodel_type, stftype, stf_format, sdep_fit, wav_format, st_format, fn_stf, fn_stloc, base;
    int nproc_x = 1, nproc_y = 2, nx = 256, ny = 256, nz = 256, nt = 1000, ipad = 0, jpad = 0, kpad = 0, na = 20, nm = 3;
    double dx = 0.5, dy = 0.5, dz = 0.5;
    float dt = 0.01f, xbeg = 0, ybeg = 0, zbeg = 0, tbeg = 0, xend = 0, yend = 0, zend = 0, clon = 0, clat = 0, phi = 0;
    float fq_min = 0.05f, fq_max = 5.0f, fq_ref = 1.0f, vcut = 0.0f;
    bool pw_mode = false, green_mode = false, bf_mode = false, earth_flattening = false;
    int ntdec_w = 10, ntdec_r = 10, ntw = 0;
    // m_pwatch replacement (main.f90:58, :148-154): per-phase device time from the library's CUDA-event stopwatches
    bool stopwatch_mode = true;
    double tim_stress = 0, tim_vel = 0, tim_halo = 0;   // seconds
    int harvest_timers();
    int ntdec_w_prg = 0;   // m_wav.f90:74, :619-621: waveform files rewritten every ntdec_w_prg steps while the run goes on
    bool sw_wav_v = false, sw_wav_u = false, sw_wav_stress = false, sw_wav_strain = false;
    float vmin = 0, vmax = 0, vmin_local = 0, vmax_local = 0, fmax = 0, fcut = 0, M0 = 0, UC = 1e-15f, zeta = 0, d2 = 0;
    float ts[8] = {}, c1[8] = {}, c2[8] = {}, d1[8] = {};
    float evlo = 0, evla = 0, evdp = 0, m0ij[6] = {}, f0[3] = {}, otim = 0, sx0 = 0, sy0 = 0;
    int exedate = 0, tz_minutes = 0;
    int field_bytes = 8;
    // ---- this rank (m_global.f90:219-388)
    int myid = 0, idx = 0, idy = 0, nxp = 0, nyp = 0, ibeg = 0, iend = 0, jbeg = 0, jend = 0;
    int ibeg_m = 0, iend_m = 0, jbeg_m = 0, jend_m = 0, kbeg_m = 0, kend_m = 0, nzm = 0, nxm = 0, nym = 0;
    int ibeg_k = 0, iend_k = 0, jbeg_k = 0, jend_k = 0, kbeg_k = 0, kend_k = 0;
    std::vector<float> xc, yc, zc;
    std::vector<float> rho, lam, mu, taup, taus, bddep;
    std::vector<int> kfs, kob, kfs_top, kfs_bot, kob_top, kob_bot, kbeg_a;
    std::vector<float> gxc, gxe, gyc, gye, gzc, gze;       // PML (4,n)
    std::vector<float> cgx_c, cgx_b, cgy_c, cgy_b, cgz_c, cgz_b;   // Cerjan
    // sources / stations
    std::vector<int> src_ijk, st_ijk;
    std::vector<double> mo, mij;
    std::vector<float> srcprm, xst, yst, zst, stlo, stla;
    std::vector<std::string> stnm;
    // device
    swpc3d_handle *dev = nullptr;
    struct SnapHost *snap = nullptr;       // snapshot products (m_snap.f90)
    int setup_snap(const IniFile &ini);
    int setup_planewave(const IniFile &ini);
    std::vector<double> pw_init[9];        // plane-wave initial fields Vx Vy Vz Sxx Syy Szz Syz Sxz Sxy over the memory box
    int snap_open_files(const std::string &dir);
    int snap_write(int it);
    int snap_close();
    std::vector<float> wav;
    std::vector<float> wav_all[4];
    double loop_seconds = 0;

    size_t i3(int k, int i, int j) const { return (size_t)(k - kbeg_m) + (size_t)nzm * ((size_t)(i - ibeg_m) + (size_t)nxm * (size_t)(j - jbeg_m)); }
    size_t i2(int i, int j) const { return (size_t)(i - ibeg_m) + (size_t)nxm * (size_t)(j - jbeg_m); }

    int setup(const IniFile &ini, int nm_, int myid_, int npx, int npy, int nt_o);
    int setup_global(const IniFile &ini, int npx, int npy, int nt_o);
    int setup_geometry();
    int setup_medium(const IniFile &ini);
    int finish_medium_3d(const IniFile &ini);
    int finish_surface(const IniFile &ini);
    bool stabilize_pending = false;
    void apply_stabilize();
    void setup_kernel();
    int setup_source(const IniFile &ini);
    int setup_absorb();
    int setup_wav(const IniFile &ini);
    // Green's-function mode (m_green.f90)
    struct Green {
        std::string stnm, fn_glst, fmt, stftype, wav_format;
        char cmp = ' ';
        float trise = 1.0f, maxdist = 1e30f, f1[3] = {0, 0, 0};
        bool bforce = false, have_src = false, finalized = false;
        int src_ijk[3] = {0, 0, 0}, ntdec_w = 10, ntw = 0, ncmp = 6;
        float src_xyz[3] = {0, 0, 0}, evlo0 = 0, evla0 = 0;
        std::vector<int> ijk, gid;                       // owned grid points (3 per point) and their ids
        std::vector<float> xg, yg, zg, lon, lat;
        std::vector<float> gf;                           // (ntw, ncmp*ng) after swpc3d_host_write_green
    } green;
    int setup_green(const IniFile &ini);
    int green_finalize();
};

int swpc3d_host::setup_global(const IniFile &ini, int npx, int npy, int nt_o) {
    benchmark_mode = ini.get_l("benchmark_mode", false);
    title = ini.get("title", "swpc3d");
    nproc_x = ini.get_i("nproc_x", 1);
    nproc_y = ini.get_i("nproc_y", 2);
    nx = ini.get_i("nx", 256); ny = ini.get_i("ny", 256); nz = ini.get_i("nz", 256);
    nt = ini.get_i("nt", 1000);
    ipad = ini.get_i("ipad", 0); jpad = ini.get_i("jpad", 0); kpad = ini.get_i("kpad", 0);
    odir = ini.get("odir", "./out");
    if (benchmark_mode) {   // m_global.f90:144-158
        dx = dy = dz = 0.5f;
        dt = 0.04f; na = 20;
        xbeg = -((float)nx / 2.0f * (float)dx);
        ybeg = -((float)ny / 2.0f * (float)dy);
        zbeg = -30 * (float)dz;
        tbeg = 0.0f; clon = 139.7604f; clat = 35.7182f; phi = 0.0f;
        abc_type = "pml";
    } else {
        dx = ini.get_d("dx", 0.5); dy = ini.get_d("dy", 0.5); dz = ini.get_d("dz", 0.5);
        dt = ini.get_s("dt", 0.01f);
        na = ini.get_i("na", 20);
        xbeg = ini.get_s("xbeg", -(float)(nx / 2) * (float)dx);
        ybeg = ini.get_s("ybeg", -(float)(ny / 2) * (float)dy);
        zbeg = ini.get_s("zbeg", -30 * (float)dz);
        tbeg = ini.get_s("tbeg", 0.0f);
        clon = ini.get_s("clon", 139.7604f); clat = ini.get_s("clat", 35.7182f); phi = ini.get_s("phi", 0.0f);
        abc_type = ini.get("abc_type", "pml");
    }
    if (npx > 0) nproc_x = npx;
    if (npy > 0) nproc_y = npy;
    if (nt_o > 0) nt = nt_o;
    xend = xbeg + nx * (float)dx; yend = ybeg + ny * (float)dy; zend = zbeg + nz * (float)dz;
    if (myid < 0 || myid >= nproc_x * nproc_y) return hfail("myid outside of nproc_x*nproc_y (assert, m_global.f90:232)");
    if (abc_type != "pml" && abc_type != "cerjan") return hfail("abc_type must be 'pml' or 'cerjan' (assert, m_absorb.f90:37)");
    UC = 1e-15f;
    const time_t now = time(nullptr);
    struct tm lt;
    localtime_r(&now, &lt);
    exedate = (int)now;
    tz_minutes = (int)(lt.tm_gmtoff / 60);
    return 0;
}

int swpc3d_host::setup_geometry() {
    idx = myid % nproc_x; idy = myid / nproc_x;
    decomp1d(nx, nproc_x, idx, nxp, ibeg, iend);
    decomp1d(ny, nproc_y, idy, nyp, jbeg, jend);
    ibeg_m = ibeg - 3; iend_m = iend + 3 + ipad; jbeg_m = jbeg - 3; jend_m = jend + 3 + jpad; kbeg_m = -2; kend_m = nz + 3 + kpad;
    nzm = kend_m - kbeg_m + 1; nxm = iend_m - ibeg_m + 1; nym = jend_m - jbeg_m + 1;
    if ((ibeg_m <= na && iend_m < na + 1) || (iend_m >= nx - na + 1 && ibeg_m > nx - na) || (jbeg_m <= na && jend_m < na + 1) ||
        (jend_m >= ny - na + 1 && jbeg_m > ny - na))
        return hfail("subdomain narrower than the absorber: the reference's absorber homogenisation (m_medium.f90:124-176) would read out of bounds");
    xc.resize(nxm); yc.resize(nym); zc.resize(nzm);
    for (int i = ibeg_m; i <= iend_m; i++) xc[i - ibeg_m] = i2x(i, xbeg, (float)dx);
    for (int j = jbeg_m; j <= jend_m; j++) yc[j - jbeg_m] = i2x(j, ybeg, (float)dy);
    for (int k = kbeg_m; k <= kend_m; k++) zc[k - kbeg_m] = i2x(k, zbeg, (float)dz);
    kbeg_a.assign((size_t)nxm * nym, 0);
    for (int j = jbeg_m; j <= jend_m; j++)
        for (int i = ibeg_m; i <= iend_m; i++)
            kbeg_a[i2(i, j)] = (i <= na || nx - na + 1 <= i || j <= na || ny - na + 1 <= j) ? 1 : nz - na + 1;
    ibeg_k = ibeg; iend_k = iend; jbeg_k = jbeg; jend_k = jend; kbeg_k = 1; kend_k = nz;
    if (abc_type == "pml") {   // m_global.f90:355-376
        if (iend <= na) ibeg_k = iend + 1; else if (ibeg <= na) ibeg_k = na + 1;
        if (ibeg >= nx - na + 1) iend_k = ibeg - 1; else if (iend >= nx - na + 1) iend_k = nx - na;
        if (jend <= na) jbeg_k = jend + 1; else if (jbeg <= na) jbeg_k = na + 1;
        if (jbeg >= ny - na + 1) jend_k = jbeg - 1; else if (jend >= ny - na + 1) jend_k = ny - na;
        kend_k = nz - na;
    }
    return 0;
}

int swpc3d_host::setup_medium(const IniFile &ini) {
    const size_t nc = (size_t)nzm * nxm * nym, n2 = (size_t)nxm * nym;
    rho.assign(nc, 0.f); lam.assign(nc, 0.f); mu.assign(nc, 0.f); taup.assign(nc, 0.f); taus.assign(nc, 0.f);
    bddep.assign(n2 * (NBD + 1), -9999.0f);
    // every supported model is laterally uniform at build time: fill one column profile, then broadcast
    std::vector<float> p_rho(nzm), p_lam(nzm), p_mu(nzm), p_qp(nzm), p_qs(nzm);
    float bd0 = 0.0f;
    bool lateral = false;   // laterally heterogeneous builder: 3-D arrays are filled directly
    bool grd_bddep = false; // ... which also wrote the boundary depths bd(:,:,0:NBD) itself
    if (benchmark_mode) {   // m_medium.f90:55-74
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
        if (vmodel_type == "uni") {   // m_vmodel_uni.f90:50-136
            const float vp0 = ini.get_s("vp0", 5.0f);
            const float vs0 = ini.get_s("vs0", vp0 / std::sqrt(3.0f));
            const float rho0 = ini.get_s("rho0", 2.7f), qp0 = ini.get_s("qp0", 1000000.0f), qs0 = ini.get_s("qs0", 1000000.0f);
            bd0 = ini.get_s("topo0", 0.0f);
            for (int q = 0; q < nzm; q++) {
                float vp1, vs1, r1;
                if (zs[q] > bd0) { vp1 = Cv[q] * vp0; vs1 = Cv[q] * vs0; r1 = rho0; p_qp[q] = qp0; p_qs[q] = qs0; }
                else if (zc[q] > 0.0f) { vp1 = Cv[q] * seawater_vel(zs[q], munk); vs1 = 0.0f; r1 = 1.0f; p_qp[q] = 1000000.0f; p_qs[q] = 1000000.0f; }
                else { vp1 = 0.0f; vs1 = 0.0f; r1 = 0.001f; p_qp[q] = 10.0f; p_qs[q] = 10.0f; }
                p_rho[q] = r1;
                p_mu[q] = r1 * vs1 * vs1;
                p_lam[q] = r1 * (vp1 * vp1 - 2 * vs1 * vs1);
            }
        } else if (vmodel_type == "lhm") {   // m_vmodel_lhm.f90:55-160
            const std::string fn = join_path(base, ini.get("fn_lhm", ""));
            std::ifstream is(fn);
            if (!is) return hfail("vmodel_lhm: cannot open " + fn + " (assert, m_vmodel_lhm.f90:72-76)");
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
            for (int l = nl - 2; l >= 0; l--)   // velocity cut-off :100-109
                if ((vp0[l] < vcut || vs0[l] < vcut) && (vp0[l] > 0 && vs0[l] > 0)) {
                    vp0[l] = vp0[l + 1]; vs0[l] = vs0[l + 1]; r0[l] = r0[l + 1]; qp0[l] = qp0[l + 1]; qs0[l] = qs0[l + 1];
                }
            bd0 = depth[0];
            for (int q = 0; q < nzm; q++) {
                if (zs[q] < depth[0]) {
                    if (zs[q] < 0.0f) { p_rho[q] = 0.001f; p_mu[q] = 0.0f; p_lam[q] = 0.0f; p_qp[q] = 10.0f; p_qs[q] = 10.0f; }
                    else {
                        const float vp1 = Cv[q] * seawater_vel(zc[q], munk);
                        p_rho[q] = 1.0f; p_mu[q] = 0.0f; p_lam[q] = 1.0f * vp1 * vp1; p_qp[q] = 1000000.0f; p_qs[q] = 1000000.0f;
                    }
                    continue;
                }
                float r1 = 0, vp1 = 0, vs1 = 0, a = 0, b = 0;
                for (int l = 0; l < nl; l++)
                    if (zs[q] >= depth[l]) { r1 = r0[l]; vp1 = Cv[q] * vp0[l]; vs1 = Cv[q] * vs0[l]; a = qp0[l]; b = qs0[l]; }
                p_rho[q] = r1; p_mu[q] = r1 * vs1 * vs1; p_lam[q] = r1 * (vp1 * vp1 - 2 * vs1 * vs1); p_qp[q] = a; p_qs[q] = b;
            }
        } else if (vmodel_type == "grd" || vmodel_type == "grd_rmed") {   // models.hpp: GMT grids (netCDF classic) + bicubic interpolation
            lateral = true;
            grd_bddep = true;
            const MediumBox mb{ibeg_m, iend_m, jbeg_m, jend_m, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
            const ModelEnv env{&ini, base, vcut, dt, dx, dy, dz, munk, ef};
            const GrdGeometry gg{nx, ny, na, xbeg, ybeg, zbeg, clon, clat, phi, xc.data(), yc.data(), bddep.data(), nz};
            if (vmodel_grd(env, mb, gg, vmodel_type == "grd_rmed")) return 1;
        } else if (vmodel_type == "lgm" || vmodel_type == "uni_rmed" || vmodel_type == "lhm_rmed" || vmodel_type == "lgm_rmed") {
            // models.hpp: these fill the 3-D arrays directly (Qp / Qs go to taup / taus, as in the reference's call)
            lateral = true;
            const MediumBox mb{ibeg_m, iend_m, jbeg_m, jend_m, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
            const ModelEnv env{&ini, base, vcut, dt, dx, dy, dz, munk, ef};
            const int rc = vmodel_type == "lgm" ? vmodel_lgm(env, mb, bd0) : vmodel_type == "uni_rmed" ? vmodel_uni_rmed(env, mb, bd0)
                         : vmodel_type == "lhm_rmed" ? vmodel_lhm_rmed(env, mb, bd0) : vmodel_lgm_rmed(env, mb, bd0);
            if (rc) return 1;
        } else {
            return hfail("vmodel_type '" + vmodel_type + "' is not available in this build ('user' is a compile-time plug-in of the reference: "
                         "link your own filler against swpc3d_b200.h's upload_medium instead)");
        }
    }
    if (!grd_bddep) for (size_t n = 0; n < n2; n++) bddep[n] = bd0;
    if (std::getenv("SWPC3D_HOST_TIMING")) std::fprintf(stderr, "[swpc3d_host] vmodel done\n");
    if (lateral) return finish_medium_3d(ini);

    // absorber homogenisation (m_medium.f90:124-190) copies columns outward; with laterally uniform input it is the
    // identity in x and y, and in z it repeats the value at k = nz-na below it
    for (int q = nz - na + 1 - kbeg_m; q < nzm; q++) {
        const int s = nz - na - kbeg_m;
        p_rho[q] = p_rho[s]; p_lam[q] = p_lam[s]; p_mu[q] = p_mu[s]; p_qp[q] = p_qp[s]; p_qs[q] = p_qs[s];
    }
    // tau-method (:193-207) and relaxed moduli (:231-271), still per depth
    relax_times(nm, ts, fq_min, fq_max);
    zeta = constq_zeta(nm, fq_min, fq_max, ts);
    if (benchmark_mode) zeta = 0.0f;
    std::vector<float> p_tp(nzm), p_tsx(nzm);
    for (int q = 0; q < nzm; q++) { p_tp[q] = nm * zeta / p_qp[q]; p_tsx[q] = nm * zeta / p_qs[q]; }
    if (nm > 0) {
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
    for (int j = 0; j < nym; j++)
        for (int i = 0; i < nxm; i++) {
            const size_t o = (size_t)nzm * ((size_t)i + (size_t)nxm * j);
            std::copy(p_rho.begin(), p_rho.end(), rho.begin() + o);
            std::copy(p_lam.begin(), p_lam.end(), lam.begin() + o);
            std::copy(p_mu.begin(), p_mu.end(), mu.begin() + o);
            std::copy(p_tp.begin(), p_tp.end(), taup.begin() + o);
            std::copy(p_tsx.begin(), p_tsx.end(), taus.begin() + o);
        }

    return finish_surface(ini);
}

// medium__setup after a laterally heterogeneous vmodel_*: absorber homogenisation (m_medium.f90:124-190), tau-method
// (:193-207) and relaxed moduli (:231-271) cell by cell
int swpc3d_host::finish_medium_3d(const IniFile &ini) {
    auto cp5 = [&](size_t d, size_t s_) { rho[d] = rho[s_]; lam[d] = lam[s_]; mu[d] = mu[s_]; taup[d] = taup[s_]; taus[d] = taus[s_]; };
    for (int i = ibeg_m; i <= na; i++)
        for (int j = jbeg_m; j <= jend_m; j++)
            for (int k = kbeg_m; k <= kend_m; k++) cp5(i3(k, i, j), i3(k, na + 1, j));
    for (int i = nx - na + 1; i <= iend_m; i++)
        for (int j = jbeg_m; j <= jend_m; j++)
            for (int k = kbeg_m; k <= kend_m; k++) cp5(i3(k, i, j), i3(k, nx - na, j));
    for (int j = jbeg_m; j <= na; j++)
        for (int i = ibeg_m; i <= iend_m; i++)
            for (int k = kbeg_m; k <= kend_m; k++) cp5(i3(k, i, j), i3(k, i, na + 1));
    for (int j = ny - na + 1; j <= jend_m; j++)
        for (int i = ibeg_m; i <= iend_m; i++)
            for (int k = kbeg_m; k <= kend_m; k++) cp5(i3(k, i, j), i3(k, i, ny - na));
    for (int j = jbeg_m; j <= jend_m; j++)
        for (int i = ibeg_m; i <= iend_m; i++)
            for (int k = nz - na + 1; k <= kend_m; k++) cp5(i3(k, i, j), i3(nz - na, i, j));
    relax_times(nm, ts, fq_min, fq_max);
    zeta = constq_zeta(nm, fq_min, fq_max, ts);
    const size_t nc = rho.size();
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < nc; n++) { taup[n] = nm * zeta / taup[n]; taus[n] = nm * zeta / taus[n]; }
    if (nm > 0) {
        const float omega = (float)(2 * PI_D * (double)fq_ref);
        std::complex<float> cc(0.0f, 0.0f);
        for (int m = 0; m < nm; m++) {
            const std::complex<double> w = std::complex<double>(0.0, 1.0) * (double)omega * (double)ts[m];
            const std::complex<double> qd = w / (1.0 - w);
            cc = cc + std::complex<float>((float)qd.real(), (float)qd.imag());
        }
        cc = std::complex<float>(cc.real() / (float)nm, cc.imag() / (float)nm);
        // chi depends on the cell only through taus / taup, which take a handful of values (one per layer): remember the
        // last evaluation per column instead of two complex square roots per cell
        const long long ncol = (long long)nxm * nym;
#pragma omp parallel for schedule(static)
        for (long long col = 0; col < ncol; col++) {
            float last_ts = -1.0f, last_tp = -1.0f, chi_mu = 1.0f, chi_lam = 1.0f;
            for (size_t n = (size_t)col * nzm; n < (size_t)(col + 1) * nzm; n++) {
                const float rb2 = mu[n], ra2 = lam[n] + 2 * mu[n];
                if (taus[n] != last_ts) { last_ts = taus[n]; chi_mu = 1.0f / (1.0f / std::sqrt(1.0f - cc * taus[n])).real(); }
                if (taup[n] != last_tp) { last_tp = taup[n]; chi_lam = 1.0f / (1.0f / std::sqrt(1.0f - cc * taup[n])).real(); }
                mu[n] = rb2 / (chi_mu * chi_mu);
                lam[n] = ra2 / (chi_lam * chi_lam) - 2 * mu[n];
            }
        }
    }
    return finish_surface(ini);
}

int swpc3d_host::finish_surface(const IniFile &ini) {
    const size_t n2 = (size_t)nxm * nym;
    // surface_detection m_medium.f90:339-394 -- evaluated per column exactly as the reference does (Q1: only
    // i in [ibeg-1, iend+2], j likewise are scanned; the outermost margin keeps kbeg-1 = 0)
    kfs.assign(n2, 0); kob.assign(n2, 0); kfs_top.assign(n2, 0); kfs_bot.assign(n2, 0); kob_top.assign(n2, 0); kob_bot.assign(n2, 0);
#pragma omp parallel for schedule(static)
    for (int j = jbeg - 1; j <= jend + 2; j++)
        for (int i = ibeg - 1; i <= iend + 2; i++)
            for (int k = 1; k <= nz - 1; k++) {
                const size_t a = i3(k, i, j), b = i3(k + 1, i, j);
                if (std::fabs(mu[a]) < EPS_SP && std::fabs(mu[b]) > EPS_SP) kob[i2(i, j)] = k;
                if (std::fabs(lam[a]) < EPS_SP && std::fabs(lam[b]) > EPS_SP) kfs[i2(i, j)] = k;
            }
    for (int j = jbeg; j <= jend; j++)
        for (int i = ibeg; i <= iend; i++) {
            int f0_ = 1 << 30, f1 = -(1 << 30), o0 = 1 << 30, o1 = -(1 << 30);
            for (int jj = j - 2; jj <= j + 3; jj++)
                for (int ii = i - 2; ii <= i + 3; ii++) {
                    f0_ = std::min(f0_, kfs[i2(ii, jj)]); f1 = std::max(f1, kfs[i2(ii, jj)]);
                    o0 = std::min(o0, kob[i2(ii, jj)]); o1 = std::max(o1, kob[i2(ii, jj)]);
                }
            kfs_top[i2(i, j)] = std::max(f0_ - 2, 1); kfs_bot[i2(i, j)] = std::min(f1 + 2, nz);
            kob_top[i2(i, j)] = std::max(o0 - 2, 1); kob_bot[i2(i, j)] = std::min(o1 + 2, nz);
        }
    // velocity_minmax :396-427 (local part)
    float vmx = -1.0f, vmn = 1e30f;
#pragma omp parallel for schedule(static) reduction(max : vmx) reduction(min : vmn)
    for (int j = jbeg; j <= jend; j++)
        for (int i = ibeg; i <= iend; i++)
            for (int k = kfs[i2(i, j)] + 1; k <= nz; k++) {
                const size_t n = i3(k, i, j);
                const float vp = std::sqrt((lam[n] + 2 * mu[n]) / rho[n]), vs = std::sqrt(mu[n] / rho[n]);
                vmx = std::max(vmx, vp);
                if (vs < EPS_SP) continue;
                vmn = std::min(vmn, vs);
            }
    vmin_local = vmin = vmn;
    vmax_local = vmax = vmx;
    // stabilize_absorber (m_medium.f90:218-222) needs the GLOBAL vmax: applied here for a single-rank run, otherwise when
    // the caller hands over the reduced values (swpc3d_host_set_minmax) or, at the latest, before the upload
    stabilize_pending = ini.get_l("stabilize_pml", false);
    if (stabilize_pending && nproc_x * nproc_y == 1) apply_stabilize();
    return 0;
}

void swpc3d_host::apply_stabilize() {
    if (!stabilize_pending) return;
    stabilize_pending = false;
    const MediumBox mb{ibeg_m, iend_m, jbeg_m, jend_m, kbeg_m, kend_m, zc.data(), rho.data(), lam.data(), mu.data(), taup.data(), taus.data()};
    stabilize_absorber(mb, kbeg_a.data(), ibeg, iend, jbeg, jend, nz, vmax);
}

void swpc3d_host::setup_kernel() {   // m_kernel.f90:58-67 (the device computes its own copy; kept for reporting)
    d2 = 0.0f;
    if (nm > 0) {
        float sum = 0.0f;
        for (int m = 0; m < nm; m++) {
            c1[m] = (2 * ts[m] - dt) / (2 * ts[m] + dt);
            c2[m] = (2) / (2 * ts[m] + dt) / nm;
            d1[m] = 2 * ts[m] / (2 * ts[m] - dt);
            sum += dt / (2 * ts[m] - dt);
        }
        d2 = sum / nm;
    }
}

// pw_setup, m_source.f90:316-466: plane P / S wave as the initial condition of all nine fields over the whole memory
// box; no source grid afterwards, M0 = 1/UC
int swpc3d_host::setup_planewave(const IniFile &ini) {
    const float pw_ztop = ini.get_s("pw_ztop", 1e30f);
    if (!(pw_ztop < zend)) return hfail("assert: pw_ztop < zend (m_source.f90:332)");
    const float pw_zlen = ini.get_s("pw_zlen", -1.0f);
    if (!(pw_zlen > 0.0f)) return hfail("assert: pw_zlen > 0 (m_source.f90:335)");
    const std::string ps = ini.get("pw_ps", "");
    const bool is_p = ps == "p" || ps == "P", is_s = ps == "s" || ps == "S";
    if (!(is_p || is_s)) return hfail("assert: pw_ps must be 'p' or 's' (m_source.f90:338)");
    const float strike = deg2rad_s(ini.get_s("pw_strike", 0.0f)), dip = deg2rad_s(ini.get_s("pw_dip", 0.0f)), rake = deg2rad_s(ini.get_s("pw_rake", 0.0f));
    stftype = ini.get("stftype", "kupper");
    const float sd = std::sin(dip), cd = std::cos(dip), sf = std::sin(strike), cf = std::cos(strike), sl = std::sin(rake), cl = std::cos(rake);
    const float c2d = std::cos(2 * dip), c2f = std::cos(2 * strike);
    const size_t nc = (size_t)nzm * nxm * nym;
    for (auto &a : pw_init) a.assign(nc, 0.0);
    const bool sp = field_bytes == 4;   // MP = SP: dx, dy, dz are single there
    auto coord = [&](float beg, int i, double d, float sub) {
        return sp ? (beg + ((float)i - 0.5f) * (float)d) - sub : (float)(((double)beg + (double)((float)i - 0.5f) * d) - (double)sub);
    };
    auto half = [&](float x, double d) { return sp ? x + (float)d / 2.0f : (float)((double)x + d / 2.0); };
    const float a = sd * sf, b = sd * cf;
    for (int j = jbeg - 3; j <= jend + 3 + jpad; j++)
        for (int i = ibeg - 3; i <= iend + 3 + ipad; i++)
            for (int k = kbeg_m; k < kbeg_m + nzm; k++) {
                const size_t n = i3(k, i, j);
                const float la0 = lam[n], mu0 = mu[n];
                const float v = is_p ? std::sqrt((la0 + 2 * mu0) / rho[n]) : std::sqrt(mu0 / rho[n]);
                if (v < EPS_SP) continue;
                const float x0 = coord(xbeg, i, dx, 0.0f), y0 = coord(ybeg, j, dy, 0.0f), z0 = coord(zbeg, k, dz, pw_ztop);
                const float x1 = half(x0, dx), y1 = half(y0, dy), z1 = half(z0, dz);
                const float stf_ii = momentrate(a * x0 - b * y0 + cd * z0, stftype, 0.0f, pw_zlen);
                const float stf_vx = momentrate(a * x1 - b * y0 + cd * z0 + dt / 2.0f * v, stftype, 0.0f, pw_zlen);
                const float stf_vy = momentrate(a * x0 - b * y1 + cd * z0 + dt / 2.0f * v, stftype, 0.0f, pw_zlen);
                const float stf_vz = momentrate(a * x0 - b * y0 + cd * z1 + dt / 2.0f * v, stftype, 0.0f, pw_zlen);
                const float stf_yz = momentrate(a * x0 - b * y1 + cd * z1, stftype, 0.0f, pw_zlen);
                const float stf_xz = momentrate(a * x1 - b * y0 + cd * z1, stftype, 0.0f, pw_zlen);
                const float stf_xy = momentrate(a * x1 - b * y1 + cd * z0, stftype, 0.0f, pw_zlen);
                float f[9];
                if (is_p) {
                    f[0] = -sd * sf * stf_vx; f[1] = sd * cf * stf_vy; f[2] = -cd * stf_vz;
                    f[3] = -(la0 + 2 * mu0 * sd * sd * sf * sf) * stf_ii / v;
                    f[4] = -(la0 + 2 * mu0 * sd * sd * cf * cf) * stf_ii / v;
                    f[5] = -(la0 + 2 * mu0 * cd * cd) * stf_ii / v;
                    f[6] = 2 * mu0 * sd * cd * cf * stf_yz / v;
                    f[7] = -(2 * mu0 * sd * cd * sf * stf_xz / v);
                    f[8] = 2 * mu0 * sd * cd * sf * stf_xy / v;
                } else {
                    f[0] = (cl * cf + sl * cd * sf) * stf_vx; f[1] = (cl * sf - sl * cd * cf) * stf_vy; f[2] = -sl * sd * stf_vz;
                    f[3] = 2 * mu0 * sd * sf * (cl * cf + sl * cd * sf) * stf_ii / v;
                    f[4] = -(2 * mu0 * sd * cf * (cl * sf - sl * cd * cf) * stf_ii / v);
                    f[5] = -(2 * mu0 * cd * sl * sd * stf_ii / v);
                    f[6] = mu0 * (cl * cd * sf - sl * c2d * cf) * stf_yz / v;
                    f[7] = mu0 * (cl * cd * cf + sl * c2d * sf) * stf_xz / v;
                    f[8] = -(mu0 * (cl * sd * c2f + 2 * sl * sd * cd * sf * cf) * stf_xy / v);
                }
                for (int q = 0; q < 9; q++) pw_init[q][n] = (double)f[q];
            }
    // wavelength condition :445-464; the MPI_MAX over the ranks is trivial here: the medium is laterally uniform
    const int kk = x2i(pw_ztop, zbeg, (float)dz);
    const size_t n = i3(kk, ibeg, jbeg);
    fcut = (is_p ? std::sqrt((lam[n] + 2 * mu[n]) / rho[n]) : std::sqrt(mu[n] / rho[n])) / pw_zlen;
    fmax = fcut * 2.0f;
    M0 = 1.0f / UC;   // fictitious scalar moment for output :77
    src_ijk.clear(); mo.clear(); mij.clear(); srcprm.clear();
    return 0;
}

int swpc3d_host::setup_source(const IniFile &ini) {   // m_source.f90:41-314
    pw_mode = ini.get_l("pw_mode", false);
    green_mode = ini.get_l("green_mode", false);
    bf_mode = ini.get_l("bf_mode", false);
    if (pw_mode && green_mode) return hfail("assert: pw_mode and green_mode are exclusive (m_source.f90:70)");
    if (pw_mode && !benchmark_mode) return setup_planewave(ini);   // :73-82
    if (green_mode && !benchmark_mode) { pw_mode = false; return 0; }   // :84-91: no source grid; M0 / fmax come from green__setup
    fn_stf = ini.get("fn_stf", "");
    stftype = ini.get("stftype", "kupper");
    if (stftype == "scosine") stftype = "cosine";
    stf_format = ini.get("stf_format", "xym0ij");
    sdep_fit = ini.get("sdep_fit", "asis");
    earth_flattening = ini.get_l("earth_flattening", false);

    struct Src { float x, y, z, t0, tr, mo, m[6]; };
    std::vector<Src> g;
    if (benchmark_mode) {   // :115-166
        stftype = "kupper"; bf_mode = false; pw_mode = false;
        Src s{};
        s.x = 0.0f; s.y = 0.0f; s.z = 5.0f; s.mo = 1e15f; s.t0 = 0.1f; s.tr = 2.0f;
        s.m[0] = s.m[1] = s.m[2] = 1 / std::sqrt(3.0f);
        g.push_back(s);
        evlo = clon; evla = clat; evdp = s.z;
        std::copy(s.m, s.m + 6, m0ij);
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
            if (bf_mode) {   // source__grid_bodyforce :741-768
                if (v.size() < 8 || !(ll || xy)) return hfail("source file: bad body-force record / invalid source type");
                if (xy) { s.x = v[0]; s.y = v[1]; } else geomap_g2c(v[0], v[1], clon, clat, phi, s.x, s.y);
                s.z = v[2]; s.t0 = v[3]; s.tr = v[4]; s.m[0] = v[5]; s.m[1] = v[6]; s.m[2] = v[7];
                if (g.empty()) { geomap_c2g(s.x, s.y, clon, clat, phi, evlo, evla); evdp = s.z; f0[0] = s.m[0]; f0[1] = s.m[1]; f0[2] = s.m[2]; otim = s.t0; }
                g.push_back(s);
                continue;
            }
            if (stf_format == "psmeca") {   // :641-666  lon lat z mzz mxx myy mxz myz mxy iex (dyn-cm)
                if (v.size() < 10) return hfail("source file: bad psmeca record");
                s.z = v[2]; s.m[2] = v[3]; s.m[0] = v[4]; s.m[1] = v[5]; s.m[4] = v[6]; s.m[3] = -v[7]; s.m[5] = -v[8];
                geomap_g2c(v[0], v[1], clon, clat, phi, s.x, s.y);
                const float M0tmp = std::sqrt(s.m[0] * s.m[0] + s.m[1] * s.m[1] + s.m[2] * s.m[2] + 2 * (s.m[4] * s.m[4] + s.m[3] * s.m[3] + s.m[5] * s.m[5])) / std::sqrt(2.0f);
                s.mo = M0tmp * powi_sp(10.0f, (int)v[9]);
                s.t0 = 0.0f;
                s.tr = (float)((double)(2 * 1.05f * 1e-8f) * std::pow((double)s.mo, 1.0 / 3.0));   // 2 x empirical half-duration
                s.mo = s.mo * 1e-7f;
                for (int q = 0; q < 6; q++) s.m[q] = s.m[q] / M0tmp;
            } else if ((ll || xy) && kind == "dsdc") {   // :585-639  x y z tbeg trise D S strike dip rake
                if (v.size() < 10) return hfail("source file: bad dsdc record");
                if (xy) { s.x = v[0]; s.y = v[1]; } else geomap_g2c(v[0], v[1], clon, clat, phi, s.x, s.y);
                s.z = v[2]; s.t0 = v[3]; s.tr = v[4];
                sdr2moment(v[7] - phi, v[8], v[9], s.m);
                const int is0 = x2i(s.x, xbeg, (float)dx), js0 = x2i(s.y, ybeg, (float)dy);
                const int ks0 = earth_flattening ? x2i((float)(-R_EARTH * std::log((R_EARTH - (double)s.z) / R_EARTH)), zbeg, (float)dz) : x2i(s.z, zbeg, (float)dz);
                // mo = 1e9 mu(ks0,is0,js0) D S on the ranks that hold the cell, MPI_MAX over ranks (:691-696).  Every
                // velocity model of this build is laterally uniform, so the owner's mu equals this rank's own column.
                if (-1 <= is0 && is0 <= nx + 3 && -1 <= js0 && js0 <= ny + 3 && -1 <= ks0 && ks0 <= nz + 3) s.mo = (1e9f * mu[i3(ks0, ibeg, jbeg)]) * v[5] * v[6];
                else s.mo = 0.0f;
            } else {
            if (!(ll || xy) || !(kind == "m0ij" || kind == "m0dc" || kind == "mwij" || kind == "mwdc"))
                return hfail("invalid source type (stf_format '" + stf_format + "', m_source.f90:668-670)");
            const size_t need = kind[2] == 'i' ? 12 : 9;
            if (v.size() < need) return hfail("source file: bad moment record (assert(ierr == 0), m_source.f90:517)");
            if (xy) { s.x = v[0]; s.y = v[1]; } else geomap_g2c(v[0], v[1], clon, clat, phi, s.x, s.y);
            s.z = v[2]; s.t0 = v[3]; s.tr = v[4];
            s.mo = kind[1] == '0' ? v[5] : seismic_moment(v[5]);
            if (kind[2] == 'i') std::copy(v.begin() + 6, v.begin() + 12, s.m);
            else sdr2moment(v[6] - phi, v[7], v[8], s.m);
            }
            if (g.empty()) {   // :675-687
                geomap_c2g(s.x, s.y, clon, clat, phi, evlo, evla);
                sx0 = s.x; sy0 = s.y; evdp = s.z; std::copy(s.m, s.m + 6, m0ij); otim = s.t0;
            }
            g.push_back(s);
        }
    }
    if (earth_flattening)
        for (Src &s : g) s.z = -(float)(R_EARTH * std::log((R_EARTH - (double)s.z) / R_EARTH));
    fcut = 0.0f;
    for (const Src &s : g) fcut = std::max(fcut, 1 / s.tr);
    fmax = 2 * fcut;
    if (bf_mode) {
        float sum = 0.0f;
        for (const Src &s : g) sum += s.m[0] * s.m[0] + s.m[1] * s.m[1] + s.m[2] * s.m[2];
        M0 = std::sqrt(sum);
        UC = UC * 1000;
    } else {
        float sum = 0.0f;
        for (const Src &s : g) sum += s.mo;
        M0 = sum;
    }
    src_ijk.clear(); mo.clear(); mij.clear(); srcprm.clear();
    const size_t n2 = (size_t)nxm * nym;
    for (const Src &s : g) {
        int is = x2i(s.x, xbeg, (float)dx), js = x2i(s.y, ybeg, (float)dy), ks = x2i(s.z, zbeg, (float)dz);
        if (!(ibeg - 2 <= is && is <= iend + 3 && jbeg - 2 <= js && js <= jend + 3 && 1 - 2 <= ks && ks <= nz + 3)) continue;   // :209-211
        float sz = s.z;
        if (sdep_fit.size() == 3 && sdep_fit[0] == 'b' && sdep_fit[1] == 'd' && std::isdigit((unsigned char)sdep_fit[2])) {   // :263-271
            sz = bddep[(size_t)(sdep_fit[2] - '0') * n2 + i2(is, js)];
            ks = x2i(sz, zbeg, (float)dz);
        }
        if (!(xbeg <= s.x && s.x <= xend && ybeg <= s.y && s.y <= yend && zbeg <= sz && sz <= zend))
            return hfail("source__setup: source outside of the model space (assert, m_source.f90:277-281)");
        src_ijk.push_back(is); src_ijk.push_back(js); src_ijk.push_back(ks);
        srcprm.push_back(s.t0); srcprm.push_back(s.tr);
        if (bf_mode) {
            mo.push_back(0.0);
            for (int q = 0; q < 3; q++) mij.push_back(field_bytes == 8 ? (double)s.m[q] / (double)M0 : (double)(s.m[q] / M0));
            for (int q = 3; q < 6; q++) mij.push_back(0.0);
        } else {
            mo.push_back(field_bytes == 8 ? (double)s.mo / (double)M0 : (double)(s.mo / M0));   // :302, real(MP)/real(SP)
            for (int q = 0; q < 6; q++) mij.push_back((double)s.m[q]);
        }
    }
    return 0;
}

int swpc3d_host::setup_absorb() {
    const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
    if (abc_type == "pml") {   // m_absorb_p.f90:75-93
        const float hx = na * fdx, hy = na * fdy, hz = na * fdz;
        gxc.assign(4 * (size_t)nxp, 0.f); gxe.assign(4 * (size_t)nxp, 0.f);
        gyc.assign(4 * (size_t)nyp, 0.f); gye.assign(4 * (size_t)nyp, 0.f);
        gzc.assign(4 * (size_t)nz, 0.f); gze.assign(4 * (size_t)nz, 0.f);
        for (int i = ibeg; i <= iend; i++) {
            damping_profile(xc[i - ibeg_m], hx, xbeg, xend, na, fcut, dt, &gxc[4 * (i - ibeg)]);
            damping_profile(xc[i - ibeg_m] + fdx / 2.0f, hx, xbeg, xend, na, fcut, dt, &gxe[4 * (i - ibeg)]);
        }
        for (int j = jbeg; j <= jend; j++) {
            damping_profile(yc[j - jbeg_m], hy, ybeg, yend, na, fcut, dt, &gyc[4 * (j - jbeg)]);
            damping_profile(yc[j - jbeg_m] + fdy / 2.0f, hy, ybeg, yend, na, fcut, dt, &gye[4 * (j - jbeg)]);
        }
        for (int k = 1; k <= nz; k++) {
            damping_profile(zc[k - kbeg_m], hz, zbeg, zend, na, fcut, dt, &gzc[4 * (k - 1)]);
            damping_profile(zc[k - kbeg_m] + fdz / 2.0f, hz, zbeg, zend, na, fcut, dt, &gze[4 * (k - 1)]);
        }
    } else {   // m_absorb_c.f90:38-101
        const float alpha = 0.09f, Lx = na * fdx, Ly = na * fdy, Lz = na * fdz;
        auto sq = [](float v) { return v * v; };
        cgx_c.assign(nxm, 1.0f); cgx_b.assign(nxm, 1.0f); cgy_c.assign(nym, 1.0f); cgy_b.assign(nym, 1.0f); cgz_c.assign(nzm, 1.0f); cgz_b.assign(nzm, 1.0f);
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
        fill(jbeg, jend, jbeg_m, ny, fdy, Ly, cgy_c, cgy_b, false);
        fill(1, nz, kbeg_m, nz, fdz, Lz, cgz_c, cgz_b, true);   // the top band stays 1 (:86-93)
    }
    return 0;
}

int swpc3d_host::setup_wav(const IniFile &ini) {   // m_wav.f90:54-271
    ntdec_w = ini.get_i("ntdec_w", 10);
    ntdec_w_prg = ini.get_i("ntdec_w_prg", 0);
    stopwatch_mode = ini.get_l("stopwatch_mode", true);
    sw_wav_v = ini.get_l("sw_wav_v", false);
    sw_wav_u = ini.get_l("sw_wav_u", false);
    sw_wav_stress = ini.get_l("sw_wav_stress", false);
    sw_wav_strain = ini.get_l("sw_wav_strain", false);
    const bool other = sw_wav_u || sw_wav_stress || sw_wav_strain;
    wav_format = ini.get("wav_format", "sac");
    st_format = ini.get("st_format", "xy");
    fn_stloc = ini.get("fn_stloc", "");
    ntdec_r = ini.get_i("ntdec_r", 10);   // m_report.f90:47
    if (ini.get_l("green_mode", false)) sw_wav_v = sw_wav_u = sw_wav_stress = sw_wav_strain = false;   // m_wav.f90:76-81: table only
    else if (!(sw_wav_v || other)) return 0;
    ntw = (int)std::floor((float)(nt - 1) / (float)ntdec_w + 1.0f);
    std::ifstream is(join_path(base, fn_stloc));
    if (!is) return 0;   // 'no station location file found' :157-161
    const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
    const size_t n2 = (size_t)nxm * nym;
    std::string line;
    while (std::getline(is, line)) {
        if (blank_or_comment(line)) continue;
        std::istringstream ls(line);
        float a, b, z;
        std::string name, zsw;
        if (!(ls >> a >> b >> z >> name >> zsw)) continue;
        name = name.substr(0, 8);
        zsw = zsw.substr(0, 3);
        float x, y, lo, la;
        if (st_format == "xy") { x = a; y = b; geomap_c2g(x, y, clon, clat, phi, lo, la); }
        else if (st_format == "ll") { lo = a; la = b; geomap_g2c(lo, la, clon, clat, phi, x, y); }
        else return hfail("unknown st_format: " + st_format);
        const int is_ = x2i(x, xbeg, fdx), js = x2i(y, ybeg, fdy);
        int ks = x2i(z, zbeg, fdz);
        if (!(i2x(1, xbeg, fdx) < x && x < i2x(nx, xbeg, fdx) && i2x(1, ybeg, fdy) < y && y < i2x(ny, ybeg, fdy) && 1 < ks && ks < nz)) continue;
        if (!(ibeg <= is_ && is_ <= iend && jbeg <= js && js <= jend)) continue;   // owner rank only (:203)
        if (zsw == "dep") ks = x2i(z, zbeg, fdz);
        else if (zsw == "fsb") ks = kfs[i2(is_, js)] + 1;
        else if (zsw == "obb") ks = kob[i2(is_, js)] + 1;
        else if (zsw == "oba") ks = kob[i2(is_, js)] - 1;
        else if (zsw.size() == 3 && zsw[0] == 'b' && zsw[1] == 'd' && std::isdigit((unsigned char)zsw[2]))
            ks = x2i(bddep[(size_t)(zsw[2] - '0') * n2 + i2(is_, js)], zbeg, fdz);
        else ks = x2i(z, zbeg, fdz);
        if (ks > nz) ks = nz - 1;
        if (ks < 1) ks = 1 + 1;
        st_ijk.push_back(is_); st_ijk.push_back(js); st_ijk.push_back(ks);
        xst.push_back(x); yst.push_back(y); zst.push_back(z); stlo.push_back(lo); stla.push_back(la); stnm.push_back(name);
    }
    return 0;
}

// green__setup, m_green.f90:67-355.  The pseudo source is a station: wav__stquery finds it on its owner rank and the
// reference broadcasts location and indices (:161-183).  A host that runs several ranks does the same with
// swpc3d_host_green_query / swpc3d_host_green_set_source; a single-rank run resolves it here.  The list of grid points
// needs the source position (green_maxdist) and is read by green_finalize().
int swpc3d_host::setup_green(const IniFile &ini) {
    if (benchmark_mode) { green_mode = false; return 0; }
    green_mode = ini.get_l("green_mode", false);
    if (!green_mode) return 0;
    Green &g = green;
    g.stnm = ini.get("green_stnm", "").substr(0, 8);
    const std::string c = ini.get("green_cmp", "");
    g.cmp = c.empty() ? ' ' : c[0];
    g.trise = ini.get_s("green_trise", 1.0f);
    g.bforce = ini.get_l("green_bforce", false);
    g.maxdist = ini.get_s("green_maxdist", 1e30f);
    if (!(g.maxdist > 0.0f)) return hfail("assert: green_maxdist > 0 (m_green.f90:124)");
    M0 = 1;
    fmax = 2.0f / g.trise;
    g.fn_glst = ini.get("fn_glst", "");
    if (!std::ifstream(join_path(base, g.fn_glst)).good()) return hfail("assert: fn_glst '" + g.fn_glst + "' exists (m_green.f90:131-132)");
    g.fmt = ini.get("green_fmt", "xyz");
    if (g.fmt != "xyz" && g.fmt != "llz") return hfail("assert: green_fmt is 'xyz' or 'llz' (m_green.f90:135)");
    g.ntdec_w = ini.get_i("ntdec_w", 10);
    g.stftype = ini.get("stftype", "kupper");
    if (g.stftype == "scosine") g.stftype = "cosine";
    g.wav_format = ini.get("wav_format", "sac");
    if (g.cmp == 'x') g.f1[0] = 1.0f;
    else if (g.cmp == 'y') g.f1[1] = 1.0f;
    else if (g.cmp == 'z') g.f1[2] = 1.0f;
    else return hfail("no matching green_cmp (m_green.f90:151-153)");
    g.ntw = (int)std::floor((float)(nt - 1) / (float)g.ntdec_w + 1.0f);
    g.ncmp = g.bforce ? 9 : 6;
    for (size_t n = 0; n < stnm.size(); n++)
        if (stnm[n] == g.stnm) {
            for (int q = 0; q < 3; q++) g.src_ijk[q] = st_ijk[3 * n + q];
            g.src_xyz[0] = xst[n]; g.src_xyz[1] = yst[n]; g.src_xyz[2] = zst[n]; g.evlo0 = stlo[n]; g.evla0 = stla[n];
            g.have_src = true;
            break;
        }
    if (nproc_x * nproc_y == 1) {
        if (!g.have_src) return hfail("assert: station '" + g.stnm + "' (green_stnm) is inside the model (m_green.f90:165)");
        return green_finalize();
    }
    return 0;
}

int swpc3d_host::green_finalize() {   // m_green.f90:188-280
    Green &g = green;
    if (g.finalized) return 0;
    if (!g.have_src) return hfail("Green's-function mode: the pseudo source is unknown on this rank -- broadcast it from its owner "
                                  "(swpc3d_host_green_query / swpc3d_host_green_set_source, mpi_bcast of m_green.f90:176-183)");
    std::ifstream is(join_path(base, g.fn_glst));
    const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
    std::string line;
    while (std::getline(is, line)) {
        if (blank_or_comment(line)) continue;
        for (auto &ch : line) if (ch == ',') ch = ' ';
        std::istringstream ls(line);
        float a, b, z, x, y, lo, la;
        int id;
        if (!(ls >> a >> b >> z >> id)) return hfail("assert: fn_glst line '" + line + "' reads as three reals and an integer (m_green.f90:205-213)");
        if (g.fmt == "xyz") { x = a; y = b; geomap_c2g(x, y, clon, clat, phi, lo, la); }
        else { lo = a; la = b; geomap_g2c(lo, la, clon, clat, phi, x, y); }
        if (!(0 <= id && id <= 99999999)) return hfail("assert: 0 <= gid <= 99999999 (m_green.f90:214)");
        const float ddx = x - g.src_xyz[0], ddy = y - g.src_xyz[1];
        if (std::sqrt(ddx * ddx + ddy * ddy) > g.maxdist) continue;
        const int ii = x2i(x, xbeg, fdx), jj = x2i(y, ybeg, fdy), kk = x2i(z, zbeg, fdz);
        if (!(ibeg <= ii && ii <= iend && jbeg <= jj && jj <= jend)) continue;
        if (!(kob[i2(ii, jj)] <= kk && kk <= nz)) continue;   // only the solid part
        g.ijk.push_back(ii); g.ijk.push_back(jj); g.ijk.push_back(kk); g.gid.push_back(id);
        g.xg.push_back(x); g.yg.push_back(y); g.zg.push_back(z); g.lon.push_back(lo); g.lat.push_back(la);
    }
    g.finalized = true;
    return 0;
}

int swpc3d_host::setup(const IniFile &ini, int nm_, int myid_, int npx, int npy, int nt_o) {
    if (nm_ < 0 || nm_ > 3) return hfail("nm must be 0..3");
    nm = nm_;
    myid = myid_;
    const bool timing = std::getenv("SWPC3D_HOST_TIMING") != nullptr;   // development aid: seconds per setup stage to stderr
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[swpc3d_host] %-8s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    if (setup_global(ini, npx, npy, nt_o)) return 1;   // main.f90:64 order
    if (setup_geometry()) return 1;
    lap("geometry");
    if (setup_medium(ini)) return 1;
    lap("medium");
    setup_kernel();
    if (setup_source(ini)) return 1;
    lap("source");
    if (setup_absorb()) return 1;
    lap("absorb");
    if (setup_snap(ini)) return 1;   // main.f90:76 snap__setup (files are created when a device is attached)
    if (setup_wav(ini)) return 1;
    if (setup_green(ini)) return 1;   // main.f90:78
    lap("snap+wav");
    return 0;
}

// ============================================================================================================
// snapshots: m_snap.f90.  netCDF classic (CDF-1) files written by a small in-tree writer (no netCDF library in the
// image): dimensions, variables, attributes and their order follow write_nc_header (:627-745), newfile_*_nc
// (:475-845), wbuf_nc (:950-979), close_nc (:2191-2204) and output__put_maxval (:2295-2348).
struct SnapProd {
    bool on = false; int sec = 0, typ = 0, n1 = 0, n2 = 0, nvar = 0, ionode = 0;
    std::string coordinate, snaptype, fname; std::vector<std::string> vname; std::string vunit;
    float vmin[4] = {0, 0, 0, 0}, vmax[4] = {0, 0, 0, 0};
    NcFile *nc = nullptr;
    FILE *snp = nullptr;   // native stream file (snp_format='native')
};
struct SnapHost {
    int idec = 1, jdec = 1, kdec = 1, ntdec_s = 10, nxs = 0, nys = 0, nzs = 0, is0 = 0, is1 = 0, js0 = 0, js1 = 0, ks0 = 0, ks1 = 0;
    int k0_xy = 0, i0_yz = 0, j0_xz = 0;
    float z0_xy = 0, x0_yz = 0, y0_xz = 0;
    std::string snp_format;
    std::vector<float> xsnp, ysnp, zsnp;
    SnapProd p[15];
    bool any = false, opened = false;
    std::vector<float> tmp;
    bool native = false;
    // asynchronous output (the reference's mpi_ireduce / mpi_wait overlap, m_snap.f90:1057-1064): the time loop only starts
    // the device-side fetch of a record; a writer thread waits for it and writes the file while the sweeps go on
    struct Job { int q, slot, rec, it; };
    std::thread writer;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> jobs;
    bool stop = false, busy = false;
    int slot_busy[15][2] = {};
    std::string werr;
    void stop_writer() {
        if (!writer.joinable()) return;
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        writer.join();
    }
    ~SnapHost() { stop_writer(); for (auto &q : p) { delete q.nc; if (q.snp) std::fclose(q.snp); } }
};

int swpc3d_host::setup_snap(const IniFile &ini) {   // m_snap.f90:95-322
    delete snap;
    snap = new SnapHost();
    SnapHost &S = *snap;
    static const char *sec_name[5] = {"xy", "xz", "yz", "fs", "ob"}, *typ_name[3] = {"ps", "v", "u"};
    for (int sec = 0; sec < 5; sec++)
        for (int typ = 0; typ < 3; typ++) {
            SnapProd &P = S.p[sec * 3 + typ];
            P.sec = sec; P.typ = typ;
            P.on = ini.get_l(std::string(sec_name[sec]) + "_" + typ_name[typ] + "%sw", false);
            S.any = S.any || P.on;
        }
    S.z0_xy = ini.get_s("z0_xy", std::max(std::min(10.0f, zend), zbeg));
    S.x0_yz = ini.get_s("x0_yz", std::max(std::min(0.0f, xend), xbeg));
    S.y0_xz = ini.get_s("y0_xz", std::max(std::min(0.0f, yend), ybeg));
    S.idec = ini.get_i("idec", 1); S.jdec = ini.get_i("jdec", 1); S.kdec = ini.get_i("kdec", 1);
    S.ntdec_s = ini.get_i("ntdec_s", 10);
    S.snp_format = ini.get("snp_format", "native");
    S.nxs = (nx + (S.idec / 2)) / S.idec; S.nys = (ny + (S.jdec / 2)) / S.jdec; S.nzs = (nz + (S.kdec / 2)) / S.kdec;
    S.xsnp.resize(S.nxs); S.ysnp.resize(S.nys); S.zsnp.resize(S.nzs);
    for (int i = 1; i <= S.nxs; i++) S.xsnp[i - 1] = i2x(i * S.idec - (S.idec / 2), xbeg, (float)dx);
    for (int j = 1; j <= S.nys; j++) S.ysnp[j - 1] = i2x(j * S.jdec - (S.jdec / 2), ybeg, (float)dy);
    for (int k = 1; k <= S.nzs; k++) S.zsnp[k - 1] = i2x(k * S.kdec - (S.kdec / 2), zbeg, (float)dz);
    S.is0 = (int)std::ceil((float)(ibeg + S.idec / 2) / (float)S.idec); S.is1 = (int)std::floor((float)(iend + S.idec / 2) / (float)S.idec);
    S.js0 = (int)std::ceil((float)(jbeg + S.jdec / 2) / (float)S.jdec); S.js1 = (int)std::floor((float)(jend + S.jdec / 2) / (float)S.jdec);
    S.ks0 = (int)std::ceil((float)(1 + S.kdec / 2) / (float)S.kdec); S.ks1 = (int)std::floor((float)(nz + S.kdec / 2) / (float)S.kdec);
    S.k0_xy = x2i(S.z0_xy, zbeg, (float)dz); S.i0_yz = x2i(S.x0_yz, xbeg, (float)dx); S.j0_xz = x2i(S.y0_xz, ybeg, (float)dy);
    if (!S.any) return 0;
    if (S.snp_format != "netcdf" && S.snp_format != "native") return hfail("snp_format '" + S.snp_format + "': expected 'native' or 'netcdf'");
    S.native = S.snp_format == "native";
    // owner of a global index along one axis (global__getnode, m_global.f90:644-683)
    auto owner = [](int n, int nproc, int g) { for (int q = 0; q < nproc; q++) { int np, b, e; decomp1d(n, nproc, q, np, b, e); if (b <= g && g <= e) return q; } return 0; };
    const int idy_xz = owner(ny, nproc_y, S.j0_xz), idx_yz = owner(nx, nproc_x, S.i0_yz), nproc = nproc_x * nproc_y;
    for (int q = 0; q < 15; q++) {
        SnapProd &P = S.p[q];
        P.n1 = P.sec == 2 ? S.nys : S.nxs;
        P.n2 = (P.sec == 1 || P.sec == 2) ? S.nzs : S.nys;
        P.nvar = P.typ == 0 ? 4 : 3;
        P.coordinate = sec_name[P.sec];
        P.snaptype = P.typ == 0 ? "ps" : (P.typ == 1 ? "v3" : "u3");
        if (P.typ == 0) { P.vname = {"div", "rot_x", "rot_y", "rot_z"}; P.vunit = "1/s"; }
        else if (P.typ == 1) { P.vname = {"Vx", "Vy", "Vz"}; P.vunit = "m/s"; }
        else { P.vname = {"Ux", "Uy", "Uz"}; P.vunit = "m"; }
        // output node, m_snap.f90:163-191: v, u, ps -> 0, 1, 2 (xy: +0, fs: +3, ob: +6) mod nproc; xz / yz within their row / column
        const int ord = P.typ == 1 ? 0 : (P.typ == 2 ? 1 : 2);
        if (P.sec == 0) P.ionode = (ord) % nproc;
        else if (P.sec == 3) P.ionode = (3 + ord) % nproc;
        else if (P.sec == 4) P.ionode = (6 + ord) % nproc;
        else if (P.sec == 1) P.ionode = (ord % nproc_x) + nproc_x * idy_xz;
        else P.ionode = idx_yz + nproc_x * (ord % nproc_y);
        P.fname = title + ".3d." + sec_name[P.sec] + "." + typ_name[P.typ] + (S.native ? ".snp" : ".nc");
    }
    return 0;
}

// newfile_{xy,xz,yz}_nc: header + medium slices on the I/O rank (every rank takes part in the sum-reduce)
int swpc3d_host::snap_open_files(const std::string &dir) {
    SnapHost &S = *snap;
    if (!S.any || S.opened) return 0;
    make_dirs(dir);
    const size_t n2m = (size_t)nxm * nym;
    for (int q = 0; q < 15; q++) {
        SnapProd &P = S.p[q];
        if (!P.on) continue;
        const bool horiz = P.sec == 0 || P.sec == 3 || P.sec == 4;
        const int nmed = horiz ? 6 : 3;
        const size_t np = (size_t)P.n1 * P.n2;
        std::vector<std::vector<float>> med((size_t)nmed, std::vector<float>(np, 0.0f));
        const bool part = (P.sec == 1) ? (jbeg <= S.j0_xz && S.j0_xz <= jend) : (P.sec == 2 ? (ibeg <= S.i0_yz && S.i0_yz <= iend) : true);
        if (part) {
            const int a0 = P.sec == 2 ? S.js0 : S.is0, a1 = P.sec == 2 ? S.js1 : S.is1;
            const int b0 = (P.sec == 1 || P.sec == 2) ? S.ks0 : S.js0, b1 = (P.sec == 1 || P.sec == 2) ? S.ks1 : S.js1;
            for (int b = b0; b <= b1; b++)
                for (int a = a0; a <= a1; a++) {
                    int i, j, k;
                    if (P.sec == 1) { i = a * S.idec - S.idec / 2; j = S.j0_xz; k = b * S.kdec - S.kdec / 2; }
                    else if (P.sec == 2) { i = S.i0_yz; j = a * S.jdec - S.jdec / 2; k = b * S.kdec - S.kdec / 2; }
                    else { i = a * S.idec - S.idec / 2; j = b * S.jdec - S.jdec / 2; k = S.k0_xy; }
                    if (P.sec == 3) k = kfs[i2(i, j)] + 1;
                    if (P.sec == 4) k = kob[i2(i, j)] + 1;
                    const size_t o = (size_t)(a - 1) + (size_t)P.n1 * (size_t)(b - 1), n = i3(k, i, j);
                    med[0][o] = rho[n]; med[1][o] = lam[n]; med[2][o] = mu[n];
                    if (horiz) {
                        med[3][o] = -bddep[i2(i, j)] * 1000;
                        geomap_c2g(S.xsnp[(size_t)a - 1], S.ysnp[(size_t)b - 1], clon, clat, phi, med[4][o], med[5][o]);
                    }
                }
        }
        (void)n2m;
        for (int m = 0; m < nmed; m++)
            if (swpc3d_reduce_sum(dev, med[(size_t)m].data(), (int64_t)np, P.ionode)) return hfail(std::string("device: ") + swpc3d_last_error());
        if (myid != P.ionode) continue;
        if (S.native) {   // newfile_{xy,xz,yz} + write_snp_header :846-892: fixed stream header, then rho, lam, mu (, topo)
            P.snp = std::fopen((dir + "/" + P.fname).c_str(), "wb");
            if (!P.snp) return hfail("cannot create " + dir + "/" + P.fname);
            const std::vector<float> &a1 = P.sec == 2 ? S.ysnp : S.xsnp, &a2 = (P.sec == 1 || P.sec == 2) ? S.zsnp : S.ysnp;
            const int e1 = P.sec == 2 ? S.jdec : S.idec, e2 = (P.sec == 1 || P.sec == 2) ? S.kdec : S.jdec;
            std::string ttl = title; ttl.resize(80, ' ');
            auto wi = [&](int32_t v) { std::fwrite(&v, 4, 1, P.snp); };
            auto wf = [&](float v) { std::fwrite(&v, 4, 1, P.snp); };
            std::fwrite("STREAMIO", 1, 8, P.snp); std::fwrite("SWPC_3D ", 1, 8, P.snp); wi(6);
            std::fwrite(ttl.data(), 1, 80, P.snp); wi(exedate);
            std::fwrite(P.coordinate.data(), 1, 2, P.snp); std::fwrite(P.snaptype.data(), 1, 2, P.snp);
            wi(P.n1); wi(P.n2); wf(a1[0]); wf(a2[0]);
            wf(a1.size() > 1 ? a1[1] - a1[0] : 0.0f); wf(a2.size() > 1 ? a2[1] - a2[0] : 0.0f);
            wf(dt * (float)S.ntdec_s); wi(na / e1); wi(na / e2); wi(nmed); wi(P.nvar);
            wf(clon); wf(clat); wf(phi); wf(1.0f); wf(1.0f); wf(1.0f);
            for (int m = 0; m < (horiz ? 4 : 3); m++) std::fwrite(med[(size_t)m].data(), 4, np, P.snp);
            continue;
        }
        NcFile *nc = new NcFile();
        P.nc = nc;
        const std::vector<float> &x1 = P.sec == 2 ? S.ysnp : S.xsnp, &x2 = (P.sec == 1 || P.sec == 2) ? S.zsnp : S.ysnp;
        const std::string n1s = P.sec == 2 ? "y" : "x", n2s = (P.sec == 1 || P.sec == 2) ? "z" : "y";
        nc->dims = {{n1s, P.n1}, {n2s, P.n2}, {"t", 0}};
        auto mm = [](const std::vector<float> &v) { float lo = v[0], hi = v[0]; for (float f : v) { lo = std::min(lo, f); hi = std::max(hi, f); } return std::make_pair(lo, hi); };
        NcVar vx1; vx1.name = n1s; vx1.dimids = {0}; vx1.data = x1;
        NcVar vx2; vx2.name = n2s; vx2.dimids = {1}; vx2.data = x2;
        vx1.atts.push_back(att_text("long_name", n1s)); vx2.atts.push_back(att_text("long_name", n2s));
        if (horiz) { vx1.atts.push_back(att_text("standard_name", "projection_x_coordinate")); vx2.atts.push_back(att_text("standard_name", "projection_y_coordinate")); }
        vx1.atts.push_back(att_text("units", "km")); vx2.atts.push_back(att_text("units", "km"));
        vx1.atts.push_back(att_floats("actual_range", {x1.front(), x1.back()})); vx2.atts.push_back(att_floats("actual_range", {x2.front(), x2.back()}));
        NcVar vt; vt.name = "t"; vt.dimids = {2}; vt.rec = true;
        vt.atts = {att_text("long_name", "t"), att_text("units", "s")};
        nc->vars = {vx1, vx2, vt};
        static const char *mname[6] = {"rho", "lambda", "mu", "topo", "lon", "lat"};
        static const char *mlong[6] = {"rho", "lambda", "mu", "Topography", "Longitude", "Latitude"};
        static const char *munit[6] = {"10^3 kg/cm^3", "10^9 Pa", "10^9 Pa", "m", "degrees_east", "degrees_north"};
        for (int m = 0; m < nmed; m++) {
            NcVar v; v.name = mname[m]; v.dimids = {1, 0}; v.data = med[(size_t)m];
            if (horiz && m < 3) v.atts.push_back(att_text("coordinates", "lat lon"));
            v.atts.push_back(att_text("long_name", mlong[m]));
            v.atts.push_back(att_text("units", munit[m]));
            if (horiz && m == 3) v.atts.push_back(att_text("coordinates", "lat lon"));
            const auto r = mm(med[(size_t)m]);
            v.atts.push_back(att_floats("actual_range", {r.first, r.second}));
            nc->vars.push_back(v);
        }
        for (int v = 0; v < P.nvar; v++) {
            NcVar sv; sv.name = P.vname[(size_t)v]; sv.dimids = {2, 1, 0}; sv.rec = true;
            sv.atts = {att_text("long_name", P.vname[(size_t)v]), att_text("units", P.vunit)};
            if (horiz) sv.atts.push_back(att_text("coordinates", "lat lon"));
            sv.atts.push_back(att_floats("actual_range", {0.0f, 0.0f}));
            nc->vars.push_back(sv);
        }
        if ((P.sec == 3 || P.sec == 4) && P.typ != 0) {   // output__put_maxval defines these at close; reserved here
            static const char *xn[3] = {"max-V", "max-H", "max-A"};
            static const char *xl[3] = {"Maximum amplitude of the vertical component", "Maximum amplitude of the horizontal components", "Maximum amplitude of the vector motion"};
            for (int v = 0; v < 3; v++) {
                NcVar xv; xv.name = xn[v]; xv.dimids = {1, 0}; xv.data.assign(np, 0.0f);
                xv.atts = {att_text("long_name", xl[v]), att_text("coordinates", "lat lon"), att_text("units", P.typ == 1 ? "m/s" : "m"), att_floats("actual_range", {0.0f, 0.0f})};
                nc->vars.push_back(xv);
            }
        }
        const int d1 = P.sec == 2 ? S.jdec : S.idec, d2 = (P.sec == 1 || P.sec == 2) ? S.kdec : S.jdec;
        const float ds1 = (float)(d1 * (P.sec == 2 ? dy : dx)), ds2 = (float)(d2 * ((P.sec == 1 || P.sec == 2) ? dz : dy));
        nc->gatts = {att_text("generated_by", "SWPC"), att_text("codetype", "SWPC_3D "), att_int("hdrver", 6), att_text("title", title), att_int("exedate", exedate),
                     att_int("ns1", P.n1), att_int("ns2", P.n2), att_floats("beg1", {x1.front()}), att_floats("beg2", {x2.front()}), att_int("na1", na / d1), att_int("na2", na / d2),
                     att_floats("ds1", {ds1}), att_floats("ds2", {ds2}), att_int("nmed", nmed), att_int("nsnp", P.nvar), att_text("coordinate", P.coordinate),
                     att_text("datatype", P.snaptype), att_floats("dt", {dt * S.ntdec_s}), att_floats("evlo", {evlo}), att_floats("evla", {evla}), att_floats("evdp", {evdp}),
                     att_floats("evx", {sx0}), att_floats("evy", {sy0}), att_floats("clon", {clon}), att_floats("clat", {clat}), att_floats("phi", {phi})};
        if (!nc->create(dir + "/" + P.fname)) return hfail("cannot create " + dir + "/" + P.fname);
    }
    // device side
    swpc3d_snap_cfg c{};
    c.idec = S.idec; c.jdec = S.jdec; c.kdec = S.kdec; c.ntdec_s = S.ntdec_s; c.nxs = S.nxs; c.nys = S.nys; c.nzs = S.nzs;
    c.is0 = S.is0; c.is1 = S.is1; c.js0 = S.js0; c.js1 = S.js1; c.ks0 = S.ks0; c.ks1 = S.ks1; c.k0_xy = S.k0_xy; c.i0_yz = S.i0_yz; c.j0_xz = S.j0_xz;
    for (int q = 0; q < 15; q++) c.sw[q] = S.p[q].on ? 1 : 0;
    c.M0 = M0; c.UC = UC;
    if (swpc3d_snap_setup(dev, &c)) return hfail(std::string("device: ") + swpc3d_last_error());
    S.opened = true;
    return 0;
}

// one record of one product into its file (wbuf_nc :950-979 / write_reduce_array2d_r): runs on the writer thread
static void snap_put_record(SnapHost &S, SnapProd &P, const float *data, int rec, float tval) {
    const size_t np = (size_t)P.n1 * P.n2;
    if (P.snp) { std::fwrite(data, 4, np * (size_t)P.nvar, P.snp); std::fflush(P.snp); return; }
    if (!P.nc) return;
    (void)S;
    P.nc->put_record(P.nc->var_index("t"), rec, &tval, 1);
    for (int v = 0; v < P.nvar; v++) {
        const float *d = data + np * (size_t)v;
        const int vi = P.nc->var_index(P.vname[(size_t)v]);
        P.nc->put_record(vi, rec, d, np);
        for (size_t n = 0; n < np; n++) { P.vmax[v] = std::max(P.vmax[v], d[n]); P.vmin[v] = std::min(P.vmin[v], d[n]); }
        *P.nc->find_att(P.nc->vars[(size_t)vi], "actual_range") = att_floats("actual_range", {P.vmin[v], P.vmax[v]});
    }
    P.nc->flush_header();
}

// snap__write(it): device step, and at output steps the record of every product: the device-side fetch (reduce onto the I/O
// rank + copy to pinned host memory, on a stream of its own) is only STARTED here; the writer thread waits for it and writes
// the file (same record index it0/ntdec_s+1, same time it0*dt, same data as the reference's one-cycle-late wbuf_nc).  A
// record's host buffer (two per product) is reused two records later: the loop waits for the writer only then.
int swpc3d_host::snap_write(int it) {
    SnapHost &S = *snap;
    if (!S.any || !S.opened) return 0;
    if (swpc3d_snap_step(dev, it)) return hfail(std::string("device: ") + swpc3d_last_error());
    if (!(S.ntdec_s > 0 && (it - 1) % S.ntdec_s == 0)) return 0;
    const int rec = it / S.ntdec_s;   // stt(3) = it0/ntdec_s + 1, zero-based here
    if (!S.writer.joinable()) {
        S.stop = false;
        S.writer = std::thread([this, &S]() {
            for (;;) {
                SnapHost::Job j;
                {
                    std::unique_lock<std::mutex> lk(S.mu);
                    S.cv.wait(lk, [&] { return S.stop || !S.jobs.empty(); });
                    if (S.jobs.empty()) return;   // stop requested and nothing left
                    j = S.jobs.front();
                    S.jobs.pop_front();
                    S.busy = true;
                }
                const float *data = nullptr;
                std::string err;
                if (swpc3d_snap_fetch_end(dev, j.q, j.slot, &data)) err = std::string("device: ") + swpc3d_last_error();
                else if (data) snap_put_record(S, S.p[j.q], data, j.rec, j.it * dt);
                {
                    std::lock_guard<std::mutex> lk(S.mu);
                    S.slot_busy[j.q][j.slot] = 0;
                    S.busy = false;
                    if (!err.empty() && S.werr.empty()) S.werr = err;
                }
                S.cv.notify_all();
            }
        });
    }
    for (int q = 0; q < 15; q++) {
        SnapProd &P = S.p[q];
        if (!P.on) continue;
        const int slot = rec & 1;
        {
            std::unique_lock<std::mutex> lk(S.mu);
            S.cv.wait(lk, [&] { return S.slot_busy[q][slot] == 0; });   // the reference's mpi_wait before reusing the buffer
            if (!S.werr.empty()) return hfail("snapshot writer: " + S.werr);
            S.slot_busy[q][slot] = 1;
        }
        if (swpc3d_snap_fetch_begin(dev, q, P.ionode, slot)) return hfail(std::string("device: ") + swpc3d_last_error());
        { std::lock_guard<std::mutex> lk(S.mu); S.jobs.push_back(SnapHost::Job{q, slot, rec, it}); }
        S.cv.notify_all();
    }
    return 0;
}

// snap__closefiles :2206-2293 (+ output__put_maxval)
int swpc3d_host::snap_close() {
    SnapHost &S = *snap;
    if (!S.any || !S.opened) return 0;
    {   // every record that was started is on disk before the maxima are fetched and the files closed
        std::unique_lock<std::mutex> lk(S.mu);
        S.cv.wait(lk, [&] { return S.jobs.empty() && !S.busy; });
        if (!S.werr.empty()) return hfail("snapshot writer: " + S.werr);
    }
    S.stop_writer();
    for (int q = 0; q < 15; q++) {
        SnapProd &P = S.p[q];
        if (!P.on) continue;
        const size_t np = (size_t)P.n1 * P.n2;
        if ((P.sec == 3 || P.sec == 4) && P.typ != 0) {
            S.tmp.assign(np * 3, 0.0f);
            if (swpc3d_snap_fetch_max(dev, q, P.ionode, S.tmp.data())) return hfail(std::string("device: ") + swpc3d_last_error());
            if (P.nc) {
                static const char *xn[3] = {"max-V", "max-H", "max-A"};
                for (int v = 0; v < 3; v++) {
                    NcVar &xv = P.nc->vars[(size_t)P.nc->var_index(xn[v])];
                    xv.data.assign(S.tmp.begin() + (long)(np * (size_t)v), S.tmp.begin() + (long)(np * (size_t)(v + 1)));
                    float lo = xv.data[0], hi = xv.data[0];
                    for (float f : xv.data) { lo = std::min(lo, f); hi = std::max(hi, f); }
                    *P.nc->find_att(xv, "actual_range") = att_floats("actual_range", {lo, hi});
                    P.nc->put_fixed(xv);
                }
            }
        }
        if (P.nc) { P.nc->flush_header(); delete P.nc; P.nc = nullptr; }
        if (P.snp) { std::fclose(P.snp); P.snp = nullptr; }
    }
    S.opened = false;
    return 0;
}

// ============================================================================================================
// C interface
extern "C" {

static int host_create(IniFile &ini, const char *base_dir, int nm, int myid, int npx, int npy, int nt, int fb, swpc3d_host **out) {
    if (!out) return hfail("null out");
    *out = nullptr;
    if (fb != 8 && fb != 4) return hfail("field_bytes must be 8 or 4");
    ini.strict = ini.get_l("strict_mode", false);   // main.f90:60-61
    swpc3d_host *h = new swpc3d_host();
    h->base = base_dir ? base_dir : "";
    h->field_bytes = fb;
    if (h->setup(ini, nm, myid, npx, npy, nt)) { delete h; return 1; }
    *out = h;
    return 0;
}

int swpc3d_host_create(const char *inf_path, const char *base_dir, int32_t nm, int32_t myid, int32_t npx, int32_t npy, int32_t nt,
                       int32_t fb, swpc3d_host **out) {
    IniFile ini;
    if (!inf_path || !IniFile::from_file(inf_path, ini)) return hfail(std::string("cannot open parameter file ") + (inf_path ? inf_path : "(null)"));
    return host_create(ini, base_dir, nm, myid, npx, npy, nt, fb, out);
}
int swpc3d_host_create_from_text(const char *text, const char *base_dir, int32_t nm, int32_t myid, int32_t npx, int32_t npy, int32_t nt,
                                 int32_t fb, swpc3d_host **out) {
    IniFile ini = IniFile::from_text(text ? text : "");
    return host_create(ini, base_dir, nm, myid, npx, npy, nt, fb, out);
}
int swpc3d_host_destroy(swpc3d_host *h) {
    if (!h) return 0;
    delete h->snap;
    if (h->dev) swpc3d_destroy(h->dev);
    delete h;
    return 0;
}

const char *swpc3d_host_last_error(void) { return g_herr.c_str(); }

int swpc3d_host_get_int(swpc3d_host *h, const char *name, int32_t *v) {
    if (!h || !name || !v) return hfail("null argument");
    const std::string n = name;
#define GI(x) if (n == #x) { *v = (int32_t)h->x; return 0; }
    GI(nx) GI(ny) GI(nz) GI(nt) GI(na) GI(nm) GI(nproc_x) GI(nproc_y) GI(myid) GI(ibeg) GI(iend) GI(jbeg) GI(jend) GI(nxp) GI(nyp)
    GI(ibeg_k) GI(iend_k) GI(jbeg_k) GI(jend_k) GI(kbeg_k) GI(kend_k) GI(ntw) GI(ntdec_w) GI(ntdec_r) GI(bf_mode) GI(nzm) GI(nxm) GI(nym)
    GI(exedate) GI(tz_minutes) GI(field_bytes)
#undef GI
    if (n == "nsrc") { *v = (int32_t)(h->src_ijk.size() / 3); return 0; }
    if (n == "nst") { *v = (int32_t)(h->st_ijk.size() / 3); return 0; }
    if (n == "green_mode") { *v = h->green_mode ? 1 : 0; return 0; }
    if (n == "ng") { *v = (int32_t)h->green.gid.size(); return 0; }
    if (n == "green_ncmp") { *v = h->green.ncmp; return 0; }
    if (n == "green_ntw") { *v = h->green.ntw; return 0; }
    return hfail("unknown int " + n);
}
int swpc3d_host_get_double(swpc3d_host *h, const char *name, double *v) {
    if (!h || !name || !v) return hfail("null argument");
    const std::string n = name;
#define GD(x) if (n == #x) { *v = (double)h->x; return 0; }
    GD(dx) GD(dy) GD(dz) GD(dt) GD(xbeg) GD(ybeg) GD(zbeg) GD(tbeg) GD(vmin) GD(vmax) GD(vmin_local) GD(vmax_local) GD(fmax) GD(fcut)
    GD(M0) GD(UC) GD(zeta) GD(d2) GD(loop_seconds) GD(evlo) GD(evla) GD(evdp) GD(clon) GD(clat) GD(phi)
#undef GD
    if (n == "c") { *v = (double)(h->dt / stable_dt((float)h->dx, (float)h->dy, (float)h->dz, h->vmax)); return 0; }   // m_fdtool.f90:99-113
    if (n == "r") {   // m_fdtool.f90:116-134
        const float dh = std::max(std::max((float)h->dx, (float)h->dy), (float)h->dz);
        *v = (double)((h->vmin / h->fmax) / dh);
        return 0;
    }
    return hfail("unknown double " + n);
}
int swpc3d_host_get_string(swpc3d_host *h, const char *name, char *buf, int32_t cap) {
    if (!h || !name || !buf || cap <= 0) return hfail("null argument");
    const std::string n = name;
    const std::string *s = nullptr;
    if (n == "title") s = &h->title; else if (n == "odir") s = &h->odir; else if (n == "abc_type") s = &h->abc_type;
    else if (n == "stftype") s = &h->stftype; else if (n == "vmodel_type") s = &h->vmodel_type; else if (n == "stf_format") s = &h->stf_format;
    else if (n == "wav_format") s = &h->wav_format;
    if (!s) return hfail("unknown string " + n);
    std::snprintf(buf, (size_t)cap, "%s", s->c_str());
    return 0;
}
int swpc3d_host_set_minmax(swpc3d_host *h, float vmin, float vmax) {
    if (!h) return hfail("null handle");
    h->vmin = vmin; h->vmax = vmax;
    h->apply_stabilize();
    return 0;
}
int swpc3d_host_set_exedate(swpc3d_host *h, int32_t exedate, int32_t tz) {
    if (!h) return hfail("null handle");
    h->exedate = exedate; h->tz_minutes = tz;
    return 0;
}

extern "C++" {
template <typename T>
static int put(const std::vector<T> &v, void *out, int64_t cap, int64_t *n) {
    if (n) *n = (int64_t)v.size();
    if (out) std::memcpy(out, v.data(), sizeof(T) * (size_t)std::min<int64_t>(cap, (int64_t)v.size()));
    return 0;
}
}
int swpc3d_host_get_array(swpc3d_host *h, const char *name, void *out, int64_t cap, int64_t *n) {
    if (!h || !name) return hfail("null argument");
    const std::string s = name;
#define GA(x) if (s == #x) return put(h->x, out, cap, n);
    GA(rho) GA(lam) GA(mu) GA(taup) GA(taus) GA(kfs) GA(kob) GA(kfs_top) GA(kfs_bot) GA(kob_top) GA(kob_bot) GA(kbeg_a)
    GA(gxc) GA(gxe) GA(gyc) GA(gye) GA(gzc) GA(gze) GA(src_ijk) GA(st_ijk) GA(mo) GA(mij) GA(srcprm) GA(xc) GA(yc) GA(zc)
    GA(stlo) GA(stla) GA(wav)
    if (s == "green_ijk") return put(h->green.ijk, out, cap, n);
    if (s == "green_gid") return put(h->green.gid, out, cap, n);
    if (s == "green_gf") return put(h->green.gf, out, cap, n);
    if (s == "wav_u") return put(h->wav_all[1], out, cap, n);
    if (s == "wav_stress") return put(h->wav_all[2], out, cap, n);
    if (s == "wav_strain") return put(h->wav_all[3], out, cap, n);
#undef GA
    static const char *fn[9] = {"init_Vx", "init_Vy", "init_Vz", "init_Sxx", "init_Syy", "init_Szz", "init_Syz", "init_Sxz", "init_Sxy"};
    for (int q = 0; q < 9; q++) if (s == fn[q]) return put(h->pw_init[q], out, cap, n);   // plane-wave initial condition (pw_mode)
    if (s == "gx_c") return put(h->cgx_c, out, cap, n);
    if (s == "gx_b") return put(h->cgx_b, out, cap, n);
    if (s == "gy_c") return put(h->cgy_c, out, cap, n);
    if (s == "gy_b") return put(h->cgy_b, out, cap, n);
    if (s == "gz_c") return put(h->cgz_c, out, cap, n);
    if (s == "gz_b") return put(h->cgz_b, out, cap, n);
    auto small = [&](const float *p) { std::vector<float> v(p, p + h->nm); return put(v, out, cap, n); };
    if (s == "ts") return small(h->ts);
    if (s == "c1") return small(h->c1);
    if (s == "c2") return small(h->c2);
    if (s == "d1") return small(h->d1);
    return hfail("unknown array " + s);
}
int swpc3d_host_station_name(swpc3d_host *h, int32_t i, char *buf9) {
    if (!h || !buf9 || i < 0 || (size_t)i >= h->stnm.size()) return hfail("bad station index");
    std::memset(buf9, 0, 9);
    std::strncpy(buf9, h->stnm[(size_t)i].c_str(), 8);
    return 0;
}

// `!$acc enter data copyin(...)` main.f90:80-113
int swpc3d_host_attach_device(swpc3d_host *h, int32_t device) {
    if (!h) return hfail("null handle");
    if (h->dev) { swpc3d_destroy(h->dev); h->dev = nullptr; }
    h->apply_stabilize();
    swpc3d_grid g{};
    g.nx = h->nx; g.ny = h->ny; g.nz = h->nz; g.nproc_x = h->nproc_x; g.nproc_y = h->nproc_y; g.myid = h->myid;
    g.ibeg = h->ibeg; g.iend = h->iend; g.jbeg = h->jbeg; g.jend = h->jend; g.ipad = h->ipad; g.jpad = h->jpad; g.kpad = h->kpad;
    g.ibeg_k = h->ibeg_k; g.iend_k = h->iend_k; g.jbeg_k = h->jbeg_k; g.jend_k = h->jend_k; g.kbeg_k = h->kbeg_k; g.kend_k = h->kend_k;
    g.na = h->na; g.nm = h->nm; g.abc_type = h->abc_type == "pml" ? SWPC3D_ABC_PML : SWPC3D_ABC_CERJAN;
    g.field_bytes = h->field_bytes; g.device = device; g.dx = h->dx; g.dy = h->dy; g.dz = h->dz; g.dt = h->dt;
#define DV(call) if (call) return hfail(std::string("device: ") + swpc3d_last_error());
    DV(swpc3d_create(&g, h->ts, &h->dev));
    DV(swpc3d_upload_medium(h->dev, h->rho.data(), h->lam.data(), h->mu.data(), h->taup.data(), h->taus.data(), h->kfs.data(), h->kob.data(),
                            h->kfs_top.data(), h->kfs_bot.data(), h->kob_top.data(), h->kob_bot.data(), h->kbeg_a.data()));
    if (g.abc_type == SWPC3D_ABC_PML) { DV(swpc3d_setup_pml(h->dev, h->gxc.data(), h->gxe.data(), h->gyc.data(), h->gye.data(), h->gzc.data(), h->gze.data())); }
    else { DV(swpc3d_setup_cerjan(h->dev, h->cgx_c.data(), h->cgx_b.data(), h->cgy_c.data(), h->cgy_b.data(), h->cgz_c.data(), h->cgz_b.data())); }
    const int nsrc = (int)(h->src_ijk.size() / 3);
    if (nsrc > 0) {
        std::vector<int> a(nsrc), b(nsrc), c(nsrc);
        std::vector<double> m[6];
        for (int q = 0; q < 6; q++) m[q].resize(nsrc);
        for (int i = 0; i < nsrc; i++) {
            a[i] = h->src_ijk[3 * i]; b[i] = h->src_ijk[3 * i + 1]; c[i] = h->src_ijk[3 * i + 2];
            for (int q = 0; q < 6; q++) m[q][i] = h->mij[6 * (size_t)i + q];
        }
        DV(swpc3d_set_sources(h->dev, nsrc, a.data(), b.data(), c.data(), h->mo.data(), m[0].data(), m[1].data(), m[2].data(), m[3].data(),
                              m[4].data(), m[5].data(), h->srcprm.data(), h->stftype.c_str(), h->bf_mode ? 1 : 0, h->tbeg));
    }
    if (h->pw_mode && !h->pw_init[0].empty()) {   // `!$acc enter data copyin(Vx..Sxy)` with the plane-wave initial condition
        const void *f[9];
        std::vector<float> f32[9];
        for (int q = 0; q < 9; q++) {
            if (h->field_bytes == 4) { f32[q].assign(h->pw_init[q].begin(), h->pw_init[q].end()); f[q] = f32[q].data(); }
            else f[q] = h->pw_init[q].data();
        }
        DV(swpc3d_upload_fields(h->dev, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8]));
        DV(swpc3d_set_option(h->dev, "pw_mode", 1));
    }
    if (h->green_mode) {   // `!$acc enter data copyin(ig, jg, kg, gf, stftype)` m_green.f90:351
        if (h->green_finalize()) return 1;
        const swpc3d_host::Green &gr = h->green;
        const int ng = (int)gr.gid.size();
        std::vector<int> a(ng), b(ng), c(ng);
        for (int i = 0; i < ng; i++) { a[i] = gr.ijk[3 * i]; b[i] = gr.ijk[3 * i + 1]; c[i] = gr.ijk[3 * i + 2]; }
        const bool is_src = (h->ibeg <= gr.src_ijk[0] && gr.src_ijk[0] <= h->iend + 1) && (h->jbeg <= gr.src_ijk[1] && gr.src_ijk[1] <= h->jend + 1) &&
                            (1 <= gr.src_ijk[2] && gr.src_ijk[2] <= h->nz);   // redefined is_src, :185-186
        DV(swpc3d_set_green(h->dev, ng, a.data(), b.data(), c.data(), gr.bforce ? 1 : 0, is_src ? 1 : 0, gr.src_ijk[0], gr.src_ijk[1], gr.src_ijk[2],
                            gr.f1[0], gr.f1[1], gr.f1[2], gr.trise, gr.stftype.c_str(), gr.ntdec_w, gr.ntw, h->tbeg));
    }
    const int nst = (int)(h->st_ijk.size() / 3);
    if (nst > 0 && (h->sw_wav_v || h->sw_wav_u || h->sw_wav_stress || h->sw_wav_strain)) {
        std::vector<int> a(nst), b(nst), c(nst);
        for (int i = 0; i < nst; i++) { a[i] = h->st_ijk[3 * i]; b[i] = h->st_ijk[3 * i + 1]; c[i] = h->st_ijk[3 * i + 2]; }
        DV(swpc3d_set_stations(h->dev, nst, a.data(), b.data(), c.data(), h->ntdec_w, h->ntw, h->M0, h->UC));
        DV(swpc3d_set_wav_products(h->dev, h->sw_wav_v, h->sw_wav_u, h->sw_wav_stress, h->sw_wav_strain));
    }
#undef DV
    return 0;
}
swpc3d_handle *swpc3d_host_handle(swpc3d_host *h) { return h ? h->dev : nullptr; }

// snapshot files: created on the I/O ranks (newfile_*_nc), one record per ntdec_s steps during swpc3d_host_run, closed here
int swpc3d_host_snap_open(swpc3d_host *h, const char *odir) {
    if (!h || !h->dev) return hfail("swpc3d_host_snap_open: no device attached");
    return h->snap_open_files(odir ? odir : h->odir.c_str());
}
int swpc3d_host_snap_close(swpc3d_host *h) {
    if (!h) return hfail("null handle");
    return h->snap_close();
}

// writer self-test (no GPU): a 3 x 2 grid, one fixed and one record variable, two records
int swpc3d_host_nc_selftest(const char *path) {
    NcFile nc;
    nc.dims = {{"x", 3}, {"z", 2}, {"t", 0}};
    NcVar x; x.name = "x"; x.dimids = {0}; x.data = {0.5f, 1.5f, 2.5f}; x.atts = {att_text("units", "km"), att_floats("actual_range", {0.5f, 2.5f})};
    NcVar t; t.name = "t"; t.dimids = {2}; t.rec = true; t.atts = {att_text("units", "s")};
    NcVar m; m.name = "rho"; m.dimids = {1, 0}; m.data = {1, 2, 3, 4, 5, 6};
    NcVar v; v.name = "Vx"; v.dimids = {2, 1, 0}; v.rec = true; v.atts = {att_text("long_name", "Vx"), att_floats("actual_range", {0.0f, 0.0f})};
    nc.vars = {x, t, m, v};
    nc.gatts = {att_text("generated_by", "SWPC"), att_int("hdrver", 6), att_floats("dt", {0.25f})};
    if (!nc.create(path ? path : "")) return hfail("cannot create file");
    for (int r = 0; r < 2; r++) {
        const float tv = 0.25f * (float)r, d[6] = {10.f * r, 10.f * r + 1, 10.f * r + 2, 10.f * r + 3, 10.f * r + 4, -10.f * r - 5};
        nc.put_record(1, r, &tv, 1);
        nc.put_record(3, r, d, 6);
    }
    *nc.find_att(nc.vars[3], "actual_range") = att_floats("actual_range", {-15.0f, 14.0f});
    nc.flush_header();
    return 0;
}

int swpc3d_host_banner(swpc3d_host *h) {   // m_report.f90:64-92
    if (!h) return hfail("null handle");
    double c, r;
    swpc3d_host_get_double(h, "c", &c);
    swpc3d_host_get_double(h, "r", &r);
    std::fprintf(stderr, "\n ------------------------------------------------------------------------------\n");
    std::fprintf(stderr, "  SWPC_3D (swpc3d_b200, B200-native time loop)%s\n", h->benchmark_mode ? " (benchmark mode) " : (h->bf_mode ? " (body force mode) " : ""));
    std::fprintf(stderr, " ------------------------------------------------------------------------------\n\n");
    std::fprintf(stderr, "  Grid Size               : %8d x %6d x %6d\n", h->nx, h->ny, h->nz);
    std::fprintf(stderr, "  MPI Partitioning        : %8d x %4d\n", h->nproc_x, h->nproc_y);
    std::fprintf(stderr, "  Stability  Condition c  : %15.3f  (c<1)\n", c);
    std::fprintf(stderr, "  Wavelength Condition r  : %15.3f  (r>5-10)\n", r);
    std::fprintf(stderr, "  Minimum velocity        : %15.3f  [km/s]\n", (double)h->vmin);
    std::fprintf(stderr, "  Maximum velocity        : %15.3f  [km/s]\n", (double)h->vmax);
    std::fprintf(stderr, "  Maximum frequency       : %15.3f  [Hz]\n\n", (double)h->fmax);
    std::fprintf(stderr, " ------------------------------------------------------------------------------\n\n");
    if (c > 1.0) return hfail("stability condition is violated (assert(c <= 1.0), m_report.f90:100-103)");
    return 0;
}

// add what the device stopwatches hold to the phase totals and restart them (they keep at most 4096 brackets)
int swpc3d_host::harvest_timers() {
    double ms = 0, n = 0;
    const char *key[3][2] = {{"ms_stress", "n_stress"}, {"ms_vel", "n_vel"}, {"ms_halo", "n_halo"}};
    double *dst[3] = {&tim_stress, &tim_vel, &tim_halo};
    for (int q = 0; q < 3; q++) {
        if (swpc3d_get_info(dev, key[q][0], &ms) || swpc3d_get_info(dev, key[q][1], &n)) return hfail(std::string("device: ") + swpc3d_last_error());
        *dst[q] += ms * n * 1e-3;
    }
    if (swpc3d_set_option(dev, "kernel_timing", 1)) return hfail(std::string("device: ") + swpc3d_last_error());
    return 0;
}

int swpc3d_host_run(swpc3d_host *h, int32_t it0, int32_t it1, int32_t verbose, float *vm, int32_t nvm, int32_t *nrec) {
    if (!h || !h->dev) return hfail("swpc3d_host_run: no device attached");
    int rec = 0;
    if (h->stopwatch_mode && swpc3d_set_option(h->dev, "kernel_timing", 1)) return hfail(std::string("device: ") + swpc3d_last_error());
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = it0; it <= it1; it++) {
        if (h->ntdec_r > 0 && it % h->ntdec_r == 0) {   // report__progress m_report.f90:120-185
            float v[3];
            if (swpc3d_vmax_global(h->dev, v)) return hfail(std::string("device: ") + swpc3d_last_error());
            const float mx = std::max(v[0], std::max(v[1], v[2]));
            if (mx * h->UC > 1e5f) return hfail("numerical divergence detected (m_report.f90:144-151)");
            for (int q = 0; q < 3; q++) v[q] = v[q] * h->UC * h->M0;
            if (vm && rec < nvm) { vm[3 * rec] = v[0]; vm[3 * rec + 1] = v[1]; vm[3 * rec + 2] = v[2]; }
            rec++;
            if (verbose && h->myid == 0) {
                const double tt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                double etas = (double)(h->nt - it) / (double)std::max(1, it - it0 + 1) * tt;
                const int eh = (int)(etas / 3600); etas -= eh * 3600.0;
                const int em = (int)(etas / 60); etas -= em * 60.0;
                std::fprintf(stderr, "  it=%07d,%6.3f s/loop, eta %03d:%02d:%02d, (%9.2E %9.2E %9.2E )\n", it, tt / std::max(1, it - it0 + 1), eh, em,
                             (int)etas, (double)v[0], (double)v[1], (double)v[2]);
            }
        }
        if (h->snap && h->snap->any && h->snap->opened) {   // main.f90:123-138 with snap__write between wav__store and the sweeps
#define DS(call) if (call) return hfail(std::string("device: ") + swpc3d_last_error());
            DS(swpc3d_green_store(h->dev, it));
            DS(swpc3d_wav_store(h->dev, it));
            if (h->snap_write(it)) return 1;
            DS(swpc3d_advance(h->dev, it));
#undef DS
        } else if (swpc3d_step(h->dev, it)) return hfail(std::string("device: ") + swpc3d_last_error());
        // wav__store's tail (m_wav.f90:619-621): the traces sampled so far, written while the run goes on (the buffers change
        // only at the next wav__store, so writing after the sweeps of this iteration gives the reference's files)
        if (h->ntdec_w_prg > 0 && (it - 1) % h->ntdec_w_prg == 0) {
            int32_t nf = 0;
            if (swpc3d_host_write_sac(h, nullptr, &nf)) return 1;
        }
        if (h->stopwatch_mode && (it - it0 + 1) % 1000 == 0 && h->harvest_timers()) return 1;
    }
    if (h->stopwatch_mode && h->harvest_timers()) return 1;
    if (swpc3d_sync(h->dev)) return hfail(std::string("device: ") + swpc3d_last_error());
    h->loop_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (nrec) *nrec = rec;
    return 0;
}

// pwatch__report (m_pwatch.f90:146-195) for this rank: <odir>/<title>.tim in the reference's table layout.  The rows are the
// phases the device stopwatches bracket -- the fused sweeps carry the reference's kernel__update_* AND absorb__update_* time,
// the exchange its global__comm_* time -- plus everything else of the time loop (sources, stations, snapshots, host).
static void mkdirs(const std::string &p);
int swpc3d_host_write_tim(swpc3d_host *h, const char *odir) {
    if (!h) return hfail("null handle");
    if (!h->stopwatch_mode) return 0;
    const std::string dir = odir ? odir : h->odir.c_str();
    mkdirs(dir);
    const std::string fn = dir + "/" + h->title + ".tim";
    FILE *fp = std::fopen(fn.c_str(), "w");
    if (!fp) return hfail("cannot write " + fn);
    const char *name[4] = {"kernel__update_stress", "kernel__update_vel", "global__comm", "others"};
    // With neighbours the exchange runs on its own stream BESIDE the core sweeps (boundary-first overlap): its stopwatch time is
    // not serial time, so it is reported but not subtracted from the loop -- "others" is what the two sweeps leave of the wall
    // clock (sources, stations, snapshots, host), and may be slightly negative when the stopwatches' brackets overlap.
    double ov = 0.0;
    if (h->dev && swpc3d_get_info(h->dev, "overlapped", &ov)) ov = 0.0;
    const double other = h->loop_seconds - h->tim_stress - h->tim_vel - (ov != 0.0 ? 0.0 : h->tim_halo);
    const double t[4] = {h->tim_stress, h->tim_vel, h->tim_halo, other};
    const double tsum = std::max(t[0] + t[1] + t[2] + t[3], 1e-30);
    std::fprintf(fp, "#   CPU     #ID       Procedure Name         Real Time[s]   Total Time[s]   Occupancy[%%]  Total Occp.[%%] \n");
    std::fprintf(fp, "# -------+-------+-------------------------+--------------+--------------+--------------+----------------\n");
    double acc = 0, racc = 0;
    for (int i = 0; i < 4; i++) {   // '(I8.5,I8.5,"    ",A22,4F15.3)'
        acc += t[i];
        racc += t[i] / tsum * 100.0;
        std::fprintf(fp, "   %05d   %05d    %22s%15.3f%15.3f%15.3f%15.3f\n", h->myid, i + 1, name[i], t[i], acc, t[i] / tsum * 100.0, racc);
    }
    std::fclose(fp);
    return 0;
}

// SAC writer: m_sac.f90:314-449 (header), m_wav.f90:346-393 (values), :782-792 (file names)
static void put_chars(char *dst, const std::string &s, int n) {
    for (int q = 0; q < n; q++) dst[q] = q < (int)s.size() ? s[q] : ' ';
}
static void mkdirs(const std::string &p) {
    for (size_t q = 1; q <= p.size(); q++)
        if (q == p.size() || p[q] == '/') mkdir(p.substr(0, q).c_str(), 0777);
}
// wav__stquery + the broadcast of m_green.f90:161-183, for hosts that run several ranks
int swpc3d_host_green_query(swpc3d_host *h, int32_t *found, int32_t ijk[3], float xyz[3], float lonlat[2]) {
    if (!h || !found) return hfail("null argument");
    if (!h->green_mode) return hfail("swpc3d_host_green_query: green_mode is off");
    *found = h->green.have_src ? 1 : 0;
    if (h->green.have_src) {
        for (int q = 0; q < 3; q++) { if (ijk) ijk[q] = h->green.src_ijk[q]; if (xyz) xyz[q] = h->green.src_xyz[q]; }
        if (lonlat) { lonlat[0] = h->green.evlo0; lonlat[1] = h->green.evla0; }
    }
    return 0;
}
int swpc3d_host_green_set_source(swpc3d_host *h, const int32_t ijk[3], const float xyz[3], const float lonlat[2]) {
    if (!h || !ijk || !xyz || !lonlat) return hfail("null argument");
    if (!h->green_mode) return hfail("swpc3d_host_green_set_source: green_mode is off");
    if (h->green.finalized) return hfail("swpc3d_host_green_set_source: the grid-point list was already read");
    for (int q = 0; q < 3; q++) { h->green.src_ijk[q] = ijk[q]; h->green.src_xyz[q] = xyz[q]; }
    h->green.evlo0 = lonlat[0]; h->green.evla0 = lonlat[1];
    h->green.have_src = true;
    return h->green_finalize();
}

// green__export, m_green.f90:553-604 (wav_format sac | csf | wav).  Headers: green__setup :282-349 on sac__init defaults.
int swpc3d_host_write_green(swpc3d_host *h, const char *odir, int32_t *nfiles) {
    if (!h) return hfail("null handle");
    if (nfiles) *nfiles = 0;
    if (!h->green_mode) return 0;
    if (!h->dev) return hfail("swpc3d_host_write_green: no device attached");
    swpc3d_host::Green &g = h->green;
    const int ng = (int)g.gid.size(), ncmp = g.ncmp, ntw = g.ntw;
    g.gf.assign((size_t)ntw * ncmp * std::max(ng, 1), 0.0f);
    if (ng > 0 && swpc3d_get_green(h->dev, g.gf.data())) return hfail(std::string("device: ") + swpc3d_last_error());
    if (g.cmp == 'z')   // positive upward for the z component (:563-565)
        for (auto &v : g.gf) v = -v;
    const std::string dir = std::string(odir ? odir : h->odir.c_str()) + "/green/" + g.stnm;
    mkdirs(dir);
    static const char *cmpn[9] = {"mxx", "myy", "mzz", "myz", "mxz", "mxy", "fx_", "fy_", "fz_"};
    const time_t tt = (time_t)h->exedate + (time_t)h->tz_minutes * 60;
    struct tm gm;
    gmtime_r(&tt, &gm);
    const float fdx = (float)h->dx, fdy = (float)h->dy, fdz = (float)h->dz;
    auto header = [&](int i, int j, unsigned char *out) {
        float f[70];
        int32_t iv[35], lv[5] = {1, 0, 1, 0, 0};
        char a[192];
        std::fill(f, f + 70, -12345.0f);
        std::fill(iv, iv + 35, -12345);
        for (int q = 0; q < 24; q++) put_chars(a + 8 * q, "-12345", 8);
        put_chars(a + 8, "-12345", 16);
        const double delta = (double)(g.ntdec_w * h->dt);
        f[0] = (float)((int)(delta * 1e7)) / 1e7f;
        f[5] = h->tbeg;
        f[31] = g.evla0; f[32] = g.evlo0; f[34] = g.src_xyz[2] * 1000;
        f[35] = g.lat[i]; f[36] = g.lon[i]; f[38] = g.zg[i];
        f[40] = g.xg[i]; f[41] = g.yg[i]; f[42] = g.zg[i];
        f[43] = i2x(g.ijk[3 * i], h->xbeg, fdx); f[44] = i2x(g.ijk[3 * i + 1], h->ybeg, fdy); f[45] = i2x(g.ijk[3 * i + 2], h->zbeg, fdz);
        f[46] = h->clon; f[47] = h->clat; f[48] = h->phi;
        f[58] = g.cmp == 'z' ? 0.0f : 90.0f;
        f[57] = g.cmp == 'x' ? 0.0f + h->phi : (g.cmp == 'y' ? 90.0f + h->phi : 0.0f);
        iv[0] = gm.tm_year + 1900; iv[1] = gm.tm_yday + 1; iv[2] = gm.tm_hour; iv[3] = gm.tm_min; iv[4] = gm.tm_sec; iv[5] = 0;
        iv[6] = 6; iv[9] = ntw; iv[15] = 1; iv[16] = j < 6 ? 7 : 6;
        char cid[16];
        std::snprintf(cid, sizeof(cid), "%08d", g.gid[i]);
        put_chars(a, g.stnm, 8);
        put_chars(a + 8, cid, 16);
        put_chars(a + 160, std::string("G_V") + g.cmp + "_" + cmpn[j], 8);
        std::memcpy(out, f, 280); std::memcpy(out + 280, iv, 140); std::memcpy(out + 420, lv, 20); std::memcpy(out + 440, a, 192);
    };
    int count = 0;
    if (g.wav_format == "sac") {
        std::vector<unsigned char> rec(632 + 4 * (size_t)ntw);
        for (int i = 0; i < ng; i++)
            for (int j = 0; j < ncmp; j++) {
                header(i, j, rec.data());
                std::memcpy(rec.data() + 632, g.gf.data() + (size_t)ntw * ((size_t)ncmp * i + j), 4 * (size_t)ntw);
                char cid[16];
                std::snprintf(cid, sizeof(cid), "%08d", g.gid[i]);
                const std::string fn = dir + "/" + h->title + "__" + cid + "__" + g.stnm + "__" + g.cmp + "__" + cmpn[j] + "__.sac";
                std::ofstream os(fn, std::ios::binary);
                if (!os) return hfail("cannot write " + fn);
                os.write((const char *)rec.data(), (std::streamsize)rec.size());
                count++;
            }
    } else if ((g.wav_format == "csf" || g.wav_format == "wav") && ng > 0) {
        char cmyid[16];
        std::snprintf(cmyid, sizeof(cmyid), "%06d", h->myid);
        const std::string fn = dir + "/" + h->title + "__" + g.stnm + "__" + g.cmp + "__" + cmyid + "__." + g.wav_format;
        std::ofstream os(fn, std::ios::binary);
        if (!os) return hfail("cannot write " + fn);
        const int32_t ntr = ng * ncmp, npts = ntw;
        std::vector<unsigned char> hd(632);
        if (g.wav_format == "csf") {   // csf__write, m_sac.f90:585-648: 'CSFD', ntrace, npts, then header + data per trace
            os.write("CSFD", 4);
            os.write((const char *)&ntr, 4);
            os.write((const char *)&npts, 4);
            for (int i = 0; i < ng; i++)
                for (int j = 0; j < ncmp; j++) {
                    header(i, j, hd.data());
                    os.write((const char *)hd.data(), 632);
                    os.write((const char *)(g.gf.data() + (size_t)ntw * ((size_t)ncmp * i + j)), 4 * (std::streamsize)ntw);
                }
        } else {
            return hfail("wav_format = 'wav' of green__export writes the compiler's in-memory sac__hdr records (m_green.f90:594-596): not portable, not supported");
        }
        count = 1;
    }
    if (nfiles) *nfiles = count;
    return 0;
}

int swpc3d_host_write_sac(swpc3d_host *h, const char *odir, int32_t *nfiles) {
    if (!h) return hfail("null handle");
    if (nfiles) *nfiles = 0;
    const int nst = (int)(h->st_ijk.size() / 3);
    const bool sw[4] = {h->sw_wav_v, h->sw_wav_u, h->sw_wav_stress, h->sw_wav_strain};
    if (!(sw[0] || sw[1] || sw[2] || sw[3]) || nst == 0 || h->ntw <= 0) return 0;
    if (!h->dev) return hfail("swpc3d_host_write_sac: no device attached");
    // wav_format other than sac / csf / tar_st / tar_node: wav__write has no branch for it and writes nothing (m_wav.f90:677-763)
    const std::string dir = std::string(odir ? odir : h->odir.c_str()) + "/wav";
    mkdirs(dir);
    static const char *cmpnm[4][6] = {{"Vx", "Vy", "Vz", "", "", ""}, {"Ux", "Uy", "Uz", "", "", ""},
                                      {"Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy"}, {"Exx", "Eyy", "Ezz", "Eyz", "Exz", "Exy"}};
    const time_t tt = (time_t)h->exedate + (time_t)h->tz_minutes * 60;   // daytim__localtime m_daytim.f90:262-264
    struct tm g;
    gmtime_r(&tt, &g);
    int count = 0;
    std::vector<float> buf;
    // `update self(wav_*)` m_wav.f90:672-675, one product at a time
    for (int prod = 0; prod < 4; prod++) {
        if (!sw[prod]) continue;
        const int ncmp = prod < 2 ? 3 : 6;
        buf.assign((size_t)h->ntw * ncmp * nst, 0.0f);
        if (swpc3d_get_wav_product(h->dev, prod, buf.data())) return hfail(std::string("device: ") + swpc3d_last_error());
        if (prod == 0) h->wav = buf;
        h->wav_all[prod] = buf;
    }
    // one SAC record = 632-byte header (sac__whdr) + ntw samples
    auto sac_record = [&](int s, int prod, int c, std::vector<unsigned char> &out) {
        const int ncmp = prod < 2 ? 3 : 6;
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
        if (h->bf_mode) { f[40] = h->f0[0]; f[41] = h->f0[1]; f[42] = h->f0[2]; }
        else for (int q = 0; q < 6; q++) f[40 + q] = h->m0ij[q];
        f[46] = h->clon; f[47] = h->clat; f[48] = h->phi;
        const float ddx = h->sx0 - h->xst[s], ddy = h->sy0 - h->yst[s];
        f[50] = std::sqrt(ddx * ddx + ddy * ddy);
        f[51] = rad2deg_s(std::atan2(h->yst[s] - h->sy0, h->xst[s] - h->sx0));
        f[52] = rad2deg_s(std::atan2(h->sy0 - h->yst[s], h->sx0 - h->xst[s]));
        if (prod < 2) {   // cmpaz / cmpinc only for the vector products (m_wav.f90:286-288, :297-299)
            f[57] = c == 0 ? 0.0f + h->phi : (c == 1 ? 90.0f + h->phi : 0.0f);
            f[58] = 90.0f;
        }
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
    // collect the records in the order of m_wav.f90:679-705 (per station: v, u, stress, strain), then write the containers
    std::vector<WavStation> sts((size_t)nst);
    for (int s = 0; s < nst; s++) {
        sts[(size_t)s].stnm = h->stnm[s];
        for (int prod = 0; prod < 4; prod++) {
            if (!sw[prod]) continue;
            for (int c = 0; c < (prod < 2 ? 3 : 6); c++) {
                WavTrace t; t.prod = prod; t.cmp = cmpnm[prod][c];
                sac_record(s, prod, c, t.rec);
                sts[(size_t)s].tr.push_back(std::move(t));
            }
        }
    }
    std::string err;
    count = write_wav_files(h->wav_format, dir, h->title, "3d", true, h->myid, h->exedate, h->ntw, sts, err);
    if (count < 0) return hfail(err);
    if (nfiles) *nfiles = count;
    return 0;
}

}   // extern "C"
