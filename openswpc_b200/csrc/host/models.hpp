// models.hpp -- host-side builders for the heavier velocity models of swpc_3d (setup-only CPU code):
//   vmodel_lgm        src/swpc_3d/m_vmodel_lgm.f90:20-175        linear-gradient layers
//   vmodel_uni_rmed   src/swpc_3d/m_vmodel_uni_rmed.f90:22-170   homogeneous + random-media volume
//   vmodel_lhm_rmed   src/swpc_3d/m_vmodel_lhm_rmed.f90:22-244   layered + one random-media volume per layer
//   vmodel_lgm_rmed   src/swpc_3d/m_vmodel_lgm_rmed.f90:22-258   gradient layers + random media (with the reference's
//                                                                whole-plane assignment quirk, :236-240)
//   rdrmed__3d        src/shared/m_rdrmed.f90:73-134             cyclic read of the volume written by gen_rmed3d
//   stabilize_absorber src/swpc_3d/m_medium.f90:273-337
// The image has no netCDF library; gen_rmed3d creates its file with NF90_CLOBBER (tools/gen_rmed3d.f90:91), i.e. the
// netCDF classic format, which ClassicNc below parses directly.  vmodel_grd / grd_rmed read GMT grids: those are accepted in
// the classic container (`gmt grdconvert in.grd out.grd=cf`, `nccopy -k classic`); netCDF-4/HDF5 grids are refused with a
// message.  vmodel_user is a compile-time plug-in of the reference and stays out of this build.
//   vmodel_grd        src/swpc_3d/m_vmodel_grd.f90:28-303 + src/shared/m_bicubic.f90
//   vmodel_grd_rmed   src/swpc_3d/m_vmodel_grd_rmed.f90:28-388
#pragma once

#include "common.hpp"

namespace {

// the memory box of one rank and the medium arrays being built, (k,i,j) with k fastest as m_medium.f90:441-452
struct MediumBox {
    int ib, ie, jb, je, kb, ke;          // ibeg_m..iend_m, jbeg_m..jend_m, kbeg_m..kend_m
    const float *zc;                     // zc(kb:ke)
    float *rho, *lam, *mu, *qp, *qs;     // the reference passes taup / taus as the Qp / Qs outputs
    int nk() const { return ke - kb + 1; }
    int ni() const { return ie - ib + 1; }
    size_t ncell() const { return (size_t)nk() * ni() * (size_t)(je - jb + 1); }
    size_t at(int k, int i, int j) const { return (size_t)(k - kb) + (size_t)nk() * ((size_t)(i - ib) + (size_t)ni() * (size_t)(j - jb)); }
};

// netCDF classic (CDF-1 / CDF-2) header: dimension lengths and, per variable, type and data offset
class ClassicNc {
  public:
    std::vector<long long> dim;
    struct Var { std::string name; int type = 0; long long begin = 0; std::vector<int> dimid; };
    std::vector<Var> var;
    std::vector<unsigned char> bytes;

    bool open(const std::string &path, std::string &err) {
        std::ifstream is(path, std::ios::binary);
        if (!is) { err = "cannot open " + path; return false; }
        bytes.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
        if (bytes.size() < 32 || bytes[0] != 'C' || bytes[1] != 'D' || bytes[2] != 'F' || (bytes[3] != 1 && bytes[3] != 2)) {
            err = path + " is not a netCDF classic (CDF-1/2) file";
            return false;
        }
        const int offw = bytes[3] == 2 ? 8 : 4;
        pos_ = 8;   // magic, numrecs
        static const int tsz[7] = {0, 1, 1, 2, 4, 4, 8};
        auto skip_atts = [&]() {
            const unsigned tag = u4(), n = u4();
            if (tag != 0x0C) return;
            for (unsigned a = 0; a < n; a++) {
                name();
                const unsigned ty = u4(), ne = u4();
                pos_ += ((size_t)ne * tsz[ty < 7 ? ty : 0] + 3) / 4 * 4;
            }
        };
        unsigned tag = u4(), n = u4();
        if (tag == 0x0A)
            for (unsigned d = 0; d < n; d++) { name(); dim.push_back(u4()); }
        skip_atts();
        tag = u4(); n = u4();
        if (tag == 0x0B)
            for (unsigned v = 0; v < n; v++) {
                Var x;
                x.name = name();
                const unsigned nd = u4();
                for (unsigned d = 0; d < nd; d++) x.dimid.push_back((int)u4());
                skip_atts();
                x.type = (int)u4();
                u4();   // vsize
                x.begin = (long long)ube(offw);
                var.push_back(x);
            }
        return true;
    }
    // element `idx` of variable v as a double, for the numeric netCDF types a GMT grid may use
    double value(int v, long long idx) const {
        const unsigned char *p = bytes.data() + var[(size_t)v].begin;
        auto be = [&](int n) { unsigned long long u = 0; for (int q = 0; q < n; q++) u = (u << 8) | p[(size_t)idx * n + q]; return u; };
        switch (var[(size_t)v].type) {
        case 5: { const uint32_t u = (uint32_t)be(4); float x; std::memcpy(&x, &u, 4); return (double)x; }
        case 6: { const uint64_t u = be(8); double x; std::memcpy(&x, &u, 8); return x; }
        case 4: return (double)(int32_t)be(4);
        case 3: return (double)(int16_t)be(2);
        default: return 0.0;
        }
    }
    float f32(long long byte_off) const {
        const unsigned char *p = bytes.data() + byte_off;
        const uint32_t u = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        float v;
        std::memcpy(&v, &u, 4);
        return v;
    }

  private:
    size_t pos_ = 0;
    unsigned long long ube(int n) {
        unsigned long long v = 0;
        for (int q = 0; q < n; q++) v = (v << 8) | bytes[pos_ + q];
        pos_ += n;
        return v;
    }
    unsigned u4() { return (unsigned)ube(4); }
    std::string name() {
        const unsigned n = u4();
        std::string s((const char *)bytes.data() + pos_, n);
        pos_ += (n + 3) / 4 * 4;
        return s;
    }
};

// m_rdrmed.f90:73-134: the volume is periodic in x and y, wraps upward for k <= 0 and repeats below its last plane
inline int rdrmed3d(const MediumBox &b, const std::string &fn, float *vol) {
    ClassicNc nc;
    std::string err;
    if (!nc.open(fn, err)) return hfail("rdrmed__3d: " + err);
    if (nc.dim.size() < 3 || nc.var.size() < 4 || nc.var[3].type != 5)
        return hfail("rdrmed__3d: " + fn + " needs dimensions x, y, z and a float volume as its 4th variable (gen_rmed3d.f90:91-126)");
    const long long nxc = nc.dim[0], nyc = nc.dim[1], nzc = nc.dim[2], beg = nc.var[3].begin;
    if (beg + 4 * nxc * nyc * nzc > (long long)nc.bytes.size()) return hfail("rdrmed__3d: " + fn + " is truncated");
    auto wrap = [](long long v, long long n) { long long r = v % n; return r <= 0 ? r + n : r; };
    // (k innermost: contiguous writes; the reads stride through the volume, which is small and cache-resident)
    const int ktop = (int)std::min<long long>(b.ke, nzc);
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = b.jb; j <= b.je; j++)
        for (int i = b.ib; i <= b.ie; i++) {
            const long long col = beg + 4 * (nxc * (wrap(j, nyc) - 1) + (wrap(i, nxc) - 1));
            float *v = vol + b.at(b.kb, i, j);
            for (int k = b.kb; k <= ktop; k++) v[k - b.kb] = nc.f32(col + 4 * nxc * nyc * ((k <= 0 ? k + nzc : k) - 1));
            for (long long k = nzc + 1; k <= b.ke; k++) v[k - b.kb] = v[(k % nzc) - b.kb];
        }
    return 0;
}

// m_rdrmed.f90:19-70 (swpc_psv): the section of tools/gen_rmed2d.f90 (dimensions x, z; 3rd variable, x fastest), periodic in
// x, wrapped upward for k <= 0 and repeated below its last row.  `b` is a one-plane box (jb == je).
inline int rdrmed2d(const MediumBox &b, const std::string &fn, float *vol) {
    ClassicNc nc;
    std::string err;
    if (!nc.open(fn, err)) return hfail("rdrmed__2d: " + err);
    if (nc.dim.size() < 2 || nc.var.size() < 3 || nc.var[2].type != 5)
        return hfail("rdrmed__2d: " + fn + " needs dimensions x, z and a float section as its 3rd variable (gen_rmed2d.f90:87-123)");
    const long long nxc = nc.dim[0], nzc = nc.dim[1], beg = nc.var[2].begin;
    if (beg + 4 * nxc * nzc > (long long)nc.bytes.size()) return hfail("rdrmed__2d: " + fn + " is truncated");
    auto wrap = [](long long v, long long n) { long long r = v % n; return r <= 0 ? r + n : r; };
    const int ktop = (int)std::min<long long>(b.ke, nzc);
#pragma omp parallel for schedule(static)
    for (int i = b.ib; i <= b.ie; i++) {
        const long long col = beg + 4 * (wrap(i, nxc) - 1);
        float *v = vol + b.at(b.kb, i, b.jb);
        for (int k = b.kb; k <= ktop; k++) v[k - b.kb] = nc.f32(col + 4 * nxc * ((k <= 0 ? k + nzc : k) - 1));
        for (long long k = nzc + 1; k <= b.ke; k++) v[k - b.kb] = v[wrap(k, nzc) - b.kb];
    }
    return 0;
}

struct LayerTable {
    std::vector<float> depth, rho, vp, vs, qp, qs;
    std::vector<std::string> rmed;
    int n() const { return (int)depth.size(); }
};

// "depth rho vp vs Qp Qs [random-media file]" rows, '#' comments; then the velocity cut-off (m_vmodel_lgm.f90:92-101)
inline int read_layer_table(const std::string &fn, bool with_rmed, float vcut, LayerTable &t) {
    std::ifstream is(fn);
    if (!is) return hfail("layer file " + fn + " not found (assert, m_vmodel_lgm.f90:56-57)");
    std::string line;
    while (std::getline(is, line)) {
        if (blank_or_comment(line)) continue;
        const std::vector<float> v = parse_reals(line);
        if (v.size() < 6) continue;
        t.depth.push_back(v[0]); t.rho.push_back(v[1]); t.vp.push_back(v[2]); t.vs.push_back(v[3]); t.qp.push_back(v[4]); t.qs.push_back(v[5]);
        std::string name;
        if (with_rmed) {   // 7th list-directed item: a (possibly quoted) file name
            std::istringstream ss(line);
            std::string tok;
            for (int q = 0; q < 7 && (ss >> tok); q++)
                if (q == 6) name = tok;
            for (auto &ch : name) if (ch == ',') ch = ' ';
            while (!name.empty() && name.back() == ' ') name.pop_back();
            if (name.size() >= 2 && (name[0] == '\'' || name[0] == '"') && name.back() == name[0]) name = name.substr(1, name.size() - 2);
        }
        t.rmed.push_back(name);
    }
    if (t.n() == 0) return hfail("no layer in " + fn);
    for (int l = t.n() - 2; l >= 0; l--)
        if ((t.vp[l] < vcut || t.vs[l] < vcut) && (t.vp[l] > 0 && t.vs[l] > 0)) {
            t.vp[l] = t.vp[l + 1]; t.vs[l] = t.vs[l + 1]; t.rho[l] = t.rho[l + 1]; t.qp[l] = t.qp[l + 1]; t.qs[l] = t.qs[l + 1];
        }
    return 0;
}

// m_fdtool.f90:15-48
inline void vcheck(float &vp, float &vs, float &rho, float xi, float vmin, float vmax, float rhomin) {
    float gamma = vp / vs;
    if (gamma < EPS_SP) gamma = std::sqrt(3.0f);
    if (vp > vmax || vs > vmax) {
        const float xi2 = (1 + xi) * vmax / vp - 1;
        vp = vmax;
        vs = vmax / gamma;
        rho = rho * (1 + 0.8f * xi2) / (1 + 0.8f * xi);
    }
    if (vp < vmin || vs < vmin) { vs = vmin; vp = vmin * gamma; }
    if (rho < rhomin) rho = rhomin;
}

struct ModelEnv {
    const IniFile *ini;
    std::string base;
    float vcut, dt;
    double dx, dy, dz;
    bool munk, flatten;
    bool psv = false;   // swpc_psv flavour of the builders: the y axis removed and a few differing expressions (cited at each use)
    // spherical depth / velocity scaling of the earth-flattening transformation (m_vmodel_lgm.f90:64-73)
    void depth(float zc, float &zs, float &cv) const {
        if (flatten) { zs = (float)(R_EARTH - R_EARTH * std::exp(-(double)zc / R_EARTH)); cv = (float)std::exp((double)zc / R_EARTH); }
        else { zs = zc; cv = 1.0f; }
    }
    float rmed_vmax() const {   // cc * dh / dt, m_vmodel_uni_rmed.f90:79-83
        const float dh = psv ? (float)(1.0 / std::sqrt(1.0 / (dx * dx) + 1.0 / (dz * dz)))   // swpc_psv/m_vmodel_uni_rmed.f90:80
                             : (float)(1.0 / std::sqrt(1.0 / (dx * dx) + 1.0 / (dy * dy) + 1.0 / (dz * dz)));
        return (6.0f / 7.0f) * dh / dt;
    }
};

inline void set_cell(const MediumBox &b, size_t n, float rho, float vp, float vs, float qp, float qs) {
    b.rho[n] = rho; b.mu[n] = rho * vs * vs; b.lam[n] = rho * (vp * vp - 2 * vs * vs); b.qp[n] = qp; b.qs[n] = qs;
}

// The reference's builders loop over k outermost; k is the FASTEST index in memory, so that order strides through all
// five arrays.  Here every builder first decides, per plane k, whether the plane is laterally uniform (air, ocean, 1-D
// models) or needs a per-cell evaluation, and the arrays are then filled column by column (k innermost, contiguous).
struct PlanePlan {
    bool uniform = true;
    float rho = 0, mu = 0, lam = 0, qp = 0, qs = 0;   // uniform planes
    int layer = -1;                                   // per-cell planes: what the cell function needs
    float cv = 1.0f;
    void set(float r, float vp, float vs, float a, float b_) { rho = r; mu = r * vs * vs; lam = r * (vp * vp - 2 * vs * vs); qp = a; qs = b_; }
};
template <typename CellFn>
inline void fill_columns(const MediumBox &b, const std::vector<PlanePlan> &pl, CellFn &&cell) {
    const int nk = b.nk();
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = b.jb; j <= b.je; j++)
        for (int i = b.ib; i <= b.ie; i++) {
            const size_t n0 = b.at(b.kb, i, j);
            for (int q = 0; q < nk; q++) {
                const PlanePlan &p = pl[(size_t)q];
                const size_t n = n0 + (size_t)q;
                if (p.uniform) { b.rho[n] = p.rho; b.mu[n] = p.mu; b.lam[n] = p.lam; b.qp[n] = p.qp; b.qs[n] = p.qs; }
                else cell(p, n);
            }
        }
}

// air / ocean plane above the first interface, shared by the lhm / lgm families; returns false inside the solid
inline bool air_or_ocean(const ModelEnv &e, float zc, float zs, float cv, float top, float &rho, float &vp, float &vs, float &qp, float &qs) {
    if (!(zs < top)) return false;
    if (zs < 0.0f) { rho = 0.001f; vp = 0.0f; vs = 0.0f; qp = 10.0f; qs = 10.0f; }
    else { rho = 1.0f; vp = cv * seawater_vel(zc, e.munk); vs = 0.0f; qp = 1000000.0f; qs = 1000000.0f; }
    return true;
}

// linear interpolation inside layer l (m_vmodel_lgm.f90:137-141)
inline void gradient_at(const LayerTable &t, int l, float zs, float cv, float &rho, float &vp, float &vs, float &qp, float &qs, bool psv_order = false) {
    const float dd = t.depth[l + 1] - t.depth[l], dz = zs - t.depth[l];
    if (psv_order) {   // swpc_psv/m_vmodel_lgm.f90:128-132 multiplies before it divides: (b - a) * dz / dd
        rho = t.rho[l] + (t.rho[l + 1] - t.rho[l]) * dz / dd;
        vp = cv * (t.vp[l] + (t.vp[l + 1] - t.vp[l]) * dz / dd);
        vs = cv * (t.vs[l] + (t.vs[l + 1] - t.vs[l]) * dz / dd);
        qp = t.qp[l] + (t.qp[l + 1] - t.qp[l]) * dz / dd;
        qs = t.qs[l] + (t.qs[l + 1] - t.qs[l]) * dz / dd;
        return;
    }
    rho = t.rho[l] + (t.rho[l + 1] - t.rho[l]) / dd * dz;
    vp = cv * (t.vp[l] + (t.vp[l + 1] - t.vp[l]) / dd * dz);
    vs = cv * (t.vs[l] + (t.vs[l + 1] - t.vs[l]) / dd * dz);
    qp = t.qp[l] + (t.qp[l + 1] - t.qp[l]) / dd * dz;
    qs = t.qs[l] + (t.qs[l + 1] - t.qs[l]) / dd * dz;
}

inline int vmodel_lgm(const ModelEnv &e, const MediumBox &b, float &bd0) {
    LayerTable t;
    if (read_layer_table(join_path(e.base, e.ini->get("fn_lhm", "")), false, e.vcut, t)) return 1;
    const int nl = t.n();
    bd0 = t.depth[0];
    std::vector<PlanePlan> pl((size_t)b.nk());
    for (int k = b.kb; k <= b.ke; k++) {
        const float zc = b.zc[k - b.kb];
        float zs, cv, rho, vp, vs, qp, qs;
        e.depth(zc, zs, cv);
        // swpc_psv/m_vmodel_lgm.f90:110 evaluates the sea-water profile at the spherical depth zs, the 3-D code at zc (:123)
        if (!air_or_ocean(e, e.psv ? zs : zc, zs, cv, t.depth[0], rho, vp, vs, qp, qs)) {
            rho = t.rho[nl - 1]; vp = cv * t.vp[nl - 1]; vs = cv * t.vs[nl - 1]; qp = t.qp[nl - 1]; qs = t.qs[nl - 1];
            for (int l = 0; l + 1 < nl; l++)
                if (t.depth[l] <= zs && zs < t.depth[l + 1]) { gradient_at(t, l, zs, cv, rho, vp, vs, qp, qs, e.psv); break; }
        }
        pl[(size_t)(k - b.kb)].set(rho, vp, vs, qp, qs);
    }
    fill_columns(b, pl, [](const PlanePlan &, size_t) {});
    return 0;
}

// the distinct random-media files of a layer table (independent_list, m_fdtool.f90:815-855) read over the memory box;
// a missing file means "no perturbation" (m_vmodel_lhm_rmed.f90:134-140)
inline int read_rmed_set(const ModelEnv &e, const MediumBox &b, const LayerTable &t, std::vector<int> &tbl, std::vector<std::vector<float>> &xi) {
    const std::string dir = e.ini->get("dir_rmed", "");
    std::vector<std::string> uniq;
    tbl.assign(t.n(), 0);
    for (int l = 0; l < t.n(); l++) {
        const std::string full = dir + "/" + t.rmed[l];
        const auto it = std::find(uniq.begin(), uniq.end(), full);
        tbl[l] = (int)(it - uniq.begin());
        if (it == uniq.end()) uniq.push_back(full);
    }
    xi.assign(uniq.size(), std::vector<float>());
    for (size_t q = 0; q < uniq.size(); q++) {
        xi[q].assign(b.ncell(), 0.0f);
        const std::string path = join_path(e.base, uniq[q]);
        if (std::ifstream(path).good() && (e.psv ? rdrmed2d(b, path, xi[q].data()) : rdrmed3d(b, path, xi[q].data()))) return 1;
    }
    return 0;
}

inline int vmodel_uni_rmed(const ModelEnv &e, const MediumBox &b, float &bd0) {
    const IniFile &ini = *e.ini;
    const float vp0 = ini.get_s("vp0", 5.0f), vs0 = ini.get_s("vs0", vp0 / std::sqrt(3.0f)), rho0 = ini.get_s("rho0", 2.7f);
    const float qp0 = ini.get_s("qp0", 1000000.0f), qs0 = ini.get_s("qs0", 1000000.0f), topo0 = ini.get_s("topo0", 0.0f);
    const float rhomin = ini.get_s("rhomin", 1.0f), vmin = e.vcut, vmax = e.rmed_vmax();
    std::vector<float> xi(b.ncell(), 0.0f);
    const std::string path = join_path(e.base, ini.get("dir_rmed", "") + "/" + ini.get("fn_rmed0", ""));
    if (std::ifstream(path).good() && (e.psv ? rdrmed2d(b, path, xi.data()) : rdrmed3d(b, path, xi.data()))) return 1;
    bd0 = topo0;
    std::vector<PlanePlan> pl((size_t)b.nk());
    for (int k = b.kb; k <= b.ke; k++) {
        PlanePlan &p = pl[(size_t)(k - b.kb)];
        const float zc = b.zc[k - b.kb];
        float zs, cv;
        e.depth(zc, zs, cv);
        if (zs > topo0) { p.uniform = false; p.cv = cv; }
        else if (zs > 0.0f) p.set(1.0f, cv * seawater_vel(zc, e.munk), 0.0f, 1000000.0f, 1000000.0f);
        else p.set(0.001f, 0.0f, 0.0f, 10.0f, 10.0f);
    }
    const float *x = xi.data();
    fill_columns(b, pl, [&](const PlanePlan &p, size_t n) {
        float rho = (1.0f + 0.8f * x[n]) * rho0, vp = (1.0f + x[n]) * p.cv * vp0, vs = (1.0f + x[n]) * p.cv * vs0;
        vcheck(vp, vs, rho, x[n], vmin, vmax, rhomin);
        set_cell(b, n, rho, vp, vs, qp0, qs0);
    });
    return 0;
}

inline int vmodel_lhm_rmed(const ModelEnv &e, const MediumBox &b, float &bd0) {
    LayerTable t;
    if (read_layer_table(join_path(e.base, e.ini->get("fn_lhm_rmed", "")), true, e.vcut, t)) return 1;
    const float rhomin = e.ini->get_s("rhomin", 1.0f), vmin = e.vcut, vmax = e.rmed_vmax();
    std::vector<int> tbl;
    std::vector<std::vector<float>> xi;
    if (read_rmed_set(e, b, t, tbl, xi)) return 1;
    bd0 = t.depth[0];
    std::vector<PlanePlan> pl((size_t)b.nk());
    for (int k = b.kb; k <= b.ke; k++) {
        PlanePlan &p = pl[(size_t)(k - b.kb)];
        const float zc = b.zc[k - b.kb];
        float zs, cv, rho = 0, vp = 0, vs = 0, qp = 0, qs = 0;
        e.depth(zc, zs, cv);
        if (e.psv) zs = zc;   // swpc_psv/m_vmodel_lhm_rmed.f90:148-179 tests the grid depth where the 3-D code tests the spherical depth
        if (air_or_ocean(e, zc, zs, cv, t.depth[0], rho, vp, vs, qp, qs)) {
            p.set(rho, vp, vs, qp, qs);
            if (!(zs < 0.0f)) p.lam = 1.0f * vp * vp;   // m_vmodel_lhm_rmed.f90:176-178 writes lam = 1.0 * vp1 * vp1 directly
            continue;
        }
        // the reference walks all layers and lets every one with zs >= depth(l) overwrite rho1 .. qs1 (:196-209): only
        // the LAST such layer survives, and it depends on k alone
        for (int l = 0; l < t.n(); l++)
            if (zs >= t.depth[l]) p.layer = l;
        if (p.layer < 0) { p.set(rho, vp, vs, qp, qs); continue; }
        p.uniform = false;
        p.cv = cv;
    }
    std::vector<const float *> xl((size_t)t.n());
    for (int l = 0; l < t.n(); l++) xl[(size_t)l] = xi[tbl[l]].data();
    fill_columns(b, pl, [&](const PlanePlan &p, size_t n) {
        const int l = p.layer;
        const float x = xl[(size_t)l][n];
        float rho = t.rho[l] * (1 + 0.8f * x), vp = p.cv * t.vp[l] * (1 + x), vs = p.cv * t.vs[l] * (1 + x);
        if (t.vp[l] > 0 && t.vs[l] > 0) vcheck(vp, vs, rho, x, vmin, vmax, rhomin);
        set_cell(b, n, rho, vp, vs, t.qp[l], t.qs[l]);
    });
    return 0;
}

inline int vmodel_lgm_rmed(const ModelEnv &e, const MediumBox &b, float &bd0) {
    LayerTable t;
    if (read_layer_table(join_path(e.base, e.ini->get("fn_lhm_rmed", "")), true, e.vcut, t)) return 1;
    const float rhomin = e.ini->get_s("rhomin", 1.0f), vmin = e.vcut, vmax = e.rmed_vmax();
    std::vector<int> tbl;
    std::vector<std::vector<float>> xi;
    if (read_rmed_set(e, b, t, tbl, xi)) return 1;
    const int nl = t.n();
    bd0 = t.depth[0];
    std::vector<PlanePlan> pl((size_t)b.nk());
    for (int k = b.kb; k <= b.ke; k++) {
        const float zc = b.zc[k - b.kb];
        float zs, cv, rho, vp, vs, qp, qs;
        e.depth(zc, zs, cv);
        if (!air_or_ocean(e, zc, zs, cv, t.depth[0], rho, vp, vs, qp, qs)) {
            // the reference assigns the WHOLE plane inside its (i,j) loops (:236-240): what survives is the value
            // computed at the last column (iend_m, jend_m) of this rank's memory box
            const size_t n = b.at(k, b.ie, b.je);
            const float xl = xi[tbl[nl - 1]][n];
            rho = t.rho[nl - 1] * (1 + 0.8f * xl); vp = cv * t.vp[nl - 1] * (1 + xl); vs = cv * t.vs[nl - 1] * (1 + xl);
            qp = t.qp[nl - 1]; qs = t.qs[nl - 1];
            for (int l = 0; l + 1 < nl; l++)
                if (t.depth[l] <= zs && zs < t.depth[l + 1]) {
                    gradient_at(t, l, zs, cv, rho, vp, vs, qp, qs);
                    const float x = xi[tbl[l]][n];
                    rho = rho * (1 + 0.8f * x); vp = vp * (1 + x); vs = vs * (1 + x);
                    if (t.vp[l] > 0 && t.vs[l] > 0) vcheck(vp, vs, rho, x, vmin, vmax, rhomin);
                    break;
                }
        }
        pl[(size_t)(k - b.kb)].set(rho, vp, vs, qp, qs);
    }
    fill_columns(b, pl, [](const PlanePlan &, size_t) {});
    return 0;
}

// m_medium.f90:273-337.  `b` carries taup / taus in qp / qs here; vmax is the global maximum P velocity.
inline void stabilize_absorber(const MediumBox &b, const int *kbeg_a /* (i,j) over the memory box */, int ibeg, int iend, int jbeg, int jend,
                               int kend, float vmax) {
    const float vmin_pml = vmax * 0.4f;   // V_DYNAMIC_RANGE
    const int LV_THICK = 20;
    // (a swpc_psv section is a one-plane box: jbeg == jend == b.jb, and its stencil has no y extent, swpc_psv/m_medium.f90:323)
    const int wy = b.jb == b.je ? 0 : 2, ey = b.jb == b.je ? 0 : 1;
    auto top_of = [&](int i, int j) {
        int k = 1 << 30;
        for (int jj = j - wy; jj <= j + wy; jj++)
            for (int ii = i - 2; ii <= i + 2; ii++) k = std::min(k, kbeg_a[(size_t)(ii - b.ib) + (size_t)b.ni() * (size_t)(jj - b.jb)]);
        return k;
    };
    for (int j = jbeg - ey; j <= jend + ey; j++)
        for (int i = ibeg - 1; i <= iend + 1; i++) {
            for (int k = top_of(i, j); k <= kend; k++) {
                const size_t n = b.at(k, i, j), m = b.at(k - 1, i, j);
                if (!(b.lam[n] < b.lam[m] || b.mu[n] < b.mu[m])) continue;
                int k2 = k + 1;
                for (; k2 <= kend; k2++)
                    if (b.lam[b.at(k2, i, j)] > b.lam[b.at(k2 - 1, i, j)] || b.mu[b.at(k2, i, j)] > b.mu[b.at(k2 - 1, i, j)]) break;
                if (k2 - k <= LV_THICK) {
                    // the 3-D code replaces the top cell of the layer only (swpc_3d/m_medium.f90:299-303), the P-SV code the
                    // whole layer k .. k2-1 (swpc_psv/m_medium.f90:337-341)
                    for (int q = k; q <= (wy == 0 ? k2 - 1 : k); q++) {
                        const size_t nq = b.at(q, i, j);
                        b.rho[nq] = b.rho[m]; b.lam[nq] = b.lam[m]; b.mu[nq] = b.mu[m]; b.qp[nq] = b.qp[m]; b.qs[nq] = b.qs[m];
                    }
                    k = k2 - 1;
                }
            }
            for (int k = top_of(i, j); k <= kend; k++) {
                const size_t n = b.at(k, i, j);
                float vs = std::sqrt(b.mu[n] / b.rho[n]);
                if (vs < EPS_SP) continue;
                if (vs < vmin_pml) {
                    vs = vmin_pml;
                    const float vp = vs * std::sqrt(3.0f);
                    b.lam[n] = b.rho[n] * (vp * vp - 2 * (vs * vs));
                    b.mu[n] = b.rho[n] * (vs * vs);
                }
            }
        }
}

// ------------------------------------------------------------------------------------------------------------
// vmodel_grd, src/swpc_3d/m_vmodel_grd.f90:28-303: layers bounded by GMT grids in geographic coordinates, interpolated
// with the bicubic patch of src/shared/m_bicubic.f90.  Grids must be netCDF *classic* files (GMT's default netCDF-4
// container needs `gmt grdconvert in.grd out.grd=cf` or `nccopy -k classic` first: there is no HDF5 in this build).
class BicubicPatch {   // m_bicubic.f90: bicubic__init_d (:72-110), bicubic__coef (:231-288), bicubic__interp_d (:141-228)
  public:
    BicubicPatch(int nx, int ny, double x0, double y0, double dx, double dy, std::vector<double> f)
        : nx_(nx), ny_(ny), x0_(x0), y0_(y0), dx_(dx), dy_(dy), f_(std::move(f)), fx_(f_.size()), fy_(f_.size()), fxy_(f_.size()) {
        ddx(f_, fx_); ddy(f_, fy_); ddx(fy_, fxy_);
        for (size_t q = 0; q < f_.size(); q++) { fx_[q] = fx_[q] * dx_; fy_[q] = fy_[q] * dy_; fxy_[q] = fxy_[q] * dx_ * dy_; }
    }
    double operator()(double xi, double yi) {
        int ii = (int)std::floor((xi - x0_) / dx_) + 1, jj = (int)std::floor((yi - y0_) / dy_) + 1;   // 1-based pixel
        if (ii < 1 || ii > nx_ - 1 || jj <= 0 || jj > ny_ - 1) {   // outside: clamp to the edge pixel and its edge
            if (ii < 1) { ii = 1; xi = x0_; }
            if (ii > nx_ - 1) { ii = nx_ - 1; xi = x0_ + (nx_ - 1) * dx_; }
            if (jj < 1) { jj = 1; yi = y0_; }
            if (jj > ny_ - 1) { jj = ny_ - 1; yi = y0_ + (ny_ - 1) * dy_; }
        }
        if (ii != ii0_ || jj != jj0_) { coefficients(ii, jj); ii0_ = ii; jj0_ = jj; }
        const double xd = (xi - (x0_ + (ii - 1) * dx_)) / dx_, yd = (yi - (y0_ + (jj - 1) * dy_)) / dy_;
        const double px[4] = {1.0, xd, xd * xd, xd * xd * xd}, py[4] = {1.0, yd, yd * yd, yd * yd * yd};
        double v = 0.0;
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < 4; i++) v = v + a_[4 * j + i] * px[i] * py[j];
        return v;
    }

  private:
    int nx_, ny_, ii0_ = 0, jj0_ = 0;
    double x0_, y0_, dx_, dy_, a_[16];
    std::vector<double> f_, fx_, fy_, fxy_;
    size_t at(int i, int j) const { return (size_t)i + (size_t)nx_ * j; }
    void ddx(const std::vector<double> &g, std::vector<double> &o) const {   // diffx :291-313
        for (int j = 0; j < ny_; j++) {
            for (int i = 1; i < nx_ - 1; i++) o[at(i, j)] = (g[at(i + 1, j)] - g[at(i - 1, j)]) / (2 * dx_);
            o[at(0, j)] = (g[at(1, j)] - g[at(0, j)]) / dx_;
            o[at(nx_ - 1, j)] = (g[at(nx_ - 1, j)] - g[at(nx_ - 2, j)]) / dx_;
        }
    }
    void ddy(const std::vector<double> &g, std::vector<double> &o) const {   // diffy :316-338
        for (int i = 0; i < nx_; i++) {
            for (int j = 1; j < ny_ - 1; j++) o[at(i, j)] = (g[at(i, j + 1)] - g[at(i, j - 1)]) / (2 * dy_);
            o[at(i, 0)] = (g[at(i, 1)] - g[at(i, 0)]) / dy_;
            o[at(i, ny_ - 1)] = (g[at(i, ny_ - 1)] - g[at(i, ny_ - 2)]) / dy_;
        }
    }
    void coefficients(int ii, int jj) {   // a(i, j) = row 4j + i of M x, x = (f, fx, fy, fxy) at the four pixel corners
        static const signed char M[16][16] = {
            {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {-3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0},
            {0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0},
            {-3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0}, {0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0},
            {9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1}, {-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1},
            {2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0},
            {-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1}, {4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1}};
        const std::vector<double> *src[4] = {&f_, &fx_, &fy_, &fxy_};
        double x[16];
        for (int q = 0; q < 4; q++) {
            const std::vector<double> &g = *src[q];
            x[4 * q] = g[at(ii - 1, jj - 1)]; x[4 * q + 1] = g[at(ii, jj - 1)]; x[4 * q + 2] = g[at(ii - 1, jj)]; x[4 * q + 3] = g[at(ii, jj)];
        }
        for (int r = 0; r < 16; r++) {
            double acc = 0.0;
            for (int l = 0; l < 16; l++) acc = acc + (double)M[r][l] * x[l];
            a_[r] = acc;
        }
    }
};

struct GrdGeometry {   // what vmodel_grd needs beyond ModelEnv / MediumBox
    int nx, ny, na;
    float xbeg, ybeg, zbeg, clon, clat, phi;
    const float *xc, *yc;       // xc(ib:ie), yc(jb:je)
    float *bddep;               // bd(ib:ie, jb:je, 0:NBD)
    int nz = 0;
};

// with_rmed: vmodel_grd_rmed (m_vmodel_grd_rmed.f90:28-388) -- each layer additionally carries a random-media volume that is
// sampled at the depth below a reference interface `reflyr` (0: the top of the volume), wrapped cyclically
inline int vmodel_grd(const ModelEnv &e, const MediumBox &b, const GrdGeometry &gg, bool with_rmed = false) {
    const IniFile &ini = *e.ini;
    const std::string fn_lst = ini.get(with_rmed ? "fn_grdlst_rmed" : "fn_grdlst", "."), dir_grd = ini.get("dir_grd", ".");
    const bool topo_flatten = ini.get_l("topo_flatten", false);
    const bool is_ocean = ini.get_l("is_ocean", true) || topo_flatten;
    const int nk = b.nk(), ni = b.ni(), nj = b.je - b.jb + 1;
    const size_t n2 = (size_t)ni * nj;
    const float fdx = (float)e.dx, fdy = (float)e.dy, fdz = (float)e.dz;
    std::vector<float> cv((size_t)nk, 1.0f);
    if (e.flatten) for (int q = 0; q < nk; q++) cv[(size_t)q] = (float)std::exp((double)b.zc[q] / R_EARTH);
    // air, then ocean (:88-133): one plan per plane, filled column by column
    std::vector<PlanePlan> pl((size_t)nk);
    for (int q = 0; q < nk; q++) {
        PlanePlan &p = pl[(size_t)q];
        const float zc = b.zc[q];
        if (is_ocean && !(zc < 0)) { const float vp = cv[(size_t)q] * seawater_vel(zc, e.munk); p.rho = 1.0f; p.lam = 1.0f * (vp * vp - 2 * 0.0f * 0.0f); p.mu = 0.0f; p.qp = p.qs = 1000000.0f; }
        else { p.rho = 0.001f; p.lam = 0.0f; p.mu = 0.0f; p.qp = p.qs = 1.0f; }
    }
    fill_columns(b, pl, [](const PlanePlan &, size_t) {});
    // geographic location of every column, clamped to the inner edge of the absorber (:137-151)
    const float x_ab = i2x(gg.na + 1, gg.xbeg, fdx), x_ae = i2x(gg.nx - gg.na, gg.xbeg, fdx);
    const float y_ab = i2x(gg.na + 1, gg.ybeg, fdy), y_ae = i2x(gg.ny - gg.na, gg.ybeg, fdy);
    std::vector<float> glon(n2), glat(n2);
    if (e.psv) {   // swpc_psv/m_vmodel_grd.f90:127-129: the section runs along y = 0 and is not clamped to the absorber edge
        for (int i = 0; i < ni; i++) geomap_c2g(gg.xc[i], 0.0f, gg.clon, gg.clat, gg.phi, glon[(size_t)i], glat[(size_t)i]);
    } else {
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++)
                geomap_c2g(std::min(std::max(gg.xc[i], x_ab), x_ae), std::min(std::max(gg.yc[j], y_ab), y_ae), gg.clon, gg.clat, gg.phi,
                           glon[(size_t)i + (size_t)ni * j], glat[(size_t)i + (size_t)ni * j]);
    }
    // layer list (:154-173): 'file' rho vp vs qp qs pid
    struct Layer { std::string fn; float rho, vp, vs, qp, qs; int pid; int reflyr = 0; };
    std::vector<Layer> L;
    LayerTable names;   // random-media file per layer, for read_rmed_set
    {
        std::ifstream is(join_path(e.base, fn_lst));
        if (!is) return hfail("vmodel_grd: cannot open the layer list " + fn_lst);
        std::string line;
        while (std::getline(is, line)) {
            if (blank_or_comment(line)) continue;
            for (auto &ch : line) if (ch == ',') ch = ' ';
            std::istringstream ls(line);
            Layer y;
            if (!(ls >> y.fn >> y.rho >> y.vp >> y.vs >> y.qp >> y.qs >> y.pid)) continue;
            const auto unquote = [](std::string &t) { if (t.size() >= 2 && (t[0] == '\'' || t[0] == '"') && t.back() == t[0]) t = t.substr(1, t.size() - 2); };
            unquote(y.fn);
            y.fn = dir_grd + "/" + y.fn;
            if (with_rmed) {
                std::string rn;
                if (!(ls >> rn >> y.reflyr)) continue;
                unquote(rn);
                names.rmed.push_back(rn);
                names.depth.push_back(0.0f);
            }
            L.push_back(y);
        }
    }
    const int ngrd = (int)L.size();
    if (ngrd == 0) return hfail("vmodel_grd: no layer in the list " + fn_lst);
    for (int n = ngrd - 2; n >= 0; n--)
        // (swpc_psv/m_vmodel_grd_rmed.f90:185 drops the positivity clause the other three variants have)
        if ((L[n].vp < e.vcut || L[n].vs < e.vcut) && ((e.psv && with_rmed) || (L[n].vp > 0 && L[n].vs > 0))) { L[n].vp = L[n + 1].vp; L[n].vs = L[n + 1].vs; L[n].rho = L[n + 1].rho; L[n].qp = L[n + 1].qp; L[n].qs = L[n + 1].qs; }
    std::vector<int> tbl;
    std::vector<std::vector<float>> xi;
    const float rhomin = ini.get_s("rhomin", 1.0f), vmin = e.vcut, vmax = e.rmed_vmax();
    if (with_rmed) {
        for (const Layer &y : L) {
            if (!(0 <= y.reflyr && y.reflyr <= ngrd)) return hfail("assert: 0 <= reflyr <= ngrd (m_vmodel_grd_rmed.f90:208)");
            if (!e.psv && !(y.vp < vmax && y.vs < vmax)) return hfail("assert: background velocity exceeds the stability limit (m_vmodel_grd_rmed.f90:342-343)");
        }
        if (read_rmed_set(e, b, names, tbl, xi)) return 1;
    }
    std::fill(gg.bddep, gg.bddep + n2 * (NBD + 1), 0.0f);
    // interface index of every layer in every column (:180-268)
    std::vector<int> kgrd(n2 * (size_t)(ngrd + 1), b.kb - 1);
    const int ktopo = x2i(0.0f - fdz / 2, gg.zbeg, fdz);
    for (int n = 1; n <= ngrd; n++) {
        ClassicNc nc;
        std::string err;
        if (!nc.open(join_path(e.base, L[n - 1].fn), err)) return hfail("vmodel_grd: " + err + " (grids must be netCDF classic files)");
        if (nc.dim.size() != 2 || nc.var.size() < 3) return hfail("vmodel_grd: " + L[n - 1].fn + " is not a 2-D grid with variables x, y, z");
        const int nlon = (int)nc.dim[0], nlat = (int)nc.dim[1];
        std::vector<double> dep((size_t)nlon * nlat);
        for (size_t q = 0; q < dep.size(); q++) dep[q] = nc.value(2, (long long)q) / 1000;   // m -> km
        const double lon0 = nc.value(0, 0), lat0 = nc.value(1, 0);
        const double dlon = (nc.value(0, nlon - 1) - lon0) / (nlon - 1), dlat = (nc.value(1, nlat - 1) - lat0) / (nlat - 1);
        BicubicPatch patch(nlon, nlat, lon0, lat0, dlon, dlat, std::move(dep));
        const int *below = kgrd.data() + n2 * (size_t)(n - 1);
        int *mine = kgrd.data() + n2 * (size_t)n;
        for (size_t q = 0; q < n2; q++) {
            float z = (float)patch((double)glon[q], (double)glat[q]);
            if (e.flatten) z = (float)(-R_EARTH * std::log((R_EARTH - (double)z) / R_EARTH));
            if (n == 1) gg.bddep[q] = z;
            if (topo_flatten) z = z - gg.bddep[q];
            int k = std::max(x2i(z - fdz / 2, gg.zbeg, fdz), below[q]);
            if (n == 1 && z > 0) k = std::max(ktopo + 2, k);   // sea column at least two cells thick
            mine[q] = k;
            if (L[n - 1].pid > 0 && L[n - 1].pid <= NBD) gg.bddep[n2 * (size_t)L[n - 1].pid + q] = z;
        }
    }
    // every layer fills everything below its interface; deeper layers overwrite (:270-288)
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long q = 0; q < (long long)n2; q++) {
        const size_t n0 = (size_t)q * (size_t)nk;
        for (int n = 1; n <= ngrd; n++) {
            const Layer &y = L[(size_t)n - 1];
            const float *x3 = with_rmed ? xi[(size_t)tbl[(size_t)n - 1]].data() + n0 : nullptr;
            const int kref = kgrd[n2 * (size_t)y.reflyr + (size_t)q];
            for (int k = std::max(kgrd[n2 * (size_t)n + (size_t)q] + 1, b.kb); k <= b.ke; k++) {
                const size_t m = n0 + (size_t)(k - b.kb);
                const float c = cv[(size_t)(k - b.kb)];
                if (with_rmed) {   // m_vmodel_grd_rmed.f90:333-360
                    int kk = k - kref + 1;   // depth index below the reference interface, cyclic
                    if (kk < b.kb) kk = kk + gg.nz;
                    if (kk > b.ke) kk = kk - b.ke;
                    if (kk < b.kb || kk > b.ke) { bad++; continue; }
                    const float x = x3[kk - b.kb];
                    float vp2 = c * y.vp * (1.0f + x), vs2 = c * y.vs * (1.0f + x), rho2 = y.rho * (1.0f + 0.8f * x);
                    if (y.vp > 0 && y.vs > 0) vcheck(vp2, vs2, rho2, x, vmin, vmax, rhomin);
                    b.rho[m] = rho2;
                    b.lam[m] = rho2 * (vp2 * vp2 - 2 * vs2 * vs2);
                    b.mu[m] = rho2 * vs2 * vs2;
                } else {
                    b.rho[m] = y.rho;
                    b.lam[m] = y.rho * (c * c) * (y.vp * y.vp - 2 * y.vs * y.vs);
                    b.mu[m] = y.rho * (c * c) * y.vs * y.vs;
                }
                b.qp[m] = y.qp;
                b.qs[m] = y.qs;
            }
        }
    }
    if (bad) return hfail("vmodel_grd_rmed: relative depth index out of the volume");
    return 0;
}

}   // namespace
