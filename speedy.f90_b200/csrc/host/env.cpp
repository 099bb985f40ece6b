// Product-side host environment: boundary data, land/sea slab constants and the daily
// solar tables.  This is start-up work that runs once on the host; the per-step path is
// entirely on the device.  Follows input_output.f90:23-92 (read, N->S flip, -999 rule),
// boundaries.f90:47-72,98-142 (forchk, fillsf), land_model.f90:50-181, sea_model.f90:80-250
// and shortwave_radiation.f90:238-329.  Input: the packed boundary file written by
// tools/pack_boundary.py (the reference's NetCDF variables, unmodified).
#include "../model.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <sys/stat.h>

namespace spd {

namespace {
// ---- the reference's own boundary files (data/bc/t30/{clim,anom}/*.nc: NetCDF-4 = HDF5) read without an HDF5 library -------------
// What load_boundary_file (input_output.f90:23-92) asks of NetCDF — a named float32 variable of a file — is, in these files, one
// CONTIGUOUS little-endian dataset.  The reader finds the variable's hard link (1-byte name length, name, 8-byte object-header address:
// the encoding of a link message whether it sits in the group's header or in its fractal heap), checks that the address holds a
// version-2 object header ('OHDR'), and reads its dataspace (dimensions), datatype (4-byte IEEE little-endian float) and data-layout
// (version 3, contiguous: address + size) messages, following continuation blocks ('OCHK').  Anything else — chunked or compressed
// storage, another type — is refused with a pointer to tools/pack_boundary.py.
struct Hdf5Var { std::vector<unsigned long long> dims; unsigned long long addr = 0, size = 0; };

unsigned long long le(const std::vector<unsigned char>& b, size_t o, int n) {
    if (o + n > b.size()) throw std::runtime_error("boundary file: truncated HDF5 structure");
    unsigned long long v = 0;
    for (int k = n - 1; k >= 0; k--) v = (v << 8) | b[o + k];
    return v;
}

bool parse_dataset_header(const std::vector<unsigned char>& b, size_t a, Hdf5Var& out, std::string& why) {
    if (a + 8 > b.size() || memcmp(&b[a], "OHDR", 4) != 0 || b[a + 4] != 2) return false;
    const unsigned flags = b[a + 5];
    size_t p = a + 6;
    if (flags & 0x20) p += 16;                 // access / modification / change / birth times
    if (flags & 0x10) p += 4;                  // attribute storage phase-change values
    const int szn = 1 << (flags & 3);
    const unsigned long long c0 = le(b, p, szn);
    p += szn;
    const int co = (flags & 0x04) ? 2 : 0;     // creation-order field of every message
    std::vector<std::pair<size_t, size_t>> chunks{{p, p + (size_t)c0}};
    bool have_space = false, have_type = false, have_layout = false;
    for (size_t i = 0; i < chunks.size() && i < 64; i++) {
        size_t q = chunks[i].first;
        const size_t end = std::min(chunks[i].second, b.size());
        while (q + 4 + co <= end) {
            const unsigned type = b[q];
            const size_t sz = (size_t)le(b, q + 1, 2);
            q += 4 + co;
            const size_t d = q;
            q += sz;
            if (q > end) break;
            if (sz < 2 && type != 0) continue;        // too short to be one of the messages read below
            if (type == 0x10) {                // continuation: offset, length of an 'OCHK' block (signature first, checksum last)
                const size_t off = (size_t)le(b, d, 8), len = (size_t)le(b, d + 8, 8);
                if (off + len <= b.size() && len >= 8 && memcmp(&b[off], "OCHK", 4) == 0) chunks.push_back({off + 4, off + len - 4});
            } else if (type == 0x01) {         // dataspace
                const unsigned ver = b[d], rank = b[d + 1];
                const size_t dp = d + (ver == 2 ? 4 : 8);
                out.dims.clear();
                for (unsigned k = 0; k < rank; k++) out.dims.push_back(le(b, dp + 8 * k, 8));
                have_space = true;
            } else if (type == 0x03) {         // datatype: class in the low nibble, bit 0 of the first flag byte = byte order
                const unsigned cls = b[d] & 0x0f, big_endian = b[d + 1] & 1;
                const unsigned long long size = le(b, d + 4, 4);
                if (cls != 1 || size != 4 || big_endian) { why = "the variable is not a 4-byte little-endian float"; return false; }
                have_type = true;
            } else if (type == 0x08) {         // data layout
                if (b[d] != 3 || b[d + 1] != 1) { why = "the variable is not stored contiguously (chunked / compressed)"; return false; }
                out.addr = le(b, d + 2, 8);
                out.size = le(b, d + 10, 8);
                have_layout = true;
            }
        }
    }
    if (!(have_space && have_type && have_layout)) { why = "dataspace / datatype / layout message missing"; return false; }
    return true;
}

Hdf5Var find_variable(const std::vector<unsigned char>& b, const std::string& file, const std::string& name) {
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (b.size() < 64 || memcmp(b.data(), sig, 8) != 0)
        throw std::runtime_error(file + ": not a NetCDF-4 / HDF5 file (the reference's data/bc/t30 files are)");
    std::string why = "no link to a dataset of that name";
    const size_t n = name.size();
    for (size_t k = 1; k + n + 8 <= b.size(); k++) {
        if (b[k - 1] != n || memcmp(&b[k], name.data(), n) != 0) continue;
        const size_t addr = (size_t)le(b, k + n, 8);
        Hdf5Var v;
        if (addr + 8 <= b.size() && parse_dataset_header(b, addr, v, why)) {
            unsigned long long cnt = 1;
            for (auto d : v.dims) cnt *= d;
            if (v.size != 4 * cnt || v.addr + v.size > b.size()) { why = "layout size does not match the dimensions"; continue; }
            return v;
        }
    }
    throw std::runtime_error(file + ": variable " + name + ": " + why + " — pack the files with tools/pack_boundary.py instead");
}

std::vector<unsigned char> read_whole(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("cannot open " + path);
    std::vector<unsigned char> b;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) b.insert(b.end(), buf, buf + n);
    fclose(fp);
    return b;
}

bool is_directory(const char* path) {
    struct stat st;
    return stat(path, &st) == 0 && S_ISDIR(st.st_mode);
}
bool file_exists(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

struct Packed {
    int ix = 0, il = 0;
    std::map<std::string, std::vector<float>> f;
    // `dir` holds the reference's files, either side by side (the run directory run.sh links them into) or as the data/bc/t30 tree
    void load_reference_tree(const std::string& dir) {
        struct Src { const char *file, *sub; std::vector<const char*> vars; };
        const Src srcs[] = {{"surface.nc", "clim", {"orog", "lsm", "alb", "vegh", "vegl"}}, {"land.nc", "clim", {"stl"}},
                            {"sea_surface_temperature.nc", "clim", {"sst"}}, {"sea_ice.nc", "clim", {"icec"}}, {"snow.nc", "clim", {"snowd"}},
                            {"soil.nc", "clim", {"swl1", "swl2"}}, {"sea_surface_temperature_anomaly.nc", "anom", {"ssta"}}};
        for (const Src& s : srcs) {
            std::string path = dir + "/" + s.file;
            if (!file_exists(path)) path = dir + "/" + s.sub + "/" + s.file;
            if (!file_exists(path)) throw std::runtime_error("boundary directory " + dir + ": " + s.file + " not found (nor under " + s.sub + "/)");
            const std::vector<unsigned char> b = read_whole(path);
            for (const char* name : s.vars) {
                const Hdf5Var v = find_variable(b, path, name);
                if (v.dims.size() < 2 || v.dims.size() > 3) throw std::runtime_error(path + ": variable " + name + " is not (lat, lon) or (time, lat, lon)");
                const int vil = (int)v.dims[v.dims.size() - 2], vix = (int)v.dims[v.dims.size() - 1];
                if (ix == 0) { ix = vix; il = vil; }
                if (vix != ix || vil != il) throw std::runtime_error(path + ": variable " + name + " has another grid than the other boundary fields");
                std::vector<float> data((size_t)(v.size / 4));
                memcpy(data.data(), b.data() + v.addr, (size_t)v.size);      // little-endian float32, as the host
                f[name] = std::move(data);
            }
        }
    }
    void load(const char* path) {
        if (is_directory(path)) { load_reference_tree(path); return; }
        FILE* fp = fopen(path, "rb");
        if (!fp) throw std::runtime_error(std::string("cannot open boundary file ") + path);
        char magic[8];
        int hdr[3];
        bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "SPDYBC01", 8) == 0 && fread(hdr, 4, 3, fp) == 3;
        if (!ok) { fclose(fp); throw std::runtime_error("bad boundary file header (expected SPDYBC01, see tools/pack_boundary.py)"); }
        ix = hdr[0]; il = hdr[1];
        for (int q = 0; q < hdr[2]; q++) {
            char name[17] = {0};
            int nrec = 0;
            if (fread(name, 1, 16, fp) != 16 || fread(&nrec, 4, 1, fp) != 1 || nrec < 1) { fclose(fp); throw std::runtime_error("truncated boundary file"); }
            std::vector<float> v((size_t)nrec * ix * il);
            if (fread(v.data(), 4, v.size(), fp) != v.size()) { fclose(fp); throw std::runtime_error("truncated boundary file"); }
            f[name] = std::move(v);
        }
        fclose(fp);
    }
    int nrec(const char* name) const {
        auto it = f.find(name);
        if (it == f.end()) throw std::runtime_error(std::string("boundary field missing: ") + name);
        return (int)(it->second.size() / ((size_t)ix * il));
    }
    // one record as the reference's load_boundary_file returns it: S->N, values <= -999 zeroed
    void get(const char* name, int rec, double* out) const {
        auto it = f.find(name);
        if (it == f.end()) throw std::runtime_error(std::string("boundary field missing: ") + name);
        const float* raw = it->second.data() + (size_t)rec * ix * il;
        for (int j = 0; j < il; j++)
            for (int i = 0; i < ix; i++) {
                double v = (double)raw[i + (size_t)ix * (il - 1 - j)];
                out[i + (size_t)ix * j] = (v <= -999) ? 0.0 : v;
            }
    }
};

// boundaries.f90:98-142
void fill_missing(double* sf, int ix, int il, double fmis) {
    std::vector<double> row(ix + 2);
    double fmean = 0.0;
    auto do_row = [&](int j) {
        int nmis = 0;
        for (int i = 0; i < ix; i++) {
            row[i + 1] = sf[i + (size_t)ix * j];
            if (sf[i + (size_t)ix * j] < fmis) { nmis++; row[i + 1] = 0.0; }
        }
        if (nmis < ix) {
            double s = 0.0;
            for (int i = 1; i <= ix; i++) s += row[i];
            fmean = s / (double)(float)(ix - nmis);
        }
        for (int i = 0; i < ix; i++)
            if (sf[i + (size_t)ix * j] < fmis) row[i + 1] = fmean;
        row[0] = row[ix];
        row[ix + 1] = row[1];
        for (int i = 0; i < ix; i++)
            if (sf[i + (size_t)ix * j] < fmis) sf[i + (size_t)ix * j] = 0.5 * (row[i] + row[i + 2]);
    };
    for (int j = il / 2 - 1; j >= 0; j--) do_row(j);     // hemisphere 1: j = il/2 .. 1
    // hemisphere 2 starts at j1+1 where j1 is still il/2 (boundaries.f90:113), i.e. 1-based il/2+1
    for (int j = il / 2; j < il; j++) do_row(j);
}

// boundaries.f90:47-72: points outside the mask are set to fset
void mask_fill(const std::vector<double>& mask, int n2d, int nf, double fset, double* field) {
    for (int jf = 0; jf < nf; jf++)
        for (int q = 0; q < n2d; q++)
            if (!(mask[q] > 0.0)) field[(size_t)jf * n2d + q] = fset;
}

// shortwave_radiation.f90:287-329
void solar_top(double tyear, double csol, const Tables& t, std::vector<double>& topsr) {
    const int il = t.d.il;
    const double pigr = (double)(2.0f * asinf(1.0f));
    const double alpha = 2.0 * pigr * tyear;
    const double ca1 = cos(alpha), sa1 = sin(alpha);
    const double ca2 = ca1 * ca1 - sa1 * sa1, sa2 = 2. * sa1 * ca1;
    const double ca3 = ca1 * ca2 - sa1 * sa2, sa3 = sa1 * ca2 + sa2 * ca1;
    const double decl = (double)0.006918f - (double)0.399912f * ca1 + (double)0.070257f * sa1 - (double)0.006758f * ca2 +
                        (double)0.000907f * sa2 - (double)0.002697f * ca3 + (double)0.001480f * sa3;
    const double fdis = (double)1.000110f + (double)0.034221f * ca1 + (double)0.001280f * sa1 + (double)0.000719f * ca2 + (double)0.000077f * sa2;
    const double cdecl_ = cos(decl), sdecl = sin(decl), tdecl = sdecl / cdecl_;
    const double csolp = csol / pigr;
    topsr.resize(il);
    for (int j = 0; j < il; j++) {
        double ch0 = std::min(1.0, std::max(-1.0, -tdecl * t.sia[j] / t.coa[j]));
        double h0 = acos(ch0), sh0 = sin(h0);
        topsr[j] = csolp * fdis * (h0 * t.sia[j] * sdecl + sh0 * t.coa[j] * cdecl_);
    }
}
}  // namespace

// The zonally symmetric daily radiation inputs take only 365 distinct values of tyear
// (date.f90:151), so they are tabulated once: solar[doy][f][j], f = fsol, ozone, ozupp, zenit, stratz
static void build_solar_tables(const Tables& t, std::vector<double>& solar) {
    const int il = t.d.il;
    solar.assign((size_t)365 * 5 * il, 0.0);
    const double solc = 342.0, epssw = (double)0.020f;
    std::vector<double> topsr;
    for (int doy = 0; doy < 365; doy++) {
        const double tyear = (double)(((float)(doy + 1) - 0.5f) / 365.0f);
        const double alpha = (double)(4.0f * asinf(1.0f)) * (tyear + (double)(10.0f / 365.0f));   // shortwave_radiation.f90:248
        const double coz1 = 1.0 * std::max(0.0, cos(alpha - 0.0));
        const double coz2 = (double)1.8f, azen = 1.0, fs0 = 6.0;
        const double rzen = -cos(alpha) * (double)23.45f * (double)asinf(1.0f) / 90.0;            // :257
        solar_top(tyear, 4.0 * solc, t, topsr);
        double* S = solar.data() + (size_t)doy * 5 * il;
        for (int j = 0; j < il; j++) {
            const double flat2 = 1.5 * (t.sia[j] * t.sia[j]) - 0.5;
            const double fsol = topsr[j];
            double ozupp = 0.5 * epssw;
            double ozone = (double)0.4f * epssw * (1.0 + coz1 * t.sia[j] + coz2 * flat2);
            const double z = 1.0 - (t.coa[j] * cos(rzen) + t.sia[j] * sin(rzen));
            const double zenit = 1.0 + azen * (z * z);
            ozupp = fsol * ozupp * zenit;
            ozone = fsol * ozone * zenit;
            S[0 * il + j] = fsol;
            S[1 * il + j] = ozone;
            S[2 * il + j] = ozupp;
            S[3 * il + j] = zenit;
            S[4 * il + j] = std::max(fs0 - fsol, 0.0);
        }
    }
}

void load_host_env(const char* bc_path, const Tables& tab, HostEnv& env) {
    Packed pk;
    pk.load(bc_path);
    const int ix = tab.d.ix, il = tab.d.il, N = ix * il;
    if (pk.ix != ix || pk.il != il) throw std::runtime_error("boundary file resolution does not match the context");
    env.ix = ix; env.il = il;
    const Consts& c = tab.c;
    // boundaries.f90:28-43
    env.phi0.resize(N); env.fmask.resize(N); env.alb0.resize(N);
    pk.get("orog", 0, env.phi0.data());
    for (auto& v : env.phi0) v = c.grav * v;
    pk.get("lsm", 0, env.fmask.data());
    pk.get("alb", 0, env.alb0.data());

    // ---- land_model.f90:50-181
    const double thrsh = (double)0.1f;
    env.fmask_l = env.fmask; env.bmask_l.assign(N, 0.0);
    for (int q = 0; q < N; q++) {
        if (env.fmask_l[q] >= thrsh) {
            env.bmask_l[q] = 1.0;
            if (env.fmask[q] > (1.0 - thrsh)) env.fmask_l[q] = 1.0;
        } else {
            env.bmask_l[q] = 0.0;
            env.fmask_l[q] = 0.0;
        }
    }
    env.stl12.resize((size_t)12 * N); env.snowd12.resize((size_t)12 * N); env.soilw12.resize((size_t)12 * N);
    for (int mth = 0; mth < 12; mth++) {
        pk.get("stl", mth, env.stl12.data() + (size_t)mth * N);
        fill_missing(env.stl12.data() + (size_t)mth * N, ix, il, 0.0);
        pk.get("snowd", mth, env.snowd12.data() + (size_t)mth * N);
    }
    mask_fill(env.bmask_l, N, 12, 273.0, env.stl12.data());
    mask_fill(env.bmask_l, N, 12, 0.0, env.snowd12.data());
    {
        std::vector<double> vegh(N), vegl(N), veg(N), swl1(N), swl2(N);
        pk.get("vegh", 0, vegh.data());
        pk.get("vegl", 0, vegl.data());
        for (int q = 0; q < N; q++) veg[q] = std::max(0.0, vegh[q] + (double)0.8f * vegl[q]);
        const double swcap = (double)0.30f, swwil = (double)0.17f;
        const int idep2 = 3;
        const double swwil2 = idep2 * swwil;
        const double rsw = 1.0 / (swcap + idep2 * (swcap - swwil));
        for (int mth = 0; mth < 12; mth++) {
            pk.get("swl1", mth, swl1.data());
            pk.get("swl2", mth, swl2.data());
            for (int q = 0; q < N; q++) {
                const double swroot = idep2 * swl2[q];
                env.soilw12[(size_t)mth * N + q] = std::min(1.0, rsw * (swl1[q] + veg[q] * std::max(0.0, swroot - swwil2)));
            }
        }
        mask_fill(env.bmask_l, N, 12, 0.0, env.soilw12.data());
    }
    {
        const double tdland = 40., flandmin = (double)(1.f / 3.f);
        const double hcapl = 1.0 * (double)2.50e+6f, hcapli = 5.0 * (double)1.93e+6f;
        env.rhcapl.resize(N); env.cdland.resize(N);
        for (int q = 0; q < N; q++) {
            const double dmask = (env.fmask_l[q] < flandmin) ? 0.0 : 1.0;
            env.rhcapl[q] = (env.alb0[q] < (double)0.4f) ? c.delt / hcapl : c.delt / hcapli;
            env.cdland[q] = dmask * tdland / (1. + dmask * tdland);
        }
    }

    // ---- sea_model.f90:80-250
    env.fmask_s.resize(N); env.bmask_s.resize(N);
    for (int q = 0; q < N; q++) {
        env.fmask_s[q] = 1.0 - env.fmask[q];
        if (env.fmask_s[q] >= thrsh) {
            env.bmask_s[q] = 1.0;
            if (env.fmask_s[q] > (1.0 - thrsh)) env.fmask_s[q] = 1.0;
        } else {
            env.bmask_s[q] = 0.0;
            env.fmask_s[q] = 0.0;
        }
    }
    env.deglat_s.resize(il);
    for (int j = 0; j < il; j++) env.deglat_s[j] = tab.radang[j] * 90.0 / (double)asinf(1.0f);
    env.sst12.resize((size_t)12 * N); env.sice12.resize((size_t)12 * N);
    for (int mth = 0; mth < 12; mth++) {
        pk.get("sst", mth, env.sst12.data() + (size_t)mth * N);
        fill_missing(env.sst12.data() + (size_t)mth * N, ix, il, 0.0);
        pk.get("icec", mth, env.sice12.data() + (size_t)mth * N);
        for (int q = 0; q < N; q++) env.sice12[(size_t)mth * N + q] = std::max(env.sice12[(size_t)mth * N + q], 0.0);
    }
    mask_fill(env.bmask_s, N, 12, 273.0, env.sst12.data());
    mask_fill(env.bmask_s, N, 12, 0.0, env.sice12.data());
    // SST anomaly window: every record resident, flipped and masked once (forchk fset = 0)
    env.nssta = pk.nrec("ssta");
    env.ssta.resize((size_t)env.nssta * N);
    {
        std::vector<double> tmp(N);
        for (int r = 0; r < env.nssta; r++) {
            pk.get("ssta", r, tmp.data());
            for (int q = 0; q < N; q++) env.ssta[(size_t)r * N + q] = (env.bmask_s[q] > 0.0) ? (float)tmp[q] : 0.0f;
        }
    }
    {
        const double depth_ml = 60., dept0_ml = 40., depth_ice = 2.5, dept0_ice = 1.5, tdsst = 90., tdice = 30.0;
        const double fseamin = (double)(1.f / 3.f);
        const double crad = (double)(asinf(1.f) / 90.f);
        env.rhcaps.resize(N); env.rhcapi.resize(N); env.cdsea.resize(N); env.cdice.resize(N);
        for (int j = 0; j < il; j++) {
            const double coslat = cos(crad * env.deglat_s[j]);
            const double hcaps = (double)4.18e+6f * (depth_ml + (dept0_ml - depth_ml) * (coslat * coslat * coslat));
            const double hcapi = (double)1.93e+6f * (depth_ice + (dept0_ice - depth_ice) * (coslat * coslat));
            for (int i = 0; i < ix; i++) {
                const int q = i + ix * j;
                // l_globe: the latitudinal smoothing of an all-ones mask leaves it at one (0.25*(1+2+1))
                const double dmask = (env.fmask_s[q] < fseamin) ? 0.0 : 1.0;
                env.rhcaps[q] = c.delt / hcaps;
                env.rhcapi[q] = c.delt / hcapi;
                env.cdsea[q] = dmask * tdsst / (1. + dmask * tdsst);
                env.cdice[q] = dmask * tdice / (1. + dmask * tdice);
            }
        }
    }
    build_solar_tables(tab, env.solar);
}

}  // namespace spd
