// Minimal writer for the NetCDF classic format (CDF-1: 'C','D','F',1; big-endian; 32-bit offsets) — what
// nf90_create(..., nf90_clobber, ...) produces in the reference's output routine (input_output.f90:133).  Only
// what that routine needs: fixed dimensions + one unlimited one, real4 variables, text attributes, one record.
// No NetCDF library is involved; files are readable by any NetCDF reader (the tests use scipy.io.netcdf_file).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace spd {

class NcClassicWriter {
public:
    // returns the dimension id; len == 0 declares the unlimited (record) dimension
    int def_dim(const std::string& name, uint32_t len) { dims_.push_back({name, len}); return (int)dims_.size() - 1; }
    // dimension ids slowest-varying first (the C / CDL order); a variable whose first dimension is the unlimited
    // one is a record variable
    int def_var(const std::string& name, std::vector<int> dimids) { vars_.push_back({name, std::move(dimids), {}, {}}); return (int)vars_.size() - 1; }
    void put_att(int var, const std::string& name, const std::string& text) { vars_[var].atts.push_back({name, text}); }
    void put_var(int var, const float* data, size_t n) {
        if (n != var_len(vars_[var])) throw std::runtime_error("netcdf writer: size mismatch for variable " + vars_[var].name);
        vars_[var].data.assign(data, data + n);
        vars_[var].ref = nullptr;
    }
    // the same without a copy: `data` must stay valid until write() has returned (the large fields of an output file)
    void put_var_ref(int var, const float* data, size_t n) {
        if (n != var_len(vars_[var])) throw std::runtime_error("netcdf writer: size mismatch for variable " + vars_[var].name);
        vars_[var].data.clear();
        vars_[var].ref = data;
    }
    void write(const std::string& path) const {
        // header size first (offsets of the data sections depend on it)
        std::vector<uint8_t> h;
        auto build = [&](uint32_t base) {
            h.clear();
            const uint8_t magic[4] = {'C', 'D', 'F', 1};
            h.insert(h.end(), magic, magic + 4);
            u32(h, 1);                                            // numrecs
            u32(h, 0x0A); u32(h, (uint32_t)dims_.size());         // NC_DIMENSION
            for (const Dim& d : dims_) { name(h, d.name); u32(h, d.len); }
            u32(h, 0); u32(h, 0);                                 // no global attributes
            u32(h, 0x0B); u32(h, (uint32_t)vars_.size());         // NC_VARIABLE
            uint32_t off = base;
            // fixed-size variables first in the file, in definition order, then the record
            std::vector<uint32_t> begin(vars_.size());
            for (int pass = 0; pass < 2; pass++)
                for (size_t i = 0; i < vars_.size(); i++)
                    if (is_record(vars_[i]) == (pass == 1)) { begin[i] = off; off += (uint32_t)(4 * var_len(vars_[i])); }
            for (size_t i = 0; i < vars_.size(); i++) {
                const Var& v = vars_[i];
                name(h, v.name);
                u32(h, (uint32_t)v.dimids.size());
                for (int d : v.dimids) u32(h, (uint32_t)d);
                if (v.atts.empty()) { u32(h, 0); u32(h, 0); }
                else {
                    u32(h, 0x0C); u32(h, (uint32_t)v.atts.size()); // NC_ATTRIBUTE
                    for (const auto& a : v.atts) {
                        name(h, a.first);
                        u32(h, 2);                                // NC_CHAR
                        name(h, a.second);                        // nelems + padded bytes, as for a name
                    }
                }
                u32(h, 5);                                        // NC_FLOAT
                u32(h, (uint32_t)(4 * var_len(v)));               // vsize (per record for record variables)
                u32(h, begin[i]);
            }
        };
        build(0);
        build((uint32_t)h.size());
        // one buffer, one write: header, then every variable byte-swapped straight into its place
        size_t total = h.size();
        for (const Var& v : vars_) total += 4 * var_len(v);
        std::vector<uint8_t> file(total);
        memcpy(file.data(), h.data(), h.size());
        size_t off = h.size();
        for (int pass = 0; pass < 2; pass++)
            for (const Var& v : vars_) {
                if (is_record(v) != (pass == 1)) continue;
                const size_t n = var_len(v);
                const float* src = v.ref ? v.ref : v.data.data();
                if (!v.ref && v.data.size() != n) throw std::runtime_error("netcdf writer: variable " + v.name + " was never written");
                uint8_t* dst = file.data() + off;
                for (size_t i = 0; i < n; i++) {
                    uint32_t w; memcpy(&w, &src[i], 4);
                    w = __builtin_bswap32(w);
                    memcpy(dst + 4 * i, &w, 4);
                }
                off += 4 * n;
            }
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("netcdf writer: cannot create " + path);
        bool ok = fwrite(file.data(), 1, file.size(), f) == file.size();
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("netcdf writer: short write to " + path);
    }

private:
    struct Dim { std::string name; uint32_t len; };
    struct Var { std::string name; std::vector<int> dimids; std::vector<std::pair<std::string, std::string>> atts; std::vector<float> data; const float* ref = nullptr; };
    std::vector<Dim> dims_;
    std::vector<Var> vars_;
    bool is_record(const Var& v) const { return !v.dimids.empty() && dims_[v.dimids[0]].len == 0; }
    size_t var_len(const Var& v) const {   // elements of the variable (of one record for record variables)
        size_t n = 1;
        for (int d : v.dimids) n *= dims_[d].len ? dims_[d].len : 1;
        return n;
    }
    static void u32(std::vector<uint8_t>& h, uint32_t v) { h.push_back((uint8_t)(v >> 24)); h.push_back((uint8_t)(v >> 16)); h.push_back((uint8_t)(v >> 8)); h.push_back((uint8_t)v); }
    static void name(std::vector<uint8_t>& h, const std::string& s) {
        u32(h, (uint32_t)s.size());
        h.insert(h.end(), s.begin(), s.end());
        while (h.size() % 4) h.push_back(0);
    }
};

}  // namespace spd
