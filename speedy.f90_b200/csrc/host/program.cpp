// The caller of the hot path: `program speedy` (speedy.f90:1-54) with the two namelist groups it reads
// (params.f90:46-70 `&params nsteps_out nstdia`, date.f90:54-71 `&date start_datetime end_datetime`), written on
// top of the library's own C ABI (no CUDA in this file): speedy_run_steps advances the device-resident state to the
// next event, speedy_check_diagnostics / speedy_write_output produce what the reference prints and writes there.
//
//   step 0      diagnostics of time level 1 (prognostics.f90:120) and output(0, ...) (prognostics.f90:123-126): first_step
//               leaves time level 1 and its geopotential untouched, so both come from the resident state
//   step k > 0  ' step =...' print when mod(k, nstdia) == 0 (diagnostics.f90:53-57), output when mod(k, nsteps_out) == 0
//               (speedy.f90:50), until the model date equals the end date (speedy.f90:27)
#include "../../../include/speedy_b200.h"
#include "../abi_util.h"
#include "../calendar.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cerrno>
#include <sys/stat.h>

namespace {

std::string lower(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}

// One namelist group as (name, values) pairs.  The subset of Fortran namelist input the reference's files use and the
// obvious variants: `&group ... /` (or `$group ... $end`), `!` comments, `name = v`, `derived%component = v`,
// `derived = v1, v2, ...` (components in declaration order), blanks or commas between values, any letter case.
struct Group {
    bool found = false;
    std::vector<std::pair<std::string, std::vector<long long>>> items;
};

Group parse_group(const std::string& text, const std::string& group) {
    Group g;
    // strip comments
    std::string t;
    bool in_comment = false;
    for (char c : text) {
        if (c == '!') in_comment = true;
        if (c == '\n') in_comment = false;
        t += in_comment ? ' ' : c;
    }
    const std::string low = lower(t);
    size_t pos = 0;
    while (pos < low.size()) {
        const size_t a = low.find_first_of("&$", pos);
        if (a == std::string::npos) break;
        size_t b = a + 1;
        while (b < low.size() && (isalnum((unsigned char)low[b]) || low[b] == '_')) b++;
        const std::string name = low.substr(a + 1, b - a - 1);
        // body up to the terminating '/' (or &end / $end)
        size_t e = b;
        while (e < low.size() && low[e] != '/' && !(low[e] == '&' || low[e] == '$')) e++;
        if (name == group) {
            g.found = true;
            const std::string body = low.substr(b, e - b);
            size_t i = 0;
            std::string cur;
            while (i < body.size()) {
                const unsigned char c = (unsigned char)body[i];
                if (isspace(c) || c == ',') { i++; continue; }
                if (isalpha(c)) {
                    size_t j = i;
                    while (j < body.size() && (isalnum((unsigned char)body[j]) || body[j] == '_' || body[j] == '%')) j++;
                    cur = body.substr(i, j - i);
                    while (j < body.size() && isspace((unsigned char)body[j])) j++;
                    if (j >= body.size() || body[j] != '=') throw std::runtime_error("namelist group " + group + ": expected '=' after " + cur);
                    g.items.push_back({cur, {}});
                    i = j + 1;
                } else if (isdigit(c) || c == '-' || c == '+') {
                    char* endp = nullptr;
                    const long long v = strtoll(body.c_str() + i, &endp, 10);
                    if (endp == body.c_str() + i) throw std::runtime_error("namelist group " + group + ": bad number");
                    if (g.items.empty()) throw std::runtime_error("namelist group " + group + ": value without a name");
                    g.items.back().second.push_back(v);
                    i = (size_t)(endp - body.c_str());
                } else {
                    throw std::runtime_error("namelist group " + group + ": unexpected character '" + std::string(1, (char)c) + "'");
                }
            }
        }
        pos = (e < low.size() && low[e] == '/') ? e + 1 : std::max(e, b + 1);
        if (e < low.size() && (low[e] == '&' || low[e] == '$')) {     // "&end" closes the group
            size_t k = e + 1;
            while (k < low.size() && isalpha((unsigned char)low[k])) k++;
            if (low.substr(e + 1, k - e - 1) == "end") pos = k;
        }
    }
    return g;
}

void assign_datetime(int* dst, const std::string& name, const std::string& var, const std::vector<long long>& v) {
    static const char* comp[5] = {"year", "month", "day", "hour", "minute"};      // date.f90:14-20
    if (name == var) {
        if (v.size() > 5) throw std::runtime_error("namelist: too many values for " + var);
        for (size_t k = 0; k < v.size(); k++) dst[k] = (int)v[k];
        return;
    }
    for (int k = 0; k < 5; k++)
        if (name == var + "%" + comp[k]) {
            if (v.size() != 1) throw std::runtime_error("namelist: " + name + " takes one value");
            dst[k] = (int)v[0];
            return;
        }
    throw std::runtime_error("namelist: unknown variable " + name + " in group date");
}

bool same_date(const int* a, const int* b) {        // date.f90:23-36 datetime_equal
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3] && a[4] == b[4];
}

// formats 2001-2003 of diagnostics.f90:71-73; diag = diag(kx,3) in Fortran order
void print_diagnostics(long long istep, const double* diag, int kx) {
    printf(" step =%6lld reke =", istep);
    for (int k = 0; k < kx; k++) printf("%8.2f", diag[k]);
    printf("\n%13s deke =", "");
    for (int k = 0; k < kx; k++) printf("%8.2f", diag[kx + k]);
    printf("\n%13s temp =", "");
    for (int k = 0; k < kx; k++) printf("%8.2f", diag[2 * kx + k]);
    printf("\n");
    fflush(stdout);
}

}  // namespace

extern "C" {

int speedy_namelist_defaults(speedy_namelist* nml) {
    API_BEGIN
    if (!nml) throw std::runtime_error("null namelist");
    nml->nsteps_out = 1;                 // params.f90:58
    nml->nstdia = 36 * 5;                // params.f90:59
    const int s[5] = {1982, 1, 1, 0, 0}, e[5] = {1982, 2, 1, 0, 0};     // date.f90:62-63
    memcpy(nml->start_datetime, s, sizeof s);
    memcpy(nml->end_datetime, e, sizeof e);
    API_END
}

int speedy_read_namelist(const char* path, speedy_namelist* nml) {
    API_BEGIN
    if (speedy_namelist_defaults(nml)) throw std::runtime_error(speedy_last_error());
    if (!path || !*path) return 0;
    FILE* fp = fopen(path, "rb");
    if (!fp) return 0;                   // `inquire(file=..., exist=...)`: a missing file keeps the defaults (params.f90:62-67)
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
    fclose(fp);
    // a file that exists must hold both groups: `read(10, nml=...)` fails at end of file otherwise
    const Group p = parse_group(text, "params"), d = parse_group(text, "date");
    if (!p.found) throw std::runtime_error(std::string(path) + ": namelist group &params not found");
    if (!d.found) throw std::runtime_error(std::string(path) + ": namelist group &date not found");
    for (const auto& it : p.items) {
        if (it.second.size() != 1) throw std::runtime_error("namelist: " + it.first + " takes one value");
        if (it.first == "nsteps_out") nml->nsteps_out = (int)it.second[0];
        else if (it.first == "nstdia") nml->nstdia = (int)it.second[0];
        else throw std::runtime_error("namelist: unknown variable " + it.first + " in group params");
    }
    for (const auto& it : d.items) {
        if (it.first.rfind("start_datetime", 0) == 0) assign_datetime(nml->start_datetime, it.first, "start_datetime", it.second);
        else if (it.first.rfind("end_datetime", 0) == 0) assign_datetime(nml->end_datetime, it.first, "end_datetime", it.second);
        else throw std::runtime_error("namelist: unknown variable " + it.first + " in group date");
    }
    API_END
}

long long speedy_steps_between(const int* start_ymdhm, const int* end_ymdhm, int nsteps) {
    // the main loop runs `do while (.not. datetime_equal(model_datetime, end_datetime))` with newdate's calendar: count its trips
    try {
        if (!start_ymdhm || !end_ymdhm) throw std::runtime_error("null date");
        if (nsteps <= 0) nsteps = 36;
        spd::DevClock c;
        spd::calendar_init(c, start_ymdhm[0], start_ymdhm[1], start_ymdhm[2], start_ymdhm[3], start_ymdhm[4], 1 << 30, nsteps);
        const long long cap = 400LL * 366 * nsteps;
        for (long long k = 0; k <= cap; k++) {
            const int now[5] = {c.year, c.month, c.day, c.hour, c.minute};
            if (same_date(now, end_ymdhm)) return k;
            spd::cal_advance(c);
        }
        throw std::runtime_error("the end date is never reached from the start date in steps of 1/nsteps day (the reference would not terminate)");
    } catch (const std::exception& e_) { spd::last_error() = e_.what(); return -1; }
}

int speedy_main_loop(speedy_ctx* ctx, const speedy_namelist* nml, const char* out_dir, int member, int verbose, long long* steps_done) {
    long long done = 0;
    auto finish = [&](int rc) { if (steps_done) *steps_done = done; return rc; };
    try {
        if (!ctx || !nml) throw std::runtime_error("null argument");
        if (nml->nsteps_out < 1 || nml->nstdia < 1) throw std::runtime_error("nsteps_out and nstdia must be positive");
        int dims[8];
        if (speedy_dims(ctx, dims)) throw std::runtime_error(speedy_last_error());
        const int kx = dims[4];
        int now[5];
        long long model_step = 0;
        if (speedy_model_date(ctx, now, &model_step)) throw std::runtime_error(speedy_last_error());
        if (model_step != 1 || !same_date(now, nml->start_datetime))
            throw std::runtime_error("speedy_main_loop: call speedy_model_init with the namelist's start date first (model_step must be 1)");
        int info[4];
        if (speedy_run_info(ctx, info)) throw std::runtime_error(speedy_last_error());
        const int nmem = info[0], nsteps_day = info[1];
        const long long total = speedy_steps_between(nml->start_datetime, nml->end_datetime, nsteps_day);
        if (total < 0) throw std::runtime_error(speedy_last_error());
        std::vector<double> diag(3 * (size_t)kx);
        const bool write = out_dir != nullptr;
        // the host's own copy of the calendar (date.f90:109-157): names the files without asking the device
        spd::DevClock hc;
        spd::calendar_init(hc, nml->start_datetime[0], nml->start_datetime[1], nml->start_datetime[2], nml->start_datetime[3], nml->start_datetime[4], 1 << 30, nsteps_day);
        std::vector<std::pair<long long, std::string>> files;          // (timestep, path) of every file of this run
        const int lo = member < 0 ? 0 : member, hi = member < 0 ? nmem : member + 1;
        auto member_dir = [&](int e) { return member < 0 ? std::string(out_dir) + "/member" + std::to_string(e) : std::string(out_dir); };
        // output(timestep, ...) of the state now at the end of the stream: conversions enqueued, files written by the library's host threads
        auto emit = [&](long long timestep) {
            const int ymdhm[5] = {hc.year, hc.month, hc.day, hc.hour, hc.minute};
            char name[40];
            snprintf(name, sizeof name, "%04d%02d%02d%02d%02d.nc", ymdhm[0], ymdhm[1], ymdhm[2], ymdhm[3], ymdhm[4]);
            for (int e = lo; e < hi; e++) {
                const std::string dir = member_dir(e);
                if (speedy_write_output_async(ctx, e, dir.c_str(), ymdhm, timestep)) throw std::runtime_error(speedy_last_error());
                files.push_back({timestep, dir.empty() ? std::string(name) : dir + "/" + name});
            }
        };
        // 'Model variables out of accepted range' (diagnostics.f90:60-69): the failing step's numbers, then stop.  The reference stops
        // inside check_diagnostics of step `bad`, before that step's output: files from there on (enqueued before the guard was polled) go
        auto range_failure = [&]() {
            long long bad = 0;
            if (speedy_range_failure(ctx, &bad, diag.data()) < 0) throw std::runtime_error(speedy_last_error());
            speedy_output_drain(ctx);
            for (const auto& f : files)
                if (f.first >= bad) remove(f.second.c_str());
            done = bad;
            print_diagnostics(done, diag.data(), kx);
            return finish(1);
        };
        // ---- step 0 (prognostics.f90:120-126): level-1 diagnostics and output(0, ...).  first_step leaves time level 1 and its
        // geopotential as they were (eps = 0, time_stepping.f90:31-32): the state initialize_prognostics wrote is still the resident one
        {
            const int rc = speedy_check_diagnostics(ctx, 1, diag.data());
            if (rc < 0) throw std::runtime_error(speedy_last_error());
            if (verbose || rc > 0) print_diagnostics(0, diag.data(), kx);     // mod(0, nstdia) == 0
            if (rc > 0) return finish(1);
            if (write) {
                if (member < 0)
                    for (int e = lo; e < hi; e++)
                        if (mkdir(member_dir(e).c_str(), 0777) && errno != EEXIST) throw std::runtime_error("cannot create " + member_dir(e));
                emit(0);
            }
        }
        // ---- main loop (speedy.f90:27-54): enqueue up to the next event; the range guard is polled at every print, every 30 simulated
        // days and at the end (it is sticky and remembers the failing step)
        long long since_poll = 0;
        while (done < total) {
            const long long to_out = write ? nml->nsteps_out - done % nml->nsteps_out : total - done;
            const long long to_dia = verbose ? nml->nstdia - done % nml->nstdia : total - done;
            long long chunk = std::min(std::min(to_out, to_dia), total - done);
            if (chunk > (1 << 20)) chunk = 1 << 20;
            if (speedy_enqueue_steps(ctx, (int)chunk)) throw std::runtime_error(speedy_last_error());
            for (long long k = 0; k < chunk; k++) spd::cal_advance(hc);
            done += chunk;
            since_poll += chunk;
            const bool print = verbose && done % nml->nstdia == 0;
            if (print || done == total || since_poll >= 30LL * nsteps_day) {
                const int rc = speedy_finish(ctx);
                if (rc < 0) throw std::runtime_error(speedy_last_error());
                if (rc > 0) return range_failure();
                since_poll = 0;
            }
            if (print) {
                // the numbers check_diagnostics(vor(:,:,:,2), ...) of this step left on the device (speedy.f90:41): nothing is recomputed
                if (speedy_range_failure(ctx, nullptr, diag.data()) < 0) throw std::runtime_error(speedy_last_error());
                print_diagnostics(done, diag.data(), kx);
            }
            if (write && done % nml->nsteps_out == 0) emit(done);
        }
        if (speedy_output_drain(ctx)) throw std::runtime_error(speedy_last_error());
        if (speedy_model_date(ctx, now, &model_step)) throw std::runtime_error(speedy_last_error());
        if (!same_date(now, nml->end_datetime)) throw std::runtime_error("speedy_main_loop: the device calendar did not arrive at the end date");
    } catch (const std::exception& e_) { spd::last_error() = e_.what(); return finish(-1); }
    return finish(0);
}

}  // extern "C"
