// speedy_b200: the reference's executable (`program speedy`, speedy.f90:1-54) on the B200 library.  Run it where the
// reference's `speedy` would be run: it reads ./namelist.nml (params.f90:62-67, date.f90:66-71; a missing file keeps the
// defaults), initialises the model from the reference's boundary files in the working directory (or --bc: a directory of them / a
// packed file), and writes yyyymmddhhmm.nc files into the working
// directory while printing the reference's start-up lines and diagnostics.  Plain C ABI only (include/speedy_b200.h).
//
//   speedy_b200 [--namelist FILE] [--bc FILE] [--out DIR] [--trunc 30|47] [--steps-per-day N] [--members M] [--member E]
//               [--sppt] [--seed S] [--device D] [--precision 0|1] [--no-output] [--quiet]
#include "../../include/speedy_b200.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static int fail(const char* what) {
    fprintf(stderr, "speedy_b200: %s: %s\n", what, speedy_last_error());
    return 2;
}

int main(int argc, char** argv) {
    std::string namelist = "namelist.nml", bc, out = ".";
    speedy_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.trunc = 30; cfg.kx = 8; cfg.ntr = 1; cfg.nmembers = 1;
    int member = 0, verbose = 1, write = 1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "speedy_b200: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--namelist") namelist = val();
        else if (a == "--bc") bc = val();
        else if (a == "--out") out = val();
        else if (a == "--trunc") cfg.trunc = atoi(val());
        else if (a == "--steps-per-day") cfg.nsteps = atoi(val());
        else if (a == "--members") cfg.nmembers = atoi(val());
        else if (a == "--member") member = atoi(val());
        else if (a == "--sppt") cfg.sppt_on = 1;
        else if (a == "--seed") cfg.seed = strtoull(val(), nullptr, 10);
        else if (a == "--device") cfg.device = atoi(val());
        else if (a == "--precision") cfg.precision = atoi(val());
        else if (a == "--no-output") write = 0;
        else if (a == "--quiet") verbose = 0;
        else { fprintf(stderr, "speedy_b200: unknown option %s\n", a.c_str()); return 2; }
    }
    if (bc.empty()) {
        // the reference's run directory holds its boundary files side by side (run.sh links them): read them where they are
        const char* env = getenv("SPEEDY_BC");
        FILE* probe = env ? nullptr : fopen("surface.nc", "rb");
        if (probe) fclose(probe);
        bc = env ? env : probe ? "." : (cfg.trunc == 30 ? "data/bc_t30.bin" : "data/bc_t47.bin");
    }
    speedy_namelist nml;
    if (speedy_read_namelist(namelist.c_str(), &nml)) return fail("namelist");
    if (verbose) {
        printf("\n  speedy.f90 main loop on the speedy_b200 library (T%d, %d member%s)\n\n", cfg.trunc, cfg.nmembers, cfg.nmembers == 1 ? "" : "s");
        // the lines initialize_params / initialize_date print (params.f90:69-70, date.f90:76-81)
        printf("nsteps_out (frequency of output)  = %5d\n", nml.nsteps_out);
        printf("nstdia (frequency of diagnostics) = %5d\n", nml.nstdia);
        const int* s = nml.start_datetime; const int* e = nml.end_datetime;
        printf("Start date: %4d/%02d/%02d %02d:%02d\n", s[0], s[1], s[2], s[3], s[4]);
        printf("  End date: %4d/%02d/%02d %02d:%02d\n", e[0], e[1], e[2], e[3], e[4]);
        fflush(stdout);
    }
    speedy_ctx* ctx = nullptr;
    if (speedy_create(&cfg, &ctx)) return fail("speedy_create");
    const int* s = nml.start_datetime;
    if (speedy_model_init(ctx, bc.c_str(), s[0], s[1], s[2], s[3], s[4])) return fail("speedy_model_init");
    long long steps = 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = speedy_main_loop(ctx, &nml, write ? out.c_str() : nullptr, member, verbose, &steps);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc < 0) return fail("speedy_main_loop");
    if (rc > 0) {
        fprintf(stderr, "Model variables out of accepted range\n");      // diagnostics.f90:68 `stop '...'`
        speedy_destroy(ctx);
        return 1;
    }
    int info[4] = {1, 36, 0, 0};
    speedy_run_info(ctx, info);
    if (verbose) printf("%lld steps (%.2f simulated days x %d member%s) in %.3f s\n", steps, (double)steps / info[1], info[0], info[0] == 1 ? "" : "s", secs);
    speedy_destroy(ctx);
    return 0;
}
