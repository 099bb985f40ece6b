/* The C ABI driven from COMPILED C exactly as the Fortran bind(C) wrappers of fortran/speedy_b200_c.f90 drive it (the image has
 * no Fortran compiler): the derived type by reference, scalars by value, NUL-terminated character arrays, the module arrays as
 * one contiguous state vector.  Test helper only (tests/test_abi_c.py).
 *   abi_drive layout                 -> sizeof / offsetof of speedy_cfg (no GPU needed)
 *   abi_drive run <bc.bin> <nsteps>  -> create, model_init, state from the resident fields, run_steps_host, date, checksum, destroy */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/speedy_b200.h"

static void check(int rc, const char* what) {
    if (rc > 0) { printf("STOP Model variables out of accepted range\n"); exit(3); }        /* b200_check */
    if (rc < 0) { printf("speedy_b200: %s failed: %s\n", what, speedy_last_error()); exit(2); }
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) {
        printf("sizeof %zu\n", sizeof(speedy_cfg));
        printf("trunc %zu\nkx %zu\nntr %zu\nnmembers %zu\ndevice %zu\nsppt_on %zu\nseed %zu\nmember_offset %zu\nnsteps %zu\nprecision %zu\n",
               offsetof(speedy_cfg, trunc), offsetof(speedy_cfg, kx), offsetof(speedy_cfg, ntr), offsetof(speedy_cfg, nmembers), offsetof(speedy_cfg, device),
               offsetof(speedy_cfg, sppt_on), offsetof(speedy_cfg, seed), offsetof(speedy_cfg, member_offset), offsetof(speedy_cfg, nsteps), offsetof(speedy_cfg, precision));
        return 0;
    }
    if (argc < 4 || strcmp(argv[1], "run")) { fprintf(stderr, "usage: abi_drive layout | run <bc.bin> <nsteps>\n"); return 1; }
    /* cfg = speedy_cfg(trunc, kx, ntr, 1, 0, 0, 0_c_long_long, 0, nsteps, 0)   */
    speedy_cfg cfg = {30, 8, 1, 1, 0, 0, 0ull, 0, 0, 0};
    speedy_ctx* ctx = NULL;
    check(speedy_create(&cfg, &ctx), "create");
    check(speedy_model_init(ctx, argv[2], 1982, 1, 1, 0, 0), "init");
    int dims[8];
    check(speedy_dims(ctx, dims), "dims");
    const size_t n = speedy_state_len(ctx);
    double* state = (double*)malloc(n * sizeof(double));
    float* out = (float*)malloc(speedy_output_len(ctx) * sizeof(float));
    /* pack_state: [vor, div, t, tr, ps], both time levels each */
    const char* names[5] = {"vor", "div", "t", "tr", "ps"};
    const size_t n3 = (size_t)2 * dims[6] * dims[5] * dims[4] * 2, n2 = (size_t)2 * dims[6] * dims[5] * 2;
    size_t o = 0;
    for (int i = 0; i < 5; i++) { const size_t len = i < 4 ? n3 : n2; check(speedy_get_field(ctx, names[i], state + o, len), names[i]); o += len; }
    if (o != n) { printf("state length mismatch %zu %zu\n", o, n); return 4; }
    check(speedy_run_steps_host(ctx, state, n, atoi(argv[3]), out), "run_steps_host");
    int ymdhm[5]; long long step;
    check(speedy_model_date(ctx, ymdhm, &step), "date");
    double s1 = 0.0, s2 = 0.0;
    for (size_t i = 0; i < n; i++) { s1 += state[i]; s2 += state[i] * state[i]; }
    printf("date %d %d %d %d %d step %lld\n", ymdhm[0], ymdhm[1], ymdhm[2], ymdhm[3], ymdhm[4], step);
    printf("checksum %.17g %.17g\n", s1, s2);
    printf("t_lowest_mean %.9g\n", (double)out[(size_t)(2 * dims[4] + dims[4] - 1) * dims[1] * dims[3]]);
    check(speedy_destroy(ctx), "destroy");
    free(state); free(out);
    return 0;
}
