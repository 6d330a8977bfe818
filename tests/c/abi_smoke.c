/* include/hs.h must be a plain C header: this file is compiled as C99 with -Wall -Werror -pedantic
 * and linked against libhs_b200.so.  It checks argument validation and, on a machine without a GPU,
 * that the library refuses to run (no CPU fallback).  Exit code 0 = as expected. */
#include <stdio.h>
#include <string.h>

#include "hs.h"

int main(void) {
    hs_config cfg;
    hs_ctx* ctx = NULL;
    int rc;
    if (hs_version() != HS_VERSION) return 10;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (uint32_t)sizeof cfg;
    cfg.width = 0; cfg.height = 8; cfg.window_size = 5; cfg.max_iterations = 100; cfg.alpha = 1.0;
    rc = hs_create(&cfg, &ctx);
    if (rc != HS_ERR_INVALID_ARG || ctx != NULL) return 11;
    if (strstr(hs_last_error(NULL), "width") == NULL) return 12;
    cfg.width = 8;
    cfg.device = -1;
    rc = hs_create(&cfg, &ctx);
    if (rc == HS_OK) {                         /* a GPU is present: one tiny solve through the C ABI */
        unsigned char prev[64], next[64];
        double u[64], v[64];
        int i;
        for (i = 0; i < 64; ++i) { prev[i] = (unsigned char)(i * 3); next[i] = (unsigned char)(i * 3 + (i & 7)); }
        rc = hs_solve(ctx, prev, 8, 0, next, 8, 0, u, 8 * sizeof(double), 0, v, 8 * sizeof(double), 0, HS_F64);
        hs_destroy(ctx);
        if (rc != HS_OK) return 13;
        printf("gpu solve ok u[9]=%g\n", u[9]);
        return 0;
    }
    if (rc != HS_ERR_CUDA || strstr(hs_last_error(NULL), "no CPU fallback") == NULL) return 14;
    printf("no gpu: %s\n", hs_last_error(NULL));
    return 0;
}
