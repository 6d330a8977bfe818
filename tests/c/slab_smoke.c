/* A plain-C caller drives a multi-GPU solve through the C ABI: ONE hs_ctx over
 * `n` row slabs (hs_config.num_devices / device_ids / decomposition), hs_solve, and the result must be
 * bit-identical to a single-device context.  Usage: slab_smoke <n> <same|distinct> [peer|nccl]
 *   same     - the n slabs all on device 0 (one cooperative launch: runs on a 1-GPU box)
 *   distinct - devices 0..n-1 (needs n GPUs; halos cross NVLink from inside the kernel, or via NCCL)
 * Exit 0 = bit-identical; 77 = not enough devices (skipped). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hs.h"

#define W 517
#define H 1030

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 2;
    int distinct = argc > 2 && strcmp(argv[2], "distinct") == 0;
    int nccl = argc > 3 && strcmp(argv[3], "nccl") == 0;
    static unsigned char prev[H * W], next[H * W];
    float* u1 = malloc(sizeof(float) * H * W); float* v1 = malloc(sizeof(float) * H * W);
    float* u2 = malloc(sizeof(float) * H * W); float* v2 = malloc(sizeof(float) * H * W);
    int32_t devs[8];
    hs_config cfg;
    hs_ctx *one = NULL, *many = NULL;
    hs_timing tm;
    unsigned s = 12345u;
    int i, rc;
    if (n < 2 || n > 8 || !u1 || !v1 || !u2 || !v2) return 2;
    for (i = 0; i < H * W; ++i) {
        s = s * 1664525u + 1013904223u;
        prev[i] = (unsigned char)(s >> 24);
        next[i] = (unsigned char)((s >> 24) + ((s >> 12) & 15) > 255 ? 255 : (s >> 24) + ((s >> 12) & 15));
    }
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (uint32_t)sizeof cfg;
    cfg.width = W; cfg.height = H; cfg.window_size = 3; cfg.max_iterations = 57; cfg.alpha = 1.0; cfg.temporal_k = 4;
    cfg.device = 0;
    rc = hs_create(&cfg, &one);
    if (rc != HS_OK) { printf("single: %s\n", hs_last_error(NULL)); return 3; }
    rc = hs_solve(one, prev, W, 0, next, W, 0, u1, W * sizeof(float), 0, v1, W * sizeof(float), 0, HS_F32);
    if (rc != HS_OK) { printf("single solve: %s\n", hs_last_error(one)); return 4; }
    hs_destroy(one);

    for (i = 0; i < n; ++i) devs[i] = distinct ? i : 0;
    cfg.num_devices = n; cfg.device_ids = devs;
    cfg.decomposition = HS_DECOMP_ROW_SLAB;
    cfg.exchange = nccl ? HS_EXCHANGE_NCCL : HS_EXCHANGE_PEER;
    rc = hs_create(&cfg, &many);
    if (rc != HS_OK) {
        printf("group create: %d %s\n", rc, hs_last_error(NULL));
        return (distinct && strstr(hs_last_error(NULL), "out of range")) ? 77 : (rc == HS_ERR_NCCL ? 78 : 5);
    }
    rc = hs_solve(many, prev, W, 0, next, W, 0, u2, W * sizeof(float), 0, v2, W * sizeof(float), 0, HS_F32);
    if (rc != HS_OK) { printf("group solve: %d %s\n", rc, hs_last_error(many)); return 6; }
    rc = hs_solve(many, prev, W, 0, next, W, 0, u2, W * sizeof(float), 0, v2, W * sizeof(float), 0, HS_F32);   /* reusable */
    if (rc != HS_OK) return 7;
    hs_get_timing(many, &tm);
    hs_destroy(many);
    if (memcmp(u1, u2, sizeof(float) * H * W) != 0 || memcmp(v1, v2, sizeof(float) * H * W) != 0) {
        printf("MISMATCH\n");
        return 1;
    }
    printf("ok: %d row slabs (%s, %s) bit-identical to one device, k=%d, %d launches, u[5000]=%g\n", n,
           distinct ? "distinct devices" : "one device", nccl ? "NCCL exchange" : "in-kernel exchange",
           (int)tm.temporal_k, (int)tm.launches, (double)u1[5000]);
    free(u1); free(v1); free(u2); free(v2);
    return 0;
}
