/*
 * Plain-C CPU oracle for the Horn-Schunck hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library
 * built from this file (oracle/_build/libhs_oracle.so).  Nothing under cpp-optical-flow_b200/
 * links or dlopens it.
 *
 * It restates /root/reference/HornSchunckOF/hornSchunck.cpp operator by operator, with the
 * OpenCV calls written out as loops (OpenCV 4.4.0 core/imgproc is the un-vendored dependency
 * that holds the arithmetic; see oracle/hs_oracle.py for the cv2 twin and the golden-PNG pin):
 *   getGradients :19-41  Sobel(prev, ksize 3, scale 1, BORDER_REFLECT_101) :27-28, next-prev :39
 *   getFlow      :43-75  u=v=0 :49-50; box kernel ones(w,w)/w^2 with anchor w-w/2-1 :53-54;
 *                        per sweep :56-74  ubar=filter2D(u) (BORDER_CONSTANT) :60-61,
 *                        c=(gx*ubar+gy*vbar+gt)/(alpha^2+gx^2+gy^2) :63-68, u=ubar-gx*c :69-73
 * Tap order and rounding follow the NumPy twin (acc += fl(1/w^2)*tap, dy-major/dx-minor), which is
 * bit-identical to cv2.filter2D in fp64 for w <= 7.  Build with -ffp-contract=off: a fused
 * multiply-add would change the last bit.
 *
 * The same body is instantiated for double (the oracle proper) and float (to separate
 * "fp32 rounding" from "kernel bug" when a GPU parity test fails).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

#define DEFINE_ORACLE(SUF, REAL)                                                                   \
    /* hornSchunck.cpp:19-41 */                                                                    \
    void hs_oracle_gradients_##SUF(const uint8_t* prev, const uint8_t* next, int H, int W,         \
                                   REAL* gx, REAL* gy, REAL* gt) {                                 \
        _Pragma("omp parallel for schedule(static)")                                               \
        for (int y = 0; y < H; ++y) {                                                              \
            const uint8_t* r0 = prev + (size_t)reflect101(y - 1, H) * W;                           \
            const uint8_t* r1 = prev + (size_t)y * W;                                              \
            const uint8_t* r2 = prev + (size_t)reflect101(y + 1, H) * W;                           \
            for (int x = 0; x < W; ++x) {                                                          \
                int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);                          \
                REAL a = r0[xm], b = r0[x], c = r0[xp];                                            \
                REAL d = r1[xm], f = r1[xp];                                                       \
                REAL g = r2[xm], h = r2[x], i = r2[xp];                                            \
                size_t o = (size_t)y * W + x;                                                      \
                gx[o] = (c + (REAL)2 * f + i) - (a + (REAL)2 * d + g);           /* :27 */         \
                gy[o] = (g + (REAL)2 * h + i) - (a + (REAL)2 * b + c);           /* :28 */         \
                gt[o] = (REAL)next[o] - (REAL)r1[x];                             /* :39 */         \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* Gradients of frames that are already REAL (the reference converts ANY input depth with    */    \
    /* convertTo(CV_64FC1) first, :23-24; for 8-bit frames this gives the same values as above). */    \
    void hs_oracle_gradients_real_##SUF(const REAL* prev, const REAL* next, int H, int W,          \
                                        REAL* gx, REAL* gy, REAL* gt) {                            \
        _Pragma("omp parallel for schedule(static)")                                               \
        for (int y = 0; y < H; ++y) {                                                              \
            const REAL* r0 = prev + (size_t)reflect101(y - 1, H) * W;                              \
            const REAL* r1 = prev + (size_t)y * W;                                                 \
            const REAL* r2 = prev + (size_t)reflect101(y + 1, H) * W;                              \
            for (int x = 0; x < W; ++x) {                                                          \
                int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);                          \
                REAL a = r0[xm], b = r0[x], c = r0[xp];                                            \
                REAL d = r1[xm], f = r1[xp];                                                       \
                REAL g = r2[xm], h = r2[x], i = r2[xp];                                            \
                size_t o = (size_t)y * W + x;                                                      \
                gx[o] = (c + (REAL)2 * f + i) - (a + (REAL)2 * d + g);           /* :27 */         \
                gy[o] = (g + (REAL)2 * h + i) - (a + (REAL)2 * b + c);           /* :28 */         \
                gt[o] = next[o] - r1[x];                                         /* :39 */         \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* The T sweeps of :56-74 on given gradient planes.  u and v live in zero-bordered planes of  */    \
    /* (H+w-1) x (W+w-1) (image at offset (a, a)), so out-of-image taps read 0 = BORDER_CONSTANT  */    \
    /* (:60-61); two such pairs ping-pong.  One pass per sweep and row: both box means (every     */    \
    /* tap multiplied by kf = fl(1/w^2) and accumulated dy-major / dx-minor, exactly the order of */    \
    /* the NumPy twin and of cv2.filter2D for w <= 7), then the update :63-73.                    */    \
    static int sweeps_##SUF(const REAL* gx, const REAL* gy, const REAL* gt, int H, int W, int w,   \
                            int iters, double alpha, REAL* u, REAL* v) {                           \
        const size_t n = (size_t)H * W;                                                            \
        const int a = w - w / 2 - 1;                                             /* :54 */         \
        const REAL kf = (REAL)1 / (REAL)(w * w);                                 /* :53 */         \
        const size_t PW = (size_t)W + w - 1, PH = (size_t)H + w - 1;                               \
        REAL* buf = (REAL*)calloc(4 * PW * PH + n, sizeof(REAL));                /* :49-50 */      \
        if (!buf) return -1;                                                                       \
        REAL* pu[2] = {buf, buf + PW * PH};                                                        \
        REAL* pv[2] = {buf + 2 * PW * PH, buf + 3 * PW * PH};                                      \
        REAL* den = buf + 4 * PW * PH;                                                             \
        const REAL a2 = (REAL)alpha * (REAL)alpha;                                                 \
        for (size_t i = 0; i < n; ++i) den[i] = a2 + gx[i] * gx[i] + gy[i] * gy[i];  /* :65-66,68 */ \
        int cur = 0;                                                                               \
        for (int it = 0; it < iters; ++it, cur ^= 1) {                           /* :56 */         \
            const REAL* restrict su = pu[cur];                                                     \
            const REAL* restrict sv = pv[cur];                                                     \
            REAL* restrict du = pu[cur ^ 1];                                                       \
            REAL* restrict dv = pv[cur ^ 1];                                                       \
            _Pragma("omp parallel for schedule(static)")                                           \
            for (int y = 0; y < H; ++y) {                                                          \
                REAL ua[W], va[W];                                                                 \
                for (int x = 0; x < W; ++x) ua[x] = va[x] = 0;                                     \
                for (int dy = 0; dy < w; ++dy) {                                                   \
                    const REAL* restrict ru = su + (size_t)(y + dy) * PW;                          \
                    const REAL* restrict rv = sv + (size_t)(y + dy) * PW;                          \
                    for (int dx = 0; dx < w; ++dx)                                                 \
                        for (int x = 0; x < W; ++x) {                                              \
                            ua[x] += kf * ru[x + dx];                            /* :60 */         \
                            va[x] += kf * rv[x + dx];                            /* :61 */         \
                        }                                                                          \
                }                                                                                  \
                const size_t o = (size_t)y * W, po = (size_t)(y + a) * PW + a;                     \
                for (int x = 0; x < W; ++x) {                                                      \
                    REAL c = (gx[o + x] * ua[x] + gy[o + x] * va[x] + gt[o + x]) / den[o + x]; /* :63-68 */ \
                    du[po + x] = ua[x] - gx[o + x] * c;                          /* :69,72 */      \
                    dv[po + x] = va[x] - gy[o + x] * c;                          /* :70,73 */      \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
        for (int y = 0; y < H; ++y) {                                                              \
            memcpy(u + (size_t)y * W, pu[cur] + (size_t)(y + a) * PW + a, (size_t)W * sizeof(REAL)); \
            memcpy(v + (size_t)y * W, pv[cur] + (size_t)(y + a) * PW + a, (size_t)W * sizeof(REAL)); \
        }                                                                                          \
        free(buf);                                                                                 \
        return 0;                                                                                  \
    }                                                                                              \
                                                                                                   \
    /* hornSchunck.cpp:43-75.  u, v: H*W outputs.  Returns 0, or -1 if out of memory. */           \
    int hs_oracle_flow_##SUF(const uint8_t* prev, const uint8_t* next, int H, int W, int w,        \
                             int iters, double alpha, REAL* u, REAL* v) {                          \
        size_t n = (size_t)H * W;                                                                  \
        REAL* g = (REAL*)malloc(3 * n * sizeof(REAL));                                             \
        if (!g) return -1;                                                                         \
        hs_oracle_gradients_##SUF(prev, next, H, W, g, g + n, g + 2 * n);        /* :46 */         \
        int rc = sweeps_##SUF(g, g + n, g + 2 * n, H, W, w, iters, alpha, u, v);                   \
        free(g);                                                                                   \
        return rc;                                                                                 \
    }                                                                                              \
                                                                                                   \
    /* the same for frames of any depth, already converted to REAL as :23-24 does */               \
    int hs_oracle_flow_real_##SUF(const REAL* prev, const REAL* next, int H, int W, int w,         \
                                  int iters, double alpha, REAL* u, REAL* v) {                     \
        size_t n = (size_t)H * W;                                                                  \
        REAL* g = (REAL*)malloc(3 * n * sizeof(REAL));                                             \
        if (!g) return -1;                                                                         \
        hs_oracle_gradients_real_##SUF(prev, next, H, W, g, g + n, g + 2 * n);                     \
        int rc = sweeps_##SUF(g, g + n, g + 2 * n, H, W, w, iters, alpha, u, v);                   \
        free(g);                                                                                   \
        return rc;                                                                                 \
    }

DEFINE_ORACLE(f64, double)
DEFINE_ORACLE(f32, float)

void hs_oracle_set_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int hs_oracle_max_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
