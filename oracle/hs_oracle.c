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
    /* one filter2D(BORDER_CONSTANT) pass, :60-61.  `pad` is a zeroed (H+w-1) x (W+w-1) scratch  */    \
    /* plane: the image is copied to its centre, so out-of-image taps read 0 (= BORDER_CONSTANT)   */    \
    /* and every pixel still accumulates kf*tap in dy-major / dx-minor order.                      */    \
    static void box_##SUF(const REAL* restrict f, REAL* restrict out, REAL* restrict pad, int H, int W, int w) {              \
        const int a = w - w / 2 - 1;                                             /* :54 */         \
        const REAL kf = (REAL)1 / (REAL)(w * w);                                 /* :53 */         \
        const size_t PW = (size_t)W + w - 1;                                                       \
        _Pragma("omp parallel for schedule(static)")                                               \
        for (int y = 0; y < H; ++y)                                                                \
            memcpy(pad + (size_t)(y + a) * PW + a, f + (size_t)y * W, (size_t)W * sizeof(REAL));   \
        _Pragma("omp parallel for schedule(static)")                                               \
        for (int y = 0; y < H; ++y) {                                                              \
            REAL* restrict o = out + (size_t)y * W;                                                    \
            for (int x = 0; x < W; ++x) o[x] = 0;                                                  \
            for (int dy = 0; dy < w; ++dy) {                                                       \
                const REAL* restrict r = pad + (size_t)(y + dy) * PW;                                    \
                for (int dx = 0; dx < w; ++dx)                                                     \
                    for (int x = 0; x < W; ++x) o[x] += kf * r[x + dx];                            \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* hornSchunck.cpp:43-75.  u, v: H*W outputs.  Returns 0, or -1 if out of memory. */           \
    int hs_oracle_flow_##SUF(const uint8_t* prev, const uint8_t* next, int H, int W, int w,        \
                             int iters, double alpha, REAL* u, REAL* v) {                          \
        size_t n = (size_t)H * W;                                                                  \
        size_t np_ = ((size_t)H + w - 1) * ((size_t)W + w - 1);                                    \
        REAL* buf = (REAL*)malloc((n * 6 + np_) * sizeof(REAL));                                   \
        if (!buf) return -1;                                                                       \
        REAL* pad = buf + 6 * n;                                                                   \
        memset(pad, 0, np_ * sizeof(REAL));                                                        \
        REAL *gx = buf, *gy = buf + n, *gt = buf + 2 * n, *den = buf + 3 * n;                      \
        REAL *ua = buf + 4 * n, *va = buf + 5 * n;                                                 \
        hs_oracle_gradients_##SUF(prev, next, H, W, gx, gy, gt);                 /* :46 */         \
        memset(u, 0, n * sizeof(REAL));                                          /* :49 */         \
        memset(v, 0, n * sizeof(REAL));                                          /* :50 */         \
        const REAL a2 = (REAL)alpha * (REAL)alpha;                                                 \
        for (size_t i = 0; i < n; ++i) den[i] = a2 + gx[i] * gx[i] + gy[i] * gy[i];                \
        for (int it = 0; it < iters; ++it) {                                     /* :56 */         \
            box_##SUF(u, ua, pad, H, W, w);                                          /* :60 */         \
            box_##SUF(v, va, pad, H, W, w);                                          /* :61 */         \
            _Pragma("omp parallel for schedule(static)")                                           \
            for (size_t i = 0; i < n; ++i) {                                                       \
                REAL c = (gx[i] * ua[i] + gy[i] * va[i] + gt[i]) / den[i];       /* :63-68 */      \
                u[i] = ua[i] - gx[i] * c;                                        /* :69,72 */      \
                v[i] = va[i] - gy[i] * c;                                        /* :70,73 */      \
            }                                                                                      \
        }                                                                                          \
        free(buf);                                                                                 \
        return 0;                                                                                  \
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
