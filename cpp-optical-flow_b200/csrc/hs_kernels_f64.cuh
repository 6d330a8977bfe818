// HS_PREC_F64: the reference's own arithmetic, operation by operation, in fp64 on the device.
//
// /root/reference/HornSchunckOF/hornSchunck.cpp runs everything in CV_64FC1: convertTo(CV_64FC1) of frames
// of ANY depth (:23-24), Sobel (:27-28), next - prev (:39), and per sweep filter2D with the kernel
// ones(w,w)/w^2 (:53,60-61), c = (gx*ubar + gy*vbar + gt) / (alpha^2 + gx^2 + gy^2) (:63-68),
// u = ubar - gx*c (:69-73).  These kernels repeat exactly that sequence of IEEE operations with explicit
// __d*_rn intrinsics (no contraction, no re-association):
//   box mean:  acc = 0;  acc += fl(1/w^2) * tap   over dy-major / dx-minor taps, zeros outside
//              (bit-identical to cv2.filter2D in fp64 for w <= 7)
//   update:    c = ((gx*ua + gy*va) + gt) / den;   u = ua - gx*c;   v = va - gy*c
//   den:       (alpha*alpha + gx*gx) + gy*gy        computed once (the reference recomputes it, same value)
// The result is BIT-IDENTICAL to the fp64 CPU restatements the tests check against (C / NumPy for any w; OpenCV's own
// filter2D / Sobel for 8-bit frames and w <= 7).
// One pixel per thread, taps through L1/L2: this is the A/B diagnostic for fp32 rounding and the
// path for frames that are not 8-bit; the throughput path is the fp32 fused kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hs_kernels.cuh"

namespace hs {

// K1 (fp64): getGradients :19-41 for frames of element type T, + the denominator of :65-66,68
template <typename T>
__global__ void __launch_bounds__(256)
k_grad_f64(const T* __restrict__ prev, const T* __restrict__ next, size_t fpitch_bytes, size_t fimg_bytes,
           double* __restrict__ gx, double* __restrict__ gy, double* __restrict__ gt, double* __restrict__ den,
           int W, int H, int pitch, long long plane, double alpha2) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const char* P = reinterpret_cast<const char*>(prev) + (size_t)blockIdx.z * fimg_bytes;
    const char* N = reinterpret_cast<const char*>(next) + (size_t)blockIdx.z * fimg_bytes;
    auto px = [&](const char* base, int yy, int xx) {                      // convertTo(CV_64FC1) :23-24
        return (double)reinterpret_cast<const T*>(base + (size_t)yy * fpitch_bytes)[xx];
    };
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const double a = px(P, ym, xm), b = px(P, ym, x), c = px(P, ym, xp);
    const double d = px(P, y, xm), f = px(P, y, xp);
    const double g = px(P, yp, xm), h = px(P, yp, x), i = px(P, yp, xp);
    const double sx = __dsub_rn(__dadd_rn(__dadd_rn(c, __dmul_rn(2.0, f)), i), __dadd_rn(__dadd_rn(a, __dmul_rn(2.0, d)), g));   // :27
    const double sy = __dsub_rn(__dadd_rn(__dadd_rn(g, __dmul_rn(2.0, h)), i), __dadd_rn(__dadd_rn(a, __dmul_rn(2.0, b)), c));   // :28
    const double st = __dsub_rn(px(N, y, x), px(P, y, x));                                                                        // :39
    const size_t o = (size_t)blockIdx.z * plane + (size_t)y * pitch + x;
    gx[o] = sx; gy[o] = sy; gt[o] = st;
    den[o] = __dadd_rn(__dadd_rn(alpha2, __dmul_rn(sx, sx)), __dmul_rn(sy, sy));                                                  // :65-66,68
}

// K2 (fp64): one sweep of :56-74, any window
__global__ void __launch_bounds__(256)
k_jacobi_f64(const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ un, double* __restrict__ vn,
             const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ gt,
             const double* __restrict__ den, int W, int H, int pitch, long long plane, int w, int a, double kf) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t base = (size_t)blockIdx.z * plane;
    const double* U = u + base;
    const double* V = v + base;
    double ua = 0.0, va = 0.0;
    for (int dy = 0; dy < w; ++dy) {                                       // filter2D, BORDER_CONSTANT :60-61
        const int yy = y + dy - a;
        const bool yin = yy >= 0 && yy < H;
        const size_t ro = (size_t)(yin ? yy : 0) * pitch;
        for (int dx = 0; dx < w; ++dx) {
            const int xx = x + dx - a;
            const bool in = yin && xx >= 0 && xx < W;
            const double tu = in ? U[ro + xx] : 0.0, tv = in ? V[ro + xx] : 0.0;
            ua = __dadd_rn(ua, __dmul_rn(kf, tu));
            va = __dadd_rn(va, __dmul_rn(kf, tv));
        }
    }
    const size_t o = base + (size_t)y * pitch + x;
    const double ix = gx[o], iy = gy[o];
    const double c = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(ix, ua), __dmul_rn(iy, va)), gt[o]), den[o]);   // :63-68
    un[o] = __dsub_rn(ua, __dmul_rn(ix, c));                                                                  // :69,72
    vn[o] = __dsub_rn(va, __dmul_rn(iy, c));                                                                  // :70,73
}

__global__ void __launch_bounds__(256)
k_narrow_f64(const double* __restrict__ a, float* __restrict__ o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = i; p < n; p += stride) o[p] = (float)a[p];
}

}  // namespace hs
