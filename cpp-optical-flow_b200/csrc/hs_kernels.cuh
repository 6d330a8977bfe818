// Horn-Schunck device kernels for sm_100a (B200).  See DESIGN.md for the data layout.
//
// Reference semantics: /root/reference/HornSchunckOF/hornSchunck.cpp
//   K1  k_grad_coeff      <- getGradients :19-41 + the loop-invariant denominator of :65-66,68
//   K2  k_jacobi_generic  <- one sweep of the hot loop :56-74, any windowSize
//   K3  k_jacobi_tile     <- k fused sweeps of :56-74 per HBM round trip (temporal blocking)
//   K5  k_widen / k_unpack_grad <- CV_64FC1 outputs (:49-50; plotFlow.cpp:72-75 reads double)
//
// Canonical arithmetic (every kernel uses exactly this, so all variants are bit-identical):
//   window sums are PAIRED: along a row, neighbours are first added in pairs that start at EVEN
//   absolute columns, a window that starts at an odd column / ends at an even one contributes that
//   element alone, and the terms are then added left to right, zeros outside the image:
//       w=3, x even:  t[x-1] + (t[x] + t[x+1])          w=3, x odd:  (t[x-1] + t[x]) + t[x+1]
//       w=5, x even:  ((t[x-2]+t[x-1]) + (t[x]+t[x+1])) + t[x+2]      x odd:  (t[x-2] + (..)) + (..)
//   the same rule combines the row sums down a column (pairs start at even absolute rows).  A pair
//   sum is shared by the two windows that contain it: 1.5 instead of 2 adds per pixel and
//   direction for w=3, 2.5 instead of 4 for w=5.
//   ubar = S_u * fl(1/w^2);  vbar likewise - except for w = 3, where the column window is always ONE pair of rows and
//   ONE single row: ubar = fma(single, kf, kf * pair) with kf = fl(1/9), so that the fused kernel multiplies each
//   shared pair sum once for the two rows that use it (6 % less FP32-pipe work per sweep, +3 % throughput)
//   The update  u' = ubar - Ix (Ix ubar + Iy vbar + It) / den,  den = alpha^2 + Ix^2 + Iy^2  (hornSchunck.cpp:63-73) is
//   evaluated with NORMALISED coefficients  P = Ix * s, Q = Iy * s, R = It * s,  s = 1 / sqrt(den):
//       t = fma(P, ubar, fma(Q, vbar, R));  u' = fma(-P, t, ubar);  v' = fma(-Q, t, vbar)
//   - algebraically the same (P t = Ix (..) / den), one multiply and one register per pixel less than the
//   (Ix, Iy, It, 1/den) form, same accuracy against the fp64 oracle (4.1e-5 px on the hardest case either way).
//   s = fl(1 / sqrt(fl(alpha^2) + (Ix^2 + Iy^2))) is rounded once (IEEE rsqrt) in K1 and stored in the `inv` plane;
//   P, Q, R are single IEEE multiplies wherever a kernel unpacks a pixel.
// Explicit __f*_rn intrinsics keep the compiler from re-associating or contracting differently
// in different kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace hs {

// Geometry shared by all kernels.  All planes of a context use the same pixel pitch.
struct Geom {
    int W;          // image width in pixels
    int H;          // rows held in the buffer
    int pitch;      // pixels per buffer row (multiple of 32)
    long long plane;  // pixels between consecutive pairs of the batch (pitch * H)
    int oy0, oy1;   // rows this context produces: [oy0, oy1)
    int grow0;      // image row of buffer row 0 (row slabs); only its parity matters (pair alignment)
};

// Gradient coefficients are exact small integers (|Ix|,|Iy| <= 1020, |It| <= 255), so the three of
// them share one 32-bit word as BIASED unsigned fields: Ix+1024 in bits 31..21, Iy+1024 in 20..10,
// It+512 in 9..0.  Unpacking is the classic magic-number conversion: OR the field into the mantissa
// of 2^23 and subtract (2^23 + bias) - shift/LOP3/FADD on the main pipes, no I2F on the XU pipe.
// The all-zero word TMA fills in for out-of-image pixels decodes to (-1024, -1024, -512); those
// pixels are forced to 0 after every sweep (masked path), so the value is never used.
__device__ __forceinline__ uint32_t pack_coef(int gx, int gy, int gt) {
    return ((uint32_t)(gx + 1024) << 21) | ((uint32_t)(gy + 1024) << 10) | (uint32_t)(gt + 512);
}
// (a & mask) | magic as ONE LOP3 (truth table 0xEA); written as asm because the compiler, given two literal
// constants, emits two LOP3s with an immediate each (2 of the 9 unpack instructions per pixel)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
    return d;
}
__device__ __forceinline__ void unpack_coef(uint32_t w, float& ix, float& iy, float& it) {
    ix = __fsub_rn(__uint_as_float((w >> 21) | 0x4B000000u), 8389632.0f);            // 2^23 + 1024
    iy = __fsub_rn(__uint_as_float(and_or(w >> 10, 0x7ffu, 0x4B000000u)), 8389632.0f);
    it = __fsub_rn(__uint_as_float(and_or(w, 0x3ffu, 0x4B000000u)), 8389120.0f);     // 2^23 + 512
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

// ------------------------------------------------------------------------------------------
// K1: spatio-temporal gradients + per-pixel coefficients, one pass over the two uint8 frames.
// Writes {Ix,Iy,It} packed into one word per pixel (pack_coef) and s = 1/sqrt(alpha^2+Ix^2+Iy^2)
// as float: 8 B of coefficients per pixel.  4 pixels per thread, vector stores.
// Frames: `frows` rows of `fpitch` bytes; buffer row y lives in frame row y + frow0 (frow0 = 1
// when a seam row sits above).  BORDER_REFLECT_101 therefore only ever triggers at true image
// borders.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_grad_coeff(const uint8_t* __restrict__ prev, const uint8_t* __restrict__ next,
             size_t fpitch, size_t fimg, int frows, int frow0,
             uint32_t* __restrict__ cpk, float* __restrict__ inv, Geom g, float alpha2) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x0 >= g.pitch || y >= g.H) return;
    const uint8_t* P = prev + (size_t)b * fimg;
    const uint8_t* N = next + (size_t)b * fimg;
    const int fy = y + frow0;
    const uint8_t* r0 = P + (size_t)reflect101(fy - 1, frows) * fpitch;
    const uint8_t* r1 = P + (size_t)fy * fpitch;
    const uint8_t* r2 = P + (size_t)reflect101(fy + 1, frows) * fpitch;
    const uint8_t* n1 = N + (size_t)fy * fpitch;

    // columns x0-1 .. x0+4 of the three rows (reflected at the image's left/right border)
    int a[6], c[6], d[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int xx = x0 - 1 + i;
        int xr = (xx < g.W + 1) ? reflect101(xx, g.W) : 0;  // columns past W+1 are never used
        a[i] = r0[xr];
        c[i] = r1[xr];
        d[i] = r2[xr];
    }
    uint32_t opk[4];
    float oinv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        int gx = 0, gy = 0, gt = 0;
        float iv = 0.f;
        if (x < g.W) {
            gx = (a[i + 2] + 2 * c[i + 2] + d[i + 2]) - (a[i] + 2 * c[i] + d[i]);          // :27
            gy = (d[i] + 2 * d[i + 1] + d[i + 2]) - (a[i] + 2 * a[i + 1] + a[i + 2]);      // :28
            gt = (int)n1[x] - c[i + 1];                                                    // :39
            const float den = __fadd_rn(alpha2, (float)(gx * gx + gy * gy));               // :65-68
            iv = __frsqrt_rn(den);                                                         // s = 1/sqrt(den)
        }
        opk[i] = pack_coef(gx, gy, gt);
        oinv[i] = iv;
    }
    const size_t o = (size_t)b * g.plane + (size_t)y * g.pitch + x0;
    *reinterpret_cast<uint4*>(cpk + o) = make_uint4(opk[0], opk[1], opk[2], opk[3]);
    *reinterpret_cast<float4*>(inv + o) = make_float4(oinv[0], oinv[1], oinv[2], oinv[3]);
}

// ------------------------------------------------------------------------------------------
// K0: the caller's preprocess() on the device (HornSchunckOF/main.cpp:11-26): 8UC3 BGR -> 8UC1 with
// OpenCV's 15-bit fixed-point luma, Y = (3735 B + 19235 G + 9798 R + 2^14) >> 15 (bit-exact with
// cv::cvtColor(COLOR_BGR2GRAY)).  4 pixels (12 bytes in, 4 out) per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bgr2gray(const uint8_t* __restrict__ bgr, size_t bgr_pitch, uint8_t* __restrict__ gray, size_t gray_pitch,
           int W, int H) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= W || y >= H) return;
    const uint8_t* src = bgr + (size_t)y * bgr_pitch + (size_t)x0 * 3;
    uint8_t* dst = gray + (size_t)y * gray_pitch + x0;
    uint32_t out = 0;
    if (x0 + 4 <= W) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);   // 12 bytes, 4-byte aligned (pitch % 4 == 0)
        const uint32_t w0 = s32[0], w1 = s32[1], w2 = s32[2];
        const uint32_t b[4] = {w0 & 255u, (w0 >> 24) & 255u, (w1 >> 16) & 255u, (w2 >> 8) & 255u};
        const uint32_t gch[4] = {(w0 >> 8) & 255u, w1 & 255u, (w1 >> 24) & 255u, (w2 >> 16) & 255u};
        const uint32_t r[4] = {(w0 >> 16) & 255u, (w1 >> 8) & 255u, w2 & 255u, (w2 >> 24) & 255u};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            out |= ((3735u * b[i] + 19235u * gch[i] + 9798u * r[i] + 16384u) >> 15) << (8 * i);
        *reinterpret_cast<uint32_t*>(dst) = out;                          // gray pitch is a multiple of 128
    } else {
        for (int i = 0; x0 + i < W; ++i)
            dst[i] = (uint8_t)((3735u * src[3 * i] + 19235u * src[3 * i + 1] + 9798u * src[3 * i + 2] + 16384u) >> 15);
    }
}

// ------------------------------------------------------------------------------------------
// the per-pixel update, shared by K2 and K3 (hornSchunck.cpp:63-73)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void hs_update_bar(float ub, float vb, float ix, float iy, float it, float s,
                                              float& un, float& vn) {
    const float P = __fmul_rn(ix, s), Q = __fmul_rn(iy, s), R = __fmul_rn(it, s);
    const float t = __fmaf_rn(P, ub, __fmaf_rn(Q, vb, R));
    un = __fmaf_rn(-P, t, ub);
    vn = __fmaf_rn(-Q, t, vb);
}
__device__ __forceinline__ void hs_update(float su, float sv, float kf, float ix, float iy,
                                          float it, float s, float& un, float& vn) {
    hs_update_bar(__fmul_rn(su, kf), __fmul_rn(sv, kf), ix, iy, it, s, un, vn);
}

// ---- "textbook" Horn-Schunck mode (HS_FLAG_TEXTBOOK; SURVEY 8f row 4, NOT a parity item: the
// reference does not compute this, BASELINE.json's prose does) --------------------------------------
//   gradients: Horn & Schunck's 2x2x2 cube, Ix = 1/4 [sum over y,t in {0,1} of I(x+1,.)-I(x,.)] etc.,
//              replicated at the right/bottom border; 4*Ix, 4*Iy, 4*It are exact integers <= 1020
//   average:   ubar = 1/6 (N,S,E,W) + 1/12 (diagonals) = 1/12 (s121 x s121)(u) - 1/3 u, zeros outside
//   canonical: h(x) = fma(2, t[x], t[x-1] + t[x+1]);  V(y) = fma(2, h[y], h[y-1] + h[y+1]);
//              ubar = fma(V, 1/12, u * (-1/3))
// Coefficients: `cpk` holds 4*Ix, 4*Iy (11-bit biased fields), the second plane holds It as float
// (It is a multiple of 1/4), and inv is recomputed per tile (33 bits do not fit one word).
#define HS_TB_W12 0.083333336f
#define HS_TB_W3 (-0.33333334f)
__device__ __forceinline__ void unpack_coef_tb(uint32_t w, float& ix, float& iy) {
    float i4, j4, unused;
    unpack_coef(w, i4, j4, unused);
    ix = __fmul_rn(i4, 0.25f);
    iy = __fmul_rn(j4, 0.25f);
}
__device__ __forceinline__ float hs_inv(float ix, float iy, float alpha2) {      // s = 1 / sqrt(den)
    return __frsqrt_rn(__fadd_rn(alpha2, __fmaf_rn(ix, ix, __fmul_rn(iy, iy))));
}
__device__ __forceinline__ float tb_bar(float V, float centre) {
    return __fmaf_rn(V, HS_TB_W12, __fmul_rn(centre, HS_TB_W3));
}

// K1 for the textbook mode
__global__ void __launch_bounds__(256)
k_grad_coeff_tb(const uint8_t* __restrict__ prev, const uint8_t* __restrict__ next,
                size_t fpitch, size_t fimg, int frows, int frow0,
                uint32_t* __restrict__ cpk, float* __restrict__ itp, Geom g) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x0 >= g.pitch || y >= g.H) return;
    const uint8_t* P = prev + (size_t)b * fimg;
    const uint8_t* N = next + (size_t)b * fimg;
    const int fy = y + frow0;
    const int fy1 = min(fy + 1, frows - 1);               // replicate at the true bottom border only
    const uint8_t* p0 = P + (size_t)fy * fpitch;
    const uint8_t* p1 = P + (size_t)fy1 * fpitch;
    const uint8_t* n0 = N + (size_t)fy * fpitch;
    const uint8_t* n1 = N + (size_t)fy1 * fpitch;
    uint32_t opk[4];
    float oit[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        int gx = 0, gy = 0, gt = 0;
        if (x < g.W) {
            const int x1 = min(x + 1, g.W - 1);
            const int a00 = p0[x], a01 = p0[x1], a10 = p1[x], a11 = p1[x1];
            const int c00 = n0[x], c01 = n0[x1], c10 = n1[x], c11 = n1[x1];
            gx = (a01 - a00) + (a11 - a10) + (c01 - c00) + (c11 - c10);      // 4 * Ix
            gy = (a10 - a00) + (a11 - a01) + (c10 - c00) + (c11 - c01);      // 4 * Iy
            gt = (c00 + c01 + c10 + c11) - (a00 + a01 + a10 + a11);          // 4 * It
        }
        opk[i] = pack_coef(gx, gy, 0);
        oit[i] = __fmul_rn((float)gt, 0.25f);
    }
    const size_t o = (size_t)b * g.plane + (size_t)y * g.pitch + x0;
    *reinterpret_cast<uint4*>(cpk + o) = make_uint4(opk[0], opk[1], opk[2], opk[3]);
    *reinterpret_cast<float4*>(itp + o) = make_float4(oit[0], oit[1], oit[2], oit[3]);
}

// K2 for the textbook mode: one weighted-average sweep, one pixel per thread
__global__ void __launch_bounds__(256)
k_jacobi_generic_tb(const float2* __restrict__ uv, float2* __restrict__ uvn,
                    const uint32_t* __restrict__ cpk, const float* __restrict__ itp, Geom g, float alpha2) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = g.oy0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= g.W || y >= g.oy1) return;
    const size_t base = (size_t)blockIdx.z * g.plane;
    const float2* P = uv + base;
    auto tap = [&](int yy, int xx) {
        return (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? P[(size_t)yy * g.pitch + xx] : make_float2(0.f, 0.f);
    };
    auto h = [&](int yy) {
        const float2 l = tap(yy, x - 1), c = tap(yy, x), r = tap(yy, x + 1);
        return make_float2(__fmaf_rn(2.f, c.x, __fadd_rn(l.x, r.x)), __fmaf_rn(2.f, c.y, __fadd_rn(l.y, r.y)));
    };
    const float2 hm = h(y - 1), h0 = h(y), hp = h(y + 1), ctr = tap(y, x);
    const float Vu = __fmaf_rn(2.f, h0.x, __fadd_rn(hm.x, hp.x));
    const float Vv = __fmaf_rn(2.f, h0.y, __fadd_rn(hm.y, hp.y));
    const size_t o = base + (size_t)y * g.pitch + x;
    float ix, iy, nu, nv;
    unpack_coef_tb(cpk[o], ix, iy);
    hs_update_bar(tb_bar(Vu, ctr.x), tb_bar(Vv, ctr.y), ix, iy, itp[o], hs_inv(ix, iy, alpha2), nu, nv);
    uvn[o] = make_float2(nu, nv);
}

// ------------------------------------------------------------------------------------------
// K2: one Jacobi sweep, any window size (runtime w, anchor a).  One pixel per thread, taps read
// through L1/L2.  This is the always-correct path (even / large windows, A/B reference for K3).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 add2s(float2 a, float2 b) {      // two scalar IEEE adds (same rounding as FADD2)
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__global__ void __launch_bounds__(256)
k_jacobi_generic(const float2* __restrict__ uv, float2* __restrict__ uvn,
                 const uint32_t* __restrict__ cpk, const float* __restrict__ inv,
                 Geom g, int w, int a, float kf) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = g.oy0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= g.W || y >= g.oy1) return;
    const size_t base = (size_t)blockIdx.z * g.plane;
    const float2* P = uv + base;
    // canonical paired sums, written as plain loops (see the header comment); .x = u, .y = v
    auto row_sum = [&](int yy) {                            // sum over columns x-a .. x-a+w-1 of row yy
        const bool yin = (yy >= 0) && (yy < g.H);
        const float2* row = P + (size_t)(yin ? yy : 0) * g.pitch;
        auto tap = [&](int xx) { return (yin && xx >= 0 && xx < g.W) ? row[xx] : make_float2(0.f, 0.f); };
        int col = x - a;
        const int hi = col + w - 1;
        float2 s = make_float2(0.f, 0.f);
        bool first = true;
        auto add = [&](float2 t) { s = first ? t : add2s(s, t); first = false; };
        if (col & 1) { add(tap(col)); ++col; }
        for (; col + 1 <= hi; col += 2) add(add2s(tap(col), tap(col + 1)));
        if (col == hi) add(tap(col));
        return s;
    };
    // the same rule down the rows y-a .. y-a+w-1
    int r = y - a;
    const int hi = r + w - 1;
    float2 s = make_float2(0.f, 0.f);
    bool first = true;
    auto add = [&](float2 t) { s = first ? t : add2s(s, t); first = false; };
    if (w != 3) {                                        // (w = 3 has its own rule below)
        if ((r + g.grow0) & 1) { add(row_sum(r)); ++r; }
        for (; r + 1 <= hi; r += 2) add(add2s(row_sum(r), row_sum(r + 1)));
        if (r == hi) add(row_sum(r));
    }
    const size_t o = base + (size_t)y * g.pitch + x;
    float ix, iy, it, nu, nv;
    unpack_coef(cpk[o], ix, iy, it);
    if (w == 3) {
        // canonical rule for w = 3: the column window y-1 .. y+1 is one aligned pair of rows and one single row;
        // mean = fma(single, kf, kf * pair) (the fused kernel shares kf * pair between two rows: one packed
        // multiply per two pixels instead of one per pixel)
        const bool odd = ((y - 1 + g.grow0) & 1) != 0;                       // the window starts on an odd image row
        const float2 single = odd ? row_sum(y - 1) : row_sum(y + 1);
        const float2 pair = odd ? add2s(row_sum(y), row_sum(y + 1)) : add2s(row_sum(y - 1), row_sum(y));
        const float ub = __fmaf_rn(single.x, kf, __fmul_rn(pair.x, kf));
        const float vb = __fmaf_rn(single.y, kf, __fmul_rn(pair.y, kf));
        hs_update_bar(ub, vb, ix, iy, it, inv[o], nu, nv);
    } else {
        hs_update(s.x, s.y, kf, ix, iy, it, inv[o], nu, nv);
    }
    uvn[o] = make_float2(nu, nv);
}

// ------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (inline PTX; SASS: UTMALDG / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// barrier among the compute warps only (the producer warp never joins it)
template <int THREADS>
__device__ __forceinline__ void compute_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Every device-side wait is bounded: a protocol bug must end in a trapped kernel, never in a GPU
// that spins forever (a hung kernel cannot be killed from the host).
constexpr unsigned long long WAIT_LIMIT_NS = 20000000000ull;   // 20 s
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && (++spins & 0x3ff) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_LIMIT_NS) __trap();
        }
    } while (!done);
}
// 3-D tiled load: coordinates (x, y, pair), innermost first; out-of-bounds elements read as 0,
// which is exactly BORDER_CONSTANT (hornSchunck.cpp:60-61) for the first fused sweep.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
// Programmatic dependent launch: consecutive sweep launches are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a CTA of launch n+1 may become resident
// as soon as an SM drains launch n; it must not touch global memory before pdl_wait() returns
// (= launch n complete and flushed).
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// 4-D tiled load (32-byte chunk, chunk, y, pair): the swizzled flow-plane map (make_map_uv in hs_api.cu)
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------------------
// K3: ALL sweeps of a solve in one launch: phases of k fused sweeps (temporal blocking) over
// staged tiles, persistent warp-specialised CTAs, dataflow synchronisation between tiles.
//
// One CTA per SM (12 compute warps + a producer warp group) walks over staged tiles of SX x SY
// pixels (SX = 128 = 32 lanes x 4 px, SY = NWARP x R rows).  Per tile:
//   * four TMA boxes (u, v, packed Ix/Iy/It, inv) land in one of TWO shared-memory stages; the
//     producer issues the boxes of the NEXT tile as soon as the compute warps have pulled this
//     tile into registers, so they stream in underneath the k sweeps of this tile
//     (mbarriers full[stage] / empty[stage])
//   * each thread keeps its 4 x R patch of u, v AND its coefficients in registers for all k sweeps
//   * per sweep: (1) row sums of the patch, horizontal neighbours by warp shuffle; (2) the rows a
//     vertical neighbour needs go to a double-buffered exchange array (it lives in the stage the
//     tile came from, which is dead once the registers are loaded); (3) the rows whose window
//     stays inside the patch are updated; (4) one named barrier of the compute warps; (5) the
//     neighbour rows come back, the remaining rows are updated in place
// The ring of pixels whose dependency cone leaves the staged tile grows by (a, w/2) per sweep,
// so after k sweeps the centre VX x VY pixels are exact and are the only ones stored.
// Pixels outside the image must be 0 at EVERY sweep (BORDER_CONSTANT): TMA zero-fill gives that
// for sweep 1, border tiles re-zero them after each sweep (`masked` path, CTA-uniform branch).
// ------------------------------------------------------------------------------------------
template <int RL, int RR, int R, int NWARP>
struct TileShape {
    static_assert(RL <= R && RR <= R, "a thread's vertical neighbours are the adjacent patches only");
    static constexpr int SX = 128;
    static constexpr int SY = NWARP * R;
    static constexpr int CTHREADS = NWARP * 32;      // compute threads
    static constexpr int THREADS = CTHREADS + 128;   // + one producer warp group (setmaxnreg works per 4 warps;
                                                     //   only its first warp does anything)
    static constexpr int NSLOT = (RL + RR < R) ? (RL + RR) : R;       // patch rows a neighbour reads
    static constexpr size_t BYTES_F32 = (size_t)SX * SY * 4;
    static constexpr size_t OFF_UV = 0;                               // offsets inside one stage; uv = {u, v} interleaved
    static constexpr size_t OFF_CPK = OFF_UV + 2 * BYTES_F32;
    static constexpr size_t OFF_INV = OFF_CPK + BYTES_F32;
    static constexpr size_t STAGE = OFF_INV + BYTES_F32;
    // exchange array [2 buf][2 field][NWARP][NSLOT][SX] floats, aliased onto the consumed stage
    static constexpr size_t EX_FIELD = (size_t)NWARP * NSLOT * SX;    // floats
    static constexpr size_t BYTES_EX = 2 * 2 * EX_FIELD * 4;
    static_assert(BYTES_EX <= STAGE, "exchange array must fit in one stage");
    static constexpr size_t OFF_BAR = 2 * STAGE;
    static constexpr size_t OFF_INFO = OFF_BAR + 64;                  // int4 x 2 x 2: decoded (x0, y0, pair, phase | slab, ...) of the staged items
    static constexpr size_t OFF_HOT = OFF_INFO + 64;                  // per-slab fields the compute warps read per tile
    static constexpr size_t HOT_BYTES = 128;
    static constexpr size_t SMEM_FOR(int maxs) { return OFF_HOT + HOT_BYTES * maxs; }
    static constexpr size_t SMEM = SMEM_FOR(4);                       // full[2], empty[2], stored[2]; item info; hot
    static constexpr uint32_t TX_BYTES = (uint32_t)STAGE;
    // slot of patch row j in the exchange array (-1: no neighbour reads it)
    __host__ __device__ static constexpr int slot(int j) {
        return (RL + RR >= R) ? j : (j < RR ? j : (j >= R - RL ? RR + (j - (R - RL)) : -1));
    }
};

// Does any of the 4 outputs of a 4-aligned group use the single element / the pair starting at
// local position `pos` (may be negative or >= 4)?  Walks the pairing rule at compile time.
template <int RL, int RR>
__host__ __device__ constexpr bool uses_term(bool pair, int pos) {
    for (int c = 0; c < 4; ++c) {
        int x = c - RL;
        const int hi = c + RR;
        if (x & 1) { if (!pair && x == pos) return true; ++x; }
        for (; x + 1 <= hi; x += 2) if (pair && x == pos) return true;
        if (x == hi && !pair && x == pos) return true;
    }
    return false;
}

// ---- packed fp32 (Blackwell FADD2 / FMUL2 / FFMA2) ---------------------------------------------
// u and v go through IDENTICAL window sums, so the fused kernel carries them as one float2 per pixel
// (.x = u, .y = v) and adds / scales both with one instruction.  Each half is rounded exactly like
// the scalar __f*_rn form (IEEE round-to-nearest per lane), so the canonical arithmetic - and with
// it bit-identity with k_jacobi_generic - is unchanged; only the instruction count drops
// (13 -> 8 FP instructions per pixel-sweep for w=3).
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 shfl_up2(float2 a) {
    return make_float2(__shfl_up_sync(0xffffffffu, a.x, 1), __shfl_up_sync(0xffffffffu, a.y, 1));
}
__device__ __forceinline__ float2 shfl_down2(float2 a) {
    return make_float2(__shfl_down_sync(0xffffffffu, a.x, 1), __shfl_down_sync(0xffffffffu, a.y, 1));
}
// the per-pixel update on a packed average (hornSchunck.cpp:63-73), same operations as hs_update_bar with the
// normalised coefficients pq = {P, Q}, r = R already formed:
//   t = fma(P, ubar, fma(Q, vbar, R));  {u', v'} = fma(-{P, Q}, {t, t}, {ubar, vbar})
__device__ __forceinline__ float2 hs_update_bar2(float2 bar, float2 pq, float r) {
    const float t = __fmaf_rn(pq.x, bar.x, __fmaf_rn(pq.y, bar.y, r));
    return __ffma2_rn(make_float2(-pq.x, -pq.y), make_float2(t, t), bar);   // FFMA2 -R.F32x2, R.F32, R.F32x2
}

// Paired window sums of 4 consecutive positions 0..3 (position 0 is even in absolute terms).
// s[i] = element at position i-4 (i = 0..11), p[i] = pair sum starting at position 2i-4 (i = 0..5);
// only the entries uses_term() asks for have to be filled in.
template <int RL, int RR>
__device__ __forceinline__ void paired_sums(const float2 (&s)[12], const float2 (&p)[6], float2 (&out)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int x = c - RL;
        const int hi = c + RR;
        float2 acc = make_float2(0.f, 0.f);
        bool first = true;
        if (x & 1) { acc = s[x + 4]; first = false; ++x; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {                 // at most 4 pairs in a window of <= 9
            if (x + 1 <= hi) {
                acc = first ? p[(x + 4) / 2] : add2(acc, p[(x + 4) / 2]);
                first = false;
                x += 2;
            }
        }
        if (x == hi) acc = first ? s[x + 4] : add2(acc, s[x + 4]);
        out[c] = acc;
    }
}

// row sums of one patch row: the lane's 4 columns; neighbours' elements / pair sums by shuffle
template <int RL, int RR>
__device__ __forceinline__ void row_sums(const float2 (&a)[4], float2 (&h)[4]) {
    float2 s[12], p[6];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[4 + i] = a[i];
    constexpr bool need_p0 = uses_term<RL, RR>(true, 0) || uses_term<RL, RR>(true, -4) || uses_term<RL, RR>(true, 4);
    constexpr bool need_p2 = uses_term<RL, RR>(true, 2) || uses_term<RL, RR>(true, -2) || uses_term<RL, RR>(true, 6);
    if (need_p0) p[2] = add2(a[0], a[1]);
    if (need_p2) p[3] = add2(a[2], a[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (uses_term<RL, RR>(false, i - 4)) s[i] = shfl_up2(a[i]);         // left lane's element i
        if (uses_term<RL, RR>(false, i + 4)) s[8 + i] = shfl_down2(a[i]);   // right lane's element i
    }
    if (uses_term<RL, RR>(true, -4)) p[0] = shfl_up2(p[2]);
    if (uses_term<RL, RR>(true, -2)) p[1] = shfl_up2(p[3]);
    if (uses_term<RL, RR>(true, 4)) p[4] = shfl_down2(p[2]);
    if (uses_term<RL, RR>(true, 6)) p[5] = shfl_down2(p[3]);
    paired_sums<RL, RR>(s, p, h);
}

// row sums of the textbook mode: h = x[-1] + 2 x[0] + x[1]
__device__ __forceinline__ void row_sums_tb(const float2 (&a)[4], float2 (&h)[4]) {
    const float2 l = shfl_up2(a[3]);
    const float2 r = shfl_down2(a[0]);
    const float2 two = make_float2(2.f, 2.f);
    h[0] = __ffma2_rn(two, a[0], add2(l, a[1]));
    h[1] = __ffma2_rn(two, a[1], add2(a[0], a[2]));
    h[2] = __ffma2_rn(two, a[2], add2(a[1], a[3]));
    h[3] = __ffma2_rn(two, a[3], add2(a[2], r));
}

// One 256-bit store (Blackwell STG.E.ENL2.256): a lane's four {u, v} pixels = 32 contiguous bytes, a
// warp's row = 1 KB with every 32-byte sector written whole by ONE instruction.  (Two 128-bit stores at a
// 32-byte lane stride half-fill every sector twice: the store phase then runs at half the L1->L2 rate.)
__device__ __forceinline__ void st_global_256(float2* dst, const float2 (&a)[4]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(a[0].x), "f"(a[0].y),
                 "f"(a[1].x), "f"(a[1].y), "f"(a[2].x), "f"(a[2].y), "f"(a[3].x), "f"(a[3].y)
                 : "memory");
}

// k sweeps on a thread's 4 x R patch.  uv = {u, v} per pixel, gxy = {P, Q}, it = R (normalised coefficients).
template <int RL, int RR, int R, int NWARP, bool MASKED, bool TB>
__device__ __forceinline__ void tile_sweeps(float2 (&uv)[R][4], const float2 (&gxy)[R][4],
                                            const float (&it)[R][4],
                                            float* ex, int k, float kf, int warp, int lane,
                                            uint32_t inmask) {
    using TS = TileShape<RL, RR, R, NWARP>;
    // the patches above / below; at the tile's top and bottom there is none: those rows only feed
    // pixels of the invalid ring, so any finite-or-not value will do - read our own slots
    const int wa = warp > 0 ? warp - 1 : 0;
    const int wb = warp < NWARP - 1 ? warp + 1 : NWARP - 1;
    const float2 kf2 = make_float2(kf, kf);
    // one sweep; PAR = parity of the sweep = which of the two exchange arrays it uses.  The loop below is unrolled
    // by two so that PAR is a compile-time constant and every exchange address is base + immediate (the address
    // arithmetic per sweep was ~10 integer instructions, half of them IMADs on the FP32 pipe)
    auto sweep = [&](auto PAR) {
        // exchange array of this sweep parity: [NWARP][NSLOT][SX] float2 = {row sum of u, row sum of v}
        float2* exs = reinterpret_cast<float2*>(ex) + (size_t)decltype(PAR)::value * TS::EX_FIELD;
        float2 h[R][4];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (TB) row_sums_tb(uv[j], h[j]);
            else row_sums<RL, RR>(uv[j], h[j]);
            if (TS::slot(j) >= 0) {  // rows a vertical neighbour will need
                // a row is stored as four quarters of 32 lanes x 8 B (pixel c of every lane): each {u, v} sum is a
                // register pair already, so these are plain STS.64 / LDS.64 (two STS.128 per row needed seven
                // MOVs to line their operands up; same number of shared-memory wavefronts either way)
                float2* o = exs + ((size_t)warp * TS::NSLOT + TS::slot(j)) * TS::SX + lane;
#pragma unroll
                for (int c = 0; c < 4; ++c) o[32 * c] = h[j][c];
            }
        }
        static_assert(R == 4, "the shared column pairs below are (0,1) and (2,3)");
        float2 q[2][4];                                // shared pair sums of the patch's own rows
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            q[0][c] = add2(h[0][c], h[1][c]);
            q[1][c] = add2(h[2][c], h[3][c]);
        }
        [[maybe_unused]] float2 qk[2][4];              // w = 3 only: the pair sums times 1/9 (dead code otherwise)
        if constexpr (RL == 1 && RR == 1 && !TB) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { qk[0][c] = __fmul2_rn(q[0][c], kf2); qk[1][c] = __fmul2_rn(q[1][c], kf2); }
        }
        // one patch row: paired box sum down the column from the row sums of rows j-RL..j+RR (the
        // patch's first row is an even image row), then the update
        auto column_term = [&](const float2 (&ab)[RL > 0 ? RL : 1][4], const float2 (&be)[RR > 0 ? RR : 1][4], int i, int c) {
            return i < 0 ? ab[i + RL][c] : (i >= R ? be[i - R][c] : h[i][c]);
        };
        auto update_row = [&](int j, const float2 (&ab)[RL > 0 ? RL : 1][4], const float2 (&be)[RR > 0 ? RR : 1][4]) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float2 bar;
                if constexpr (TB) {   // weighted average: V = h[j-1] + 2 h[j] + h[j+1], ubar = V/12 - u/3
                    const float2 V = __ffma2_rn(make_float2(2.f, 2.f), h[j][c], add2(column_term(ab, be, j - 1, c), column_term(ab, be, j + 1, c)));
                    bar = __ffma2_rn(V, make_float2(HS_TB_W12, HS_TB_W12), __fmul2_rn(uv[j][c], make_float2(HS_TB_W3, HS_TB_W3)));
                } else {
                    float2 sum = make_float2(0.f, 0.f);
                    bool first = true;
                    int i = j - RL;
                    const int hi = j + RR;
                    if (i & 1) { sum = column_term(ab, be, i, c); first = false; ++i; }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (i + 1 <= hi) {
                            // pair (i, i+1), i even: inside the patch it is one of the two shared sums
                            float2 pr;
                            if (i == 0) pr = q[0][c];
                            else if (i == 2) pr = q[1][c];
                            else pr = add2(column_term(ab, be, i, c), column_term(ab, be, i + 1, c));
                            sum = first ? pr : add2(sum, pr);
                            first = false;
                            i += 2;
                        }
                    }
                    if (i == hi) {
                        const float2 tt = column_term(ab, be, i, c);
                        sum = first ? tt : add2(sum, tt);
                    }
                    if constexpr (RL == 1 && RR == 1) {
                        // w = 3 (canonical rule, see the header comment): the column window is one pair and one single
                        // row; mean = fma(single, kf, kf * pair), the scaled pair sums being shared by two rows
                        const float2 single = (j == 0) ? ab[0][c] : (j == 1 ? h[2][c] : (j == 2 ? h[1][c] : be[0][c]));
                        bar = __ffma2_rn(single, kf2, qk[j >> 1][c]);
                    } else {
                        bar = __fmul2_rn(sum, kf2);
                    }
                }
                float2 n = hs_update_bar2(bar, gxy[j][c], it[j][c]);
                if (MASKED) {
                    const bool in = (inmask >> (j * 4 + c)) & 1u;
                    n.x = in ? n.x : 0.f;
                    n.y = in ? n.y : 0.f;
                }
                uv[j][c] = n;
            }
        };
        float2 ab[RL > 0 ? RL : 1][4], be[RR > 0 ? RR : 1][4];
        // rows whose window stays inside the patch need nothing from the neighbours: do them while
        // the exchange rows of the other warps are still on their way
#pragma unroll
        for (int j = RL; j < R - RR; ++j) update_row(j, ab, be);
        compute_sync<TS::CTHREADS>();
        // neighbour rows: the last RL rows of the patch above, the first RR rows of the patch below
#pragma unroll
        for (int i = 0; i < RL; ++i) {
            const float2* o = exs + ((size_t)wa * TS::NSLOT + TS::slot(R - RL + i)) * TS::SX + lane;
#pragma unroll
            for (int c = 0; c < 4; ++c) ab[i][c] = o[32 * c];
        }
#pragma unroll
        for (int i = 0; i < RR; ++i) {
            const float2* o = exs + ((size_t)wb * TS::NSLOT + TS::slot(i)) * TS::SX + lane;
#pragma unroll
            for (int c = 0; c < 4; ++c) be[i][c] = o[32 * c];
        }
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (j < RL || j >= R - RR) update_row(j, ab, be);
    };
    int s = 0;
    for (; s + 2 <= k; s += 2) {
        sweep(std::integral_constant<int, 0>{});
        sweep(std::integral_constant<int, 1>{});
    }
    if (s < k) sweep(std::integral_constant<int, 0>{});
}

#ifdef HS_TILE_PROFILE
// Debug build only (tools/tile_profile.py): per-CTA cycle counts, summed over the tiles of a launch:
// [0] wait for TMA, [1] smem->registers + unpack + barrier, [2] sweeps, [3] stores, [4] tiles,
// [5] TMA issue -> landed (sum), [6] (max), [7] consumer barrier -> TMA issue of the next tile.
__device__ long long g_tile_prof[256][8];
#define HS_PROF_T(var) const long long var = clock64()
#define HS_PROF_ADD(slot, a, b) if (tid == 0) g_tile_prof[blockIdx.x][slot] += (b) - (a)
#else
#define HS_PROF_T(var)
#define HS_PROF_ADD(slot, a, b)
#endif

// One row slab (or a whole image) of a launch: its planes, TMA maps, tile geometry and - for the
// row-slab decomposition over several GPUs - its two seams.  Host-computed, passed by value.
//
// Seams.  A slab's buffer holds its own rows [oy0, oy1) plus halo rows above / below that belong to
// the neighbouring slabs (other GPUs).  Nothing is exchanged between launches: a tile that finishes
// a phase stores the rows the neighbour's next phase will read STRAIGHT INTO THE NEIGHBOUR'S HALO
// ROWS (peer memory over NVLink, plain vector stores), and its producer warp then publishes the
// tile's phase count in the neighbour's `inbox` (fence.acq_rel.sys + st.release.sys).  The
// neighbour's producer warp polls its own inbox (local memory) next to the local `done[]` counters
// before it issues the TMA loads of a seam tile.  The rule is the one used inside a GPU - a tile
// may start phase p once every tile whose rows it reads, or that reads the rows it will overwrite,
// has finished phase p-1 - so the read-after-write on the halo rows and the write-after-read on the
// ping-pong plane are both covered.  Flags are ABSOLUTE phase counts since hs_prepare (`phase_base`
// + phase of the launch + 1), so they never have to be reset while neighbours are running.
constexpr int SEAM_JMAX = 2;   // tile rows of a slab that can touch one seam (halos fit in one tile pitch)

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (host computes mul, shr).  The producer
// warp decodes one work item per tile on a latency chain of its own; real divisions there cost more
// than the TMA issue they precede.
struct FastDiv {
    uint32_t mul, shr, d;
    __host__ __device__ static FastDiv make(uint32_t den) {
        FastDiv f; f.d = den; f.mul = 0; f.shr = 0;
        if (den > 1) {
            uint32_t lg = 0;
            while ((1ull << lg) < den) ++lg;
            const uint32_t pw = 31 + lg;
            f.mul = (uint32_t)(((1ull << pw) + den - 1) / den);
            f.shr = pw - 32;
        }
        return f;
    }
    __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr); }
};
struct SlabDesc {
    CUtensorMap tm_uv[2], tm_cpk, tm_inv;   // tm_uv: the flow plane in 32-byte chunks of 4 pixels ({u, v} interleaved), SWIZZLE_32B
    float2* uv[2];
    Geom g;                    // oy0 / oy1 = the rows this launch produces
    int hyt, vy;               // halo rows above the stored centre of a tile; rows of the centre
    int tiles_y, ntiles;       // tile rows; tiles_x * tiles_y * batch
    int tile0;                 // index of the slab's first tile in done[] and in the item stream
    FastDiv fd_per_img;        // / (tiles_x * tiles_y)
    int reverse;               // walk the tile rows bottom-up (odd slabs: both sides of a seam are then
                               //   processed at the same end of a phase, a whole phase before they are needed)
    // --- seams; null pointers = no neighbour on that side -----------------------------------
    float2* up_uv[2];          // the planes of the slab above / below (peer memory)
    float2* dn_uv[2];
    int up_dy, dn_dy;          // buffer row y of this slab is buffer row y + dy of that neighbour
    int push_up, push_dn;      // the first push_up / last push_dn produced rows are halo rows of the neighbour
    int* inbox;                // [2][SEAM_JMAX][tiles_x]: phase counts published by the slab above ([0]) / below ([1])
    int* out_up; int* out_dn;  // where this slab publishes: the neighbours' inbox blocks for this side
    int jt, jb;                // tile rows [0, jt) touch the seam above, [tiles_y - jb, tiles_y) the seam below
    int up_j, dn_j;            // flag rows to wait for in inbox[0] / inbox[1] (the neighbour's jb / jt); 0 = no seam
};

// What the compute warps need from a SlabDesc for every tile, copied to shared memory once per launch:
// reading it from the kernel-parameter constant bank missed the constant cache about once per tile
// (ncu: ~5 % of all stall samples on three LDCs of the store epilogue).
struct SlabHot {
    float2* uv[2];
    float2* up_uv[2];
    float2* dn_uv[2];
    long long plane;
    int W, H, pitch, oy0, oy1, hyt, vy;
    int push_up, push_dn, up_dy, dn_dy;
    int seam;                  // any neighbour to push rows to
};
static_assert(sizeof(SlabHot) <= 128, "SlabHot must fit HOT_BYTES");

// geometry of one launch (host-computed)
template <int MAXS>
struct LaunchDesc {
    SlabDesc s[MAXS];
    int nslabs;
    int k;                 // sweeps fused per phase (tile geometry is sized for this)
    int sweeps;            // total sweeps of this launch = phases * k (the last phase may be short)
    int cur;               // phase p reads planes [cur ^ (p & 1)] and writes the other pair
    int phase_base;        // phases completed since hs_prepare (seam flags are absolute)
    int hxl, vx;           // halo columns left of / columns of the stored centre of a tile
    int tiles_x;           // tile columns (all slabs have the same width)
    FastDiv fd_tiles_x;
    int ntiles;            // all tiles of all slabs
    int* done;             // per-tile phase counters of this launch (zeroed by the host)
    float kf, alpha2;
};

// gpu-scope flag helpers for the inter-CTA dataflow of a multi-phase launch
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// system scope: flags and halo rows that cross NVLink (row-slab seams)
__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// generic <-> async proxy ordering for GLOBAL memory only: FENCE.VIEW.ASYNC.G, no MEMBAR (the
// unqualified fence.proxy.async also emits a MEMBAR.ALL.GPU, ~1.5k cycles on the producer's chain)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// acquire loads: LDG.STRONG + CCTL.IVALL, again no MEMBAR (a relaxed load + fence.acq_rel would wait
// for every outstanding store of the SM, which an acquire has no use for)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int* p, int v) {
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One launch = `phases` phases of up to k sweeps over every tile.  Phase p reads the flow planes
// (p & 1) and writes the planes ((p & 1) ^ 1).  There is NO grid-wide barrier between phases:
// tile t may start phase p as soon as t and its 8 neighbours have finished phase p-1 (their
// per-tile counters `done[]` say so), which covers both the read-after-write on the halo and the
// write-after-read on the plane that is overwritten.  Work is one phase-major stream of
// (phase, tile) items dealt round robin to the CTAs, so a tile's inputs were finished about one
// whole phase (ntiles / #CTAs rounds) before its turn and the TMA prefetch of the next tile
// keeps overlapping the current one.
// All CTAs must be co-resident (grid <= #SMs, cooperative launch) because they wait on one another.
//
// Warp roles: warps 0..NWARP-1 compute; warp NWARP is the producer.  The producer owns everything
// with long latency that is not arithmetic: it polls the neighbours' counters, fences, issues the
// TMA boxes of the next item into the free stage (empty[] -> full[] mbarriers) and publishes this
// CTA's finished tiles (stored[] mbarrier -> fence -> counter), polling all of that without ever
// blocking on one duty, so a CTA can never hold back a tile somebody else is waiting for.
//
// MAXS = 1 is the production instantiation (one slab or whole image per launch and GPU).  MAXS > 1
// runs several row slabs of ONE device in one cooperative launch: the seam protocol (peer stores,
// system-scope flags) is then exercised end to end on a single GPU, without kernels of different
// launches waiting on each other.
template <int RL, int RR, int R, int NWARP, bool TB = false, int MAXS = 1>
__global__ void __launch_bounds__(NWARP * 32 + 128, 1)
k_jacobi_tile(const __grid_constant__ LaunchDesc<MAXS> d) {
    using TS = TileShape<RL, RR, R, NWARP>;
    static_assert(!TB || (RL == 1 && RR == 1), "the textbook average is a 3x3 stencil");
    static_assert(R * 4 <= 32, "in-image mask is one 32-bit word per thread");
    static_assert(RL <= 4 && RR <= 4, "horizontal neighbours come from the adjacent lane only");
    static_assert(NWARP % 4 == 0, "setmaxnreg acts on whole warp groups");
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TS::OFF_BAR);  // TMA bytes landed      [2]
    uint64_t* empty = full + 2;                                        // stage may be refilled [2]
    uint64_t* stored = full + 4;                                       // tile results stored   [2]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int G = gridDim.x;
    const int phases = (d.sweeps + d.k - 1) / d.k;
    const int total = phases * d.ntiles;              // work items of the launch, phase-major: g = p*ntiles + t
    const int items = (total - (int)blockIdx.x + G - 1) / G;   // ... of this CTA (>= 1: grid <= ntiles)
    int* __restrict__ done = d.done;

    // CTA c takes items c, c+G, c+2G, ... of the global phase-major stream.  Unless G divides the
    // tile count this rotates the tile -> CTA assignment from phase to phase, so the slower border
    // tiles (masked path) are shared by everybody instead of pinning a few CTAs that all their
    // neighbours then have to wait for.
    // t: index in done[] (all slabs); by: the REAL tile row; pub_up / pub_dn: where a seam tile publishes
    // in the neighbour's inbox (-1: not a seam tile)
    struct Item { int p, t, s, b, bx, by, x0, y0, pub_up, pub_dn; };
    int seq_p = 0, seq_t = (int)blockIdx.x;           // the stream is walked in order: no division by ntiles
    auto next_item = [&]() {
        Item it;
        it.p = seq_p;
        const int t = seq_t;
        seq_t += G;
        if (seq_t >= d.ntiles) { seq_t -= d.ntiles; ++seq_p; }   // G <= ntiles: at most one wrap
        it.s = 0;
        if (MAXS > 1) {
#pragma unroll
            for (int q = 1; q < MAXS; ++q)
                if (q < d.nslabs && t >= d.s[q].tile0) it.s = q;
        }
        const SlabDesc& S = d.s[it.s];
        const int per_img = d.tiles_x * S.tiles_y;
        const int lt = t - S.tile0;
        it.b = S.fd_per_img.div(lt);
        const int r = lt - it.b * per_img;
        const int row = d.fd_tiles_x.div(r);          // position in this slab's walking order
        it.bx = r - row * d.tiles_x;
        it.by = S.reverse ? S.tiles_y - 1 - row : row;
        it.t = S.tile0 + it.b * per_img + it.by * d.tiles_x + it.bx;
        it.x0 = it.bx * d.vx - d.hxl;                 // staged tile origin (may be negative)
        it.y0 = S.g.oy0 + it.by * S.vy - S.hyt;
        it.pub_up = (S.out_up && it.by < S.jt) ? it.by * d.tiles_x + it.bx : -1;
        it.pub_dn = (S.out_dn && it.by >= S.tiles_y - S.jb) ? (it.by - (S.tiles_y - S.jb)) * d.tiles_x + it.bx : -1;
        return it;
    };

    if (tid < MAXS && tid < d.nslabs) {               // descriptor fields of the hot loop: constant bank -> shared memory
        const SlabDesc& S = d.s[tid];
        SlabHot h;
        for (int i = 0; i < 2; ++i) { h.uv[i] = S.uv[i]; h.up_uv[i] = S.up_uv[i]; h.dn_uv[i] = S.dn_uv[i]; }
        h.plane = S.g.plane; h.W = S.g.W; h.H = S.g.H; h.pitch = S.g.pitch; h.oy0 = S.g.oy0; h.oy1 = S.g.oy1;
        h.hyt = S.hyt; h.vy = S.vy; h.push_up = S.push_up; h.push_dn = S.push_dn; h.up_dy = S.up_dy; h.dn_dy = S.dn_dy;
        h.seam = (S.up_uv[0] != nullptr) || (S.dn_uv[0] != nullptr);
        *reinterpret_cast<SlabHot*>(smem + TS::OFF_HOT + TS::HOT_BYTES * tid) = h;
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(&stored[0], NWARP);
        mbar_init(&stored[1], NWARP);
        fence_mbar_init();
    }
    __syncthreads();

    // 512 threads are launched with 128 registers each (the whole register file).  The producer warp
    // group hands most of its share back and the three compute warp groups take it.
    if (warp >= NWARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp > NWARP) return;
        // ================================ producer warp ================================
        // Start coordinates are multiples of 4 floats = 16 B: UTMALDG faults ("illegal instruction")
        // on sm_100a when the innermost start offset is not 16-byte aligned (tools/tma_probe.cu).
        auto issue = [&](const Item& it, int stage) {   // lane 0 only
            unsigned char* st = smem + (size_t)stage * TS::STAGE;
            const SlabDesc& S = d.s[it.s];
            const int rd = (d.cur ^ it.p) & 1;            // planes this phase reads
            // the decoded item travels with its boxes: the compute warps read it after the full[]
            // wait instead of each redoing the divisions (the arrive below releases this store)
            int4* info = reinterpret_cast<int4*>(smem + TS::OFF_INFO) + 2 * stage;
            info[0] = make_int4(it.x0, it.y0, it.b, it.p);
            if (MAXS > 1) info[1] = make_int4(it.s, 0, 0, 0);
            mbar_expect_tx(&full[stage], TS::TX_BYTES);
            tma_load_3d(st + TS::OFF_CPK, &S.tm_cpk, &full[stage], it.x0, it.y0, it.b);
            tma_load_3d(st + TS::OFF_INV, &S.tm_inv, &full[stage], it.x0, it.y0, it.b);
            tma_load_4d(st + TS::OFF_UV, &S.tm_uv[rd], &full[stage], 0, it.x0 >> 2, it.y0, it.b);   // x0 is a multiple of 4 (may be negative)
        };
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < MAXS; ++q)
                if (q < d.nslabs) {
                    tma_prefetch_desc(&d.s[q].tm_uv[0]);
                    tma_prefetch_desc(&d.s[q].tm_uv[1]);
                    tma_prefetch_desc(&d.s[q].tm_cpk);
                    tma_prefetch_desc(&d.s[q].tm_inv);
                }
        }
        pdl_launch_dependents();      // the next launch may queue up behind us (it waits in pdl_wait)
        pdl_wait();                   // whatever wrote the planes we read (K1, a sweep launch, a halo copy) is done
        // A seam tile depends on the neighbouring slab (another GPU) even in the first phase of a launch:
        // that slab's previous launch must have delivered its last phase.  Which tiles are seam tiles:
        auto seam_sides = [&](const Item& it, bool& up, bool& dn) {
            const SlabDesc& S = d.s[it.s];
            up = S.up_j > 0 && it.by < S.jt;
            dn = S.dn_j > 0 && it.by >= S.tiles_y - S.jb;
        };
        // Wait until everything item `nx` reads (or will overwrite) is finished: the 3 x 3 tiles around it
        // in its own slab (done[], phases of this launch) and, for a seam tile, the facing tiles of the
        // neighbouring slab (inbox, absolute phase counts).  `on_spin` runs while waiting.
        auto wait_deps = [&](const Item& nx, auto&& on_spin) {
            bool up, dn;
            seam_sides(nx, up, dn);
            const bool seam = (up || dn) && (d.phase_base + nx.p > 0);
            if (nx.p == 0 && !seam) return;
            const SlabDesc& S = d.s[nx.s];
            const int per_img = d.tiles_x * S.tiles_y;
            unsigned long long t0 = 0;
            for (uint32_t spins = 0;; ++spins) {
                int dep = 0x7fffffff;                 // lanes 0..8: one neighbour counter each
                if (lane < 9) {
                    const int yy = nx.by + lane / 3 - 1, xx = nx.bx + lane % 3 - 1;
                    if (nx.p > 0 && yy >= 0 && yy < S.tiles_y && xx >= 0 && xx < d.tiles_x)
                        dep = ld_acquire_gpu(done + S.tile0 + nx.b * per_img + yy * d.tiles_x + xx);
                } else if (seam && lane < 9 + 6 * SEAM_JMAX) {   // lanes 9..: the neighbours' flags for 3 tile columns
                    const int q = lane - 9;
                    const int side = q / (3 * SEAM_JMAX), j = (q / 3) % SEAM_JMAX, xx = nx.bx + q % 3 - 1;
                    const bool want = side == 0 ? (up && j < S.up_j) : (dn && j < S.dn_j);
                    if (want && xx >= 0 && xx < d.tiles_x)
                        dep = ld_acquire_sys(S.inbox + (side * SEAM_JMAX + j) * d.tiles_x + xx) - d.phase_base;
                }
                if (__reduce_min_sync(0xffffffffu, dep) >= nx.p) break;   // phase p-1 is complete around us
                on_spin();
                // no sleep: the counter loads take an L2 round trip each, that is pause enough, and
                // a tile with little slack (ntiles / #CTAs is < 3 rounds at 1080p) should be
                // picked up the moment its last neighbour is published
                if ((spins & 0xff) == 0xff) {
                    const unsigned long long now = global_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > WAIT_LIMIT_NS) __trap();
                }
            }
            // Every lane that observed a counter did so with an ACQUIRE load; a warp barrier carries the
            // ordering over to lane 0, whose proxy fence extends it to the async-proxy (TMA) reads
            // issued below (the publisher has the matching proxy fence before its release).
            __syncwarp();
            if (lane == 0) fence_proxy_async_global();
        };
        // the last three decoded items: publishing lags issuing by at most two
        Item it_a, it_b, it_c = next_item();              // items mi-2, mi-1, mi
        it_a = it_b = it_c;
        wait_deps(it_c, [] {});
        if (lane == 0) issue(it_c, 0);

        // The events the producer reacts to alternate strictly in time: "item m-2 stored" -> publish
        // it;  "stage free" (the compute warps passed the barrier of item m-1) -> issue item m.
        // Both are mbarrier waits that hardware-suspend the warp, so it reacts within ~100 cycles
        // and steals no issue slots.  The only polling loop is the wait for the neighbours'
        // counters; it keeps publishing meanwhile, so this CTA never sits on a finished tile that
        // somebody else is waiting for (no cycle in the wait graph: dependencies always point to
        // items that are earlier in every CTA's order).
        bool any_seam = false;
#pragma unroll
        for (int q = 0; q < MAXS; ++q)
            if (q < d.nslabs && (d.s[q].out_up || d.s[q].out_dn)) any_seam = true;
        const bool publish = phases > 1 || any_seam;
        int np = 0;                   // next item to publish
        int cur_mi = 0;               // index of the item held in it_c
        auto do_publish = [&](int n) {
            // all compute warps have stored item n (their arrive is ordered after their stores);
            // make those stores visible GPU-wide, then bump the tile's counter
            if (lane == 0) {
                const int age = cur_mi - n;               // 0, 1 or 2 (selected field by field: stays in registers)
#define HS_PICK(f) (age == 0 ? it_c.f : (age == 1 ? it_b.f : it_a.f))
                const int t = HS_PICK(t), p = HS_PICK(p), si = HS_PICK(s), pu = HS_PICK(pub_up), pd = HS_PICK(pub_dn);
#undef HS_PICK
                const bool seam = (pu & pd) != -1;
                // the tile's rows were written through the generic proxy and will be read by TMA (async
                // proxy) on other SMs / GPUs: proxy fence, then ONE release (a MEMBAR that also covers the
                // compute warps' stores: they happen-before this thread through the stored[] mbarrier)
                fence_proxy_async_global();
                if (!seam) {
                    st_release_gpu(done + t, p + 1);
                } else {
                    // seam tiles: the halo rows are in the neighbour's memory once the system-scope fence
                    // completes; then tell the local and the neighbour's producer warps
                    fence_acq_rel_sys();
                    st_relaxed_gpu(done + t, p + 1);
                    const SlabDesc& S = d.s[MAXS > 1 ? si : 0];
                    if (pu >= 0) st_relaxed_sys(S.out_up + pu, d.phase_base + p + 1);
                    if (pd >= 0) st_relaxed_sys(S.out_dn + pd, d.phase_base + p + 1);
                }
            }
        };
        for (int mi = 1; mi < items; ++mi) {
            // items <= mi-3 are published (step (2) of the previous round), so slot a can be recycled
            it_a = it_b; it_b = it_c; it_c = next_item();
            cur_mi = mi;
            const Item& nx = it_c;
            // (1) Dependencies first, while the compute warps are still busy with item mi-2/mi-1:
            // they are normally satisfied a whole phase ahead, and the L2 round trips of the
            // counter loads and of the two fences (~1.5k cycles each under load) stay off the
            // critical path between "stage free" and the TMA issue.
#ifdef HS_TILE_PROFILE
            const long long prof_t0 = clock64();
#endif
            wait_deps(nx, [&] {
                if (publish && np < mi && mbar_test(&stored[np & 1], (np >> 1) & 1)) {
                    do_publish(np);
                    ++np;
                }
            });
#ifdef HS_TILE_PROFILE
            const long long prof_t1 = clock64();
#endif
            // (2) items <= mi-2 are stored by now or about to be: block on them, publish promptly
            if (publish)
                for (; np <= mi - 2; ++np) {
                    mbar_wait(&stored[np & 1], (np >> 1) & 1);
                    do_publish(np);
                }
            // (3) the stage of item mi is free once the compute warps passed the barrier of item mi-1
            if (mi >= 2) mbar_wait(&empty[mi & 1], ((mi - 2) >> 1) & 1);
            if (lane == 0) {
                issue(nx, mi & 1);
#ifdef HS_TILE_PROFILE
                if (mi >= 2) {
                    g_tile_prof[blockIdx.x][7] += clock64() - reinterpret_cast<volatile long long*>(smem + TS::OFF_BAR + 48)[0];
                    g_tile_prof[blockIdx.x][5] += prof_t1 - prof_t0;      // decode + dependency wait + fences
                }
#endif
            }
            __syncwarp();
        }
        if (publish)
            for (; np < items; ++np) {
                mbar_wait(&stored[np & 1], (np >> 1) & 1);
                do_publish(np);
            }
        return;
    }

    // ================================ compute warps ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int row0 = warp * R;                        // first tile row of this thread's patch
    for (int n = 0; n < items; ++n) {
        const int stage = n & 1;
        unsigned char* st = smem + (size_t)stage * TS::STAGE;
        const float4* s_uv = reinterpret_cast<const float4*>(st + TS::OFF_UV);   // two pixels {u,v,u,v} per float4
        const uint32_t* s_cpk = reinterpret_cast<const uint32_t*>(st + TS::OFF_CPK);
        const float* s_inv = reinterpret_cast<const float*>(st + TS::OFF_INV);

        HS_PROF_T(pt0);
        mbar_wait(&full[stage], (n >> 1) & 1);
        HS_PROF_T(pt1);

        const int4 info = reinterpret_cast<const int4*>(smem + TS::OFF_INFO)[2 * stage];   // decoded by the producer
        const int tx0 = info.x, ty0 = info.y, b = info.z, cur_p = info.w;
        const int si = MAXS > 1 ? reinterpret_cast<const int4*>(smem + TS::OFF_INFO)[2 * stage + 1].x : 0;
        const SlabHot& S = *reinterpret_cast<const SlabHot*>(smem + TS::OFF_HOT + TS::HOT_BYTES * si);
        const int gW = S.W, gH = S.H;
        const int gx0 = tx0 + lane * 4;
        const int gy0 = ty0 + row0;
        const bool tile_inside = (tx0 >= 0) && (tx0 + TS::SX <= gW) && (ty0 >= 0) && (ty0 + TS::SY <= gH);
        // which of the patch pixels lie inside the image (bit j*4+c); interior tiles never look at it
        uint32_t inmask = 0xffffffffu;
        if (!tile_inside) {
            inmask = 0;
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool in = (gy0 + j >= 0) && (gy0 + j < gH) && (gx0 + c >= 0) && (gx0 + c < gW);
                    inmask |= (in ? 1u : 0u) << (j * 4 + c);
                }
        }

        const int swz = ((smem_u32(s_uv) >> 7) ^ (lane >> 2)) & 1;   // address bit 7 of the lane's chunk (rows are 1 KB)
        float2 uv[R][4], gxy[R][4];        // {u, v} and the normalised gradients {P, Q} per pixel: packed-fp32 operands
        float it[R][4];                    // R = It * s
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int so = (row0 + j) * TS::SX + lane * 4;
            // the lane's 32-byte chunk; SWIZZLE_32B put its first half into the upper 16 bytes in every other
            // 128-byte line (lanes 4-7 of each group of 8): reading "my first half" is then conflict-free
            const float4 q0 = s_uv[so / 2 + swz], q1 = s_uv[so / 2 + (swz ^ 1)];
            const float4 qi = *reinterpret_cast<const float4*>(s_inv + so);
            const uint4 qc = *reinterpret_cast<const uint4*>(s_cpk + so);
            uv[j][0] = make_float2(q0.x, q0.y); uv[j][1] = make_float2(q0.z, q0.w);
            uv[j][2] = make_float2(q1.x, q1.y); uv[j][3] = make_float2(q1.z, q1.w);
            const float sc[4] = {qi.x, qi.y, qi.z, qi.w};     // second plane: s = 1 / sqrt(den) (textbook mode: It)
            const uint32_t wc[4] = {qc.x, qc.y, qc.z, qc.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {      // (Ix, Iy, It) -> (P, Q, R): the same three multiplies as hs_update_bar
                float ix, iy, itv, sv;
                if (TB) {                      // s from the gradients
                    unpack_coef_tb(wc[c], ix, iy);
                    itv = sc[c];
                    sv = hs_inv(ix, iy, d.alpha2);
                } else {
                    unpack_coef(wc[c], ix, iy, itv);
                    sv = sc[c];
                }
                gxy[j][c] = __fmul2_rn(make_float2(ix, iy), make_float2(sv, sv));
                it[j][c] = __fmul_rn(itv, sv);
            }
        }
        // Everyone has (a) finished the previous tile - its exchange scratch in the OTHER stage is
        // dead - and (b) pulled this tile out of THIS stage, which now becomes the exchange scratch.
        fence_proxy_async();      // our generic-proxy scratch accesses of the other stage, before TMA rewrites it
        compute_sync<TS::CTHREADS>();
#ifdef HS_TILE_PROFILE
        if (tid == 0) reinterpret_cast<volatile long long*>(smem + TS::OFF_BAR + 48)[0] = clock64();
#endif
        if (tid == 0 && n >= 1) mbar_arrive(&empty[stage ^ 1]);   // producer may refill it with item n+1
        HS_PROF_T(pt2);

        float* s_ex = reinterpret_cast<float*>(st);
        const int kk = min(d.k, d.sweeps - cur_p * d.k);
        if (tile_inside)
            tile_sweeps<RL, RR, R, NWARP, false, TB>(uv, gxy, it, s_ex, kk, d.kf, warp, lane, inmask);
        else
            tile_sweeps<RL, RR, R, NWARP, true, TB>(uv, gxy, it, s_ex, kk, d.kf, warp, lane, inmask);

        HS_PROF_T(pt3);
        // store the exact centre of the tile into the other pair of planes
        const int lx = lane * 4;
        const int seam = S.seam;
        if (lx >= d.hxl && lx < d.hxl + d.vx && gx0 < gW) {
            const int wr = (d.cur ^ cur_p ^ 1) & 1;       // planes this phase writes
            const int pitch = S.pitch, oy1 = S.oy1, hyt = S.hyt, ylim = S.hyt + S.vy;
            float2* UV = S.uv[wr] + (size_t)b * S.plane;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int ly = row0 + j;
                const int gy = gy0 + j;
                if (ly >= hyt && ly < ylim && gy < oy1) {
                    st_global_256(UV + (size_t)gy * pitch + gx0, uv[j]);
                }
            }
            // row-slab seams: rows that are halo rows of a neighbouring slab also go straight into that
            // slab's planes (peer memory over NVLink; all slabs flip their planes in lock step).  A
            // CTA-uniform branch: tiles of a slab without neighbours never see this code.
            if (seam) {
                float2* const PU = S.up_uv[wr];
                float2* const PD = S.dn_uv[wr];
                const int up_lim = S.oy0 + S.push_up, dn_lim = oy1 - S.push_dn;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int ly = row0 + j;
                    const int gy = gy0 + j;
                    if (ly >= hyt && ly < ylim && gy < oy1) {
                        if (PU && gy < up_lim) st_global_256(PU + (size_t)(gy + S.up_dy) * pitch + gx0, uv[j]);
                        if (PD && gy >= dn_lim) st_global_256(PD + (size_t)(gy + S.dn_dy) * pitch + gx0, uv[j]);
                    }
                }
            }
        }
        if (phases > 1 || seam) {     // tell the producer this warp's share of the tile is stored
            __syncwarp();
            if (lane == 0) mbar_arrive(&stored[stage]);
        }
        HS_PROF_T(pt4);
        HS_PROF_ADD(0, pt0, pt1); HS_PROF_ADD(1, pt1, pt2); HS_PROF_ADD(2, pt2, pt3); HS_PROF_ADD(3, pt3, pt4);
        HS_PROF_ADD(4, 0, 1);
    }
}

// ------------------------------------------------------------------------------------------
// K5: outputs.  fp32 planes -> fp32/fp64 dense-ish host layout staging (pitch kept), and the
// gradient planes back out of the packed coefficient layout.
// ------------------------------------------------------------------------------------------
// {u, v} interleaved -> two planes of T (double: the reference's CV_64FC1; float: plain split)
template <typename T>
__global__ void __launch_bounds__(256)
k_split(const float2* __restrict__ uv, T* __restrict__ oa, T* __restrict__ ob, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = i; p < n; p += stride) {
        const float2 q = uv[p];
        oa[p] = (T)q.x;
        ob[p] = (T)q.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_unpack_grad_tb(const uint32_t* __restrict__ cpk, const float* __restrict__ itp, T* __restrict__ gx,
                 T* __restrict__ gy, T* __restrict__ gt, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = i; p < n; p += stride) {
        float ix, iy;
        unpack_coef_tb(cpk[p], ix, iy);
        gx[p] = (T)ix;
        gy[p] = (T)iy;
        gt[p] = (T)itp[p];
    }
}

// max(|a1-a0|, |b1-b0|) over a pitched plane pair (rows [row0,row1), W columns, `batch` images):
// the residual of the early-exit extension.  Non-negative floats order like their bit patterns,
// so the grid-wide maximum is an atomicMax on the raw bits; NaN counts as +inf.
__global__ void __launch_bounds__(256)
k_max_abs_diff(const float2* __restrict__ p0, const float2* __restrict__ p1, Geom g, int batch,
               unsigned int* __restrict__ out) {
    const int rows = g.oy1 - g.oy0;
    const long long per = (long long)rows * g.W;
    const long long n = per * batch;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const long long r = i - (long long)b * per;
        const int y = g.oy0 + (int)(r / g.W), x = (int)(r % g.W);
        const size_t o = (size_t)b * g.plane + (size_t)y * g.pitch + x;
        const float2 q0 = p0[o], q1 = p1[o];
        float d = fmaxf(fabsf(q1.x - q0.x), fabsf(q1.y - q0.y));
        if (!(d == d) || !(q1.x == q1.x) || !(q1.y == q1.y)) d = __int_as_float(0x7f800000);
        m = fmaxf(m, d);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// Flow samples on the plot grid: plotFlow::plotBresenhamLine reads u, v only at rows/columns that
// are multiples of `delta` (plotFlow.cpp:70-75); this gathers just those (row-major grid order).
__global__ void __launch_bounds__(256)
k_sample_grid(const float2* __restrict__ uv, double* __restrict__ ou,
              double* __restrict__ ov, int pitch, int row0, int delta, int ny, int nx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    const int gy = i / nx, gx = i - gy * nx;
    const size_t o = (size_t)(row0 + gy * delta) * pitch + (size_t)gx * delta;
    const float2 q = uv[o];
    ou[i] = (double)q.x;
    ov[i] = (double)q.y;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_unpack_grad(const uint32_t* __restrict__ cpk, T* __restrict__ gx, T* __restrict__ gy,
              T* __restrict__ gt, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = i; p < n; p += stride) {
        float ix, iy, it;
        unpack_coef(cpk[p], ix, iy, it);
        gx[p] = (T)ix;
        gy[p] = (T)iy;
        gt[p] = (T)it;
    }
}

}  // namespace hs
