// hornSchunck.cpp - drop-in replacement for HornSchunckOF/hornSchunck.cpp of
// liuyang9609/Cpp-Optical-Flow.  Same file name, same class, same public surface:
//     int windowSize, maxIterations; double alpha;               (reference :10-11)
//     hornSchunck(int inpWindowSize, int inpMaxIterations, double inpAlpha)      (:13-17)
//     void getGradients(cv::Mat, cv::Mat, cv::Mat&, cv::Mat&, cv::Mat&)          (:19-41)
//     void getFlow(cv::Mat, cv::Mat, cv::Mat&, cv::Mat&)                         (:43-75)
// so the reference's driver compiles unchanged against it (HornSchunckOF/main.cpp:8 includes this
// file by name, :97-98 is the only call site).  Nothing is computed here: both methods forward to
// the C ABI of libhs_b200.so (include/hs.h), which runs hand-written sm_100a CUDA kernels.
// There is no CPU fallback - without the library or without a B200 the calls throw cv::Exception,
// the same way the reference throws from inside OpenCV on bad input.
//
// Outputs are freshly allocated, continuous CV_64FC1 H x W Mats (the caller's headers are
// replaced, like `u = cv::Mat::zeros(...)` / `u = uAvg - uUpdateConst` do upstream); inputs may be
// non-continuous (ROI views): Mat::step is passed through.  Frames of ANY depth are accepted, as
// upstream (convertTo(CV_64FC1), :23-24): 8UC1 - what main.cpp's preprocess() produces - and any
// other depth whose values are all integers in [0,255] run on the fast fp32 path (within 1e-4 px of
// the reference's fp64 result); everything else (CV_32F in [0,1], 16-bit, negative values ...) runs
// on the library's HS_PREC_F64 path in its own depth, which repeats the reference's fp64 arithmetic
// bit for bit.  Set `precision = HS_PREC_F64` to force that path for 8-bit frames too.
#ifndef HS_B200_HORNSCHUNCK_ADAPTER
#define HS_B200_HORNSCHUNCK_ADAPTER

#include "opencv2/imgproc/imgproc.hpp"
#include "opencv2/highgui/highgui.hpp"
#include <opencv2/core/core.hpp>
#include <stdlib.h>
#include <stdio.h>
#include <iostream>
#include <string>
#include <vector>

#include "hs.h"

class hornSchunck{
public:
    int windowSize, maxIterations;
    double alpha;
    int precision = HS_PREC_F32;      // extension (not in the reference): HS_PREC_F64 = fp64 arithmetic for every frame

    hornSchunck(int inpWindowSize, int inpMaxIterations, double inpAlpha)
        : windowSize(inpWindowSize), maxIterations(inpMaxIterations), alpha(inpAlpha) {}

    // `hornSchunck hs = hornSchunck(w, T, alpha);` (main.cpp:97) needs copy/move; a copy starts
    // without a device context and creates its own on first use.
    hornSchunck(const hornSchunck& o) : windowSize(o.windowSize), maxIterations(o.maxIterations), alpha(o.alpha) {}
    hornSchunck& operator=(const hornSchunck& o) {
        if (this != &o) { release(); windowSize = o.windowSize; maxIterations = o.maxIterations; alpha = o.alpha; }
        return *this;
    }
    ~hornSchunck() { release(); }

    void getGradients(cv::Mat imagePrev, cv::Mat imageNext, cv::Mat &gradX, cv::Mat &gradY, cv::Mat &gradT){
        cv::Mat p8, n8;
        int fdt = HS_FRAME_U8;
        checkPair(imagePrev, imageNext, p8, n8, fdt);
        hs_ctx* ctx = context(p8.cols, p8.rows, fdt);
        cv::Mat gx(p8.rows, p8.cols, CV_64FC1), gy(p8.rows, p8.cols, CV_64FC1), gt(p8.rows, p8.cols, CV_64FC1);
        check(hs_gradients(ctx, p8.data, (size_t)p8.step, n8.data, (size_t)n8.step,
                           gx.data, gy.data, gt.data, (size_t)gx.step, HS_F64), ctx);
        gradX = gx; gradY = gy; gradT = gt;
    }

    void getFlow(cv::Mat imagePrev, cv::Mat imageNext, cv::Mat &u, cv::Mat &v){
        cv::Mat p8, n8;
        int fdt = HS_FRAME_U8;
        checkPair(imagePrev, imageNext, p8, n8, fdt);
        hs_ctx* ctx = context(p8.cols, p8.rows, fdt);
        cv::Mat uu(p8.rows, p8.cols, CV_64FC1), vv(p8.rows, p8.cols, CV_64FC1);
        check(hs_solve(ctx, p8.data, (size_t)p8.step, 0, n8.data, (size_t)n8.step, 0,
                       uu.data, (size_t)uu.step, 0, vv.data, (size_t)vv.step, 0, HS_F64), ctx);
        u = uu; v = vv;
    }

private:
    hs_ctx* ctx_ = nullptr;
    int cw_ = 0, ch_ = 0, cwin_ = 0, cit_ = 0, cfdt_ = 0, cprec_ = 0;
    double calpha_ = 0;

    void release() { if (ctx_) { hs_destroy(ctx_); ctx_ = nullptr; } }

    static void check(int rc, const hs_ctx* ctx) {
        if (rc != HS_OK) {
            const char* msg = hs_last_error(ctx);
            CV_Error(rc == HS_ERR_INVALID_ARG ? cv::Error::StsBadArg : cv::Error::GpuApiCallError,
                     std::string("hs_b200: ") + (msg ? msg : "unknown error"));
        }
    }

    // one context per (geometry, parameters); the public fields may be changed between calls
    hs_ctx* context(int width, int height, int fdt) {
        const int prec = (fdt != HS_FRAME_U8 || precision == HS_PREC_F64) ? HS_PREC_F64 : HS_PREC_F32;
        if (ctx_ && cw_ == width && ch_ == height && cwin_ == windowSize && cit_ == maxIterations && calpha_ == alpha &&
            cfdt_ == fdt && cprec_ == prec)
            return ctx_;
        release();
        hs_config cfg = {};
        cfg.struct_size = sizeof(cfg);
        cfg.width = width; cfg.height = height;
        cfg.window_size = windowSize; cfg.max_iterations = maxIterations; cfg.alpha = alpha;
        cfg.batch = 1; cfg.device = -1;
        cfg.precision = prec; cfg.frame_dtype = fdt;
        hs_ctx* c = nullptr;
        check(hs_create(&cfg, &c), nullptr);
        ctx_ = c; cw_ = width; ch_ = height; cwin_ = windowSize; cit_ = maxIterations; calpha_ = alpha;
        cfdt_ = fdt; cprec_ = prec;
        return ctx_;
    }

    // lossless narrowing to 8-bit; false when some value is not an integer in [0, 255]
    template <typename T>
    static bool narrow(const cv::Mat& src, cv::Mat& dst) {
        dst = cv::Mat(src.rows, src.cols, CV_8UC1);
        for (int y = 0; y < src.rows; ++y) {
            const T* s = src.ptr<T>(y);
            unsigned char* d = dst.ptr<unsigned char>(y);
            for (int x = 0; x < src.cols; ++x) {
                const T val = s[x];
                if (!(val >= (T)0 && val <= (T)255) || (T)(unsigned char)val != val) return false;
                d[x] = (unsigned char)val;
            }
        }
        return true;
    }

    static bool to8u(const cv::Mat& src, cv::Mat& dst) {
        switch (src.depth()) {
            case CV_8U:  dst = src; return true;                 // shared header, no copy
            case CV_8S:  return narrow<signed char>(src, dst);
            case CV_16U: return narrow<unsigned short>(src, dst);
            case CV_16S: return narrow<short>(src, dst);
            case CV_32S: return narrow<int>(src, dst);
            case CV_32F: return narrow<float>(src, dst);
            case CV_64F: return narrow<double>(src, dst);
            default: CV_Error(cv::Error::StsUnsupportedFormat, "hs_b200: unsupported frame depth");
        }
        return false;
    }

    static int frameDtype(int depth) {
        switch (depth) {
            case CV_8U: return HS_FRAME_U8;   case CV_8S: return HS_FRAME_S8;
            case CV_16U: return HS_FRAME_U16; case CV_16S: return HS_FRAME_S16;
            case CV_32S: return HS_FRAME_S32; case CV_32F: return HS_FRAME_F32;
            case CV_64F: return HS_FRAME_F64;
            default: CV_Error(cv::Error::StsUnsupportedFormat, "hs_b200: unsupported frame depth");
        }
        return HS_FRAME_U8;
    }

    // a, b -> the frames the library gets (shared headers where possible) and their hs_frame_dtype
    static void checkPair(const cv::Mat& a, const cv::Mat& b, cv::Mat& fa, cv::Mat& fb, int& fdt) {
        if (a.empty() || b.empty()) CV_Error(cv::Error::StsBadArg, "hs_b200: empty frame");
        if (a.channels() != 1 || b.channels() != 1)
            CV_Error(cv::Error::StsBadArg, "hs_b200: frames must be single-channel (see preprocess(), main.cpp:11-26)");
        if (a.rows != b.rows || a.cols != b.cols)
            CV_Error(cv::Error::StsUnmatchedSizes, "hs_b200: Image sizes are different (main.cpp:70-73)");
        if (to8u(a, fa) && to8u(b, fb)) { fdt = HS_FRAME_U8; return; }
        if (a.depth() != b.depth()) CV_Error(cv::Error::StsUnmatchedFormats, "hs_b200: the two frames must have the same depth");
        fa = a; fb = b;                                          // any other depth: fp64 path, frames as they are
        fdt = frameDtype(a.depth());
    }
};

#endif  // HS_B200_HORNSCHUNCK_ADAPTER
