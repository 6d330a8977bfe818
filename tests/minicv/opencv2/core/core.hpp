// minicv: a few dozen lines of cv::Mat, just enough to COMPILE AND RUN the C++ adapter
// (cpp-optical-flow_b200/adapter/hornSchunck.cpp) in an image that has no OpenCV C++ SDK.
// TEST INFRASTRUCTURE ONLY - it implements no image processing and is never shipped.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(d, cn) (CV_MAT_DEPTH(d) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {
namespace Error { enum { StsBadArg = -5, StsUnmatchedFormats = -205, StsUnmatchedSizes = -209, StsUnsupportedFormat = -210, GpuApiCallError = -217 }; }
class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct Size { int width, height; bool operator!=(const Size& o) const { return width != o.width || height != o.height; } };

class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    unsigned char* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), flags(type), data((unsigned char*)ext) {
        step = st ? st : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; flags = type; step = (size_t)c * elemSize();
        owner_.reset(new unsigned char[step * (size_t)r], std::default_delete<unsigned char[]>());
        data = owner_.get();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * (size_t)r); return m; }
    // ROI view: rows [y0,y1), cols [x0,x1) - shares the buffer, keeps the parent's step
    Mat roi(int y0, int y1, int x0, int x1) const {
        Mat m; m.rows = y1 - y0; m.cols = x1 - x0; m.flags = flags; m.step = step; m.owner_ = owner_;
        m.data = data + (size_t)y0 * step + (size_t)x0 * elemSize();
        return m;
    }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { static const int s[] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)s[depth()] * channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Size size() const { return Size{cols, rows}; }
    template <typename T> T* ptr(int y) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
private:
    std::shared_ptr<unsigned char> owner_;
};
}  // namespace cv
#define CV_Error(code, msg) throw cv::Exception((code), (msg))
