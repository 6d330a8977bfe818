// Timing / dump driver for the UNMODIFIED reference class -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiled by `make -C oracle ref` only where an OpenCV C++ SDK exists (pkg-config opencv4); the
// reference's hornSchunck.cpp is included from where it lies (-I$(HS_REFERENCE_DIR)/HornSchunckOF),
// nothing of it is copied into this repository.  The build image has no OpenCV C++ SDK, so there
// oracle/_ref/ stays empty and the cv2 restatement (oracle/hs_oracle.py) is the CPU reference.
//   hs_ref <prev.raw> <next.raw> <rows> <cols> <windowSize> <maxIterations> <alpha> [out.bin]
// prints one line:  seconds=<wall seconds of getFlow> threads=<cv::getNumThreads()>
#include "hornSchunck.cpp"   // the reference's file (HornSchunckOF/hornSchunck.cpp:8-76)

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>

static std::vector<unsigned char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 8) {
        std::fprintf(stderr, "usage: hs_ref prev.raw next.raw rows cols windowSize maxIterations alpha [out.bin]\n");
        return 2;
    }
    const int rows = std::atoi(argv[3]), cols = std::atoi(argv[4]);
    std::vector<unsigned char> a = slurp(argv[1]), b = slurp(argv[2]);
    if ((int)a.size() != rows * cols || (int)b.size() != rows * cols) return 3;
    cv::Mat prev(rows, cols, CV_8UC1, a.data()), next(rows, cols, CV_8UC1, b.data());
    hornSchunck hs = hornSchunck(std::atoi(argv[5]), std::atoi(argv[6]), std::atof(argv[7]));   // main.cpp:97
    cv::Mat u, v;
    const auto t0 = std::chrono::steady_clock::now();
    hs.getFlow(prev, next, u, v);                                                                // main.cpp:98
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("seconds=%.6f threads=%d\n", s, cv::getNumThreads());
    if (argc > 8) {
        std::ofstream o(argv[8], std::ios::binary);
        o.write((const char*)u.data, (std::streamsize)(u.total() * u.elemSize()));
        o.write((const char*)v.data, (std::streamsize)(v.total() * v.elemSize()));
    }
    return 0;
}
