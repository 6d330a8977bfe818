// C++ host-side test of the drop-in adapter: the same three lines the reference's driver runs
// (HornSchunckOF/main.cpp:93-98) against cpp-optical-flow_b200/adapter/hornSchunck.cpp, compiled
// with the minicv stub instead of OpenCV.  Reads two raw uint8 frames, writes u, v (float64) and the
// gradients, so the pytest side can compare with the oracle.
//   adapter_driver <prev.raw> <next.raw> <rows> <cols> <windowSize> <maxIterations> <alpha> <out.bin> [roi|f32|f64prec]
//     roi     - non-continuous ROI views of the frames
//     f32     - the frames as CV_32F in [0,1] (value / 255): the reference takes any depth (hornSchunck.cpp:23-24)
//     f64prec - 8-bit frames, hs.precision = HS_PREC_F64 (the reference's own fp64 arithmetic)
#include "hornSchunck.cpp"
#include <cstdio>
#include <fstream>
#include <vector>

static std::vector<unsigned char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 9) { std::fprintf(stderr, "usage\n"); return 2; }
    int rows = atoi(argv[3]), cols = atoi(argv[4]);
    std::vector<unsigned char> a = slurp(argv[1]), b = slurp(argv[2]);
    cv::Mat imagePrev(rows, cols, CV_8UC1, a.data()), imageNext(rows, cols, CV_8UC1, b.data());
    const std::string mode = argc > 9 ? argv[9] : "";
    if (mode == "roi") {   // exercise non-continuous inputs: drop a 3-pixel border through ROI views
        imagePrev = imagePrev.roi(3, rows - 3, 3, cols - 3);
        imageNext = imageNext.roi(3, rows - 3, 3, cols - 3);
    }
    if (mode == "f32") {   // what cv::Mat::convertTo(CV_32F, 1.0 / 255) gives
        cv::Mat fp(rows, cols, CV_32FC1), fn(rows, cols, CV_32FC1);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                fp.at<float>(y, x) = (float)(imagePrev.at<unsigned char>(y, x) * (1.0 / 255));
                fn.at<float>(y, x) = (float)(imageNext.at<unsigned char>(y, x) * (1.0 / 255));
            }
        imagePrev = fp; imageNext = fn;
    }
    try {
        cv::Mat u, v;
        int windowSize = atoi(argv[5]);
        int maxIterations = atoi(argv[6]);
        double alpha = atof(argv[7]);
        hornSchunck hs = hornSchunck(windowSize, maxIterations, alpha);     // main.cpp:97
        if (mode == "f64prec") hs.precision = HS_PREC_F64;
        hs.getFlow(imagePrev, imageNext, u, v);                            // main.cpp:98
        cv::Mat gx, gy, gt;
        hs.getGradients(imagePrev, imageNext, gx, gy, gt);
        if (u.type() != CV_64FC1 || !u.isContinuous() || u.rows != imagePrev.rows || u.cols != imagePrev.cols) return 3;
        std::ofstream o(argv[8], std::ios::binary);
        for (cv::Mat* m : {&u, &v, &gx, &gy, &gt}) o.write((const char*)m->data, (std::streamsize)(m->step * m->rows));
        std::printf("ok %d %d u(0,0)=%g\n", u.rows, u.cols, u.at<double>(0, 0));
    } catch (const cv::Exception& e) {
        std::printf("cv::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
