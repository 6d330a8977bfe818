// How much do pitched (2-D) host<->device copies cost against flat ones for the frame sizes of the
// configs?  Pinned host memory, one stream, cudaEvent timing, median of 20.
//   nvcc -O2 -o tools/memcpy_probe tools/memcpy_probe.cu && ./tools/memcpy_probe
#include <algorithm>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

static float med(std::vector<float>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main() {
    struct Case { const char* name; int W, H; } cases[] = {{"kitti 1242x375", 1242, 375}, {"1080p", 1920, 1080}};
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (auto& c : cases) {
        for (int es : {1, 8}) {
            const size_t row = (size_t)c.W * es, dpitch = ((c.W + 127) / 128 * 128) * (size_t)es;
            void *h, *d;
            cudaMallocHost(&h, row * c.H); cudaMalloc(&d, dpitch * c.H);
            for (int dir = 0; dir < 2; ++dir) {
                std::vector<float> t2, t1, t2x2, t1x2;
                for (int it = 0; it < 20; ++it) {
                    float ms;
                    cudaEventRecord(e0, s);
                    if (dir == 0) cudaMemcpy2DAsync(d, dpitch, h, row, row, c.H, cudaMemcpyHostToDevice, s);
                    else cudaMemcpy2DAsync(h, row, d, dpitch, row, c.H, cudaMemcpyDeviceToHost, s);
                    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); t2.push_back(ms);
                    cudaEventRecord(e0, s);
                    if (dir == 0) cudaMemcpyAsync(d, h, row * c.H, cudaMemcpyHostToDevice, s);
                    else cudaMemcpyAsync(h, d, row * c.H, cudaMemcpyDeviceToHost, s);
                    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); t1.push_back(ms);
                    // two copies back to back (prev + next, or u + v)
                    cudaEventRecord(e0, s);
                    for (int q = 0; q < 2; ++q) {
                        if (dir == 0) cudaMemcpy2DAsync(d, dpitch, h, row, row, c.H, cudaMemcpyHostToDevice, s);
                        else cudaMemcpy2DAsync(h, row, d, dpitch, row, c.H, cudaMemcpyDeviceToHost, s);
                    }
                    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); t2x2.push_back(ms);
                    cudaEventRecord(e0, s);
                    for (int q = 0; q < 2; ++q) {
                        if (dir == 0) cudaMemcpyAsync(d, h, row * c.H, cudaMemcpyHostToDevice, s);
                        else cudaMemcpyAsync(h, d, row * c.H, cudaMemcpyDeviceToHost, s);
                    }
                    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); t1x2.push_back(ms);
                }
                printf("%-15s %d B/px %s: pitched %.1f us, flat %.1f us | two copies: pitched %.1f us, flat %.1f us  (%.2f MB each)\n",
                       c.name, es, dir == 0 ? "H2D" : "D2H", 1e3 * med(t2), 1e3 * med(t1), 1e3 * med(t2x2), 1e3 * med(t1x2), row * c.H / 1e6);
            }
            cudaFreeHost(h); cudaFree(d);
        }
    }
    return 0;
}
