// Throughput of the shuffle / shared-memory data path of one SM on sm_100a: cycles per warp instruction,
// SM-wide, for SHFL, LDS.128, STS.128 (conflict-free) with 4, 8, 12 and 16 warps resident.  The fused Jacobi
// kernel's sweep is 16 SHFL + 4 STS.128 + 4 LDS.128 per warp; this tells what that costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lsu_probe tools/lsu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 1024;
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return q;
}
__device__ __forceinline__ void sts128(float4* p, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int OP>
__global__ void __launch_bounds__(512) probe(float* out, long long* cyc) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.5f + i;
    float4* my = sm + warp * 256 + lane;                 // 4 KB per warp, lane stride 16 B: conflict-free
    for (int i = 0; i < 8; ++i) my[32 * i] = make_float4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (OP == 0) {            // 16 independent shuffles
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
        } else if (OP == 1) {     // 4 x LDS.128 (16 registers)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = lds128(&my[32 * ((i + it) & 7)]);
                v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
        } else if (OP == 2) {     // 4 x STS.128
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sts128(&my[32 * ((i + it) & 7)], v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (OP == 3) {     // the sweep's mix: 16 SHFL + 4 STS.128 + 4 LDS.128
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sts128(&my[32 * i], v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = lds128(&my[32 * (i + 4)]);
                v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 1.2345f) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(probe<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096);
    printf("%-40s", name);
    for (int warps : {4, 8, 12, 16}) {
        probe<OP><<<148, warps * 32, 16 * 4096>>>(out, cyc);
        probe<OP><<<148, warps * 32, 16 * 4096>>>(out, cyc);
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("  %2d warps: %.2f cyc/inst/SM", warps, avg / ((double)ITERS * instr * warps));
    }
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}
int main() {
    printf("SM-wide cycles per warp instruction (1.0 = one instruction per clock for the whole SM)\n");
    run<0>("SHFL.UP (32-bit)", 16);
    run<1>("LDS.128 conflict-free", 4);
    run<2>("STS.128 conflict-free", 4);
    run<3>("16 SHFL + 4 STS.128 + 4 LDS.128 (per instr)", 24);
    return 0;
}
