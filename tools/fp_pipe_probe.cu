// Issue-rate probe for the FP32 pipes of sm_100a: how many cycles does one SM sub-partition need per
// warp instruction of FADD / FFMA / FMUL and of their packed forms FADD2 / FMUL2 / FFMA2?
// The fused Jacobi kernel's sweep is made of exactly these; the answer is its real ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp_pipe_probe tools/fp_pipe_probe.cu
//   ./tools/fp_pipe_probe            (prints cycles per warp-instruction per sub-partition)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NACC = 16;     // independent accumulators per thread (dependency distance 16 instructions)
constexpr int ITERS = 2048;

template <int OP>
__global__ void __launch_bounds__(1024) probe(float2* out, long long* cyc, float seed) {
    // a[]: 16 independent accumulators (32 registers); b, c: loop-invariant operands; d[]: a second set of
    // scalar accumulators for the mixed tests
    float2 a[NACC];
    float d[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = make_float2(seed * (i + 1) + threadIdx.x, seed * (i + 2)); d[i] = seed * i; }
    const float2 b = make_float2(1.0f + seed, 1.0f - seed), c = make_float2(seed * 0.5f, seed * 0.25f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) a[i].x = __fadd_rn(a[i].x, b.x);                                 // FADD
            if (OP == 1) a[i].x = __fmul_rn(a[i].x, b.x);                                 // FMUL
            if (OP == 2) a[i].x = __fmaf_rn(a[i].x, b.x, c.x);                            // FFMA, 3 registers
            if (OP == 3) a[i] = __fadd2_rn(a[i], b);                                      // FADD2
            if (OP == 4) a[i] = __fmul2_rn(a[i], b);                                      // FMUL2
            if (OP == 5) a[i] = __ffma2_rn(a[i], b, c);                                   // FFMA2, 3 register pairs
            if (OP == 6) { a[i].x = __fadd_rn(a[i].x, b.x); d[i] = __fadd_rn(d[i], b.y); }                      // 2 x FADD (unpaired registers)
            if (OP == 7) { a[i].x = __fmaf_rn(a[i].x, b.x, c.x); d[i] = __fmaf_rn(d[i], b.y, c.y); }            // 2 x FFMA (unpaired)
            if (OP == 8) a[i] = __fadd2_rn(a[i], a[(i + 5) % NACC]);                      // FADD2, both operands accumulators
            if (OP == 9) a[i] = __ffma2_rn(make_float2(-b.x, -b.y), make_float2(d[i], d[i]), a[i]);             // the kernel's update form
            if (OP == 10) { a[i] = __fadd2_rn(a[i], b); d[i] = __fadd_rn(d[i], b.y); }                          // FADD2 + FADD
            if (OP == 11) { a[i] = __fadd2_rn(a[i], b); d[i] = __fmaf_rn(d[i], b.y, c.x); }                     // FADD2 + FFMA
            if (OP == 12) { a[i] = __fadd2_rn(a[i], b); d[i] = __fmaf_rn(d[i], b.y, c.x); d[i] = __fmul_rn(d[i], c.y); }  // FADD2 + FFMA + FMUL
            if (OP == 13) { a[i] = __fadd2_rn(a[i], b); d[i] = __int_as_float(__float_as_int(d[i]) + (__float_as_int(b.y) ^ i)); }   // FADD2 + integer op
            if (OP == 14) { a[i].x = __fadd_rn(a[i].x, b.x); d[i] = __int_as_float(__float_as_int(d[i]) + (__float_as_int(b.y) ^ i)); }   // FADD + integer op
        }
    }
    const long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { s.x += a[i].x + d[i]; s.y += a[i].y; }
    if (s.x == 123.456f) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_step) {
    float2* out; long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(float2));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    printf("%-28s", name);
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int threads = warps_per_smsp * 4 * 32;
        probe<OP><<<148, threads>>>(out, cyc, 1e-3f);
        probe<OP><<<148, threads>>>(out, cyc, 1e-3f);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148;
        const double per = avg / ((double)ITERS * NACC * instr_per_step * warps_per_smsp);
        printf("  %dw/smsp: %.2f cyc/inst", warps_per_smsp, per);
    }
    printf("\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    printf("cycles per warp-instruction per SM sub-partition (lower bound 1.0 = one issue slot)\n");
    run<0>("FADD", 1);
    run<1>("FMUL", 1);
    run<2>("FFMA r,r,r", 1);
    run<3>("FADD2", 1);
    run<4>("FMUL2", 1);
    run<5>("FFMA2 r,r,r", 1);
    run<6>("2 x FADD (same work as FADD2)", 2);
    run<7>("2 x FFMA (same as FFMA2)", 2);
    run<8>("FADD2 acc,acc", 1);
    run<9>("FFMA2 -b, {c,c}, a", 1);
    run<10>("FADD2 + FADD (per pair)", 1);
    run<11>("FADD2 + FFMA (per pair)", 1);
    run<12>("FADD2 + FFMA + FMUL (per triple)", 1);
    run<13>("FADD2 + int (per pair)", 1);
    run<14>("FADD + int (per pair)", 1);
    return 0;
}
