// Micro-benchmark: per-SM issue rate of FFMA (3-register), FFMA (immediate addend), FFMA2 (fma.rn.f32x2) and
// MUFU on sm_100a.  Decides how the coverage-gain inner loop should be written.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fma_rate tools/ubench/fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *out, long long *cycles, float x, float y)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float2 xx = make_float2(x, x * 1.0001f), yy = make_float2(y, y * 0.999f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, 1.25f);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long acc, xv, yv;
                asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[i]), "f"(a[i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(xx.x), "f"(xx.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(yv) : "f"(yy.x), "f"(yy.y));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(xv), "l"(yv));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(acc));
            }
        } else if (MODE == 3) {  // horner-like: coefficient operand differs per fma (4 distinct registers)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, a[(i + 5) & 15]);
        } else if (MODE == 4) {  // MUFU.EX2
#pragma unroll
            for (int i = 0; i < 16; ++i) asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 5) {  // 12 FFMA + 1 MUFU mix
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = fmaf(a[i], x, y);
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[12]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter, float *out, long long *cyc)
{
    k<MODE><<<148, 1024>>>(out, cyc, 1.0001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, cyc, 1.0001f, 1e-7f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += h[i];
    mean /= 148;
    double ops = double(ITERS) * per_iter * 1024;
    printf("%-28s cycles/SM %.0f  lane-ops/clk/SM %.1f  warp-instr/clk/SMSP %.3f  time %.3f ms (%.0f MHz)\n", name, mean,
           ops / mean, ops / mean / 128.0 * (MODE == 2 ? 0.5 : 1.0), ms, mean / ms * 1e-3);
}

int main()
{
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA a=a*x+y (3 regs)", 16, out, cyc);
    run<1>("FFMA a=a*x+imm", 16, out, cyc);
    run<2>("FFMA2 f32x2 (ops=2/lane)", 16, out, cyc);
    run<3>("FFMA a=a*x+b (4 regs)", 16, out, cyc);
    run<4>("MUFU.EX2", 16, out, cyc);
    run<5>("12 FFMA + 1 MUFU", 13, out, cyc);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
