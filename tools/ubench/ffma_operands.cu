// Micro-benchmark: FFMA issue rate on sm_100a as a function of how the multiplier operand changes between consecutive
// instructions (operand-reuse cache) with per-instruction distinct addend registers (the Horner pattern of covgain.cu).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma_operands tools/ubench/ffma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int NF = 32;   // FFMAs per loop iteration
constexpr int NCH = 8;   // independent chains

template <int MODE>
__device__ __forceinline__ int mul_index(int i)
{
    return MODE == 0 ? 0 : MODE == 1 ? (i & 1) : MODE == 2 ? ((i >> 1) & 1) : MODE == 3 ? (i & 3) : MODE == 4 ? ((i >> 2) & 3)
         : MODE == 5 ? ((i >> 3) & 3) : 0;
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(float *out, long long *cycles, const float *cin)
{
    float c[NF], x[4], a[NCH];
#pragma unroll
    for (int i = 0; i < NF; ++i) c[i] = cin[i + (threadIdx.x & 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = cin[40 + i];
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) a[i % NCH] = fmaf(a[i % NCH], x[mul_index<MODE>(i)], c[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int THREADS>
void run(const char *name, float *out, long long *cyc, const float *cin)
{
    for (int rep = 0; rep < 2; ++rep) {
        k<MODE, THREADS><<<148, THREADS>>>(out, cyc, cin);
        cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += h[i];
    mean /= 148;
    const double warp_instr = double(ITERS) * NF * (THREADS / 32) / 4.0;  // per SMSP
    printf("%-44s warps/SMSP %d  FFMA/clk/SMSP %.3f  (loop ceiling %.3f)\n", name, THREADS / 128, warp_instr / mean,
           double(NF) / (NF + 3));
}

int main()
{
    float *out, *cin;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&cin, 64 * 4);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1e-3f * (i + 1);
    for (int i = 40; i < 44; ++i) h[i] = 0.999f + 1e-4f * i;
    cudaMemcpy(cin, h, sizeof(h), cudaMemcpyHostToDevice);
#define RUN(M, NAME) run<M, 512>(NAME, out, cyc, cin); run<M, 1024>(NAME, out, cyc, cin)
    RUN(0, "one multiplier");
    RUN(1, "2 multipliers alternating every FFMA");
    RUN(2, "2 multipliers alternating every 2 FFMA");
    RUN(3, "4 multipliers cyclic every FFMA");
    RUN(4, "4 multipliers, runs of 4");
    RUN(5, "4 multipliers, runs of 8");
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
