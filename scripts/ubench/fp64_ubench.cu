// FP64 pipe microbenchmarks for B200 (sm_100a): dependent-chain latency, throughput vs warps/SMSP and ILP,
// co-issue of FP64 with integer/move instructions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int ILP, int MIX>
__global__ void k_chain(double *out, int iters, double a, double b, int z)
{
    double x[ILP];
    int y0 = threadIdx.x, y1 = threadIdx.x * 3;
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                x[j] = fma(x[j], a, b);
                if (MIX == 1) { y0 = y0 * z + y1; }                  // one IMAD per DFMA
                if (MIX == 2) { y0 = y0 * z + y1; y1 = y1 * z + y0; } // two IMADs per DFMA
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + y0 + y1;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int ILP, int MIX>
void run(const char *name, int warps_per_sm, int nsm, double *out)
{
    // one block per SM with warps_per_sm warps
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<ILP, MIX><<<nsm, warps_per_sm * 32>>>(out, 10, 0.999999, 1e-7, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_chain<ILP, MIX><<<nsm, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-7, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clk; cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    double dfma_per_warp = (double)iters * 8 * ILP;
    double warp_dfma_per_smsp = dfma_per_warp * warps_per_sm / 4.0;
    printf("%-10s warps/SM %2d ILP %d mix %d : %.2f cyc per DFMA per warp; SMSP FP64 issue interval %.2f cyc (ideal 2.0); %.2f TFLOP/s\n",
           name, warps_per_sm, ILP, MIX, clk / dfma_per_warp, clk / warp_dfma_per_smsp,
           dfma_per_warp * warps_per_sm * nsm * 32 * 2 / (ms * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double *out; cudaMalloc(&out, 8 * 2048 * nsm);
    printf("%s SMs %d\n", p.name, nsm);
    run<1, 0>("lat", 1, nsm, out);
    run<1, 0>("lat", 4, nsm, out);
    run<2, 0>("ilp", 4, nsm, out);
    run<3, 0>("ilp", 4, nsm, out);
    run<4, 0>("ilp", 4, nsm, out);
    run<8, 0>("ilp", 4, nsm, out);
    run<1, 0>("w", 8, nsm, out);
    run<2, 0>("w", 8, nsm, out);
    run<3, 0>("w", 8, nsm, out);
    run<4, 0>("w", 8, nsm, out);
    run<1, 0>("w", 12, nsm, out);
    run<2, 0>("w", 12, nsm, out);
    run<3, 0>("w", 12, nsm, out);
    run<1, 0>("w", 16, nsm, out);
    run<2, 0>("w", 16, nsm, out);
    run<1, 0>("w", 32, nsm, out);
    run<4, 1>("mix", 8, nsm, out);
    run<4, 2>("mix", 8, nsm, out);
    run<2, 1>("mix", 8, nsm, out);
    run<2, 2>("mix", 8, nsm, out);
    run<2, 1>("mix", 12, nsm, out);
    run<2, 2>("mix", 12, nsm, out);
    run<4, 1>("mix", 16, nsm, out);
    run<4, 2>("mix", 16, nsm, out);
    return 0;
}
