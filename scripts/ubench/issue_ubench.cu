// Which instruction classes steal FP64 throughput on B200?  ILP-4 DFMA chains (inline PTX, program order kept)
// with NUM independent non-FP64 instructions per DEN DFMAs (4 independent chains each, never latency-limited).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_ubench issue_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NUM, int DEN>
__global__ void k_mix(double *out, int iters, double a, double b, int z, float fz)
{
    __shared__ double sm[4 * 512];
    double x[4], l[4];
    int y[4]; float f[4];
    for (int j = 0; j < 4; j++) { x[j] = threadIdx.x + j; y[j] = threadIdx.x * (j + 1); f[j] = threadIdx.x + j; l[j] = 0; sm[threadIdx.x + 512 * j] = j; }
    __syncthreads();
    unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 8;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));
                if ((r * 4 + j) % DEN < NUM) {
                    if (KIND == 1) asm volatile("mad.lo.s32 %0, %0, %1, 7;" : "+r"(y[j]) : "r"(z));
                    if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f[j]) : "f"(fz));
                    if (KIND == 3) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l[j]) : "r"(sa + 4096 * j) : "memory");
                    if (KIND == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[j]) : "r"(z));
                    if (KIND == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(y[j]) : "r"(z));
                    if (KIND == 7) asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p; }" : "+f"(f[j]) : "f"(fz), "r"(z));
                    if (KIND == 8) asm volatile("st.shared.f64 [%1], %0;" : : "d"(x[j]), "r"(sa + 4096 * j) : "memory");
                    if (KIND == 9) asm volatile("mul.f64 %0, %0, %1;" : "+d"(l[j]) : "d"(a));
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 4; j++) s += x[j] + y[j] + f[j] + l[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int KIND, int NUM, int DEN>
void run(const char *name, int warps_per_sm, int nsm, double *out)
{
    int iters = 2000;
    k_mix<KIND, NUM, DEN><<<nsm, warps_per_sm * 32>>>(out, 10, 0.999999, 1e-7, 3, 0.99f);
    cudaDeviceSynchronize();
    k_mix<KIND, NUM, DEN><<<nsm, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-7, 3, 0.99f);
    cudaDeviceSynchronize();
    double clk; cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    double dfma_per_smsp = (double)iters * 32 * warps_per_sm / 4.0;
    printf("%-10s warps/SM %2d: %d other per %d DFMA -> %.2f cyc per DFMA per SMSP\n", name, warps_per_sm, NUM, DEN, clk / dfma_per_smsp);
}
#define ALL(K, name) run<K, 1, 4>(name, w, nsm, out); run<K, 1, 2>(name, w, nsm, out); run<K, 1, 1>(name, w, nsm, out);
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double *out; cudaMalloc(&out, 8 * 2048 * nsm);
    for (int w = 8; w <= 16; w += 4) {
        run<0, 0, 1>("DFMA only", w, nsm, out);
        ALL(1, "IMAD") ALL(2, "FFMA") ALL(3, "LDS.64") ALL(5, "LOP") ALL(6, "IADD") ALL(7, "FSEL") ALL(8, "STS.64") ALL(9, "DMUL")
    }
    return 0;
}
