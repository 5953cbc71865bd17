// What limits a DMMA.8x8x4 stream shaped like the SYRK / TRMM k-step (8 warps per SM, 32 accumulator pairs per warp,
// 8 A and 4 B operands per k-step)?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/dmma_shape scratch/dmma_shape.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: registers only; 1: + barrier every 4 k-steps; 2: + operands from shared memory; 3: both; 4: mode 3 + 12 DADD
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters) {
    __shared__ double sm[16 * 132 * 2];
    for (int i = threadIdx.x; i < 16 * 132 * 2; i += 256) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm0 = (wid >> 2) * 64, wn0 = (wid & 3) * 32;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = i; acc[i][j][1] = -j; }
    double a[8], b[4], am[8], bm[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0 + (lane + i) * 1e-6; am[i] = 1e-3 * i; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { b[j] = 1.0 - (lane + j) * 1e-6; bm[j] = 1e-3 * j; }
    for (int it = 0; it < iters; ++it) {
        if ((MODE & 1) && (it & 3) == 0) __syncthreads();
        const int k0 = (it & 3) * 4;
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sm[(k0 + t) * 132 + wm0 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sm[16 * 132 + (k0 + t) * 132 + wn0 + j * 8 + g];
        }
        if (MODE & 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] -= am[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] -= bm[j];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE>
void run(double* out, int sms) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 256>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<sms, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d: %.2f TFLOP/s\n", MODE, (double)sms * 8 * iters * 32 * 512.0 / ms * 1e-9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, p.multiProcessorCount * 256 * 8);
    run<0>(out, p.multiProcessorCount); run<1>(out, p.multiProcessorCount); run<2>(out, p.multiProcessorCount);
    run<3>(out, p.multiProcessorCount); run<7>(out, p.multiProcessorCount);
    return 0;
}
