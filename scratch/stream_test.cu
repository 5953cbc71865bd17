#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
// one CTA per 2 MB matrix; each warp walks 16x32 tiles with 16 independent 8-byte loads per lane, DEPTH tiles in flight
template <int DEPTH>
__global__ void __launch_bounds__(256, 1) stream_kernel(const double* __restrict__ Aall, int N, int reps, double* out, long long* cyc) {
    const double* A = Aall + (size_t)blockIdx.x * N * N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tr = N / 16, tc = N / 32, total = tr * tc;
    double acc = 0.0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = wid; t < total; t += 8 * DEPTH) {
            double av[DEPTH][16];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                int tt = t + 8 * d;
                if (tt < total) {
                    const double* p = A + (size_t)(tt / tc) * 16 * N + (tt % tc) * 32 + lane;
#pragma unroll
                    for (int q = 0; q < 16; ++q) av[d][q] = p[(size_t)q * N];
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) av[d][q] = 0.0;
                }
            }
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int q = 0; q < 16; ++q) acc += av[d][q];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int DEPTH>
void run(const double* A, int N, int S, double* out, long long* cyc) {
    int reps = 8;
    stream_kernel<DEPTH><<<S, 256>>>(A, N, reps, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    stream_kernel<DEPTH><<<S, 256>>>(A, N, reps, out, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)(N / 16) * (N / 32) * 4096.0 * reps;
    printf("S=%4d depth=%d: %.3f ms, block0 %lld cycles, %.1f B/clk/SM, %.1f GB/s per CTA, total %.2f TB/s (%s)\n", S, DEPTH, ms, c, bytes / c,
           bytes / (ms * 1e-3) / 1e9, bytes * S / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int N = 512;
    for (int S : {1, 148, 592}) {
        double* A; cudaMalloc(&A, (size_t)S * N * N * 8); cudaMemset(A, 0, (size_t)S * N * N * 8);
        double* out; cudaMalloc(&out, (size_t)S * 256 * 8); long long* cyc; cudaMalloc(&cyc, S * 8);
        run<1>(A, N, S, out, cyc); run<2>(A, N, S, out, cyc); run<4>(A, N, S, out, cyc);
        cudaFree(A); cudaFree(out); cudaFree(cyc);
    }
}
