// DMMA.8x8x4 dependent-issue latency and throughput vs independent chains per warp / warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, long long* cyc) {
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-9;
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = i; c[i][1] = -i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// shuffle throughput (64-bit shuffles)
__global__ void ks(double* out, int iters, long long* cyc) {
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], (threadIdx.x * 5 + i) & 31);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int threads, double* out, long long* cyc) {
    const int iters = 4096;
    k<CH><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<CH><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * CH);
    printf("chains %d warps/SM %2d: %.2f cycles per DMMA per warp -> %.2f cycles per DMMA per SM\n", CH, threads / 32, per,
           per / (threads / 32));
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int th : {32, 128, 256, 512, 1024}) { run<1>(th, out, cyc); run<2>(th, out, cyc); run<4>(th, out, cyc); run<8>(th, out, cyc); }
    for (int th : {32, 128, 512}) {
        ks<<<148, th>>>(out, 4096, cyc); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl64: warps/SM %d: %.2f cycles per 64-bit shuffle per warp -> %.2f per SM\n", th / 32, (double)h / (4096 * 8),
               (double)h / (4096 * 8) / (th / 32));
    }
    return 0;
}
