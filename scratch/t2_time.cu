#define T2_TIMING 1
#ifndef NBX
#define NBX 8
#define NWX 8
#define RBX 16
#endif
#include "../mcos_b200/csrc/tridiag2.cu"
#include <vector>
#include <cstdlib>
int main(int argc, char** argv) {
    int N = 500, S = argc > 1 ? atoi(argv[1]) : 148;
    std::vector<double> h((size_t)N * N);
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) { double v = (i == j) ? 1.0 : 0.3 * ((rand() % 2001) / 1000.0 - 1.0); h[(size_t)i*N+j] = v; h[(size_t)j*N+i] = v; }
    double *A, *d, *e, *tau; cudaMalloc(&A, (size_t)S*N*N*8); cudaMalloc(&d, (size_t)S*N*8); cudaMalloc(&e, (size_t)S*N*8); cudaMalloc(&tau, (size_t)S*N*8);
    for (int s = 0; s < S; ++s) cudaMemcpy(A + (size_t)s*N*N, h.data(), (size_t)N*N*8, cudaMemcpyHostToDevice);
    size_t smem = t2_smem<NBX, NWX>(N);
    cudaFuncSetAttribute(tridiag2_kernel<NBX, NWX, RBX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tridiag2_kernel<NBX, NWX, RBX><<<S, NWX * 32, smem>>>(A, N, d, e, tau);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("S=%d N=%d: %.3f ms (%s)\n", S, N, ms, cudaGetErrorString(cudaGetLastError()));
}
