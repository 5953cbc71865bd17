// FP64 peak micro-benchmarks for the roofline denominators MEASURED_PEAKS.json lacks
// (DFMA vector pipe, DMMA tensor pipe, DADD/DMUL non-fused, FP64 div/exp/log).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_dfma(double* out, int iters, double a0) {
    double a = a0, b = 1.0000001;
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fma(r[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmuladd(double* out, int iters, double a0) {
    double a = a0, b = 1.0000001;
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { double m = __dmul_rn(r[i], b); r[i] = __dadd_rn(m, a); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ddiv(double* out, int iters, double a0) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 1e-3 + i + 1.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = a0 / r[i] + 1.0;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dexp(double* out, int iters, double a0) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = -(threadIdx.x * 1e-3 + i + 0.5);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = exp(r[i]) * a0 - 1.5;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dsqrt(double* out, int iters, double a0) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = (threadIdx.x * 1e-3 + i + 0.5);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = sqrt(r[i]) + a0;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m8n8k4: each warp-level instruction = 8*8*4 = 256 MACs = 512 flops.
__global__ void k_dmma884(double* out, int iters) {
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-9;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m16n8k8: 16*8*8 = 1024 MACs = 2048 flops per warp instruction. A: 4 regs, B: 2 regs, C: 4 regs.
__global__ void k_dmma1688(double* out, int iters) {
    double a0 = threadIdx.x * 1e-6, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double b0 = 1.0 + threadIdx.x * 1e-9, b1 = b0 * 0.5;
    double c[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m16n8k16: 2048 MACs. A: 8 regs, B: 4 regs, C: 4 regs.
__global__ void k_dmma16816(double* out, int iters) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-9 * (i + 1);
    double c[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double4* __restrict__ in, double4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"smem_optin\": %zu, \"l2\": %d,\n", p.name, sms, p.sharedMemPerBlockOptin, p.l2CacheSize);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 4096;
    for (int bps = 1; bps <= 2; ++bps) {
        for (int threads = 256; threads <= 1024; threads *= 2) {
            int grid = sms * bps;
            double n_thr = (double)grid * threads;
            float ms;
            ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, iters, 0.5); }, 5);
            printf(" \"dfma_tflops_b%d_t%d\": %.2f,\n", bps, threads, n_thr * iters * 16 * 2 / ms * 1e-9);
            ms = time_ms([&] { k_dmma884<<<grid, threads>>>(out, iters); }, 5);
            printf(" \"dmma884_tflops_b%d_t%d\": %.2f,\n", bps, threads, (n_thr / 32) * iters * 8 * 512.0 / ms * 1e-9);
            ms = time_ms([&] { k_dmma1688<<<grid, threads>>>(out, iters); }, 5);
            printf(" \"dmma1688_tflops_b%d_t%d\": %.2f,\n", bps, threads, (n_thr / 32) * iters * 6 * 2048.0 / ms * 1e-9);
            ms = time_ms([&] { k_dmma16816<<<grid, threads>>>(out, iters); }, 5);
            printf(" \"dmma16816_tflops_b%d_t%d\": %.2f,\n", bps, threads, (n_thr / 32) * iters * 6 * 4096.0 / ms * 1e-9);
        }
    }
    {
        int grid = sms * 2, threads = 512; double n_thr = (double)grid * threads; float ms;
        ms = time_ms([&] { k_dmuladd<<<grid, threads>>>(out, iters, 0.5); }, 5);
        printf(" \"dmul_dadd_tinstr\": %.2f,\n", n_thr * iters * 16 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_ddiv<<<grid, threads>>>(out, 512, 0.5); }, 5);
        printf(" \"ddiv_tops\": %.3f,\n", n_thr * 512 * 8 / ms * 1e-9);
        ms = time_ms([&] { k_dexp<<<grid, threads>>>(out, 512, 0.5); }, 5);
        printf(" \"dexp_tops\": %.3f,\n", n_thr * 512 * 8 / ms * 1e-9);
        ms = time_ms([&] { k_dsqrt<<<grid, threads>>>(out, 512, 0.5); }, 5);
        printf(" \"dsqrt_tops\": %.3f,\n", n_thr * 512 * 8 / ms * 1e-9);
    }
    {
        size_t bytes = (size_t)2 << 30;  // 2 GiB each way
        double4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        CK(cudaMemset(a, 1, bytes));
        float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, bytes / sizeof(double4)); }, 5);
        printf(" \"copy_gbs\": %.1f,\n", 2.0 * bytes / ms * 1e-6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); }, 5);
        printf(" \"memcpy_d2d_gbs\": %.1f,\n", 2.0 * bytes / ms * 1e-6);
        // pinned host <-> device
        void* h; CK(cudaMallocHost(&h, 256 << 20));
        ms = time_ms([&] { CK(cudaMemcpyAsync(a, h, 256 << 20, cudaMemcpyHostToDevice)); }, 3);
        printf(" \"h2d_gbs\": %.1f,\n", (256 << 20) / ms * 1e-6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(h, a, 256 << 20, cudaMemcpyDeviceToHost)); }, 3);
        printf(" \"d2h_gbs\": %.1f,\n", (256 << 20) / ms * 1e-6);
    }
    printf(" \"clock_khz\": %d}\n", p.clockRate);
    return 0;
}
