#ifdef DBGP
#define MCOSB_DEBUG_MPFIT 1
#endif
#include "../mcos_b200/csrc/denoise.cu"
#include <vector>
int main() {
    FILE* f = fopen("scratch/dbg_lam.bin", "rb");
    double hdr[2]; fread(hdr, 8, 2, f);
    int N = (int)hdr[0], T = (int)hdr[1];
    std::vector<double> lam(N); fread(lam.data(), 8, N, f); fclose(f);
    double* dl; cudaMalloc(&dl, N * 8); cudaMemcpy(dl, lam.data(), N * 8, cudaMemcpyHostToDevice);
    double* dv; int *dn, *ds; cudaMalloc(&dv, 8); cudaMalloc(&dn, 4); cudaMalloc(&ds, 4); cudaMemset(ds, 0, 4);
    mpfit_kernel<<<1, MP_THREADS, N * 8>>>(dl, N, (double)T / N, 0.25, dv, dn, ds);
    cudaDeviceSynchronize();
    double v; int n, s; cudaMemcpy(&v, dv, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&s, ds, 4, cudaMemcpyDeviceToHost);
    printf("var=%.17g nf=%d status=%d err=%s\n", v, n, s, cudaGetErrorString(cudaGetLastError()));
}
