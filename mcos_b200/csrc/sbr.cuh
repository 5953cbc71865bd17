// Two-stage tridiagonalisation (successive band reduction): shared constants and launchers.
//   stage 1  sy2sb.cu   dense symmetric -> band of half bandwidth SBR_B (blocked Householder, DMMA)
//   stage 2  sb2st.cu   band -> tridiagonal (bulge chasing in shared memory), reflectors kept for the eigenvectors
//   back     q2back_kernel (sb2st.cu) + backtransform_kernel (eigvec3.cu): x = Q1 Q2 y for the signal eigenvectors
#pragma once
#include "common.cuh"

#define SBR_B 8
#ifndef SBR_STAGE1_DEFAULT
#define SBR_STAGE1_DEFAULT(N) 3   // which dense -> band kernel serves size N by default (3 = batched, sy2sb3.cu: measured
                                  // faster than the one-CTA kernels at every N from 128 to 1024, profiles/sy2sb3_r02.md)
#endif
#define SBR_MIN_N 128   // below this the one-stage kernels are used (mcosb_dev_eig_prepare)

// Each returns 1 if launched, 0 if the problem size is outside its range (nothing launched), < 0 on error.
int mcosb_launch_sy2sb(mcosb_ctx* ctx, int S, double* A, double* tau);    // round-1 kernel (N <= 696), MCOSB_SY2SB_OLD=1
int mcosb_launch_sy2sb2(mcosb_ctx* ctx, int S, double* A, double* tau);   // two matrices per SM / N <= 1024 (sy2sb2.cu)
bool sbr2_supported(mcosb_ctx* ctx);
int mcosb_launch_sy2sb3(mcosb_ctx* ctx, int S, double* A, double* tau);   // batched multi-kernel form (sy2sb3.cu), N <= 1024
bool sbr3_supported(mcosb_ctx* ctx);
size_t sbr3_scratch_bytes(int N);
int sbr_stage1_kind(int N);   // 1 one CTA per matrix (sy2sb.cu), 2 sy2sb2.cu, 3 batched (sy2sb3.cu)
int mcosb_launch_stage1(mcosb_ctx* ctx, int S, double* A, double* tau);       // the one in use
int mcosb_launch_sb2st(mcosb_ctx* ctx, int S, const double* A, double* d, double* e, double* Q2);
int mcosb_launch_q2back(mcosb_ctx* ctx, int S, const double* Q2, const int* nf, double* Wy);
bool mcosb_sbr_supported(mcosb_ctx* ctx);
