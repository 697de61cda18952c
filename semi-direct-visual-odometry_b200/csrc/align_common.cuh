// align_common.cuh -- optimiser state and small helpers shared by the alignment kernels (generic kernel sparse_align.cu, cluster fast path sparse_align_v3.cu)
#pragma once
#include "ctx.h"
#include "math.cuh"

namespace {

using svo::Pose;

struct Ctrl {
    Pose pose, pre_pose;
    double R[9], t[3];
    double E[28];     // last evaluation: H (21, upper triangle row-major), g (6), chi2
    double curE[28];  // accepted evaluation (LM)
    double preChi2;
    double sigma;
    double first_sigma;
    double lambda, nu;
    double dx[6];
    double rmse;
    int n_eval, cur_n, nvis;
    int status, it, done, success;
    int evals_total, iters_total, evals_level, iters_level;
    int first;
};

__constant__ int c_pairA[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ int c_pairB[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

__device__ __forceinline__ void set_Rt(Ctrl* c)
{
    svo::quat_to_R(c->pose.q, c->R);
    c->t[0] = c->pose.t[0];
    c->t[1] = c->pose.t[1];
    c->t[2] = c->pose.t[2];
}

__device__ __forceinline__ void expand_H(const double* E, double* H, double* g)
{
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++, k++) {
            H[a * 6 + b] = E[k];
            H[b * 6 + a] = E[k];
        }
#pragma unroll
    for (int a = 0; a < 6; a++) g[a] = E[21 + a];
}

// dx = (H + diag_add I)^-1 g with H, g packed in E (21 + 6): register-resident fast path, pivoted LDLT fallback
__device__ __noinline__ void solve6(const double* E, double diag_add, double* dx)
{
    if (svo::ldlt6_nopivot(E, diag_add, E + 21, dx)) return;
    double H[36], g[6];
    expand_H(E, H, g);
    for (int i = 0; i < 6; i++) H[i * 6 + i] += diag_add;
    svo::ldlt_solve<6>(H, g, dx);
}

}  // namespace
