// Single-launch path for small problems (small.cu): arguments and result block shared with the host orchestration (api.cu).
#pragma once
#include <stdint.h>

#define SSLAPB_SMALL_MAXN 256      // persons / objects at most
#define SSLAPB_SMALL_CAP 12288     // CSR entries at most (12 B each in shared memory)

struct SslapbSmallResult {
    int status;                    // 0 solved | 1 fewer values than rows | 2 maximum matching < N | 3 take the general path | 6 empty row
    int cardinality;               // -1 when the feasibility check did not run
    float start_eps, final_eps, target_eps;
    int eCE, soln_found;
    long long its;
    int nreductions, n_assigned;
    double obj64;
    int nnz;
};

struct SslapbSmallArgs {
    const double *mat;             // dense input (N x M, row-major, entries < 0 invalid) when dense != 0
    const int *rows, *cols_in;     // COO input otherwise: element k at rows[k * stride], cols_in[k * stride] (int32)
    const double *val;
    long long stride;
    int nnz, dense;
    int N, M, negate, hk;
    float eps_start;
    long long max_iter;
    int *sol_out;                  // device, N
    double *price_out;             // device, M
    SslapbSmallResult *res;        // device
};
