// Parameter / result blocks of the batched small-problem kernels (auction.cu: one warp per problem, round 1;
// batch.cu: round 2), shared with the host orchestration (api.cu).
#pragma once

struct SslapbBatchMeta {
    float start_eps, final_eps, target_eps;
    int eCE, soln_found, stop_reason;
    long long its, nreductions, n_assigned;
};

// per object (global column id): price and owner in ONE 16-byte record, so that a bidder's price gather also brings the
// owner it would evict (batch.cu; the reference keeps p[] and object_to_person[] apart, auction_.pyx:220,232)
struct __align__(16) SslapbBatchRec { double price; int owner; int pad; };

struct SslapbBatchParams {
    int P;
    const long long *rowoff, *coloff;     // P+1 prefix sums of the problems' row / column counts
    const long long *rowptr;              // global CSR
    const int *cols;                      // global column ids
    const double *vals;                   // sign-folded
    const float *eps_start;               // per problem, <= 0 => C/2 (nullable)
    long long max_iter;
    double *price; int *owner; unsigned long long *bestkey; int *winpos;      // per global column
    int *p2o, *list, *mover, *bidj; double *bidv, *chosen;                    // per global row
    SslapbBatchMeta *meta;
    SslapbBatchRec *brec;                 // per global column (batch.cu only)
};
