// Flag blocks shared between the build / Hopcroft-Karp kernels and the host orchestration (api.cu).
#pragma once

struct SslapbBuildFlags {
    int unsorted;                 // rows not non-decreasing (the reference's precondition, auction_.pyx:33-48)
    int out_of_range;             // index outside [0,N) x [0,M)
    int empty_rows;               // some row in [0,N) has no entry (infeasible; UB in the reference)
    int maxdeg;                   // longest row (written by the row-maximum pass)
    unsigned long long maxabs;    // bits of max |a_ij| (non-negative doubles order like uint64), max_val :123-134
    long long nnz;                // dense path: number of valid (>= 0) entries
};

struct SslapbHkFlags {
    int found;        // BFS reached a free right vertex at this level
    int grew;         // BFS labelled at least one new left vertex
    int augmented;    // successful augmentations in this phase
    int matched;      // greedy initial matches
};

// Control block of the device-resident Hopcroft-Karp loop (hopcroft.cu: sslapb_hk_persistent_kernel)
struct SslapbHkCtrl {
    unsigned bar;             // grid barrier: monotone arrival counter
    int found;                // 0, or 1 + the BFS level at which a tree reached a free right vertex in this phase
    int cnt[3];               // frontier sizes, rotating by level (cnt[L % 3] = size of level L)
    int nroots;               // free left vertices at the start of the phase
    int augmented;            // (unused)
    int matched;              // size of the matching
    int phases, levels;       // instrumentation: phases run, BFS levels in total
    int watchdog;             // 1: a barrier wait exceeded the limit
};
