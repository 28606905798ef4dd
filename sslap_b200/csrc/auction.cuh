// Device-side state of one auction solve (sslap_b200).  Host code lives in api.cu; kernels in auction.cu.
#pragma once
#include "common.cuh"

// Control block in device memory: the only cross-CTA communication channel of the persistent kernel.
struct SslapbCtrl {
    // ---- line 0: the words the waiting CTAs poll (up to 147 CTAs, ~100 loads per microsecond on this one L2 line while
    // CTA 0 runs a phase alone) — nothing the working CTAs touch per round may share it
    unsigned bar_count;       // grid barrier: arrivals
    unsigned bar_gen;         // grid barrier: generation
    int abort_flag;           // set by the watchdog (a barrier waited longer than watchdog_ns) or an internal assert
    int pad0[29];
    // ---- line 1 onwards: round / phase state
    int nu;                   // number of unassigned persons (auction_.pyx:198 num_unassigned)
    int done;                 // 0 running | 1 target-eps CS holds (:275,309) | 2 eps < target (:280) | 3 max_iter (:309)
    float eps;                // current eps, float32 as in the reference (:180)
    float target_eps;         // float32(1/N) (:247)
    float theta;              // 0.15f (:248)
    long long its;            // nits (:273)
    long long max_iter;
    int nreductions;          // :292
    int ece_viol;             // set by the eCE sweep when some person violates eps-CS
    int tie_flag;             // some atomicMax saw an equal bid this round -> run the position tie-break pass
    int ece_final;            // meta['eCE'] (:297); -1 until known
    long long rounds_grid, rounds_warp, rounds_solo;   // instrumentation: rounds executed per regime
    long long rounds_mid;                              // rounds of the mid regime (CTA 0 alone, 33..t_mid bidders, hot phases)
    unsigned long long pmin_key[2];                    // order-preserving image of a LOWER bound of every price (slot = phase & 1);
                                                       // prices never decrease, so a phase-start minimum stays valid all phase
    unsigned long long pmax_key;                       // running maximum of the prices (atomicMax by every winner): heuristic only
    long long prune_second_pass;                       // instrumentation: rows that needed the second (exactness) gather pass
    unsigned long long t_begin, t_end;                 // %globaltimer at kernel start / end
    unsigned dbg[16];                                  // (unused)
                                                       // ((round << 3) | barrier index), reported when the watchdog fires
    unsigned long long prof[8];                        // ns spent (CTA 0 view): 0 grid bid, 1 grid tie+assign, 2 grid compaction,
                                                       // 3 warp regime, 4 solo regime, 5 eCE/phase change, 6 mid regime, 7 barriers of the grid regime
    long long rounds_sharded;                          // row-sharded solve: rounds whose bidding was split over the ranks
    unsigned long long xchg_ns;                        // ... ns CTA 0 spent in their cross-GPU exchange barrier (signal + wait)
    unsigned long long sharded_ns;                     // ... ns of those rounds in total (CTA 0 view)
    unsigned long long sweep_ns[2];                    // in-situ bidding step of full-frontier rounds (nu == N): total ns / count
    double tol;                                        // eps-CS tolerance of eCE_satisfied (1e-7, auction_.pyx:16; 0 in strict mode)
    // ---- hot lists (hot.cu)
    int hot_mode;                                      // this eps-phase decides bids from the hot lists first (set by the probing round)
    int hot_probe_fail;                                // bids of the phase's first round (every person bids) the hot list could not decide
    long long hot_grid[2], hot_tail[2];                // instrumentation: bids decided by the hot list / handed to the full-row sweep
    long long hot_last[2];                             // hot_grid[] at the lead thread's last look (adaptive on/off)
    long long rounds_nohole;                           // grid rounds that ended without compaction and final barrier (no hole, no tie)
};

// One hot-list entry (hot.cu): 32 per person, 512 bytes per row, lane t of a warp reads entry t.
struct __align__(16) SslapbHotEnt {
    int col;                  // object; padding entries: column 0 with a = -inf, idx = -1
    int idx;                  // index of the entry inside its CSR row (tie rule: the LAST maximal entry wins, auction_.pyx:351)
    double a;                 // value (sign-folded)
};

// Per-object record (32 B = one sector): everything a bidder needs to know about the object it wins, so that the
// warp-list regimes need only two dependent memory hops per round (row entries -> records) instead of four
// (owner -> its row offsets -> row entries -> prices).  `price` mirrors price[] (which the grid regime gathers from,
// 8 B per object, L1 friendly); start/deg describe the CSR row of the current owner.
struct __align__(32) SslapbObjRec {
    long long start;          // first CSR entry of the owner's row
    int owner;                // object_to_person (auction_.pyx:232); -1 = unowned
    int deg;                  // number of entries of the owner's row
    double price;             // copy of price[j]
    long long pad;
};

struct SslapbAuctionParams {
    int N, M;
    const long long *rowptr;  // N+1 row start offsets (i_starts_stops, auction_.pyx:223)
    const int *cols;          // flat_j (:229); 16-byte aligned base, >= 4 entries of slack after nnz
    const double *vals;       // sign-folded values ('min' negated, :236-237); same alignment/slack
    const double *rowmax;     // per person: max_j a_ij (static; the pruned sweeps gather only entries within the price
                              // spread of it)
    const SslapbHotEnt *hot;  // hot lists (hot.cu): the 32 largest entries of every row; nullptr = off
    const double *hthr;       // per person: entries with a <= hthr are NOT in the hot list (NaN: every entry is)
    double *rest;             // per person: upper bound of a_ik - p_k over the entries outside the hot list (+inf: unknown)
    double *price;            // p (:220)
    SslapbObjRec *rec;        // per-object record incl. object_to_person (:232)
    int *p2o;                 // person_to_object (:231)
    int *list;                // unassigned_people (:260), positions [0,nu)
    int *mover;               // grid-mode compaction: k-th live entry right of the new count
    int *bidj;                // per list position: object bid on   (objects_bidded, :324)
    double *bidv;             // per list position: bid value       (bids, :325)
    unsigned long long *bidkey;  // per object: order-preserving image of the best bid (best_bids, :255); 0 = none
    int *winpos;              // per object: smallest list position among equal best bids (tie rule of :379)
    int *hole_count;          // per CTA: holes produced in its chunk of positions
    double *chosen;           // per person: sum of (folded) values of entries equal to its object (get_obj, :504-521)
    SslapbCtrl *ctrl;
    int t_small;              // nu <= t_small (<= 32) -> CTA 0 runs the round alone (warp-list regime)
    int t_mid;                // 32 < nu <= t_mid in a hot-list phase -> CTA 0 runs the round alone (mid regime); 0 = off
    int pad_mid;
    unsigned long long watchdog_ns;
    // ---- row-sharded solve over several GPUs (instance of auction_sharded.cu; nranks == 1: off).  Persons are split into
    // nnz-balanced contiguous row ranges; every rank holds the whole state and CSR, but in rounds with nu > t_shard it sweeps
    // only the bidders of its own rows and stores their bids straight into EVERY rank's exchange buffer over NVLink.
    int nranks, rank;
    int t_shard;
    const int *rowsplit;      // nranks + 1 row boundaries (device; written by sslapb_row_split_kernel)
    const unsigned long long *xtab;   // per rank r: [3r] flag block, [3r+1] bidj[2][xcap], [3r+2] bidv[2][xcap] (peer-mapped addresses)
    long long xcap;           // capacity (rows) of one parity half of the exchange buffers
    unsigned xround_base;     // sharded rounds completed by earlier solves on this communicator (flags are monotone)
};

#ifndef SSLAPB_MID
#define SSLAPB_MID 256           // capacity (bidders) of the mid regime's shared-memory list; option "t_mid" is at most this
#endif
#define SSLAPB_MAX_RANKS 8
