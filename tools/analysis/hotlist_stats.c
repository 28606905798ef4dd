/* Analysis tool (not product, not a test): how deep into a row's value-sorted entries do the best and second-best
 * candidates of a bid sit?  Decides whether a per-row "hot list" of the H largest values (L2-resident, one gather per
 * lane) can serve the few-bidder / chain rounds exactly: the hot-list result is exact when the second-best net value
 * found inside the list exceeds fl(a_(H+1) - pmin), a_(H+1) = the largest value NOT in the list, pmin = a lower bound of
 * all prices (taken at the start of the eps-phase).   gcc -O2 -o hotlist_stats hotlist_stats.c -lm  (see hotlist_stats.py) */
#include <stdio.h>
struct auction_s;
static void stats_bid(const void *s, int nb, int i, double vbest, double wi, int jbest);
#define ORACLE_BID_HOOK(s, nb, i, vbest, wi, jbest) stats_bid((s), (nb), (i), (vbest), (wi), (jbest))
#include "../../oracle/sslap_oracle.c"

#define NH 4
static const int HS[NH] = {8, 16, 24, 32};
static double *g_hmin[NH];
static double g_pmin;
static long long g_nb_hist[40];
static int g_prev_nb = -1; static long long g_grid_rounds, g_grid_same, g_hist_same[8], g_hist_all[8];
static long long g_ph_h[64][NH];
static double *g_restR; static long long g_ph_tail[64], g_ph_pass[64], g_ph_passR[64], g_ph_chain[64], g_ph_chain_pass[64];
static int *g_seen; static int g_phase; static long long g_tail_first;
static double *g_restS[NH], *g_restD[NH];
static long long g_tailS[NH], g_tailD[NH], g_gridS[NH], g_gridD[NH];
static long long g_tail_bids, g_tail_pass[NH], g_grid_bids, g_grid_pass[NH], g_tail_short[NH];
static int cmp_desc(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return x < y ? 1 : (x > y ? -1 : 0); }

static void stats_bid(const void *sv, int nb, int i, double vbest, double wi, int jbest)
{
    const auction_t *s = (const auction_t *)sv;
    (void)vbest; (void)jbest;
    if (!g_hmin[0]) {
        for (int h = 0; h < NH; ++h) g_hmin[h] = (double *)malloc(sizeof(double) * s->N);
        double *tmp = (double *)malloc(sizeof(double) * 1000000);
        for (int r = 0; r < s->N; ++r) {
            const int64_t st = s->rowptr[r], dg = s->rowptr[r + 1] - st;
            memcpy(tmp, s->val + st, sizeof(double) * dg);
            qsort(tmp, dg, sizeof(double), cmp_desc);
            for (int h = 0; h < NH; ++h) g_hmin[h][r] = dg > HS[h] ? tmp[HS[h]] : -INFINITY;
        }
        free(tmp);
    }
    if (nb == s->N && i == s->unassigned[0]) {       /* first bid of an eps-phase: price lower bound */
        g_pmin = INFINITY;
        for (int j = 0; j < s->M; ++j) if (s->p[j] < g_pmin) g_pmin = s->p[j];
    }
    if (nb == s->N && i == s->unassigned[0]) {       /* per-row bounds of everything outside the hot list, at the phase start */
        double *tv = (double *)malloc(sizeof(double) * 1000000), *ta = (double *)malloc(sizeof(double) * 1000000);
        for (int h = 0; h < NH; ++h) {
            if (!g_restS[h]) { g_restS[h] = (double *)malloc(sizeof(double) * s->N); g_restD[h] = (double *)malloc(sizeof(double) * s->N); }
        }
        for (int r = 0; r < s->N; ++r) {
            const int64_t st = s->rowptr[r], dg = s->rowptr[r + 1] - st;
            for (int64_t k = 0; k < dg; ++k) tv[k] = s->val[st + k] - s->p[s->cols[st + k]];
            memcpy(ta, tv, sizeof(double) * dg);
            qsort(ta, dg, sizeof(double), cmp_desc);
            for (int h = 0; h < NH; ++h) {
                g_restD[h][r] = dg > HS[h] ? ta[HS[h]] : -INFINITY;
                double m = -INFINITY;                 /* static list: entries with a <= a_(H+1) are outside */
                for (int64_t k = 0; k < dg; ++k) if (s->val[st + k] <= g_hmin[h][r] && tv[k] > m) m = tv[k];
                g_restS[h][r] = m;
            }
        }
        free(tv); free(ta);
    }
    if (!g_seen) g_seen = (int *)calloc(s->N, sizeof(int));
    if (nb == s->N && i == s->unassigned[0]) ++g_phase;
    if (nb <= 32) { if (g_seen[i] != g_phase) { g_seen[i] = g_phase; ++g_tail_first; } }
    if (nb == s->N && i == s->unassigned[0]) {
        if (!g_restR) g_restR = (double *)malloc(sizeof(double) * s->N);
        memcpy(g_restR, g_restS[3], sizeof(double) * s->N);
    }
    if (nb <= 32) {
        ++g_ph_tail[g_phase];
        for (int h = 0; h < NH; ++h) g_ph_h[g_phase][h] += wi > g_restS[h][i];
        g_ph_pass[g_phase] += wi > g_restS[3][i];
        if (nb == 1) { ++g_ph_chain[g_phase]; g_ph_chain_pass[g_phase] += wi > g_restR[i]; }
        if (wi > g_restR[i]) ++g_ph_passR[g_phase];
        else {                                        /* fallback = full sweep: refresh the bound with the current prices */
            const int64_t st = s->rowptr[i], dg = s->rowptr[i + 1] - st;
            double m = -INFINITY;
            for (int64_t k = 0; k < dg; ++k) if (s->val[st + k] <= g_hmin[3][i]) { double v = s->val[st + k] - s->p[s->cols[st + k]]; if (v > m) m = v; }
            g_restR[i] = m;
        }
    }
    if (i == s->unassigned[0]) {                     /* first bid of a round */
        if (nb <= 32) ++g_nb_hist[nb];
        if (g_prev_nb > 32) {
            int b = g_prev_nb <= 128 ? 0 : g_prev_nb <= 512 ? 1 : g_prev_nb <= 2048 ? 2 : g_prev_nb <= 8192 ? 3 : 4;
            ++g_grid_rounds; ++g_hist_all[b];
            if (nb == g_prev_nb) { ++g_grid_same; ++g_hist_same[b]; }
        }
        g_prev_nb = nb;
    }
    const int tail = nb <= 32;
    for (int h = 0; h < NH; ++h) {
        if (tail) { g_tailS[h] += wi > g_restS[h][i]; g_tailD[h] += wi > g_restD[h][i]; }
        else { g_gridS[h] += wi > g_restS[h][i]; g_gridD[h] += wi > g_restD[h][i]; }
    }
    if (tail) ++g_tail_bids; else ++g_grid_bids;
    for (int h = 0; h < NH; ++h) {
        const double bound = g_hmin[h][i] - g_pmin;
        const int pass = wi > bound;
        if (tail) { g_tail_pass[h] += pass; g_tail_short[h] += (g_hmin[h][i] == -INFINITY); } else g_grid_pass[h] += pass;
    }
}

void stats_report(void)
{
    printf("rounds by frontier size:"); for (int k = 1; k <= 32; ++k) printf(" %d:%lld", k, g_nb_hist[k]); printf("\n");
    printf("grid rounds %lld, of which the frontier size did not change (no hole): %lld\n", g_grid_rounds, g_grid_same);
    for (int b = 0; b < 5; ++b) printf("  frontier bucket %d: rounds %lld no-hole %lld\n", b, g_hist_all[b], g_hist_same[b]);
    for (int p = 1; p <= g_phase; ++p) printf("phase %2d: tail bids %8lld  pass(static32) %.4f  pass(with refresh) %.4f   chain bids %8lld pass(refresh) %.4f\n", p, g_ph_tail[p], (double)g_ph_pass[p] / (g_ph_tail[p] ? g_ph_tail[p] : 1), (double)g_ph_passR[p] / (g_ph_tail[p] ? g_ph_tail[p] : 1), g_ph_chain[p], (double)g_ph_chain_pass[p] / (g_ph_chain[p] ? g_ph_chain[p] : 1));
    for (int p = 1; p <= g_phase; ++p) printf("phase %2d static H=8/16/24/32: %.4f %.4f %.4f %.4f\n", p, (double)g_ph_h[p][0] / (g_ph_tail[p] ? g_ph_tail[p] : 1), (double)g_ph_h[p][1] / (g_ph_tail[p] ? g_ph_tail[p] : 1), (double)g_ph_h[p][2] / (g_ph_tail[p] ? g_ph_tail[p] : 1), (double)g_ph_h[p][3] / (g_ph_tail[p] ? g_ph_tail[p] : 1));
    printf("tail bids whose person bids for the first time in this phase's tail: %lld (phases %d)\n", g_tail_first, g_phase);
    printf("tail bids (rounds with <= 32 bidders): %lld   grid bids: %lld\n", g_tail_bids, g_grid_bids);
    for (int h = 0; h < NH; ++h)
        printf("H=%2d  tail exact-from-hot-list %.4f (rows with deg<=H: %.4f)   grid %.4f\n", HS[h],
               (double)g_tail_pass[h] / (double)(g_tail_bids ? g_tail_bids : 1), (double)g_tail_short[h] / (double)(g_tail_bids ? g_tail_bids : 1),
               (double)g_grid_pass[h] / (double)(g_grid_bids ? g_grid_bids : 1));
    for (int h = 0; h < NH; ++h)
        printf("H=%2d  per-row bound at phase start: static list tail %.4f grid %.4f | dynamic list tail %.4f grid %.4f\n", HS[h],
               (double)g_tailS[h] / (double)g_tail_bids, (double)g_gridS[h] / (double)g_grid_bids,
               (double)g_tailD[h] / (double)g_tail_bids, (double)g_gridD[h] / (double)g_grid_bids);
}
