// Second instance of the persistent auction kernel: the same source with the cooperative sweep of very long bidder rows
// compiled into the multi-bidder regime (SSLAPB_LONG_ROWS, see multi_rounds / coop_multi_phase in auction.cu).  It is a
// separate translation unit so that the instance for CSRs without such rows stays byte-identical — the few-bidder loops
// are latency-bound down to instruction placement.  api.cu picks the instance from the longest row of the CSR.
#define SSLAPB_LONG_ROWS 1
#include "auction.cu"
