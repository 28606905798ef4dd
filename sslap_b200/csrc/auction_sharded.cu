// Fourth instance of the persistent auction kernel: the row-sharded multi-GPU solve.  Same source as
// auction.cu with the sharded bidding step and the cross-GPU exchange barrier compiled in; a separate translation unit so
// that the single-GPU instance stays byte-identical (the latency-bound loops are sensitive to code placement).
#define SSLAPB_SHARDED 1
#include "auction.cu"
