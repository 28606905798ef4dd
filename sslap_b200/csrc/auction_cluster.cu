// Third instance of the persistent auction kernel (opt-in, sslapb_set_option "t_cluster" > 0): the same source with the
// cluster regime compiled in — for frontiers of t_small < nu <= t_cluster bidders the CTAs of cluster 0 run the rounds
// alone with hardware cluster barriers (spread_round<true> in auction.cu).  A separate translation unit for the same
// reason as auction_long.cu: the default instance stays byte-identical.
#define SSLAPB_CLUSTER_REGIME 1
#include "auction.cu"
