// The device-resident Monte-Carlo kernel (mlp_tc_kernel<MODE_FUSED>) as two CTAs of two groups per SM: with programmatic
// dependent launch the next launch's CTA moves into an SM while the other CTA of the previous launch is still at work, so
// the SM is never empty between the back-to-back launches of a Monte-Carlo run.  Same source as k_mlp_tc.cu.
#define NVD_TC_GROUPS 2
#include "k_mlp_tc.cu"
