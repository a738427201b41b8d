// k_pipe instantiations, group 7: the IAMFB_ARITH_FMA variants (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 7
#define IAMFB_PIPE_THIS_FMA 1
#include "iamfb_pipe_tu.inc"
