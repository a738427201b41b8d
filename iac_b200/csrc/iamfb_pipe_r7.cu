// k_pipe instantiations, group 7, animated-gain variant k_pipe<SIG, true> (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 7
#define IAMFB_PIPE_THIS_RAMPS true
#define IAMFB_PIPE_THIS_FMA 1
#include "iamfb_pipe_tu.inc"
