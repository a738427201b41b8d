// k_pipe instantiations, group 2, animated-gain variant k_pipe<SIG, true> (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 2
#define IAMFB_PIPE_THIS_RAMPS true
#include "iamfb_pipe_tu.inc"
