// k_pipe instantiations, group 5, animated-gain variant k_pipe<SIG, true> (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 5
#define IAMFB_PIPE_THIS_RAMPS true
#include "iamfb_pipe_tu.inc"
