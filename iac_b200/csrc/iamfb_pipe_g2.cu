// k_pipe instantiations, group 2 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 2
#include "iamfb_pipe_tu.inc"
