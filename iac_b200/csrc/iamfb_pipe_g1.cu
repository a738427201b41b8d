// k_pipe instantiations, group 1 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 1
#include "iamfb_pipe_tu.inc"
