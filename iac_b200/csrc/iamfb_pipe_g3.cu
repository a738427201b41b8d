// k_pipe instantiations, group 3 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 3
#include "iamfb_pipe_tu.inc"
