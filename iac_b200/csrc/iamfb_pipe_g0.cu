// k_pipe instantiations, group 0 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 0
#include "iamfb_pipe_tu.inc"
