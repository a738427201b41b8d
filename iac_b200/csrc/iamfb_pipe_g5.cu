// k_pipe instantiations, group 5 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 5
#include "iamfb_pipe_tu.inc"
