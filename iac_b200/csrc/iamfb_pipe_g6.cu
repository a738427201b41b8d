// k_pipe instantiations, group 6 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 6
#include "iamfb_pipe_tu.inc"
