// k_pipe instantiations, group 4 (see iamfb_pipe_tu.inc)
#define IAMFB_PIPE_THIS_GROUP 4
#include "iamfb_pipe_tu.inc"
