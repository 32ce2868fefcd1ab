// fp32 instantiation of the sweep-pipelined red-black kernel (see apde_rb_pipe.inl)
#define RBP_T float
#define RBP_FN run_pipe_f32
#include "apde_rb_pipe.inl"
