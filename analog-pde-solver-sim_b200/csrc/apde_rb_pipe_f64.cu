// fp64 instantiation of the sweep-pipelined red-black kernel (see apde_rb_pipe.inl)
#define RBP_T double
#define RBP_FN run_pipe_f64
#include "apde_rb_pipe.inl"
