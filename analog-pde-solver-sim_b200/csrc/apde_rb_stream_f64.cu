// fp64 instantiation of the temporally blocked red-black sweep (see apde_rb_stream.inl)
#define RBS_T double
#define RBS_FN run_stream_f64
#define RBS_BAND_FN run_stream_band_f64
#include "apde_rb_stream.inl"
