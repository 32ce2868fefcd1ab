// fp32 instantiation of the temporally blocked red-black sweep (see apde_rb_stream.inl)
#define RBS_T float
#define RBS_FN run_stream_f32
#define RBS_BAND_FN run_stream_band_f32
#include "apde_rb_stream.inl"
