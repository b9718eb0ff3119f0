// la_all.cu -- single translation unit of libliteattn_b200.so (no relocatable device code needed;
// the watchdog symbol in la_ptx.cuh is then defined exactly once).
#include <algorithm>
#include "la_fwd_sm100.cu"
#include "la_skip_update.cu"
#include "la_combine.cu"
#include "la_rope_cast.cu"
#include "la_list_codec.cu"
#include "la_api.cu"
