// esrp_host.h — host-side helpers shared by the translation units of libesrp.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace esrp {
// printf-style; stores a thread-local message for esrp_last_error() and returns 1.
int set_error(const char* fmt, ...);
// Encode a 4-D TMA map over an NHWC bf16 tensor; returns 0 on success.
int make_nhwc_tmap(CUtensorMap* tm, const void* ptr, int n, int h, int w, int c_total, int kc,
                   int box_w, int box_h);
int sm_count();
}  // namespace esrp
