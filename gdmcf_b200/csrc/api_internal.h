// Internal helpers shared by the translation units of libgdmcf_sm100.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include "../../include/gdmcf_sm100.h"

namespace gd {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
// cudaGetLastError() after a launch; does not synchronise.
int cuda_check_launch(const char* kernel);

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace gd
