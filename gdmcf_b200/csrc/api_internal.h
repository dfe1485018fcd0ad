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

// Programmatic dependent launch, OFF unless GDMCF_PDL=1. Measured on B200 (profiles/r2_pdl_sweep.txt): inside the replayed
// CUDA graph it buys nothing on the denoise + rank chain (1.165 vs 1.168 ms) and costs 0.5 ms on the training step, so
// kernel boundaries inside a graph are not where the contractions' fixed cost comes from. When enabled every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with gd::pdl_entry() (common.cuh), i.e.
// griddepcontrol.launch_dependents + griddepcontrol.wait. The next kernel's CTAs are scheduled while this one still runs
// (launch latency, barrier / TMEM / tensormap set-up overlap its tail), and griddepcontrol.wait holds every read and write
// of dependent data until the predecessor grid has completed and flushed. The step is ~160 launches replayed from a CUDA
// graph; stream capture records the attribute as programmatic dependency edges.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in cuda_check_launch()
}

// Kernels whose CTAs wait for one another inside ONE launch (grid barriers of a single resident wave: user_tower_kernel,
// lightgcn_bf16_kernel) are launched cooperatively: the driver either makes every CTA resident at once or fails the launch
// (rc -3) — a partially resident grid can never spin on CTAs that were not scheduled.
template <typename... KArgs, typename... Args>
inline void launch_kernel_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace gd
