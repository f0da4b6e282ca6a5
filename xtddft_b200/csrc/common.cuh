// Common definitions for the sigma-vector library (sm_100a only).
#pragma once
#include <cuda.h>          // CUtensorMap and enums only; the driver entry point is resolved at run time
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace xtd {

// ---- error codes returned across the C-ABI (negative = error) --------------------------------------
enum : int {
  XTD_OK = 0,
  XTD_ERR_CUDA = -1,         // a CUDA runtime/driver call failed (message via xtd_last_error)
  XTD_ERR_ARG = -2,          // bad argument / shape / state
  XTD_ERR_ALIGN = -3,        // pointer or leading dimension violates the 16-byte TMA rule
  XTD_ERR_NOMEM = -4,        // workspace too small / allocation failed
  XTD_ERR_STATE = -5,        // call order violated (e.g. sigma before finalize)
  XTD_ERR_UNSUPPORTED = -6
};

extern thread_local char g_last_error[512];

#define XTD_SET_ERR(...) snprintf(::xtd::g_last_error, sizeof(::xtd::g_last_error), __VA_ARGS__)

#define XTD_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      XTD_SET_ERR("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return XTD_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define XTD_TRY(call)            \
  do {                           \
    int r__ = (call);            \
    if (r__ != XTD_OK) return r__; \
  } while (0)

#define XTD_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      XTD_SET_ERR(__VA_ARGS__);      \
      return (code);                 \
    }                                \
  } while (0)

inline long round_up(long x, long m) { return (x + m - 1) / m * m; }
inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// leading dimensions of every internal matrix are multiples of 16 doubles (128 B): satisfies the TMA
// 16-byte stride rule and keeps rows line-aligned.
constexpr int LD_ALIGN = 16;
inline long pad_ld(long n) { return round_up(n < 1 ? 1 : n, LD_ALIGN); }

// launch counter (the bench reports how many of our kernels ran in the timed region)
extern unsigned long long g_launch_count;
#define XTD_COUNT_LAUNCH() (++::xtd::g_launch_count)

}  // namespace xtd
