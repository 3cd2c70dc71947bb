// dev_common.cuh -- shared device/host helpers for the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include "b200_internal.h"

namespace b200 {

extern std::atomic<long long> g_launches;

#define B200_CUDA_OK(expr)                                                      \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            ::b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,     \
                              cudaGetErrorString(_e));                          \
            return B200_ERR_CUDA;                                               \
        }                                                                       \
    } while (0)

#define B200_LAUNCH_CHECK()                                                     \
    do {                                                                        \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);             \
        B200_CUDA_OK(cudaGetLastError());                                       \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: call
// sites remember it per device, so a process that drives several GPUs sets it
// on each of them.
struct AttrOnce {
    unsigned long long done = 0;
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if ((done >> d) & 1ull) return false;
        done |= 1ull << d;
        return true;
    }
};

constexpr int32_t kWorstScore = (int32_t)0xE0000000;
constexpr int32_t kWorstDistI = (int32_t)0x80000000;
constexpr int kShift = B200_SENSCR_SHIFT;

// sphinxbase logmath_add (logmath.c:391-436) on the shift-10 byte table; the
// table pointer may be shared or constant memory.  `zero` = MIN_INT32 >> 12.
__device__ __forceinline__ int32_t logadd_tab(const uint8_t *tab, int32_t x, int32_t y) {
    const int32_t zero = (int32_t)0x80000000 >> (kShift + 2);
    if (x <= zero) return y;
    if (y <= zero) return x;
    int32_t d, r;
    if (x > y) { d = x - y; r = x; } else { d = y - x; r = y; }
    if (d < 0 || d >= 256) return r;
    return r + (int32_t)tab[d];
}

// tied_mgau_common.h:104-121 fast_logmath_add on negated scores.  The reference
// indexes its 256-byte table with |x - y| unchecked; ptm_mgau.c:267-288
// normalises with the MINIMUM of the codebooks' top-1 scores, so normalised
// scores go negative and |x - y| can exceed 255 -- the reference then reads
// whatever follows the table on its heap (its scores become history dependent;
// measured: 8 % of PTM scores differ between a fresh decoder and one that has
// already decoded another utterance).  We return the intended value instead:
// beyond the table the correction term is 0.
__device__ __forceinline__ int32_t fast_logadd_neg(const uint8_t *tab, int32_t mlx, int32_t mly) {
    int32_t d, r;
    if (mlx > mly) { d = mlx - mly; r = mly; } else { d = mly - mlx; r = mlx; }
    return r - (d < 256 ? (int32_t)tab[d] : 0);
}

__device__ __forceinline__ int32_t clamp16(int32_t v) {
    return v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
}

}  // namespace b200
