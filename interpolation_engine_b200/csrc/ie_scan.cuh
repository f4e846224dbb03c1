// CTA-level exclusive scan of per-thread output lengths, and the two ways a tile finds its place in a compacted
// arena without a second pass over HBM: a decoupled look-back across ordered tiles (Merrill & Garland style; the escape
// kernel, whose output must stay in input order) or one atomic claim per tile (the resolve kernel: results carry their
// own offsets, so tile order in the arena is free).
#pragma once
#include <cstdint>

#include "ie_kernels.h"

namespace ie_scan {

constexpr uint64_t FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_state(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int NT>
struct TileSmemT {
    uint32_t tile;
    uint64_t warp_tot[NT / 32];
    uint64_t base;
};

// Tile ids are handed out in CTA start order, so every predecessor of a tile is already resident
// or finished when it waits on it: forward progress does not depend on the block scheduler.
template <int NT>
__device__ __forceinline__ uint32_t acquire_tile(TileSmemT<NT>& sm, uint32_t* tile_counter) {
    if (threadIdx.x == 0) sm.tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    return sm.tile;
}

// For kernels that do not order their tiles: the thread's exclusive prefix
// within the tile, the tile total rounded up to round_mask + 1 and (optionally) the exact total.
template <int NT>
__device__ __forceinline__ uint64_t local_scan(TileSmemT<NT>& sm, uint64_t len, uint64_t round_mask, uint64_t* tile_total_rounded,
                                               uint64_t* tile_total_raw = nullptr) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += y;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    uint64_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const uint64_t x = sm.warp_tot[w];
        if (w < (int)warp) warp_off += x;
        tile_total += x;
    }
    if (tile_total_raw) *tile_total_raw = tile_total;
    *tile_total_rounded = (tile_total + round_mask) & ~round_mask;
    return warp_off + incl - len;
}
// Claims `bytes` of the output arena for this tile (completion order) and returns the range's start on
// every thread.  Must be called by all NT threads.
template <int NT>
__device__ __forceinline__ uint64_t allocate(TileSmemT<NT>& sm, uint64_t* counter, uint64_t bytes) {
    if (threadIdx.x == 0) sm.base = atomicAdd(reinterpret_cast<unsigned long long*>(counter), (unsigned long long)bytes);
    __syncthreads();
    return sm.base;
}

// Wide look-back: every lane inspects LB consecutive predecessors per round (independent loads in flight),
// so a round trip to L2 covers 32 * LB tiles.  Used by kernels whose tiles all advance at the same pace
// (many tiles sit between "aggregate published" and "inclusive prefix published" at any time).
#ifdef IE_PHASE_TIMING
__device__ unsigned long long g_lb_stats[4];  // look-backs, rounds, spin iterations, cycles spent spinning
#define LB_STAT(k, v) do { if (lane == 0) atomicAdd(&g_lb_stats[k], (unsigned long long)(v)); } while (0)
#else
#define LB_STAT(k, v) do { } while (0)
#endif
template <int NT, int LB>
__device__ __forceinline__ uint64_t lookback_wide(TileSmemT<NT>& sm, uint64_t* tile_state, uint32_t tile, uint64_t tile_total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        uint64_t base = 0;
        if (tile != 0) {
            int64_t j = (int64_t)tile - 1;  // nearest predecessor not yet accounted for
            LB_STAT(0, 1);
            // Tiles publish roughly in id order: one lane waits for the nearest predecessor before the warp
            // looks at the window, instead of 32 * LB lanes polling the same few cache lines of fresh states
            // (the polls of all resident CTAs queue up in front of the very stores they are waiting for).
            if (lane == 0) {
                while ((ld_state(tile_state + j) >> 62) == 0) { LB_STAT(3, 1); __nanosleep(100); }
            }
            __syncwarp();
            for (;;) {
                LB_STAT(1, 1);
                // lane l owns predecessors j - l*LB, j - l*LB - 1, ... (nearest first)
                uint64_t sv[LB];
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    const int64_t idx = j - (int64_t)lane * LB - k;
                    sv[k] = idx >= 0 ? ld_state(tile_state + idx) : FLAG_INC;  // before tile 0: inclusive prefix 0
                }
                uint64_t sum = 0;
                bool inc = false;
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    const int64_t idx = j - (int64_t)lane * LB - k;
                    while ((sv[k] >> 62) == 0) {
#ifdef IE_PHASE_TIMING
                        atomicAdd(&g_lb_stats[2], 1ull);
#endif
                        __nanosleep(20); sv[k] = ld_state(tile_state + idx);
                    }
                    if (!inc) { sum += sv[k] & VAL_MASK; inc = (sv[k] >> 62) == 2; }
                }
                const uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, inc);
                const int first = inc_mask ? (__ffs(inc_mask) - 1) : 31;
                uint64_t v = ((int)lane <= first) ? sum : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
                base += v;
                if (inc_mask) break;
                j -= 32 * LB;
            }
            if (lane == 0) st_state(tile_state + tile, FLAG_INC | (base + tile_total));
        }
        if (lane == 0) sm.base = base;
    }
    __syncthreads();
    return sm.base;
}

}  // namespace ie_scan
