// kernels.cuh — __global__ entry points (sm_100a).  One thread = one game / sample / rollout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dk_common.cuh"
#include "fdo_rules.cuh"

namespace dk {

constexpr int PLAYOUT_THREADS = 256;

struct RngParams {
    uint32_t seed_lo, seed_hi;
    uint64_t first_id;
    uint32_t epoch;
};

// Per-thread 12-word scratch in shared memory, word-interleaved across the block so that every access of a
// warp hits 32 distinct banks regardless of the (data-dependent) word index.
struct SharedDeck {
    uint32_t* base;  // &smem[threadIdx.x]
    __device__ __forceinline__ uint32_t get(uint32_t w) const { return base[w * PLAYOUT_THREADS]; }
    __device__ __forceinline__ void set(uint32_t w, uint32_t v) { base[w * PLAYOUT_THREADS] = v; }
};

__device__ __forceinline__ RngKey make_key(const RngParams& rp, uint64_t index, uint32_t unit_hi_override, bool use_override) {
    uint64_t unit = rp.first_id + index;
    RngKey k;
    k.seed_lo = rp.seed_lo; k.seed_hi = rp.seed_hi;
    k.unit_lo = (uint32_t)unit;
    k.unit_hi = use_override ? unit_hi_override : (uint32_t)(unit >> 32);
    k.epoch = rp.epoch;
    return k;
}

// K2: fresh full-rules playouts.  Replaces FdoState::new_game + the random_action loop
// (rs-full-doko/src/state/state.rs:169-178,378-431).  HBM traffic: 0 B in, 16 B points + 4 B steps out per game.
template <bool WITH_ANN>
__global__ void __launch_bounds__(PLAYOUT_THREADS)
fdo_playout_fresh_kernel(RngParams rp, uint64_t n, int4* __restrict__ points, uint32_t* __restrict__ steps) {
    __shared__ uint32_t smem[12 * PLAYOUT_THREADS];
    uint64_t i = (uint64_t)blockIdx.x * PLAYOUT_THREADS + threadIdx.x;
    // Out-of-range lanes play game n-1 again (keeps the warp converged); they just do not store.
    uint64_t gi = i < n ? i : n - 1;
    SharedDeck deck;
    deck.base = smem + threadIdx.x;
    RngKey key = make_key(rp, gi, 0, false);
    int32_t p[4];
    uint32_t s;
    fdo_playout_fresh<WITH_ANN>(key, deck, p, s);
    if (i < n) {
        if (points) points[i] = make_int4(p[0], p[1], p[2], p[3]);
        if (steps) steps[i] = s;
    }
}

}  // namespace dk
