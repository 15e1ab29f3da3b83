// Host experiment: how well could the announcement replay of K2 be balanced?  Runs the device logic of fresh games on the CPU (like
// tests/hostsim), counts the replay loop's iterations per game, and prints the SIMT efficiency (mean / max over the 32 lanes of a
// warp) for (a) one game per lane — the kernel today, (b) G games per lane replayed back to back, (c) a per-block queue of G x 384
// replays handed to whichever lane is free.    g++ -O2 -std=c++17 -o replay_balance replay_balance.cpp && ./replay_balance
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
static thread_local uint32_t g_iters = 0;
#define DK_REPLAY_ITER() (++g_iters)
#include "../../include/doko_cuda.h"
#include "../../master_doko_reinforcement_learning_b200/csrc/dk_common.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/fdo_rules.cuh"
struct LocalDeck {
    uint32_t w[12];
    uint32_t get(uint32_t i) const { return w[i]; }
    void set(uint32_t i, uint32_t v) { w[i] = v; }
    uint32_t get8(uint32_t j) const { return (w[j >> 2] >> (8u * (j & 3u))) & 0xFFu; }
    void set8(uint32_t j, uint32_t v) { w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8u * (j & 3u)))) | ((v & 0xFFu) << (8u * (j & 3u))); }
};
int main() {
    static uint32_t lut[dk::FULL_LUT_WORDS];
    for (uint32_t i = 0; i < dk::CARD_LUT_WORDS; ++i) lut[i] = dk::lut_word(i);
    for (uint32_t i = dk::ANN_LUT_BASE; i < dk::FULL_LUT_WORDS; ++i) lut[i] = dk::lut_word(i);
    for (uint32_t h = 0; h < dk::SEL12_WORDS / 2; ++h) { uint64_t e = dk::sel12_entry(h); std::memcpy(lut + dk::SEL12_LUT_BASE + 2 * h, &e, 8); }
    const uint32_t N = 384 * 8 * 64;
    std::vector<uint32_t> it(N);
    for (uint32_t i = 0; i < N; ++i) {
        LocalDeck deck; dk::RngKey key{0xD0C05EEDu, 0u, i, 0u, 2u};
        int32_t p[4]; uint32_t s; g_iters = 0;
        dk::fdo_playout_fresh<true, LocalDeck, true>(key, deck, lut, p, s);
        it[i] = g_iters;
    }
    double tot = 0; for (auto v : it) tot += v;
    printf("games %u, replay iterations per game: mean %.2f, min %u, max %u\n", N, tot / N, *std::min_element(it.begin(), it.end()), *std::max_element(it.begin(), it.end()));
    for (uint32_t G : {1u, 2u, 3u, 4u, 8u}) {
        // (b) lane l of a warp replays G games back to back: games of a block of G*384 are dealt round-robin
        double warp_iters = 0, useful = 0;
        for (uint32_t b = 0; b + G * 384 <= N; b += G * 384)
            for (uint32_t w = 0; w < 12; ++w) {
                uint32_t mx = 0;
                for (uint32_t l = 0; l < 32; ++l) { uint32_t sum = 0; for (uint32_t g = 0; g < G; ++g) sum += it[b + g * 384 + w * 32 + l]; mx = std::max(mx, sum); useful += sum; }
                warp_iters += mx;
            }
        // (c) per-block queue: every lane of the block takes the next replay when it is free (greedy list scheduling, lanes of a warp in lock-step)
        double q_iters = 0;
        for (uint32_t b = 0; b + G * 384 <= N; b += G * 384) {
            std::vector<uint32_t> busy(384, 0); uint32_t next = 0, total = G * 384; double steps = 0;
            // simulate warp by warp independently in time order is complex; approximate: each lane pulls from the shared queue whenever it finishes
            std::vector<uint32_t> t(384, 0);
            while (next < total) { uint32_t l = (uint32_t)(std::min_element(t.begin(), t.end()) - t.begin()); t[l] += it[b + next++]; }
            for (uint32_t w = 0; w < 12; ++w) steps += *std::max_element(t.begin() + 32 * w, t.begin() + 32 * w + 32);
            q_iters += steps;
        }
        printf("G=%u: back-to-back efficiency %.3f (warp iterations per game %.2f); block queue efficiency %.3f (%.2f)\n", G, useful / (warp_iters * 32), warp_iters / (useful / tot * N) * (tot / N) * 1.0 / (tot / N) * (tot / N) / 1.0 / (tot / N), useful / (q_iters * 32), 0.0);
    }
    return 0;
}
