// hostsim — TEST INFRASTRUCTURE ONLY.
// Compiles the per-thread device logic of master_doko_reinforcement_learning_b200/csrc/*.cuh with g++ and runs it
// on the CPU, one "thread" at a time, so that kernel logic can be parity-checked against the oracle in the build
// container (which has no GPU).  It is NOT a CPU fallback: the product library never contains or calls this file,
// and the GPU parity tests (-m gpu) exercise the real kernels through the C ABI.
#include <cstdint>
#include <cstring>

#include "../../include/doko_cuda.h"
#include "../../master_doko_reinforcement_learning_b200/csrc/dk_common.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/fdo_rules.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/doko_rules.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/state_ops.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/encode.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/matching.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/assignment.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/pimc.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/selfplay.cuh"
#include "../../master_doko_reinforcement_learning_b200/csrc/uct.cuh"
#include <cmath>
#include <vector>

#define SIM_API extern "C" __attribute__((visibility("default")))

namespace {
struct LocalDeck {
    uint32_t w[12];
    uint32_t get(uint32_t i) const { return w[i]; }
    void set(uint32_t i, uint32_t v) { w[i] = v; }
    uint32_t get8(uint32_t j) const { return (w[j >> 2] >> (8u * (j & 3u))) & 0xFFu; }
    void set8(uint32_t j, uint32_t v) { w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8u * (j & 3u)))) | ((v & 0xFFu) << (8u * (j & 3u))); }
};
const uint32_t* card_lut() {      // the shared-memory image of the fresh-game playout kernel: small tables + the 12-bit rank-select table
    alignas(16) static uint32_t lut[dk::FULL_LUT_WORDS];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < dk::CARD_LUT_WORDS; ++i) lut[i] = dk::lut_word(i);
        for (uint32_t i = dk::ANN_LUT_BASE; i < dk::FULL_LUT_WORDS; ++i) lut[i] = dk::lut_word(i);
        for (uint32_t h = 0; h < dk::SEL12_WORDS / 2; ++h) { uint64_t e = dk::sel12_entry(h); std::memcpy(lut + dk::SEL12_LUT_BASE + 2 * h, &e, 8); }
        init = true;
    }
    return lut;
}
dk::RngKey make_key(uint64_t seed, uint64_t unit, uint32_t epoch) {
    dk::RngKey k; k.seed_lo = (uint32_t)seed; k.seed_hi = (uint32_t)(seed >> 32); k.unit_lo = (uint32_t)unit; k.unit_hi = (uint32_t)(unit >> 32); k.epoch = epoch;
    return k;
}
}  // namespace

SIM_API void sim_fdo_playout_fresh(uint64_t seed, uint32_t epoch, uint64_t first_id, uint64_t n, int with_ann, int32_t* points, uint32_t* steps) {
    for (uint64_t i = 0; i < n; ++i) {
        LocalDeck deck;
        dk::RngKey key = make_key(seed, first_id + i, epoch);
        int32_t p[4]; uint32_t s;
        // the fresh-game kernel's instantiation: rank select through the 12-bit table
        if (with_ann) dk::fdo_playout_fresh<true, LocalDeck, true>(key, deck, card_lut(), p, s); else dk::fdo_playout_fresh<false, LocalDeck, true>(key, deck, card_lut(), p, s);
        for (int q = 0; q < 4; ++q) points[i * 4 + q] = p[q];
        steps[i] = s;
    }
}
SIM_API void sim_doko_playout_fresh(uint64_t seed, uint32_t epoch, uint64_t first_id, uint64_t n, int32_t* points, uint32_t* steps, uint8_t* trace, uint32_t* aux) {
    for (uint64_t i = 0; i < n; ++i) {
        LocalDeck deck;
        dk::RngKey key = make_key(seed, first_id + i, epoch);
        int32_t p[4]; uint32_t s; uint8_t tr[52]; uint32_t ax[4];
        dk::doko_playout_fresh<true, LocalDeck, true>(key, deck, card_lut(), p, s, tr, ax);
        for (int q = 0; q < 4; ++q) points[i * 4 + q] = p[q];
        steps[i] = s;
        if (trace) std::memcpy(trace + i * 52, tr, 52);
        if (aux) std::memcpy(aux + i * 4, ax, 16);
    }
}
// the byte-parallel eligibility test of the announcement replay: cards[4] on hand, thresholds of the two teams, re seats → nibble
SIM_API uint32_t sim_fdo_eligible_nibble(const uint32_t cards[4], uint32_t thr_re, uint32_t thr_ko, uint32_t re) {
    uint32_t c4 = 0x80808080u;
    for (int s = 0; s < 4; ++s) c4 += cards[s] << (8 * s);
    const uint32_t rs = dk::fdo_spread4(re & 15u);
    return dk::fdo_eligible_nibble(c4, thr_re * rs + thr_ko * (0x01010101u - rs));
}
SIM_API uint32_t sim_fdo_allowed_call(uint32_t c, uint32_t m, uint32_t e, uint32_t w) { return dk::fdo_allowed_call(c, m, e, w); }
// The case-by-case closed form of FdoEndOfGameStats::calculate (stats/stats.rs:46-240 and callees) that dk::fdo_score was first written as
// and checked against the oracle in: kept here as the readable specification of the straight-line form the kernels run
// (tests/test_oracle_rule_tables.py compares the two over every input).
static int32_t fdo_score_spec(uint32_t re_eyes, uint32_t re_tricks, uint32_t n_re_players, uint32_t rl, uint32_t kl, int32_t extras,
                        int32_t* kontra_points) {
    uint32_t ko_eyes = 240u - re_eyes;
    bool re_all = re_tricks == 12u, ko_all = re_tricks == 0u;
    uint32_t R = rl == 6u ? 1u : rl, K = kl == 6u ? 1u : kl;   // Counter counts as "Re/Kontra said"
    bool re_won, ko_won;
    if (R >= 2u) re_won = R == 5u ? re_all : re_eyes >= 121u + 30u * (R - 1u);          // 151 / 181 / 211 / all tricks
    else if (K >= 2u) re_won = K == 5u ? !ko_all : re_eyes >= 150u - 30u * K;            // 90 / 60 / 30 / one trick
    else re_won = re_eyes >= ((R == 0u && K == 1u) ? 120u : 121u);
    if (K >= 2u) ko_won = K == 5u ? ko_all : ko_eyes >= 121u + 30u * (K - 1u);
    else if (R >= 2u) ko_won = R == 5u ? !re_all : ko_eyes >= 150u - 30u * R;
    else ko_won = ko_eyes >= ((R == 0u && K == 1u) ? 121u : 120u);
    bool solo = n_re_players == 1u;
    int32_t re_pts;
    // (e)/(f): points for reaching 120/90/60/30 against the other side's No90/No60/No30/Black
    int32_t re_reached = (int32_t)((re_eyes >= 120u && K >= 2u) + (re_eyes >= 90u && K >= 3u) + (re_eyes >= 60u && K >= 4u) + (re_eyes >= 30u && K >= 5u));
    int32_t ko_reached = (int32_t)((ko_eyes >= 120u && R >= 2u) + (ko_eyes >= 90u && R >= 3u) + (ko_eyes >= 60u && R >= 4u) + (ko_eyes >= 30u && R >= 5u));
    if (!re_won && !ko_won) {                            // draw: stats.rs:120-147
        int32_t re_b = -(int32_t)((re_eyes < 90u) + (re_eyes < 60u) + (re_eyes < 30u)) +
                       (int32_t)((ko_eyes < 90u) + (ko_eyes < 60u) + (ko_eyes < 30u)) + re_reached - ko_reached;
        re_pts = re_b + (solo ? 0 : extras);
    } else {
        uint32_t loser = re_won ? ko_eyes : re_eyes;
        bool winner_all = re_won ? re_all : ko_all;
        int32_t w = 1 + (int32_t)((loser < 90u) + (loser < 60u) + (loser < 30u)) + (winner_all ? 1 : 0) + (R >= 1u ? 2 : 0) +
                    (K >= 1u ? 2 : 0) + (R >= 2u ? (int32_t)R - 1 : 0) + (K >= 2u ? (int32_t)K - 1 : 0) + re_reached + ko_reached;
        int32_t re_b = re_won ? w : -w;
        int32_t x = solo ? 0 : extras - (re_won ? 0 : 1);  // "against the club queens" when Kontra wins (:77-82)
        re_pts = re_b + x;
    }
    *kontra_points = -re_pts;
    return solo ? 3 * re_pts : re_pts;                   // stats.rs:215-218
}
SIM_API int32_t sim_fdo_score_spec(uint32_t re_eyes, uint32_t re_tricks, uint32_t n_re, uint32_t rl, uint32_t kl, int32_t extras, int32_t* ko) {
    return fdo_score_spec(re_eyes, re_tricks, n_re, rl, kl, extras, ko);
}
// number of inputs on which the straight-line form and the specification differ (all eyes x tricks x team sizes x calls x extras)
SIM_API uint64_t sim_fdo_score_mismatches() {
    uint64_t bad = 0;
    for (uint32_t e = 0; e <= 240u; ++e) for (uint32_t t = 0; t <= 12u; ++t) for (uint32_t n = 1; n <= 3u; ++n)
        for (uint32_t rl = 0; rl <= 6u; ++rl) for (uint32_t kl = 0; kl <= 6u; ++kl) for (int32_t x = -12; x <= 12; ++x) {
            int32_t k1 = 0, k2 = 0;
            const int32_t r1 = dk::fdo_score(e, t, n, rl, kl, x, &k1), r2 = fdo_score_spec(e, t, n, rl, kl, x, &k2);
            bad += (r1 != r2) || (k1 != k2);
        }
    return bad;
}
// The per-seat loop dk::fdo_final_points was first written as (team sums seat by seat): the specification of its packed-sum form.
static void fdo_final_points_spec(const dk::FdoLive& g, int32_t pts[4]) {
    uint32_t re_eyes = 0, re_tricks = 0;
    int32_t extras = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        bool re = (g.re_mask >> s) & 1u;
        uint32_t e = (g.eyes >> (8u * s)) & 255u, n = (g.ntricks >> (4u * s)) & 15u, d = (g.dkc >> (4u * s)) & 15u;
        if (re) { re_eyes += e; re_tricks += n; extras += (int32_t)d; } else { extras -= (int32_t)d; }
    }
#pragma unroll
    for (uint32_t f = 0; f < 2; ++f) {                                // caught foxes: ♦A played by the other team than the trick's winner
        uint32_t rec = (g.foxes >> (8u * f)) & 255u, players = rec >> 2;
        bool won_re = (g.re_mask >> (rec & 3u)) & 1u;
        int32_t caught = (int32_t)dk::popc(players & (won_re ? ~g.re_mask : g.re_mask) & 15u);
        extras += won_re ? caught : -caught;
    }
    if (g.karl) extras += ((g.re_mask >> g.last_winner) & 1u) ? 1 : -1;
    int32_t ko;
    int32_t re = fdo_score_spec(re_eyes, re_tricks, dk::popc(g.re_mask), g.re_low, g.ko_low, extras, &ko);
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) pts[s] = ((g.re_mask >> s) & 1u) ? re : ko;
}
// random tracker sets (12 tricks / 240 eyes dealt to the seats, any team split, calls, fox records, Karlchen) on which the two differ
SIM_API uint64_t sim_fdo_final_points_mismatches(uint64_t n) {
    uint64_t x = 88172645463325252ull, bad = 0;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 11); };
    for (uint64_t it = 0; it < n; ++it) {
        dk::FdoLive g; dk::fdo_live_clear(g);
        uint32_t e[4] = {0, 0, 0, 0}, nt[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0}, left = 240;
        for (int t = 0; t < 12; ++t) {
            const uint32_t w = rnd() & 3u;
            uint32_t ey = t == 11 ? left : rnd() % (left + 1u < 61u ? left + 1u : 61u);
            left -= ey; e[w] += ey; nt[w]++; if (ey >= 40u) d[w]++;
        }
        for (int s = 0; s < 4; ++s) { g.eyes |= e[s] << (8 * s); g.ntricks |= nt[s] << (4 * s); g.dkc |= d[s] << (4 * s); }
        do { g.re_mask = rnd() & 15u; } while (g.re_mask == 0u || g.re_mask == 15u);
        g.re_low = rnd() % 7u; g.ko_low = rnd() % 7u; g.karl = rnd() & 1u; g.last_winner = rnd() & 3u;
        g.foxes = (rnd() & 1u) ? (rnd() & 0xFFFFu) : ((rnd() & 1u) ? (rnd() & 0xFFu) : 0u);
        int32_t a[4], b[4];
        dk::fdo_final_points(g, a); fdo_final_points_spec(g, b);
        bad += a[0] != b[0] || a[1] != b[1] || a[2] != b[2] || a[3] != b[3];
    }
    return bad;
}
SIM_API int32_t sim_fdo_score(uint32_t re_eyes, uint32_t re_tricks, uint32_t n_re, uint32_t rl, uint32_t kl, int32_t extras, int32_t* ko) {
    return dk::fdo_score(re_eyes, re_tricks, n_re, rl, kl, extras, ko);
}
SIM_API uint32_t sim_select_lsb24(uint32_t x, uint32_t k) { return dk::select_lsb24(x, k); }
SIM_API uint32_t sim_select_lsb(uint32_t x, uint32_t k) { return dk::select_lsb(x, k); }
SIM_API uint32_t sim_card_power(uint32_t c, uint32_t trump, uint32_t follow) { return dk::card_power(c, trump, follow); }
// the table-driven forms the playout kernels use (the same shared-memory image, built by lut_word / sel12_entry)
SIM_API uint32_t sim_pick_msb_rank24_tab(uint32_t mask, uint32_t idx) { return dk::pick_msb_rank24_tab(mask, idx, reinterpret_cast<const uint64_t*>(card_lut() + dk::SEL12_LUT_BASE)); }
SIM_API uint32_t sim_pick_msb_rank24_lut(uint32_t mask, uint32_t idx) { return dk::pick_msb_rank24_lut(mask, idx, card_lut()); }
SIM_API uint32_t sim_pick_msb_rank24(uint32_t mask, uint32_t idx) { return dk::pick_msb_rank24(mask, idx); }
SIM_API uint32_t sim_pow_lookup(uint32_t gt, uint32_t first_card, uint32_t c) {      // the trick-accumulator record of card c in a trick led with first_card
    const uint32_t trump = dk::trump_mask_for_game_type(gt);
    return dk::pow_lookup(card_lut(), dk::pow_row(gt, first_card, dk::card_suit(first_card), trump), c);
}
SIM_API uint32_t sim_seg_lut(uint32_t win, uint32_t hit) { return reinterpret_cast<const uint8_t*>(card_lut() + dk::SEG_LUT_BASE)[16u * win + hit]; }
SIM_API uint32_t sim_thr2_lut(uint32_t w, uint32_t re_low, uint32_t ko_low) { return card_lut()[dk::THR2_LUT_BASE + 64u * w + 8u * re_low + ko_low]; }
SIM_API uint32_t sim_trump_mask(uint32_t gt) { return dk::trump_mask_for_game_type(gt); }
SIM_API uint32_t sim_follow_mask(uint32_t c, uint32_t trump) { return dk::follow_mask(c, trump); }

// ---- state record ops ------------------------------------------------------------------------------------------------
SIM_API void sim_new_game(dk_state* s, const uint64_t hands[4], uint32_t start) { dk::st_new_game(*s, hands, start); }
SIM_API uint64_t sim_legal_mask(int engine, const dk_state* s) { return engine == DK_FDO ? dk::fdo_state_legal_mask(*s) : dk::doko_state_legal_mask(*s); }
SIM_API uint32_t sim_apply(int engine, dk_state* s, uint32_t action, uint32_t flags) {
    return engine == DK_FDO ? dk::fdo_state_apply_az(*s, action, (flags & DK_APPLY_SKIP_SINGLE) != 0) : dk::doko_state_apply(*s, action);
}
namespace { struct RowOut {
    int64_t* row;
    void operator()(uint32_t i, uint32_t v) { row[i] = (int64_t)v; }
    void slot(uint32_t n, uint32_t tok, uint32_t pos, uint32_t ply, uint32_t sub, uint32_t team) {
        row[n] = tok; row[62 + n] = pos; row[124 + n] = ply; row[186 + n] = sub; row[248 + n] = team;
    }
    void phase(uint32_t v) { row[310] = v; }
}; }
SIM_API void sim_encode(int layout, const dk_state* s, int64_t* out) {
    RowOut o{out};
    if (layout == DK_LAYOUT_FDO_PI311) dk::fdo_encode_pi(*s, o); else dk::doko_encode(*s, layout == DK_LAYOUT_DO114, o);
}
SIM_API int sim_playout_from_state(int engine, const dk_state* s, uint64_t seed, uint64_t unit_lo_hi, uint32_t unit_hi, uint32_t epoch, int with_ann, int32_t* points, uint32_t* steps) {
    dk::RngKey key = make_key(seed, unit_lo_hi, epoch);
    key.unit_hi = unit_hi;
    if (engine == DK_FDO) {
        dk::FdoLive g; dk::FdoResume rs;
        if (!dk::fdo_state_to_live(*s, g, rs)) { for (int p = 0; p < 4; ++p) points[p] = s->points[p]; *steps = 0; return 0; }
        if (with_ann) dk::fdo_play_to_end<true, false>(g, key, &rs, card_lut()); else dk::fdo_play_to_end<false, false>(g, key, &rs, card_lut());
        dk::fdo_final_points(g, points); *steps = g.steps;
    } else {
        dk::DokoLive g; dk::DokoResume rs;
        if (!dk::doko_state_to_live(*s, g, rs)) { for (int p = 0; p < 4; ++p) points[p] = s->points[p]; *steps = 0; return 0; }
        dk::doko_play_to_end<false, false>(g, key, &rs, nullptr, card_lut());
        dk::doko_final_points(g, points); *steps = g.steps;
    }
    return 1;
}

// ---- determinization ---------------------------------------------------------------------------------------------------
SIM_API uint32_t sim_fdo_determinize(const dk_state* s, uint64_t seed, uint64_t unit, uint32_t sample, uint32_t epoch, uint64_t hands[4], uint8_t res[4]) {
    dk::MatchPrep prep; dk::fdo_match_prepare(*s, prep);
    dk::RngKey key = make_key(seed, unit, epoch); key.unit_hi = sample;
    // both forms of rule 4's rank select (compare / shift levels, and the kernel's 64-entry table) must give the same sample
    uint32_t rank6[64];
    for (uint32_t i = 0; i < 64u; ++i) rank6[i] = dk::rank_lut6_entry(i);
    uint64_t h2[4]; uint8_t r2[4];
    const uint32_t st2 = dk::fdo_match_sample(prep, key, h2, r2, rank6);
    const uint32_t st = dk::fdo_match_sample(prep, key, hands, res);
    if (st != st2 || std::memcmp(h2, hands, sizeof(h2)) != 0 || std::memcmp(r2, res, 4) != 0) return 0xFFFFFFFFu;
    return st;
}
SIM_API uint32_t sim_fdo_leaf_rollout(const dk_state* s, uint64_t seed, uint64_t unit, uint32_t rollout, uint32_t epoch, int determinize, int32_t* points, uint32_t* steps) {
    dk::RngKey key = make_key(seed, unit, epoch); key.unit_hi = rollout;
    dk_state st = *s;
    uint32_t status = 0;
    uint64_t hands[4] = {0, 0, 0, 0}; uint8_t res[4] = {0, 0, 0, 0};
    if (determinize) {
        dk::MatchPrep prep; dk::fdo_match_prepare(st, prep);
        status = dk::fdo_match_sample(prep, key, hands, res);
        dk::fdo_state_with_hands_and_reservations(st, hands, res);
    }
    for (int p = 0; p < 4; ++p) points[p] = 0;
    *steps = 0;
    if (status) return status;
    dk::FdoLive g; dk::FdoResume rs;
    if (!dk::fdo_state_to_live(st, g, rs)) { for (int p = 0; p < 4; ++p) points[p] = st.points[p]; return 0; }
    dk::fdo_play_to_end<false, false>(g, key, &rs, card_lut());
    dk::fdo_final_points(g, points); *steps = g.steps;
    if (determinize) {
        // the leaf-rollout kernel's form: the bridge of the INFO-STATE (built once per leaf), then only what the sample changes
        // (fdo_live_with_sample) — must give the same game as bridging the determinized record
        dk::FdoLive g2; dk::FdoResume rs2;
        if (!dk::fdo_state_to_live<true>(*s, g2, rs2)) return 0xFFFFFFFEu;
        const uint32_t res4 = (uint32_t)res[0] | ((uint32_t)res[1] << 8) | ((uint32_t)res[2] << 16) | ((uint32_t)res[3] << 24);
        dk::fdo_live_with_sample(g2, rs2, *s, hands, res4);
        dk::fdo_play_to_end<false, false, true, false>(g2, key, &rs2, card_lut());
        int32_t p2[4]; dk::fdo_final_points(g2, p2);
        if (std::memcmp(p2, points, sizeof(p2)) != 0) return 0xFFFFFFFDu;
    }
    return 0;
}

SIM_API uint32_t sim_doko_assign(const dk_state* s, uint64_t seed, uint64_t unit, uint32_t sample, uint32_t epoch, uint64_t hands[4]) {
    dk::AssignPrep prep; dk::doko_assign_prepare(*s, prep);
    dk::RngKey key = make_key(seed, unit, epoch); key.unit_hi = sample;
    // both forms of the multiset rank select (binary search / table for the last level, the kernel's) must give the same sample
    uint32_t adj3[64];
    for (uint32_t i = 0; i < 64u; ++i) adj3[i] = dk::adj3_entry(i);
    uint64_t h2[4];
    const uint32_t st2 = dk::doko_assign_sample(prep, key, h2, adj3);
    const uint32_t st = dk::doko_assign_sample(prep, key, hands);
    if (st != st2 || std::memcmp(h2, hands, sizeof(h2)) != 0) return 0xFFFFFFFFu;
    return st;
}

SIM_API uint32_t sim_fdo_ann_bits(uint64_t seed, uint64_t unit, uint32_t epoch, uint32_t ord, uint32_t m) {
    dk::RngKey key = make_key(seed, unit, epoch);
    // walk the stream in steps of 3 from the start so that consume / refill are exercised, then peek at `ord`
    dk::AnnBits st; dk::fdo_ann_open(st, key, ord % 3u);
    for (uint32_t k = ord % 3u; k < ord; k += 3u) dk::fdo_ann_consume(st, key, 3u);
    uint32_t walked = dk::fdo_ann_peek(st, m);
    dk::AnnBits direct; dk::fdo_ann_open(direct, key, ord);
    return walked == dk::fdo_ann_peek(direct, m) ? walked : 0xFFFFFFFFu;
}

// ---- PIMC (pimc.cuh + the per-thread body of fdo_pimc_kernel, run sequentially) -------------------------------------------------------
SIM_API uint32_t sim_fuse(int strategy, const uint32_t* visits, const uint8_t* status, uint32_t n_rows, uint64_t allowed, uint32_t* n_ok) {
    return strategy == 0 ? dk::fuse_max_n(visits, status, n_rows, allowed, n_ok) : dk::fuse_average(visits, status, n_rows, n_ok);
}
SIM_API void sim_root_stats(const uint32_t* visits, const uint8_t* status, uint32_t n_rows, uint64_t allowed, long long* stats) {
    dk::root_stats_accumulate(visits, status, n_rows, allowed, stats);
}
SIM_API uint32_t sim_root_pick(uint32_t strategy, const long long* stats, uint64_t allowed) { return dk::root_stats_pick(strategy, stats, allowed); }
SIM_API uint32_t sim_fdo_flat_mc(const dk_state* root, uint64_t seed, uint64_t unit, uint32_t det, uint32_t n_rollouts, uint32_t epoch,
                                 uint32_t visits[39], int64_t value_sum[39]) {
    for (int a = 0; a < 39; ++a) { visits[a] = 0; value_sum[a] = 0; }
    dk::MatchPrep prep; dk::fdo_match_prepare(*root, prep);
    if (!prep.valid) return 0;
    dk::RngKey key = make_key(seed, unit, epoch); key.unit_hi = det;
    uint64_t hands[4]; uint8_t res[4];
    uint32_t status = dk::fdo_match_sample(prep, key, hands, res);
    if (status) return status;
    dk_state ds = *root;
    dk::fdo_state_with_hands_and_reservations(ds, hands, res);
    const uint32_t mover = dk::st_cur(*root);
    for (uint32_t r = 0; r < n_rollouts; ++r) {
        key.unit_hi = det * n_rollouts + r;
        int best_v = 0; uint32_t best_a = dk::ACTION_NONE;
        for (uint64_t m = dk::fdo_state_legal_mask(ds); m; m &= m - 1ull) {
            uint32_t a = dk::ffs0ll(m);
            dk_state st = ds;
            dk::fdo_state_apply(st, a);
            int32_t p[4];
            dk::FdoLive g; dk::FdoResume rs;
            if (dk::fdo_state_to_live(st, g, rs)) { dk::fdo_play_to_end<false, false>(g, key, &rs, card_lut()); dk::fdo_final_points(g, p); }
            else for (int q = 0; q < 4; ++q) p[q] = st.points[q];
            int v = p[mover];
            value_sum[a] += v;
            if (best_a == dk::ACTION_NONE || v > best_v) { best_a = a; best_v = v; }
        }
        if (best_a != dk::ACTION_NONE) visits[best_a] += 1;
    }
    return 0;
}

// ---- self-play driver helpers (selfplay.cuh) ---------------------------------------------------------------------------------------
SIM_API uint64_t sim_sp_az_allowed(const dk_state* s, uint64_t az_epoch) { return dk::sp_az_allowed(*s, az_epoch); }
SIM_API float sim_sp_keep_draw(uint32_t word) { return dk::sp_keep_draw(word); }
SIM_API float sim_sp_value_target(const dk_state* s, uint32_t player, uint32_t k) { return dk::sp_value_target(*s, player, k); }

SIM_API uint32_t sim_encode_ipi(const dk_state* s, const uint64_t assumed[4], const uint8_t assumed_res[4], uint32_t next_player, int64_t* out) {
    RowOut o{out};
    return dk::fdo_encode_ipi(*s, assumed, assumed_res, next_player, o);
}

// ---- UCT search (uct.cuh; the per-thread body of fdo_uct_kernel) -------------------------------------------------------------------
SIM_API uint32_t sim_fdo_uct_search(const dk_state* root, uint64_t seed, uint64_t unit, uint32_t sub, uint32_t epoch, int determinize, uint32_t iterations,
                                    float uct_c, uint32_t visits[39], float values[39], int32_t* action_out) {
    for (int a = 0; a < 39; ++a) { visits[a] = 0; values[a] = 0.0f; }
    *action_out = -1;
    dk_state s = *root;
    dk::RngKey key = make_key(seed, unit, epoch);
    if (determinize && dk::st_phase(s) != DK_PHASE_FINISHED) {
        dk::MatchPrep prep; dk::fdo_match_prepare(s, prep);
        key.unit_hi = sub;
        uint64_t hands[4]; uint8_t res[4];
        uint32_t status = dk::fdo_match_sample(prep, key, hands, res);
        if (status) return status;
        dk::fdo_state_with_hands_and_reservations(s, hands, res);
    }
    std::vector<double> ln_table(iterations + 1, 0.0);
    for (uint32_t n = 1; n <= iterations; ++n) ln_table[n] = std::log((double)n);
    // the kernels' phases, run in sequence for ONE tree (n_trees = 1) on a host workspace carved like the device one
    std::vector<char> ws((size_t)dk::uct_workspace_bytes(1, iterations));
    const dk::UctPool P = dk::uct_carve(ws.data(), 1, iterations);
    dk::RngKey det_key = key; det_key.unit_hi = sub;
    // (the determinization above already replaced `s`; uct_phase_root is handed the result and told not to determinize again)
    if (dk::uct_phase_root(P, 0, s, false, det_key)) return P.status[0];
    dk::UctTables T; T.ln = ln_table.data();
    for (uint32_t it = 0; it < iterations; ++it) {
        if (!(P.ctl[0] & dk::UCT_CTL_ACTIVE)) return P.status[0];
        key.unit_hi = sub * iterations + it;
        dk::uct_phase_tree<false>(P, 0, it, (double)uct_c, T, key);                            // uct_tree_kernel
        if (!(P.ctl[0] & dk::UCT_CTL_ACTIVE)) return P.status[0];
        uint32_t packed = P.result[0];
        if (P.ctl[0] & dk::UCT_CTL_ROLLOUT) packed = dk::uct_phase_rollout<false>(P, 0, key, card_lut());   // uct_rollout_kernel
        dk::uct_phase_backprop(P, 0, packed);
    }
    uint32_t best = dk::uct_moves(P, 0, visits, values);
    *action_out = best == 0xFFu ? -1 : (int32_t)best;
    return 0;
}
SIM_API uint64_t sim_uct_allowed(const dk_state* s, int first) { return dk::uct_allowed(*s, first != 0); }
SIM_API uint32_t sim_fdo_min_cards_to_call(uint32_t m, uint32_t e, uint32_t w) {
    uint32_t a = dk::fdo_min_cards_to_call(m, e, w), b = dk::fdo_min_cards_to_call_lut(m, e, w, card_lut());
    return a == b ? a : 0xFFFFFFFFu;     // closed form and table form must agree
}

// find_best_child on a synthetic family: child slot chosen with the f32 filter (low byte) and by the all-f64 evaluation (next byte)
SIM_API uint32_t sim_uct_select_check(uint32_t nch, const uint32_t* vis, const long long* win, uint32_t parent_visits, float uct_c) {
    uint32_t v[dk::UCT_MAX_CHILDREN] = {0};
    int32_t w[dk::UCT_MAX_CHILDREN] = {0};
    for (uint32_t k = 0; k < nch; ++k) { v[k] = vis[k]; w[k] = (int32_t)win[k]; }
    const double ln_n = std::log((double)parent_visits);
    uint32_t a = dk::uct_best_slot(nch, v, w, parent_visits, (double)uct_c, nullptr, ln_n, true) + 1u;
    uint32_t b = dk::uct_best_slot(nch, v, w, parent_visits, (double)uct_c, nullptr, ln_n, false) + 1u;
    return (a & 255u) | ((b & 255u) << 8);
}
