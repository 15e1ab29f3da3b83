// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// C API over the CPU restatement, loaded by tests/ (ctypes), __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` leg.  The product library (libdoko_cuda.so) never links this.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../include/doko_cuda.h"
#include "assignment.hpp"
#include "bitflag.hpp"
#include "doko.hpp"
#include "encode.hpp"
#include "fdo.hpp"
#include "matching.hpp"
#include "mcts.hpp"
#include "pimc.hpp"
#include "replay.hpp"
#include "rng.hpp"
#include "selfplay.hpp"

using namespace oracle;

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {
thread_local char g_err[256] = {0};
template <class F> int guarded(F&& f) {
    try { f(); return 0; } catch (const std::exception& e) { std::strncpy(g_err, e.what(), sizeof g_err - 1); return 1; }
}
int ann_bit_to_code(int bit) {
    switch (bit) { case fdo::A_NONE: return DK_ANN_NONE; case fdo::A_RE_CONTRA: return DK_ANN_RE_CONTRA; case fdo::A_NO90: return DK_ANN_NO90;
        case fdo::A_NO60: return DK_ANN_NO60; case fdo::A_NO30: return DK_ANN_NO30; case fdo::A_BLACK: return DK_ANN_BLACK;
        case fdo::A_COUNTER_RE_CONTRA: return DK_ANN_COUNTER; }
    throw std::runtime_error("bad announcement bit");
}
int ann_code_to_bit(int code) {
    static const int b[7] = {fdo::A_NONE, fdo::A_RE_CONTRA, fdo::A_NO90, fdo::A_NO60, fdo::A_NO30, fdo::A_BLACK, fdo::A_COUNTER_RE_CONTRA};
    if (code < 0 || code > 6) throw std::runtime_error("bad announcement code");
    return b[code];
}
}  // namespace

ORC_API const char* orc_last_error() { return g_err; }

// ---- bit utilities / RNG ----------------------------------------------------------------------
ORC_API uint64_t orc_select_by_rank(uint64_t v, uint64_t r) { return select_by_rank(v, r); }
ORC_API void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox4x32::block(ctr, key, out); }
ORC_API uint32_t orc_philox_word(uint64_t seed, uint32_t unit_lo, uint32_t unit_hi, uint32_t epoch, uint32_t site, uint32_t ordinal) {
    return philox_word(seed, unit_lo, unit_hi, epoch, site, ordinal);
}
// `count` consecutive draws of sub-stream `site` of the parity stream, starting at ordinal `first` (inside a trick, `chain_mul` is the
// product of the counts already drawn from the trick's word: Rng::set_card_position).
ORC_API void orc_philox_draws(uint64_t seed, uint32_t unit_lo, uint32_t unit_hi, uint32_t epoch, uint32_t site, uint32_t first, uint32_t chain_mul,
                              int count, const uint32_t* n, uint32_t* out) {
    PhiloxStream r(seed, unit_lo, unit_hi, epoch);
    if (site == SITE_CARD) r.set_card_position(first, chain_mul); else r.set_ordinal((Site)site, first);
    for (int i = 0; i < count; ++i) out[i] = r.below((Site)site, n[i]);
}
// `count` (draw, chained draw) pairs of sub-stream `site`: below(site, n0[i]) and then below_chained(site, n1[i]) from the same word.
ORC_API void orc_philox_pair_draws(uint64_t seed, uint32_t unit_lo, uint32_t unit_hi, uint32_t epoch, uint32_t site, uint32_t first, int count,
                                   const uint32_t* n0, const uint32_t* n1, uint32_t* out0, uint32_t* out1) {
    PhiloxStream r(seed, unit_lo, unit_hi, epoch);
    r.set_ordinal((Site)site, first);
    for (int i = 0; i < count; ++i) { out0[i] = r.below((Site)site, n0[i]); out1[i] = r.below_chained((Site)site, n1[i]); }
}
ORC_API void orc_smallrng_u64(uint64_t seed, int count, uint64_t* out) { SmallRngStream r(seed); for (int i = 0; i < count; ++i) out[i] = r.next_u64(); }
ORC_API void orc_smallrng_ranges(uint64_t seed, int count, const uint32_t* n, uint32_t* out) {
    SmallRngStream r(seed); for (int i = 0; i < count; ++i) out[i] = r.range_u32(n[i]);
}
// rs-doko/src/hand/hand_random.rs:68-83 and rs-full-doko/src/hand/hand.rs:559-585 (shuffle only, no start player)
ORC_API void orc_smallrng_distribute_cards(uint64_t seed, int engine, uint64_t hands[4]) {
    SmallRngStream r(seed);
    if (engine == DK_DOKO) doko::distribute_cards(r, hands);
    else { fdo::Hand h[4]; fdo::randomly_distributed(r, h); for (int p = 0; p < 4; ++p) hands[p] = h[p].bits; }
}
// rs-doko/src/util/bitflag/bitflag.rs:190-206: six `slow` picks (usize range over to_vec) then four fast picks.
ORC_API void orc_smallrng_bitflag_picks(uint64_t seed, uint64_t bitflag, int n_slow, int n_fast, uint64_t* out) {
    SmallRngStream r(seed);
    uint64_t v[64]; int m = bitflag_to_vec(bitflag, v);
    int k = 0;
    for (int i = 0; i < n_slow; ++i) out[k++] = v[r.range_u32((uint32_t)m)];
    for (int i = 0; i < n_fast; ++i) out[k++] = random_single(bitflag, r, SITE_CARD);
}

// ---- rs-full-doko state handles -------------------------------------------------------------------
ORC_API void* orc_fdo_new(const uint64_t hands[4], int start_player) {
    fdo::Hand h[4]; for (int p = 0; p < 4; ++p) h[p].bits = hands[p];
    return new fdo::State(fdo::State::new_game_from_hand_and_start_player(h, start_player));
}
ORC_API void* orc_fdo_new_game_philox(uint64_t seed, uint64_t unit, uint32_t epoch) {
    PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
    return new fdo::State(fdo::State::new_game(r));
}
ORC_API void* orc_fdo_new_game_smallrng(uint64_t seed) { SmallRngStream r(seed); return new fdo::State(fdo::State::new_game(r)); }
ORC_API void* orc_fdo_clone(const void* h) { return new fdo::State(*(const fdo::State*)h); }
ORC_API void orc_fdo_free(void* h) { delete (fdo::State*)h; }
ORC_API int orc_fdo_play(void* h, int action) { return guarded([&] { ((fdo::State*)h)->play_action(action); }); }
ORC_API uint64_t orc_fdo_allowed(const void* h) { uint64_t m = 0; guarded([&] { m = ((const fdo::State*)h)->allowed_actions(); }); return m; }
// out: [phase, current_player, game_type, card_index, n_tricks, team_tag, wedding_player, solved_idx, re_players,
//       re_lowest(code), contra_lowest(code), turns_without, ann_start, n_announcements, current_allowed(bits),
//       eyes[4], num_tricks[4], points[4], n_play_actions, is_solo, re_eyes, kontra_eyes, re_points, kontra_points]  (36 ints)
ORC_API void orc_fdo_info(const void* h, int32_t* o) {
    const fdo::State& s = *(const fdo::State*)h;
    int k = 0;
    o[k++] = s.current_phase; o[k++] = s.current_player; o[k++] = s.game_type; o[k++] = s.card_index; o[k++] = s.n_tricks;
    o[k++] = s.team_state.tag; o[k++] = s.team_state.wedding_player; o[k++] = s.team_state.solved_trick_index; o[k++] = (int)s.team_state.re_players;
    o[k++] = ann_bit_to_code(s.announcements.re_lowest); o[k++] = ann_bit_to_code(s.announcements.contra_lowest);
    o[k++] = s.announcements.turns_without; o[k++] = s.announcements.starting_player; o[k++] = s.announcements.n;
    o[k++] = (int)s.announcements.current_allowed;
    for (int p = 0; p < 4; ++p) o[k++] = (int)s.player_eyes[p];
    for (int p = 0; p < 4; ++p) o[k++] = (int)s.player_num_tricks[p];
    for (int p = 0; p < 4; ++p) o[k++] = s.end_of_game_stats.player_points[p];
    o[k++] = (int)s.n_play_actions; o[k++] = s.end_of_game_stats.is_solo; o[k++] = (int)s.end_of_game_stats.re_eyes;
    o[k++] = (int)s.end_of_game_stats.kontra_eyes; o[k++] = s.end_of_game_stats.re_points; o[k++] = s.end_of_game_stats.kontra_points;
}
ORC_API void orc_fdo_hands(const void* h, uint64_t out[4]) { for (int p = 0; p < 4; ++p) out[p] = ((const fdo::State*)h)->hands[p].bits; }
// tricks: out[t*6 + {0..3 cards, 4 start, 5 winner}] (-1 = none)
ORC_API void orc_fdo_tricks(const void* h, int32_t* out) {
    const fdo::State& s = *(const fdo::State*)h;
    for (int t = 0; t < 12; ++t) {
        for (int k = 0; k < 4; ++k) out[t * 6 + k] = (t < s.n_tricks && k < s.tricks[t].len) ? s.tricks[t].cards[k] : -1;
        out[t * 6 + 4] = t < s.n_tricks ? s.tricks[t].starting_player : -1;
        out[t * 6 + 5] = t < s.n_tricks ? s.tricks[t].winning_player : -1;
    }
}
// additional point details of a finished game: [present, against_cq, doko_re, doko_kontra, fuchs_re, fuchs_kontra, karl_re, karl_kontra]
ORC_API void orc_fdo_additional(const void* h, int32_t* o) {
    const fdo::AdditionalPointsDetails& d = ((const fdo::State*)h)->end_of_game_stats.additional;
    o[0] = d.present; o[1] = d.against_club_queens; o[2] = d.doko_re; o[3] = d.doko_kontra; o[4] = d.fuchs_re; o[5] = d.fuchs_kontra;
    o[6] = d.karlchen_re; o[7] = d.karlchen_kontra;
}
ORC_API void orc_fdo_visible_reservations(const void* h, int observing_player, int32_t out[4]) {
    int v[4]; fdo::get_visible_reservations(((const fdo::State*)h)->reservations_round, observing_player, v);
    for (int p = 0; p < 4; ++p) out[p] = v[p];
}
ORC_API int orc_fdo_encode_pi(const void* h, int64_t out[311]) { return guarded([&] { fdo::encode_state_pi(*(const fdo::State*)h, out); }); }
// random step with the Philox parity stream; returns the action or -1 when finished
ORC_API int orc_fdo_random_step_philox(void* h, uint64_t seed, uint64_t unit, uint32_t epoch, int with_announcements, uint32_t ann_ordinal) {
    fdo::State& s = *(fdo::State*)h;
    PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
    s.position_streams(r);
    r.set_ordinal(SITE_ANNOUNCEMENT, ann_ordinal);
    int a = -1;
    bool fin = with_announcements ? s.random_action_for_current_player(r, &a) : s.random_action_for_current_player_no_announcement(r, &a);
    return fin ? -1 : a;
}

// One lock-step env step with the SITE_STEP contract (word 0, unit = game id, epoch = caller's step counter):
// allowed actions (minus the five call actions unless with_announcements) → MSB-rank pick → play_action
// [→ skip_single: keep playing while exactly one non-call action is legal, full_doko.rs:127-151].  Returns the action or -1.
ORC_API int orc_fdo_step_site(void* h, uint64_t seed, uint64_t unit, uint32_t epoch, int with_announcements, int skip_single) {
    fdo::State& s = *(fdo::State*)h;
    if (s.current_phase == fdo::PH_FINISHED) return -1;
    uint64_t allowed = s.allowed_actions();
    if (!with_announcements) allowed &= ~fdo::ANNOUNCEMENT_CALL_ACTIONS;
    uint32_t w = philox_word(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch, SITE_STEP, 0);
    uint64_t bit = select_by_rank(allowed, mul_shift(w, popcount64(allowed)));
    int a = __builtin_ctzll(bit);
    s.play_action(a);
    if (skip_single) {
        for (;;) {
            if (s.current_phase == fdo::PH_FINISHED) break;
            uint64_t m = s.allowed_actions() & ~fdo::ANNOUNCEMENT_CALL_ACTIONS;
            if (popcount64(m) != 1) break;
            s.play_action(__builtin_ctzll(m));
        }
    }
    return a;
}

// ---- dk_state export / import (layout documented in include/doko_cuda.h) -----------------------------
static void export_fdo(const fdo::State& s, dk_state* o) {
    std::memset(o, 0, sizeof *o);
    for (int p = 0; p < 4; ++p) o->hands[p] = s.hands[p].bits;
    std::memset(o->cards, 0xFF, 48);
    int ci = 0;
    uint32_t tr = 0;
    for (int t = 0; t < s.n_tricks; ++t) {
        tr |= (uint32_t)s.tricks[t].starting_player << (2 * t);
        for (int k = 0; k < s.tricks[t].len; ++k) o->cards[ci++] = (uint8_t)s.tricks[t].cards[k];
    }
    for (int a = 0; a < 12; ++a) o->announcements[a] = 0xFFFF;
    for (int a = 0; a < s.announcements.n; ++a) {
        const fdo::AnnouncementOccurrence& oc = s.announcements.occ[a];
        o->announcements[a] = (uint16_t)(oc.card_index | (oc.player << 6) | (ann_bit_to_code(oc.announcement) << 8));
    }
    for (int i = 0; i < 4; ++i) o->reservations[i] = i < s.reservations_round.len ? (uint8_t)s.reservations_round.r[i] : (uint8_t)DK_RES_NONE;
    tr |= (uint32_t)s.n_tricks << 24; tr |= (uint32_t)s.announcements.n << 28;
    o->tricks = tr;
    uint16_t nt = 0;
    for (int p = 0; p < 4; ++p) { o->eyes[p] = (uint8_t)s.player_eyes[p]; nt |= (uint16_t)(s.player_num_tricks[p] << (4 * p)); }
    o->num_tricks = nt;
    o->card_index = (uint8_t)s.card_index;
    o->n_reservations = (uint8_t)s.reservations_round.len;
    for (int p = 0; p < 4; ++p) o->points[p] = s.current_phase == fdo::PH_FINISHED ? (int8_t)s.end_of_game_stats.player_points[p] : 0;
    uint32_t m = 0;
    m |= (uint32_t)s.current_phase;
    m |= (uint32_t)(s.current_player < 0 ? 0 : s.current_player) << 2;
    m |= (uint32_t)s.reservations_round.starting_player << 4;
    m |= (uint32_t)(s.game_type < 0 ? DK_GT_NONE : s.game_type) << 6;
    m |= (uint32_t)s.team_state.tag << 10;
    m |= (uint32_t)(s.team_state.wedding_player < 0 ? 0 : s.team_state.wedding_player) << 12;
    m |= (uint32_t)(s.team_state.tag == fdo::TS_WEDDING_SOLVED ? s.team_state.solved_trick_index : 0) << 14;
    m |= (uint32_t)(s.team_state.has_re_players() ? s.team_state.re_players : 0) << 16;
    m |= (uint32_t)ann_bit_to_code(s.announcements.re_lowest) << 20;
    m |= (uint32_t)ann_bit_to_code(s.announcements.contra_lowest) << 23;
    m |= (uint32_t)s.announcements.turns_without << 26;
    m |= (uint32_t)s.announcements.starting_player << 29;
    o->meta = m;
}
static fdo::State import_fdo(const dk_state* o) {
    fdo::State s;
    uint32_t m = o->meta, tr = o->tricks;
    s.reservations_round.starting_player = (int8_t)((m >> 4) & 3);
    for (int i = 0; i < o->n_reservations; ++i) s.reservations_round.play_reservation(o->reservations[i]);
    for (int p = 0; p < 4; ++p) { s.hands[p].bits = o->hands[p]; s.player_eyes[p] = o->eyes[p]; s.player_num_tricks[p] = (o->num_tricks >> (4 * p)) & 15; }
    s.card_index = o->card_index;
    s.current_phase = (int)(m & 3);
    s.current_player = s.current_phase == fdo::PH_FINISHED ? -1 : (int)((m >> 2) & 3);
    int gt = (int)((m >> 6) & 15); s.game_type = gt == DK_GT_NONE ? fdo::GT_NONE : gt;
    if (s.reservations_round.is_completed()) s.reservation_result = fdo::winning_player_in_reservation_round(s.reservations_round);
    s.team_state.tag = (int)((m >> 10) & 3);
    bool wed = s.team_state.tag == fdo::TS_WEDDING_UNSOLVED || s.team_state.tag == fdo::TS_WEDDING_SOLVED;
    s.team_state.wedding_player = wed ? (int)((m >> 12) & 3) : -1;
    s.team_state.solved_trick_index = (int)((m >> 14) & 3);
    s.team_state.re_players = (m >> 16) & 15;
    s.announcements.re_lowest = ann_code_to_bit((int)((m >> 20) & 7));
    s.announcements.contra_lowest = ann_code_to_bit((int)((m >> 23) & 7));
    s.announcements.turns_without = (int)((m >> 26) & 7);
    s.announcements.starting_player = (int)((m >> 29) & 3);
    s.announcements.n = (int)((tr >> 28) & 15);
    for (int a = 0; a < s.announcements.n; ++a) {
        uint16_t v = o->announcements[a];
        s.announcements.occ[a].card_index = (uint8_t)(v & 63); s.announcements.occ[a].player = (uint8_t)((v >> 6) & 3);
        s.announcements.occ[a].announcement = (uint8_t)ann_code_to_bit((v >> 8) & 7);
    }
    s.n_tricks = (int)((tr >> 24) & 15);
    int ci = 0;
    for (int t = 0; t < s.n_tricks; ++t) {
        s.tricks[t] = fdo::Trick::empty((int)((tr >> (2 * t)) & 3));
        for (int k = 0; k < 4 && ci < s.card_index; ++k) s.tricks[t].play_card(o->cards[ci++], s.game_type);
    }
    if (s.current_phase == fdo::PH_ANNOUNCEMENT) {
        uint32_t lens[4]; s.hand_lens(lens);
        s.announcements.current_allowed = fdo::calc_allowed_announcements(s.current_player, (int)lens[s.current_player], s.team_state,
                                                                          s.announcements.re_lowest, s.announcements.contra_lowest);
    }
    if (s.current_phase == fdo::PH_FINISHED) {
        s.end_of_game_stats = fdo::calculate_end_of_game_stats(s.player_eyes, s.player_num_tricks, s.team_state.re_players,
                                                               s.announcements.re_lowest, s.announcements.contra_lowest, s.tricks, s.n_tricks);
    }
    return s;
}
ORC_API void orc_fdo_export(const void* h, dk_state* out) { export_fdo(*(const fdo::State*)h, out); }
ORC_API void* orc_fdo_import(const dk_state* in) { fdo::State* s = nullptr; guarded([&] { s = new fdo::State(import_fdo(in)); }); return s; }

// ---- rs-doko state handles -------------------------------------------------------------------------------
static void export_doko(const doko::State& s, dk_state* o) {
    std::memset(o, 0, sizeof *o);
    for (int p = 0; p < 4; ++p) o->hands[p] = s.hands[p];
    std::memset(o->cards, 0xFF, 48);
    for (int a = 0; a < 12; ++a) o->announcements[a] = 0xFFFF;
    int ci = 0, nt = 0; uint32_t tr = 0;
    for (int t = 0; t < 12; ++t) {
        if (!s.tricks[t].present) break;
        nt++;
        tr |= (uint32_t)s.tricks[t].start_player << (2 * t);
        for (int k = 0; k < 4; ++k) if (s.tricks[t].cards[k] >= 0) o->cards[ci++] = (uint8_t)s.tricks[t].cards[k];
    }
    tr |= (uint32_t)nt << 24;
    o->tricks = tr;
    int nres = s.reservations_round.len();
    for (int i = 0; i < 4; ++i) o->reservations[i] = i < nres ? (uint8_t)s.reservations_round.r[i] : (uint8_t)DK_RES_NONE;
    uint16_t ntr = 0;
    for (int p = 0; p < 4; ++p) { o->eyes[p] = (uint8_t)s.player_eyes[p]; ntr |= (uint16_t)(s.player_num_tricks[p] << (4 * p)); }
    o->num_tricks = ntr; o->card_index = (uint8_t)ci; o->n_reservations = (uint8_t)nres;
    int phase = s.current_phase == doko::PH_RESERVATION ? DK_PHASE_RESERVATION : (s.current_phase == doko::PH_PLAYCARD ? DK_PHASE_PLAY_CARD : DK_PHASE_FINISHED);
    for (int p = 0; p < 4; ++p) o->points[p] = phase == DK_PHASE_FINISHED ? (int8_t)s.end_of_game_stats.player_points[p] : 0;
    uint32_t m = 0;
    m |= (uint32_t)phase;
    m |= (uint32_t)(s.current_player < 0 ? 0 : s.current_player) << 2;
    m |= (uint32_t)s.reservations_round.start_player << 4;
    int gt = !s.has_reservation_result ? DK_GT_NONE : (s.wedding_player_result >= 0 ? DK_GT_WEDDING : DK_GT_NORMAL);
    m |= (uint32_t)gt << 6;
    m |= (uint32_t)s.team_state.tag << 10;
    m |= (uint32_t)(s.team_state.wedding_player < 0 ? 0 : s.team_state.wedding_player) << 12;
    m |= (uint32_t)(s.team_state.tag == doko::TS_WEDDING_SOLVED ? s.team_state.solved_trick_index : 0) << 14;
    m |= (uint32_t)(s.team_state.is_final() ? s.team_state.re_players : 0) << 16;
    o->meta = m;
}
ORC_API void* orc_doko_new(const uint64_t hands[4], int start_player) { return new doko::State(doko::State::new_game_from_hand_and_start_player(hands, start_player)); }
ORC_API void* orc_doko_new_game_philox(uint64_t seed, uint64_t unit, uint32_t epoch) {
    PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
    return new doko::State(doko::State::new_game(r));
}
ORC_API void* orc_doko_new_game_smallrng_play(uint64_t seed, int n_random_actions) {
    SmallRngStream r(seed);
    doko::State* s = new doko::State(doko::State::new_game(r));
    for (int i = 0; i < n_random_actions; ++i) s->random_action_for_current_player(r);
    return s;
}
ORC_API void* orc_doko_clone(const void* h) { return new doko::State(*(const doko::State*)h); }
ORC_API void orc_doko_free(void* h) { delete (doko::State*)h; }
ORC_API int orc_doko_play(void* h, int action) { return guarded([&] { ((doko::State*)h)->play_action(action); }); }
ORC_API uint64_t orc_doko_allowed(const void* h) { return ((const doko::State*)h)->allowed_actions(); }
// [phase(doko enum), current_player, trick_index, team_tag, wedding_player, solved_idx, re_players, eyes[4], ntricks[4], points[4], n_play_actions, start_player]
ORC_API void orc_doko_info(const void* h, int32_t* o) {
    const doko::State& s = *(const doko::State*)h; int k = 0;
    o[k++] = s.current_phase; o[k++] = s.current_player; o[k++] = s.current_trick_index; o[k++] = s.team_state.tag;
    o[k++] = s.team_state.wedding_player; o[k++] = s.team_state.solved_trick_index; o[k++] = (int)s.team_state.re_players;
    for (int p = 0; p < 4; ++p) o[k++] = (int)s.player_eyes[p];
    for (int p = 0; p < 4; ++p) o[k++] = (int)s.player_num_tricks[p];
    for (int p = 0; p < 4; ++p) o[k++] = s.end_of_game_stats.player_points[p];
    o[k++] = (int)s.n_play_actions; o[k++] = s.reservations_round.start_player;
}
ORC_API void orc_doko_hands(const void* h, uint64_t out[4]) { for (int p = 0; p < 4; ++p) out[p] = ((const doko::State*)h)->hands[p]; }
ORC_API int orc_doko_encode(const void* h, int with_reservations, int64_t* out) { return doko::encode_state(*(const doko::State*)h, out, with_reservations != 0); }
ORC_API void orc_doko_export(const void* h, dk_state* out) { export_doko(*(const doko::State*)h, out); }
ORC_API int orc_doko_random_step_philox(void* h, uint64_t seed, uint64_t unit, uint32_t epoch) {
    doko::State& s = *(doko::State*)h;
    PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
    s.position_streams(r);
    int a = -1; bool fin = s.random_action_for_current_player(r, &a);
    return fin ? -1 : a;
}

// ---- pure rule functions for the table-style known-answer tests ---------------------------------------------
ORC_API int orc_fdo_card_to_color(int card, int gt) { return fdo::card_to_color(card, gt); }
ORC_API void orc_fdo_color_masks(int gt, uint64_t out[5]) { fdo::ColorMasks m = fdo::get_color_masks_for_game_type(gt); for (int i = 0; i < 5; ++i) out[i] = m.m[i]; }
ORC_API int orc_fdo_is_greater_in_trick(int cur, int prev, int color, int gt) { return fdo::is_greater_in_trick(cur, prev, (fdo::Color)color, gt); }
ORC_API uint32_t orc_fdo_card_eyes(int c) { return fdo::card_eyes(c); }
ORC_API int orc_fdo_trick_winner(int start, const int32_t cards[4], int gt, int32_t* winning_card, int32_t* eyes, int32_t* color) {
    fdo::Trick t = fdo::Trick::empty(start);
    for (int k = 0; k < 4; ++k) t.play_card(cards[k], gt);
    if (winning_card) *winning_card = t.winning_card;
    if (eyes) *eyes = (int)t.eyes();
    if (color) *color = t.color(gt);
    return t.winning_player;
}
// hand ops: op 0 add, 1 remove, 2 remove_ignore, 3 remove_both, 4 add_ignore; returns 1 on panic
ORC_API int orc_fdo_hand_op(uint64_t* bits, int op, int card) {
    fdo::Hand h; h.bits = *bits;
    int rc = guarded([&] { switch (op) { case 0: h.add(card); break; case 1: h.remove(card); break; case 2: h.remove_ignore(card); break;
                                         case 3: h.remove_both(card); break; case 4: h.add_ignore(card); break; } });
    *bits = h.bits; return rc;
}
ORC_API uint64_t orc_fdo_hand_plus(uint64_t a, uint64_t b) { fdo::Hand x, y; x.bits = a; y.bits = b; return x.plus_hand(y).bits; }
ORC_API uint64_t orc_fdo_hand_minus(uint64_t a, uint64_t b) { fdo::Hand x, y; x.bits = a; y.bits = b; uint64_t r = 0; guarded([&] { r = x.minus_hand(y).bits; }); return r; }
ORC_API uint64_t orc_fdo_hand_remove_color(uint64_t a, int color, int gt) { fdo::Hand x; x.bits = a; x.remove_color((fdo::Color)color, gt); return x.bits; }
ORC_API int orc_fdo_hand_iter(uint64_t a, int32_t* out) { fdo::Hand x; x.bits = a; int c[48]; int n = x.iter(c); for (int i = 0; i < n; ++i) out[i] = c[i]; return n; }
// reservation round: reservations in play order; out = [kind, player, reservation, game_type]
ORC_API void orc_fdo_reservation_result(int start, int n, const int32_t* res, int32_t out[4]) {
    fdo::ReservationRound rr; rr.starting_player = (int8_t)start; for (int i = 0; i < n; ++i) rr.play_reservation(res[i]);
    fdo::ReservationResult r = fdo::winning_player_in_reservation_round(rr);
    out[0] = r.kind; out[1] = r.player; out[2] = r.reservation; out[3] = fdo::to_game_type(r);
}
ORC_API void orc_fdo_visible(int start, int n, const int32_t* res, int observing, int32_t out[4]) {
    fdo::ReservationRound rr; rr.starting_player = (int8_t)start; for (int i = 0; i < n; ++i) rr.play_reservation(res[i]);
    int v[4]; fdo::get_visible_reservations(rr, observing, v); for (int p = 0; p < 4; ++p) out[p] = v[p];
}
// team_resolve: rr = [kind, player, reservation]; tricks[t*5+{0..3 cards,4 start}] (card -1 = absent); out = [tag, wedding_player, solved_idx, re_players]
ORC_API void orc_fdo_team_resolve(const int32_t rr_in[3], int n_tricks, const int32_t* tricks, const uint64_t hands[4], int32_t out[4]) {
    fdo::ReservationResult rr; rr.kind = rr_in[0]; rr.player = rr_in[1]; rr.reservation = rr_in[2];
    fdo::Trick ts[12];
    for (int t = 0; t < n_tricks; ++t) { ts[t] = fdo::Trick::empty(tricks[t * 5 + 4]); for (int k = 0; k < 4; ++k) if (tricks[t * 5 + k] >= 0) ts[t].play_card(tricks[t * 5 + k], fdo::GT_NORMAL); }
    fdo::Hand h[4]; for (int p = 0; p < 4; ++p) h[p].bits = hands[p];
    fdo::TeamState r = fdo::team_resolve(rr, ts, n_tricks, h);
    out[0] = r.tag; out[1] = r.wedding_player; out[2] = r.solved_trick_index; out[3] = (int)r.re_players;
}
// announcements (bit-valued sets / Option = 0)
ORC_API uint32_t orc_fdo_all_higher_than(int lowest) { return fdo::all_higher_than(lowest); }
ORC_API int orc_fdo_cards_possible_last(uint32_t prev, int wedding_solved) { return fdo::cards_possible_for_last_announcement(prev, wedding_solved); }
ORC_API uint32_t orc_fdo_internal_calc_allowed(int n_cards, uint32_t prev_team, int wedding_solved, int enemy_possible) {
    return fdo::internal_calc_allowed_announcements(n_cards, prev_team, wedding_solved, enemy_possible);
}
ORC_API uint32_t orc_fdo_calc_allowed(int player, int n_cards, int tag, int wedding_player, int solved_idx, uint32_t re_players, int re_lowest, int contra_lowest) {
    fdo::TeamState ts; ts.tag = tag; ts.wedding_player = wedding_player; ts.solved_trick_index = solved_idx; ts.re_players = re_players;
    uint32_t r = 0; guarded([&] { r = fdo::calc_allowed_announcements(player, n_cards, ts, re_lowest, contra_lowest); }); return r;
}
// scoring
ORC_API int orc_fdo_re_won(uint32_t re_eyes, uint32_t re_prev, uint32_t k_prev, int re_all, int k_all) { return fdo::re_won(re_eyes, re_prev, k_prev, re_all, k_all); }
ORC_API int orc_fdo_kontra_won(uint32_t k_eyes, uint32_t re_prev, uint32_t k_prev, int re_all, int k_all) { return fdo::kontra_won(k_eyes, re_prev, k_prev, re_all, k_all); }
ORC_API void orc_fdo_basic_winning_points(uint32_t we, uint32_t le, int wall, uint32_t re_prev, uint32_t k_prev, uint32_t re_eyes, uint32_t k_eyes, int32_t out[25]) {
    int w, l, d[23]; fdo::basic_winning_points(we, le, wall, re_prev, k_prev, re_eyes, k_eyes, w, l, d);
    out[0] = w; out[1] = l; for (int i = 0; i < 23; ++i) out[2 + i] = d[i];
}
ORC_API void orc_fdo_basic_draw_points(uint32_t re_prev, uint32_t k_prev, uint32_t re_eyes, uint32_t k_eyes, int32_t out[16]) {
    int r, k, d[14]; fdo::basic_draw_points(re_prev, k_prev, re_eyes, k_eyes, r, k, d);
    out[0] = r; out[1] = k; for (int i = 0; i < 14; ++i) out[2 + i] = d[i];
}
// end of game stats from raw inputs; tricks[t*5+{0..3 cards, 4 start}] (winners computed under game type Normal like
// FdoTrick::existing); out = [re_eyes, kontra_eyes, re_points, kontra_points, pts[4], is_solo, has_win, has_draw, add[8], win[23], draw[14]]
ORC_API void orc_fdo_end_of_game_stats(const uint32_t eyes[4], const uint32_t ntricks[4], uint32_t re_players, int re_lowest, int contra_lowest,
                                       const int32_t* tricks, int32_t* out) {
    fdo::Trick ts[12];
    for (int t = 0; t < 12; ++t) { ts[t] = fdo::Trick::empty(tricks[t * 5 + 4]); for (int k = 0; k < 4; ++k) ts[t].play_card(tricks[t * 5 + k], fdo::GT_NORMAL); }
    fdo::EndOfGameStats s = fdo::calculate_end_of_game_stats(eyes, ntricks, re_players, re_lowest, contra_lowest, ts, 12);
    int k = 0;
    out[k++] = (int)s.re_eyes; out[k++] = (int)s.kontra_eyes; out[k++] = s.re_points; out[k++] = s.kontra_points;
    for (int p = 0; p < 4; ++p) out[k++] = s.player_points[p];
    out[k++] = s.is_solo; out[k++] = s.has_winning_details; out[k++] = s.has_draw_details;
    out[k++] = s.additional.present; out[k++] = s.additional.against_club_queens; out[k++] = s.additional.doko_re; out[k++] = s.additional.doko_kontra;
    out[k++] = s.additional.fuchs_re; out[k++] = s.additional.fuchs_kontra; out[k++] = s.additional.karlchen_re; out[k++] = s.additional.karlchen_kontra;
    for (int i = 0; i < 23; ++i) out[k++] = s.winning_details[i];
    for (int i = 0; i < 14; ++i) out[k++] = s.draw_details[i];
}
// rs-doko pure functions
ORC_API int orc_doko_card_to_color(int c) { return doko::card_to_color_in_normal_game(c); }
ORC_API int orc_doko_is_greater(int cur, int prev, int color) { return doko::is_greater_in_trick_in_normal_game(cur, prev, (doko::Color)color); }
ORC_API uint64_t orc_doko_allowed_actions(int phase, int color, uint64_t hand) { return doko::calculate_allowed_actions_in_normal_game(phase, (doko::Color)color, hand); }
ORC_API int orc_doko_trick_winner(int start, const int32_t cards[4], int32_t* eyes) {
    doko::Trick t; t.present = true; t.start_player = (int8_t)start; for (int k = 0; k < 4; ++k) t.play_card(cards[k]);
    if (eyes) *eyes = (int)t.eyes();
    return t.winner();
}
ORC_API void orc_doko_end_of_game_stats(uint32_t re_players, const uint32_t eyes[4], const uint32_t ntricks[4], int32_t out[10]) {
    doko::EndOfGameStats s = doko::calculate_end_of_game_stats(re_players, eyes, ntricks);
    out[0] = s.winning_team; out[1] = s.is_solo; out[2] = (int)s.re_eyes; out[3] = (int)s.kontra_eyes; out[4] = s.re_points; out[5] = s.kontra_points;
    for (int p = 0; p < 4; ++p) out[6 + p] = s.player_points[p];
}
ORC_API uint64_t orc_doko_hand_op(uint64_t h, int op, int card) { uint64_t r = h; guarded([&] { r = op == 0 ? doko::hand_add(h, card) : doko::hand_remove(h, card); }); return r; }

// ---- determinization -------------------------------------------------------------------------------------------
// card_matching on `h` with Philox unit (unit, sample): returns status; hands[4], reservations[4] (DK_RES_* / 0xFF)
ORC_API int orc_fdo_card_matching_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t sample, uint32_t epoch, uint64_t hands[4], uint8_t res[4]) {
    const fdo::State& s = *(const fdo::State*)h;
    PhiloxStream r(seed, (uint32_t)unit, sample, epoch);
    fdo::Hand oh[4]; int ores[4]; int status = 1;
    guarded([&] { status = fdo::card_matching(s, r, oh, ores); });
    for (int p = 0; p < 4; ++p) { hands[p] = oh[p].bits; res[p] = ores[p] < 0 ? (uint8_t)DK_RES_NONE : (uint8_t)ores[p]; }
    return status;
}
ORC_API int orc_fdo_is_consistent(const void* h, const uint64_t hands[4], const uint8_t res[4]) {
    fdo::Hand ah[4]; int ar[4];
    for (int p = 0; p < 4; ++p) { ah[p].bits = hands[p]; ar[p] = res[p] == DK_RES_NONE ? (int)fdo::R_NONE : (int)res[p]; }
    int rc = -1; guarded([&] { rc = fdo::is_consistent(*(const fdo::State*)h, ah, ar); }); return rc;
}
ORC_API void* orc_fdo_with_hands_and_reservations(const void* h, const uint64_t hands[4], const uint8_t res[4]) {
    fdo::Hand ah[4]; int ar[4];
    for (int p = 0; p < 4; ++p) { ah[p].bits = hands[p]; ar[p] = res[p] == DK_RES_NONE ? (int)fdo::R_NONE : (int)res[p]; }
    return new fdo::State(fdo::with_hands_and_reservations(*(const fdo::State*)h, ah, ar));
}
// random_rollout from a state (no-announcement policy), Philox unit (unit, rollout)
ORC_API void orc_fdo_random_rollout_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t rollout, uint32_t epoch, int with_announcements,
                                           int32_t points[4], uint32_t* steps) {
    fdo::State s = *(const fdo::State*)h;
    PhiloxStream r(seed, (uint32_t)unit, rollout, epoch);
    s.position_streams(r);
    uint32_t before = s.n_play_actions;
    for (;;) { bool fin = with_announcements ? s.random_action_for_current_player(r) : s.random_action_for_current_player_no_announcement(r); if (fin) break; }
    for (int p = 0; p < 4; ++p) points[p] = s.end_of_game_stats.player_points[p];
    if (steps) *steps = s.n_play_actions - before;
}

// rs-doko-assignment sample_assignment_full on `h` with Philox unit (unit, sample): returns status; hands[4]
ORC_API int orc_doko_sample_assignment_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t sample, uint32_t epoch, uint64_t hands[4]) {
    PhiloxStream r(seed, (uint32_t)unit, sample, epoch);
    int status = 1;
    guarded([&] { status = doko::sample_assignment_full(*(const doko::State*)h, r, hands); });
    return status;
}
// are_hands_consistent is not restated; the samples are checked with the properties below instead (tests).

// One leaf rollout: [card_matching → clone_with_different_hands_and_reservations →] random_rollout, all on the Philox unit (unit, rollout).
// Returns the determinization status (0 ok).
ORC_API int orc_fdo_leaf_rollout_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t rollout, uint32_t epoch, int determinize,
                                        int32_t points[4], uint32_t* steps) {
    fdo::State s = *(const fdo::State*)h;
    PhiloxStream r(seed, (uint32_t)unit, rollout, epoch);
    int status = 0;
    if (determinize) {
        fdo::Hand oh[4]; int ores[4];
        status = fdo::card_matching(s, r, oh, ores);
        s = fdo::with_hands_and_reservations(s, oh, ores);
    }
    s.position_streams(r);
    uint32_t before = s.n_play_actions;
    if (status == 0) {
        for (;;) { if (s.random_action_for_current_player_no_announcement(r)) break; }
        for (int p = 0; p < 4; ++p) points[p] = s.end_of_game_stats.player_points[p];
    } else for (int p = 0; p < 4; ++p) points[p] = 0;
    if (steps) *steps = s.n_play_actions - before;
    return status;
}

// ---- bulk playouts (parity checks at scale + CPU baseline) --------------------------------------------------------
// Plays games first_id .. first_id+n-1 from fresh Philox deals.  points[n*4], steps[n] (may be NULL), aux[n*8] (may be NULL):
// aux = [game_type, re_players, re_lowest code, contra_lowest code, eyes0..3];  trace (may be NULL): trace_stride bytes per game, 0xFF padded.
static void fdo_playout_range(uint64_t seed, uint32_t epoch, uint64_t first, uint64_t count, uint64_t base, int with_ann,
                              int32_t* points, uint32_t* steps, int32_t* aux, uint8_t* trace, int trace_stride) {
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t unit = first + i; uint64_t o = unit - base;
        PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
        fdo::State s = fdo::State::new_game(r);
        int k = 0;
        if (trace) std::memset(trace + o * trace_stride, 0xFF, trace_stride);
        for (;;) {
            int a;
            bool fin = with_ann ? s.random_action_for_current_player(r, &a) : s.random_action_for_current_player_no_announcement(r, &a);
            if (fin) break;
            if (trace && k < trace_stride) trace[o * trace_stride + k] = (uint8_t)a;
            k++;
        }
        if (points) for (int p = 0; p < 4; ++p) points[o * 4 + p] = s.end_of_game_stats.player_points[p];
        if (steps) steps[o] = s.n_play_actions;
        if (aux) {
            aux[o * 8 + 0] = s.game_type; aux[o * 8 + 1] = (int)s.team_state.re_players;
            aux[o * 8 + 2] = ann_bit_to_code(s.announcements.re_lowest); aux[o * 8 + 3] = ann_bit_to_code(s.announcements.contra_lowest);
            for (int p = 0; p < 4; ++p) aux[o * 8 + 4 + p] = (int)s.player_eyes[p];
        }
    }
}
static void doko_playout_range(uint64_t seed, uint32_t epoch, uint64_t first, uint64_t count, uint64_t base,
                               int32_t* points, uint32_t* steps, int32_t* aux, uint8_t* trace, int trace_stride) {
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t unit = first + i; uint64_t o = unit - base;
        PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
        doko::State s = doko::State::new_game(r);
        int k = 0;
        if (trace) std::memset(trace + o * trace_stride, 0xFF, trace_stride);
        for (;;) {
            int a; bool fin = s.random_action_for_current_player(r, &a);
            if (fin) break;
            if (trace && k < trace_stride) trace[o * trace_stride + k] = (uint8_t)a;
            k++;
        }
        if (points) for (int p = 0; p < 4; ++p) points[o * 4 + p] = s.end_of_game_stats.player_points[p];
        if (steps) steps[o] = s.n_play_actions;
        if (aux) {
            aux[o * 8 + 0] = s.wedding_player_result >= 0 ? DK_GT_WEDDING : DK_GT_NORMAL; aux[o * 8 + 1] = (int)s.team_state.re_players;
            aux[o * 8 + 2] = 0; aux[o * 8 + 3] = 0;
            for (int p = 0; p < 4; ++p) aux[o * 8 + 4 + p] = (int)s.player_eyes[p];
        }
    }
}
// Returns wall seconds.  n_threads <= 0 → hardware_concurrency.  Static contiguous partition of game ids
// (the shape of the reference's rayon harness, rs-doko-experiments/src/experiment_0_5.rs:97-147).
ORC_API double orc_playout_philox(int engine, int with_announcements, uint64_t seed, uint32_t epoch, uint64_t first_id, uint64_t n, int n_threads,
                                  int32_t* points, uint32_t* steps, int32_t* aux, uint8_t* trace, int trace_stride) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    uint64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t lo = (uint64_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        th.emplace_back([=] {
            if (engine == DK_FDO) fdo_playout_range(seed, epoch, first_id + lo, hi - lo, first_id, with_announcements, points, steps, aux, trace, trace_stride);
            else doko_playout_range(seed, epoch, first_id + lo, hi - lo, first_id, points, steps, aux, trace, trace_stride);
        });
    }
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
ORC_API int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- raw FdoAnnouncements object for the protocol step tests (announcement.rs:228-573) ---------------------------------------
ORC_API void* orc_ann_new() { return new fdo::Announcements(); }
ORC_API void orc_ann_free(void* a) { delete (fdo::Announcements*)a; }
// op 0 = start_round(card_index, player, ...), op 1 = play_announcement(player, ann bit, card_index, ...)
// out = [result kind (0 NextPlayerIs, 1 RoundIsOver), result player, n calls, starting_player, turns, re_lowest bit, contra_lowest bit, allowed set,
//        call[k] = player | card_index<<8 | ann<<16 ...]
ORC_API int orc_ann_step(void* a, int op, int player, int ann, int card_index, const uint32_t lens[4], int tag, int wedding_player, int solved_idx,
                         uint32_t re_players, int32_t* out) {
    fdo::Announcements& A = *(fdo::Announcements*)a;
    fdo::TeamState ts; ts.tag = tag; ts.wedding_player = wedding_player; ts.solved_trick_index = solved_idx; ts.re_players = re_players;
    return guarded([&] {
        fdo::ProgressResult r = op == 0 ? A.start_round(card_index, player, lens, ts) : A.play_announcement(player, ann, card_index, lens, ts);
        out[0] = r.kind == fdo::ROUND_IS_OVER; out[1] = r.player; out[2] = A.n; out[3] = A.starting_player; out[4] = A.turns_without;
        out[5] = A.re_lowest; out[6] = A.contra_lowest; out[7] = (int)A.current_allowed;
        for (int k = 0; k < A.n; ++k) out[8 + k] = A.occ[k].player | (A.occ[k].card_index << 8) | (A.occ[k].announcement << 16);
    });
}

// sample_assignment on raw inputs (the form of the reference's own tests, assignment.rs:908-1186):
// tricks[t*5 + {0..3 cards (-1 = none), 4 start}] for t < n_tricks.
ORC_API int orc_doko_sample_assignment_raw(int marriage, int n_tricks, const int32_t* tricks, uint64_t own_hand, const uint32_t lens[4], int observer,
                                           uint64_t seed, uint64_t unit, uint32_t sample, uint32_t epoch, uint64_t hands[4]) {
    doko::Trick ts[12];
    for (int t = 0; t < n_tricks; ++t) {
        ts[t].present = true; ts[t].start_player = (int8_t)tricks[t * 5 + 4];
        for (int k = 0; k < 4; ++k) if (tricks[t * 5 + k] >= 0) ts[t].play_card(tricks[t * 5 + k]);
    }
    size_t l[4] = {lens[0], lens[1], lens[2], lens[3]};
    PhiloxStream r(seed, (uint32_t)unit, sample, epoch);
    int status = 1;
    guarded([&] { status = doko::sample_assignment(marriage, ts, own_hand, l, observer, r, hands); });
    return status;
}

// ---- PIMC move decision ----------------------------------------------------------------------
ORC_API int orc_fuse_max_n(const uint32_t* visits, uint64_t n_rows, uint64_t allowed) { return pimc::fuse_max_n(visits, (size_t)n_rows, allowed); }
ORC_API int orc_fuse_average(const uint32_t* visits, uint64_t n_rows) { return pimc::fuse_average(visits, (size_t)n_rows); }
ORC_API int orc_fdo_flat_mc_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t det, uint32_t n_rollouts, uint32_t epoch,
                                   uint32_t visits[39], int64_t value_sum[39]) {
    int status = 1;
    guarded([&] { status = pimc::flat_mc(*(const fdo::State*)h, seed, unit, det, n_rollouts, epoch, visits, value_sum); });
    return status;
}

// ---- AlphaZero self-play driver ----------------------------------------------------------------
// One game `unit` dealt from the Philox stream (epoch first_epoch), played by self_play with the uniform stand-in search.
// Returns the number of rows (<= max_rows are written); *turns_out = number of turns.
ORC_API int orc_selfplay_uniform(uint64_t seed, uint64_t unit, uint64_t az_epoch, float keep_prob, uint32_t first_epoch, int max_rows,
                                 int64_t* states_out, float* policy_out, float* value_out, uint8_t* player_out, uint16_t* turn_out,
                                 uint8_t* forced_out, uint32_t* turns_out, int32_t points_out[4]) {
    PhiloxStream deal(seed, (uint32_t)unit, (uint32_t)(unit >> 32), first_epoch);
    fdo::State s = fdo::State::new_game(deal);
    std::vector<selfplay::Row> rows;
    uint32_t turns = 0;
    int rc = guarded([&] {
        turns = selfplay::self_play(s, (size_t)az_epoch, keep_prob, seed, unit, first_epoch,
            [&](const fdo::State& st, uint64_t allowed, uint32_t turn, float* policy) {
                return selfplay::uniform_search(seed, unit, first_epoch, st, allowed, turn, policy); }, rows);
    });
    if (rc) return -1;
    if (turns_out) *turns_out = turns;
    for (size_t i = 0; i < rows.size() && (int)i < max_rows; ++i) {
        std::memcpy(states_out + i * 311, rows[i].state, sizeof rows[i].state);
        std::memcpy(policy_out + i * 39, rows[i].policy, sizeof rows[i].policy);
        std::memcpy(value_out + i * 4, rows[i].value, sizeof rows[i].value);
        player_out[i] = rows[i].player; turn_out[i] = rows[i].turn; forced_out[i] = rows[i].forced;
    }
    if (points_out) {
        // replay to the end to report the final points (value targets already carry them rotated)
        for (int p = 0; p < 4; ++p) points_out[p] = 0;
        if (!rows.empty()) for (int j = 0; j < 4; ++j) points_out[(rows[0].player + j) % 4] = (int32_t)(rows[0].value[j] * 8.0f);
    }
    return (int)rows.size();
}
ORC_API uint64_t orc_fdo_az_allowed(const void* h, int is_secondary, uint64_t epoch) { return selfplay::az_allowed(*(const fdo::State*)h, is_secondary != 0, (size_t)epoch); }

// encode_state_ipi: assumed hands (bitboards by absolute seat), assumed reservations (DK_RES_* / DK_RES_NONE)
ORC_API int orc_fdo_encode_ipi(const void* h, const uint64_t assumed_hands[4], const uint8_t assumed_res[4], int next_player, int64_t out[311]) {
    fdo::Hand ah[4]; int ar[4];
    for (int p = 0; p < 4; ++p) { ah[p].bits = assumed_hands[p]; ar[p] = assumed_res[p] == DK_RES_NONE ? (int)fdo::R_NONE : (int)assumed_res[p]; }
    return guarded([&] { fdo::encode_state_ipi(*(const fdo::State*)h, ah, ar, next_player, out); });
}

// bincode DBRecord bytes (oracle/replay.hpp) for n rows
ORC_API void orc_replay_records(uint64_t n, const int64_t* states, const float* value, const float* policy, uint8_t* out) {
    for (uint64_t r = 0; r < n; ++r) replay::serialize_record(states + r * 311, value + r * 4, policy + r * 39, out + r * replay::RECORD_BYTES);
}

// ---- UCT search --------------------------------------------------------------------------------------
// Tree (unit, sub) over `h` [after card_matching determinization `sub` when determinize != 0].  Returns the determinization status
// (0 ok); *action_out = move with the most visits (-1 when the root has no children).
ORC_API int orc_fdo_uct_search_philox(const void* h, uint64_t seed, uint64_t unit, uint32_t sub, uint32_t epoch, int determinize, uint32_t iterations,
                                      float uct_c, uint32_t visits[39], float values[39], int32_t* action_out) {
    fdo::State s = *(const fdo::State*)h;
    int status = 0;
    for (int a = 0; a < 39; ++a) { visits[a] = 0; values[a] = 0.0f; }
    if (action_out) *action_out = -1;
    int rc = guarded([&] {
        if (determinize && s.current_phase != fdo::PH_FINISHED) {
            PhiloxStream rm(seed, (uint32_t)unit, sub, epoch);
            fdo::Hand oh[4]; int ores[4];
            status = fdo::card_matching(s, rm, oh, ores);
            if (status == 0) s = fdo::with_hands_and_reservations(s, oh, ores);
        }
        if (status == 0) {
            std::vector<mcts::Move> moves = mcts::search(s, (double)uct_c, iterations, seed, unit, sub, epoch);
            int best = mcts::moves_to_arrays(moves, visits, values);
            if (action_out) *action_out = best;
        }
    });
    return rc ? -1 : status;
}
ORC_API uint64_t orc_fdo_mc_allowed(const void* h, int first_expansion) { return mcts::allowed_actions(*(const fdo::State*)h, first_expansion != 0); }

// ---- CPU baselines of the other configs: the oracle timed on host threads, static partition of the unit ids ------------
// kind: 3 = card_matching samples (config 3, rs-full-doko), 30 = sample_assignment (config 3, rs-doko), 5 = lock-step env step + encode_state_pi
//       (config 5), 4 = determinized leaf rollouts (config 4), 6 = UCT iterations (N3), 7 = flat Monte-Carlo PIMC rollouts (N2).
// Units are mid-game info-states made like BASELINE config 3 (games advanced by random steps to card_index 8/16/24/32 round-robin).
// Returns wall seconds; *work_out = number of work items done (samples / step-encodes / rollouts / iterations).
ORC_API double orc_cpu_baseline(int kind, uint64_t seed, uint64_t n_units, uint32_t per_unit, int n_threads, uint64_t* work_out) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    std::vector<uint64_t> work((size_t)n_threads, 0);
    auto make_fdo = [&](uint64_t unit) {
        PhiloxStream r(seed, (uint32_t)unit, 0, 0);
        fdo::State s = fdo::State::new_game(r);
        const int target = 8 * (1 + (int)(unit & 3));
        while (s.current_phase != fdo::PH_FINISHED && s.card_index < target) s.random_action_for_current_player(r);
        return s;
    };
    auto make_doko = [&](uint64_t unit) {
        PhiloxStream r(seed, (uint32_t)unit, 0, 0);
        doko::State s = doko::State::new_game(r);
        const int target = 8 * (1 + (int)(unit & 3)) + 4;
        for (int k = 0; k < target; ++k) if (s.random_action_for_current_player(r)) break;
        return s;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    uint64_t per = (n_units + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t lo = (uint64_t)t * per, hi = lo + per > n_units ? n_units : lo + per;
        if (lo >= hi) break;
        th.emplace_back([&, t, lo, hi] {
            uint64_t done = 0;
            volatile int64_t sink = 0;
            for (uint64_t unit = lo; unit < hi; ++unit) {
                try {
                    if (kind == 30) {
                        doko::State s = make_doko(unit);
                        for (uint32_t k = 0; k < per_unit; ++k) {
                            PhiloxStream r(seed, (uint32_t)unit, k, 1); uint64_t hands[4];
                            sink += doko::sample_assignment_full(s, r, hands) + (int64_t)(hands[0] & 1); done++;
                        }
                        continue;
                    }
                    fdo::State s = make_fdo(unit);
                    if (s.current_phase == fdo::PH_FINISHED) continue;
                    if (kind == 3) {
                        for (uint32_t k = 0; k < per_unit; ++k) {
                            PhiloxStream r(seed, (uint32_t)unit, k, 1); fdo::Hand oh[4]; int ores[4];
                            sink += fdo::card_matching(s, r, oh, ores) + (int64_t)(oh[0].bits & 1); done++;
                        }
                    } else if (kind == 4) {
                        for (uint32_t k = 0; k < per_unit; ++k) {
                            PhiloxStream r(seed, (uint32_t)unit, k, 1); fdo::Hand oh[4]; int ores[4];
                            if (fdo::card_matching(s, r, oh, ores)) continue;
                            fdo::State d = fdo::with_hands_and_reservations(s, oh, ores);
                            d.position_streams(r);
                            for (;;) { if (d.random_action_for_current_player_no_announcement(r)) break; }
                            sink += d.end_of_game_stats.player_points[0]; done++;
                        }
                    } else if (kind == 5) {
                        int64_t obs[311];
                        for (uint32_t k = 0; k < per_unit && s.current_phase != fdo::PH_FINISHED; ++k) {
                            PhiloxStream r(seed, (uint32_t)unit, 0, 100 + k);
                            uint64_t allowed = s.allowed_actions();
                            uint64_t bit = random_single(allowed, r, SITE_STEP);
                            s.play_action(__builtin_ctzll(bit));
                            fdo::encode_state_pi(s, obs); sink += obs[0]; done++;
                        }
                    } else if (kind == 6) {
                        std::vector<mcts::Move> m = mcts::search(s, 1.4, per_unit, seed, unit, 0, 1);
                        sink += (int64_t)m.size(); done += per_unit;
                    } else if (kind == 7) {
                        uint32_t visits[39]; int64_t values[39];
                        if (pimc::flat_mc(s, seed, unit, 0, per_unit, 1, visits, values) == 0)
                            done += (uint64_t)per_unit * (uint64_t)__builtin_popcountll(s.allowed_actions());
                    }
                } catch (const std::exception&) {}
            }
            work[(size_t)t] = done;
        });
    }
    for (auto& x : th) x.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t total = 0;
    for (uint64_t w : work) total += w;
    if (work_out) *work_out = total;
    return sec;
}

// ---- bulk checkers for the full-size GPU parity tests (tests/test_gpu_parity_at_size.py) -------------------------------------------
// A batch of oracle states made by a recipe, exported as dk_state records for the GPU and kept as oracle objects for the checks:
// game first_id + i is dealt from the Philox stream (unit, epoch) and advanced with the random policy (announcements included).
//   mode 0  BASELINE config 3: until card_index reaches 8 / 16 / 24 / 32 (round-robin over i)
//   mode 1  the reference's determinization soak (rs-full-doko-cmd/src/main.rs:190-283 tests card_matching at EVERY state of a game):
//           (i * 2654435761 >> 16) % 96 actions, never the last card — reservation-phase, announcement-phase and card-phase states
//   mode 2  late-game states for the support-set test: until card_index reaches 34 + i % 13 (<= 10 hidden cards)
// rs-doko (engine 0): 4 reservations + the same number of cards.
namespace {
struct BulkStates { int engine = DK_FDO; std::vector<fdo::State> f; std::vector<doko::State> d; };
template <class F> void parallel_for(uint64_t n, int n_threads, F&& body) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> th;
    uint64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t lo = (uint64_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        th.emplace_back([=, &body] { for (uint64_t i = lo; i < hi; ++i) body(i); });
    }
    for (auto& x : th) x.join();
}
}  // namespace
ORC_API void* orc_bulk_make(int engine, uint64_t n, uint64_t seed, uint64_t first_id, uint32_t epoch, int mode, dk_state* recs_out, int n_threads) {
    BulkStates* b = new BulkStates();
    b->engine = engine;
    if (engine == DK_FDO) b->f.resize((size_t)n); else b->d.resize((size_t)n);
    parallel_for(n, n_threads, [&](uint64_t i) {
        const uint64_t unit = first_id + i;
        PhiloxStream r(seed, (uint32_t)unit, (uint32_t)(unit >> 32), epoch);
        const int quota = (int)((((uint32_t)i * 2654435761u) >> 16) % 96u);
        const int target = mode == 2 ? 34 + (int)(i % 13) : 8 * (1 + (int)(i & 3));
        if (engine == DK_FDO) {
            fdo::State s = fdo::State::new_game(r);
            if (mode == 0 || mode == 2) { while (s.current_phase != fdo::PH_FINISHED && s.card_index < target) s.random_action_for_current_player(r); }
            else for (int k = 0; k < quota; ++k) {
                if (s.current_phase == fdo::PH_PLAYCARD && s.card_index == 47) break;
                if (s.random_action_for_current_player(r)) break;
            }
            b->f[(size_t)i] = s;
            if (recs_out) export_fdo(s, recs_out + i);
        } else {
            doko::State s = doko::State::new_game(r);
            const int steps = mode == 0 ? 4 + target : (quota < 51 ? quota : 51);
            for (int k = 0; k < steps; ++k) if (s.random_action_for_current_player(r)) break;
            b->d[(size_t)i] = s;
            if (recs_out) export_doko(s, recs_out + i);
        }
    });
    return b;
}
ORC_API void orc_bulk_free(void* h) { delete (BulkStates*)h; }
// card_matching (engine 1) / sample_assignment_full (engine 0): S samples per state on the stream (first_id + i, first_sub + s, epoch).
// hands_out[(i*S+s)*4], res_out[(i*S+s)*4] (DK_RES_* / 0xFF; engine 1 only), status_out[i*S+s], consistent_out[i*S+s] (engine 1, nullable:
// is_consistent's reason code, 0 = consistent, -1 = not checked because the sample failed).
ORC_API double orc_bulk_determinize(const void* h, uint32_t S, uint64_t seed, uint64_t first_id, uint32_t epoch, uint32_t first_sub, uint64_t* hands_out,
                                    uint8_t* res_out, uint8_t* status_out, int8_t* consistent_out, int n_threads) {
    const BulkStates& b = *(const BulkStates*)h;
    const uint64_t n = b.engine == DK_FDO ? b.f.size() : b.d.size();
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, n_threads, [&](uint64_t i) {
        const uint64_t unit = first_id + i;
        for (uint32_t s = 0; s < S; ++s) {
            const uint64_t o = i * S + s;
            PhiloxStream r(seed, (uint32_t)unit, first_sub + s, epoch);
            if (b.engine == DK_FDO) {
                fdo::Hand oh[4]; int ores[4] = {-1, -1, -1, -1}; int status = 2;
                const fdo::State& st = b.f[(size_t)i];
                if (st.current_phase != fdo::PH_FINISHED) { try { status = fdo::card_matching(st, r, oh, ores); } catch (const std::exception&) { status = 9; } }
                for (int p = 0; p < 4; ++p) { hands_out[o * 4 + p] = status == 2 ? 0 : oh[p].bits; if (res_out) res_out[o * 4 + p] = (status == 2 || ores[p] < 0) ? (uint8_t)DK_RES_NONE : (uint8_t)ores[p]; }
                status_out[o] = (uint8_t)status;
                if (consistent_out) {
                    int rc = -1;
                    if (status == 0) { int ar[4]; for (int p = 0; p < 4; ++p) ar[p] = ores[p] < 0 ? (int)fdo::R_NONE : ores[p]; try { rc = fdo::is_consistent(st, oh, ar); } catch (const std::exception&) { rc = 99; } }
                    consistent_out[o] = (int8_t)rc;
                }
            } else {
                uint64_t hands[4] = {0, 0, 0, 0}; int status = 9;
                try { status = doko::sample_assignment_full(b.d[(size_t)i], r, hands); } catch (const std::exception&) {}
                for (int p = 0; p < 4; ++p) hands_out[o * 4 + p] = hands[p];
                status_out[o] = (uint8_t)status;
            }
        }
    });
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// [card_matching →] random_rollout (no-announcement policy) R times per state on the stream (first_id + i, first_sub + r, epoch):
// sums_out[i*4 + seat] = exact sum of player_points (failed samples add 0), the quantity dk_leaf_rollouts returns.
ORC_API double orc_bulk_leaf_rollouts(const void* h, uint32_t R, uint64_t seed, uint64_t first_id, uint32_t epoch, uint32_t first_sub, int determinize,
                                      int64_t* sums_out, int n_threads) {
    const BulkStates& b = *(const BulkStates*)h;
    const uint64_t n = b.f.size();
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, n_threads, [&](uint64_t i) {
        int64_t acc[4] = {0, 0, 0, 0};
        for (uint32_t k = 0; k < R; ++k) {
            int32_t pts[4]; 
            if (orc_fdo_leaf_rollout_philox(&b.f[(size_t)i], seed, first_id + i, first_sub + k, epoch, determinize, pts, nullptr) == 0)
                for (int p = 0; p < 4; ++p) acc[p] += pts[p];
        }
        for (int p = 0; p < 4; ++p) sums_out[i * 4 + p] = acc[p];
    });
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// Lock-step env step (dk_step_random_encode) on every state of the batch: returns the actions; states advance in place.
ORC_API void orc_bulk_step(void* h, uint64_t seed, uint64_t first_id, uint32_t epoch, int with_announcements, int skip_single, uint8_t* action_out,
                           dk_state* recs_out, int64_t* obs_out, int n_threads) {
    BulkStates& b = *(BulkStates*)h;
    parallel_for(b.f.size(), n_threads, [&](uint64_t i) {
        fdo::State& s = b.f[(size_t)i];
        int a = 0xFF;
        if (s.current_phase != fdo::PH_FINISHED) a = orc_fdo_step_site(&s, seed, first_id + i, epoch, with_announcements, skip_single);
        if (action_out) action_out[i] = (uint8_t)a;
        if (recs_out) export_fdo(s, recs_out + i);
        if (obs_out) fdo::encode_state_pi(s, obs_out + i * 311);
    });
}

// ---- support set of the determinizer (tests/test_card_matching_support.py) -------------------------------------------------------
// Every assignment of the hidden cards to the three hidden seats (hand sizes kept) for which SOME assignment of the hidden
// reservations passes is_consistent — the full support a consistent sampler may draw from; card_matching's greedy rules reach a
// subset of it.  Late-game states only (the number of assignments is a multinomial).  out_hands[k*4 + seat]; returns the number of
// members, or -1 when it exceeds `max`.
ORC_API int64_t orc_fdo_enumerate_consistent_hands(const void* h, int64_t max, uint64_t* out_hands) {
    const fdo::State& s = *(const fdo::State*)h;
    const int obs = s.observing_player();
    int count[24] = {0};
    for (int p = 0; p < 4; ++p) {
        if (p == obs) continue;
        for (int c = 0; c < 24; ++c) count[c] += (int)((s.hands[p].bits >> c) & 1) + (int)((s.hands[p].bits >> (c + 24)) & 1);
    }
    int seats[3], cap[3], k = 0;
    for (int p = 0; p < 4; ++p) if (p != obs) { seats[k] = p; cap[k] = (int)s.hands[p].len(); k++; }
    int visible[4]; fdo::get_visible_reservations(s.reservations_round, obs, visible);
    int have[3][24] = {{0}};
    int64_t n_out = 0;
    bool overflow = false;
    auto emit = [&]() {
        fdo::Hand hands[4];
        hands[obs] = s.hands[obs];
        for (int j = 0; j < 3; ++j) {
            uint64_t bits = 0;
            for (int c = 0; c < 24; ++c) { if (have[j][c] >= 1) bits |= 1ull << c; if (have[j][c] == 2) bits |= 1ull << (c + 24); }
            hands[seats[j]].bits = bits;
        }
        // hidden reservations: a NotRevealed seat declared a solo or a wedding — try both kinds
        int opts[4][2], nopt[4];
        for (int p = 0; p < 4; ++p) {
            const int vr = visible[p];
            if (vr == fdo::VR_NONE_YET) { opts[p][0] = fdo::R_NONE; nopt[p] = 1; }
            else if (vr == fdo::VR_NOT_REVEALED) { opts[p][0] = fdo::R_DIAMONDS_SOLO; opts[p][1] = fdo::R_WEDDING; nopt[p] = 2; }
            else if (vr == fdo::VR_HEALTHY) { opts[p][0] = fdo::R_HEALTHY; nopt[p] = 1; }
            else if (vr == fdo::VR_WEDDING) { opts[p][0] = fdo::R_WEDDING; nopt[p] = 1; }
            else { opts[p][0] = vr - fdo::VR_DIAMONDS_SOLO + fdo::R_DIAMONDS_SOLO; nopt[p] = 1; }
        }
        bool ok = false;
        for (int a = 0; a < nopt[0] && !ok; ++a) for (int b = 0; b < nopt[1] && !ok; ++b) for (int c = 0; c < nopt[2] && !ok; ++c) for (int d = 0; d < nopt[3] && !ok; ++d) {
            int res[4] = {opts[0][a], opts[1][b], opts[2][c], opts[3][d]};
            ok = fdo::is_consistent(s, hands, res) == 0;
        }
        if (!ok) return;
        if (n_out >= max) { overflow = true; return; }
        for (int p = 0; p < 4; ++p) out_hands[n_out * 4 + p] = hands[p].bits;
        n_out++;
    };
    std::function<void(int)> rec = [&](int c) {
        if (overflow) return;
        while (c < 24 && count[c] == 0) c++;
        if (c == 24) { if (cap[0] == 0 && cap[1] == 0 && cap[2] == 0) emit(); return; }
        const int n = count[c];
        for (int a = 0; a <= n; ++a) for (int b = 0; a + b <= n; ++b) {
            const int d = n - a - b;
            if (a > cap[0] || b > cap[1] || d > cap[2]) continue;
            have[0][c] = a; have[1][c] = b; have[2][c] = d; cap[0] -= a; cap[1] -= b; cap[2] -= d;
            rec(c + 1);
            cap[0] += a; cap[1] += b; cap[2] += d; have[0][c] = have[1][c] = have[2][c] = 0;
        }
    };
    rec(0);
    return overflow ? -1 : n_out;
}
