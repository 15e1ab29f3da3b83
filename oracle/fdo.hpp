// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the reference's FULL-RULES Doppelkopf engine `rs-full-doko`.
// Every function cites the reference file:line it follows.  The data model deliberately mirrors
// the reference's (trick list, announcement occurrence list, 48-bit hand bitboards) and NOT the
// packed/bit-sliced layout of the CUDA kernels, so that the two are independent implementations.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include "bitflag.hpp"
#include "rng.hpp"

namespace oracle {
namespace fdo {

// ---- basic enums --------------------------------------------------------------------------
// FdoCard: one-hot 1<<i, i = suit*6 + rank; suits ♦0 ♥1 ♣2 ♠3; ranks 9,10,J,Q,K,A
// (rs-full-doko/src/card/cards.rs:7-36).  Here a card is its index i (0..23).
enum Card : int {
    D9, D10, DJ, DQ, DK, DA, H9, H10, HJ, HQ, HK, HA, C9, C10, CJ, CQ, CK, CA, S9, S10, SJ, SQ, SK, SA,
    CARD_COUNT = 24, CARD_NONE = -1
};
// FdoColor (basic/color.rs:4-13)
enum Color : int { TRUMP = 0, DIAMOND = 1, HEART = 2, SPADE = 3, CLUB = 4, COLOR_NONE = -1 };
// FdoPhase (basic/phase.rs:6-11)
enum Phase : int { PH_RESERVATION = 0, PH_ANNOUNCEMENT = 1, PH_PLAYCARD = 2, PH_FINISHED = 3 };
// FdoGameType (game_type/game_type.rs:6-20)
enum GameType : int {
    GT_NORMAL = 0, GT_WEDDING, GT_DIAMONDS_SOLO, GT_HEARTS_SOLO, GT_SPADES_SOLO, GT_CLUBS_SOLO,
    GT_TRUMPLESS_SOLO, GT_QUEENS_SOLO, GT_JACKS_SOLO, GT_NONE = -1
};
// FdoReservation (reservation/reservation.rs:11-24) — NOTE order differs from the action order.
enum Reservation : int {
    R_HEALTHY = 0, R_WEDDING, R_DIAMONDS_SOLO, R_HEARTS_SOLO, R_SPADES_SOLO, R_CLUBS_SOLO,
    R_QUEENS_SOLO, R_JACKS_SOLO, R_TRUMPLESS_SOLO, R_NONE = -1
};
// FdoVisibleReservation (reservation/reservation.rs:66-82)
enum VisibleReservation : int {
    VR_WEDDING = 0, VR_HEALTHY = 1, VR_NOT_REVEALED = 2, VR_DIAMONDS_SOLO = 3, VR_HEARTS_SOLO = 4,
    VR_SPADES_SOLO = 5, VR_CLUBS_SOLO = 6, VR_QUEENS_SOLO = 7, VR_JACKS_SOLO = 8,
    VR_TRUMPLESS_SOLO = 9, VR_NONE_YET = 10
};
// FdoAnnouncement (announcement/announcement.rs:12-22): bit values; 0 = Option::None.
enum Announcement : int {
    A_NONE = 0, A_RE_CONTRA = 1, A_NO90 = 2, A_NO60 = 4, A_NO30 = 8, A_BLACK = 16,
    A_COUNTER_RE_CONTRA = 32, A_NO_ANNOUNCEMENT = 64
};
// FdoAction indices (action/action.rs:9-54,57-155): bit i of the 39-bit legal mask.
enum ActionIndex : int {
    ACT_CARD0 = 0, ACT_RES_HEALTHY = 24, ACT_RES_WEDDING = 25, ACT_RES_DIAMONDS = 26, ACT_RES_HEARTS = 27,
    ACT_RES_SPADES = 28, ACT_RES_CLUBS = 29, ACT_RES_TRUMPLESS = 30, ACT_RES_QUEENS = 31, ACT_RES_JACKS = 32,
    ACT_ANN_RE_CONTRA = 33, ACT_ANN_NO90 = 34, ACT_ANN_NO60 = 35, ACT_ANN_NO30 = 36, ACT_ANN_BLACK = 37,
    ACT_NO_ANNOUNCEMENT = 38, ACT_COUNT = 39
};
// team state tag (team/team_logic.rs:11-28)
enum TeamTag : int { TS_IN_RESERVATIONS = 0, TS_WEDDING_UNSOLVED = 1, TS_WEDDING_SOLVED = 2, TS_NO_WEDDING = 3 };
// reservation result kind (reservation/reservation_winning_logic.rs:9-13)
enum ResResultKind : int { RR_NONE_YET = -1, RR_NO_RESERVATION = 0, RR_SOLO = 1, RR_WEDDING = 2 };

constexpr uint64_t SECOND_CARD_SHIFT_OFFSET = 24;  // hand/hand.rs:107
inline uint64_t card_bit(int c) { return 1ull << c; }
inline int player_next(int p, int i) { return (p + i) % 4; }  // player/player.rs:67-72

// ---- eyes: card/card_to_eyes.rs:8-40 ----------------------------------------------------------
inline uint32_t card_eyes(int c) {
    static const uint32_t by_rank[6] = {0, 10, 2, 3, 4, 11};  // 9,10,J,Q,K,A
    return by_rank[c % 6];
}

// ---- card_to_color: card/card_to_color.rs:11-258 (one 24-entry table per game type) ----------
inline Color card_to_color(int card, int game_type) {
    // T trump, D ♦, H ♥, S ♠, C ♣; index = card id (♦9..♦A, ♥9..♥A, ♣9..♣A, ♠9..♠A)
    static const char* normal    = "TTTTTT" "HTTTHH" "CCTTCC" "SSTTSS";  // :16-47 (Normal, Wedding, ♦-Solo)
    static const char* hearts    = "DDTTDD" "TTTTTT" "CCTTCC" "SSTTSS";  // :50-81
    static const char* spades    = "DDTTDD" "HTTTHH" "CCTTCC" "TTTTTT";  // :84-115
    static const char* clubs     = "DDTTDD" "HTTTHH" "TTTTTT" "SSTTSS";  // :118-149
    static const char* trumpless = "DDDDDD" "HHHHHH" "CCCCCC" "SSSSSS";  // :151-182
    static const char* queens    = "DDDTDD" "HHHTHH" "CCCTCC" "SSSTSS";  // :184-215
    static const char* jacks     = "DDTDDD" "HHTHHH" "CCTCCC" "SSTSSS";  // :217-248
    const char* t;
    switch (game_type) {
        case GT_NORMAL: case GT_WEDDING: case GT_DIAMONDS_SOLO: t = normal; break;
        case GT_HEARTS_SOLO: t = hearts; break;
        case GT_SPADES_SOLO: t = spades; break;
        case GT_CLUBS_SOLO: t = clubs; break;
        case GT_TRUMPLESS_SOLO: t = trumpless; break;
        case GT_QUEENS_SOLO: t = queens; break;
        case GT_JACKS_SOLO: t = jacks; break;
        default: throw std::runtime_error("card_to_color: bad game type");
    }
    switch (t[card]) {
        case 'T': return TRUMP; case 'D': return DIAMOND; case 'H': return HEART;
        case 'S': return SPADE; default: return CLUB;
    }
}

// ---- get_color_masks_for_game_type: card/card_color_masks.rs:7-246 ----------------------------
// Returns (trump, diamond, heart, spade, club) 24-bit masks, written out per game type as in the
// reference (NOT derived from card_to_color; the test-suite checks both agree).
struct ColorMasks { uint64_t m[5]; };
inline uint64_t mk(std::initializer_list<int> cs) { uint64_t r = 0; for (int c : cs) r |= card_bit(c); return r; }
inline ColorMasks build_color_masks_for_game_type(int game_type) {
    const uint64_t jacks = mk({DJ, HJ, SJ, CJ}), queens = mk({DQ, HQ, SQ, CQ});
    const uint64_t d4 = mk({D9, DK, D10, DA}), h3 = mk({H9, HK, HA}), s4 = mk({S9, SK, S10, SA}), c4 = mk({C9, CK, C10, CA});
    ColorMasks r;
    switch (game_type) {
        case GT_NORMAL: case GT_WEDDING: case GT_DIAMONDS_SOLO:        // :11-43
            r.m[TRUMP] = d4 | jacks | queens | card_bit(H10);
            r.m[DIAMOND] = 0; r.m[HEART] = h3; r.m[SPADE] = s4; r.m[CLUB] = c4; break;
        case GT_HEARTS_SOLO:                                           // :44-77
            r.m[TRUMP] = h3 | jacks | queens | card_bit(H10);
            r.m[DIAMOND] = d4; r.m[HEART] = 0; r.m[SPADE] = s4; r.m[CLUB] = c4; break;
        case GT_SPADES_SOLO:                                           // :78-111
            r.m[TRUMP] = s4 | jacks | queens | card_bit(H10);
            r.m[DIAMOND] = d4; r.m[HEART] = h3; r.m[SPADE] = 0; r.m[CLUB] = c4; break;
        case GT_CLUBS_SOLO:                                            // :112-145
            r.m[TRUMP] = c4 | jacks | queens | card_bit(H10);
            r.m[DIAMOND] = d4; r.m[HEART] = h3; r.m[SPADE] = s4; r.m[CLUB] = 0; break;
        case GT_TRUMPLESS_SOLO:                                        // :146-178
            r.m[TRUMP] = 0;
            r.m[DIAMOND] = mk({D9, DJ, DQ, DK, D10, DA}); r.m[HEART] = mk({H9, HJ, HQ, HK, H10, HA});
            r.m[SPADE] = mk({S9, SJ, SQ, SK, S10, SA}); r.m[CLUB] = mk({C9, CJ, CQ, CK, C10, CA}); break;
        case GT_QUEENS_SOLO:                                           // :179-211
            r.m[TRUMP] = queens;
            r.m[DIAMOND] = mk({D9, DJ, DK, D10, DA}); r.m[HEART] = mk({H9, HJ, HK, H10, HA});
            r.m[SPADE] = mk({S9, SJ, SK, S10, SA}); r.m[CLUB] = mk({C9, CJ, CK, C10, CA}); break;
        case GT_JACKS_SOLO:                                            // :212-243
            r.m[TRUMP] = jacks;
            r.m[DIAMOND] = mk({D9, DQ, DK, D10, DA}); r.m[HEART] = mk({H9, HQ, HK, H10, HA});
            r.m[SPADE] = mk({S9, SQ, SK, S10, SA}); r.m[CLUB] = mk({C9, CK, C10, CA, CQ}); break;
        default: throw std::runtime_error("color masks: bad game type");
    }
    return r;
}

// (the written-out masks above, evaluated once per game type: the reference's `match` returns constants)
inline const ColorMasks& get_color_masks_for_game_type(int game_type) {
    static const struct Table { ColorMasks t[GT_JACKS_SOLO + 1]; Table() { for (int g = GT_NORMAL; g <= GT_JACKS_SOLO; ++g) t[g] = build_color_masks_for_game_type(g); } } tab;
    if (game_type < GT_NORMAL || game_type > GT_JACKS_SOLO) throw std::runtime_error("color masks: bad game type");
    return tab.t[game_type];
}

// ---- is_greater_in_trick: card/card_in_trick_logic.rs:17-141 ----------------------------------
inline int trump_to_rank(int card) {  // :37-77 (one table for every game type)
    switch (card) {
        case D9: return 0; case DK: return 1; case D10: return 2; case DA: return 3;
        case H9: return 0; case HK: return 1; case HA: return 2;
        case S9: return 0; case SK: return 1; case S10: return 2; case SA: return 3;
        case C9: return 0; case CK: return 1; case C10: return 2; case CA: return 3;
        case DJ: return 4; case HJ: return 5; case SJ: return 6; case CJ: return 7;
        case DQ: return 8; case HQ: return 9; case SQ: return 10; case CQ: return 11;
        case H10: return 12;
    }
    throw std::runtime_error("trump_to_rank");
}
inline bool is_greater_in_trick(int current, int previous, Color trick_color, int game_type) {
    Color cc = card_to_color(current, game_type), pc = card_to_color(previous, game_type);
    bool ct = cc == TRUMP, pt = pc == TRUMP;
    if (ct && !pt) return true;                                           // :102-104
    if (!ct && pt) return false;                                          // :107-109
    if (ct && pt) return trump_to_rank(previous) < trump_to_rank(current);  // :113-115, :78-81 strict
    bool c_tc = cc == trick_color, p_tc = pc == trick_color;
    if (c_tc && !p_tc) return true;                                       // :124-126
    if (c_tc && p_tc) return card_eyes(current) > card_eyes(previous);    // :129-135 strict
    return false;
}

// ---- hand: hand/hand.rs:20-272, hand/hand_iter.rs:13-49 ----------------------------------------
struct Hand {
    uint64_t bits = 0;
    bool contains(int c) const { return (bits & card_bit(c)) || (bits & (card_bit(c) << 24)); }     // :207-209
    bool contains_both(int c) const { return (bits & card_bit(c)) && (bits & (card_bit(c) << 24)); }  // :212-214
    void add(int c) {                                                     // :217-230 (A first, then B)
        if (bits & card_bit(c)) bits |= card_bit(c) << 24; else bits |= card_bit(c);
    }
    void add_ignore(int c) { if (!contains_both(c)) add(c); }             // :232-236
    void remove(int c) {                                                  // :239-247 (B first, then A)
        if (bits & (card_bit(c) << 24)) bits &= ~(card_bit(c) << 24);
        else if (bits & card_bit(c)) bits &= ~card_bit(c);
        else throw std::runtime_error("Karte nicht in Hand");
    }
    void remove_ignore(int c) { if (contains(c)) remove(c); }             // :250-254
    void remove_both(int c) {                                             // :256-266
        if (contains(c)) { remove(c); if (contains(c)) remove(c); }
        else throw std::runtime_error("Karte nicht in Hand");
    }
    uint32_t len() const { return popcount64(bits); }                     // :269-271
    bool contains_card_of_color(Color color, int gt) const {              // :24-42
        uint64_t m = get_color_masks_for_game_type(gt).m[color];
        return (bits & m) != 0 || (bits & (m << 24)) != 0;
    }
    void remove_color(Color color, int gt) {                              // :45-62
        uint64_t m = get_color_masks_for_game_type(gt).m[color];
        bits &= ~(m | (m << 24));
    }
    Hand plus_hand(Hand other) const {                                    // :69-83
        Hand n = *this;
        for (int c = 0; c < 24; ++c) {
            if (other.contains_both(c)) { n.add(c); n.add(c); }
            else if (other.contains(c)) n.add(c);
        }
        return n;
    }
    Hand minus_hand(Hand other) const {                                   // :86-102
        Hand n = *this;
        for (int c = 0; c < 24; ++c) {
            if (other.contains_both(c)) n.remove_both(c);
            else if (other.contains(c)) n.remove(c);
        }
        return n;
    }
    // FdoHandIter (hand_iter.rs:16-37): ascending bit position over all 48 bits; both copies map
    // to the same card.  Returns the number of yielded cards.
    int iter(int* out) const {
        int n = 0; uint64_t b = bits;
        while (b) { int pos = __builtin_ctzll(b); b &= b - 1; out[n++] = pos < 24 ? pos : pos - 24; }
        return n;
    }
};

// AVAILABLE_CARDS (hand/hand.rs:116-141): [c0,c0,c1,c1,...,c23,c23]
inline void available_cards(uint8_t* a) { for (int i = 0; i < 48; ++i) a[i] = (uint8_t)(i / 2); }

// FdoHand::randomly_distributed (hand/hand.rs:174-188)
inline void randomly_distributed(Rng& rng, Hand hands[4]) {
    uint8_t cards[48];
    available_cards(cards);
    rng.shuffle48(cards);
    for (int p = 0; p < 4; ++p) { hands[p] = Hand(); for (int j = 0; j < 12; ++j) hands[p].add(cards[p * 12 + j]); }
}

// ---- trick: trick/trick.rs:11-126, trick/trick_winning_player_logic.rs:15-45 -------------------
struct Trick {
    int8_t cards[4] = {-1, -1, -1, -1};
    int8_t len = 0;
    int8_t starting_player = 0;
    int8_t winning_player = -1;
    int8_t winning_card = -1;
    static Trick empty(int start) { Trick t; t.starting_player = (int8_t)start; return t; }
    bool is_completed() const { return len == 4; }
    Color color(int gt) const { return len == 0 ? COLOR_NONE : card_to_color(cards[0], gt); }  // :31-37
    int player_at(int i) const { return player_next(starting_player, i); }
    uint32_t eyes() const { uint32_t e = 0; for (int i = 0; i < len; ++i) e += card_eyes(cards[i]); return e; }  // :74-82
    void calc_winner(int gt, int& wp, int& wc) const {                   // trick_winning_player_logic.rs:15-45
        Color tc = color(gt);
        wc = -1; wp = -1;
        for (int i = 0; i < len; ++i) {
            if (wc < 0) { wc = cards[i]; wp = player_at(i); continue; }
            if (is_greater_in_trick(cards[i], wc, tc, gt)) { wc = cards[i]; wp = player_at(i); }
        }
    }
    void play_card(int card, int gt) {                                    // :84-104
        cards[len++] = (int8_t)card;
        if (is_completed()) { int wp, wc; calc_winner(gt, wp, wc); winning_player = (int8_t)wp; winning_card = (int8_t)wc; }
    }
    static Trick existing(int start, std::initializer_list<int> cs) {    // :55-68 (game type Normal)
        Trick t = empty(start); for (int c : cs) t.play_card(c, GT_NORMAL); return t;
    }
};

// ---- reservations: reservation/*.rs -------------------------------------------------------------
struct ReservationRound {             // reservation_round.rs:10-68 (PlayerOrientedVec<FdoReservation>)
    int8_t r[4] = {-1, -1, -1, -1};   // in play order from starting_player
    int8_t len = 0;
    int8_t starting_player = 0;
    bool is_completed() const { return len == 4; }
    void play_reservation(int res) { r[len++] = (int8_t)res; }
    int get(int player) const {       // Index<FdoPlayer>: po_vec.rs:96-108 ; -1 when not yet played
        int i = ((4 - starting_player) + player) % 4;
        return i < len ? r[i] : -1;
    }
};
struct ReservationResult { int kind = RR_NONE_YET; int player = -1; int reservation = -1; };

// winning_player_in_reservation_round (reservation_winning_logic.rs:36-73): FIRST solo in seat
// order from the start player wins immediately; else the LAST wedding; else none.
inline ReservationResult winning_player_in_reservation_round(const ReservationRound& rr) {
    int wedding_player = -1;
    for (int i = 0; i < rr.len; ++i) {
        int player = player_next(rr.starting_player, i);
        int res = rr.r[i];
        if (res == R_WEDDING) wedding_player = player;
        else if (res == R_HEALTHY) {}
        else { ReservationResult x; x.kind = RR_SOLO; x.player = player; x.reservation = res; return x; }
    }
    ReservationResult x;
    if (wedding_player < 0) x.kind = RR_NO_RESERVATION; else { x.kind = RR_WEDDING; x.player = wedding_player; }
    return x;
}
inline int to_game_type(const ReservationResult& rr) {                   // :16-32
    switch (rr.kind) {
        case RR_NO_RESERVATION: return GT_NORMAL;
        case RR_WEDDING: return GT_WEDDING;
        case RR_SOLO:
            switch (rr.reservation) {
                case R_DIAMONDS_SOLO: return GT_DIAMONDS_SOLO; case R_HEARTS_SOLO: return GT_HEARTS_SOLO;
                case R_SPADES_SOLO: return GT_SPADES_SOLO; case R_CLUBS_SOLO: return GT_CLUBS_SOLO;
                case R_QUEENS_SOLO: return GT_QUEENS_SOLO; case R_JACKS_SOLO: return GT_JACKS_SOLO;
                case R_TRUMPLESS_SOLO: return GT_TRUMPLESS_SOLO;
            }
    }
    throw std::runtime_error("to_game_type");
}
// get_visible_reservations (visible_reservations_logic.rs:7-70); out[] indexed by ABSOLUTE seat.
inline void get_visible_reservations(const ReservationRound& rr, int observing_player, int out[4]) {
    for (int p = 0; p < 4; ++p) out[p] = VR_NONE_YET;
    bool completed = rr.is_completed();
    bool higher_made = false;
    for (int i = 0; i < rr.len; ++i) {
        int player = player_next(rr.starting_player, i);
        int res = rr.r[i];
        if (res == R_HEALTHY) out[player] = VR_HEALTHY;
        else if (res == R_WEDDING) out[player] = (completed || player == observing_player) ? VR_WEDDING : VR_NOT_REVEALED;
        else {
            if (completed && !higher_made) {
                switch (res) {
                    case R_DIAMONDS_SOLO: out[player] = VR_DIAMONDS_SOLO; break;
                    case R_HEARTS_SOLO: out[player] = VR_HEARTS_SOLO; break;
                    case R_SPADES_SOLO: out[player] = VR_SPADES_SOLO; break;
                    case R_CLUBS_SOLO: out[player] = VR_CLUBS_SOLO; break;
                    case R_QUEENS_SOLO: out[player] = VR_QUEENS_SOLO; break;
                    case R_JACKS_SOLO: out[player] = VR_JACKS_SOLO; break;
                    case R_TRUMPLESS_SOLO: out[player] = VR_TRUMPLESS_SOLO; break;
                }
                higher_made = true;
            } else out[player] = VR_NOT_REVEALED;
        }
    }
}

// ---- team state: team/team_logic.rs:11-157 -------------------------------------------------------
struct TeamState {
    int tag = TS_IN_RESERVATIONS;
    int wedding_player = -1;
    int solved_trick_index = 0;
    uint32_t re_players = 0;  // bit p
    bool is_final() const { return tag == TS_WEDDING_SOLVED || tag == TS_NO_WEDDING; }
    bool has_re_players() const { return is_final(); }
};
inline TeamState team_resolve(const ReservationResult& rr, const Trick* tricks, int n_tricks, const Hand hands[4]) {
    TeamState ts;
    switch (rr.kind) {
        case RR_NO_RESERVATION: {                                         // :45-57
            for (int p = 0; p < 4; ++p) if (hands[p].contains(CQ)) ts.re_players |= 1u << p;
            ts.tag = TS_NO_WEDDING; return ts;
        }
        case RR_WEDDING: {                                                // :59-112
            int wedding_player = rr.player;
            int partner = -1, partner_trick = -1, completed = 0;
            for (int i = 0; i < 3; ++i) {
                if (i >= n_tricks) break;
                if (!tricks[i].is_completed()) break;
                completed += 1;
                int winner = tricks[i].winning_player;
                if (winner != wedding_player) { partner = winner; partner_trick = i; break; }
            }
            ts.wedding_player = wedding_player;
            if (partner >= 0) {
                ts.tag = TS_WEDDING_SOLVED; ts.solved_trick_index = partner_trick;
                ts.re_players = (1u << wedding_player) | (1u << partner);
            } else if (completed == 3) {
                ts.tag = TS_WEDDING_SOLVED; ts.solved_trick_index = 2; ts.re_players = 1u << wedding_player;
            } else ts.tag = TS_WEDDING_UNSOLVED;
            return ts;
        }
        case RR_SOLO: {                                                   // :113-120
            ts.tag = TS_NO_WEDDING; ts.re_players = 1u << rr.player; return ts;
        }
    }
    throw std::runtime_error("team_resolve");
}

// ---- announcements: announcement/*.rs --------------------------------------------------------------
// FdoAnnouncementSet::all_higher_than (announcement_set.rs:25-64)
inline uint32_t all_higher_than(int lowest) {
    switch (lowest) {
        case A_RE_CONTRA: return A_RE_CONTRA;
        case A_NO90: return A_RE_CONTRA | A_NO90;
        case A_NO60: return A_RE_CONTRA | A_NO90 | A_NO60;
        case A_NO30: return A_RE_CONTRA | A_NO90 | A_NO60 | A_NO30;
        case A_BLACK: return A_RE_CONTRA | A_NO90 | A_NO60 | A_NO30 | A_BLACK;
        case A_COUNTER_RE_CONTRA: return A_COUNTER_RE_CONTRA;
        default: return 0;
    }
}
// calc_number_of_cards_announcement_possible_for_last_announcement (calc_announcement.rs:21-48); -1 = None
inline int cards_possible_for_last_announcement(uint32_t prev, int wedding_solved /* -1 none */) {
    int w = wedding_solved < 0 ? 0 : wedding_solved;
    if (prev & A_BLACK) return 7 - w;
    if (prev & A_NO30) return 8 - w;
    if (prev & A_NO60) return 9 - w;
    if (prev & A_NO90) return 10 - w;
    if (prev & A_RE_CONTRA) return 11 - w;
    return -1;
}
// internal_calc_allowed_annoucements (calc_announcement.rs:51-173)
inline uint32_t internal_calc_allowed_announcements(int n_cards, uint32_t prev_team, int wedding_solved, int enemy_cards_possible) {
    int w = wedding_solved < 0 ? 0 : wedding_solved;
    int t_re = 11 - w, t90 = 10 - w, t60 = 9 - w, t30 = 8 - w, tbl = 7 - w;
    uint32_t rem = 0;
    bool p_re = n_cards >= t_re;
    bool p90 = p_re || (n_cards >= t90 && (prev_team & A_RE_CONTRA));
    bool p60 = p90 || (n_cards >= t60 && (prev_team & A_NO90));
    bool p30 = p60 || (n_cards >= t30 && (prev_team & A_NO60));
    bool pbl = p30 || (n_cards >= tbl && (prev_team & A_NO30));
    if (p_re) rem |= A_RE_CONTRA;
    if (p90) rem |= A_NO90;
    if (p60) rem |= A_NO60;
    if (p30) rem |= A_NO30;
    if (pbl) rem |= A_BLACK;
    rem &= ~(prev_team & (A_RE_CONTRA | A_NO90 | A_NO60 | A_NO30 | A_BLACK));          // :105-119
    if (rem & A_RE_CONTRA) rem &= ~(uint32_t)(A_NO90 | A_NO60 | A_NO30 | A_BLACK);     // :122-127
    if (rem & A_NO90) rem &= ~(uint32_t)(A_NO60 | A_NO30 | A_BLACK);
    if (rem & A_NO60) rem &= ~(uint32_t)(A_NO30 | A_BLACK);
    if (rem & A_NO30) rem &= ~(uint32_t)A_BLACK;
    bool regular_allowed = rem != 0;                                                  // :144
    if (enemy_cards_possible >= 0 && !regular_allowed) {                              // :146-170
        bool by_cards = n_cards >= enemy_cards_possible - 1;
        bool counter_already = prev_team & A_COUNTER_RE_CONTRA;
        bool regular_already = prev_team & A_RE_CONTRA;
        bool re_already = counter_already || regular_already;
        if (by_cards && !counter_already && !re_already) rem |= A_COUNTER_RE_CONTRA;
    }
    return rem;
}
// calc_allowed_announcements (calc_announcement.rs:176-228)
inline uint32_t calc_allowed_announcements(int player, int n_cards, const TeamState& ts, int re_lowest, int contra_lowest) {
    int wedding_solved = -1;
    switch (ts.tag) {
        case TS_IN_RESERVATIONS: throw std::runtime_error("should not happen");
        case TS_WEDDING_UNSOLVED: return 0;
        case TS_WEDDING_SOLVED: wedding_solved = ts.solved_trick_index; break;
        case TS_NO_WEDDING: break;
    }
    bool is_re = (ts.re_players >> player) & 1;
    int own = is_re ? re_lowest : contra_lowest;
    int enemy = is_re ? contra_lowest : re_lowest;
    int enemy_possible = cards_possible_for_last_announcement(all_higher_than(enemy), wedding_solved);
    return internal_calc_allowed_announcements(n_cards, all_higher_than(own), wedding_solved, enemy_possible);
}

struct AnnouncementOccurrence { uint8_t card_index, player, announcement; };
enum ProgressKind { NEXT_PLAYER_IS, ROUND_IS_OVER };
struct ProgressResult { ProgressKind kind; int player; };

struct Announcements {                                                   // announcement.rs:46-215
    AnnouncementOccurrence occ[12];
    int n = 0;
    int re_lowest = A_NONE, contra_lowest = A_NONE;
    int turns_without = 0;
    int starting_player = 0;
    uint32_t current_allowed = 0;

    void internal_play(int player, int ann, int card_index, const TeamState& ts) {  // :177-215
        if (ann == A_NO_ANNOUNCEMENT) { turns_without += 1; return; }
        turns_without = 0;
        if (n < 12) { occ[n].card_index = (uint8_t)card_index; occ[n].player = (uint8_t)player; occ[n].announcement = (uint8_t)ann; n++; }  // push result ignored (:201)
        if (!ts.has_re_players()) throw std::runtime_error("announcement without teams");
        if ((ts.re_players >> player) & 1) re_lowest = ann; else contra_lowest = ann;
    }
    ProgressResult internal_progress(int player, int card_index, const uint32_t n_cards[4], const TeamState& ts) {  // :130-175
        for (;;) {
            if (turns_without == 4) { current_allowed = 0; return {ROUND_IS_OVER, starting_player}; }
            uint32_t allowed = calc_allowed_announcements(player, (int)n_cards[player], ts, re_lowest, contra_lowest);
            if (allowed == 0) internal_play(player, A_NO_ANNOUNCEMENT, card_index, ts);
            else { current_allowed = allowed; return {NEXT_PLAYER_IS, player}; }
            player = player_next(player, 1);
        }
    }
    ProgressResult start_round(int card_index, int start, const uint32_t n_cards[4], const TeamState& ts) {  // :83-102
        turns_without = 0; starting_player = start;
        return internal_progress(start, card_index, n_cards, ts);
    }
    ProgressResult play_announcement(int player, int ann, int card_index, const uint32_t n_cards[4], const TeamState& ts) {  // :104-128
        internal_play(player, ann, card_index, ts);
        return internal_progress(player_next(player, 1), card_index, n_cards, ts);
    }
};

// ---- scoring: stats/**/*.rs ------------------------------------------------------------------------
struct AnnFlags { bool re, u90, u60, u30, black; uint32_t len; };
inline AnnFlags ann_flags(uint32_t set) {
    AnnFlags f;
    f.re = (set & A_RE_CONTRA) || (set & A_COUNTER_RE_CONTRA);
    f.u90 = set & A_NO90; f.u60 = set & A_NO60; f.u30 = set & A_NO30; f.black = set & A_BLACK;
    f.len = (uint32_t)__builtin_popcount(set);
    return f;
}
// re_won (stats/win_conditions/re_won.rs:5-116)
inline bool re_won(uint32_t re_eyes, uint32_t re_prev, uint32_t kontra_prev, bool re_all, bool kontra_all) {
    AnnFlags r = ann_flags(re_prev), k = ann_flags(kontra_prev);
    bool only_re_none = r.len == 0, only_re = r.re && !r.u90, only_re90 = r.u90 && !r.u60, only_re60 = r.u60 && !r.u30,
         only_re30 = r.u30 && !r.black, only_re_black = r.black;
    bool only_k_none = k.len == 0, only_k = k.re && !k.u90, only_k90 = k.u90 && !k.u60, only_k60 = k.u60 && !k.u30,
         only_k30 = k.u30 && !k.black, only_k_black = k.black;
    (void)re_all;
    if (re_eyes >= 121 && only_re_none && only_k_none) return true;
    if (re_eyes >= 121 && only_re && !k.re) return true;
    if (re_eyes >= 121 && only_re && only_k) return true;
    if (re_eyes >= 120 && !r.re && only_k) return true;
    if (re_eyes >= 151 && only_re90) return true;
    if (re_eyes >= 181 && only_re60) return true;
    if (re_eyes >= 211 && only_re30) return true;
    if (re_all && only_re_black) return true;
    if (re_eyes >= 90 && only_k90 && !r.u90) return true;
    if (re_eyes >= 60 && only_k60 && !r.u90) return true;
    if (re_eyes >= 30 && only_k30 && !r.u90) return true;
    if (!kontra_all && only_k_black && !r.u90) return true;
    return false;
}
// kontra_won (stats/win_conditions/kontra_won.rs:4-113)
inline bool kontra_won(uint32_t kontra_eyes, uint32_t re_prev, uint32_t kontra_prev, bool re_all, bool kontra_all) {
    AnnFlags r = ann_flags(re_prev), k = ann_flags(kontra_prev);
    bool only_re_none = r.len == 0, only_re = r.re && !r.u90, only_re90 = r.u90 && !r.u60, only_re60 = r.u60 && !r.u30,
         only_re30 = r.u30 && !r.black, only_re_black = r.black;
    bool only_k_none = k.len == 0, only_k = k.re && !k.u90, only_k90 = k.u90 && !k.u60, only_k60 = k.u60 && !k.u30,
         only_k30 = k.u30 && !k.black, only_k_black = k.black;
    if (kontra_eyes >= 120 && only_re_none && only_k_none) return true;
    if (kontra_eyes >= 120 && only_re && !k.re) return true;
    if (kontra_eyes >= 120 && only_re && only_k) return true;
    if (kontra_eyes >= 121 && !r.re && only_k) return true;
    if (kontra_eyes >= 151 && only_k90) return true;
    if (kontra_eyes >= 181 && only_k60) return true;
    if (kontra_eyes >= 211 && only_k30) return true;
    if (kontra_all && only_k_black) return true;
    if (kontra_eyes >= 90 && only_re90 && !k.u90) return true;
    if (kontra_eyes >= 60 && only_re60 && !k.u90) return true;
    if (kontra_eyes >= 30 && only_re30 && !k.u90) return true;
    if (!re_all && only_re_black && !k.u90) return true;
    return false;
}
// FdoBasicWinningPointsDetails::calculate (stats/basic_points/basic_winning_points.rs:48-284)
// details[23] in the reference struct's field order.
inline void basic_winning_points(uint32_t winner_eyes, uint32_t looser_eyes, bool winner_all_tricks, uint32_t re_prev,
                                 uint32_t kontra_prev, uint32_t re_eyes, uint32_t kontra_eyes, int& winner_pts, int& looser_pts,
                                 int details[23]) {
    (void)winner_eyes;
    AnnFlags r = ann_flags(re_prev), k = ann_flags(kontra_prev);
    for (int i = 0; i < 23; ++i) details[i] = 0;
    int w = 0, l = 0;
    auto add = [&](int idx, int v) { w += v; l -= v; details[idx] = v; };
    add(0, 1);
    if (looser_eyes < 90) add(1, 1);
    if (looser_eyes < 60) add(2, 1);
    if (looser_eyes < 30) add(3, 1);
    if (winner_all_tricks) add(4, 1);
    if (r.re) add(5, 2);
    if (k.re) add(6, 2);
    if (r.u90) add(7, 1);
    if (r.u60) add(8, 1);
    if (r.u30) add(9, 1);
    if (r.black) add(10, 1);
    if (k.u90) add(11, 1);
    if (k.u60) add(12, 1);
    if (k.u30) add(13, 1);
    if (k.black) add(14, 1);
    if (re_eyes >= 120 && k.u90) add(15, 1);
    if (re_eyes >= 90 && k.u60) add(16, 1);
    if (re_eyes >= 60 && k.u30) add(17, 1);
    if (re_eyes >= 30 && k.black) add(18, 1);
    if (kontra_eyes >= 120 && r.u90) add(19, 1);
    if (kontra_eyes >= 90 && r.u60) add(20, 1);
    if (kontra_eyes >= 60 && r.u30) add(21, 1);
    if (kontra_eyes >= 30 && r.black) add(22, 1);
    winner_pts = w; looser_pts = l;
}
// FdoBasicDrawPointsDetails::calculate (stats/basic_points/basic_draw_points.rs:27-174); details[14]
inline void basic_draw_points(uint32_t re_prev, uint32_t kontra_prev, uint32_t re_eyes, uint32_t kontra_eyes, int& re_pts,
                              int& kontra_pts, int details[14]) {
    AnnFlags r = ann_flags(re_prev), k = ann_flags(kontra_prev);
    for (int i = 0; i < 14; ++i) details[i] = 0;
    int re = 0, ko = 0;
    if (re_eyes < 90) { ko++; re--; details[3] = 1; }
    if (kontra_eyes < 90) { re++; ko--; details[0] = 1; }
    if (re_eyes < 60) { ko++; re--; details[4] = 1; }
    if (kontra_eyes < 60) { re++; ko--; details[1] = 1; }
    if (re_eyes < 30) { ko++; re--; details[5] = 1; }
    if (kontra_eyes < 30) { re++; ko--; details[2] = 1; }
    if (re_eyes >= 120 && k.u90) { re++; ko--; details[6] = 1; }
    if (re_eyes >= 90 && k.u60) { re++; ko--; details[7] = 1; }
    if (re_eyes >= 60 && k.u30) { re++; ko--; details[8] = 1; }
    if (re_eyes >= 30 && k.black) { re++; ko--; details[9] = 1; }
    if (kontra_eyes >= 120 && r.u90) { ko++; re--; details[10] = 1; }
    if (kontra_eyes >= 90 && r.u60) { ko++; re--; details[11] = 1; }
    if (kontra_eyes >= 60 && r.u30) { ko++; re--; details[12] = 1; }
    if (kontra_eyes >= 30 && r.black) { ko++; re--; details[13] = 1; }
    re_pts = re; kontra_pts = ko;
}
inline bool is_re(uint32_t re_players, int p) { return (re_players >> p) & 1; }
// calc_number_of_doppelkopf (stats/additional_points/doppelkopf.rs:8-26)
inline void calc_number_of_doppelkopf(uint32_t re_players, const Trick* tricks, int n, int& re, int& kontra) {
    re = kontra = 0;
    for (int i = 0; i < n; ++i)
        if (tricks[i].eyes() >= 40) { if (is_re(re_players, tricks[i].winning_player)) re++; else kontra++; }
}
// calc_fuchs_gefangen (stats/additional_points/fuchs_gefangen.rs:9-60)
inline void calc_fuchs_gefangen(uint32_t re_players, const Trick* tricks, int n, int& re, int& kontra) {
    re = kontra = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < tricks[i].len; ++j) {
            bool is_fuchs = tricks[i].cards[j] == DA;
            bool played_re = is_re(re_players, tricks[i].player_at(j));
            bool won_re = is_re(re_players, tricks[i].winning_player);
            if (is_fuchs && played_re != won_re) { if (won_re) re++; else kontra++; }
        }
}
// calc_trick_karlchen (stats/additional_points/last_trick_karlchen.rs:6-27)
inline void calc_trick_karlchen(uint32_t re_players, const Trick* tricks, int& re, int& kontra) {
    re = kontra = 0;
    const Trick& last = tricks[11];
    if (last.winning_card == CJ) { if (is_re(re_players, last.winning_player)) re = 1; else kontra = 1; }
}
struct AdditionalPointsDetails {
    bool present = false;
    bool against_club_queens = false;
    int doko_re = 0, doko_kontra = 0, fuchs_re = 0, fuchs_kontra = 0;
    bool karlchen_re = false, karlchen_kontra = false;
};
// FdoAdditionalPointsDetails::calculate (stats/additional_points/additional_points.rs:25-127);
// winning_team: 0 Re, 1 Kontra, -1 none.
inline void additional_points(uint32_t re_players, int winning_team, const Trick* tricks, int n, int& re_pts, int& kontra_pts,
                              AdditionalPointsDetails& d) {
    int dre, dko, fre, fko, kre, kko;
    calc_number_of_doppelkopf(re_players, tricks, n, dre, dko);
    calc_fuchs_gefangen(re_players, tricks, n, fre, fko);
    calc_trick_karlchen(re_players, tricks, kre, kko);
    int re = 0, ko = 0;
    d = AdditionalPointsDetails(); d.present = true;
    if (winning_team == 1) { ko += 1; re -= 1; d.against_club_queens = true; }
    re += dre; ko -= dre; re -= dko; ko += dko;
    re += fre; ko -= fre; re -= fko; ko += fko;
    if (kre > 0) { re += 1; ko -= 1; }
    if (kko > 0) { re -= 1; ko += 1; }
    d.doko_re = dre; d.doko_kontra = dko; d.fuchs_re = fre; d.fuchs_kontra = fko; d.karlchen_re = kre > 0; d.karlchen_kontra = kko > 0;
    re_pts = re; kontra_pts = ko;
}
struct EndOfGameStats {
    bool present = false;
    uint32_t re_players = 0;
    bool is_solo = false;
    uint32_t player_eyes[4] = {0, 0, 0, 0};
    uint32_t re_eyes = 0, kontra_eyes = 0;
    int re_points = 0, kontra_points = 0;
    int player_points[4] = {0, 0, 0, 0};
    bool has_winning_details = false; int winning_details[23] = {0};
    bool has_draw_details = false; int draw_details[14] = {0};
    AdditionalPointsDetails additional;
};
// FdoEndOfGameStats::calculate (stats/stats.rs:46-240)
inline EndOfGameStats calculate_end_of_game_stats(const uint32_t player_eyes[4], const uint32_t player_num_tricks[4],
                                                  uint32_t re_players, int re_lowest, int contra_lowest, const Trick* tricks, int n_tricks) {
    uint32_t re_tricks = 0, kontra_tricks = 0, re_eyes = 0, kontra_eyes = 0;
    for (int p = 0; p < 4; ++p) {
        if (is_re(re_players, p)) { re_eyes += player_eyes[p]; re_tricks += player_num_tricks[p]; }
        else { kontra_eyes += player_eyes[p]; kontra_tricks += player_num_tricks[p]; }
    }
    bool re_all = re_tricks == 12, kontra_all = kontra_tricks == 12;
    uint32_t re_prev = all_higher_than(re_lowest), kontra_prev = all_higher_than(contra_lowest);
    bool rw = re_won(re_eyes, re_prev, kontra_prev, re_all, kontra_all);
    bool kw = kontra_won(kontra_eyes, re_prev, kontra_prev, re_all, kontra_all);
    bool is_solo = __builtin_popcount(re_players) == 1;
    int winning_team = rw ? 0 : (kw ? 1 : -1);
    bool none_won = !rw && !kw;
    EndOfGameStats s; s.present = true;
    int re_game, kontra_game;
    if (none_won) {                                                      // :120-147
        int rb, kb; basic_draw_points(re_prev, kontra_prev, re_eyes, kontra_eyes, rb, kb, s.draw_details);
        int rx = 0, kx = 0;
        if (!is_solo) additional_points(re_players, winning_team, tricks, n_tricks, rx, kx, s.additional);
        re_game = rb + rx; kontra_game = kb + kx; s.has_draw_details = true;
    } else {                                                             // :148-212
        uint32_t winner_eyes = rw ? re_eyes : kontra_eyes, looser_eyes = rw ? kontra_eyes : re_eyes;
        bool winner_all = rw ? re_all : (kw ? kontra_all : false);
        int wp, lp; basic_winning_points(winner_eyes, looser_eyes, winner_all, re_prev, kontra_prev, re_eyes, kontra_eyes, wp, lp, s.winning_details);
        int rx = 0, kx = 0;
        if (!is_solo) additional_points(re_players, winning_team, tricks, n_tricks, rx, kx, s.additional);
        re_game = (rw ? wp : lp) + rx;
        kontra_game = (kw ? wp : lp) + kx;
        s.has_winning_details = true;
    }
    if (is_solo) re_game *= 3;                                           // :215-218
    s.re_players = re_players; s.is_solo = is_solo;
    for (int p = 0; p < 4; ++p) { s.player_eyes[p] = player_eyes[p]; s.player_points[p] = is_re(re_players, p) ? re_game : kontra_game; }
    s.re_eyes = re_eyes; s.kontra_eyes = kontra_eyes; s.re_points = re_game; s.kontra_points = kontra_game;
    return s;
}

// ---- actions: action/action.rs, action/allowed_actions.rs -------------------------------------------
inline int reservation_to_action(int res) {
    switch (res) {
        case R_HEALTHY: return ACT_RES_HEALTHY; case R_WEDDING: return ACT_RES_WEDDING;
        case R_DIAMONDS_SOLO: return ACT_RES_DIAMONDS; case R_HEARTS_SOLO: return ACT_RES_HEARTS;
        case R_SPADES_SOLO: return ACT_RES_SPADES; case R_CLUBS_SOLO: return ACT_RES_CLUBS;
        case R_TRUMPLESS_SOLO: return ACT_RES_TRUMPLESS; case R_QUEENS_SOLO: return ACT_RES_QUEENS;
        case R_JACKS_SOLO: return ACT_RES_JACKS;
    }
    throw std::runtime_error("reservation_to_action");
}
inline int action_to_reservation(int act) {                              // action.rs:265-291
    switch (act) {
        case ACT_RES_HEALTHY: return R_HEALTHY; case ACT_RES_WEDDING: return R_WEDDING;
        case ACT_RES_DIAMONDS: return R_DIAMONDS_SOLO; case ACT_RES_HEARTS: return R_HEARTS_SOLO;
        case ACT_RES_SPADES: return R_SPADES_SOLO; case ACT_RES_CLUBS: return R_CLUBS_SOLO;
        case ACT_RES_TRUMPLESS: return R_TRUMPLESS_SOLO; case ACT_RES_QUEENS: return R_QUEENS_SOLO;
        case ACT_RES_JACKS: return R_JACKS_SOLO;
    }
    throw std::runtime_error("action_to_reservation");
}
inline int action_to_announcement(int act) {                             // action.rs:293-310 (33 is ALWAYS ReContra)
    switch (act) {
        case ACT_ANN_RE_CONTRA: return A_RE_CONTRA; case ACT_ANN_NO90: return A_NO90; case ACT_ANN_NO60: return A_NO60;
        case ACT_ANN_NO30: return A_NO30; case ACT_ANN_BLACK: return A_BLACK; case ACT_NO_ANNOUNCEMENT: return A_NO_ANNOUNCEMENT;
    }
    throw std::runtime_error("action_to_announcement");
}
// FdoAllowedActions::calculate_allowed_actions (action/allowed_actions.rs:68-169); returns 39-bit mask
inline uint64_t calculate_allowed_actions(int phase, Color trick_color, Hand hand, int game_type, uint32_t allowed_announcements) {
    switch (phase) {
        case PH_RESERVATION: {                                           // :76-96
            uint64_t a = 1ull << ACT_RES_HEALTHY;
            if (hand.contains_both(CQ)) a |= 1ull << ACT_RES_WEDDING;
            a |= (1ull << ACT_RES_JACKS) | (1ull << ACT_RES_QUEENS) | (1ull << ACT_RES_DIAMONDS) | (1ull << ACT_RES_HEARTS) |
                 (1ull << ACT_RES_SPADES) | (1ull << ACT_RES_CLUBS) | (1ull << ACT_RES_TRUMPLESS);
            return a;
        }
        case PH_PLAYCARD: {                                              // :97-140
            uint64_t single = (hand.bits | (hand.bits >> 24)) & 0xFFFFFFull;
            if (trick_color == COLOR_NONE) return single;
            uint64_t mask = get_color_masks_for_game_type(game_type).m[trick_color];
            if (single & mask) return single & mask;
            return single;
        }
        case PH_ANNOUNCEMENT: {                                          // :144-166
            uint64_t a = 0;
            if ((allowed_announcements & A_COUNTER_RE_CONTRA) || (allowed_announcements & A_RE_CONTRA)) a |= 1ull << ACT_ANN_RE_CONTRA;
            if (allowed_announcements & A_NO90) a |= 1ull << ACT_ANN_NO90;
            if (allowed_announcements & A_NO60) a |= 1ull << ACT_ANN_NO60;
            if (allowed_announcements & A_NO30) a |= 1ull << ACT_ANN_NO30;
            if (allowed_announcements & A_BLACK) a |= 1ull << ACT_ANN_BLACK;
            a |= 1ull << ACT_NO_ANNOUNCEMENT;
            return a;
        }
        default: throw std::runtime_error("Es sind keine Aktionen mehr möglich, da das Spiel beendet ist.");  // :141-143
    }
}
constexpr uint64_t ANNOUNCEMENT_CALL_ACTIONS = (1ull << ACT_ANN_RE_CONTRA) | (1ull << ACT_ANN_NO90) | (1ull << ACT_ANN_NO60) |
                                              (1ull << ACT_ANN_NO30) | (1ull << ACT_ANN_BLACK);

// ---- state: state/state.rs:24-518 ----------------------------------------------------------------
struct State {
    ReservationRound reservations_round;
    Trick tricks[12];
    int n_tricks = 0;
    Hand hands[4];
    Announcements announcements;
    int card_index = 0;
    int current_player = -1;       // -1 = None
    int current_phase = PH_RESERVATION;
    ReservationResult reservation_result;
    int game_type = GT_NONE;
    uint32_t player_eyes[4] = {0, 0, 0, 0};
    uint32_t player_num_tricks[4] = {0, 0, 0, 0};
    TeamState team_state;
    EndOfGameStats end_of_game_stats;
    uint32_t n_play_actions = 0;   // oracle-side bookkeeping: number of play_action calls ("game steps")

    static State new_game_from_hand_and_start_player(const Hand hands[4], int start_player) {  // :125-166
        State s;
        s.reservations_round.starting_player = (int8_t)start_player;
        for (int p = 0; p < 4; ++p) s.hands[p] = hands[p];
        s.current_player = start_player;
        return s;
    }
    static State new_game(Rng& rng) {                                    // :169-178 (start player FIRST, then shuffle)
        int start = (int)rng.start_player();
        Hand hands[4];
        randomly_distributed(rng, hands);
        return new_game_from_hand_and_start_player(hands, start);
    }
    void hand_lens(uint32_t out[4]) const { for (int p = 0; p < 4; ++p) out[p] = hands[p].len(); }

    void progress_next_card_or_announcement() {                          // :184-206
        uint32_t lens[4]; hand_lens(lens);
        ProgressResult r = announcements.start_round(card_index, current_player, lens, team_state);
        current_phase = r.kind == NEXT_PLAYER_IS ? PH_ANNOUNCEMENT : PH_PLAYCARD;
        current_player = r.player;
    }
    Color current_trick_color() const {                                  // :360-372 (None when 12 tricks exist!)
        if (n_tricks > 0 && n_tricks < 12) return tricks[n_tricks - 1].color(game_type);
        return COLOR_NONE;
    }
    uint64_t allowed_actions() const {
        if (current_phase == PH_FINISHED) return 0;                      // observation: :486-488
        return calculate_allowed_actions(current_phase, current_trick_color(), hands[current_player], game_type,
                                         announcements.current_allowed);
    }

    void play_action(int action) {                                       // :208-358
        if (current_phase == PH_FINISHED) throw std::runtime_error("play_action on a finished game");
        n_play_actions++;
        int cp = current_player;
        if (action >= ACT_RES_HEALTHY && action <= ACT_RES_JACKS) {      // Reservation :216-250
            reservations_round.play_reservation(action_to_reservation(action));
            int np = player_next(cp, 1);
            current_player = np;
            if (reservations_round.is_completed()) {
                reservation_result = winning_player_in_reservation_round(reservations_round);
                game_type = to_game_type(reservation_result);
                if (!team_state.is_final()) team_state = team_resolve(reservation_result, tricks, n_tricks, hands);
                tricks[n_tricks++] = Trick::empty(np);
                progress_next_card_or_announcement();
            }
        } else if (action >= ACT_ANN_RE_CONTRA && action <= ACT_NO_ANNOUNCEMENT) {   // Announcement :252-272
            uint32_t lens[4]; hand_lens(lens);
            ProgressResult r = announcements.play_announcement(cp, action_to_announcement(action), card_index, lens, team_state);
            if (r.kind == NEXT_PLAYER_IS) current_player = r.player;
            else { current_phase = PH_PLAYCARD; current_player = r.player; }
        } else if (action >= 0 && action < 24) {                        // Card :274-357
            int card = action;
            hands[cp].remove(card);
            Trick& t = tricks[n_tricks - 1];
            t.play_card(card, game_type);
            card_index += 1;
            if (t.is_completed()) {
                int wp = t.winning_player;
                player_eyes[wp] += t.eyes();
                player_num_tricks[wp] += 1;
                if (!team_state.is_final()) team_state = team_resolve(reservation_result, tricks, n_tricks, hands);
                if (n_tricks == 12) {
                    current_phase = PH_FINISHED; current_player = -1;
                    end_of_game_stats = calculate_end_of_game_stats(player_eyes, player_num_tricks, team_state.re_players,
                                                                    announcements.re_lowest, announcements.contra_lowest, tricks, n_tricks);
                    return;
                }
                current_player = wp;
                tricks[n_tricks++] = Trick::empty(wp);
                progress_next_card_or_announcement();
                return;
            }
            current_player = player_next(cp, 1);
            progress_next_card_or_announcement();
        } else throw std::runtime_error("Invalid action");
    }

    // Chained card draws (rng.hpp PhiloxStream::chain): the product of the numbers of legal card types the plays already made in the
    // running trick chose from (1 at the start of a trick) — each recomputed from this state: the seat's hand with its card back in
    // it under the colour rule of its turn.  Rng::set_card_position(card_index, card_chain_mul()) resumes the stream inside a trick.
    uint32_t card_chain_mul() const {
        if (n_tricks == 0 || current_phase == PH_FINISHED) return 1;
        const Trick& t = tricks[n_tricks - 1];
        if (t.len == 4) return 1;
        uint32_t mul = 1;
        for (int k = 0; k < t.len; ++k) {
            Hand h = hands[t.player_at(k)];
            h.add(t.cards[k]);
            const Color col = (k > 0 && n_tricks < 12) ? card_to_color(t.cards[0], game_type) : COLOR_NONE;
            mul *= (uint32_t)popcount64(calculate_allowed_actions(PH_PLAYCARD, col, h, game_type, 0));
        }
        return mul;
    }
    void position_streams(Rng& rng) const {      // sub-streams of the state-derived sites at this state's ordinals
        rng.set_card_position((uint32_t)card_index, card_chain_mul());
        rng.set_ordinal(SITE_RESERVATION, (uint32_t)reservations_round.len);
    }
    static Site site_for_phase(int phase) {
        return phase == PH_RESERVATION ? SITE_RESERVATION : (phase == PH_ANNOUNCEMENT ? SITE_ANNOUNCEMENT : SITE_CARD);
    }
    // random_action_for_current_player (:378-399).  Returns true if the game was already finished.
    bool random_action_for_current_player(Rng& rng, int* action_out = nullptr) {
        if (current_phase == PH_FINISHED) return true;
        uint64_t allowed = allowed_actions();
        uint64_t bit = random_single(allowed, rng, site_for_phase(current_phase));
        int a = __builtin_ctzll(bit);
        if (action_out) *action_out = a;
        play_action(a);
        return false;
    }
    // random_action_for_current_player_no_announcement (:401-431)
    bool random_action_for_current_player_no_announcement(Rng& rng, int* action_out = nullptr) {
        if (current_phase == PH_FINISHED) return true;
        uint64_t allowed = allowed_actions() & ~ANNOUNCEMENT_CALL_ACTIONS;
        uint64_t bit = random_single(allowed, rng, site_for_phase(current_phase));
        int a = __builtin_ctzll(bit);
        if (action_out) *action_out = a;
        play_action(a);
        return false;
    }
    int observing_player() const { return current_player < 0 ? 0 : current_player; }  // :437-439
    int wedding_player_if_wedding_announced() const {                    // :444-448
        return (team_state.tag == TS_WEDDING_UNSOLVED || team_state.tag == TS_WEDDING_SOLVED) ? team_state.wedding_player : -1;
    }
};

// McFullDokoEnvState::random_rollout (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:198-220):
// clone, random no-announcement actions to the end, reward = player_points as f64.
inline void random_rollout(const State& s, Rng& rng, double rewards[4], uint32_t* steps = nullptr) {
    State r = s;
    uint32_t before = r.n_play_actions;
    for (;;) { if (r.random_action_for_current_player_no_announcement(rng)) break; }
    for (int p = 0; p < 4; ++p) rewards[p] = (double)r.end_of_game_stats.player_points[p];
    if (steps) *steps = r.n_play_actions - before;
}

}  // namespace fdo
}  // namespace oracle
