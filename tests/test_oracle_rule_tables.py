"""The reference's table-style known-answer tests (fixture tests/golden/rule_tables.json, extracted by
tests/golden/make_rule_tables.py) replayed on the oracle, and — where a device function exists — on the device logic."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import hostsim_lib

T = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rule_tables.json")))


def test_is_greater_in_trick_all_game_types(orc):
    """card_in_trick_logic.rs:226-1023 — 648 rows x colours x cards x game types."""
    n = 0
    for row in T["is_greater_in_trick"]:
        for gt in row["game_types"]:
            for col in row["colors"]:
                for prev in row["previous"]:
                    assert bool(orc.orc_fdo_is_greater_in_trick(row["current"], prev, col, gt)) == row["greater"], (row, gt, col, prev)
                    n += 1
    assert n > 10000


def test_device_card_power_matches_is_greater_exhaustively(orc):
    """power(b) > power(a) <=> is_greater_in_trick(b, a, colour of the lead) for every lead/a/b/game type."""
    sim = hostsim_lib.load()
    for gt in range(9):
        trump = sim.sim_trump_mask(gt)
        for lead in range(24):
            follow = sim.sim_follow_mask(lead, trump)
            col = orc.orc_fdo_card_to_color(lead, gt)
            for a in range(24):
                # `a` is a possible current best only if it is trump or follows the lead
                pa = sim.sim_card_power(a, trump, follow)
                if pa == 0:
                    continue
                for b in range(24):
                    pb = sim.sim_card_power(b, trump, follow)
                    assert (pb > pa) == bool(orc.orc_fdo_is_greater_in_trick(b, a, col, gt)), (gt, lead, a, b)


def test_card_to_color_and_masks(orc):
    for card, gt, col in T["card_to_color"]:
        assert orc.orc_fdo_card_to_color(card, gt) == col
    assert len(T["card_to_color"]) == 216
    sim = hostsim_lib.load()
    # card_color_masks.rs:249-315: masks agree with card_to_color for every card x game type; device trump/follow masks too
    for gt in range(9):
        m = (C.c_uint64 * 5)()
        orc.orc_fdo_color_masks(gt, m)
        for card in range(24):
            col = orc.orc_fdo_card_to_color(card, gt)
            for k in range(5):
                assert bool((m[k] >> card) & 1) == (k == col)
            assert sim.sim_follow_mask(card, sim.sim_trump_mask(gt)) == m[col]
        assert sim.sim_trump_mask(gt) == m[0]
    # rs-doko/src/card/card_color_masks.rs:36-41
    orc.orc_fdo_color_masks(0, m)
    assert [m[0], m[2], m[3], m[4]] == [0b00110000_11000011_10111111, 0b00000000_00001100_01000000, 0b11001100_00000000_00000000, 0b00000011_00110000_00000000]


def test_eyes(orc):
    for card, e in T["eyes"]:
        assert orc.orc_fdo_card_eyes(card) == e


def test_win_conditions(orc):
    for e, rp, kp, ra, ka, exp in T["re_won"]:
        assert bool(orc.orc_fdo_re_won(e, rp, kp, int(ra), int(ka))) == exp
    for e, rp, kp, ra, ka, exp in T["kontra_won"]:
        assert bool(orc.orc_fdo_kontra_won(e, rp, kp, int(ra), int(ka))) == exp
    assert len(T["re_won"]) >= 18 and len(T["kontra_won"]) >= 30


def test_basic_points(orc):
    for row in T["basic_winning_points"]:
        out = (C.c_int32 * 25)()
        a = row["args"]
        orc.orc_fdo_basic_winning_points(a[0], a[1], int(a[2]), a[3], a[4], a[5], a[6], out)
        assert list(out) == [row["winner"], row["loser"]] + row["details"]
    for row in T["basic_draw_points"]:
        out = (C.c_int32 * 16)()
        orc.orc_fdo_basic_draw_points(*row["args"], out)
        assert list(out) == [row["re"], row["kontra"]] + row["details"]


def test_calc_allowed_announcements(orc):
    for r in T["calc_allowed_announcements"]:
        got = orc.orc_fdo_calc_allowed(r["player"], r["n_cards"], r["tag"], r["wedding_player"], r["solved_idx"], r["re_players"], r["re_lowest"], r["contra_lowest"])
        assert got == r["expected"], r
    for r in T["internal_calc_allowed_announcements"]:
        assert orc.orc_fdo_internal_calc_allowed(r["n_cards"], r["prev"], r["wedding_solved"], r["enemy_possible"]) == r["expected"], r


LEVEL_BITS = [0, 1, 2, 4, 8, 16, 32]


def test_device_allowed_call_matches_oracle_exhaustively(orc):
    """Closed form (fdo_allowed_call) vs the literal set logic for every (cards, own lowest, enemy lowest, wedding shift)."""
    sim = hostsim_lib.load()
    act = {0: 0, 1: 1, 2: 2, 4: 3, 8: 4, 16: 5, 32: 1}
    for c in range(0, 13):
        for m in range(7):
            for e in range(7):
                for w in (0, 1, 2):
                    tag, si = (2, w) if w else (3, 0)
                    exp = orc.orc_fdo_calc_allowed(0, c, tag, 0, si, 0b0001, LEVEL_BITS[m], LEVEL_BITS[e])
                    assert bin(exp).count("1") <= 1
                    assert sim.sim_fdo_allowed_call(c, m, e, w) == act[exp], (c, m, e, w)


def test_device_min_cards_to_call_is_the_threshold_of_allowed_call():
    """fdo_min_cards_to_call(m, e, w) = the smallest hand size for which fdo_allowed_call is non-empty (99 = never), exhaustively."""
    sim = hostsim_lib.load()
    for m in range(7):
        for e in range(7):
            for w in (0, 1, 2):
                ok = [c for c in range(0, 13) if sim.sim_fdo_allowed_call(c, m, e, w) != 0]
                thr = sim.sim_fdo_min_cards_to_call(m, e, w)
                if ok:
                    assert thr == min(ok) and ok == list(range(min(ok), 13)), (m, e, w, thr, ok)
                else:
                    assert thr > 12, (m, e, w, thr)


def test_device_byte_parallel_eligibility_matches_the_per_seat_comparison():
    """The announcement replay decides "cards on hand >= threshold of the seat's team" for four seats with one byte-wise subtraction and
    one multiplication (fdo_eligible_nibble): equal to the per-seat comparison for every hand-size pattern the game produces (all seats
    hold c or c - 1 cards), both teams' thresholds incl. 99 = never, and every Re mask."""
    import numpy as np

    sim = hostsim_lib.load()
    thresholds = list(range(0, 13)) + [99]
    for cmax in range(1, 13):
        for played in range(16):
            cards = np.array([cmax - ((played >> s) & 1) for s in range(4)], dtype=np.uint32)
            for thr_re in thresholds:
                for thr_ko in thresholds:
                    for re in range(16):
                        want = sum(1 << s for s in range(4) if cards[s] >= (thr_re if (re >> s) & 1 else thr_ko))
                        assert sim.sim_fdo_eligible_nibble(hostsim_lib.ptr(cards), thr_re, thr_ko, re) == want, (cmax, played, thr_re, thr_ko, re)


def test_device_lookup_tables_match_the_closed_forms():
    """The shared-memory tables of the playout kernels against the functions they replace:
    rank select (12-bit table and 64-entry table vs the popcount binary search, which test_select_by_rank pins against the reference's
    select_by_rank), card strength + eyes (vs card_power / the eyes table), call thresholds of both teams, caller of a segment."""
    import numpy as np

    sim = hostsim_lib.load()
    prng = np.random.default_rng(5)
    masks = [int(x) for x in prng.integers(1, 1 << 24, size=3000)] + [1, 1 << 23, (1 << 24) - 1, 0xFFF, 0xFFF000, 0x800001]
    for m in masks:
        if bin(m).count("1") > 12:                                  # a hand holds at most 12 card types
            m &= int(prng.integers(1, 1 << 24)) | 1
        for idx in range(bin(m).count("1")):
            want = sim.sim_pick_msb_rank24(m, idx)
            assert sim.sim_pick_msb_rank24_tab(m, idx) == want and sim.sim_pick_msb_rank24_lut(m, idx) == want, (hex(m), idx)
            assert want == [b for b in range(23, -1, -1) if (m >> b) & 1][idx]
    eyes = [0, 10, 2, 3, 4, 11]
    for gt in range(9):
        trump = sim.sim_trump_mask(gt)
        for first in range(24):
            follow = sim.sim_follow_mask(first, trump)
            for c in range(24):
                v = sim.sim_pow_lookup(gt, first, c)
                # the trick accumulator's record: eyes | ♦A bit | empty position field | card id | strength (dk_common.cuh pow_lut_entry)
                assert v >> 16 == sim.sim_card_power(c, trump, follow) and v & 255 == eyes[c % 6], (gt, first, c)
                assert (v >> 11) & 31 == c and (v >> 9) & 3 == 0 and ((v >> 8) & 1) == (1 if c == 5 else 0), (gt, first, c)
            # what the one-max trick winner rests on: within a row, two different card types never share a positive strength, and the
            # card that leads has a positive one
            row = [sim.sim_pow_lookup(gt, first, c) >> 16 for c in range(24)]
            positive = [x for x in row if x > 0]
            assert len(positive) == len(set(positive)) and row[first] > 0, (gt, first)
    for w in (0, 1, 2):
        for re_low in range(7):
            for ko_low in range(7):
                v = sim.sim_thr2_lut(w, re_low, ko_low)
                assert v & 255 == sim.sim_fdo_min_cards_to_call(re_low, ko_low, w) and v >> 8 == sim.sim_fdo_min_cards_to_call(ko_low, re_low, w)
    for win in range(16):
        seats = [d for d in range(4) if (win >> d) & 1]
        for hit in range(1, 1 << len(seats)):
            j = (hit & -hit).bit_length() - 1                       # the first eligible seat whose decision bit is set
            assert sim.sim_seg_lut(win, hit) == (seats[j] | (j << 2)), (win, hit)


def test_device_score_matches_oracle_exhaustively(orc):
    """Closed-form scoring (fdo_score) vs the literal stats.rs restatement over all calls x eyes x trick extremes x team sizes."""
    sim = hostsim_lib.load()
    tricks = np.zeros(60, dtype=np.int32)   # 12 dull tricks (no extras): ♥9 x4 is not a legal trick but scoring only reads eyes/cards
    for t in range(12):
        tricks[t * 5:t * 5 + 5] = [6, 6, 10, 10, 0]
    tp = tricks.ctypes.data_as(C.c_void_p)
    out = (C.c_int32 * 56)()
    for rl in range(7):
        for kl in range(7):
            for re_eyes in list(range(0, 241, 3)) + [29, 30, 59, 60, 89, 90, 119, 120, 121, 150, 151, 180, 181, 210, 211, 240]:
                for re_tricks in ((0, 5) if re_eyes == 0 else (12, 7) if re_eyes == 240 else (6,)):
                    for n_re in (1, 2):
                        re_players = 0b0001 if n_re == 1 else 0b0011
                        eyes = (C.c_uint32 * 4)(re_eyes, 0, 240 - re_eyes, 0)
                        ntr = (C.c_uint32 * 4)(re_tricks, 0, 12 - re_tricks, 0)
                        orc.orc_fdo_end_of_game_stats(eyes, ntr, re_players, LEVEL_BITS[rl], LEVEL_BITS[kl], tp, out)
                        exp_re, exp_ko = out[2], out[3]
                        extras_oracle = 0
                        ko = C.c_int32()
                        got_re = sim.sim_fdo_score(re_eyes, re_tricks, n_re, rl, kl, extras_oracle, C.byref(ko))
                        assert (got_re, ko.value) == (exp_re, exp_ko), (rl, kl, re_eyes, re_tricks, n_re)


def test_straight_line_score_equals_its_specification():
    """dk::fdo_score (straight-line: clamps over eyes / 30) == the case-by-case closed form it replaced, over EVERY input:
    241 eyes x 13 trick counts x 3 team sizes x 7 x 7 lowest calls x 25 extras = 3.8e7 combinations."""
    sim = hostsim_lib.load()
    assert sim.sim_fdo_score_mismatches() == 0
    # and the team sums by masked multiplication (fdo_final_points) == the seat-by-seat loop, on 4e6 random tracker sets
    assert sim.sim_fdo_final_points_mismatches(4_000_000) == 0


def test_end_of_game_stats_real_games(orc):
    """stats.rs:258,382,489,638 — four real games."""
    for r in T["end_of_game_stats"]:
        tricks = np.zeros(60, dtype=np.int32)
        for t, tr in enumerate(r["tricks"]):
            tricks[t * 5:t * 5 + 4] = tr["cards"]
            tricks[t * 5 + 4] = tr["start"]
        out = (C.c_int32 * 56)()
        orc.orc_fdo_end_of_game_stats((C.c_uint32 * 4)(*r["eyes"]), (C.c_uint32 * 4)(*r["num_tricks"]), r["re_players"], r["re_lowest"],
                                      r["contra_lowest"], tricks.ctypes.data_as(C.c_void_p), out)
        o = list(out)
        assert o[0:4] == [r["re_eyes"], r["kontra_eyes"], r["re_points"], r["kontra_points"]]
        assert o[4:8] == r["player_points"]
        assert bool(o[8]) == r["is_solo"]
        if r["additional"] is None:
            assert o[11] == 0
        else:
            a = r["additional"]
            assert o[11:19] == [1, int(a["against_club_queens"]), a["number_of_doppelkopf_re"], a["number_of_doppelkopf_kontra"], a["fuchs_gefangen_re"],
                                a["fuchs_gefangen_kontra"], int(a["karlchen_last_trick_re"]), int(a["karlchen_last_trick_kontra"])]
    assert len(T["end_of_game_stats"]) == 4


def test_rs_doko_is_greater_and_select_by_rank(orc):
    for a, b, col, exp in T["doko_is_greater"]:
        assert bool(orc.orc_doko_is_greater(a, b, col)) == exp
    sim = hostsim_lib.load()
    for v, r, e in T["select_by_rank"]:
        assert orc.orc_select_by_rank(v, r) == e
        assert 1 << sim.sim_select_lsb(v, bin(v).count("1") - 1 - r) == e
    # MSB-rank select: device binary search == the reference's branch-free ladder on random masks
    rng = np.random.default_rng(0)
    for _ in range(3000):
        v = int(rng.integers(1, 1 << 24))
        for r in range(bin(v).count("1")):
            assert 1 << sim.sim_select_lsb24(v, bin(v).count("1") - 1 - r) == orc.orc_select_by_rank(v, r)
