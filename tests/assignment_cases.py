"""Shared helpers for the rs-doko-assignment support-set cases (tests/golden/assignment_sets.json)."""
import json
import os

import numpy as np

from oracle_lib import DK_STATE_DTYPE, hand_from_cards

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "assignment_sets.json")))["cases"]


def case_record(case):
    """A dk_state (DK_DOKO) that presents the case's information to the sampler: history, observer hand, hand sizes.
    The hidden hands are filled with the remaining cards in any order — the sampler only reads their sizes."""
    rec = np.zeros(1, dtype=DK_STATE_DTYPE)
    played = [c for t in case["tricks"] for c in t["cards"]]
    left = [2] * 24
    for c in played + case["hand"]:
        left[c] -= 1
    pool = [c for c in range(24) for _ in range(left[c])]
    hands = []
    for p in range(4):
        if p == case["observer"]:
            hands.append(hand_from_cards(case["hand"]))
        else:
            take, pool = pool[:case["lens"][p]], pool[case["lens"][p]:]
            hands.append(hand_from_cards(take))
    assert not pool
    rec["hands"][0] = hands
    rec["cards"][0][:] = 0xFF
    rec["cards"][0][:len(played)] = played
    rec["announcements"][0][:] = 0xFFFF
    rec["reservations"][0][:] = 1                     # DoReservation::Healthy
    tr = 0
    for t, trick in enumerate(case["tricks"]):
        tr |= trick["start"] << (2 * t)
    rec["tricks"][0] = tr | (len(case["tricks"]) << 24)
    rec["card_index"][0] = len(played)
    rec["n_reservations"][0] = 4
    if case["marriage"] >= 0:
        gt, tag, wed = 1, 2, case["marriage"]
    else:
        gt, tag, wed = 0, 3, 0
    rec["meta"][0] = 2 | (case["observer"] << 2) | (case["tricks"][0]["start"] << 4) | (gt << 6) | (tag << 10) | (wed << 12)
    return rec


def canonical(hands):
    """4 bitboards → tuple of sorted card lists (what the reference's HashSet dedups)."""
    out = []
    for h in hands:
        cards = []
        for c in range(24):
            cards += [c] * (((h >> c) & 1) + ((h >> (c + 24)) & 1))
        out.append(tuple(cards))
    return tuple(out)


def expected_set(case):
    return {tuple(tuple(h) for h in a) for a in case["assignments"]}
