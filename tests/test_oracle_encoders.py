"""Known-answer tests that pin the oracle's encoders (and through them SmallRng emulation + rs-doko rules).

* encode_state_pi: the reference's hand-built 311-entry test (rs-doko-networks/src/full_doko/var1/encode_pi.rs:237-793),
  fixture tests/golden/encode_pi_vector.json.  Checked on the oracle AND on the device logic (hostsim).
* encode_state: seed 0 → new_game + 15 random actions → exact 110 tokens (rs-doko-embeddings/src/encode_state.rs:349-597).
  This vector also pins the rand 0.9.0 SmallRng emulation (seed_from_u64, range sampling, shuffle) end to end.
"""
import ctypes as C
import json
import os

import numpy as np

import hostsim_lib
import oracle_lib
from oracle_lib import DK_STATE_DTYPE, Doko, Fdo, card_id, hand_from_cards

G = os.path.join(os.path.dirname(__file__), "golden")


def pi_record(path=None):
    v = json.load(open(path or os.path.join(G, "encode_pi_vector.json")))
    st = v["state"]
    rec = np.zeros(1, dtype=DK_STATE_DTYPE)
    r = rec[0]
    r["hands"] = [hand_from_cards(h) for h in st["hands"]]
    r["cards"] = 0xFF
    tr = 0
    ci = 0
    for t, trick in enumerate(st["tricks"]):
        tr |= trick["start"] << (2 * t)
        for c in trick["cards"]:
            rec["cards"][0][ci] = c
            ci += 1
    assert ci == st["card_index"]
    r["announcements"] = 0xFFFF
    for a, c in enumerate(st["calls"]):
        rec["announcements"][0][a] = c["card_index"] | (c["player"] << 6) | (c["level"] << 8)
    r["reservations"] = st["reservations"]
    r["tricks"] = tr | (len(st["tricks"]) << 24) | (len(st["calls"]) << 28)
    r["eyes"] = st["eyes"]
    r["num_tricks"] = sum(n << (4 * p) for p, n in enumerate(st["num_tricks"]))
    r["card_index"] = st["card_index"]
    r["n_reservations"] = 4
    re_mask = sum(1 << p for p in st["re_players"])
    r["meta"] = (st["phase"] | (st["current_player"] << 2) | (st["start"] << 4) | (st["game_type"] << 6) | (3 << 10) | (re_mask << 16) |
                 (st["re_lowest"] << 20) | (st["contra_lowest"] << 23) | (st["turns_without"] << 26) | (st["ann_start"] << 29))
    return rec, np.array(v["expected"], dtype=np.int64)


def test_encode_state_pi_reference_vector_oracle(orc):
    rec, expected = pi_record()
    o = Fdo.from_dk_state(orc, rec)
    assert np.array_equal(o.encode_pi(), expected)


def test_encode_state_pi_reference_vector_device_logic():
    sim = hostsim_lib.load()
    rec, expected = pi_record()
    out = np.zeros(311, dtype=np.int64)
    sim.sim_encode(2, hostsim_lib.ptr(rec), hostsim_lib.ptr(out))
    assert np.array_equal(out, expected)


# rs-doko-embeddings/src/encode_state.rs:372-594, transcribed: cards as names, seats relative to the seat to move (TOP)
SEED0_PLAYED = ["SK", "SA", "SA", "S9", "H9", "SJ", "HA", "HA", "C10", "SJ", "CK"]
SEED0_HANDS = [["H10", "CQ", "DQ", "DJ", "CA", "CA", "S10", "S10", "S9", "H9"],
               ["SQ", "SQ", "HQ", "DQ", "DJ", "DK", "CK", "C9", "C9"],
               ["CJ", "CJ", "D10", "D10", "DK", "D9", "D9", "HK", "HK"],
               ["H10", "CQ", "HQ", "HJ", "HJ", "DA", "DA", "C10", "SK"]]


def seed0_expected():
    rel = lambda cur, tgt: (tgt + 4 - cur) % 4
    TOP, LEFT, RIGHT = 2, 1, 3
    v = [1, rel(TOP, LEFT) + 1, rel(TOP, LEFT) + 1, rel(TOP, TOP) + 1, rel(TOP, RIGHT) + 1] + [0] * 9
    v += [card_id(c) + 1 for c in SEED0_PLAYED] + [0] * (48 - len(SEED0_PLAYED))
    for h in SEED0_HANDS:
        v += [card_id(c) + 1 for c in h] + [0] * (12 - len(h))
    return np.array(v, dtype=np.int64)


def test_encode_state_seed0_reference_vector(orc):
    h = orc.orc_doko_new_game_smallrng_play(0, 15)
    o = Doko(orc, h)
    exp = seed0_expected()
    assert len(exp) == 110
    assert np.array_equal(o.encode(False), exp)
    # and the device logic on the same state
    sim = hostsim_lib.load()
    rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
    out = np.zeros(110, dtype=np.int64)
    sim.sim_encode(0, hostsim_lib.ptr(rec), hostsim_lib.ptr(out))
    assert np.array_equal(out, exp)


def test_distribute_cards_seed42_reference_vector(orc):
    """rs-full-doko/src/hand/hand.rs:559-585 — SmallRng seed 42 shuffle → four exact hand bitboards."""
    hands = (C.c_uint64 * 4)()
    orc.orc_smallrng_distribute_cards(42, 1, hands)
    assert [int(x) for x in hands] == [0b0000000010000000000000110100100110000110110001, 0b0001000000100001000000000001010101100001100101,
                                       0b1000000000000000001000011110110000000000011110, 0b0000001000011000000000100010001001011110000010]
