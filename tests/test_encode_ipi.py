"""encode_state_ipi: the reference's hand-built 311-entry known answer
(rs-doko-networks/src/full_doko/var1/encode_ipi.rs:322-889 → tests/golden/encode_ipi_vector.json) on the oracle and on the device
logic, and device logic vs oracle on random games with partial guesses of the hidden hands / reservations."""
import json
import os

import numpy as np
import pytest

import hostsim_lib
from oracle_lib import DK_STATE_DTYPE, Fdo, hand_from_cards
from test_oracle_encoders import pi_record

G = os.path.join(os.path.dirname(__file__), "golden")
SEED = 0x1F1


def ipi_case():
    v = json.load(open(os.path.join(G, "encode_ipi_vector.json")))
    rec, _ = pi_record(os.path.join(G, "encode_ipi_vector.json"))
    hands = [hand_from_cards(h) for h in v["assumed_hands"]]
    res = [0xFF if r is None else r for r in v["assumed_reservations"]]
    return rec, hands, res, v["next_player"], np.array(v["expected"], dtype=np.int64)


def test_reference_vector_oracle(orc):
    rec, hands, res, nxt, expected = ipi_case()
    o = Fdo.from_dk_state(orc, rec)
    assert np.array_equal(o.encode_ipi(hands, res, nxt), expected)


def test_reference_vector_device_logic():
    sim = hostsim_lib.load()
    rec, hands, res, nxt, expected = ipi_case()
    out = np.zeros(311, dtype=np.int64)
    err = sim.sim_encode_ipi(hostsim_lib.ptr(rec), hostsim_lib.ptr(np.array(hands, dtype=np.uint64)), hostsim_lib.ptr(np.array(res, dtype=np.uint8)), nxt,
                             hostsim_lib.ptr(out))
    assert err == 0 and np.array_equal(out, expected)


def partial_guess(prng, o):
    """Random subset of every hidden hand (copy-aware), random guesses for some reservations, random seat to guess for."""
    hands = o.hands()
    cur = max(o.info()["current_player"], 0)
    assumed = []
    for p in range(4):
        bits = [b for b in range(48) if (hands[p] >> b) & 1]
        # FdoHand keeps copy A in the low plane: a guessed hand is built with add(), so one copy of a card sits in plane A
        cards = sorted(b % 24 for b in bits)
        keep = [c for c in cards if prng.random() < 0.5] if p != cur else cards
        assumed.append(hand_from_cards(keep))
    res = [int(prng.integers(0, 9)) if prng.random() < 0.6 else 0xFF for _ in range(4)]
    return assumed, res, int(prng.integers(0, 4))


def test_device_logic_matches_oracle_on_random_games(orc):
    sim = hostsim_lib.load()
    prng = np.random.default_rng(3)
    n = 0
    tokens_seen = set()
    for g in range(40):
        o = Fdo.new_game_philox(orc, SEED, g, 0)
        while o.allowed():
            rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
            assumed, res, nxt = partial_guess(prng, o)
            want = o.encode_ipi(assumed, res, nxt)
            out = np.zeros(311, dtype=np.int64)
            err = sim.sim_encode_ipi(hostsim_lib.ptr(rec), hostsim_lib.ptr(np.array(assumed, dtype=np.uint64)), hostsim_lib.ptr(np.array(res, dtype=np.uint8)),
                                     nxt, hostsim_lib.ptr(out))
            assert err == 0 and np.array_equal(out, want), (g, n)
            tokens_seen.update(want[:4].tolist())
            m = o.allowed()
            legal = [a for a in range(39) if (m >> a) & 1]
            a = int(prng.choice(legal))
            if g % 3 == 0 and (m >> 24) & 1:
                a = 25 if (m >> 25) & 1 else 24
            o.play(a)
            n += 1
    assert n > 2000
    assert {25, 34, 35} <= tokens_seen          # Healthy, NotRevealed and NoneYet all occurred in the reservation slots


def test_guess_larger_than_the_real_hand_is_an_error(orc):
    """The reference panics on `hand.len() - assumed_hands[player].len()` (usize underflow, encode_ipi.rs:158)."""
    sim = hostsim_lib.load()
    o = Fdo.new_game_philox(orc, SEED, 1, 0)
    for _ in range(30):
        m = o.allowed()
        o.play([a for a in range(39) if (m >> a) & 1][0])
    rec = np.array([o.export()], dtype=DK_STATE_DTYPE)
    cur = o.info()["current_player"]
    victim = (cur + 1) % 4
    assumed = [0, 0, 0, 0]
    assumed[victim] = hand_from_cards(list(range(12)))     # 12 guessed cards, fewer are left
    with pytest.raises(RuntimeError):
        o.encode_ipi(assumed, [0xFF] * 4, 0)
    out = np.zeros(311, dtype=np.int64)
    err = sim.sim_encode_ipi(hostsim_lib.ptr(rec), hostsim_lib.ptr(np.array(assumed, dtype=np.uint64)), hostsim_lib.ptr(np.array([0xFF] * 4, dtype=np.uint8)), 0,
                             hostsim_lib.ptr(out))
    assert err == 1
