"""RNG pins: Philox4x32-10 against the Random123 known-answer vectors; oracle Philox stream == device Philox (hostsim)."""
import ctypes as C

import numpy as np

import hostsim_lib

KAT = [  # Random123 kat_vectors, philox4x32 10 rounds: counter, key, expected
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers(orc):
    for ctr, key, exp in KAT:
        out = (C.c_uint32 * 4)()
        orc.orc_philox_block((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert list(out) == exp


def test_philox_word_layout(orc):
    """word k of a site = word (k & 3) of block (k >> 2) with counter (unit_lo, unit_hi, site << 16 | block, epoch)."""
    seed, unit_lo, unit_hi, epoch = 0x0123456789ABCDEF, 77, 5, 9
    for site in range(8):
        for k in range(13):
            out = (C.c_uint32 * 4)()
            orc.orc_philox_block((C.c_uint32 * 4)(unit_lo, unit_hi, (site << 16) | (k >> 2), epoch),
                                 (C.c_uint32 * 2)(seed & 0xFFFFFFFF, seed >> 32), out)
            assert orc.orc_philox_word(seed, unit_lo, unit_hi, epoch, site, k) == out[k & 3]


def _draws(orc, seed, unit_lo, unit_hi, epoch, site, first, chain_mul, ns):
    n = np.array(ns, dtype=np.uint32)
    out = np.zeros(len(ns), dtype=np.uint32)
    orc.orc_philox_draws(seed, unit_lo, unit_hi, epoch, site, first, chain_mul, len(ns), n.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return [int(x) for x in out]


def test_chained_draws_closed_form(orc):
    """DESIGN.md "Philox parity contract", written out independently of the oracle's stream object:
    deal  — draw s (0 = start seat, s >= 1 the shuffle step i = 48 - s) takes word min(s // 3, 11) for s <= 36, chained;
    cards — the draw at card_index ci takes word ci // 4 (its trick), the trick's four draws chained; a stream positioned inside a trick
            continues at word * (product of the counts already drawn);
    rule 4 of card_matching — its k-th use takes word k // 3, chained;
    a chained draw over n after the counts n_0..n_{k-1}: idx = ((word * n_0 * .. * n_{k-1} mod 2^32) * n) >> 32;
    the other sites take one word per draw: idx = (word * n) >> 32."""
    seed, ul, uh, ep = 0x0123456789ABCDEF, 4711, 3, 6
    word = lambda site, k: orc.orc_philox_word(seed, ul, uh, ep, site, k)
    M = 1 << 32
    # deal (site 0): the 37 draws that decide the hands
    ns = [4] + [49 - s for s in range(1, 37)]
    exp, mul, cur = [], 1, None
    for s, n in enumerate(ns):
        w = min(s // 3, 11)
        if w != cur:
            cur, mul = w, 1
        exp.append((((word(0, w) * mul) % M) * n) >> 32)
        mul = (mul * n) % M
    assert _draws(orc, seed, ul, uh, ep, 0, 0, 1, ns) == exp
    assert all(0 <= e < n for e, n in zip(exp, ns))
    # cards (site 3): 11 tricks of four draws with ragged counts; then the same stream entered inside every trick
    rng = np.random.default_rng(3)
    ns = [int(x) for x in rng.integers(1, 13, size=44)]
    exp = []
    for ci, n in enumerate(ns):
        mul = int(np.prod(ns[ci & ~3:ci], dtype=np.uint64)) % M if ci & 3 else 1
        exp.append((((word(3, ci >> 2) * mul) % M) * n) >> 32)
    assert _draws(orc, seed, ul, uh, ep, 3, 0, 1, ns) == exp
    for first in range(44):
        mul = int(np.prod(ns[first & ~3:first], dtype=np.uint64)) if first & 3 else 1
        assert _draws(orc, seed, ul, uh, ep, 3, first, mul, ns[first:]) == exp[first:], first
    # card_matching's rule 4 (site 4): use k takes word k // 3, three chained draws per word
    ns = [int(x) for x in rng.integers(1, 37, size=20)]
    exp = []
    for k, n in enumerate(ns):
        mul = int(np.prod(ns[k - k % 3:k], dtype=np.uint64)) % M
        exp.append((((word(4, k // 3) * mul) % M) * n) >> 32)
    assert _draws(orc, seed, ul, uh, ep, 4, 0, 1, ns) == exp
    # the rs-doko sampler (site 6): one word per card — the card draw, then the seat draw chained to it
    n0 = np.array(rng.integers(1, 37, size=15), dtype=np.uint32)
    n1 = np.array(rng.integers(1, 4, size=15), dtype=np.uint32)
    o0, o1 = np.zeros(15, dtype=np.uint32), np.zeros(15, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    orc.orc_philox_pair_draws(seed, ul, uh, ep, 6, 0, 15, p(n0), p(n1), p(o0), p(o1))
    for k in range(15):
        w = word(6, k)
        assert int(o0[k]) == (w * int(n0[k])) >> 32 and int(o1[k]) == ((w * int(n0[k])) % M * int(n1[k])) >> 32
    # one word per draw everywhere else (reservations, hidden reservations, the rs-doko sampler's card draw — its seat draw is chained to
    # the card's word: tests/test_hostsim_assignment.py pins it against the device logic —, lock-step step, keep, expand)
    for site in (1, 5, 6, 7, 8, 9):
        ns = [int(x) for x in rng.integers(1, 40, size=9)]
        assert _draws(orc, seed, ul, uh, ep, site, 2, 1, ns) == [(word(site, 2 + k) * n) >> 32 for k, n in enumerate(ns)]


def test_chained_draws_are_jointly_uniform():
    """The chain's claim: the draws taken from one word are independent and uniform up to n_0 * .. * n_k / 2^32.  Counted over a
    regular grid of 2^24 words for a trick-like chain (12, 11, 7, 5) and a deal-like chain (48, 47, 46): every joint cell holds its
    share of the grid to within 1.5 % or 2.5 words (the grid has 3631 resp. 162 words per cell; a wrong chain — the same word reused, or a
    fixed multiplier instead of the drawn count — misses by tens of percent)."""
    w = (np.arange(1 << 24, dtype=np.uint64) * np.uint64(256) + np.uint64(97))
    for ns in ((12, 11, 7, 5), (48, 47, 46)):
        v, cell = w.copy(), np.zeros(len(w), dtype=np.int64)
        for n in ns:
            p = v * np.uint64(n)
            cell = cell * n + (p >> np.uint64(32)).astype(np.int64)
            v = p & np.uint64(0xFFFFFFFF)
        cells = int(np.prod(ns))
        counts = np.bincount(cell, minlength=cells)
        exp = len(w) / cells
        assert len(counts) == cells and abs(counts - exp).max() <= max(0.015 * exp, 2.5), (ns, counts.min(), counts.max(), exp)


def test_deals_of_the_stream_are_uniform(orc):
    """40 000 deals from the parity stream (three chained draws per word): the start seat is uniform and every (seat, card type) holds
    0.5 copies on average (hypergeometric variance 12 * 2/48 * 46/48 * 36/47), all 100 marginals within 4.5 sigma."""
    from oracle_lib import Fdo

    n = 40_000
    cnt, start = np.zeros((4, 24)), np.zeros(4)
    for i in range(n):
        o = Fdo.new_game_philox(orc, 0xD0C05EED, 10_000_000 + i, 7)
        start[o.info()["current_player"]] += 1
        for s, h in enumerate(o.hands()):
            x = int(h)
            cnt[s] += [((x >> c) & 1) + ((x >> (24 + c)) & 1) for c in range(24)]
    assert np.abs(cnt / n - 0.5).max() < 4.5 * np.sqrt(0.36702 / n)
    assert np.abs(start / n - 0.25).max() < 4.5 * np.sqrt(0.25 * 0.75 / n)


def test_smallrng_stale_vectors_are_documented(orc):
    """The three rs-doko SmallRng vectors recorded under rand 0.9.0-alpha.2 do NOT hold under the pinned rand 0.9.0:
    they contradict the two vectors that do reproduce (tests/test_oracle_encoders.py).  Record what rand 0.9.0 semantics give."""
    hands = (C.c_uint64 * 4)()
    orc.orc_smallrng_distribute_cards(42, 0, hands)
    stale = [0b000000000001000000000010001101001111001000011, 0b100000000000000000100001100001110100010100110,
             0b000000000000000000000111011010000010111001001, 0b000000000000000010000100010110111000100111000]
    assert [int(h) for h in hands] != stale                      # rs-doko/src/hand/hand_random.rs:68-83 (stale)
    # ... while the SAME shuffle reproduces the rs-full-doko vector (hand.rs:559-585), because both crates share the deal code path
    assert [int(h) for h in hands] == [0b0000000010000000000000110100100110000110110001, 0b0001000000100001000000000001010101100001100101,
                                       0b1000000000000000001000011110110000000000011110, 0b0000001000011000000000100010001001011110000010]


def test_announcement_bit_stream_window(orc):
    """Device window extraction (incl. windows that straddle a 128-decision Philox block, which real games never reach) ==
    bit k of the stream = bit (k & 31) of word (k >> 5)."""
    sim = hostsim_lib.load()
    sim.sim_fdo_ann_bits.restype = C.c_uint32
    sim.sim_fdo_ann_bits.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    seed, unit, epoch = 0xABCDEF0123, 4242, 3
    bit = lambda k: (orc.orc_philox_word(seed, unit, 0, epoch, 2, k >> 5) >> (k & 31)) & 1
    for ord_ in list(range(0, 140)) + [250, 253, 254, 255, 256, 381, 383]:
        for m in (1, 2, 3, 4):
            exp = sum(bit(ord_ + i) << i for i in range(m))
            assert sim.sim_fdo_ann_bits(seed, unit, epoch, ord_, m) == exp, (ord_, m)
