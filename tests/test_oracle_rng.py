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
