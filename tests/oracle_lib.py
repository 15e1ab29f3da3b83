"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

SUITS = {"D": 0, "H": 1, "C": 2, "S": 3, "♦": 0, "♥": 1, "♣": 2, "♠": 3}
RANKS = {"9": 0, "10": 1, "J": 2, "Q": 3, "K": 4, "A": 5}
CARD_NAMES = [s + r for s in "DHCS" for r in ("9", "10", "J", "Q", "K", "A")]
RUST_CARD_NAMES = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
GAME_TYPES = ["Normal", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "TrumplessSolo", "QueensSolo", "JacksSolo"]
RESERVATIONS = ["Healthy", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "QueensSolo", "JacksSolo", "TrumplessSolo"]
RES_ACTION = {"Healthy": 24, "Wedding": 25, "DiamondsSolo": 26, "HeartsSolo": 27, "SpadesSolo": 28, "ClubsSolo": 29,
              "TrumplessSolo": 30, "QueensSolo": 31, "JacksSolo": 32}
ANN_ACTION = {"ReContra": 33, "No90": 34, "No60": 35, "No30": 36, "Black": 37, "NoAnnouncement": 38}
ANN_BITS = {None: 0, "ReContra": 1, "No90": 2, "No60": 4, "No30": 8, "Black": 16, "CounterReContra": 32, "NoAnnouncement": 64}
COLORS = {"Trump": 0, "Diamond": 1, "Heart": 2, "Spade": 3, "Club": 4, "T": 0, "♦": 1, "♥": 2, "♠": 3, "♣": 4}
PLAYERS = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3, "B": 0, "L": 1, "T": 2, "R": 3}


def card_id(name):
    """'D9', 'H10', '♦9', '♥10' → 0..23."""
    return SUITS[name[0]] * 6 + RANKS[name[1:]]


def hand_from_cards(cards):
    """FdoHand::from_vec / hand_from_vec: add → copy A first, then copy B."""
    h = 0
    for c in cards:
        if isinstance(c, str):
            c = card_id(c)
        if (h >> c) & 1:
            h |= 1 << (c + 24)
        else:
            h |= 1 << c
    return h


class DkState(C.Structure):
    _fields_ = [("hands", C.c_uint64 * 4), ("cards", C.c_uint8 * 48), ("announcements", C.c_uint16 * 12),
                ("reservations", C.c_uint8 * 4), ("tricks", C.c_uint32), ("eyes", C.c_uint8 * 4), ("num_tricks", C.c_uint16),
                ("card_index", C.c_uint8), ("n_reservations", C.c_uint8), ("points", C.c_int8 * 4), ("meta", C.c_uint32)]


assert C.sizeof(DkState) == 128

DK_STATE_DTYPE = np.dtype([("hands", "<u8", 4), ("cards", "u1", 48), ("announcements", "<u2", 12), ("reservations", "u1", 4),
                           ("tricks", "<u4"), ("eyes", "u1", 4), ("num_tricks", "<u2"), ("card_index", "u1"),
                           ("n_reservations", "u1"), ("points", "i1", 4), ("meta", "<u4")])
assert DK_STATE_DTYPE.itemsize == 128


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB)
            for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))):
        build()
    L = C.CDLL(LIB)
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
    L.orc_last_error.restype = C.c_char_p
    L.orc_select_by_rank.restype = u64
    L.orc_select_by_rank.argtypes = [u64, u64]
    L.orc_philox_word.restype = u32
    L.orc_philox_word.argtypes = [u64, u32, u32, u32, u32, u32]
    L.orc_philox_pair_draws.restype = None
    L.orc_philox_pair_draws.argtypes = [u64, u32, u32, u32, u32, u32, i32, vp, vp, vp, vp]
    L.orc_philox_draws.restype = None
    L.orc_philox_draws.argtypes = [u64, u32, u32, u32, u32, u32, u32, i32, vp, vp]
    for name in ("orc_fdo_new", "orc_fdo_new_game_philox", "orc_fdo_new_game_smallrng", "orc_fdo_clone", "orc_fdo_import",
                 "orc_doko_new", "orc_doko_new_game_philox", "orc_doko_new_game_smallrng_play", "orc_doko_clone",
                 "orc_fdo_with_hands_and_reservations"):
        getattr(L, name).restype = vp
    L.orc_fdo_new.argtypes = [C.POINTER(u64), i32]
    L.orc_doko_new.argtypes = [C.POINTER(u64), i32]
    L.orc_fdo_new_game_philox.argtypes = [u64, u64, u32]
    L.orc_doko_new_game_philox.argtypes = [u64, u64, u32]
    L.orc_fdo_new_game_smallrng.argtypes = [u64]
    L.orc_doko_new_game_smallrng_play.argtypes = [u64, i32]
    for name in ("orc_fdo_clone", "orc_fdo_free", "orc_doko_clone", "orc_doko_free"):
        getattr(L, name).argtypes = [vp]
    L.orc_fdo_import.argtypes = [vp]
    L.orc_fdo_play.argtypes = [vp, i32]
    L.orc_doko_play.argtypes = [vp, i32]
    L.orc_fdo_allowed.restype = u64
    L.orc_fdo_allowed.argtypes = [vp]
    L.orc_doko_allowed.restype = u64
    L.orc_doko_allowed.argtypes = [vp]
    for name in ("orc_fdo_info", "orc_fdo_hands", "orc_fdo_tricks", "orc_fdo_additional", "orc_fdo_encode_pi", "orc_fdo_export",
                 "orc_doko_info", "orc_doko_hands", "orc_doko_export"):
        getattr(L, name).argtypes = [vp, vp]
    L.orc_fdo_visible_reservations.argtypes = [vp, i32, vp]
    L.orc_doko_encode.argtypes = [vp, i32, vp]
    L.orc_fdo_random_step_philox.argtypes = [vp, u64, u64, u32, i32, u32]
    L.orc_fdo_step_site.argtypes = [vp, u64, u64, u32, i32, i32]
    L.orc_doko_sample_assignment_philox.argtypes = [vp, u64, u64, u32, u32, vp]
    L.orc_fdo_leaf_rollout_philox.argtypes = [vp, u64, u64, u32, u32, i32, vp, vp]
    L.orc_doko_random_step_philox.argtypes = [vp, u64, u64, u32]
    L.orc_fdo_hand_plus.restype = u64
    L.orc_fdo_hand_plus.argtypes = [u64, u64]
    L.orc_fdo_hand_minus.restype = u64
    L.orc_fdo_hand_minus.argtypes = [u64, u64]
    L.orc_fdo_hand_remove_color.restype = u64
    L.orc_fdo_hand_remove_color.argtypes = [u64, i32, i32]
    L.orc_fdo_hand_iter.argtypes = [u64, vp]
    L.orc_fdo_hand_op.argtypes = [C.POINTER(u64), i32, i32]
    L.orc_doko_hand_op.restype = u64
    L.orc_doko_hand_op.argtypes = [u64, i32, i32]
    L.orc_doko_allowed_actions.restype = u64
    L.orc_doko_allowed_actions.argtypes = [i32, i32, u64]
    L.orc_fdo_all_higher_than.restype = u32
    L.orc_fdo_internal_calc_allowed.restype = u32
    L.orc_fdo_internal_calc_allowed.argtypes = [i32, u32, i32, i32]
    L.orc_fdo_calc_allowed.restype = u32
    L.orc_fdo_calc_allowed.argtypes = [i32, i32, i32, i32, i32, u32, i32, i32]
    L.orc_fdo_card_matching_philox.argtypes = [vp, u64, u64, u32, u32, vp, vp]
    L.orc_fdo_is_consistent.argtypes = [vp, vp, vp]
    L.orc_fdo_with_hands_and_reservations.argtypes = [vp, vp, vp]
    L.orc_fdo_random_rollout_philox.argtypes = [vp, u64, u64, u32, u32, i32, vp, vp]
    L.orc_playout_philox.restype = dbl
    L.orc_playout_philox.argtypes = [i32, i32, u64, u32, u64, u64, i32, vp, vp, vp, vp, i32]
    L.orc_smallrng_distribute_cards.argtypes = [u64, i32, vp]
    L.orc_smallrng_bitflag_picks.argtypes = [u64, u64, i32, i32, vp]
    L.orc_smallrng_ranges.argtypes = [u64, i32, vp, vp]
    L.orc_selfplay_uniform.argtypes = [u64, u64, u64, C.c_float, u32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orc_fdo_az_allowed.restype = u64
    L.orc_fdo_az_allowed.argtypes = [vp, i32, u64]
    L.orc_fdo_encode_ipi.argtypes = [vp, vp, vp, i32, vp]
    L.orc_replay_records.argtypes = [u64, vp, vp, vp, vp]
    L.orc_fdo_uct_search_philox.argtypes = [vp, u64, u64, u32, u32, i32, u32, C.c_float, vp, vp, vp]
    L.orc_fdo_mc_allowed.restype = u64
    L.orc_fdo_mc_allowed.argtypes = [vp, i32]
    L.orc_fuse_max_n.argtypes = [vp, u64, u64]
    L.orc_fuse_average.argtypes = [vp, u64]
    L.orc_fdo_flat_mc_philox.argtypes = [vp, u64, u64, u32, u32, u32, vp, vp]
    L.orc_bulk_make.restype = vp
    L.orc_bulk_make.argtypes = [i32, u64, u64, u64, u32, i32, vp, i32]
    L.orc_bulk_free.argtypes = [vp]
    L.orc_bulk_determinize.restype = dbl
    L.orc_bulk_determinize.argtypes = [vp, u32, u64, u64, u32, u32, vp, vp, vp, vp, i32]
    L.orc_bulk_leaf_rollouts.restype = dbl
    L.orc_bulk_leaf_rollouts.argtypes = [vp, u32, u64, u64, u32, u32, i32, vp, i32]
    L.orc_bulk_step.argtypes = [vp, u64, u64, u32, i32, i32, vp, vp, vp, i32]
    L.orc_fdo_enumerate_consistent_hands.restype = C.c_int64
    L.orc_fdo_enumerate_consistent_hands.argtypes = [vp, C.c_int64, vp]
    _lib = L
    return L


# ---- convenience wrappers -------------------------------------------------------------------------------
FDO_INFO_FIELDS = ["phase", "current_player", "game_type", "card_index", "n_tricks", "team_tag", "wedding_player", "solved_idx",
                   "re_players", "re_lowest", "contra_lowest", "turns_without", "ann_start", "n_announcements", "current_allowed"]


class Fdo:
    """Handle on one oracle rs-full-doko state."""

    def __init__(self, L, handle):
        self.L, self.h = L, C.c_void_p(handle)

    @classmethod
    def from_hands(cls, L, hands, start):
        arr = (C.c_uint64 * 4)(*hands)
        return cls(L, L.orc_fdo_new(arr, start))

    @classmethod
    def new_game_philox(cls, L, seed, unit, epoch=0):
        return cls(L, L.orc_fdo_new_game_philox(seed, unit, epoch))

    @classmethod
    def from_dk_state(cls, L, rec):
        buf = np.ascontiguousarray(rec).tobytes()
        h = L.orc_fdo_import(C.c_char_p(buf))
        if not h:
            raise RuntimeError(L.orc_last_error().decode())
        return cls(L, h)

    def clone(self):
        return Fdo(self.L, self.L.orc_fdo_clone(self.h))

    def __del__(self):
        try:
            self.L.orc_fdo_free(self.h)
        except Exception:
            pass

    def play(self, action):
        if self.L.orc_fdo_play(self.h, action):
            raise RuntimeError(self.L.orc_last_error().decode())

    def allowed(self):
        return int(self.L.orc_fdo_allowed(self.h))

    def info(self):
        o = (C.c_int32 * 36)()
        self.L.orc_fdo_info(self.h, o)
        o = list(o)
        d = dict(zip(FDO_INFO_FIELDS, o[:15]))
        d["eyes"], d["num_tricks"], d["points"] = o[15:19], o[19:23], o[23:27]
        d["n_play_actions"], d["is_solo"], d["re_eyes"], d["kontra_eyes"], d["re_points"], d["kontra_points"] = o[27:33]
        return d

    def hands(self):
        o = (C.c_uint64 * 4)()
        self.L.orc_fdo_hands(self.h, o)
        return [int(x) for x in o]

    def tricks(self):
        o = (C.c_int32 * 72)()
        self.L.orc_fdo_tricks(self.h, o)
        return np.array(o, dtype=np.int32).reshape(12, 6)

    def additional(self):
        o = (C.c_int32 * 8)()
        self.L.orc_fdo_additional(self.h, o)
        return list(o)

    def encode_pi(self):
        o = (C.c_int64 * 311)()
        if self.L.orc_fdo_encode_pi(self.h, o):
            raise RuntimeError(self.L.orc_last_error().decode())
        return np.array(o, dtype=np.int64)

    def encode_ipi(self, assumed_hands, assumed_res, next_player):
        out = (C.c_int64 * 311)()
        rc = self.L.orc_fdo_encode_ipi(self.h, (C.c_uint64 * 4)(*assumed_hands), (C.c_uint8 * 4)(*assumed_res), next_player, out)
        if rc:
            raise RuntimeError(self.L.orc_last_error().decode())
        return np.array(out, dtype=np.int64)

    def export(self):
        rec = np.zeros(1, dtype=DK_STATE_DTYPE)
        self.L.orc_fdo_export(self.h, rec.ctypes.data_as(C.c_void_p))
        return rec[0]

    def random_step(self, seed, unit, epoch=0, with_announcements=True, ann_ordinal=0):
        return self.L.orc_fdo_random_step_philox(self.h, seed, unit, epoch, int(with_announcements), ann_ordinal)

    def step_site(self, seed, unit, epoch, with_announcements=True, skip_single=False):
        return self.L.orc_fdo_step_site(self.h, seed, unit, epoch, int(with_announcements), int(skip_single))

    def card_matching(self, seed, unit, sample, epoch=0):
        hands = (C.c_uint64 * 4)()
        res = (C.c_uint8 * 4)()
        st = self.L.orc_fdo_card_matching_philox(self.h, seed, unit, sample, epoch, hands, res)
        return st, [int(x) for x in hands], list(res)

    def is_consistent(self, hands, res):
        return self.L.orc_fdo_is_consistent(self.h, (C.c_uint64 * 4)(*hands), (C.c_uint8 * 4)(*res))

    def consistent_hands(self, max_members=200_000):
        """Every assignment of the hidden cards (hand sizes kept) that passes is_consistent with some hidden reservations: uint64 [k,4]."""
        out = np.zeros((max_members, 4), dtype=np.uint64)
        k = self.L.orc_fdo_enumerate_consistent_hands(self.h, max_members, out.ctypes.data_as(C.c_void_p))
        if k < 0:
            raise RuntimeError("support set larger than max_members")
        return out[:k].copy()

    def leaf_rollout(self, seed, unit, rollout, epoch=0, determinize=True):
        pts = (C.c_int32 * 4)()
        steps = C.c_uint32()
        st = self.L.orc_fdo_leaf_rollout_philox(self.h, seed, unit, rollout, epoch, int(determinize), pts, C.byref(steps))
        return st, list(pts), steps.value

    def uct_search(self, seed, unit, sub, iterations, uct_c, epoch=0, determinize=False):
        """(status, visits[39] u32, values[39] f32, action) of oracle/mcts.hpp search."""
        visits = np.zeros(39, dtype=np.uint32)
        values = np.zeros(39, dtype=np.float32)
        action = C.c_int32()
        vp = C.c_void_p
        st = self.L.orc_fdo_uct_search_philox(self.h, seed, unit, sub, epoch, int(determinize), iterations, uct_c, visits.ctypes.data_as(vp),
                                              values.ctypes.data_as(vp), C.byref(action))
        if st < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return st, visits, values, action.value

    def flat_mc(self, seed, unit, det, n_rollouts, epoch=0):
        """(status, visits[39], value_sum[39]) of oracle/pimc.hpp flat_mc."""
        visits = np.zeros(39, dtype=np.uint32)
        values = np.zeros(39, dtype=np.int64)
        st = self.L.orc_fdo_flat_mc_philox(self.h, seed, unit, det, n_rollouts, epoch, visits.ctypes.data_as(C.c_void_p), values.ctypes.data_as(C.c_void_p))
        return st, visits, values

    def rollout(self, seed, unit, rollout, epoch=0, with_announcements=False):
        pts = (C.c_int32 * 4)()
        steps = C.c_uint32()
        self.L.orc_fdo_random_rollout_philox(self.h, seed, unit, rollout, epoch, int(with_announcements), pts, C.byref(steps))
        return list(pts), steps.value


DOKO_INFO_FIELDS = ["phase", "current_player", "trick_index", "team_tag", "wedding_player", "solved_idx", "re_players"]


class Doko:
    """Handle on one oracle rs-doko state."""

    def __init__(self, L, handle):
        self.L, self.h = L, C.c_void_p(handle)

    @classmethod
    def from_hands(cls, L, hands, start):
        return cls(L, L.orc_doko_new((C.c_uint64 * 4)(*hands), start))

    @classmethod
    def new_game_philox(cls, L, seed, unit, epoch=0):
        return cls(L, L.orc_doko_new_game_philox(seed, unit, epoch))

    def __del__(self):
        try:
            self.L.orc_doko_free(self.h)
        except Exception:
            pass

    def play(self, action):
        if self.L.orc_doko_play(self.h, action):
            raise RuntimeError(self.L.orc_last_error().decode())

    def allowed(self):
        return int(self.L.orc_doko_allowed(self.h))

    def info(self):
        o = (C.c_int32 * 21)()
        self.L.orc_doko_info(self.h, o)
        o = list(o)
        d = dict(zip(DOKO_INFO_FIELDS, o[:7]))
        d["eyes"], d["num_tricks"], d["points"], d["n_play_actions"], d["start_player"] = o[7:11], o[11:15], o[15:19], o[19], o[20]
        return d

    def hands(self):
        o = (C.c_uint64 * 4)()
        self.L.orc_doko_hands(self.h, o)
        return [int(x) for x in o]

    def encode(self, with_reservations=False):
        o = (C.c_int64 * 114)()
        n = self.L.orc_doko_encode(self.h, int(with_reservations), o)
        return np.array(o[:n], dtype=np.int64)

    def export(self):
        rec = np.zeros(1, dtype=DK_STATE_DTYPE)
        self.L.orc_doko_export(self.h, rec.ctypes.data_as(C.c_void_p))
        return rec[0]

    def random_step(self, seed, unit, epoch=0):
        return self.L.orc_doko_random_step_philox(self.h, seed, unit, epoch)

    def sample_assignment(self, seed, unit, sample, epoch=0):
        hands = (C.c_uint64 * 4)()
        st = self.L.orc_doko_sample_assignment_philox(self.h, seed, unit, sample, epoch, hands)
        return st, [int(x) for x in hands]


def playout_philox(L, engine, n, seed, first_id=0, epoch=0, with_announcements=True, n_threads=0, want_aux=False, trace_stride=0):
    """Bulk oracle playouts from fresh Philox deals.  Returns dict(points, steps, aux, trace, seconds)."""
    points = np.zeros((n, 4), dtype=np.int32)
    steps = np.zeros(n, dtype=np.uint32)
    aux = np.zeros((n, 8), dtype=np.int32) if want_aux else None
    trace = np.zeros((n, trace_stride), dtype=np.uint8) if trace_stride else None
    sec = L.orc_playout_philox(engine, int(with_announcements), seed, epoch, first_id, n, n_threads,
                               points.ctypes.data_as(C.c_void_p), steps.ctypes.data_as(C.c_void_p),
                               aux.ctypes.data_as(C.c_void_p) if want_aux else None,
                               trace.ctypes.data_as(C.c_void_p) if trace_stride else None, trace_stride)
    return dict(points=points, steps=steps, aux=aux, trace=trace, seconds=sec)


def fuse(L, strategy, visits, allowed_mask=0):
    """PolicyFusionMaxN (strategy 0) / PolicyFusionAverageStrategy (1) over rows [n][39] of successful samples (oracle/pimc.hpp)."""
    v = np.ascontiguousarray(visits, dtype=np.uint32).reshape(-1, 39)
    if strategy == 0:
        return L.orc_fuse_max_n(v.ctypes.data_as(C.c_void_p), v.shape[0], allowed_mask)
    return L.orc_fuse_average(v.ctypes.data_as(C.c_void_p), v.shape[0])


def selfplay_uniform(L, seed, unit, az_epoch, keep_prob, first_epoch=0, max_rows=250):
    """oracle/selfplay.hpp self_play of one Philox-dealt game with the uniform stand-in search.
    Returns dict(states [r,311] i64, policy [r,39] f32, value [r,4] f32, player, turn, forced, turns, points)."""
    st = np.zeros((max_rows, 311), dtype=np.int64)
    po = np.zeros((max_rows, 39), dtype=np.float32)
    va = np.zeros((max_rows, 4), dtype=np.float32)
    pl = np.zeros(max_rows, dtype=np.uint8)
    tu = np.zeros(max_rows, dtype=np.uint16)
    fo = np.zeros(max_rows, dtype=np.uint8)
    turns = C.c_uint32()
    pts = (C.c_int32 * 4)()
    vp = C.c_void_p
    n = L.orc_selfplay_uniform(seed, unit, az_epoch, keep_prob, first_epoch, max_rows, st.ctypes.data_as(vp), po.ctypes.data_as(vp), va.ctypes.data_as(vp),
                               pl.ctypes.data_as(vp), tu.ctypes.data_as(vp), fo.ctypes.data_as(vp), C.byref(turns), pts)
    assert 0 <= n <= max_rows, L.orc_last_error()
    return {"states": st[:n], "policy": po[:n], "value": va[:n], "player": pl[:n], "turn": tu[:n], "forced": fo[:n], "turns": turns.value,
            "points": list(pts)}


def replay_records(L, states, value, policy):
    """bincode DBRecord bytes [n, 2684] (oracle/replay.hpp)."""
    st = np.ascontiguousarray(states, dtype=np.int64)
    va = np.ascontiguousarray(value, dtype=np.float32)
    po = np.ascontiguousarray(policy, dtype=np.float32)
    out = np.zeros((st.shape[0], 2684), dtype=np.uint8)
    vp = C.c_void_p
    L.orc_replay_records(st.shape[0], st.ctypes.data_as(vp), va.ctypes.data_as(vp), po.ctypes.data_as(vp), out.ctypes.data_as(vp))
    return out


class Bulk:
    """A batch of oracle states made by a recipe (oracle/capi.cpp orc_bulk_make) + their dk_state records (self.recs, numpy [n] of
    DK_STATE_DTYPE) for the GPU.  mode 0 = BASELINE config 3 mid-game states, mode 1 = states at every stage of a game (the reference's
    determinization soak, rs-full-doko-cmd/src/main.rs:190-283)."""

    def __init__(self, L, engine, n, seed, first_id=0, epoch=0, mode=0, n_threads=0):
        self.L, self.engine, self.n, self.seed, self.first_id = L, engine, n, seed, first_id
        self.recs = np.zeros(n, dtype=DK_STATE_DTYPE)
        self.h = C.c_void_p(L.orc_bulk_make(engine, n, seed, first_id, epoch, mode, self.recs.ctypes.data_as(C.c_void_p), n_threads))

    def __del__(self):
        try:
            self.L.orc_bulk_free(self.h)
        except Exception:
            pass

    def bytes(self):
        """uint8 [n,128] copy of the records (what the tests upload)."""
        return np.frombuffer(self.recs.tobytes(), dtype=np.uint8).reshape(self.n, 128).copy()

    def determinize(self, S, epoch, first_sub=0, want_consistent=True, n_threads=0):
        """(hands u64 [n,S,4], res u8 [n,S,4], status u8 [n,S], consistent i8 [n,S] or None, seconds)."""
        hands = np.zeros((self.n, S, 4), dtype=np.uint64)
        res = np.full((self.n, S, 4), 0xFF, dtype=np.uint8)
        status = np.zeros((self.n, S), dtype=np.uint8)
        cons = np.zeros((self.n, S), dtype=np.int8) if (want_consistent and self.engine == 1) else None
        vp = C.c_void_p
        sec = self.L.orc_bulk_determinize(self.h, S, self.seed, self.first_id, epoch, first_sub, hands.ctypes.data_as(vp), res.ctypes.data_as(vp),
                                          status.ctypes.data_as(vp), cons.ctypes.data_as(vp) if cons is not None else None, n_threads)
        return hands, res, status, cons, sec

    def leaf_rollouts(self, R, epoch, first_sub=0, determinize=True, n_threads=0):
        sums = np.zeros((self.n, 4), dtype=np.int64)
        sec = self.L.orc_bulk_leaf_rollouts(self.h, R, self.seed, self.first_id, epoch, first_sub, int(determinize), sums.ctypes.data_as(C.c_void_p), n_threads)
        return sums, sec

    def step(self, epoch, with_announcements=True, skip_single=False, want_recs=False, want_obs=False, n_threads=0):
        """One lock-step env step of every state (dk_step_random_encode): (actions u8 [n], recs or None, obs i64 [n,311] or None)."""
        act = np.zeros(self.n, dtype=np.uint8)
        recs = np.zeros(self.n, dtype=DK_STATE_DTYPE) if want_recs else None
        obs = np.zeros((self.n, 311), dtype=np.int64) if want_obs else None
        vp = C.c_void_p
        self.L.orc_bulk_step(self.h, self.seed, self.first_id, epoch, int(with_announcements), int(skip_single), act.ctypes.data_as(vp),
                             recs.ctypes.data_as(vp) if want_recs else None, obs.ctypes.data_as(vp) if want_obs else None, n_threads)
        return act, recs, obs


def fx_hash_record(rec_bytes, last_action=0xFF):
    """FxHasher64 (fxhash 0.2.1, the hasher of AzEnvState::id(), rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:156-167) over the
    sixteen little-endian u64 words of a 128-byte record and the last action: hash = (rotl(hash, 5) ^ word) * 0x517cc1b727220a95."""
    words = list(np.frombuffer(bytes(rec_bytes), dtype="<u8")) + [last_action]
    h, M = 0, (1 << 64) - 1
    for w in words:
        h = ((((h << 5) | (h >> 59)) & M) ^ int(w)) * 0x517CC1B727220A95 & M
    return h
