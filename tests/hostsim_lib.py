"""ctypes binding of tests/hostsim/libhostsim.so — TEST INFRASTRUCTURE ONLY (CPU execution of the per-thread kernel logic)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.path.join(HERE, "hostsim")
LIB = os.path.join(DIR, "libhostsim.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", DIR])
        # DK_HOSTSIM_LIB: swap in an instrumented build (make -C tests/hostsim asan; run with LD_PRELOAD=libasan) — the
        # bounds/UB check of the per-thread kernel logic on this pool, where compute-sanitizer is not available.
        L = C.CDLL(os.environ.get("DK_HOSTSIM_LIB", LIB))
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.sim_legal_mask.restype = u64
        L.sim_legal_mask.argtypes = [i32, vp]
        L.sim_apply.restype = u32
        L.sim_apply.argtypes = [i32, vp, u32, u32]
        L.sim_new_game.argtypes = [vp, vp, u32]
        L.sim_encode.argtypes = [i32, vp, vp]
        L.sim_playout_from_state.argtypes = [i32, vp, u64, u64, u32, u32, i32, vp, vp]
        L.sim_fdo_playout_fresh.argtypes = [u64, u32, u64, u64, i32, vp, vp]
        L.sim_doko_playout_fresh.argtypes = [u64, u32, u64, u64, vp, vp, vp, vp]
        L.sim_fdo_determinize.restype = u32
        L.sim_fdo_determinize.argtypes = [vp, u64, u64, u32, u32, vp, vp]
        L.sim_doko_assign.restype = u32
        L.sim_doko_assign.argtypes = [vp, u64, u64, u32, u32, vp]
        L.sim_fdo_leaf_rollout.restype = u32
        L.sim_fdo_leaf_rollout.argtypes = [vp, u64, u64, u32, u32, i32, vp, vp]
        L.sim_sp_az_allowed.restype = u64
        L.sim_sp_az_allowed.argtypes = [vp, u64]
        L.sim_sp_keep_draw.restype = C.c_float
        L.sim_sp_keep_draw.argtypes = [u32]
        L.sim_sp_value_target.restype = C.c_float
        L.sim_sp_value_target.argtypes = [vp, u32, u32]
        L.sim_encode_ipi.restype = u32
        L.sim_encode_ipi.argtypes = [vp, vp, vp, u32, vp]
        L.sim_fdo_uct_search.restype = u32
        L.sim_fdo_uct_search.argtypes = [vp, u64, u64, u32, u32, i32, u32, C.c_float, vp, vp, vp]
        L.sim_uct_select_check.restype = u32
        L.sim_uct_select_check.argtypes = [u32, vp, vp, u32, C.c_float]
        L.sim_uct_allowed.restype = u64
        L.sim_uct_allowed.argtypes = [vp, i32]
        L.sim_fuse.restype = u32
        L.sim_fuse.argtypes = [i32, vp, vp, u32, u64, vp]
        L.sim_root_stats.argtypes = [vp, vp, u32, u64, vp]
        L.sim_root_pick.restype = u32
        L.sim_root_pick.argtypes = [u32, vp, u64]
        L.sim_fdo_flat_mc.restype = u32
        L.sim_fdo_flat_mc.argtypes = [vp, u64, u64, u32, u32, u32, vp, vp]
        L.sim_fdo_score.restype = i32
        L.sim_fdo_score.argtypes = [u32, u32, u32, u32, u32, i32, vp]
        L.sim_fdo_final_points_mismatches.restype = u64
        L.sim_fdo_final_points_mismatches.argtypes = [u64]
        L.sim_fdo_score_mismatches.restype = u64
        L.sim_fdo_score_mismatches.argtypes = []
        L.sim_fdo_eligible_nibble.restype = u32
        L.sim_fdo_eligible_nibble.argtypes = [vp, u32, u32, u32]
        for n in ("sim_fdo_min_cards_to_call", "sim_fdo_allowed_call", "sim_select_lsb24", "sim_select_lsb", "sim_card_power", "sim_trump_mask", "sim_follow_mask",
                  "sim_pick_msb_rank24_tab", "sim_pick_msb_rank24_lut", "sim_pick_msb_rank24", "sim_pow_lookup", "sim_seg_lut", "sim_thr2_lut"):
            getattr(L, n).restype = u32
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
