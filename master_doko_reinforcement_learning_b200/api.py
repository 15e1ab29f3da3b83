"""ctypes binding of libdoko_cuda.so + thin torch-tensor helpers.  Plumbing only: all game logic runs in CUDA."""
import ctypes as C
import os

import numpy as np

from . import _build

DK_DOKO, DK_FDO = 0, 1
DK_PLAYOUT_WITH_ANNOUNCEMENTS = 1
DK_APPLY_SKIP_SINGLE = 1
N_ACTIONS = 39          # FdoAction::COUNT
ACTION_NONE = 0xFF
FUSE_MAX_N, FUSE_AVERAGE = 0, 1
ROOT_STATS = 80
REPLAY_RECORD_BYTES = 2684
DK_LAYOUT_DO110, DK_LAYOUT_DO114, DK_LAYOUT_FDO_PI311 = 0, 1, 2
OBS_LEN = {DK_LAYOUT_DO110: 110, DK_LAYOUT_DO114: 114, DK_LAYOUT_FDO_PI311: 311}
STATUS = {0: "DK_OK", 1: "DK_ERR_INVALID_ARGUMENT", 2: "DK_ERR_CUDA", 3: "DK_ERR_NO_DEVICE", 4: "DK_ERR_NCCL", 5: "DK_ERR_UNSUPPORTED"}

# numpy view of include/doko_cuda.h:dk_state (128 bytes)
DK_STATE_DTYPE = np.dtype([("hands", "<u8", 4), ("cards", "u1", 48), ("announcements", "<u2", 12), ("reservations", "u1", 4),
                           ("tricks", "<u4"), ("eyes", "u1", 4), ("num_tricks", "<u2"), ("card_index", "u1"),
                           ("n_reservations", "u1"), ("points", "i1", 4), ("meta", "<u4")])
assert DK_STATE_DTYPE.itemsize == 128


class DokoCudaError(RuntimeError):
    pass


class DkRng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_id", C.c_uint64), ("epoch", C.c_uint32), ("first_sub", C.c_uint32)]


class DkPlayoutStats(C.Structure):
    """include/doko_cuda.h:dk_playout_stats (2160 bytes)."""
    _fields_ = [("games", C.c_uint64), ("game_steps", C.c_uint64), ("point_sum", C.c_int64 * 4), ("point_sq_sum", C.c_uint64 * 4),
                ("wins", C.c_uint64 * 4), ("step_hist", C.c_uint64 * 256)]

    def as_dict(self):
        return {"games": int(self.games), "game_steps": int(self.game_steps), "point_sum": [int(x) for x in self.point_sum],
                "point_sq_sum": [int(x) for x in self.point_sq_sum], "wins": [int(x) for x in self.wins],
                "step_hist": np.array(self.step_hist, dtype=np.uint64)}


assert C.sizeof(DkPlayoutStats) == 2160
STATS_WORDS = 270
AZ_MIN_EPOCH = 10


def unpack_points(packed):
    """dk_unpack_points for a uint16 numpy array [n] -> int32 [n,4] (the packed host-transfer form of dk_playout_host_packed)."""
    v = np.asarray(packed, dtype=np.uint16)
    a = (v & 0xFF).astype(np.int8).astype(np.int32)
    same = (v >> 8) & 7
    k = 1 + (same & 1) + ((same >> 1) & 1) + ((same >> 2) & 1)
    k = k.astype(np.int32)
    b = np.where(k == 4, a, -(k * a) // np.maximum(4 - k, 1))
    out = np.empty(v.shape + (4,), dtype=np.int32)
    out[..., 0] = a
    for j in (1, 2, 3):
        out[..., j] = np.where((same >> (j - 1)) & 1, a, b)
    return out


def library_path():
    """In-tree libdoko_cuda.so; DOKO_CUDA_LIB points at another build of the same ABI (used by the tuning experiments)."""
    return os.environ.get("DOKO_CUDA_LIB") or _build.LIB_PATH


_LIB = None


def load_library():
    """dlopen libdoko_cuda.so.  Fails loudly when it is missing — there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise DokoCudaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(needs nvcc); master_doko_reinforcement_learning_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, sz, u32, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
    L.dk_version.restype = C.c_char_p
    L.dk_last_error.restype = C.c_char_p
    L.dk_last_error.argtypes = [vp]
    L.dk_init.argtypes = [i32, C.POINTER(vp)]
    L.dk_destroy.argtypes = [vp]
    L.dk_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)]
    L.dk_synchronize.argtypes = [vp, vp]
    L.dk_uct_workspace_bytes.restype = sz
    L.dk_uct_workspace_bytes.argtypes = [sz, sz]
    L.dk_launch_count.restype = u64
    L.dk_launch_count.argtypes = [vp]
    L.dk_playout.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), vp, vp, vp]
    L.dk_playout_host.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), vp, vp]
    L.dk_playout_host_compact.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), vp, vp]
    L.dk_playout_host_packed.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), vp, vp]
    L.dk_playout_summary.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), vp, i32, vp]
    L.dk_playout_summary_host.argtypes = [vp, i32, u32, sz, vp, C.POINTER(DkRng), C.POINTER(DkPlayoutStats)]
    L.dk_legal_mask_az.argtypes = [vp, sz, vp, i32, u64, vp, vp, vp]
    L.dk_state_id.argtypes = [vp, sz, vp, vp, vp, vp]
    L.dk_random_action.argtypes = [vp, i32, sz, vp, C.POINTER(DkRng), u32, vp, vp]
    L.dk_playout_trace.argtypes = [vp, i32, sz, C.POINTER(DkRng), vp, vp, vp, vp]
    for name, args in (
        ("dk_new_games", [vp, i32, sz, C.POINTER(DkRng), vp, vp]),
        ("dk_from_deals", [vp, i32, sz, vp, vp, vp, vp]),
        ("dk_legal_mask", [vp, i32, sz, vp, vp, vp]),
        ("dk_apply", [vp, i32, sz, vp, vp, u32, vp, vp]),
        ("dk_terminal", [vp, i32, sz, vp, vp, vp, vp]),
        ("dk_encode", [vp, i32, sz, vp, vp, sz, vp]),
        ("dk_step_random_encode", [vp, sz, vp, C.POINTER(DkRng), u32, vp, sz, vp, vp]),
        ("dk_encode_narrow", [vp, i32, i32, sz, vp, vp, vp]),
        ("dk_step_random_encode_narrow", [vp, sz, vp, C.POINTER(DkRng), u32, i32, vp, vp, vp]),
        ("dk_determinize", [vp, i32, sz, sz, vp, C.POINTER(DkRng), vp, vp, vp, vp]),
        ("dk_leaf_rollouts", [vp, sz, sz, i32, vp, C.POINTER(DkRng), vp, vp]),
        ("dk_encode_ipi", [vp, sz, vp, vp, vp, vp, vp, sz, vp, vp]),
        ("dk_pimc_evaluate", [vp, sz, sz, sz, vp, C.POINTER(DkRng), vp, vp, vp, vp]),
        ("dk_fuse", [vp, i32, sz, sz, vp, vp, vp, vp, vp, vp]),
        ("dk_pimc_root_stats", [vp, sz, sz, vp, vp, vp, vp, i32, vp]),
        ("dk_pimc_pick", [vp, i32, sz, vp, vp, vp, vp]),
        ("dk_sp_create", [vp, sz, vp, C.POINTER(vp)]),
        ("dk_sp_destroy", [vp]),
        ("dk_sp_reset", [vp, vp]),
        ("dk_sp_begin_turn", [vp, sz, vp, u64, C.c_float, u32, C.POINTER(DkRng), vp]),
        ("dk_sp_turn_view", [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        ("dk_sp_uniform_search", [vp, C.POINTER(DkRng), vp, vp, vp]),
        ("dk_sp_end_turn", [vp, vp, vp, vp, vp, vp]),
        ("dk_sp_finalize", [vp, vp, vp]),
        ("dk_sp_counts", [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), vp]),
        ("dk_uct_search", [vp, sz, sz, i32, sz, C.c_float, vp, C.POINTER(DkRng), vp, sz, vp, vp, vp, vp, vp]),
        ("dk_pack_replay_records", [vp, sz, vp, vp, vp, vp, vp]),
        ("dk_comm_unique_id", [vp, vp]),
        ("dk_comm_init", [vp, i32, i32, vp]),
        ("dk_comm_destroy", [vp]),
        ("dk_allreduce_root_stats", [vp, sz, vp, vp]),
    ):
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    _LIB = L
    return L


def _ptr(t):
    """Device/host pointer of a torch tensor, numpy array or None.  The C ABI takes dense buffers: a non-contiguous view would be read
    and written with the wrong strides, so it is refused here (alignment is checked by the library itself)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        if not t.flags["C_CONTIGUOUS"]:
            raise DokoCudaError("non-contiguous numpy array passed to the C ABI")
        return t.ctypes.data_as(C.c_void_p)
    if not t.is_contiguous():
        raise DokoCudaError("non-contiguous tensor passed to the C ABI")
    return C.c_void_p(t.data_ptr())


class DokoCuda:
    """One context per process/GPU (dk_init).  Methods mirror the C ABI one to one."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = C.c_void_p()
        st = self.L.dk_init(int(device), C.byref(self.ctx))
        if st != 0:
            raise DokoCudaError(f"dk_init(device={device}) failed with {STATUS.get(st, st)}: an sm_100 (B200) GPU is required; "
                                "there is no CPU fallback")
        self.device = int(device)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.dk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, what):
        if st != 0:
            raise DokoCudaError(f"{what}: {STATUS.get(st, st)}: {self.L.dk_last_error(self.ctx).decode()}")

    @staticmethod
    def rng(seed, first_id=0, epoch=0, first_sub=0):
        return DkRng(int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_id), int(epoch), int(first_sub))

    def _stream(self):
        import torch

        # torch's default stream is the legacy NULL stream; the C ABI reserves NULL for "the context's own stream",
        # so pass the explicit cudaStreamLegacy handle (0x1) to stay ordered with torch work and torch CUDA events.
        # The stream is the current one of THIS context's device (not of torch's current device).
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream or 1)

    def _on_device(self, *tensors):
        """Device tensors handed to the library must live on the context's GPU."""
        for t in tensors:
            if t is not None and not isinstance(t, np.ndarray) and (not t.is_cuda or t.device.index != self.device):
                raise DokoCudaError(f"tensor on {t.device} passed to the context of cuda:{self.device}")

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self.L.dk_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), "dk_device_info")
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    def launch_count(self):
        return int(self.L.dk_launch_count(self.ctx))

    def synchronize(self, stream=None):
        """Waits for the work the wrapper queued: by default the torch stream every method launches on."""
        self._check(self.L.dk_synchronize(self.ctx, stream if stream is not None else self._stream()), "dk_synchronize")

    # ---- playouts ----------------------------------------------------------------------------------------
    def playout(self, engine, n, rng, states=None, flags=0, points_out=None, steps_out=None, stream=None):
        """Device-resident playouts (dk_playout).  points_out int32 [n,4], steps_out uint32/int32 [n] (torch, cuda)."""
        import torch

        dev = torch.device("cuda", self.device)
        if points_out is None:
            points_out = torch.empty((n, 4), dtype=torch.int32, device=dev)
        if steps_out is None:
            steps_out = torch.empty((n,), dtype=torch.int32, device=dev)
        self._check(self.L.dk_playout(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), _ptr(points_out), _ptr(steps_out),
                                      stream if stream is not None else self._stream()), "dk_playout")
        return points_out, steps_out

    def playout_host(self, engine, n, rng, states=None, flags=0, points_out=None, steps_out=None):
        """Host-buffer playouts (dk_playout_host): numpy in/out, copies inside the call."""
        if points_out is None:
            points_out = np.empty((n, 4), dtype=np.int32)
        if steps_out is None:
            steps_out = np.empty((n,), dtype=np.uint32)
        self._check(self.L.dk_playout_host(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), _ptr(points_out), _ptr(steps_out)),
                    "dk_playout_host")
        return points_out, steps_out

    def playout_host_compact(self, engine, n, rng, states=None, flags=0, points_out=None, steps_out=None):
        """dk_playout_host_compact: int8 [n,4] points + uint8 [n] steps in host buffers (numpy or pinned torch tensors)."""
        if points_out is None:
            points_out = np.empty((n, 4), dtype=np.int8)
        if steps_out is None:
            steps_out = np.empty((n,), dtype=np.uint8)
        self._check(self.L.dk_playout_host_compact(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), _ptr(points_out), _ptr(steps_out)),
                    "dk_playout_host_compact")
        return points_out, steps_out

    def playout_host_packed(self, engine, n, rng, states=None, flags=0, points_out=None, steps_out=None, want_steps=True):
        """dk_playout_host_packed: uint16 [n] packed points (+ uint8 [n] steps) in host buffers; decode with unpack_points."""
        if points_out is None:
            points_out = np.empty((n,), dtype=np.uint16)
        if steps_out is None and want_steps:
            steps_out = np.empty((n,), dtype=np.uint8)
        self._check(self.L.dk_playout_host_packed(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), _ptr(points_out), _ptr(steps_out)),
                    "dk_playout_host_packed")
        return points_out, steps_out

    def playout_summary(self, engine, n, rng, states=None, flags=0, stats=None, accumulate=False, stream=None):
        """dk_playout_summary: device-reduced statistics as an int64 cuda tensor [270] (word layout of dk_playout_stats)."""
        import torch

        self._on_device(states, stats)
        if stats is None:
            stats = torch.zeros((STATS_WORDS,), dtype=torch.int64, device=self._dev())
        self._check(self.L.dk_playout_summary(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), _ptr(stats), int(accumulate),
                                              stream if stream is not None else self._stream()), "dk_playout_summary")
        return stats

    def playout_summary_host(self, engine, n, rng, states=None, flags=0):
        """dk_playout_summary_host: the batch's statistics as a DkPlayoutStats struct in host memory (2160 bytes cross PCIe)."""
        out = DkPlayoutStats()
        self._check(self.L.dk_playout_summary_host(self.ctx, engine, flags, n, _ptr(states), C.byref(rng), C.byref(out)), "dk_playout_summary_host")
        return out

    # ---- state records ---------------------------------------------------------------------------------------------------
    def _dev(self):
        import torch

        return torch.device("cuda", self.device)

    def alloc_states(self, n):
        """Device buffer of n dk_state records (torch uint8 [n,128])."""
        import torch

        return torch.empty((n, 128), dtype=torch.uint8, device=self._dev())

    def new_games(self, engine, n, rng, out=None, stream=None):
        out = self.alloc_states(n) if out is None else out
        self._check(self.L.dk_new_games(self.ctx, engine, n, C.byref(rng), _ptr(out), stream if stream is not None else self._stream()), "dk_new_games")
        return out

    def from_deals(self, engine, hands, start, out=None, stream=None):
        """hands: int64/uint64 cuda tensor [n,4]; start: uint8 cuda tensor [n]."""
        n = hands.shape[0]
        out = self.alloc_states(n) if out is None else out
        self._check(self.L.dk_from_deals(self.ctx, engine, n, _ptr(hands), _ptr(start), _ptr(out), stream if stream is not None else self._stream()),
                    "dk_from_deals")
        return out

    def legal_mask(self, engine, states, out=None, stream=None):
        import torch

        n = states.shape[0]
        out = torch.empty((n,), dtype=torch.int64, device=self._dev()) if out is None else out
        self._check(self.L.dk_legal_mask(self.ctx, engine, n, _ptr(states), _ptr(out), stream if stream is not None else self._stream()), "dk_legal_mask")
        return out

    def legal_mask_az(self, states, is_secondary, az_epoch, want_count=True, stream=None):
        """AzEnvState::allowed_actions_by_action_index(is_secondary, epoch) as masks (int64 [n]) and number_of_allowed_actions(epoch) (uint8 [n])."""
        import torch

        self._on_device(states)
        n = states.shape[0]
        mask = torch.empty((n,), dtype=torch.int64, device=self._dev())
        cnt = torch.empty((n,), dtype=torch.uint8, device=self._dev()) if want_count else None
        self._check(self.L.dk_legal_mask_az(self.ctx, n, _ptr(states), int(bool(is_secondary)), int(az_epoch), _ptr(mask), _ptr(cnt),
                                            stream if stream is not None else self._stream()), "dk_legal_mask_az")
        return mask, cnt

    def random_action(self, engine, states, rng, flags=0, stream=None):
        """FdoAllowedActions::random over the legal set without playing it (uint8 [n]; ACTION_NONE for finished games)."""
        import torch

        self._on_device(states)
        n = states.shape[0]
        out = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_random_action(self.ctx, engine, n, _ptr(states), C.byref(rng), flags, _ptr(out),
                                            stream if stream is not None else self._stream()), "dk_random_action")
        return out

    def state_id(self, states, last_action=None, stream=None):
        """AzEnvState::id(): FxHasher64 over the record (+ last action), int64 [n] (bit pattern of the u64)."""
        import torch

        self._on_device(states, last_action)
        n = states.shape[0]
        out = torch.empty((n,), dtype=torch.int64, device=self._dev())
        self._check(self.L.dk_state_id(self.ctx, n, _ptr(states), _ptr(last_action), _ptr(out), stream if stream is not None else self._stream()), "dk_state_id")
        return out

    def apply(self, engine, states, actions, flags=0, err_out=None, stream=None):
        """In-place play_action.  actions: uint8 cuda tensor [n].  Returns err flags (uint8 [n])."""
        import torch

        n = states.shape[0]
        err_out = torch.empty((n,), dtype=torch.uint8, device=self._dev()) if err_out is None else err_out
        self._check(self.L.dk_apply(self.ctx, engine, n, _ptr(states), _ptr(actions), flags, _ptr(err_out), stream if stream is not None else self._stream()),
                    "dk_apply")
        return err_out

    def terminal(self, engine, states, stream=None):
        import torch

        n = states.shape[0]
        done = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        pts = torch.empty((n, 4), dtype=torch.int32, device=self._dev())
        self._check(self.L.dk_terminal(self.ctx, engine, n, _ptr(states), _ptr(done), _ptr(pts), stream if stream is not None else self._stream()), "dk_terminal")
        return done, pts

    def encode(self, layout, states, out=None, row_stride=None, stream=None):
        import torch

        n = states.shape[0]
        row_stride = OBS_LEN[layout] if row_stride is None else row_stride
        out = torch.empty((n, row_stride), dtype=torch.int64, device=self._dev()) if out is None else out
        self._check(self.L.dk_encode(self.ctx, layout, n, _ptr(states), _ptr(out), row_stride, stream if stream is not None else self._stream()), "dk_encode")
        return out

    def step_random_encode(self, states, rng, flags=DK_PLAYOUT_WITH_ANNOUNCEMENTS, obs_out=None, row_stride=311, action_out=None, want_obs=True, stream=None):
        import torch

        n = states.shape[0]
        if want_obs and obs_out is None:
            obs_out = torch.empty((n, row_stride), dtype=torch.int64, device=self._dev())
        if action_out is None:
            action_out = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_step_random_encode(self.ctx, n, _ptr(states), C.byref(rng), flags, _ptr(obs_out), row_stride, _ptr(action_out),
                                                 stream if stream is not None else self._stream()), "dk_step_random_encode")
        return obs_out, action_out

    def encode_narrow(self, layout, states, dtype=None, out=None, stream=None):
        """dk_encode_narrow: the rows of `encode` as int32 (default) or uint8 — dense [n, len], the values are identical."""
        import torch

        n = states.shape[0]
        dtype = (out.dtype if out is not None else torch.int32) if dtype is None else dtype
        if dtype not in (torch.int32, torch.uint8):
            raise ValueError("narrow observation rows are int32 or uint8")
        out = torch.empty((n, OBS_LEN[layout]), dtype=dtype, device=self._dev()) if out is None else out
        if out.dtype != dtype or tuple(out.shape) != (n, OBS_LEN[layout]):
            raise ValueError("out must be a dense [n, len] tensor of the requested dtype")
        self._check(self.L.dk_encode_narrow(self.ctx, layout, 4 if dtype == torch.int32 else 1, n, _ptr(states), _ptr(out),
                                            stream if stream is not None else self._stream()), "dk_encode_narrow")
        return out

    def step_random_encode_narrow(self, states, rng, flags=DK_PLAYOUT_WITH_ANNOUNCEMENTS, dtype=None, obs_out=None, action_out=None, stream=None):
        """dk_step_random_encode_narrow: one lock-step env step + the new states' 311-token rows as int32 (default) or uint8."""
        import torch

        n = states.shape[0]
        dtype = (obs_out.dtype if obs_out is not None else torch.int32) if dtype is None else dtype
        if dtype not in (torch.int32, torch.uint8):
            raise ValueError("narrow observation rows are int32 or uint8")
        obs_out = torch.empty((n, 311), dtype=dtype, device=self._dev()) if obs_out is None else obs_out
        if obs_out.dtype != dtype or tuple(obs_out.shape) != (n, 311):
            raise ValueError("obs_out must be a dense [n, 311] tensor of the requested dtype")
        if action_out is None:
            action_out = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_step_random_encode_narrow(self.ctx, n, _ptr(states), C.byref(rng), flags, 4 if dtype == torch.int32 else 1, _ptr(obs_out),
                                                        _ptr(action_out), stream if stream is not None else self._stream()), "dk_step_random_encode_narrow")
        return obs_out, action_out

    def playout_trace(self, engine, n, rng, stream=None):
        import torch

        pts = torch.empty((n, 4), dtype=torch.int32, device=self._dev())
        trace = torch.empty((n, 52), dtype=torch.uint8, device=self._dev())
        aux = torch.empty((n, 4), dtype=torch.int32, device=self._dev())
        self._check(self.L.dk_playout_trace(self.ctx, engine, n, C.byref(rng), _ptr(pts), _ptr(trace), _ptr(aux), stream if stream is not None else self._stream()),
                    "dk_playout_trace")
        return pts, trace, aux

    # ---- determinization / leaf rollouts ----------------------------------------------------------------------------------
    def determinize(self, engine, states, samples_per_info, rng, stream=None):
        """card_matching samples: returns (hands int64 [n,S,4], reservations uint8 [n,S,4], status uint8 [n,S])."""
        import torch

        n = states.shape[0]
        hands = torch.empty((n, samples_per_info, 4), dtype=torch.int64, device=self._dev())
        res = torch.empty((n, samples_per_info, 4), dtype=torch.uint8, device=self._dev())
        status = torch.empty((n, samples_per_info), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_determinize(self.ctx, engine, n, samples_per_info, _ptr(states), C.byref(rng), _ptr(hands), _ptr(res), _ptr(status),
                                          stream if stream is not None else self._stream()), "dk_determinize")
        return hands, res, status

    def leaf_rollouts(self, states, rollouts_per_leaf, rng, determinize=True, out=None, stream=None):
        """Exact integer sums of player_points over R rollouts per leaf: int64 [n,4]."""
        import torch

        n = states.shape[0]
        out = torch.empty((n, 4), dtype=torch.int64, device=self._dev()) if out is None else out
        self._check(self.L.dk_leaf_rollouts(self.ctx, n, rollouts_per_leaf, int(determinize), _ptr(states), C.byref(rng), _ptr(out),
                                            stream if stream is not None else self._stream()), "dk_leaf_rollouts")
        return out

    def encode_ipi(self, states, assumed_hands, assumed_reservations, next_player, out=None, row_stride=None, stream=None):
        """encode_state_ipi rows (int64 [n,row_stride]) + err flags (uint8 [n]).  assumed_hands int64 [n,4], assumed_reservations uint8 [n,4],
        next_player uint8 [n] — cuda tensors."""
        import torch

        n = states.shape[0]
        row_stride = row_stride or 311
        out = torch.empty((n, row_stride), dtype=torch.int64, device=self._dev()) if out is None else out
        err = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_encode_ipi(self.ctx, n, _ptr(states), _ptr(assumed_hands), _ptr(assumed_reservations), _ptr(next_player), _ptr(out),
                                         row_stride, _ptr(err), stream if stream is not None else self._stream()), "dk_encode_ipi")
        return out, err

    # ---- PIMC move decision ------------------------------------------------------------------------------
    def pimc_evaluate(self, states, n_det, n_rollouts, rng, want_values=True, stream=None):
        """Flat Monte-Carlo PIMC: (visits uint32→int32 view [n,n_det,39], value_sum int64 [n,n_det,39] or None, status uint8 [n,n_det])."""
        import torch

        n = states.shape[0]
        visits = torch.empty((n, n_det, N_ACTIONS), dtype=torch.int32, device=self._dev())
        values = torch.empty((n, n_det, N_ACTIONS), dtype=torch.int64, device=self._dev()) if want_values else None
        status = torch.empty((n, n_det), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_pimc_evaluate(self.ctx, n, n_det, n_rollouts, _ptr(states), C.byref(rng), _ptr(visits),
                                            _ptr(values) if want_values else None, _ptr(status),
                                            stream if stream is not None else self._stream()), "dk_pimc_evaluate")
        return visits, values, status

    def fuse(self, strategy, visits, allowed, status=None, stream=None):
        """PolicyFusionMaxN (FUSE_MAX_N) / PolicyFusionAverageStrategy (FUSE_AVERAGE) per root: (action uint8 [n], n_success int32 [n])."""
        import torch

        n, rows = visits.shape[0], visits.shape[1]
        action = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        n_ok = torch.empty((n,), dtype=torch.int32, device=self._dev())
        self._check(self.L.dk_fuse(self.ctx, strategy, n, rows, _ptr(visits), _ptr(status) if status is not None else None, _ptr(allowed),
                                   _ptr(action), _ptr(n_ok), stream if stream is not None else self._stream()), "dk_fuse")
        return action, n_ok

    def pimc_root_stats(self, visits, allowed, status=None, out=None, accumulate=False, stream=None):
        """Additive int64 root statistics [n, ROOT_STATS] of this rank's determinizations (MaxN rank sums | visit sums | successes)."""
        import torch

        n, rows = visits.shape[0], visits.shape[1]
        out = torch.zeros((n, ROOT_STATS), dtype=torch.int64, device=self._dev()) if out is None else out
        self._check(self.L.dk_pimc_root_stats(self.ctx, n, rows, _ptr(visits), _ptr(status) if status is not None else None, _ptr(allowed),
                                              _ptr(out), int(accumulate), stream if stream is not None else self._stream()), "dk_pimc_root_stats")
        return out

    def pimc_pick(self, strategy, stats, allowed, stream=None):
        import torch

        n = stats.shape[0]
        action = torch.empty((n,), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_pimc_pick(self.ctx, strategy, n, _ptr(stats), _ptr(allowed), _ptr(action),
                                        stream if stream is not None else self._stream()), "dk_pimc_pick")
        return action

    def uct_search(self, states, iterations, uct_c, rng, trees_per_root=1, determinize=False, workspace=None, stream=None):
        """One UCT tree per (state, d): (visits int32 [n,T,39], values float32 [n,T,39], action uint8 [n,T], status uint8 [n,T])."""
        import torch

        n = states.shape[0]
        need = self.L.dk_uct_workspace_bytes(n * trees_per_root, iterations)
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = torch.empty(((need + 15) // 16 * 2,), dtype=torch.int64, device=self._dev())
        visits = torch.empty((n, trees_per_root, N_ACTIONS), dtype=torch.int32, device=self._dev())
        values = torch.empty((n, trees_per_root, N_ACTIONS), dtype=torch.float32, device=self._dev())
        action = torch.empty((n, trees_per_root), dtype=torch.uint8, device=self._dev())
        status = torch.empty((n, trees_per_root), dtype=torch.uint8, device=self._dev())
        self._check(self.L.dk_uct_search(self.ctx, n, trees_per_root, int(determinize), iterations, uct_c, _ptr(states), C.byref(rng), _ptr(workspace),
                                         workspace.numel() * workspace.element_size(), _ptr(visits), _ptr(values), _ptr(action), _ptr(status),
                                         stream if stream is not None else self._stream()), "dk_uct_search")
        return visits, values, action, status

    def pack_replay_records(self, states, value, policy, out=None, stream=None):
        """bincode DBRecord bytes (uint8 [n, 2684]) of experience rows — the values the reference's replay buffer stores."""
        import torch

        n = states.shape[0]
        out = torch.empty((n, REPLAY_RECORD_BYTES), dtype=torch.uint8, device=self._dev()) if out is None else out
        self._check(self.L.dk_pack_replay_records(self.ctx, n, _ptr(states), _ptr(value), _ptr(policy), _ptr(out),
                                                  stream if stream is not None else self._stream()), "dk_pack_replay_records")
        return out

    def self_play(self, max_games, capacity):
        """Lock-step AlphaZero self-play driver with its experience buffers."""
        return SelfPlay(self, max_games, capacity)

    # ---- multi-GPU root statistics ----------------------------------------------------------------------------------------
    def comm_init(self, group=None):
        """Create the library's NCCL communicator over the ranks of a torch.distributed group (torch is only the rendezvous that
        carries the NCCL id); a single process without torch.distributed gets a one-rank communicator."""
        import torch.distributed as dist

        ident = (C.c_char * 128)()
        if not (dist.is_available() and dist.is_initialized()):
            self._check(self.L.dk_comm_unique_id(self.ctx, ident), "dk_comm_unique_id")
            self._check(self.L.dk_comm_init(self.ctx, 1, 0, ident), "dk_comm_init")
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if rank == 0:
            self._check(self.L.dk_comm_unique_id(self.ctx, ident), "dk_comm_unique_id")
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        self._check(self.L.dk_comm_init(self.ctx, world, rank, ident), "dk_comm_init")

    def comm_destroy(self):
        self._check(self.L.dk_comm_destroy(self.ctx), "dk_comm_destroy")

    def allreduce_root_stats(self, values, stream=None):
        """In-place sum over ranks of an int64 cuda tensor (ncclAllReduce on the library's communicator)."""
        self._check(self.L.dk_allreduce_root_stats(self.ctx, values.numel(), _ptr(values), stream if stream is not None else self._stream()),
                    "dk_allreduce_root_stats")
        return values


class DkSpBuffers(C.Structure):
    _fields_ = [("states", C.c_void_p), ("policy", C.c_void_p), ("value", C.c_void_p), ("player", C.c_void_p), ("game", C.c_void_p),
                ("capacity", C.c_size_t)]


SP_DONE, SP_FORCED, SP_KEPT, SP_DROPPED = 1, 2, 4, 8
SP_SEARCH_FORCED = 1


class SelfPlay:
    """Host mirror of `self_play` (rs-doko-alpha-zero/src/alpha_zero/train/self_play.rs:19-207) for a batch of games in lock-step.

    The experience buffers (states int64 [capacity,311], policy float32 [capacity,39], value float32 [capacity,4], player uint8,
    game int32) are torch cuda tensors owned by this object; rows [0, rows()) are valid."""

    def __init__(self, dk, max_games, capacity):
        import torch

        self.dk, self.max_games, self.capacity = dk, int(max_games), int(capacity)
        dev = dk._dev()
        self.states = torch.empty((capacity, OBS_LEN[DK_LAYOUT_FDO_PI311]), dtype=torch.int64, device=dev)
        self.policy = torch.empty((capacity, N_ACTIONS), dtype=torch.float32, device=dev)
        self.value = torch.zeros((capacity, 4), dtype=torch.float32, device=dev)
        self.player = torch.empty((capacity,), dtype=torch.uint8, device=dev)
        self.game = torch.empty((capacity,), dtype=torch.int32, device=dev)
        self.search_policy = torch.zeros((max_games, N_ACTIONS), dtype=torch.float32, device=dev)
        self.search_action = torch.zeros((max_games,), dtype=torch.uint8, device=dev)
        self.err = torch.zeros((max_games,), dtype=torch.uint8, device=dev)
        bufs = DkSpBuffers(_ptr(self.states), _ptr(self.policy), _ptr(self.value), _ptr(self.player), _ptr(self.game), capacity)
        self.h = C.c_void_p()
        dk._check(dk.L.dk_sp_create(dk.ctx, max_games, C.byref(bufs), C.byref(self.h)), "dk_sp_create")
        self.n = 0

    def close(self):
        if self.h:
            self.dk.L.dk_sp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.dk._check(self.dk.L.dk_sp_reset(self.h, self.dk._stream()), "dk_sp_reset")

    def begin_turn(self, states, az_epoch, keep_prob, rng, flags=0):
        self.n = states.shape[0]
        self.dk._check(self.dk.L.dk_sp_begin_turn(self.h, self.n, _ptr(states), az_epoch, keep_prob, flags, C.byref(rng), self.dk._stream()),
                       "dk_sp_begin_turn")

    def turn_view(self):
        """(allowed int64 [n], flags uint8 [n], rows int64 [n]) — views of the driver's device arrays, valid until the next begin_turn."""
        import torch

        a, f, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.dk._check(self.dk.L.dk_sp_turn_view(self.h, C.byref(a), C.byref(f), C.byref(r)), "dk_sp_turn_view")
        return (_wrap_device(a.value, (self.n,), torch.int64, self.dk._dev()), _wrap_device(f.value, (self.n,), torch.uint8, self.dk._dev()),
                _wrap_device(r.value, (self.n,), torch.int64, self.dk._dev()))

    def uniform_search(self, rng):
        self.dk._check(self.dk.L.dk_sp_uniform_search(self.h, C.byref(rng), _ptr(self.search_policy), _ptr(self.search_action), self.dk._stream()),
                       "dk_sp_uniform_search")
        return self.search_policy, self.search_action

    def end_turn(self, states, policy=None, action=None):
        policy = self.search_policy if policy is None else policy
        action = self.search_action if action is None else action
        self.dk._check(self.dk.L.dk_sp_end_turn(self.h, _ptr(states), _ptr(policy), _ptr(action), _ptr(self.err), self.dk._stream()), "dk_sp_end_turn")
        return self.err

    def finalize(self, states):
        self.dk._check(self.dk.L.dk_sp_finalize(self.h, _ptr(states), self.dk._stream()), "dk_sp_finalize")

    def counts(self):
        r, d, u = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.dk._check(self.dk.L.dk_sp_counts(self.h, C.byref(r), C.byref(d), C.byref(u), self.dk._stream()), "dk_sp_counts")
        return r.value, d.value, u.value


def _wrap_device(ptr, shape, dtype, device):
    """torch view of device memory owned by the library (no copy), via the CUDA array interface."""
    import torch

    typestr = {torch.int64: "<i8", torch.uint8: "|u1", torch.int32: "<i4", torch.float32: "<f4"}[dtype]

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}
    return torch.as_tensor(h, device=device)
