"""master_doko_reinforcement_learning_b200 — B200-native batched Doppelkopf simulator.

Host-side mirror (Python) of the reference's game-state seam for ONE hot path: it binds the
C ABI of ``libdoko_cuda.so`` (include/doko_cuda.h) with ctypes and moves device memory with torch.  There is no
CPU fallback: importing works anywhere, but every compute call needs the CUDA library and an sm_100 GPU and
raises ``DokoCudaError`` otherwise.
"""
from .api import (DokoCuda, DokoCudaError, DK_DOKO, DK_FDO, DK_PLAYOUT_WITH_ANNOUNCEMENTS, DK_APPLY_SKIP_SINGLE,  # noqa: F401
                  DK_LAYOUT_DO110, DK_LAYOUT_DO114, DK_LAYOUT_FDO_PI311, DK_STATE_DTYPE, N_ACTIONS, ACTION_NONE, FUSE_MAX_N,
                  FUSE_AVERAGE, ROOT_STATS, library_path, load_library)

__all__ = ["DokoCuda", "DokoCudaError", "DK_DOKO", "DK_FDO", "DK_PLAYOUT_WITH_ANNOUNCEMENTS", "DK_APPLY_SKIP_SINGLE",
           "DK_LAYOUT_DO110", "DK_LAYOUT_DO114", "DK_LAYOUT_FDO_PI311", "DK_STATE_DTYPE", "N_ACTIONS", "ACTION_NONE", "FUSE_MAX_N", "FUSE_AVERAGE", "ROOT_STATS",
           "library_path", "load_library"]
