"""hierarchicalkarting_b200 — B200-native (sm_100a CUDA) planning kernels for HierarchicalKarting's hot path:
the batched feedback LQ Nash game solver (reference: Assets/Karting/Scripts/AI/LQR/) and leaf-parallel rollouts of the
discrete race game (reference: Assets/Karting/Scripts/AI/MCTS/), behind the C-ABI of include/hk_abi.h."""
from . import abi  # noqa: F401

__all__ = ["abi", "lqr", "mcts", "tracks", "scenarios"]
