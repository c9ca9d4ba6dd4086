"""tante_b200 -- B200-native (sm_100a) implementation of TANTE's hot path.

Host-side mirror of the reference interface (`models.TANTE`, the trainers' `rollout_model`)
over a C-ABI CUDA library (include/tante_b200.h).  No CPU fallback.
"""
from .tante import TANTE, TanteMetadata  # noqa: F401
from .rollout import R_Evaler, Evaler, rollout_eval  # noqa: F401
from .trainer import R_Trainer, Trainer  # noqa: F401
from .metrics import MSE, NMSE, L2RE, NNMSE, RMSE, NRMSE, VMSE, VRMSE  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__all__ = ["TANTE", "TanteMetadata", "R_Evaler", "Evaler", "R_Trainer", "Trainer", "rollout_eval", "FusedAdamW",
           "MSE", "NMSE", "L2RE", "NNMSE", "RMSE", "NRMSE", "VMSE", "VRMSE"]
