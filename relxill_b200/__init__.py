"""relxill_b200 — B200-native (sm_100a) implementation of relxill's spectrum-evaluation hot path.

Python here is a thin host-side mirror of the reference's local-model interface; the product is
librelxill_b200.so (hand-written CUDA behind a C ABI, include/relxill_b200.h).
"""
from .api import (Batch, LocalModel, ModelEvalFailed, ModelNotFound, PARAM_NAMES, batch_eval, default_energy_grid,
                  default_params, init, init_devices, last_eval_reuse, lmod, num_devices, num_params, get_xill_grid,
                  set_cache, set_num_zones, set_sharding, set_xill_grid, shutdown)

__all__ = ["Batch", "LocalModel", "ModelEvalFailed", "ModelNotFound", "PARAM_NAMES", "batch_eval",
           "default_energy_grid", "default_params", "init", "init_devices", "last_eval_reuse", "lmod", "num_devices",
           "num_params", "get_xill_grid", "set_cache", "set_num_zones", "set_sharding", "set_xill_grid", "shutdown"]
