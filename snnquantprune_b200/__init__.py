"""snnquantprune_b200 -- B200-native (sm_100a) quantized + pruned spiking-layer
forward pass, a drop-in for the hot path of SNNQuantPrune's TCJA-SNN.

Importing works on a CPU-only box (the driver's build check); any compute call
needs libsnnqp.so and an sm_100 GPU and fails loudly otherwise."""
from . import _lib  # noqa: F401
from .quant import DuQ, prune, gaussian_init, max_init, QuantConfig  # noqa: F401
from .flax_qconv import QuantConv  # noqa: F401
from .flax_qdense import QuantDense  # noqa: F401
from .spiking_learning import SpikingBlock, multi_step_LIF, atan, BatchNorm  # noqa: F401
from .models import CextNet, ModelConfig, eval_step  # noqa: F401
from .pack import pack_cextnet  # noqa: F401
from .engine import CextNetEngine  # noqa: F401
from .eval import evaluate  # noqa: F401

__all__ = ["DuQ", "prune", "gaussian_init", "max_init", "QuantConfig", "QuantConv", "QuantDense",
           "SpikingBlock", "multi_step_LIF", "atan", "BatchNorm", "CextNet", "ModelConfig",
           "eval_step", "pack_cextnet", "CextNetEngine", "evaluate"]
