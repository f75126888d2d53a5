"""autodiff_b200 — B200-native execution backend behind bgavran/autodiff's graph API.

Drop-in for the reference's `import autodiff as ad` on its data-parallel hot path: the same `Node`, `Variable`,
ops, `grad`, `checkpoint`, `Module`, `NN` and optimizers (reference autodiff/__init__.py:1-7), with node
evaluation running as hand-written sm_100a kernels (libadb200.so) on device-resident buffers.  There is no CPU
fallback: evaluating a node without a B200 raises.
"""
from .core.node import Node, Variable, ConstFill, add_context
from .core.ops import *            # noqa: F401,F403
from .core.reshape import *        # noqa: F401,F403
from .core.high_level_ops import * # noqa: F401,F403
from .core.wrappers import checkpoint
from .core.conv import Col2Im, Conv2D, Im2Col
from .core.grad import grad
from .device import DeviceArray, ewise_stats, force_interpreter, init, launch_count, pinned_empty, precision, set_precision, synchronize, to_device, to_host, use_stream
from .core.engine import eval_many, evaluate_many
