"""`checkpoint(fn)` - gradient rematerialisation (API of reference autodiff/core/wrappers.py:5-15)."""
from .grad import grad
from .node import Node


class _Checkpointed(Node):
    """The sub-graph `fn(*args)` as ONE opaque node of the outer graph: nothing of it is kept between the forward value and
    the backward pass - both rebuild the inner graph from `fn` (the reference does the same with two lambdas on a bare
    Node).  Beyond the reference: the value stays on the device, and `shape` is inferred from the inner graph so that a
    default-seed `grad()` through the node works."""

    def __init__(self, fn, args):
        super().__init__(list(args), name=fn.__name__)
        self._fn = fn
        try:
            self.shape = self._inner().shape
        except Exception:
            self.shape = None

    def _inner(self):
        return self._fn(*self.children)

    def _eval_device(self):
        return self._inner().eval_device()

    def _partial_derivative(self, wrt, previous_grad):
        return grad(self._inner(), [wrt], previous_grad=previous_grad)[0]


def checkpoint(fn):
    def build(*fn_args):
        return _Checkpointed(fn, fn_args)

    build.__name__ = getattr(fn, "__name__", "checkpoint")
    return build
