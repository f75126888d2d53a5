from .grad import grad
from .node import Node


def checkpoint(fn):
    """Gradient rematerialisation wrapper, reference autodiff/core/wrappers.py:5-15: the wrapped sub-graph is an
    opaque node whose forward builds + evaluates a throw-away inner graph (on the device) and whose gradient
    rebuilds it.  Superset behaviour: `shape` is inferred from fn's graph so default-seed grad() works."""
    def wrap_in_primitive(*fn_args):
        op = Node(children=fn_args, name=fn.__name__)
        op._eval = lambda: fn(*fn_args)()
        op._partial_derivative = lambda wrt, previous_grad: grad(fn(*fn_args), [wrt], previous_grad=previous_grad)[0]
        try:
            op.shape = fn(*fn_args).shape
        except Exception:
            op.shape = None
        return op

    return wrap_in_primitive
