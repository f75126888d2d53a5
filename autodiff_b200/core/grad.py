import collections

from .node import ConstFill, Variable
from .ops import Add  # noqa: F401  (grad accumulates with `+=`, i.e. Add nodes, as the reference does)
from .utils import gc_paused, reverse_topo_sort


def grad(top_node, wrt_list, previous_grad=None):
    """Graph -> graph reverse-mode transform, semantics of reference autodiff/core/grad.py:9-38.

    Returns, for every node in wrt_list, a NEW graph whose evaluation is d(top_node)/d(node); because partial
    derivatives are themselves graphs, grad(grad(...)) works.  No arithmetic happens here (except the eager
    evaluation inside Slice's partial derivative, a reference quirk kept on purpose).

    :param previous_grad: incoming gradient of top_node; default ones(top_node.shape) (never materialised)
    """
    assert isinstance(wrt_list, list) or isinstance(wrt_list, tuple)
    if previous_grad is None:
        previous_grad = ConstFill(tuple(top_node.shape), 1.0, name=add_sum_name(top_node))

    with gc_paused():
        dct = collections.defaultdict(lambda: Variable(0))
        dct[top_node] += previous_grad

        for node in reverse_topo_sort(top_node):
            for child in set(node.children):
                dct[child] += node.partial_derivative(wrt=child, previous_grad=dct[node])

        return [dct[wrt] for wrt in wrt_list]


def add_sum_name(node):
    return "'" + node.name + "' grad_sum"
