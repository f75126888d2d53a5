def reverse_topo_sort(top_node):
    """Nodes of the top_node subtree in reverse topological order — same order as the reference's recursive
    DFS (autodiff/core/utils.py:1-17) but iterative and O(n) (the reference copies the list at every node)."""
    order = []
    visited = {top_node}
    stack = [(top_node, iter(top_node.children))]
    while stack:
        node, it = stack[-1]
        for child in it:
            if child not in visited:
                visited.add(child)
                stack.append((child, iter(child.children)))
                break
        else:
            order.append(node)
            stack.pop()
    return reversed(order)


class gc_paused:
    """Pauses the cyclic garbage collector for a graph-sized burst of allocations.  Graphs are DAGs held by child
    references only (no cycles), so reference counting frees them; the generational collector would just re-scan the
    tens of thousands of live nodes of a second-order graph over and over (measured: ~15 % of graph-build time)."""
    __slots__ = ("was",)

    def __enter__(self):
        import gc
        self.was = gc.isenabled()
        if self.was:
            gc.disable()

    def __exit__(self, *exc):
        if self.was:
            import gc
            gc.enable()
        return False
