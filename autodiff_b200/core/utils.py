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
