def tree_map(f, tree):
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    return f(tree)

def tree_reduce(f, tree, init):
    acc = init
    if isinstance(tree, dict):
        for v in tree.values():
            acc = tree_reduce(f, v, acc)
        return acc
    return f(acc, tree)
