"""CPU model of the BVH builder's tree stage (j3d_b200/csrc/build.cu, tree_fit_kernel / upper_tree_kernel / climb_kernel).

The CUDA builder no longer runs Karras' searches on the keys: inside a block of 256 sorted triangles a node's range and
split come from a table of the deltas of ADJACENT keys by binary lifting, its box from a sparse table of the leaf boxes, and
ranges are handed down by the collapse instead of being stored.  The GPU tests pin the result (identical tree, identical
pixels); this file pins the ARGUMENT on small inputs, in plain Python, so that it runs without a GPU:

  * lifting over adjacent deltas == Karras' doubling + binary searches (range, split, children), ties broken by position,
  * the window rule: a range of fewer than 256 pairs is exact inside a window of 256 pairs on either side, anything longer
    is deferred,
  * range min / max over leaf boxes == the bottom-up fit,
  * ranges derived top-down from (parent range, left child number) == the ranges of the nodes,
  * at most two top nodes per tree level per block (the bound the climber list is sized for).

Nothing here is imported by the product.
"""
import numpy as np
import pytest

BLOCK = 256


def clz64(x):
    return 64 - int(x).bit_length()


def clz32(x):
    return 32 - int(x).bit_length()


def delta(keys, i, j):
    """build.cu delta(): common prefix of the masked keys, equal keys told apart by their position."""
    n = len(keys)
    if j < 0 or j >= n:
        return -1
    a, b = int(keys[i]), int(keys[j])
    if a != b:
        return clz64(a ^ b)
    return 64 + clz32(i ^ j)


def karras(keys, i):
    """The search of radix_tree_kernel (Karras 2012) as the builder ran it until round 2: (lo, hi, gamma)."""
    d = 1 if delta(keys, i, i + 1) - delta(keys, i, i - 1) >= 0 else -1
    dmin = delta(keys, i, i - d)
    lmax = 2
    while delta(keys, i, i + lmax * d) > dmin:
        lmax <<= 1
    l = 0
    s = lmax >> 1
    while s >= 1:
        if delta(keys, i, i + (l + s) * d) > dmin:
            l += s
        s >>= 1
    j = i + l * d
    dnode = delta(keys, i, j)
    sp, sh, t = 0, 1, (l + 1) >> 1
    while True:
        if delta(keys, i, i + (sp + t) * d) > dnode:
            sp += t
        if t == 1:
            break
        sh += 1
        t = (l + (1 << sh) - 1) >> sh
    gamma = i + sp * d + min(d, 0)
    return min(i, j), max(i, j), gamma


def adjacent_table(keys, q0, npairs):
    """s_dt of tree_fit_kernel: level 0 = 1 + delta(q, q + 1) for q = q0 .. q0 + npairs - 1 (0 past the array), level v = min over 2^v."""
    n = len(keys)
    lv0 = [0 if (q < 0 or q + 1 >= n) else 1 + delta(keys, q, q + 1) for q in range(q0, q0 + npairs)]
    tab = [lv0]
    for v in range(1, 9):
        prev = tab[-1]
        tab.append([min(prev[p], prev[min(p + (1 << (v - 1)), npairs - 1)]) for p in range(npairs)])
    return tab


def lifted(keys, k, s):
    """Node k of the block starting at s from the delta table, exactly as tree_fit_kernel does it.  None = deferred."""
    npairs = 3 * BLOCK
    tab = adjacent_table(keys, s - BLOCK, npairs)
    tid = k - s
    pf, pb = BLOCK + tid, BLOCK + tid - 1
    fwd, bwd = tab[0][pf], tab[0][pb]
    assert fwd != bwd
    d = 1 if fwd > bwd else -1
    dmin = min(fwd, bwd)
    if d > 0:
        pos = pf
        for v in range(8, -1, -1):
            if pos + (1 << v) <= npairs and tab[v][pos] > dmin:
                pos += 1 << v
        l = pos - pf
    else:
        pos = pb
        for v in range(8, -1, -1):
            if pos - (1 << v) + 1 >= 0 and tab[v][pos - (1 << v) + 1] > dmin:
                pos -= 1 << v
        l = pb - pos
    if l >= BLOCK:
        return None
    p0 = pf if d > 0 else pb - l + 1
    p1 = p0 + l - 1
    lv = l.bit_length() - 1
    dnode = min(tab[lv][p0], tab[lv][p1 - (1 << lv) + 1])
    pos = p0
    for v in range(7, -1, -1):
        if pos + (1 << v) - 1 <= p1 and tab[v][pos] > dnode:
            pos += 1 << v
    gamma = s - BLOCK + pos
    lo, hi = (k, k + l) if d > 0 else (k - l, k)
    return lo, hi, gamma


def key_sets():
    rng = np.random.default_rng(11)
    out = {}
    out["random_39bit"] = np.sort(rng.integers(0, 1 << 39, 700, dtype=np.uint64) << np.uint64(25))
    out["few_values"] = np.sort(rng.integers(0, 9, 900, dtype=np.uint64) << np.uint64(30))        # long runs of equal keys
    out["all_equal"] = np.full(600, 5 << 40, dtype=np.uint64)                                     # the tree is built on positions only
    out["clustered"] = np.sort(np.concatenate([rng.integers(0, 1 << 12, 500, dtype=np.uint64), (np.uint64(1) << np.uint64(47)) + rng.integers(0, 4, 90, dtype=np.uint64)]) << np.uint64(8))
    out["ramp"] = (np.arange(1030, dtype=np.uint64) * np.uint64(3)) << np.uint64(20)              # a strip: deep, one-sided subtrees
    out["tiny"] = np.array([1 << 30, 1 << 30, 7 << 30], dtype=np.uint64)
    return out


@pytest.mark.parametrize("name", sorted(key_sets().keys()))
def test_lifting_equals_karras(name):
    keys = key_sets()[name]
    n = len(keys)
    deferred = 0
    for k in range(n - 1):
        lo, hi, gamma = karras(keys, k)
        s = (k // BLOCK) * BLOCK
        got = lifted(keys, k, s)
        if hi - lo >= BLOCK:            # 256 pairs or more: the window cannot tell, the node goes to upper_tree_kernel
            assert got is None
            deferred += 1
        else:
            assert got == (lo, hi, gamma), (name, k)
    assert deferred <= max(1, 2 * n // BLOCK + 2)


def build_tree(keys):
    n = len(keys)
    first_leaf = n - 1
    nodes = []
    for k in range(n - 1):
        lo, hi, gamma = karras(keys, k)
        left = first_leaf + gamma if lo == gamma else gamma
        right = first_leaf + gamma + 1 if hi == gamma + 1 else gamma + 1
        nodes.append((lo, hi, left, right))
    return nodes


@pytest.mark.parametrize("name", ["random_39bit", "few_values", "all_equal", "ramp"])
def test_boxes_ranges_and_top_nodes(name):
    keys = key_sets()[name]
    n = len(keys)
    first_leaf = n - 1
    rng = np.random.default_rng(3)
    lmn = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    lmx = lmn + rng.uniform(0, 0.1, (n, 3)).astype(np.float32)
    nodes = build_tree(keys)
    # bottom-up fit (what refit did child by child) against range min / max (what the sparse table answers)
    box = {}

    def fit(node):
        if node >= first_leaf:
            return lmn[node - first_leaf], lmx[node - first_leaf]
        if node not in box:
            a, b = fit(nodes[node][2]), fit(nodes[node][3])
            box[node] = (np.minimum(a[0], b[0]), np.maximum(a[1], b[1]))
        return box[node]

    import sys
    sys.setrecursionlimit(10000)
    for k, (lo, hi, _, _) in enumerate(nodes):
        mn, mx = fit(k)
        assert (mn == lmn[lo:hi + 1].min(0)).all() and (mx == lmx[lo:hi + 1].max(0)).all()
    # ranges handed down by the collapse: children of a node that owns [lo, hi] own [lo, gamma] and [gamma + 1, hi],
    # gamma read off the left child's number
    stack = [(0, 0, n - 1)]
    seen = 0
    while stack:
        node, lo, hi = stack.pop()
        if node >= first_leaf:
            assert lo == hi == node - first_leaf
            continue
        assert (lo, hi) == nodes[node][:2]
        seen += 1
        left, right = nodes[node][2], nodes[node][3]
        gamma = left - first_leaf if left >= first_leaf else left
        stack.append((left, lo, gamma))
        stack.append((right, gamma + 1, hi))
    assert seen == n - 1
    # top nodes of a block (in-block node or leaf whose parent straddles the block): at most two per level of the tree
    parent = {}
    depth = {0: 0}
    order = [0]
    for node in order:
        for c in nodes[node][2:]:
            parent[c] = node
            depth[c] = depth[node] + 1
            if c < first_leaf:
                order.append(c)
    for s in range(0, n, BLOCK):
        e = min(s + BLOCK, n)
        inside = lambda x: (s <= x - first_leaf < e) if x >= first_leaf else (nodes[x][0] >= s and nodes[x][1] < e)
        tops = [x for x in list(range(n - 1)) + [first_leaf + i for i in range(n)] if inside(x) and (x == 0 or not inside(parent[x]))]
        per_level = {}
        for x in tops:
            per_level[depth[x]] = per_level.get(depth[x], 0) + 1
        assert max(per_level.values()) <= 2
        assert len(tops) <= 160   # CLIMB_SLOTS
