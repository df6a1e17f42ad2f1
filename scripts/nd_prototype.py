"""Prototype (sizing study) of the nested-dissection hierarchy behind the sparse exact preconditioner.
Not product code: the product implementation is dpo_b200/csrc/nd_precond.{h,cpp}.  Usage:
    python scripts/nd_prototype.py sphere2500 [agents]
"""
import sys, os, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpo_b200 import posegraph as pg


def adjacency(n, p1, p2):
    adj = [set() for _ in range(n)]
    for a, b in zip(p1, p2):
        if a != b:
            adj[a].add(b); adj[b].add(a)
    return [np.array(sorted(s), dtype=np.int64) for s in adj]


def bfs_levels(adj, nodes_mask, start):
    level = {start: 0}
    order = [start]
    q = collections.deque([start])
    while q:
        u = q.popleft()
        for v in adj[u]:
            if nodes_mask[v] and v not in level:
                level[v] = level[u] + 1
                order.append(v); q.append(v)
    return level, order


def bisect(adj, nodes, mask):
    """vertex separator of the sub-graph induced by `nodes` (mask = membership). returns (A, B, S)"""
    nodes = list(nodes)
    # connected components first
    seen = set(); comps = []
    for s in nodes:
        if s in seen: continue
        lv, order = bfs_levels(adj, mask, s)
        seen.update(order); comps.append(order)
    if len(comps) > 1:
        comps.sort(key=len, reverse=True)
        A, B = [], []
        for c in comps:
            (A if len(A) <= len(B) else B).extend(c)
        return A, B, []
    # pseudo-peripheral start
    s = nodes[0]
    for _ in range(4):
        lv, order = bfs_levels(adj, mask, s)
        far = order[-1]
        if far == s: break
        s = far
    lv, order = bfs_levels(adj, mask, s)
    nl = max(lv.values()) + 1
    cnt = np.zeros(nl, dtype=np.int64)
    for v in order: cnt[lv[v]] += 1
    cum = np.cumsum(cnt)
    n = len(nodes)
    best, bk = None, None
    for k in range(1, nl - 1):
        a, b = cum[k - 1], n - cum[k]
        if min(a, b) < 0.25 * n: continue
        score = cnt[k] + 0.02 * abs(a - b)
        if best is None or score < best: best, bk = score, k
    if bk is None:
        bk = int(np.searchsorted(cum, n // 2))
        bk = min(max(bk, 1), nl - 2) if nl >= 3 else None
    if bk is None:
        return nodes[: n // 2], [], nodes[n // 2:]   # degenerate (clique-ish): everything else is "separator"
    A = [v for v in order if lv[v] < bk]
    B = [v for v in order if lv[v] > bk]
    S = [v for v in order if lv[v] == bk]
    # thin the separator: a separator vertex with no neighbour in B moves to A (and vice versa)
    inB = set(B); inA = set(A)
    S2 = []
    for v in S:
        nb = any((u in inB) for u in adj[v] if mask[u])
        na = any((u in inA) for u in adj[v] if mask[u])
        if nb and na: S2.append(v)
        elif na or not nb: A.append(v); inA.add(v)
        else: B.append(v); inB.add(v)
    return A, B, S2


def nested_dissection(adj, n, leaf_size):
    """returns list of tree nodes: dict(parent, level, sep (list), leaf(bool))"""
    tree = []
    mask = np.zeros(n, dtype=bool)

    def rec(nodes, parent, level):
        idx = len(tree)
        tree.append(dict(parent=parent, level=level, own=None, children=[]))
        if parent >= 0: tree[parent]["children"].append(idx)
        if len(nodes) <= leaf_size:
            tree[idx]["own"] = list(nodes); return
        mask[:] = False; mask[nodes] = True
        A, B, S = bisect(adj, nodes, mask)
        if len(A) == 0 or len(B) == 0:
            tree[idx]["own"] = list(nodes); return
        tree[idx]["own"] = list(S)
        rec(A, idx, level + 1); rec(B, idx, level + 1)

    rec(list(range(n)), -1, 0)
    return tree


def macro_stats(tree, adj, cuts, dh=4, r=5):
    """cuts: sorted list of ND levels where a new macro node starts (always contains 0)."""
    depth = max(t["level"] for t in tree) + 1
    cuts = sorted(set(cuts))
    # macro id of a tree node = the ancestor at the largest cut level <= its level
    macro_of = [None] * len(tree)
    for i, t in enumerate(tree):
        a = i
        lvl = t["level"]
        target = max(c for c in cuts if c <= lvl)
        while tree[a]["level"] > target: a = tree[a]["parent"]
        macro_of[i] = a
    macros = sorted(set(macro_of))
    own = {m: [] for m in macros}
    for i, t in enumerate(tree): own[macro_of[i]].extend(t["own"])
    mparent = {}
    for m in macros:
        p = tree[m]["parent"]
        mparent[m] = macro_of[p] if p >= 0 else -1
    node_macro = {}
    for m in macros:
        for v in own[m]: node_macro[v] = m
    # ancestors sets
    def ancestors(m):
        out = []
        while mparent[m] >= 0:
            m = mparent[m]; out.append(m)
        return out
    # boundary by symbolic elimination in post-order (children before parents): process macros by decreasing level
    order = sorted(macros, key=lambda m: -tree[m]["level"])
    bnd = {}
    kids = {m: [] for m in macros}
    for m in macros:
        if mparent[m] >= 0: kids[mparent[m]].append(m)
    for m in order:
        s = set()
        mine = set(own[m])
        for v in own[m]:
            for u in adj[v]:
                if u not in mine and node_macro[u] != m: s.add(u)
        for c in kids[m]: s |= bnd[c]
        s -= mine
        # only ancestors may remain
        anc = set(ancestors(m))
        s = {u for u in s if node_macro[u] in anc}
        bnd[m] = s
    rows = []
    total = 0
    for m in macros:
        s, b = len(own[m]) * dh, len(bnd[m]) * dh
        byt = (s * s + 2 * s * b) * 8
        total += byt
        rows.append((tree[m]["level"], s, b, byt))
    return rows, total


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "sphere2500"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if name.startswith("grid"):
        dims = tuple(int(x) for x in name[4:].split("x"))
        edges, n, _ = pg.synthetic_grid_graph(*dims, edges_per_pose=4.0, seed=0)
    else:
        edges, n = pg.read_g2o_file(os.path.join(root, "data", name + ".g2o"))
    p1, p2 = np.asarray(edges.p1), np.asarray(edges.p2)
    if k > 1:
        per = n // k
        own = np.minimum(np.arange(n) // per, k - 1)
        sel = (own[p1] == 0) & (own[p2] == 0)
        p1, p2 = p1[sel], p2[sel]
        n = int((own == 0).sum())
    adj = adjacency(n, p1, p2)
    tree = nested_dissection(adj, n, leaf_size=int(os.environ.get("LEAF", "12")))
    depth = max(t["level"] for t in tree) + 1
    per_level = collections.Counter()
    cnt_level = collections.Counter()
    for t in tree:
        per_level[t["level"]] += len(t["own"]); cnt_level[t["level"]] += 1
    print(f"{name}: n={n} tree nodes={len(tree)} depth={depth}")
    for l in range(depth):
        print(f"  level {l}: nodes {cnt_level[l]:4d} own poses {per_level[l]:6d}")
    import itertools
    best = []
    levels = list(range(1, depth))
    for ncut in (1, 2, 3, 4):
        for cs in itertools.combinations(levels, ncut):
            rows, total = macro_stats(tree, adj, [0] + list(cs))
            smax = max(r[1] for r in rows)
            # time model: per macro level fwd+bwd barrier 1.0us each + bytes / 5 TB/s, top level once
            nlev = ncut + 1
            t = (2 * nlev - 1) * 1.0 + total / 5e6
            best.append((t, cs, total, smax, len(rows)))
    best.sort()
    for t, cs, total, smax, nm in best[:8]:
        print(f"  cuts {cs}: est {t:.1f} us/apply, {total/1e6:.1f} MB, max s {smax}, macro nodes {nm}")
    t, cs, total, smax, nm = best[0]
    rows, total = macro_stats(tree, adj, [0] + list(cs))
    agg = collections.defaultdict(list)
    for lvl, s, b, byt in rows: agg[lvl].append((s, b, byt))
    for lvl in sorted(agg):
        a = agg[lvl]
        print(f"    macro level@{lvl}: {len(a)} nodes, s avg {np.mean([x[0] for x in a]):.0f} max {max(x[0] for x in a)}, "
              f"b avg {np.mean([x[1] for x in a]):.0f} max {max(x[1] for x in a)}, MB {sum(x[2] for x in a)/1e6:.2f}")


if __name__ == "__main__":
    main()
