"""Newick helpers shared by the tests and tests/golden/make_goldens.py (test infrastructure).

canonical_newick() maps any Newick string of an unrooted binary tree with integer leaf labels to a
unique representative: the tree is re-rooted at the leaf with the smallest label, and at every
internal node the subtrees are ordered by their smallest leaf label.  Two trees are the same
unrooted labelled topology iff their canonical strings are equal (BASELINE.json: "the same
canonicalised tree set (Newick, sorted)").
"""


def parse_newick(s):
    """Return a nested structure: int for a leaf, list of children for an internal node."""
    s = s.strip().rstrip(";")
    pos = 0

    def node():
        nonlocal pos
        if s[pos] == "(":
            pos += 1
            kids = [node()]
            while s[pos] == ",":
                pos += 1
                kids.append(node())
            assert s[pos] == ")", (s, pos)
            pos += 1
            return kids
        start = pos
        while pos < len(s) and s[pos] not in ",()":
            pos += 1
        return int(s[start:pos])

    t = node()
    assert pos == len(s), (s, pos)
    return t


def _adjacency(t):
    """Unrooted adjacency lists; leaves are ('L', label), internal nodes ('I', k)."""
    adj = {}
    counter = [0]

    def link(a, b):
        adj.setdefault(a, []).append(b)
        adj.setdefault(b, []).append(a)

    def build(n):
        if isinstance(n, int):
            return ("L", n)
        me = ("I", counter[0])
        counter[0] += 1
        for k in n:
            link(me, build(k))
        return me

    top = build(t)
    # Suppress a degree-2 top node (a rooted binary Newick) so that the tree is properly unrooted.
    if top[0] == "I" and len(adj[top]) == 2:
        a, b = adj.pop(top)
        adj[a].remove(top)
        adj[b].remove(top)
        link(a, b)
    return adj


def canonical_newick(s):
    adj = _adjacency(parse_newick(s))
    leaves = [n for n in adj if n[0] == "L"]
    root = min(leaves, key=lambda n: n[1])

    def render(n, parent):
        """Returns (min label, string) for the subtree hanging off n away from parent."""
        if n[0] == "L":
            return n[1], str(n[1])
        subs = sorted(render(k, n) for k in adj[n] if k != parent)
        return subs[0][0], "(" + ",".join(x[1] for x in subs) + ")"

    (nb,) = adj[root]
    if nb[0] == "L":  # two-leaf tree
        return "(%d,%d);" % (root[1], nb[1])
    subs = sorted(render(k, nb) for k in adj[nb] if k != root)
    return "(" + str(root[1]) + "," + ",".join(x[1] for x in subs) + ");"


def canonical_set(newicks):
    return sorted(canonical_newick(t) for t in newicks)
