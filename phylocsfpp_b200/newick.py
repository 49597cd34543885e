"""Species tree: Newick parsing, --species reduction and flattening.

Host-side mirror of the reference's src/newick.hpp.  The flattened arrays are what crosses the C-ABI
(`pcsf_model_desc`, include/phylocsf_b200.h).

Reference semantics kept (file:line relative to the reference repo):
  * labels are lower-cased, branch lengths are parsed as doubles from [0-9.]+   (newick.hpp:31-91)
  * strictly binary trees; the root has branch length 0                          (newick.hpp:94-98)
  * flatten: leaves get ids 0..nl-1 in left-to-right DFS order, inner nodes get nl..n-1 in post-order,
    root = n-1; flattened branch lengths are stored as *float*                   (newick.hpp:100-229)
  * reduce(subset): single-child chains are merged, adding branch lengths in double, the root keeps
    length 0                                                                     (newick.hpp:286-363)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class Node:
    label: str = ""
    branch_length: float = 0.0  # double, as newick_node::branch_length
    left: Optional["Node"] = None
    right: Optional["Node"] = None
    parent: Optional["Node"] = None
    id: int = -999

    @property
    def is_leaf(self) -> bool:
        return self.left is None


class _Parser:
    def __init__(self, text: str):
        self.s = "".join(text.split())  # newick_open strips all whitespace (newick.hpp:155)
        self.i = 0

    def peek(self) -> str:
        return self.s[self.i] if self.i < len(self.s) else ""

    def number(self) -> float:
        j = self.i
        while j < len(self.s) and (self.s[j].isdigit() or self.s[j] == "."):
            j += 1
        tok = self.s[self.i:j]
        self.i = j
        return float(tok)  # std::stod

    def subtree(self, parent: Optional[Node]) -> Node:
        node = Node(parent=parent)
        if self.peek() == "(":
            self.i += 1
            node.left = self.subtree(node)
            assert self.peek() == ",", f"expected ',' at {self.i}"
            self.i += 1
            node.right = self.subtree(node)
            assert self.peek() == ")", f"expected ')' at {self.i} (only binary trees are supported)"
            self.i += 1
            if self.peek() == ":":
                self.i += 1
                node.branch_length = self.number()
        else:
            j = self.i
            while j < len(self.s) and self.s[j] not in "(),:":
                j += 1
            node.label = self.s[self.i:j].lower()
            self.i = j
            assert self.peek() == ":", f"leaf without branch length at {self.i}"
            self.i += 1
            node.branch_length = self.number()
        return node


def parse(text: str) -> Node:
    """newick_parse (newick.hpp:94): returns the root; asserts root.branch_length == 0."""
    p = _Parser(text)
    root = p.subtree(None)
    assert p.peek() in ("", ";")
    assert root.branch_length == 0.0
    return root


def leaves(node: Node) -> List[Node]:
    if node.is_leaf:
        return [node]
    return leaves(node.left) + leaves(node.right)


def _overlap(node: Node, subset) -> int:
    # newick_overlap_size (newick.hpp:264-270)
    if node.is_leaf:
        return 1 if node.label in subset else 0
    return _overlap(node.left, subset) + _overlap(node.right, subset)


def reduce(node: Node, subset) -> None:
    """newick_reduce (newick.hpp:286-363), in place."""
    while True:
        if node.is_leaf:
            return
        ol = _overlap(node.left, subset)
        orr = _overlap(node.right, subset)
        if ol == 0 or orr == 0:
            keep = node.right if ol == 0 else node.left
            node.left, node.right = keep.left, keep.right
            if node.left is not None:
                node.left.parent = node
                node.right.parent = node
            else:
                node.label = keep.label
            if node.parent is not None:  # the root carries no branch length
                node.branch_length += keep.branch_length
            continue  # newick_reduce(node, subset) again on the merged node
        reduce(node.left, subset)
        reduce(node.right, subset)
        return


@dataclass
class FlatTree:
    """std::vector<newick_elem> (newick.hpp:20-29) as arrays, plus the pointer tree's double lengths."""
    nl: int
    n: int
    child1: np.ndarray       # int16[n], -1 for leaves
    child2: np.ndarray       # int16[n]
    parent: np.ndarray       # int16[n], -1 for the root
    sibling: np.ndarray      # int16[n], -1 for the root
    branch_len: np.ndarray   # float32[n]  (newick_elem::branch_length, used by the likelihood)
    branch_len_f64: np.ndarray  # float64[n] (newick_node::branch_length, used by the BLS score)
    labels: List[str] = field(default_factory=list)  # "" for inner nodes


def flatten(root: Node) -> FlatTree:
    """newick_flatten (newick.hpp:218-229)."""
    lv = leaves(root)
    nl = len(lv)
    n = 2 * nl - 1
    for i, leaf in enumerate(lv):
        leaf.id = i
    counter = [nl]

    def annotate(node: Node):
        if not node.is_leaf:
            annotate(node.left)
            annotate(node.right)
            node.id = counter[0]
            counter[0] += 1

    annotate(root)
    assert counter[0] == n and root.id == n - 1
    t = FlatTree(nl, n, np.full(n, -1, np.int16), np.full(n, -1, np.int16), np.full(n, -1, np.int16),
                 np.full(n, -1, np.int16), np.zeros(n, np.float32), np.zeros(n, np.float64), [""] * n)

    def fill(node: Node):
        i = node.id
        t.branch_len[i] = np.float32(node.branch_length)
        t.branch_len_f64[i] = node.branch_length
        t.labels[i] = node.label
        if node.parent is not None:
            t.parent[i] = node.parent.id
            sib = node.parent.right if node.parent.left is node else node.parent.left
            t.sibling[i] = sib.id
        if not node.is_leaf:
            t.child1[i] = node.left.id
            t.child2[i] = node.right.id
            fill(node.left)
            fill(node.right)

    fill(root)
    return t


def strahler_order(t: FlatTree):
    """Evaluation order used by the CUDA prune kernel: children-before-parents DFS that always finishes
    the subtree needing more live partials first, so the number of simultaneously live inner partials is
    the tree's Strahler number (3-4 for the built-in trees).  Any children-first order gives the same
    values (fixed_lik.hpp:135-157 uses id order nl..n-1); this one minimises on-chip state.
    Returns (order: list of inner node ids, need: dict node -> live slots needed)."""
    need = {}

    def walk(i: int) -> int:
        if t.child1[i] < 0:
            need[i] = 0
            return 0
        a, b = walk(int(t.child1[i])), walk(int(t.child2[i]))
        need[i] = max(1, a, b) if a != b else max(1, a + 1) if a > 0 else 1
        return need[i]

    walk(t.n - 1)
    order: List[int] = []

    def emit(i: int):
        if t.child1[i] < 0:
            return
        c1, c2 = int(t.child1[i]), int(t.child2[i])
        first, second = (c1, c2) if need[c1] >= need[c2] else (c2, c1)
        emit(first)
        emit(second)
        order.append(i)

    emit(t.n - 1)
    return order, need
