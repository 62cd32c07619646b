"""Plain-Python decoding tree used by the oracle and the tests.  TEST INFRASTRUCTURE ONLY.

It carries exactly the attributes the reference metadata builder reads from ``TreeCache`` /
``TreeNode`` (``/root/reference/DeFT/deft/tree_decoding/tree_cache.py:94-129, 147-190``) and can be
frozen to / thawed from flat integer arrays so that trees built by the *reference* ``TreeCache``
(in ``oracle/gen_golden.py``) travel inside ``tests/golden/*.npz``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


class PlainNode:
    def __init__(self, id: int) -> None:
        self.id = id
        self.parent: Optional["PlainNode"] = None
        self.children: Dict[int, "PlainNode"] = {}
        self.kv_indices: List[int] = []
        self.refs: set = set()
        self.paused = False
        self.node_indices_id: Optional[int] = None

    def __hash__(self) -> int:
        return hash(self.id)

    def __eq__(self, other) -> bool:
        return self is other


class PlainTree:
    def __init__(self) -> None:
        self.root: Optional[PlainNode] = None
        self.nodes: Dict[int, PlainNode] = {}
        self.leaves: Dict[int, PlainNode] = {}


def freeze(tree) -> Dict[str, np.ndarray]:
    """Any duck-typed tree (reference TreeCache, deft_b200 TreeCache, PlainTree) -> flat arrays.

    Nodes are emitted in DFS pre-order with children in dict order, so that ``thaw`` re-creates
    the same child insertion order.
    """
    ids, parents, kv_off, kv, refs_off, refs, tix = [], [], [0], [], [0], [], []

    def visit(n, parent_id):
        ids.append(n.id); parents.append(parent_id)
        kv.extend(int(x) for x in n.kv_indices); kv_off.append(len(kv))
        refs.extend(sorted(r.id for r in n.refs)); refs_off.append(len(refs))
        tix.append(-1 if getattr(n, "node_indices_id", None) is None else int(n.node_indices_id))
        for c in n.children.values():
            visit(c, n.id)

    visit(tree.root, -1)
    i64 = lambda x: np.asarray(x, dtype=np.int64)
    return dict(node_id=i64(ids), node_parent=i64(parents), kv_off=i64(kv_off), kv=i64(kv),
                refs_off=i64(refs_off), refs=i64(refs), leaf_ids=i64(sorted(tree.leaves.keys())),
                node_indices_id=i64(tix))


def thaw(a) -> PlainTree:
    t = PlainTree()
    for i, nid in enumerate(a["node_id"].tolist()):
        n = PlainNode(nid)
        n.kv_indices = a["kv"][a["kv_off"][i] : a["kv_off"][i + 1]].tolist()
        tix = int(a["node_indices_id"][i]) if "node_indices_id" in a else -1
        n.node_indices_id = None if tix < 0 else tix
        t.nodes[nid] = n
        p = int(a["node_parent"][i])
        if p < 0:
            t.root = n
        else:
            n.parent = t.nodes[p]
            t.nodes[p].children[nid] = n
    for lid in a["leaf_ids"].tolist():
        t.leaves[lid] = t.nodes[lid]
    for i, nid in enumerate(a["node_id"].tolist()):
        t.nodes[nid].refs = {t.nodes[r] for r in a["refs"][a["refs_off"][i] : a["refs_off"][i + 1]].tolist()}
    return t
