"""Pure-Python replay of the workload scripts on a fresh page pool.  TEST / BENCH INFRASTRUCTURE ONLY.

``bench.py --impl reference`` times the CPU port of the reference's sequence-based path and must not load the
CUDA library, so it cannot grow its trees through ``deft_b200.TreeCache``.  This restates, for the script operations
only, what the reference's ``TreeCache`` does to page lists on a FRESH pool
(``/root/reference/DeFT/deft/tree_decoding/tree_cache.py:192-283, 336-372`` with the first-free allocator of
``deft/memory_pool.py:74-80``): ``init_prompt`` takes pages ``0..P-1``, ``branch`` creates children with the next
node ids, ``alloc`` hands the next free pages to the leaves in ascending id order.
``tests/test_tree_cache.py`` checks the page lists against ``deft_b200.TreeCache`` on every workload.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple


class SimNode:
    def __init__(self, id: int, parent: Optional["SimNode"]) -> None:
        self.id = id
        self.parent = parent
        self.children: Dict[int, "SimNode"] = {}
        self.kv_indices: List[int] = []


class SimTree:
    def __init__(self) -> None:
        self.root: Optional[SimNode] = None
        self.nodes: Dict[int, SimNode] = {}
        self.leaves: Dict[int, SimNode] = {}
        self.next_page = 0
        self.node_cnt = 1

    def _take(self, n: int) -> List[int]:
        pages = list(range(self.next_page, self.next_page + n))
        self.next_page += n
        return pages

    def _leaves_sorted(self) -> List[SimNode]:
        return sorted(self.leaves.values(), key=lambda x: x.id)

    def _branch(self, node: SimNode, k: int) -> None:
        self.leaves.pop(node.id)
        for _ in range(k):
            child = SimNode(self.node_cnt, node)
            self.node_cnt += 1
            node.children[child.id] = child
            self.nodes[child.id] = child
            self.leaves[child.id] = child

    def replay(self, script: Sequence[Tuple]) -> "SimTree":
        for op in script:
            kind = op[0]
            if kind == "init":
                self.root = SimNode(0, None)
                self.nodes[0] = self.leaves[0] = self.root
                self.root.kv_indices = self._take(op[1])
            elif kind == "branch_all":
                for leaf in self._leaves_sorted():
                    self._branch(leaf, op[1])
            elif kind == "branch_counts":
                for leaf, c in zip(self._leaves_sorted(), op[1]):
                    if c > 0:
                        self._branch(leaf, c)
            elif kind == "step":
                for _ in range(op[1]):
                    leaves = self._leaves_sorted()
                    for leaf, page in zip(leaves, self._take(len(leaves))):
                        leaf.kv_indices.append(page)
            elif kind == "step_new":
                fresh = [leaf for leaf in self._leaves_sorted() if not leaf.kv_indices]
                for leaf, page in zip(fresh, self._take(len(fresh))):
                    leaf.kv_indices.append(page)
            else:
                raise ValueError(kind)
        return self
