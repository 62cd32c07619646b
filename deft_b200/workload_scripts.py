"""Synthetic decoding trees of the shapes BASELINE.json names, as SCRIPTS of ``TreeCache`` operations (pure Python).

This module imports nothing of the package, so that ``bench.py --impl reference`` (which must not load the CUDA
library) can read the same scripts by file path.  ``deft_b200.workloads`` re-exports everything here and adds the
replay through a ``TreeCache``.

Operations: ``("init", n)`` prompt of n tokens; ``("branch_all", k)`` every leaf gets k children;
``("step", s)`` s decode steps (one token + one page per leaf: ``leaf.append_token`` + ``tree.alloc``,
tree_generate.py:109); ``("branch_counts", [c...])`` the i-th leaf by id gets c_i children (0: stays a leaf);
``("step_new",)`` one token + one page for every leaf that has none yet (a token tree: one token per node).
Geometry is Llama-3-8B: H=32 query heads, HKV=8, D=128, fp16.
"""
from __future__ import annotations

import heapq
from typing import Dict, List, Sequence, Tuple

LLAMA3_8B = dict(H=32, HKV=8, D=128, layers=32)

Op = Tuple

def medusa_tree(width: int = 6, depth: int = 5, n_nodes: int = 63, p0: float = 0.6, decay: float = 0.4) -> List[Tuple[int, ...]]:
    """Sparse top-k token tree (SURVEY.md 8d cfg 3b): nodes are paths of child ranks ``(r1, .., rk)``, ``k <= depth``,
    ``r < width``; the ``n_nodes`` paths with the largest score ``prod p[r]``, ``p[r] = p0 * decay**r``, chosen best
    first (a parent always scores above its children, ties by path) -- how Medusa derives its ``Tree_Structure``."""
    p = [p0 * decay ** r for r in range(width)]
    heap = [(-p[r], (r,)) for r in range(width)]
    heapq.heapify(heap)
    chosen: List[Tuple[int, ...]] = []
    while heap and len(chosen) < n_nodes:
        s, path = heapq.heappop(heap)
        chosen.append(path)
        if len(path) < depth:
            for r in range(width):
                heapq.heappush(heap, (s * p[r], path + (r,)))
    return chosen


def _medusa_script(prompt: int, width: int = 6, depth: int = 5, n_nodes: int = 63) -> List[Op]:
    chosen = set(medusa_tree(width, depth, n_nodes))
    script: List[Op] = [("init", prompt)]
    level: List[Tuple[int, ...]] = [()]           # current leaves by ascending id = creation order
    finished: List[Tuple[int, ...]] = []          # leaves of earlier levels that stay leaves (smaller ids)
    for _ in range(depth):
        counts = [sum(1 for r in range(width) if path + (r,) in chosen) for path in level]
        if not any(counts):
            break
        # leaves by id: the childless ones of earlier levels first (created earlier), then this level's
        script.append(("branch_counts", [0] * len(finished) + counts))
        script.append(("step_new",))
        finished += [path for path, c in zip(level, counts) if c == 0]
        level = [path + (r,) for path, c in zip(level, counts) for r in range(c)]
    return script


# name -> (script, description)
WORKLOADS: Dict[str, Tuple[List[Op], str]] = {
    "cfg1": ([("init", 512), ("branch_all", 1), ("step", 1)], "Llama-3-8B single sequence, prompt=512, 1 branch"),
    "cfg2": ([("init", 4096)] + [("branch_all", 2), ("step", 16)] * 6,
             "Llama-3-8B DeFT-Flatten paged, prompt=4096, tree depth=6, 64 leaves, 16 tokens/node"),
    "cfg3": ([("init", 2048), ("branch_all", 64), ("step", 1)],
             "Llama-3-8B speculative-decoding flat tree (the reference's own mock), prompt=2048, 64 one-token leaves"),
    "cfg3b": (_medusa_script(2048), "Llama-3-8B Medusa-style sparse token tree, width=6 depth=5, 63 one-token nodes "
                                    "(best-first by path score), prompt=2048, queries = leaves"),
    "cfg4": ([("init", 8192)] + [("branch_all", 2), ("step", 16)] * 8,
             "Llama-3-8B reasoning tree, prompt=8192, depth=8, 256 leaves, 16 tokens/node"),
}


def _simulate(script: Sequence[Op]) -> Tuple[int, int, int, int]:
    """(unique KV tokens, leaves, longest root->leaf path, nodes) of a script, without building anything."""
    leaves: List[List[int]] = []          # per leaf: [own tokens, path tokens above it]
    total = nodes = 0
    for op in script:
        if op[0] == "init":
            leaves, total, nodes = [[op[1], 0]], op[1], 1
        elif op[0] in ("branch_all", "branch_counts"):
            counts = [op[1]] * len(leaves) if op[0] == "branch_all" else list(op[1])
            kept = [lf for lf, c in zip(leaves, counts) if c == 0]
            new = [[0, lf[0] + lf[1]] for lf, c in zip(leaves, counts) for _ in range(c)]
            nodes += len(new)
            leaves = kept + new           # ids ascend in creation order
        elif op[0] == "step":
            for lf in leaves:
                lf[0] += op[1]
            total += op[1] * len(leaves)
        elif op[0] == "step_new":
            for lf in leaves:
                if lf[0] == 0:
                    lf[0] = 1
                    total += 1
    return total, len(leaves), max(lf[0] + lf[1] for lf in leaves), nodes


def unique_kv_tokens(name: str) -> int:
    return _simulate(WORKLOADS[name][0])[0]


def n_leaves(name: str) -> int:
    return _simulate(WORKLOADS[name][0])[1]


def max_path_len(name: str) -> int:
    return _simulate(WORKLOADS[name][0])[2]


def n_nodes(name: str) -> int:
    return _simulate(WORKLOADS[name][0])[3]


def algorithmic_bytes(name: str, H: int = 32, HKV: int = 8, D: int = 128) -> int:
    """Per layer-call: every unique KV token once (K and V) + Q read + O write."""
    return unique_kv_tokens(name) * 2 * HKV * D * 2 + 2 * n_leaves(name) * H * D * 2


def algorithmic_flops(name: str, H: int = 32, D: int = 128) -> int:
    """Per layer-call: sum over nodes of len(node) * |queries attending it| * H * 4 D (QK^T and PV)."""
    # every leaf attends its whole root->leaf path: sum over nodes len * |Q(node)| = sum over leaves of path length
    leaves: List[List[int]] = []
    for op in WORKLOADS[name][0]:
        if op[0] == "init":
            leaves = [[op[1], 0]]
        elif op[0] in ("branch_all", "branch_counts"):
            counts = [op[1]] * len(leaves) if op[0] == "branch_all" else list(op[1])
            leaves = ([lf for lf, c in zip(leaves, counts) if c == 0]
                      + [[0, lf[0] + lf[1]] for lf, c in zip(leaves, counts) for _ in range(c)])
        elif op[0] == "step":
            for lf in leaves:
                lf[0] += op[1]
        elif op[0] == "step_new":
            for lf in leaves:
                lf[0] = lf[0] or 1
    return sum(lf[0] + lf[1] for lf in leaves) * H * 4 * D
