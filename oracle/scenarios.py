"""Tree-building scripts shared by ``oracle/gen_golden.py`` and the tests.  TEST INFRASTRUCTURE ONLY.

A scenario is a list of operations replayed through a ``TreeCache``-shaped API (``init_prompt``,
``branch``, ``alloc``, ``cut``, ``merge_nodes``, ``reset_node_KV`` and ``leaf.append_token``).
``gen_golden.py`` replays them through the REFERENCE ``TreeCache``
(``/root/reference/DeFT/deft/tree_decoding/tree_cache.py:147-403``) and records the resulting page
tables; the tests replay the same scripts through ``deft_b200.tree_cache.TreeCache`` and require
bit-identical pages.  The usage pattern follows ``/root/reference/DeFT/tests/model/test_DeFT_kernel.py:66-117``
and the speculative mock in ``deft/tree_decoding/generation/branch_func_example.py:411-440``.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

Op = Tuple


def _leaves_sorted(tree):
    return sorted(tree.leaves.values(), key=lambda n: n.id)


def replay(tree, script: List[Op], make_ids) -> None:
    """``make_ids(n)`` returns the prompt-id container the API expects (torch tensor for the reference)."""
    for op in script:
        kind = op[0]
        if kind == "init":
            tree.init_prompt(make_ids(op[1]))
        elif kind == "branch":          # ("branch", k-th leaf by id, fan-out)
            tree.branch(_leaves_sorted(tree)[op[1]], op[2])
        elif kind == "branch_all":      # every current leaf gets op[1] children
            for leaf in _leaves_sorted(tree):
                tree.branch(leaf, op[1])
        elif kind == "step":            # op[1] decode steps: one token + one page per leaf
            for _ in range(op[1]):
                for leaf in tree.leaves.values():
                    leaf.append_token(7)
                tree.alloc()
        elif kind == "cut":
            tree.cut(_leaves_sorted(tree)[op[1]])
        elif kind == "spec":            # mock speculative verification: squeeze op[1] leaves into the root
            leaves = list(tree.leaves.values())
            before = len(tree.root.kv_indices)
            for i in range(op[1]):
                tree.merge_nodes(tree.root, leaves[i], pruneB_flag=False)
            diff = len(tree.root.kv_indices) - before
            for leaf in leaves:
                tree.reset_node_KV(leaf, diff)
        else:
            raise ValueError(kind)


# name -> (geometry, pool sizes, script)
SCENARIOS: Dict[str, dict] = {
    # balanced binary tree, toy heads
    "toy_binary": dict(H=8, HKV=2, D=64, pool=420, max_ctx=256,
                       script=[("init", 150), ("branch_all", 2), ("step", 5), ("branch_all", 2), ("step", 5),
                               ("branch_all", 2), ("step", 5)]),
    # > 32 queries: exercises the 32-query sub-block split
    "wide40": dict(H=8, HKV=2, D=64, pool=420, max_ctx=256,
                   script=[("init", 131), ("branch", 0, 40), ("step", 2)]),
    # ragged node lengths around the 128-token block edge, pruning, unbalanced fan-out
    "ragged_cut": dict(H=8, HKV=2, D=64, pool=1300, max_ctx=700,
                       script=[("init", 127), ("branch", 0, 3), ("step", 1), ("branch", 1, 2), ("step", 128),
                               ("cut", 0), ("branch", 0, 2), ("step", 129), ("cut", 2), ("branch", 2, 2), ("step", 3)]),
    # speculative mock: root pages become non-contiguous, leaves are re-allocated
    "spec_merge": dict(H=8, HKV=2, D=64, pool=420, max_ctx=256,
                       script=[("init", 140), ("branch", 0, 12), ("step", 1), ("spec", 3), ("step", 1), ("spec", 2),
                               ("step", 1)]),
    # Llama-3-8B head geometry, 400-token prompt + flat leaves (the reference's own kernel test shape, scaled)
    "llama_flat8": dict(H=32, HKV=8, D=128, pool=460, max_ctx=450,
                        script=[("init", 400), ("branch", 0, 8), ("step", 1)]),
    # single sequence (BASELINE configs[0] shape, scaled)
    "single_seq": dict(H=8, HKV=2, D=64, pool=300, max_ctx=300,
                       script=[("init", 200), ("branch", 0, 1), ("step", 1)]),
    # tree-index mode (node -> page table)
    "tree_index": dict(H=8, HKV=2, D=64, pool=420, max_ctx=256, tree_index=True,
                       script=[("init", 150), ("branch_all", 3), ("step", 4), ("branch", 1, 2), ("step", 2)]),
}

# table-only scenarios at BASELINE.json sizes (no attention run: the interpreter is too slow there)
TABLE_SCENARIOS: Dict[str, dict] = {
    "cfg2_tables": dict(script=[("init", 4096)] + [("branch_all", 2), ("step", 16)] * 6, pool=4096 + 2016 + 64, max_ctx=4300),
    "cfg3a_tables": dict(script=[("init", 2048), ("branch", 0, 64), ("step", 1)], pool=2048 + 64 + 64, max_ctx=2100),
}
