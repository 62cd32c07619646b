"""The synthetic workloads (``workload_scripts``) replayed through a ``TreeCache``.

A script is replayed through any ``TreeCache``-shaped object -- ``deft_b200.TreeCache`` here, the reference's
``deft.tree_decoding.tree_cache.TreeCache`` in ``tools/ref_triton_probe.py`` -- so page tables follow the reference
allocator exactly (prompt pages ``0..P-1`` contiguous, step-``t`` leaf pages ``P + t*n_leaves + rank``), as
SURVEY.md 8(d) prescribes.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .memory_pool import ReqToTokenPool, TokenToKVPool
from .tree_cache import TreeCache
from .workload_scripts import (LLAMA3_8B, WORKLOADS, Op, _simulate, algorithmic_bytes, algorithmic_flops,  # noqa: F401
                               max_path_len, medusa_tree, n_leaves, n_nodes, unique_kv_tokens)


def _leaves_sorted(tree):
    return sorted(tree.leaves.values(), key=lambda x: x.id)


def replay(tree, script: Sequence[Op], make_ids=lambda n: torch.arange(n, dtype=torch.int32)) -> None:
    """Replays a script through ``tree`` (ours or the reference's ``TreeCache``)."""
    for op in script:
        kind = op[0]
        if kind == "init":
            tree.init_prompt(make_ids(op[1]))
        elif kind == "branch_all":
            for leaf in _leaves_sorted(tree):
                tree.branch(leaf, op[1])
        elif kind == "branch_counts":
            leaves = _leaves_sorted(tree)
            assert len(leaves) == len(op[1]), (len(leaves), len(op[1]))
            for leaf, c in zip(leaves, op[1]):
                if c > 0:
                    tree.branch(leaf, c)
        elif kind == "step":
            for _ in range(op[1]):
                for leaf in tree.leaves.values():
                    leaf.append_token(7)
                tree.alloc()
        elif kind == "step_new":
            # what TreeCache.alloc does (tree_cache.py:261-283), for the leaves that hold no page yet
            fresh = [leaf for leaf in _leaves_sorted(tree) if len(leaf.kv_indices) == 0]
            locs = tree.token_to_kv_pool.alloc(len(fresh))
            assert locs is not None
            table = tree.req_to_token_pool.req_to_token
            for leaf, loc in zip(fresh, locs.tolist()):
                leaf.append_token(7)
                leaf.append_index(int(loc))
                table[tree.leaf_to_req[leaf.id], leaf.positions[-1]] = int(loc)
        else:
            raise ValueError(kind)


def build_forest(name: str, n_trees: int, layers: int, device="cuda", H: int = 32, HKV: int = 8, D: int = 128,
                 headroom: int = 64) -> List[TreeCache]:
    """``n_trees`` independent trees of one workload over ONE page pool (BASELINE cfg 5: batched decoding).

    Trees are grown one after the other, so tree ``t`` owns the pages ``[t * unique, (t + 1) * unique)`` with
    the same relative layout as a stand-alone tree (prompt contiguous, decode pages strided by its leaves).
    ``headroom``: free pages left in the pool per tree (a decode loop takes one per leaf per step).
    """
    script = WORKLOADS[name][0]
    unique, leaves, path, nodes = _simulate(script)
    # a branch hands the parent's request slot to its first child: a tree never holds more slots than it ends up with leaves
    r2t = ReqToTokenPool(size=leaves * n_trees + 8, max_context_len=path + headroom + 8, device=device)
    kvp = TokenToKVPool(size=(unique + headroom) * n_trees, dtype=torch.float16, head_num=HKV, head_dim=D,
                        layer_num=layers, device=device)
    trees = []
    for _ in range(n_trees):
        tree = TreeCache(torch.float16, HKV, D, layers, r2t, kvp, None, True, False)
        replay(tree, script)
        trees.append(tree)
    return trees


def build_tree(name: str, layers: int, device="cuda", H: int = 32, HKV: int = 8, D: int = 128,
               headroom: int = 64) -> TreeCache:
    return build_forest(name, 1, layers, device, H, HKV, D, headroom)[0]


