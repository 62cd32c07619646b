"""Synthetic decoding trees of the shapes BASELINE.json names, built through ``deft_b200.TreeCache``.

Page tables therefore follow the reference allocator exactly (prompt pages ``0..P-1`` contiguous,
step-``t`` leaf pages ``P + t*n_leaves + rank``), as SURVEY.md 8(d) prescribes.  Geometry is
Llama-3-8B: H=32 query heads, HKV=8, D=128, fp16.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .memory_pool import ReqToTokenPool, TokenToKVPool
from .tree_cache import TreeCache

LLAMA3_8B = dict(H=32, HKV=8, D=128, layers=32)

# name -> (prompt, [(fan-out applied to every leaf, decode steps after it), ...], description)
WORKLOADS: Dict[str, Tuple[int, List[Tuple[int, int]], str]] = {
    "cfg1": (512, [(1, 1)], "Llama-3-8B single sequence, prompt=512, 1 branch"),
    "cfg2": (4096, [(2, 16)] * 6, "Llama-3-8B DeFT-Flatten paged, prompt=4096, tree depth=6, 64 leaves, 16 tokens/node"),
    "cfg3": (2048, [(64, 1)], "Llama-3-8B speculative-decoding flat tree, prompt=2048, 64 one-token leaves"),
    "cfg4": (8192, [(2, 16)] * 8, "Llama-3-8B reasoning tree, prompt=8192, depth=8, 256 leaves, 16 tokens/node"),
}


def unique_kv_tokens(name: str) -> int:
    prompt, levels, _ = WORKLOADS[name]
    total, leaves = prompt, 1
    for fan, steps in levels:
        leaves *= fan
        total += leaves * steps
    return total


def n_leaves(name: str) -> int:
    leaves = 1
    for fan, _ in WORKLOADS[name][1]:
        leaves *= fan
    return leaves


def build_tree(name: str, layers: int, device="cuda", H: int = 32, HKV: int = 8, D: int = 128,
               headroom: int = 64) -> TreeCache:
    prompt, levels, _ = WORKLOADS[name]
    size = unique_kv_tokens(name) + headroom
    r2t = ReqToTokenPool(size=max(2 * n_leaves(name), 8), max_context_len=prompt + sum(s for _, s in levels) + 8,
                         device=device)
    kvp = TokenToKVPool(size=size, dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=layers, device=device)
    tree = TreeCache(torch.float16, HKV, D, layers, r2t, kvp, None, True, False)
    tree.init_prompt(torch.arange(prompt, dtype=torch.int32))
    for fan, steps in levels:
        for leaf in sorted(tree.leaves.values(), key=lambda x: x.id):
            tree.branch(leaf, fan)
        for _ in range(steps):
            for leaf in tree.leaves.values():
                leaf.append_token(7)
            tree.alloc()
    return tree


def build_forest(name: str, n_trees: int, layers: int, device="cuda", H: int = 32, HKV: int = 8, D: int = 128,
                 headroom: int = 64) -> List[TreeCache]:
    """``n_trees`` independent trees of one workload over ONE page pool (BASELINE cfg 5: batched decoding).

    Trees are grown one after the other, so tree ``t`` owns the pages ``[t * unique, (t + 1) * unique)`` with
    the same relative layout as a stand-alone tree (prompt contiguous, decode pages strided by its leaves).
    """
    prompt, levels, _ = WORKLOADS[name]
    size = unique_kv_tokens(name) * n_trees + headroom
    r2t = ReqToTokenPool(size=max(2 * n_leaves(name) * n_trees, 8), max_context_len=prompt + sum(s for _, s in levels) + 8,
                         device=device)
    kvp = TokenToKVPool(size=size, dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=layers, device=device)
    trees = []
    for _ in range(n_trees):
        tree = TreeCache(torch.float16, HKV, D, layers, r2t, kvp, None, True, False)
        tree.init_prompt(torch.arange(prompt, dtype=torch.int32))
        for fan, steps in levels:
            for leaf in sorted(tree.leaves.values(), key=lambda x: x.id):
                tree.branch(leaf, fan)
            for _ in range(steps):
                for leaf in tree.leaves.values():
                    leaf.append_token(7)
                tree.alloc()
        trees.append(tree)
    return trees


def algorithmic_bytes(name: str, H: int = 32, HKV: int = 8, D: int = 128) -> int:
    """Per layer-call: every unique KV token once (K and V) + Q read + O write."""
    return unique_kv_tokens(name) * 2 * HKV * D * 2 + 2 * n_leaves(name) * H * D * 2
