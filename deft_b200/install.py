"""Drop-in installation into an importable reference checkout.

The reference binds the two operators by name at import time
(``deft/layers/attention/deft_attention.py:7-10``), so the replacement rebinds those names in that
module's globals (and in ``deft.layers.attention.tree_attention`` for direct callers).  With
``metadata=True`` it also swaps ``TreeMetadata.from_tree_cache`` / ``from_tree_cache_node`` for the
C++ builder, which attaches the native work plan to the tables it returns.
``run_DeFT_llama_paged.py`` itself needs no edit: call ``deft_b200.install.install()`` before it runs
(e.g. from ``sitecustomize`` or ``python -c "import deft_b200.install as i; i.install(); import runpy; ..."``).
"""
from __future__ import annotations

import importlib
from typing import List

from . import attention, tree_cache

_PATCHED: List[tuple] = []


def _swap(module, name: str, value) -> None:
    # the raw attribute (for a class: the classmethod object itself, not the bound method getattr would build)
    old = vars(module)[name] if name in vars(module) else getattr(module, name)
    _PATCHED.append((module, name, old))
    setattr(module, name, value)


def install(metadata: bool = True) -> None:
    """Rebind the reference's tree-attention entry points to this package."""
    ta = importlib.import_module("deft.layers.attention.tree_attention")
    _swap(ta, "tree_attention_fwd", attention.tree_attention_fwd)
    _swap(ta, "tree_attention_subtree_fwd", attention.tree_attention_subtree_fwd)
    try:   # needs a GPU at import (context_flashattention_nopad.py:10)
        da = importlib.import_module("deft.layers.attention.deft_attention")
        _swap(da, "tree_attention_fwd", attention.tree_attention_fwd)
        _swap(da, "tree_attention_subtree_fwd", attention.tree_attention_subtree_fwd)
        _swap(da, "token_attention_fwd", attention.token_attention_fwd)       # Radix / seq mode (:174)
    except Exception:  # pragma: no cover - import of the caller failed; direct users are still patched
        pass
    if metadata:
        tc = importlib.import_module("deft.tree_decoding.tree_cache")
        _swap(tc.TreeMetadata, "from_tree_cache", classmethod(
            lambda cls, tree, tile_num=8, max_q_len=32, max_block_len=-1:
            _with_ref_block_config(tc, tree_cache.TreeMetadata.from_tree_cache, tree, tile_num, max_q_len, max_block_len)))
        _swap(tc.TreeMetadata, "from_tree_cache_node", classmethod(
            lambda cls, tree, tile_num=8, max_q_len=32, max_block_len=-1:
            _with_ref_block_config(tc, tree_cache.TreeMetadata.from_tree_cache_node, tree, tile_num, max_q_len, max_block_len)))


def _with_ref_block_config(tc, fn, tree, tile_num, max_q_len, max_block_len):
    # the CLI mutates the REFERENCE module's BLOCK_CONFIG (run_DeFT_llama_paged.py:145-150): honour it
    tree_cache.BLOCK_CONFIG.update(tc.BLOCK_CONFIG)
    return fn(tree, tile_num, max_q_len, max_block_len)


def uninstall() -> None:
    while _PATCHED:
        module, name, old = _PATCHED.pop()
        setattr(module, name, old)
