"""deft_b200 -- B200-native (sm_100a) tree-attention decode for DeFT.

Only what the tree-attention path needs lives here:

* ``csrc/``        hand-written CUDA kernels + the C ABI (``include/deft_b200.h``) -> ``lib/libdeft_b200.so``
* ``attention``    ``tree_attention_fwd`` / ``tree_attention_subtree_fwd`` with the reference's signatures
* ``tree_cache``   ``TreeCache`` / ``TreeMetadata`` mirror (host bookkeeping, C++ table builder)
* ``memory_pool``  paged KV pool with the reference's layout and allocation order
* ``install``      rebinding of the reference's operator names to this package (drop-in)

Importing the package loads the shared library and fails loudly if it is missing.
"""
import sys as _sys

if "deft_b200.build" in getattr(_sys, "orig_argv", []):      # `python -m deft_b200.build`: nothing to load yet
    from . import build as _build
    print(_build.build(force="--force" in _sys.argv, verbose="-v" in _sys.argv))
    raise SystemExit(0)

from . import _lib  # noqa: F401,E402  (raises ImportError when libdeft_b200.so is absent)
from .attention import kv_append, token_attention_fwd, tree_attention_fwd, tree_attention_subtree_fwd  # noqa: F401
from .decode_step import DecodeStepGraph, DecodeStepPipeline  # noqa: F401,E402
from .memory_pool import ReqToTokenPool, TokenToKVPool, TreeIndexPool  # noqa: F401
from .tree_cache import (BLOCK_CONFIG, KVCacheUpdater, TreeCache, TreeMetadata, TreeNode,  # noqa: F401
                         get_global_tree_metadata, register_tree_metadata, unregister_tree_metadata)

__version__ = "0.1.0"
