"""Decoding-tree state and the metadata that drives tree attention (host side).

Mirrors, for the paged tree-attention path, the interface of the reference's
``deft/tree_decoding/tree_cache.py``: ``TreeNode`` (:94-129), ``KVCacheUpdater`` (:53-91),
``TreeCache`` (:147-584: ``init_prompt`` / ``new_node`` / ``alloc`` / ``merge_nodes`` /
``reset_node_KV`` / ``branch`` / ``cut`` / ``add_ref`` / ``remove_ref``), the ``TreeMetadata``
dataclass (:591-616) with ``from_tree_cache`` (:618-881) / ``from_tree_cache_node`` (:883-1018) and
the module-level registry (:1021-1052).  Page ids, leaf order and every table are bit-identical to
the reference's on the same sequence of operations (tests/test_tree_cache.py, tests/test_tables.py).

What is different underneath:
  * all bookkeeping is host-side; no ``.item()`` device round trips (the reference syncs once per
    leaf per step in ``alloc``);
  * ``from_tree_cache`` flattens the tree to arrays and calls the C++ builder
    (``deft_b200_build_tables``), which returns the reference tables AND the native work plan in one
    packed buffer that is uploaded with ONE host->device copy;
  * ``KVCacheUpdater.update`` scatters K and V in one launch (``deft_b200_kv_append``).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field
from itertools import chain
from typing import Any, Dict, List, Optional, Set

import numpy as np
import torch

from . import _lib
from .memory_pool import ReqToTokenPool, TokenToKVPool, TreeIndexPool

BLOCK_CONFIG = {"BLOCK_LEN": 128, "MAX_BLOCK_LEN": -1}   # tree_cache.py:587
TRAVERSAL_CONFIG = {"METHOD": "dfs"}
NODE_SPLIT = 256   # tokens per work item when a Node entry is longer (csrc/common.cuh kNodeSplit)


class KVCacheUpdater:
    """Writes this step's K/V rows into the pages handed out by ``alloc`` / ``init_prompt``."""

    def __init__(self, token_to_kv_pool: TokenToKVPool, cache_loc: torch.Tensor, is_prompt: bool) -> None:
        self.use_paged_memory = True
        self.token_to_kv_pool = token_to_kv_pool
        self.cache_loc = cache_loc                     # int32, host
        self.is_prompt = is_prompt
        self._loc_dev: Optional[torch.Tensor] = None

    def device_loc(self) -> torch.Tensor:
        if self._loc_dev is None:
            self._loc_dev = self.cache_loc.to(self.token_to_kv_pool.device, non_blocking=True)
        return self._loc_dev

    def update(self, layer_id: int, cache_k: torch.Tensor, cache_v: torch.Tensor) -> None:
        from .attention import kv_append   # local import: attention imports this module
        kv = self.token_to_kv_pool.kv_data[layer_id]
        kv_append(kv, cache_k, cache_v, self.device_loc())


_PAUSE_EPOCH = [0]     # moved whenever a node's ``paused`` flag changes (part of flatten_tree's cache key)
_PAGES_EPOCH = [0]     # moved when a node outside any TreeCache swaps or appends to its page list


class TreeNode:
    @property
    def paused(self) -> bool:
        return self._paused

    @property
    def kv_indices(self) -> List[int]:
        return self._kv_indices

    @kv_indices.setter
    def kv_indices(self, value: List[int]) -> None:
        """A page list swapped for another object (init_prompt, reset_node_KV, or the caller's own code): whatever
        mirrors the tree's pages (``TreeCache.native_tree``) is out of date."""
        self._kv_indices = value
        self._touch_pages()

    def _touch_pages(self) -> None:
        owner = self._owner
        if owner is not None:
            owner._pages_epoch += 1
        else:
            _PAGES_EPOCH[0] += 1

    @paused.setter
    def paused(self, value: bool) -> None:
        self._paused = bool(value)
        _PAUSE_EPOCH[0] += 1

    def __init__(self, id: int, node_indices_id: Optional[int] = None,
                 node_indices: Optional[torch.Tensor] = None) -> None:
        self.id = id
        self._owner = None                    # the TreeCache that made the node (told when the page list changes)
        self.children: Dict[int, "TreeNode"] = {}
        self.token_ids: List[int] = []
        self.positions: List[int] = []
        self.position_offset = 0
        self._kv_indices: List[int] = []
        self.parent: Optional["TreeNode"] = None
        self.refs: Set["TreeNode"] = set()
        self._paused = False
        self.node_indices_id = node_indices_id
        self.node_indices = node_indices
        self.index_pool = None                # set by TreeCache in tree-index mode: told when node_indices is written
        self.cumulative_logprob = 0.0

    def get_len(self) -> int:
        return len(self.token_ids)

    def append_token(self, token: int, logprob: Optional[float] = None) -> None:
        self.positions.append(self.position_offset + len(self.token_ids))
        self.token_ids.append(token)
        if logprob is not None:
            self.cumulative_logprob += logprob

    def append_index(self, index: int) -> None:
        self._kv_indices.append(index)
        self._touch_pages()                  # (``TreeCache.alloc`` appends in bulk and tells the mirror itself)
        if self.node_indices is not None:
            self.node_indices[len(self._kv_indices) - 1] = index
            if self.index_pool is not None:
                self.index_pool.touch()


class TreeCache:
    """Tree topology + per-node page lists over a paged KV pool."""

    def __init__(self, dtype: torch.dtype, head_num: int, head_dim: int, layer_num: int,
                 req_to_token_pool: ReqToTokenPool, token_to_kv_pool: TokenToKVPool,
                 tree_index_pool: Optional[TreeIndexPool] = None, use_paged_memory: bool = True,
                 use_tree_index: bool = False) -> None:
        if not use_paged_memory:
            raise NotImplementedError("deft_b200 covers the paged tree-attention path only "
                                      "(the unpaged modes are the reference's baselines)")
        assert token_to_kv_pool is not None and req_to_token_pool is not None
        if use_tree_index:
            assert tree_index_pool is not None
        self.node_cnt = 1
        self.root: Optional[TreeNode] = None
        self.nodes: Dict[int, TreeNode] = {}
        self.leaves: Dict[int, TreeNode] = {}
        self.leaf_to_req: Dict[int, int] = {}
        self.paused_nodes: Set[int] = set()
        self.leaf_to_q: Dict[int, int] = {}
        self.req_to_token_pool = req_to_token_pool
        self.token_to_kv_pool = token_to_kv_pool
        self.tree_index_pool = tree_index_pool
        self.use_paged_memory = True
        self.use_tree_index = use_tree_index
        self.layer_num = layer_num
        self.deleted_token_num = 0
        self._topo_version = 0      # bumped by every change of nodes / leaves / refs: flatten_tree keeps the walk
        self._pages_epoch = 0       # bumped when a node's page list changes other than through alloc()
        self._mirror: Optional["_NativeTree"] = None

    # ---- construction -------------------------------------------------------------------
    def _take_index_row(self):
        if not self.use_tree_index:
            return None, None
        rows = self.tree_index_pool.alloc(1)
        assert rows is not None
        rid = int(rows[0])
        return rid, self.tree_index_pool.node_to_kv[rid]

    def init_prompt(self, prompt_ids) -> KVCacheUpdater:
        ids = prompt_ids.tolist() if hasattr(prompt_ids, "tolist") else list(prompt_ids)
        n = len(ids)
        rid, row = self._take_index_row()
        self._topo_version += 1
        root = TreeNode(0, rid, row)
        root._owner = self
        root.index_pool = self.tree_index_pool if self.use_tree_index else None
        root.token_ids = ids
        root.positions = list(range(n))
        self.root = root
        self.nodes[0] = root
        self.leaves[0] = root
        self.add_ref(root)
        req = self.req_to_token_pool.alloc(1)
        assert req is not None
        req_id = int(req[0])
        self.leaf_to_req[0] = req_id
        cache_loc = self.token_to_kv_pool.alloc(n)
        assert cache_loc is not None
        root.kv_indices = cache_loc.tolist()
        self.req_to_token_pool.req_to_token[req_id, :n] = cache_loc
        if row is not None:
            row[:n] = cache_loc
            self.tree_index_pool.touch()
        return KVCacheUpdater(self.token_to_kv_pool, cache_loc, is_prompt=True)

    def new_node(self, parent: TreeNode) -> TreeNode:
        self._topo_version += 1
        rid, row = self._take_index_row()
        node = TreeNode(self.node_cnt, rid, row)
        node._owner = self
        node.index_pool = self.tree_index_pool if self.use_tree_index else None
        self.node_cnt += 1
        node.parent = parent
        node.position_offset = parent.position_offset + len(parent.positions)
        parent.children[node.id] = node
        self.nodes[node.id] = node
        return node

    def alloc(self) -> KVCacheUpdater:
        """One new page per leaf, leaves in ascending id order (tree_cache.py:261-283)."""
        leaves = [self.leaves[i] for i in sorted(self.leaves)]
        out_cache_loc = self.token_to_kv_pool.alloc(len(leaves))
        assert out_cache_loc is not None
        locs = out_cache_loc.tolist()
        mirror = self._mirror
        in_sync = mirror is not None and mirror.key == self._mirror_key()
        if self.use_tree_index:
            for leaf, loc in zip(leaves, locs):
                leaf.append_index(loc)
        else:
            for leaf, loc in zip(leaves, locs):
                leaf._kv_indices.append(loc)
        if in_sync:       # the native mirror follows: one call with the step's pages (deft_b200_tree_append)
            mirror.append(locs)
            mirror.key = self._mirror_key()
        # the page table of the sequence-based baseline, in one indexed store (the host tensor and the array share memory)
        table = self.req_to_token_pool.req_to_token.numpy()
        l2r = self.leaf_to_req
        table[[l2r[leaf.id] for leaf in leaves], [leaf.positions[-1] for leaf in leaves]] = locs
        return KVCacheUpdater(self.token_to_kv_pool, out_cache_loc, is_prompt=False)

    # ---- native mirror --------------------------------------------------------------------
    def _mirror_key(self):
        return (self._topo_version, self._pages_epoch, _PAUSE_EPOCH[0], _PAGES_EPOCH[0], id(self.root))

    def native_tree(self) -> "_NativeTree":
        """The C++ mirror of this tree (``deft_tree_t``), brought up to date: a decode loop that only
        calls ``alloc`` between builds keeps it in step with one ``deft_b200_tree_append`` per ``alloc``; anything else
        (branch, cut, merge, pause, a swapped page list) is followed by a full ``deft_b200_tree_set`` here."""
        m = self._mirror
        if m is None:
            m = self._mirror = _NativeTree()
        key = self._mirror_key()
        # (the sum catches in-place edits of a page list that change its length; code that rewrites pages in place
        # without going through the tree calls invalidate_native_tree())
        if m.key != key or m.n_pages != sum(map(len, m.lists)):
            m.sync(self, flatten_tree(self))
            m.key = key
        return m

    def invalidate_native_tree(self) -> None:
        """Forces a full hand-over of the tree at the next table build."""
        self._pages_epoch += 1

    # ---- mutation -------------------------------------------------------------------------
    def merge_nodes(self, node_A: TreeNode, node_B: TreeNode, pruneB_flag: Optional[bool] = True) -> None:
        """Append B's tokens and pages to A (speculative verification, tree_cache.py:300-327)."""
        for token_id in node_B.token_ids:
            # the reference appends the position twice (once here, once in append_token)
            node_A.positions.append(node_A.position_offset + len(node_A.token_ids))
            node_A.append_token(token=token_id)
        for kv_idx in node_B.kv_indices:
            node_A.append_index(index=kv_idx)
        self.token_to_kv_pool.add_refs(node_B.kv_indices)
        if pruneB_flag:
            self.cut(node_B)

    def reset_node_KV(self, node: TreeNode, diff: int) -> None:
        self.token_to_kv_pool.free(node.kv_indices)
        node.kv_indices = []
        node.position_offset += diff
        node.positions = [pos + diff for pos in node.positions]

    def branch(self, node: TreeNode, branch_cnt: int) -> List[TreeNode]:
        assert node.id in self.leaves
        self._topo_version += 1
        self.leaves.pop(node.id)
        path_len = node.positions[-1] + 1
        req = self.leaf_to_req.pop(node.id)
        new_nodes: List[TreeNode] = []
        for i in range(branch_cnt):
            child = self.new_node(node)
            new_nodes.append(child)
            self.leaves[child.id] = child
            if i == 0:
                self.leaf_to_req[child.id] = req          # the first child inherits the parent's slot
            else:
                fresh = self.req_to_token_pool.alloc(1)
                assert fresh is not None
                fresh_id = int(fresh[0])
                self.req_to_token_pool.copy(req, fresh_id, path_len)
                self.leaf_to_req[child.id] = fresh_id
        self.remove_ref(node)
        for child in new_nodes:
            self.add_ref(child)
        return new_nodes

    def cut(self, node: TreeNode, record_deleted: bool = False) -> List[TreeNode]:
        assert len(node.children) == 0
        assert node.id in self.leaves
        self._topo_version += 1
        self.leaves.pop(node.id)
        self.remove_ref(node)
        self.req_to_token_pool.free(self.leaf_to_req.pop(node.id))
        assert len(node.refs) == 0
        deleted: List[TreeNode] = []
        cur: Optional[TreeNode] = node
        while cur is not None and len(cur.refs) == 0:
            deleted.append(self.nodes.pop(cur.id))
            self.token_to_kv_pool.free(cur.kv_indices)
            if self.use_tree_index:
                assert cur.node_indices_id is not None
                self.tree_index_pool.free(cur.node_indices_id)
            parent = cur.parent
            if parent is not None:
                parent.children.pop(cur.id)
            cur = parent
        if record_deleted:
            self.deleted_token_num += sum(len(d.token_ids) for d in deleted)
        return deleted

    def add_ref(self, node: TreeNode) -> None:
        self._topo_version += 1
        ref = node
        cur: Optional[TreeNode] = node
        while cur is not None:
            cur.refs.add(ref)
            cur = cur.parent

    def remove_ref(self, node: TreeNode) -> None:
        self._topo_version += 1
        ref = node
        cur: Optional[TreeNode] = node
        while cur is not None:
            cur.refs.remove(ref)
            cur = cur.parent

    def free(self) -> None:
        self._topo_version += 1
        self.root = None
        self.nodes.clear()
        self.leaves.clear()
        self.node_cnt = 0

    def get_tree_token_number(self) -> int:
        return sum(len(n.token_ids) for n in self.nodes.values()) + self.deleted_token_num


# ------------------------------------------------------------------------------------------------
# metadata
# ------------------------------------------------------------------------------------------------
def _node_pages(node) -> np.ndarray:
    """``node.kv_indices`` as int64, converted once and extended as the list grows.

    The page list of a node only ever grows by appends (``append_index``, tree_cache.py:126) or is replaced by a
    new list object (:210, :332); the cache is keyed on the list object and its first/last cached entries, and a
    list that shrank or was swapped is converted afresh.  (4096 prompt pages cost 0.25 ms to convert: more than
    the C++ builder takes for the whole tree.)
    """
    lst = node.kv_indices
    n = len(lst)
    c = getattr(node, "_kv_np", None)
    if c is not None:
        src, arr, m, first, last, view = c
        if src is lst and m <= n and (m == 0 or (first == lst[0] and last == lst[m - 1])):
            if m == n:
                return view
            if n > arr.shape[0]:
                grown = np.empty(max(n + 64, 2 * arr.shape[0]), dtype=np.int64)
                grown[:m] = arr[:m]
                arr = grown
            if n == m + 1:
                arr[m] = lst[m]              # a decode step: one page more
            else:
                arr[m:n] = lst[m:n]
            view = arr[:n]
            node._kv_np = (lst, arr, n, first if m else lst[0], lst[n - 1], view)
            return view
    arr = np.empty(n + 64, dtype=np.int64)
    arr[:n] = lst
    view = arr[:n]
    try:
        node._kv_np = (lst, arr, n, lst[0] if n else 0, lst[n - 1] if n else 0, view)
    except AttributeError:        # a node type with __slots__: no cache
        pass
    return view


def _walk(tree):
    """DFS pre-order walk: (nodes, parent, q_off, qs, tix, leaf_to_q)."""
    leaf_to_q = {lid: i for i, lid in enumerate(sorted(tree.leaves))}      # leaves is keyed by leaf id
    nodes: List[Any] = []
    parent: List[int] = []
    qs: List[int] = []
    q_len: List[int] = []
    tix: List[int] = []
    stack = [(tree.root, -1)]
    while stack:
        node, par = stack.pop()
        if node.paused:
            continue
        me = len(parent)
        nodes.append(node)
        parent.append(par)
        n0 = len(qs)
        qs.extend([leaf_to_q[r.id] for r in node.refs if not r.paused])
        q_len.append(len(qs) - n0)
        tix.append(-1 if node.node_indices_id is None else node.node_indices_id)
        if node.children:
            for child in reversed(list(node.children.values())):
                stack.append((child, me))
    q_off = np.zeros(len(parent) + 1, dtype=np.int64)
    np.cumsum(q_len, out=q_off[1:])
    return (nodes, np.asarray(parent, dtype=np.int32), q_off, np.asarray(qs, dtype=np.int64),
            np.asarray(tix, dtype=np.int64), leaf_to_q)


def flatten_tree(tree) -> Dict[str, Any]:
    """Any ``TreeCache``-shaped object -> the flat arrays ``deft_b200_build_tables`` takes.

    Nodes in DFS pre-order, children in dict (creation) order -- the visiting order of
    tree_cache.py:725-791; queries = rank of each ref'd leaf by ascending leaf id (:650-652).
    A decode step that only appended pages (``alloc``) keeps the topology: for our ``TreeCache`` (which counts
    its structural changes) the walk is kept and only the page lists are gathered again.
    """
    ver = getattr(tree, "_topo_version", None)
    topo = None
    if ver is not None:
        key = (ver, _PAUSE_EPOCH[0], id(tree.root))
        c = getattr(tree, "_flat_topo", None)
        if c is not None and c[0] == key:
            topo = c[1]
    if topo is None:
        topo = _walk(tree)
        if ver is not None:
            tree._flat_topo = (key, topo)
    nodes, parent, q_off, qs, tix, leaf_to_q = topo
    kv_arrs = [_node_pages(n) for n in nodes]
    kv_off = np.zeros(len(nodes) + 1, dtype=np.int64)
    np.cumsum([a.shape[0] for a in kv_arrs], out=kv_off[1:])
    kv = np.concatenate(kv_arrs) if kv_arrs else np.zeros(0, dtype=np.int64)
    return dict(parent=parent, kv_off=kv_off, kv=kv, q_off=q_off, qs=qs, tix=tix, leaf_to_q=leaf_to_q)


def flatten_forest(trees) -> Dict[str, Any]:
    """Several independent trees over ONE page pool -> one set of flat arrays (batched decoding).

    Trees are laid one after the other in DFS pre-order; the queries of tree ``t`` follow those of tree
    ``t - 1`` (its leaves ranked by id, offset by the number of leaves before it).  The reference has no
    batched mode (one tree per ``TreeMetadata``); this is the multi-tree extension SURVEY.md 8(d) cfg 5 asks for.
    """
    parts = [flatten_tree(t) for t in trees]
    n_nodes = np.fromiter((len(p["parent"]) for p in parts), dtype=np.int64, count=len(parts))
    n_kv = np.fromiter((len(p["kv"]) for p in parts), dtype=np.int64, count=len(parts))
    n_qs = np.fromiter((len(p["qs"]) for p in parts), dtype=np.int64, count=len(parts))
    n_q = np.fromiter((len(p["leaf_to_q"]) for p in parts), dtype=np.int64, count=len(parts))

    def starts(counts: np.ndarray) -> np.ndarray:          # exclusive prefix sums
        out = np.zeros(len(counts), dtype=np.int64)
        np.cumsum(counts[:-1], out=out[1:])
        return out

    node_off, kv_base, qs_base, q_base = starts(n_nodes), starts(n_kv), starts(n_qs), starts(n_q)
    parent = np.concatenate([p["parent"] for p in parts]).astype(np.int64)
    parent = np.where(parent >= 0, parent + np.repeat(node_off, n_nodes), -1).astype(np.int32)
    kv = np.concatenate([p["kv"] for p in parts])
    qs = np.concatenate([p["qs"] for p in parts]) + np.repeat(q_base, n_qs)
    kv_off = np.zeros(int(n_nodes.sum()) + 1, dtype=np.int64)
    kv_off[1:] = np.concatenate([p["kv_off"][1:] for p in parts]) + np.repeat(kv_base, n_nodes)
    q_off = np.zeros(int(n_nodes.sum()) + 1, dtype=np.int64)
    q_off[1:] = np.concatenate([p["q_off"][1:] for p in parts]) + np.repeat(qs_base, n_nodes)
    leaf_to_q: Dict[Any, int] = {}
    for i, p in enumerate(parts):
        b = int(q_base[i])
        for leaf, q in p["leaf_to_q"].items():
            leaf_to_q[(i, leaf)] = q + b
    return dict(parent=parent, kv_off=kv_off, kv=kv, q_off=q_off, qs=qs, tix=np.concatenate([p["tix"] for p in parts]),
                leaf_to_q=leaf_to_q)


class _Staging:
    """Ring of pinned host buffers for the one-copy table upload."""

    def __init__(self, depth: int = 4) -> None:
        self.bufs: List[Optional[torch.Tensor]] = [None] * depth
        self.events: List[Optional[torch.cuda.Event]] = [None] * depth
        self.i = 0

    def reserve(self, n: int) -> torch.Tensor:
        """The next pinned buffer of the ring, at least ``n`` bytes, no longer read by an earlier copy."""
        slot = self.i
        self.i = (self.i + 1) % len(self.bufs)
        if self.events[slot] is not None:
            self.events[slot].synchronize()
            self.events[slot] = None
        buf = self.bufs[slot]
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1 << 20), dtype=torch.uint8, pin_memory=True)
            self.bufs[slot] = buf
        self._slot = slot
        return buf

    def send(self, buf: torch.Tensor, n: int, device: torch.device, into: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One async H2D copy of the first ``n`` bytes of the reserved buffer; into the first bytes of ``into`` when
        given (a persistent device buffer: the tables then keep their addresses from one decode step to the next),
        else into a fresh tensor."""
        out = into[:n] if into is not None else torch.empty(n, dtype=torch.uint8, device=device)
        out.copy_(buf[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[self._slot] = ev
        return out

    def upload(self, src: np.ndarray, device: torch.device, into: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = src.nbytes
        buf = self.reserve(n)
        buf[:n].numpy()[:] = src
        return self.send(buf, n, device, into)


_STAGING = _Staging()
_PLAN_REGISTRY: Dict[int, "weakref.ReferenceType[TreeMetadata]"] = {}


_SM_COUNT: Dict[int, int] = {}


def sm_count(device: Optional[torch.device]) -> int:
    """CTAs the native plan is balanced for: one per SM of the device that will run it (148 on B200)."""
    if device is None or device.type != "cuda":
        return 148
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


class _NativeTree:
    """Handle + bookkeeping of one tree's C++ mirror (``deft_tree_t``); owned by ``TreeCache.native_tree``."""

    def __init__(self) -> None:
        self.handle = _lib.lib.deft_b200_tree_new()
        if not self.handle:
            raise MemoryError("deft_b200_tree_new failed")
        self.key = None                  # TreeCache._mirror_key() the mirror was last in step with
        self.lists: List[List[int]] = [] # the page lists of the walked nodes (their lengths are checked against n_pages)
        self.n_pages = -1
        self.leaf_idx = np.zeros(0, dtype=np.int32)   # DFS index of every leaf, leaves in ascending id order (-1: not walked)
        self.n_live = 0
        self.leaf_to_q: Dict[int, int] = {}
        self.syncs = 0                   # full hand-overs so far (a decode loop that only allocs needs one)

    def sync(self, tree, flat: Dict[str, Any]) -> None:
        """Full hand-over of the tree (``deft_b200_tree_set``) from the arrays ``flatten_tree`` made."""
        self.syncs += 1
        use_tix = bool(getattr(tree, "use_tree_index", False))
        _lib.check(_lib.lib.deft_b200_tree_set(self.handle, len(flat["parent"]), flat["parent"].ctypes.data,
                                               flat["kv_off"].ctypes.data, flat["kv"].ctypes.data, flat["q_off"].ctypes.data,
                                               flat["qs"].ctypes.data, flat["tix"].ctypes.data if use_tix else None,
                                               len(flat["leaf_to_q"])))
        nodes = tree._flat_topo[1][0]
        self.lists = [n.kv_indices for n in nodes]
        self.n_pages = int(flat["kv_off"][-1])
        at = {id(n): i for i, n in enumerate(nodes)}
        self.leaf_idx = np.asarray([at.get(id(tree.leaves[l]), -1) for l in sorted(tree.leaves)], dtype=np.int32)
        self.n_live = int((self.leaf_idx >= 0).sum())
        self.leaf_to_q = flat["leaf_to_q"]

    def append(self, pages: List[int]) -> None:
        """A decode step: one page per leaf, leaves in ascending id order (``TreeCache.alloc``)."""
        arr = np.asarray(pages, dtype=np.int64)
        assert arr.shape[0] == self.leaf_idx.shape[0]
        _lib.check(_lib.lib.deft_b200_tree_append(self.handle, arr.shape[0], self.leaf_idx.ctypes.data, arr.ctypes.data))
        self.n_pages += self.n_live

    def __del__(self) -> None:
        if getattr(self, "handle", None) and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.deft_b200_tree_free(self.handle)
            self.handle = None


def _native_trees_enabled() -> bool:
    return os.environ.get("DEFT_NATIVE_TREE", "1") != "0"


def mirror_flat(trees) -> Optional[Dict[str, Any]]:
    """The builder's input as native mirrors instead of flat arrays, when every tree keeps one (our ``TreeCache``)."""
    if not _native_trees_enabled() or not all(hasattr(t, "native_tree") for t in trees):
        return None
    mirrors = [t.native_tree() for t in trees]
    if len(trees) == 1:
        return dict(trees=mirrors, leaf_to_q=mirrors[0].leaf_to_q, query_num=len(mirrors[0].leaf_to_q))
    # (tree index, leaf id) -> query: kept while no tree of the forest changes its leaves
    key = tuple(id(m.leaf_to_q) for m in mirrors)
    cached = getattr(trees[0], "_forest_leaf_to_q", None)
    if cached is None or cached[0] != key:
        leaf_to_q: Dict[Any, int] = {}
        base = 0
        for i, m in enumerate(mirrors):
            for leaf, q in m.leaf_to_q.items():
                leaf_to_q[(i, leaf)] = q + base
            base += len(m.leaf_to_q)
        cached = trees[0]._forest_leaf_to_q = (key, leaf_to_q, base, [m.leaf_to_q for m in mirrors])   # (keeps the ids alive)
    return dict(trees=mirrors, leaf_to_q=cached[1], query_num=cached[2])


class TableLayout:
    """Capacity-padded packing of the tables (``deft_layout_t``): one per decode loop.  While every table fits its
    region the packed buffer keeps its offsets from one decode step to the next -- what a captured CUDA graph of the
    step needs -- although every step appends a page per leaf.  ``version`` moves when a region had to grow."""

    def __init__(self, native_only: bool = False) -> None:
        """``native_only``: builds with this layout make the native unit plan only -- the reference's twelve tables and
        the item / group plans of the warp-FMA path stay empty (a third of the build and two thirds of the upload).
        For decode loops that call the operators themselves on a tensor-core geometry (``DecodeStepGraph``)."""
        self.handle = _lib.lib.deft_b200_layout_new()
        if not self.handle:
            raise MemoryError("deft_b200_layout_new failed")
        self.native_only = bool(native_only)
        if native_only:
            _lib.lib.deft_b200_layout_set_native_only(self.handle, 1)

    @property
    def version(self) -> int:
        return int(_lib.lib.deft_b200_layout_version(self.handle))

    def __del__(self) -> None:
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None and getattr(_lib, "lib", None) is not None:     # (None at interpreter shutdown)
            _lib.lib.deft_b200_layout_free(h)


def build_tables_host(flat: Dict[str, Any], max_q_len: int = 32, max_block_len: int = -1,
                      block_len: int = 128, tree_index_max_ctx: int = 0, node_split: int = NODE_SPLIT,
                      hkv: int = 0, n_ctas: int = 148, reserve=None, layout: Optional[TableLayout] = None,
                      fresh_page: Optional[np.ndarray] = None):
    """Runs the C++ builder; returns (packed bytes as numpy uint8, directory, scalars).

    ``reserve(nbytes) -> uint8 tensor`` (optional) supplies the destination -- the pinned staging buffer of the
    upload -- so that the tables are copied once, straight out of the builder; the first return value is then
    ``(tensor, nbytes)`` instead of an array."""
    query_num = flat["query_num"] if "trees" in flat else len(flat["leaf_to_q"])
    use_tix = tree_index_max_ctx > 0
    if fresh_page is not None:
        fresh_page = np.ascontiguousarray(fresh_page, dtype=np.int32)
        assert fresh_page.shape == (query_num,), "fresh_page: one page per query"
    if "trees" in flat:       # native mirrors (TreeCache.native_tree): nothing of the tree crosses the ABI again
        mirrors = flat["trees"]
        handles = (C.c_void_p * len(mirrors))(*[m.handle for m in mirrors])
        h = _lib.lib.deft_b200_build_tables_trees(handles, len(mirrors), tree_index_max_ctx, block_len, max_q_len,
                                                  max_block_len, node_split, hkv, n_ctas,
                                                  layout.handle if layout is not None else None,
                                                  fresh_page.ctypes.data if fresh_page is not None else None)
    else:
        h = _lib.lib.deft_b200_build_tables(len(flat["parent"]), flat["parent"].ctypes.data, flat["kv_off"].ctypes.data,
                                            flat["kv"].ctypes.data, flat["q_off"].ctypes.data, flat["qs"].ctypes.data,
                                            flat["tix"].ctypes.data if use_tix else None,
                                            tree_index_max_ctx, query_num, block_len, max_q_len, max_block_len, node_split,
                                            hkv, n_ctas, layout.handle if layout is not None else None,
                                            fresh_page.ctypes.data if fresh_page is not None else None)
    if not h:
        raise _lib.DeftError(f"deft_b200_build_tables failed: {_lib.last_error()}")
    try:
        nbytes = _lib.lib.deft_b200_tables_bytes(h)
        src = _lib.lib.deft_b200_tables_data(h)
        if reserve is not None:
            buf = reserve(nbytes)
            C.memmove(buf.data_ptr(), src, nbytes)
            data = (buf, nbytes)
        else:
            data = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(src)).copy()
        meta = np.empty(2 * _lib.T_COUNT + _lib.N_SCALARS, dtype=np.int64)
        _lib.check(_lib.lib.deft_b200_tables_directory(h, meta.ctypes.data))
        _lib.check(_lib.lib.deft_b200_tables_scalars(h, meta.ctypes.data + 16 * _lib.T_COUNT))
    finally:
        _lib.lib.deft_b200_tables_free(h)
    return data, meta[: 2 * _lib.T_COUNT].reshape(-1, 2), meta[2 * _lib.T_COUNT:]


@dataclass
class TreeMetadata:
    """Field names and meanings of the reference dataclass (tree_cache.py:591-616)."""
    query_num: int
    node_num: int
    total_kv_len: int
    leaf_to_q: Dict[int, int]
    node_q: torch.Tensor
    node_kv: torch.Tensor
    node_q_len: torch.Tensor
    node_kv_len: torch.Tensor
    node_q_offset: torch.Tensor
    node_kv_offset: torch.Tensor
    block_len: int
    block_q: torch.Tensor
    block_q_cnts: torch.Tensor
    block_q_offset: torch.Tensor
    block_bitmasks: torch.Tensor
    block_kv: torch.Tensor
    block_lens: torch.Tensor
    # native extras (not in the reference): the packed device buffer and the two work plans
    packed: Optional[torch.Tensor] = field(default=None, repr=False)
    flat_plan: Optional[_lib.Plan] = field(default=None, repr=False)
    node_plan: Optional[_lib.Plan] = field(default=None, repr=False)
    host_tables: Optional[Dict[str, np.ndarray]] = field(default=None, repr=False)
    layout: bytes = field(default=b"", repr=False)     # directory + scalars: equal layouts = equal pointers and counts

    @classmethod
    def _assemble(cls, tree, flat, max_q_len: int, max_block_len: int, tree_index: bool,
                  device_buffer: Optional[torch.Tensor] = None, table_layout: Optional[TableLayout] = None,
                  fresh_page=None) -> "TreeMetadata":
        block_len = BLOCK_CONFIG["BLOCK_LEN"]
        if fresh_page is not None:
            fresh_page = (fresh_page.detach().cpu().numpy() if isinstance(fresh_page, torch.Tensor) else np.asarray(fresh_page)).astype(np.int32)
        if max_block_len == -1:
            max_block_len = BLOCK_CONFIG["MAX_BLOCK_LEN"]
        max_ctx = tree.tree_index_pool.node_to_kv.shape[1] if tree_index else 0
        pool = tree.token_to_kv_pool           # ours, or the reference's (which has no .device)
        device = getattr(pool, "device", None) or pool.kv_data[0].device
        hkv = int(pool.kv_data[0].shape[2]) if len(pool.kv_data) else 0     # kv_data[l] is [size, 2, HKV, D]
        on_gpu = device.type == "cuda"
        data, directory, scalars = build_tables_host(flat, max_q_len, max_block_len, block_len, max_ctx,
                                                     hkv=hkv, n_ctas=sm_count(device),
                                                     reserve=_STAGING.reserve if on_gpu else None, layout=table_layout,
                                                     fresh_page=fresh_page)
        if on_gpu:
            buf, nbytes = data
            if device_buffer is not None and device_buffer.numel() < nbytes:
                device_buffer = None           # too small: a fresh tensor, like without it
            packed = _STAGING.send(buf, nbytes, device, device_buffer)
        else:
            packed = torch.from_numpy(data)
        directory = directory.tolist()
        p64 = packed.view(torch.int64)         # every table starts on a 256-byte boundary of the packed buffer
        t = {name: p64[directory[i][0] >> 3: (directory[i][0] >> 3) + directory[i][1]]
             for i, name in enumerate(_lib.T_NAMES[:12])}
        base = packed.data_ptr()
        dir_bytes = np.asarray(directory, dtype=np.int64).tobytes()

        U = _lib.T_NAMES.index("u_units")

        def addr(i: int) -> Optional[int]:
            return base + directory[i][0] if directory[i][1] > 0 else None

        def plan(first: int, rows: int) -> _lib.Plan:
            """Item/group layer of one operator + the native unit layer (shared by both operators)."""
            return _lib.Plan(items=base + directory[first][0], groups=base + directory[first + 1][0],
                             csr_off=base + directory[first + 2][0], csr_rows=base + directory[first + 3][0],
                             n_items=directory[first][1], n_groups=directory[first + 1][1],
                             n_part_rows=rows, n_units=directory[U][1],
                             units=addr(U), u_csr_off=addr(U + 1), u_csr_rows=addr(U + 2), u_kv=addr(U + 3), u_blk=addr(U + 8),
                             u_mask=addr(U + 4), u_q=addr(U + 5), u_job_off=addr(U + 6), u_jobs=addr(U + 7),
                             n_unit_slots=int(scalars[9]), n_ctas=int(scalars[7]), hkv=hkv, paired=int(scalars[8]),
                             fresh=int(scalars[10]))

        if tree_index:
            null = torch.empty(0, dtype=torch.int64, device=device)
            tix = tree.tree_index_pool
            t["node_kv"] = (tix.device_table() if on_gpu and hasattr(tix, "device_table") else tix.node_to_kv).view(-1)
            for k in ("block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"):
                t[k] = null
        # what a captured graph of the step depends on: where the tables sit, and the few scalars the launches bake in
        # (grid, cluster pairing, workspace carving).  With a TableLayout the offsets are capacity-padded, and a step
        # that only appended pages keeps this key although every count has moved.
        if table_layout is not None:
            key = np.asarray([d[0] for d in directory] + [int(scalars[0]), int(scalars[7]), int(scalars[8]), int(scalars[9]), int(scalars[10]),
                                                          table_layout.version, base], dtype=np.int64).tobytes()
        else:
            key = dir_bytes + scalars.tobytes() + int(base).to_bytes(8, "little")
        meta = cls(query_num=int(scalars[0]), node_num=int(scalars[1]), total_kv_len=int(scalars[2]),
                   leaf_to_q=flat["leaf_to_q"], block_len=int(scalars[3]), packed=packed,
                   flat_plan=None if tree_index else plan(12, int(scalars[4])), node_plan=plan(16, int(scalars[5])),
                   layout=key, **t)
        if on_gpu:
            register_plan(meta)
        return meta

    @classmethod
    def from_tree_cache(cls, tree, tile_num: int = 8, max_q_len: int = 32, max_block_len: int = -1,
                        device_buffer: Optional[torch.Tensor] = None, table_layout: Optional[TableLayout] = None,
                        fresh_page=None) -> "TreeMetadata":
        """Reference signature (tree_cache.py:618-625) + ``device_buffer``: an optional persistent uint8 CUDA tensor
        the packed tables are uploaded into, so that consecutive decode steps find them at the same addresses
        (what a captured CUDA graph of the step needs, see ``decode_step.DecodeStepGraph``); ``table_layout``:
        capacity-padded packing; ``fresh_page``: this step's page per query (``TreeCache.alloc().cache_loc``) -- the
        native plan then reads those tokens from the step's activations (fused KV append: ``attention.Append``)."""
        return cls._assemble(tree, mirror_flat([tree]) or flatten_tree(tree), max_q_len, max_block_len, tree_index=False,
                             device_buffer=device_buffer, table_layout=table_layout, fresh_page=fresh_page)

    @classmethod
    def from_forest(cls, trees, max_q_len: int = 32, max_block_len: int = -1,
                    device_buffer: Optional[torch.Tensor] = None, table_layout: Optional[TableLayout] = None,
                    fresh_page=None) -> "TreeMetadata":
        """One metadata object (tables + native plan) for several trees sharing one KV pool: the operators
        then attend the whole batch in one launch.  ``leaf_to_q`` is keyed by ``(tree index, leaf id)``."""
        trees = list(trees)
        assert trees and all(t.token_to_kv_pool is trees[0].token_to_kv_pool for t in trees), \
            "the trees of a forest share one TokenToKVPool"
        return cls._assemble(trees[0], mirror_flat(trees) or flatten_forest(trees), max_q_len, max_block_len, tree_index=False,
                             device_buffer=device_buffer, table_layout=table_layout, fresh_page=fresh_page)

    @classmethod
    def from_tree_cache_node(cls, tree, tile_num: int = 8, max_q_len: int = 32, max_block_len: int = -1) -> "TreeMetadata":
        assert tree.use_tree_index and tree.tree_index_pool is not None
        return cls._assemble(tree, mirror_flat([tree]) or flatten_tree(tree), max_q_len, max_block_len, tree_index=True)


def register_plan(meta: TreeMetadata) -> None:
    """Lets the operator find the native plan from the table pointers it is called with."""
    keys = [meta.node_q.data_ptr()]
    if meta.flat_plan is not None:
        keys.append(meta.block_q.data_ptr())
    ref = weakref.ref(meta)
    for k in keys:
        _PLAN_REGISTRY[k] = ref
    weakref.finalize(meta, lambda ks=tuple(keys), r=ref: [_PLAN_REGISTRY.pop(k, None) for k in ks
                                                       if _PLAN_REGISTRY.get(k) is r])


def lookup_plan(table: torch.Tensor) -> Optional[TreeMetadata]:
    ref = _PLAN_REGISTRY.get(table.data_ptr())
    return ref() if ref is not None else None


GLOBAL_TREE_METADATA: Optional[TreeMetadata] = None
GLOBAL_TREE_CACHE: Optional[TreeCache] = None


def register_tree_metadata(tree_metadata: TreeMetadata) -> None:
    global GLOBAL_TREE_METADATA
    GLOBAL_TREE_METADATA = tree_metadata


def unregister_tree_metadata() -> None:
    global GLOBAL_TREE_METADATA
    GLOBAL_TREE_METADATA = None


def get_global_tree_metadata() -> TreeMetadata:
    assert GLOBAL_TREE_METADATA is not None
    return GLOBAL_TREE_METADATA


def register_tree_cache(tree_cache: TreeCache) -> None:
    global GLOBAL_TREE_CACHE
    GLOBAL_TREE_CACHE = tree_cache


def unregister_tree_cache() -> None:
    global GLOBAL_TREE_CACHE
    GLOBAL_TREE_CACHE = None


def get_global_tree_cache() -> TreeCache:
    assert GLOBAL_TREE_CACHE is not None
    return GLOBAL_TREE_CACHE
