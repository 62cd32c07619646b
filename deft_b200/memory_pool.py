"""Paged KV pool and per-sequence page table, host-managed.

Mirrors the interface of the reference's ``deft/memory_pool.py`` (``ReqToTokenPool`` :11-45,
``TokenToKVPool`` :48-108) for the tree-attention path: page = one token, ``kv_data[layer]`` is
``[size, 2(K/V), HKV, D]`` fp16 in HBM, ``alloc(n)`` returns the first ``n`` free pages in ascending
order.  Unlike the reference, the allocator state lives on the HOST (numpy), so allocation never
launches a kernel or synchronises the device; only the KV bytes live in HBM.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

IndexLike = Union[torch.Tensor, np.ndarray, List[int], int]


def _as_index(x: IndexLike) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().astype(np.int64).reshape(-1)
    return np.asarray(x, dtype=np.int64).reshape(-1)


class ReqToTokenPool:
    """Slot allocator + ``req_to_token [size, max_context_len] int32`` page table (memory_pool.py:11-45).

    The table is only read by the sequence-based (Radix) baseline; it is kept on the host and
    mirrored to the device on demand by :meth:`device_table`.
    """

    def __init__(self, size: int, max_context_len: int, device: Union[str, torch.device] = "cuda") -> None:
        self.device = torch.device(device)
        self.mem_state = np.ones(size, dtype=bool)
        self.can_use_mem_size = size
        self.req_to_token = torch.zeros((size, max_context_len), dtype=torch.int32)

    def alloc(self, need_size: int) -> Optional[torch.Tensor]:
        if need_size > self.can_use_mem_size:
            return None
        sel = np.flatnonzero(self.mem_state)[:need_size]
        self.mem_state[sel] = False
        self.can_use_mem_size -= need_size
        return torch.from_numpy(sel.astype(np.int32))

    def free(self, free_index: IndexLike) -> None:
        idx = _as_index(free_index)
        self.can_use_mem_size += len(idx)
        self.mem_state[idx] = True

    def copy(self, from_req: int, to_req: int, copy_len: int) -> None:
        self.req_to_token[to_req, :copy_len] = self.req_to_token[from_req, :copy_len]

    def clear(self) -> None:
        self.mem_state[:] = True
        self.can_use_mem_size = len(self.mem_state)

    def device_table(self) -> torch.Tensor:
        return self.req_to_token.to(self.device, non_blocking=True)


class TokenToKVPool:
    """Refcounted page allocator + the KV pages themselves (memory_pool.py:48-108)."""

    def __init__(self, size: int, dtype: torch.dtype, head_num: int, head_dim: int, layer_num: int,
                 device: Union[str, torch.device] = "cuda") -> None:
        self.device = torch.device(device)
        self.mem_state = np.zeros(size, dtype=np.int16)
        self.alloc_ct = 0
        self._low = 0          # no free page below this index (first-free search starts here instead of at page 0)
        # [size, key/value, head_num, head_dim] per layer -- the layout the kernels index
        self.kv_data = [torch.empty((size, 2, head_num, head_dim), dtype=dtype, device=self.device)
                        for _ in range(layer_num)]

    def get_key_buffer(self, layer_id: int) -> torch.Tensor:
        return self.kv_data[layer_id][:, 0]

    def get_value_buffer(self, layer_id: int) -> torch.Tensor:
        return self.kv_data[layer_id][:, 1]

    def alloc(self, need_size: int) -> Optional[torch.Tensor]:
        """The first ``need_size`` free pages in ascending order (memory_pool.py:74-80), found window by window from
        the lowest page that can be free -- the reference scans the whole pool with ``nonzero`` at every call."""
        n, lo, found, got = len(self.mem_state), self._low, [], 0
        win = max(4096, 4 * need_size)
        first_free = None
        while lo < n and got < need_size:
            idx = np.flatnonzero(self.mem_state[lo: lo + win] == 0)
            if len(idx):
                if first_free is None:
                    first_free = lo + int(idx[0])
                found.append(idx[: need_size - got] + lo)
                got += len(found[-1])
            lo += win
        self._low = first_free if first_free is not None else n
        if got < need_size:
            return None
        sel = found[0] if len(found) == 1 else np.concatenate(found)
        self.add_refs(sel)
        return torch.from_numpy(sel.astype(np.int32))

    def free(self, free_index: IndexLike) -> int:
        return self.decrease_refs(free_index)

    def used_size(self) -> int:
        return int(np.count_nonzero(self.mem_state))

    def available_size(self) -> int:
        return int(np.count_nonzero(self.mem_state == 0))

    def add_refs(self, token_index: IndexLike) -> None:
        idx = _as_index(token_index)
        self.alloc_ct += len(idx)
        # the reference's `mem_state[token_index] += 1` (memory_pool.py:92-94) is an indexed read-modify-write: a page
        # named twice in the list is incremented ONCE; plain fancy indexing has the same semantics (np.add.at has not)
        self.mem_state[idx] += 1

    def decrease_refs(self, token_index: IndexLike) -> int:
        idx = _as_index(token_index)
        self.alloc_ct -= len(idx)
        self.mem_state[idx] -= 1           # once per distinct page, like the reference (memory_pool.py:96-98)
        if len(idx):
            self._low = min(self._low, int(idx.min()))
        return int(np.count_nonzero(self.mem_state[idx] == 0))

    def clear(self) -> None:
        self.mem_state[:] = 0
        self.alloc_ct = 0
        self._low = 0


class TreeIndexPool:
    """Node -> page table of tree-index mode (tree_decoding/tree_index_pool.py:11-50), host-managed."""

    def __init__(self, size: int, max_context_len: int, device: Union[str, torch.device] = "cuda") -> None:
        self.device = torch.device(device)
        self.mem_state = np.ones(size, dtype=bool)
        self.can_use_mem_size = size
        self.node_to_kv = torch.zeros((size, max_context_len), dtype=torch.int32)
        self._dev: Optional[torch.Tensor] = None      # device mirror of node_to_kv ...
        self._dev_version = -1                        # ... as of this version of the host table
        self.version = 0                              # bumped by whoever writes node_to_kv (TreeCache does)

    def alloc(self, need_size: int) -> Optional[torch.Tensor]:
        if need_size > self.can_use_mem_size:
            return None
        sel = np.flatnonzero(self.mem_state)[:need_size]
        self.mem_state[sel] = False
        self.can_use_mem_size -= need_size
        return torch.from_numpy(sel.astype(np.int32))

    def free(self, free_index: IndexLike) -> None:
        idx = _as_index(free_index)
        self.can_use_mem_size += len(idx)
        self.mem_state[idx] = True

    def clear(self) -> None:
        self.mem_state[:] = True
        self.can_use_mem_size = len(self.mem_state)

    def get_offset(self, node_id: int) -> int:
        return node_id * self.node_to_kv.shape[1]

    def touch(self) -> None:
        """The host table has been written (rows are written through ``TreeNode.node_indices`` views)."""
        self.version += 1

    def device_table(self) -> torch.Tensor:
        """Persistent device mirror of ``node_to_kv`` (the reference keeps the table on the GPU); uploaded again only
        after the host table changed, through pinned memory so that the copy is asynchronous."""
        if self._dev is None or self._dev_version != self.version:
            if self._dev is None:
                self._dev = torch.empty(self.node_to_kv.shape, dtype=torch.int32, device=self.device)
                self._pinned = torch.empty(self.node_to_kv.shape, dtype=torch.int32).pin_memory() if self.device.type == "cuda" else None
            if self._pinned is not None:
                torch.cuda.current_stream(self.device).synchronize()      # (the previous upload has left the staging buffer)
                self._pinned.copy_(self.node_to_kv)
                self._dev.copy_(self._pinned, non_blocking=True)
            else:
                self._dev.copy_(self.node_to_kv)
            self._dev_version = self.version
        return self._dev
