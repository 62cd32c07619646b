"""A decode step's attention work as CUDA graphs.

Per layer a decode step does ``KVCacheUpdater.update`` (append this step's K/V rows, tree_cache.py:67-76) and one
tree-attention call (deft_attention.py:110-151 / 72-108).  From Python that is three launches and ~45 us of host
time per layer -- more than the kernels take -- so the step is captured ONCE into CUDA graphs (one per chunk of
layers, so that host<->device copies of the neighbouring chunks can overlap) and replayed for as long as the table
LAYOUT (sizes, offsets, counts, addresses: ``TreeMetadata.layout``) is the same; the table CONTENTS (page ids, masks,
job lists) are free to change from step to step because they live in one persistent device buffer the graphs point
into.  A different layout re-captures.  This is SURVEY.md 8(f) item 4.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import attention
from .tree_cache import TreeMetadata


class DecodeStepGraph:
    def __init__(self, kv_pool, qkv: torch.Tensor, out: torch.Tensor, cache_loc: torch.Tensor, num_heads: int,
                 num_kv_heads: int, head_dim: int, mode: str = "flatten", chunk: int = 8,
                 table_bytes: int = 8 << 20) -> None:
        """``qkv``: [layers, nq, (H + 2 HKV) D] fp16 device buffer the fused projections land in; ``out``: [layers, nq,
        H, D]; ``cache_loc``: [nq] int32 device buffer with this step's page per query (all three keep their
        addresses; their contents change every step).  ``mode``: flatten | node | node_chunk."""
        assert mode in ("flatten", "node", "node_chunk")
        self.kv_pool, self.qkv, self.out, self.loc = kv_pool, qkv, out, cache_loc
        self.H, self.HKV, self.D, self.mode = num_heads, num_kv_heads, head_dim, mode
        self.layers = qkv.shape[0]
        self.chunk = max(1, min(chunk, self.layers))
        self.n_chunks = (self.layers + self.chunk - 1) // self.chunk
        self.tables = torch.empty(table_bytes, dtype=torch.uint8, device=qkv.device)
        self._graphs: Dict[bytes, List[torch.cuda.CUDAGraph]] = {}
        self.captures = 0

    # ---- tables -------------------------------------------------------------------------------
    def metadata(self, trees) -> TreeMetadata:
        """C++ builder + ONE async upload into the persistent table buffer (grown, and the graphs dropped, if the
        tables outgrow it)."""
        single = not isinstance(trees, (list, tuple))
        for _ in range(2):
            m = (TreeMetadata.from_tree_cache(trees, device_buffer=self.tables) if single
                 else TreeMetadata.from_forest(trees, device_buffer=self.tables))
            if m.packed.data_ptr() == self.tables.data_ptr():
                return m
            self.tables = torch.empty(2 * m.packed.numel(), dtype=torch.uint8, device=self.qkv.device)
            self._graphs.clear()
        return m

    # ---- the work of one layer ----------------------------------------------------------------
    def _layer(self, l: int, m: TreeMetadata) -> None:
        H, HKV, D = self.H, self.HKV, self.D
        nq = self.qkv.shape[1]
        row = self.qkv[l]
        attention.kv_append(self.kv_pool.kv_data[l], row[:, H * D: (H + HKV) * D].view(nq, HKV, D),
                            row[:, (H + HKV) * D:].view(nq, HKV, D), self.loc)
        q = row[:, : H * D].view(nq, H, D)
        K, V = self.kv_pool.get_key_buffer(l), self.kv_pool.get_value_buffer(l)
        if self.mode == "flatten":
            attention.tree_attention_subtree_fwd(q, K, V, self.out[l], m.block_len, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
        else:
            attention.tree_attention_fwd(q, K, V, self.out[l], m.node_kv, m.node_kv_offset, m.node_kv_len, m.node_q,
                                         m.node_q_offset, m.node_q_len)

    def _capture(self, m: TreeMetadata) -> List[torch.cuda.CUDAGraph]:
        for l in range(self.layers):          # eager once: sizes the workspace outside the capture
            self._layer(l, m)
        torch.cuda.current_stream().synchronize()
        graphs = []
        for c in range(self.n_chunks):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for l in range(c * self.chunk, min(self.layers, (c + 1) * self.chunk)):
                    self._layer(l, m)
            graphs.append(g)
        self.captures += 1
        return graphs

    # ---- one step -------------------------------------------------------------------------------
    def run(self, m: TreeMetadata, before_chunk: Optional[Callable[[int], None]] = None,
            after_chunk: Optional[Callable[[int], None]] = None) -> None:
        """Replays the step on the current stream.  ``before_chunk(c)`` / ``after_chunk(c)`` run on the host right
        before / after chunk ``c`` is enqueued (event waits for the chunk's inputs, event records for its outputs)."""
        assert m.packed is not None and m.packed.data_ptr() == self.tables.data_ptr(), \
            "build the step's tables with DecodeStepGraph.metadata()"
        graphs = self._graphs.get(m.layout)
        if graphs is None:
            graphs = self._graphs[m.layout] = self._capture(m)
        for c, g in enumerate(graphs):
            if before_chunk is not None:
                before_chunk(c)
            g.replay()
            if after_chunk is not None:
                after_chunk(c)
