"""A decode step's attention work as CUDA graphs that survive the tree growing.

Per layer a decode step does ``KVCacheUpdater.update`` (append this step's K/V rows, tree_cache.py:67-76) and one
tree-attention call (deft_attention.py:110-151 / 72-108).  From Python that is three launches and ~45 us of host
time per layer -- more than the kernels take -- so the step is captured ONCE into CUDA graphs (one per chunk of
layers, so that host<->device copies of the neighbouring chunks can overlap) and replayed.

What a captured launch holds is device ADDRESSES (tables, workspace, activations) and a few scalars (grid, cluster
pairing); every count the kernels need they read from the tables themselves (job records, CSR bounds).  A real
decode loop appends a page per leaf every step (``TreeCache.alloc``, tree_generate.py:109), so every table grows a
little: the step therefore builds its tables with a ``TableLayout`` (capacity-padded regions in ONE persistent device
buffer, ``deft_layout_t``) -- the offsets only move when a table outgrows its region (the small native tables take
100 % of headroom), and only then is the step captured again.  The workspace is owned by the step and sized for the layout's slot capacity.
This is SURVEY.md 8(f) items 1 (tables that follow the tree incrementally) and 4 (graph'd decode step).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Callable, List, Optional

import torch

from . import _lib, attention
from .tree_cache import BLOCK_CONFIG, TableLayout, TreeMetadata


class DecodeStepGraph:
    MAX_LAYOUTS = 4      # captured layouts kept (least recently used goes first)

    def __init__(self, kv_pool, qkv: torch.Tensor, out: torch.Tensor, cache_loc: torch.Tensor, num_heads: int,
                 num_kv_heads: int, head_dim: int, mode: str = "flatten", chunk=8,
                 table_bytes: int = 8 << 20, fused_append: bool = True, reference_tables: bool = False) -> None:
        """``qkv``: [layers, nq, (H + 2 HKV) D] fp16 device buffer the fused projections land in; ``out``: [layers, nq,
        H, D]; ``cache_loc``: [nq] int32 device buffer with this step's page per query (all three keep their
        addresses; their contents change every step).  ``mode``: flatten | node | node_chunk.  ``chunk``: layers per CUDA
        graph, or the list of chunk sizes.  ``fused_append``: the
        step's K/V rows are read by the attention straight from ``qkv`` and written to their pages by its second
        kernel (``metadata(trees, cache_loc=...)`` marks them), instead of a ``kv_append`` launch per layer.
        ``reference_tables``: also build and upload the reference's int64 tables (``block_q``, ``node_kv`` ... of the
        ``TreeMetadata`` that ``metadata()`` returns); the step itself only reads the native plan, so by default -- on
        the geometries of the tensor-core kernels -- they are left empty."""
        assert mode in ("flatten", "node", "node_chunk")
        self.kv_pool, self.qkv, self.out, self.loc = kv_pool, qkv, out, cache_loc
        self.H, self.HKV, self.D, self.mode = num_heads, num_kv_heads, head_dim, mode
        self.layers = qkv.shape[0]
        # (the fused append lives in the tensor-core kernels: other geometries append with kv_append launches)
        self.fused_append = fused_append and head_dim in (64, 128) and num_heads // num_kv_heads in (1, 2, 4)
        if isinstance(chunk, (list, tuple)):          # chunk sizes given one by one (e.g. a short first and last chunk)
            sizes = [int(c) for c in chunk]
            assert sizes and min(sizes) >= 1 and sum(sizes) == self.layers, "chunk sizes must add up to the layers"
        else:
            size = max(1, min(int(chunk), self.layers))
            sizes = [min(size, self.layers - l) for l in range(0, self.layers, size)]
        self.bounds = [0]
        for c in sizes:
            self.bounds.append(self.bounds[-1] + c)
        self.n_chunks = len(sizes)
        self.tables = torch.empty(table_bytes, dtype=torch.uint8, device=qkv.device)
        tensor_core = head_dim in (64, 128) and num_heads // num_kv_heads in (1, 2, 4)
        self.table_layout = TableLayout(native_only=tensor_core and not reference_tables)
        self.workspace = torch.empty(1 << 20, dtype=torch.uint8, device=qkv.device)
        self._graphs: "OrderedDict[bytes, List[torch.cuda.CUDAGraph]]" = OrderedDict()
        self._capture_stream: Optional[torch.cuda.Stream] = None
        self.captures = 0

    # ---- tables -------------------------------------------------------------------------------
    def metadata(self, trees, cache_loc=None) -> TreeMetadata:
        """C++ builder + ONE async upload into the persistent table buffer (grown, and the graphs dropped, if the
        tables outgrow it).  ``cache_loc``: this step's page per query on the HOST (``TreeCache.alloc().cache_loc``, the
        trees' one after the other) -- needed for the fused append; without it the step appends with ``kv_append``."""
        single = not isinstance(trees, (list, tuple))
        fresh = cache_loc if self.fused_append else None
        if self.mode == "node_chunk":
            BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
        try:
            for _ in range(2):
                m = (TreeMetadata.from_tree_cache(trees, device_buffer=self.tables, table_layout=self.table_layout, fresh_page=fresh)
                     if single else
                     TreeMetadata.from_forest(trees, device_buffer=self.tables, table_layout=self.table_layout, fresh_page=fresh))
                if m.packed.data_ptr() == self.tables.data_ptr():
                    return m
                self.tables = torch.empty(2 * m.packed.numel(), dtype=torch.uint8, device=self.qkv.device)
                self._graphs.clear()
            return m
        finally:
            if self.mode == "node_chunk":
                BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1

    def _plan(self, m: TreeMetadata):
        return m.flat_plan if self.mode == "flatten" else m.node_plan

    def _size_workspace(self, m: TreeMetadata) -> None:
        """The step's own partial-softmax buffer, sized for the layout's slot CAPACITY (a graph keeps its address)."""
        nq = self.qkv.shape[1]
        plan = self._plan(m)
        if self.mode == "flatten":
            need = _lib.lib.deft_b200_flatten_workspace_bytes(nq, self.H, self.HKV, self.D, m.block_q.numel(),
                                                              m.block_q_cnts.numel(), C.byref(plan))
        else:
            need = _lib.lib.deft_b200_node_workspace_bytes(nq, self.H, self.HKV, self.D, m.node_q.numel(),
                                                           m.node_kv_offset.numel(), m.node_kv.numel(), C.byref(plan))
        if self.workspace.numel() < need:
            self.workspace = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=self.qkv.device)
            self._graphs.clear()          # (the captured launches point into the old buffer)

    # ---- the work of one layer ----------------------------------------------------------------
    def _layer(self, l: int, m: TreeMetadata) -> None:
        H, HKV, D = self.H, self.HKV, self.D
        nq = self.qkv.shape[1]
        row = self.qkv[l]
        k_new, v_new = row[:, H * D: (H + HKV) * D].view(nq, HKV, D), row[:, (H + HKV) * D:].view(nq, HKV, D)
        append = (k_new, v_new, self.loc) if self._plan(m).fresh else None
        if append is None:
            attention.kv_append(self.kv_pool.kv_data[l], k_new, v_new, self.loc)
        q = row[:, : H * D].view(nq, H, D)
        K, V = self.kv_pool.get_key_buffer(l), self.kv_pool.get_value_buffer(l)
        if self.mode == "flatten":
            attention.tree_attention_subtree_fwd(q, K, V, self.out[l], m.block_len, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens,
                                                 plan=m.flat_plan, workspace=self.workspace, append=append)
        else:
            attention.tree_attention_fwd(q, K, V, self.out[l], m.node_kv, m.node_kv_offset, m.node_kv_len, m.node_q,
                                         m.node_q_offset, m.node_q_len, plan=m.node_plan, workspace=self.workspace, append=append)

    def _capture(self, m: TreeMetadata) -> List[torch.cuda.CUDAGraph]:
        """The step's launches into one CUDA graph per chunk of layers.  A re-capture happens in the middle of a decode
        loop (a table outgrew its region), so it must not stall it: ``capture_begin`` / ``capture_end`` on a side
        stream, without the device-wide synchronize + ``empty_cache`` that ``torch.cuda.graph`` puts in front of every
        capture (measured: 3-20 ms per capture, a 0.8 s stall once) -- the launches allocate nothing."""
        cur = torch.cuda.current_stream()
        if self.captures == 0:
            self._layer(0, m)                 # eager once: launch attributes, tensor maps and errors outside the capture
            cur.synchronize()
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=self.qkv.device)
        side = self._capture_stream
        side.wait_stream(cur)
        graphs = []
        with torch.cuda.stream(side):
            for c in range(self.n_chunks):
                g = torch.cuda.CUDAGraph()
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    for l in range(self.bounds[c], self.bounds[c + 1]):
                        self._layer(l, m)
                finally:
                    g.capture_end()
                graphs.append(g)
        cur.wait_stream(side)
        self.captures += 1
        return graphs

    # ---- one step -------------------------------------------------------------------------------
    def run(self, m: TreeMetadata, before_chunk: Optional[Callable[[int], None]] = None,
            after_chunk: Optional[Callable[[int], None]] = None) -> None:
        """Replays the step on the current stream.  ``before_chunk(c)`` / ``after_chunk(c)`` run on the host right
        before / after chunk ``c`` is enqueued (event waits for the chunk's inputs, event records for its outputs)."""
        assert m.packed is not None and m.packed.data_ptr() == self.tables.data_ptr(), \
            "build the step's tables with DecodeStepGraph.metadata()"
        self._size_workspace(m)
        key = m.layout + self.workspace.data_ptr().to_bytes(8, "little")
        graphs = self._graphs.get(key)
        if graphs is None:
            graphs = self._graphs[key] = self._capture(m)
            while len(self._graphs) > self.MAX_LAYOUTS:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        for c, g in enumerate(graphs):
            if before_chunk is not None:
                before_chunk(c)
            g.replay()
            if after_chunk is not None:
                after_chunk(c)


class DecodeStepPipeline:
    """Two ``DecodeStepGraph`` used alternately, so that the tables of step t+1 are built and uploaded while the layers
    of step t run on the GPU.

    What the tables of a decode step depend on is the TREE (topology and pages), not the token values the previous
    step sampled: ``TreeCache.alloc`` hands every leaf a page whatever its token turns out to be (tree_cache.py:401-444).
    So as long as the branch controller leaves the topology alone (the common step), the host work of step t+1
    (alloc, flatten_tree, C++ builder, one H2D copy) can run under the GPU time of step t.  Each half owns its table
    buffer, layout, ``cache_loc`` buffer and workspace; they share the activation / output buffers and the KV pool
    (the steps themselves stay serial on the stream).  A step whose topology changed after ``prepare`` simply calls
    ``prepare`` again before ``run`` -- that step then pays the table build in line, like the unpipelined step."""

    def __init__(self, kv_pool, qkv: torch.Tensor, out: torch.Tensor, num_heads: int, num_kv_heads: int,
                 head_dim: int, **kw) -> None:
        nq = qkv.shape[1]
        self.halves = [DecodeStepGraph(kv_pool, qkv, out, torch.zeros(nq, dtype=torch.int32, device=qkv.device),
                                       num_heads, num_kv_heads, head_dim, **kw) for _ in range(2)]
        self._turn = 0
        self._ready = None

    @property
    def captures(self) -> int:
        return sum(h.captures for h in self.halves)

    def prepare(self, trees, cache_loc: Optional[torch.Tensor] = None, fused: bool = True) -> TreeMetadata:
        """Host side of the NEXT step: tables + their upload (and this step's pages, ``cache_loc`` on the host) into the
        half that the running step does not read.  Enqueued on the current stream, i.e. behind the running step.
        ``fused=False``: the step appends with ``kv_append`` launches (the tables do not mark the fresh tokens)."""
        half = self.halves[self._turn]
        m = half.metadata(trees, cache_loc=cache_loc if fused else None)
        if cache_loc is not None:
            half.loc.copy_(cache_loc, non_blocking=True)
        self._ready = (half, m)
        return m

    def run(self, before_chunk: Optional[Callable[[int], None]] = None,
            after_chunk: Optional[Callable[[int], None]] = None) -> TreeMetadata:
        """Replays the prepared step and hands the turn to the other half."""
        assert self._ready is not None, "DecodeStepPipeline.prepare() first"
        half, m = self._ready
        self._ready = None
        self._turn ^= 1
        half.run(m, before_chunk=before_chunk, after_chunk=after_chunk)
        return m
