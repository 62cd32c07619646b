"""The tree-attention operators, with the reference's Python signatures.

``tree_attention_fwd`` (DeFT-Node / Node-Chunk / Tree-Index) and ``tree_attention_subtree_fwd``
(DeFT-Flatten) keep the positional signatures of ``deft/layers/attention/tree_attention.py:14-25``
and ``:552-568`` so that ``DeFTAttention.deft_node_forward`` / ``deft_flatten_forward``
(``deft_attention.py:94-105, 136-148``) can call them unchanged.  Each call is a thin shim over the
C ABI (``deft_b200_node_fwd`` / ``deft_b200_flatten_fwd``): it passes raw device pointers, strides
and the current CUDA stream.  There is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import Plan

_SUPPORTED_HEAD_DIMS = (16, 32, 64, 128)     # the reference's own list (tree_attention.py:100,305,582)

try:                                      # raw handle of the current stream without building a Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                    # pragma: no cover - older / newer torch: public (slower) API
    def _raw_stream(index: int) -> int:
        return torch.cuda.current_stream(index).cuda_stream


def _current_stream(t: torch.Tensor) -> int:
    return _raw_stream(t.device.index if t.device.index is not None else torch.cuda.current_device())


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on_device(t: torch.Tensor):
    """The C ABI launches on the current device: switch to the tensors' device around a call made from another one
    (a process driving several GPUs); nothing in the common single-device case."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)
_WORKSPACES: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(device: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch for the partial-softmax buffers; grows, never shrinks.

    Calls on one stream are ordered, so reusing the buffer between calls is race-free (the
    reference allocates and zero-fills its partial buffers on every call).
    """
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _check_qkvo(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor) -> Tuple[int, int, int, int]:
    if not (q.is_cuda and k.is_cuda and v.is_cuda and o.is_cuda):
        raise _lib.DeftError("deft_b200 tree attention needs CUDA tensors (there is no CPU fallback)")
    assert q.dtype == torch.float16 and k.dtype == torch.float16 and v.dtype == torch.float16 and o.dtype == torch.float16
    nq, H, D = q.shape
    HKV = k.shape[1]
    assert D in _SUPPORTED_HEAD_DIMS, f"head_dim {D} not in {_SUPPORTED_HEAD_DIMS}"
    assert k.shape[2] == D and v.shape == k.shape and tuple(o.shape) == (nq, H, D) and H % HKV == 0
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and o.stride(2) == 1
    assert k.stride() == v.stride(), "K and V views must share one layout (kv_data[layer][:, 0] / [:, 1])"
    return nq, H, HKV, D


def _append_struct(append, nq: int, HKV: int, D: int) -> "_lib.Append":
    new_k, new_v, loc = append
    assert new_k.is_cuda and new_v.is_cuda and loc.is_cuda and loc.dtype == torch.int32 and loc.is_contiguous()
    assert new_k.dtype == torch.float16 and tuple(new_k.shape) == (nq, HKV, D) == tuple(new_v.shape) and loc.numel() == nq
    assert new_k.stride() == new_v.stride() and new_k.stride(2) == 1
    return _lib.Append(new_k.data_ptr(), new_v.data_ptr(), new_k.stride(0), new_k.stride(1), loc.data_ptr())


def _i64(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.int64 and t.is_cuda
    return t if t.is_contiguous() else t.contiguous()


def _flat_tables(block_q, block_q_cnts, block_q_offset, block_bitmasks, block_kv, block_lens, plan):
    """Pointers and counts of the Flatten tables + the plan to use; cached on the TreeMetadata that owns them
    (the tables of a decode step are immutable and shared by the 32 layer-calls of the step)."""
    from .tree_cache import lookup_plan
    meta = lookup_plan(block_q)
    owned = meta is not None and meta.flat_plan is not None and meta.block_kv.data_ptr() == block_kv.data_ptr()
    if owned and plan is None:
        cached = meta.__dict__.get("_flat_call")
        if cached is not None:
            return cached
    bq, bc, bo = _i64(block_q), _i64(block_q_cnts), _i64(block_q_offset)
    bm, bk, bl = _i64(block_bitmasks), _i64(block_kv), _i64(block_lens)
    use = plan if plan is not None else (meta.flat_plan if owned else None)
    args = (bq.data_ptr(), bq.numel(), bc.data_ptr(), bo.data_ptr(), bl.data_ptr(), bc.numel(), bm.data_ptr(),
            bk.data_ptr(), C.byref(use) if use is not None else None, {}, (bq, bc, bo, bm, bk, bl, use))
    if owned and plan is None:
        meta.__dict__["_flat_call"] = args
    return args


def tree_attention_subtree_fwd(query_states: torch.Tensor, key_buffer: torch.Tensor, value_buffer: torch.Tensor,
                               output: torch.Tensor, block_len: int, block_q: torch.Tensor,
                               block_q_cnts: torch.Tensor, block_q_offset: torch.Tensor,
                               block_bitmasks: torch.Tensor, block_kv: torch.Tensor, block_lens: torch.Tensor,
                               plan: Optional[Plan] = None, workspace: Optional[torch.Tensor] = None,
                               append: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None) -> None:
    """DeFT-Flatten attention; writes ``output`` in place (tree_attention.py:552-667).  ``workspace``: an optional
    caller-owned uint8 CUDA tensor for the partial-softmax buffers (a captured CUDA graph keeps its address); by
    default a per-(device, stream) buffer that grows on demand.  ``append`` = ``(new_k, new_v, cache_loc)``: fused KV
    append -- the tables must have been built with ``fresh_page``; this step's tokens are read from ``new_k / new_v``
    ([nq, HKV, D] views of the fused qkv output) and written to the pool pages ``cache_loc`` (int32, device)."""
    nq, H, HKV, D = _check_qkvo(query_states, key_buffer, value_buffer, output)
    (p_bq, n_partials, p_bc, p_bo, p_bl, n_blocks, p_bm, p_bk, plan_ref, ws_need, _keep) = _flat_tables(
        block_q, block_q_cnts, block_q_offset, block_bitmasks, block_kv, block_lens, plan)
    stream = _current_stream(query_states)
    geom = (nq, H, HKV, D)
    need = ws_need.get(geom)
    if need is None:
        need = ws_need[geom] = _lib.lib.deft_b200_flatten_workspace_bytes(nq, H, HKV, D, n_partials, n_blocks, plan_ref)
    ws = workspace if workspace is not None else _workspace(query_states.device, stream, need)
    qs, ks, os_ = query_states.stride(), key_buffer.stride(), output.stride()
    with _on_device(query_states):
        if append is None:
            rc = _lib.lib.deft_b200_flatten_fwd(
                query_states.data_ptr(), qs[0], qs[1], key_buffer.data_ptr(), value_buffer.data_ptr(), ks[0], ks[1],
                key_buffer.shape[0], output.data_ptr(), os_[0], os_[1], nq, H, HKV, D, int(block_len),
                p_bq, n_partials, p_bc, p_bo, p_bl, n_blocks, p_bm, p_bk, plan_ref, ws.data_ptr(), ws.numel(), stream)
        else:
            rc = _lib.lib.deft_b200_flatten_fwd_append(
                query_states.data_ptr(), qs[0], qs[1], key_buffer.data_ptr(), value_buffer.data_ptr(), ks[0], ks[1],
                key_buffer.shape[0], output.data_ptr(), os_[0], os_[1], nq, H, HKV, D, int(block_len),
                p_bq, n_partials, p_bc, p_bo, p_bl, n_blocks, p_bm, p_bk, plan_ref, C.byref(_append_struct(append, nq, HKV, D)),
                ws.data_ptr(), ws.numel(), stream)
    if rc:
        _lib.check(rc)


def tree_attention_fwd(query_states: torch.Tensor, key_buffer: torch.Tensor, value_buffer: torch.Tensor,
                       output: torch.Tensor, KV_indices: torch.Tensor, KV_indices_offset: torch.Tensor,
                       KV_len: torch.Tensor, KVMapQ_List: torch.Tensor, KVMapQ_List_Offset: torch.Tensor,
                       KVMapQ_List_Len: torch.Tensor, plan: Optional[Plan] = None,
                       workspace: Optional[torch.Tensor] = None,
                       append: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None) -> None:
    """DeFT-Node / Node-Chunk / Tree-Index attention; writes ``output`` in place (tree_attention.py:14-68)."""
    from .tree_cache import lookup_plan
    nq, H, HKV, D = _check_qkvo(query_states, key_buffer, value_buffer, output)
    assert KV_indices.is_cuda and KV_indices.dtype in (torch.int64, torch.int32)
    KV_indices = KV_indices if KV_indices.is_contiguous() else KV_indices.contiguous()
    kv_off, kv_len = _i64(KV_indices_offset), _i64(KV_len)
    node_q, q_off, q_len = _i64(KVMapQ_List), _i64(KVMapQ_List_Offset), _i64(KVMapQ_List_Len)
    n_partials, n_entries = node_q.numel(), kv_off.numel()
    # int64 node_kv holds exactly sum(kv_len) pages, which bounds the split of long entries; the int32
    # tree-index table gives no such bound, and its entries are <= 128 tokens anyway -> no split
    total_kv_bound = KV_indices.numel() if KV_indices.dtype == torch.int64 else 0
    if plan is None:
        meta = lookup_plan(node_q)
        if meta is not None and meta.node_plan is not None and meta.node_kv.data_ptr() == KV_indices.data_ptr():
            plan = meta.node_plan
    stream = _current_stream(query_states)
    need = _lib.lib.deft_b200_node_workspace_bytes(nq, H, HKV, D, n_partials, n_entries, total_kv_bound,
                                                   C.byref(plan) if plan is not None else None)
    ws = workspace if workspace is not None else _workspace(query_states.device, stream, need)
    args = (query_states.data_ptr(), query_states.stride(0), query_states.stride(1),
            key_buffer.data_ptr(), value_buffer.data_ptr(), key_buffer.stride(0), key_buffer.stride(1), key_buffer.shape[0],
            output.data_ptr(), output.stride(0), output.stride(1), nq, H, HKV, D,
            KV_indices.data_ptr(), KV_indices.element_size(), kv_off.data_ptr(), kv_len.data_ptr(), node_q.data_ptr(),
            n_partials, q_off.data_ptr(), q_len.data_ptr(), n_entries, total_kv_bound,
            C.byref(plan) if plan is not None else None)
    with _on_device(query_states):
        if append is None:
            rc = _lib.lib.deft_b200_node_fwd(*args, ws.data_ptr(), ws.numel(), stream)
        else:
            rc = _lib.lib.deft_b200_node_fwd_append(*args, C.byref(_append_struct(append, nq, HKV, D)), ws.data_ptr(), ws.numel(), stream)
    _lib.check(rc)


_SEQ_CONST: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}


def token_attention_fwd(q: torch.Tensor, k_buffer: torch.Tensor, v_buffer: torch.Tensor, o: torch.Tensor,
                        req_to_token: torch.Tensor, b_req_idx: torch.Tensor, b_start_loc: torch.Tensor,
                        b_seq_len: torch.Tensor, max_len_in_batch: int, other_kv_index=None,
                        total_num_tokens=None, att_m=None) -> None:
    """Sequence-based decode attention (Radix Attention / ``--mode seq --mem paged``), reference signature of
    ``deft/layers/attention/token_attention.py:297-335``: query ``i`` attends the first ``b_seq_len[i]`` pages of
    row ``b_req_idx[i]`` of ``req_to_token``, independently of every other query (no prefix sharing).

    Runs on the same sm_100a kernels as the tree operators: every query is one entry of a Node-style table whose
    KV list is its row of the page table (int32, read in place); the tables are a few device-side index ops, no
    host synchronisation.  ``b_start_loc``, ``other_kv_index``, ``total_num_tokens`` and ``att_m`` (the
    reference's logits scratch) are accepted and unused.
    """
    nq, H, HKV, D = _check_qkvo(q, k_buffer, v_buffer, o)
    assert req_to_token.is_cuda and req_to_token.dtype == torch.int32 and req_to_token.stride(1) == 1
    assert b_req_idx.numel() == nq and b_seq_len.numel() == nq
    dev_idx = q.device.index if q.device.index is not None else torch.cuda.current_device()
    const = _SEQ_CONST.get((nq, dev_idx))
    if const is None:
        const = _SEQ_CONST[(nq, dev_idx)] = (torch.arange(nq, dtype=torch.int64, device=q.device),
                                             torch.ones(nq, dtype=torch.int64, device=q.device))
    ids, ones = const
    kv_off = b_req_idx.to(torch.int64) * req_to_token.stride(0)
    kv_len = b_seq_len.to(torch.int64)
    total_kv_bound = nq * int(max_len_in_batch)       # long sequences are cut into 256-token items on the device
    stream = _current_stream(q)
    need = _lib.lib.deft_b200_node_workspace_bytes(nq, H, HKV, D, nq, nq, total_kv_bound, None)
    ws = _workspace(q.device, stream, need)
    with _on_device(q):
        rc = _lib.lib.deft_b200_node_fwd(
            q.data_ptr(), q.stride(0), q.stride(1), k_buffer.data_ptr(), v_buffer.data_ptr(), k_buffer.stride(0),
            k_buffer.stride(1), k_buffer.shape[0], o.data_ptr(), o.stride(0), o.stride(1), nq, H, HKV, D,
            req_to_token.data_ptr(), 4, kv_off.data_ptr(), kv_len.data_ptr(), ids.data_ptr(), nq, ids.data_ptr(),
            ones.data_ptr(), nq, total_kv_bound, None, ws.data_ptr(), ws.numel(), stream)
    _lib.check(rc)


def kv_append(kv_layer: torch.Tensor, cache_k: torch.Tensor, cache_v: torch.Tensor, cache_loc: torch.Tensor) -> None:
    """``key_buffer[cache_loc] = cache_k; value_buffer[cache_loc] = cache_v`` in one launch.

    ``kv_layer`` is ``TokenToKVPool.kv_data[layer]`` ``[size, 2, HKV, D]`` (tree_cache.py:67-76).
    """
    if not (kv_layer.is_cuda and cache_k.is_cuda and cache_v.is_cuda and cache_loc.is_cuda):
        raise _lib.DeftError("kv_append needs CUDA tensors (there is no CPU fallback)")
    assert cache_loc.dtype == torch.int32 and cache_loc.is_contiguous()
    assert kv_layer.dtype == torch.float16 and cache_k.dtype == torch.float16 and cache_v.dtype == torch.float16
    n, HKV, D = cache_k.shape
    ks, ls = cache_k.stride(), kv_layer.stride()          # K view = kv_layer[:, 0], V view = kv_layer[:, 1]
    assert cache_v.shape == cache_k.shape and ks == cache_v.stride() and ks[2] == 1 and ls[3] == 1
    assert cache_loc.numel() == n and kv_layer.shape[1] == 2 and kv_layer.shape[2] == HKV and kv_layer.shape[3] == D
    stream = _current_stream(kv_layer)
    k_ptr = kv_layer.data_ptr()
    with _on_device(kv_layer):
        rc = _lib.lib.deft_b200_kv_append(k_ptr, k_ptr + ls[1] * 2, ls[0], ls[2], cache_k.data_ptr(), cache_v.data_ptr(),
                                          ks[0], ks[1], cache_loc.data_ptr(), n, HKV, D, stream)
    if rc:
        _lib.check(rc)
