"""CPU oracle for the DeFT tree-attention decode path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference (LINs-lab/DeFT @ 728525ec) for the
one path this repository accelerates.  Nothing under ``deft_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do,
and only as the checker / reported CPU baseline, never as the product path.

Parity pinning: the reference ships no golden vectors for this path, so the oracle
is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/gen_golden.py`` imports the reference
from ``/root/reference`` (Triton kernels under ``TRITON_INTERPRET=1``, ``TreeCache`` /
``TreeMetadata`` on CPU) and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function here against those files.

Reference lines restated (paths under ``/root/reference/DeFT/deft``):

* ``build_tables``           -> ``tree_decoding/tree_cache.py:618-881`` (``TreeMetadata.from_tree_cache``)
* ``build_tables_tree_index``-> ``tree_decoding/tree_cache.py:883-1018`` (``from_tree_cache_node``)
* ``flatten_stage1``         -> ``layers/attention/tree_attention.py:860-976`` (kernel2)
* ``node_stage1``            -> ``layers/attention/tree_attention.py:170-293``
* ``stage2_reduce``          -> ``layers/attention/tree_attention.py:297-416, 420-445, 485-546``
* ``flatten_fwd`` / ``node_fwd`` -> ``tree_attention.py:552-667`` / ``:14-68``
* ``leaf_paths`` / ``exact_attention`` -> the per-leaf torch check in
  ``tests/model/test_DeFT_kernel.py:212-276`` (fp64 here)
* ``seq_attention``          -> per-leaf semantics of ``layers/attention/token_attention.py:297-335``
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

BLOCK_LEN = 128  # tree_cache.py:587  BLOCK_CONFIG["BLOCK_LEN"]


# --------------------------------------------------------------------------------------------
# metadata tables
# --------------------------------------------------------------------------------------------
def _chunks(seq: Sequence[int], n: int):
    for i in range(0, len(seq), n):
        yield seq[i : i + n]


def _offsets(lens: List[int]) -> np.ndarray:
    # tree_cache.py:822-833 -- cat([0], cumsum(len)[:-1])
    out = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        out[1:] = np.cumsum(np.asarray(lens[:-1], dtype=np.int64))
    return out


def leaf_order(tree) -> Dict[int, int]:
    """tree_cache.py:650-652 -- query index of a leaf = its rank by ascending leaf id."""
    return {leaf.id: i for i, leaf in enumerate(sorted(tree.leaves.values(), key=lambda n: n.id))}


def build_tables(tree, max_q_len: int = 32, max_block_len: int = -1, block_len: int = BLOCK_LEN) -> Dict[str, object]:
    """Node tables and Flatten tables for one tree, as int64 numpy arrays.

    ``tree`` is duck-typed like the reference ``TreeCache``: ``.root``, ``.leaves`` and nodes with
    ``.id .kv_indices .children(dict, insertion ordered) .refs(set of leaf nodes) .paused``.
    """
    leaf_to_q = leaf_order(tree)
    node_q: List[int] = []
    node_kv: List[int] = []
    node_q_len: List[int] = []
    node_kv_len: List[int] = []
    block_q: List[int] = []
    block_q_cnts: List[int] = []
    block_kv: List[int] = []
    block_masks: List[int] = []
    block_lens: List[int] = []
    total_kv_len = 0

    # open block state (tree_cache.py:654-658)
    seg_tokens: List[int] = []          # page ids gathered so far
    seg_lens: List[int] = []            # tokens per segment
    seg_qsets: List[set] = []           # attending queries per segment
    union: set = set()

    def close_block() -> None:          # tree_cache.py:661-723
        n_live = len(seg_tokens)
        toks = list(seg_tokens)
        lens = list(seg_lens)
        qsets = list(seg_qsets)
        if n_live < block_len:
            toks += [-1] * (block_len - n_live)
            lens.append(block_len - n_live)
            qsets.append(set())
        for sub in _chunks(sorted(union), max_q_len):
            pos = {q: i for i, q in enumerate(sub)}
            block_q.extend(sub)
            block_q_cnts.append(len(sub))
            block_kv.extend(toks)
            block_lens.append(n_live)
            for qs, ln in zip(qsets, lens):
                bits = 0
                for q in qs:
                    if q in pos:
                        bits |= 1 << pos[q]
                block_masks.extend([bits] * ln)
        seg_tokens.clear(); seg_lens.clear(); seg_qsets.clear(); union.clear()

    def visit(node) -> None:            # tree_cache.py:725-791
        nonlocal total_kv_len
        assert len(node.refs) > 0
        if node.paused:
            return
        kv = sorted(node.kv_indices)
        total_kv_len += len(kv)
        q = sorted(leaf_to_q[r.id] for r in node.refs if not r.paused)
        step = len(kv) if max_block_len == -1 else max_block_len
        kv_parts = [kv[i : i + step] for i in range(0, len(kv), step)]  # raises on len(kv)==0, like the reference
        for qs in _chunks(q, max_q_len):
            for part in kv_parts:
                node_q.extend(qs); node_q_len.append(len(qs))
                node_kv.extend(part); node_kv_len.append(len(part))
        room = block_len - len(seg_tokens)
        done = 0
        while done < len(kv):
            if len(kv) - done < room:
                piece = kv[done:]
                seg_tokens.extend(piece); seg_lens.append(len(piece)); seg_qsets.append(set(q)); union.update(q)
                break
            piece = kv[done : done + room]
            seg_tokens.extend(piece); seg_lens.append(len(piece)); seg_qsets.append(set(q)); union.update(q)
            close_block()
            done += room
            room = block_len
        for child in node.children.values():
            visit(child)

    visit(tree.root)
    if seg_lens:                        # tree_cache.py:797-798
        close_block()

    i64 = lambda x: np.asarray(x, dtype=np.int64)
    return dict(
        query_num=len(leaf_to_q), node_num=len(node_q_len), total_kv_len=total_kv_len, leaf_to_q=leaf_to_q,
        block_len=block_len,
        node_q=i64(node_q), node_kv=i64(node_kv), node_q_len=i64(node_q_len), node_kv_len=i64(node_kv_len),
        node_q_offset=_offsets(node_q_len), node_kv_offset=_offsets(node_kv_len),
        block_q=i64(block_q), block_q_cnts=i64(block_q_cnts), block_q_offset=_offsets(block_q_cnts),
        block_bitmasks=i64(block_masks), block_kv=i64(block_kv), block_lens=i64(block_lens),
    )


def build_tables_tree_index(tree, max_ctx: int, max_q_len: int = 32, max_block_len: int = -1) -> Dict[str, object]:
    """``from_tree_cache_node``: node_kv is the whole node->page table, offsets index into it.

    ``node.node_indices_id`` is the row of the table a node owns; the table itself
    (``TreeIndexPool.node_to_kv``, int32 ``[size, max_ctx]``) is the caller's.
    """
    leaf_to_q = leaf_order(tree)
    node_q: List[int] = []; node_q_len: List[int] = []; node_kv_offset: List[int] = []; node_kv_len: List[int] = []
    total = 0

    def visit(node) -> None:
        nonlocal total
        assert len(node.refs) > 0
        if node.paused:
            return
        base = node.node_indices_id * max_ctx          # tree_index_pool.py:47-49
        n = len(node.kv_indices)
        total += n
        q = sorted(leaf_to_q[r.id] for r in node.refs if not r.paused)
        step = n if max_block_len == -1 else max_block_len
        parts = [(base + i, min(step, n - i)) for i in range(0, n, step)]
        for qs in _chunks(q, max_q_len):
            for off, ln in parts:
                node_q.extend(qs); node_q_len.append(len(qs)); node_kv_offset.append(off); node_kv_len.append(ln)
        for child in node.children.values():
            visit(child)

    visit(tree.root)
    i64 = lambda x: np.asarray(x, dtype=np.int64)
    return dict(query_num=len(leaf_to_q), node_num=len(node_q_len), total_kv_len=total, leaf_to_q=leaf_to_q,
                node_q=i64(node_q), node_q_len=i64(node_q_len), node_q_offset=_offsets(node_q_len),
                node_kv_offset=i64(node_kv_offset), node_kv_len=i64(node_kv_len))


# --------------------------------------------------------------------------------------------
# attention arithmetic (fp16 inputs, fp32 accumulation, as the Triton kernels)
# --------------------------------------------------------------------------------------------
def _f32(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32)


def flatten_stage1(q, K, V, t, block_len: int = BLOCK_LEN):
    """One softmax partial per (head, block_q slot).  q [nq,H,D]; K,V [pool,HKV,D] (fp16 arrays).

    Returns partial_o [H, P, D] fp32 and partial_lse [H, P] fp32, P = len(block_q).
    """
    nq, H, D = q.shape
    HKV = K.shape[1]
    G = H // HKV
    P = len(t["block_q"])
    po = np.zeros((H, P, D), np.float32)
    pl = np.zeros((H, P), np.float32)
    scale = np.float32(1.0 / math.sqrt(D))
    rows = np.arange(32, dtype=np.int64)
    for b in range(len(t["block_q_cnts"])):
        cnt = int(t["block_q_cnts"][b]); off = int(t["block_q_offset"][b]); ln = int(t["block_lens"][b])
        qi = t["block_q"][off : off + cnt]
        pages = t["block_kv"][b * block_len : b * block_len + ln]
        bits = t["block_bitmasks"][b * block_len : b * block_len + ln]
        allow = ((bits[None, :] >> rows[:cnt, None]) & 1).astype(bool)        # [cnt, ln]
        k = _f32(K[pages])                                                    # [ln, HKV, D]
        v = _f32(V[pages])
        qq = _f32(q[qi]).reshape(cnt, HKV, G, D)
        s = np.einsum("chgd,nhd->hgcn", qq, k, optimize=True) * scale         # [HKV,G,cnt,ln]
        s = np.where(allow[None, None], s, -np.inf)
        m = s.max(-1)
        p = np.exp(s - m[..., None])
        l = p.sum(-1)
        acc = np.einsum("hgcn,nhd->hgcd", p, v, optimize=True)
        po[:, off : off + cnt] = (acc / l[..., None]).reshape(H, cnt, D)
        pl[:, off : off + cnt] = (m + np.log(l)).reshape(H, cnt)
    return po, pl


def node_stage1(q, K, V, node_kv, node_kv_offset, node_kv_len, node_q, node_q_offset, node_q_len, step: int = 16):
    """Node / Node-Chunk / Tree-Index stage 1: online softmax over ``step``-token slices."""
    nq, H, D = q.shape
    HKV = K.shape[1]
    G = H // HKV
    P = len(node_q)
    po = np.zeros((H, P, D), np.float32)
    pl = np.zeros((H, P), np.float32)
    scale = np.float32(1.0 / math.sqrt(D))
    for e in range(len(node_kv_len)):
        qo, qn = int(node_q_offset[e]), int(node_q_len[e])
        ko, kn = int(node_kv_offset[e]), int(node_kv_len[e])
        qq = _f32(q[node_q[qo : qo + qn]]).reshape(qn, HKV, G, D)
        m = np.full((HKV, G, qn), -np.inf, np.float32)
        l = np.zeros((HKV, G, qn), np.float32)
        acc = np.zeros((HKV, G, qn, D), np.float32)
        for s0 in range(0, kn, step):
            pages = np.asarray(node_kv[ko + s0 : ko + min(s0 + step, kn)], dtype=np.int64)
            k = _f32(K[pages]); v = _f32(V[pages])
            s = np.einsum("chgd,nhd->hgcn", qq, k, optimize=True) * scale
            m_new = np.maximum(m, s.max(-1))
            p = np.exp(s - m_new[..., None])
            a = np.exp(m - m_new)
            acc = acc * a[..., None] + np.einsum("hgcn,nhd->hgcd", p, v, optimize=True)
            l = l * a + p.sum(-1)
            m = m_new
        po[:, qo : qo + qn] = (acc / l[..., None]).reshape(H, qn, D)
        pl[:, qo : qo + qn] = (m + np.log(l)).reshape(H, qn)
    return po, pl


def stage2_reduce(slot_to_q, po, pl, nq: int, faithful_fp16: bool = True):
    """Partial-softmax combine.  ``faithful_fp16`` follows the reference's arithmetic: row max
    floored at 0 (zero-initialised atomic max), weights accumulated into an fp16 output in slot
    order (the GPU's atomic order is unspecified; slot order is one legal order), then ``O / L``.
    With ``faithful_fp16=False`` the merge is done in fp32 with the true maximum (what the CUDA
    stage 2 of this repository does) and rounded to fp16 once.
    """
    H, P, D = po.shape
    slot_to_q = np.asarray(slot_to_q, dtype=np.int64)
    if faithful_fp16:
        rmax = np.zeros((H, nq), np.float32)
        np.maximum.at(rmax, (slice(None), slot_to_q), pl)
        w = np.exp(pl - rmax[:, slot_to_q])
        L = np.zeros((H, nq), np.float32)
        np.add.at(L, (slice(None), slot_to_q), w)
        O = np.zeros((H, nq, D), np.float16)
        contrib = (w[..., None] * po).astype(np.float16)
        for s in range(P):                       # fp16 read-modify-write per slot
            O[:, slot_to_q[s]] = (O[:, slot_to_q[s]].astype(np.float16) + contrib[:, s]).astype(np.float16)
        out = (O.astype(np.float32) / L[..., None]).astype(np.float16)
    else:
        rmax = np.full((H, nq), -np.inf, np.float32)
        np.maximum.at(rmax, (slice(None), slot_to_q), pl)
        w = np.exp(pl - rmax[:, slot_to_q])
        L = np.zeros((H, nq), np.float32)
        np.add.at(L, (slice(None), slot_to_q), w)
        O = np.zeros((H, nq, D), np.float32)
        np.add.at(O, (slice(None), slot_to_q), w[..., None] * po)
        out = (O / L[..., None]).astype(np.float16)
    return np.ascontiguousarray(out.transpose(1, 0, 2))      # [nq, H, D]


def flatten_fwd(q, K, V, t, faithful_fp16: bool = True):
    po, pl = flatten_stage1(q, K, V, t, t.get("block_len", BLOCK_LEN))
    return stage2_reduce(t["block_q"], po, pl, q.shape[0], faithful_fp16)


def node_fwd(q, K, V, t, faithful_fp16: bool = True, node_kv=None):
    kv = t["node_kv"] if node_kv is None else node_kv
    po, pl = node_stage1(q, K, V, kv, t["node_kv_offset"], t["node_kv_len"], t["node_q"], t["node_q_offset"], t["node_q_len"])
    return stage2_reduce(t["node_q"], po, pl, q.shape[0], faithful_fp16)


# --------------------------------------------------------------------------------------------
# per-leaf ("seq") semantics
# --------------------------------------------------------------------------------------------
def leaf_paths(tree) -> List[np.ndarray]:
    """Root->leaf page list of every leaf, in query order (ascending leaf id)."""
    out = []
    for leaf in sorted(tree.leaves.values(), key=lambda n: n.id):
        chain = []
        n = leaf
        while n is not None:
            chain.append(list(n.kv_indices))
            n = n.parent
        out.append(np.asarray([p for part in reversed(chain) for p in part], dtype=np.int64))
    return out


def exact_attention(q, K, V, paths, dtype=np.float64) -> np.ndarray:
    """softmax(q k^T / sqrt(D)) v per leaf over its own path, GQA by head // (H/HKV).  [nq,H,D] ``dtype``."""
    nq, H, D = q.shape
    HKV = K.shape[1]
    G = H // HKV
    out = np.zeros((nq, H, D), dtype)
    for i, pages in enumerate(paths):
        k = K[pages].astype(dtype); v = V[pages].astype(dtype)
        qq = q[i].astype(dtype).reshape(HKV, G, D)
        s = np.einsum("hgd,nhd->hgn", qq, k) / math.sqrt(D)
        s -= s.max(-1, keepdims=True)
        p = np.exp(s)
        p /= p.sum(-1, keepdims=True)
        out[i] = np.einsum("hgn,nhd->hgd", p, v).reshape(H, D)
    return out


def seq_attention(q, K, V, paths) -> np.ndarray:
    """The sequence-based baseline (Radix / Flash-Decoding semantics): every leaf re-reads its whole
    path.  fp32 arithmetic, fp16 result.  This is what ``bench.py`` times as the CPU baseline."""
    return exact_attention(q, K, V, paths, np.float32).astype(np.float16)
