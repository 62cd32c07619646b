"""Generate ``tests/golden/*.npz`` by RUNNING THE REFERENCE.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference is mounted read-only at /root/reference and is never
copied):

    TRITON_INTERPRET=1 python oracle/gen_golden.py

What it records, per scenario of ``oracle/scenarios.py``:

* the tree the reference ``TreeCache`` ends up with after the scripted branch / decode / cut /
  speculative-merge operations (page ids are the reference allocator's),
* every table ``TreeMetadata.from_tree_cache`` returns, for ``MAX_BLOCK_LEN = -1`` (Node, Flatten)
  and ``= 128`` (Node-Chunk), and ``from_tree_cache_node`` in tree-index mode,
* the outputs of the reference Triton operators ``tree_attention_subtree_fwd`` (Flatten, plus its
  stage-1 partials), ``tree_attention_fwd`` (Node, Node-Chunk, Tree-Index) and
  ``token_attention_fwd`` (Radix / seq baseline), executed by the Triton interpreter on CPU.

The reference hard-codes ``device="cuda"`` in its torch factory calls; the shim below strips that
keyword.  Nothing else of the reference is altered.
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("TRITON_INTERPRET", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/DeFT")

import numpy as np
import torch

for _name in ["tensor", "empty", "ones", "zeros", "full", "arange"]:
    _f = getattr(torch, _name)
    setattr(torch, _name, (lambda f: lambda *a, **k: f(*a, **{kk: v for kk, v in k.items()
                                                             if not (kk == "device" and v == "cuda")}))(_f))
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.nvtx.range_push = torch.cuda.nvtx.range_pop = lambda *a, **k: None

from deft.memory_pool import ReqToTokenPool, TokenToKVPool  # noqa: E402
from deft.tree_decoding.tree_index_pool import TreeIndexPool  # noqa: E402
from deft.tree_decoding.tree_cache import TreeCache, TreeMetadata, BLOCK_CONFIG  # noqa: E402
from deft.layers.attention import tree_attention as ta  # noqa: E402
from deft.layers.attention.token_attention import token_attention_fwd  # noqa: E402

from oracle.plain_tree import freeze  # noqa: E402
from oracle.scenarios import SCENARIOS, TABLE_SCENARIOS, replay  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


def build(cfg, layer_num=1):
    H, HKV, D = cfg.get("H", 8), cfg.get("HKV", 2), cfg.get("D", 64)
    r2t = ReqToTokenPool(size=128, max_context_len=cfg["max_ctx"])
    kvp = TokenToKVPool(size=cfg["pool"], dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=layer_num)
    tix = TreeIndexPool(size=64, max_context_len=cfg["max_ctx"]) if cfg.get("tree_index") else None
    if tix is not None:
        tix.node_to_kv.zero_()
    r2t.req_to_token.zero_()
    tree = TreeCache(torch.float16, HKV, D, layer_num, r2t, kvp, tix, True, tix is not None)
    replay(tree, cfg["script"], lambda n: torch.arange(1, n + 1, dtype=torch.int32))
    return tree, r2t, kvp, tix


def tables(tree, mbl, prefix, out):
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = mbl
    m = TreeMetadata.from_tree_cache(tree)
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    for k in TABLE_KEYS:
        out[f"{prefix}{k}"] = getattr(m, k).numpy().astype(np.int64)
    out[f"{prefix}scalars"] = np.asarray([m.query_num, m.node_num, m.total_kv_len, m.block_len], dtype=np.int64)
    out[f"{prefix}leaf_to_q"] = np.asarray(sorted(m.leaf_to_q.items()), dtype=np.int64).reshape(-1, 2)
    return m


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    for name, cfg in SCENARIOS.items():
        torch.manual_seed(0)
        tree, r2t, kvp, tix = build(cfg)
        H, HKV, D = cfg["H"], cfg["HKV"], cfg["D"]
        kvp.kv_data[0].normal_()
        nq = len(tree.leaves)
        q_full = torch.randn(nq, (H + 2 * HKV) * D, dtype=torch.float16)
        q = q_full[:, : H * D].view(nq, H, D)                      # strided view, as the model's qkv split
        K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
        out = {f"tree_{k}": v for k, v in freeze(tree).items()}
        out["geom"] = np.asarray([H, HKV, D, cfg["pool"], cfg["max_ctx"]], dtype=np.int64)
        out["q"] = q.contiguous().numpy()
        out["kv_pool"] = kvp.kv_data[0].numpy()
        out["mem_state"] = kvp.mem_state.numpy().astype(np.int64)

        m = tables(tree, -1, "t_", out)
        captured = {}
        stage2 = ta.DeFT_splitBynode_Triton_stage2

        def spy(map_, po, pl, o_):
            captured["po"], captured["pl"] = po.clone(), pl.clone()
            return stage2(map_, po, pl, o_)

        ta.DeFT_splitBynode_Triton_stage2 = spy
        o = torch.zeros(nq, H, D, dtype=torch.float16)
        ta.tree_attention_subtree_fwd(q, K, V, o, m.block_len, m.block_q, m.block_q_cnts, m.block_q_offset,
                                      m.block_bitmasks, m.block_kv, m.block_lens)
        ta.DeFT_splitBynode_Triton_stage2 = stage2
        out["o_flatten"] = o.numpy().copy()
        out["flatten_partial_o"] = captured["po"].numpy()
        out["flatten_partial_lse"] = captured["pl"].numpy()

        o = torch.zeros(nq, H, D, dtype=torch.float16)
        ta.tree_attention_fwd(q, K, V, o, m.node_kv, m.node_kv_offset, m.node_kv_len, m.node_q, m.node_q_offset, m.node_q_len)
        out["o_node"] = o.numpy().copy()

        mc = tables(tree, 128, "tc_", out)
        o = torch.zeros(nq, H, D, dtype=torch.float16)
        ta.tree_attention_fwd(q, K, V, o, mc.node_kv, mc.node_kv_offset, mc.node_kv_len, mc.node_q, mc.node_q_offset, mc.node_q_len)
        out["o_node_chunk"] = o.numpy().copy()

        if tix is not None:
            BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
            mi = TreeMetadata.from_tree_cache_node(tree)
            BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
            out["node_to_kv"] = tix.node_to_kv.numpy().copy()
            for k in ["node_q", "node_q_len", "node_q_offset", "node_kv_offset", "node_kv_len"]:
                out[f"ti_{k}"] = getattr(mi, k).numpy().astype(np.int64)
            o = torch.zeros(nq, H, D, dtype=torch.float16)
            ta.tree_attention_fwd(q, K, V, o, mi.node_kv, mi.node_kv_offset, mi.node_kv_len, mi.node_q, mi.node_q_offset, mi.node_q_len)
            out["o_tree_index"] = o.numpy().copy()

        # Radix / seq baseline through the per-sequence page table (token_attention.py:297-335)
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        seq_lens = torch.tensor([l.positions[-1] + 1 for l in leaves], dtype=torch.int32)
        req_idx = torch.tensor([tree.leaf_to_req[l.id] for l in leaves], dtype=torch.int32)
        out["req_to_token"] = r2t.req_to_token.numpy().copy()
        out["req_idx"] = req_idx.numpy(); out["seq_lens"] = seq_lens.numpy()
        if name != "spec_merge":        # the speculative mock leaves req_to_token stale (reference behaviour)
            start_loc = torch.zeros(nq, dtype=torch.int32)
            start_loc[1:] = torch.cumsum(seq_lens[:-1], 0)
            o = torch.zeros(nq, H, D, dtype=torch.float16)
            token_attention_fwd(q, K, V, o, r2t.req_to_token, req_idx, start_loc, seq_lens, int(seq_lens.max()),
                                int(r2t.req_to_token[req_idx[0], seq_lens[0] - 1]), int(seq_lens.sum()))
            out["o_seq"] = o.numpy().copy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        print(name, "nq", nq, "blocks", len(m.block_lens), "partials", len(m.block_q), flush=True)

    for name, cfg in TABLE_SCENARIOS.items():
        tree, r2t, kvp, tix = build(dict(cfg, H=8, HKV=1, D=16))
        out = {f"tree_{k}": v for k, v in freeze(tree).items()}
        tables(tree, -1, "t_", out)
        tables(tree, 128, "tc_", out)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        print(name, "nodes", len(tree.nodes), flush=True)


if __name__ == "__main__":
    main()
