"""Randomised decoding trees through the whole CUDA path (native tiler, job plan, both kernels) vs per-leaf attention.

Shapes the fixed workloads do not have: ragged node lengths around the 128-token tile edge, pruned branches (page
holes the allocator refills), uneven fan-out, more than two query slots, forests of unequal trees, a fused append on
top.  Needs a B200.
"""
import random

import numpy as np
import pytest
import torch

from oracle import deft_oracle as orc

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-3, 1e-2
TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]
GEOMS = [(32, 8, 128), (8, 2, 64), (16, 4, 128), (8, 8, 64), (4, 2, 128), (4, 2, 16), (8, 1, 32)]


def random_tree(rng, pool, r2t, layers, HKV, D):
    from deft_b200.tree_cache import TreeCache
    tree = TreeCache(torch.float16, HKV, D, layers, r2t, pool, None, True, False)
    tree.init_prompt(torch.arange(rng.choice([1, 17, 127, 128, 129, 300, 515]), dtype=torch.int32))

    def step(n):
        for _ in range(n):
            for leaf in tree.leaves.values():
                leaf.append_token(7)
            tree.alloc()

    for _ in range(rng.randint(1, 4)):
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        for leaf in rng.sample(leaves, k=max(1, len(leaves) // 2)):
            if len(tree.leaves) < 70:
                tree.branch(leaf, rng.choice([1, 2, 2, 3, 5, 9]))
        step(rng.choice([1, 2, 3, 16, 40]))
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        if len(leaves) > 3 and rng.random() < 0.5:
            tree.cut(rng.choice(leaves))            # frees pages: later allocations fill the holes
            step(1)
    if any(len(leaf.kv_indices) == 0 for leaf in tree.leaves.values()):
        step(1)
    return tree


def per_leaf(q, K, V, paths):
    nq, H, D = q.shape
    HKV = K.shape[1]
    out = torch.empty(nq, H, D, dtype=torch.float32, device=q.device)
    for i, p in enumerate(paths):
        idx = torch.as_tensor(p, device=q.device)
        k = K[idx].float().repeat_interleave(H // HKV, dim=1)
        v = V[idx].float().repeat_interleave(H // HKV, dim=1)
        s = torch.einsum("hd,nhd->hn", q[i].float(), k) / D ** 0.5
        out[i] = torch.einsum("hn,nhd->hd", torch.softmax(s, dim=-1), v)
    return out


@pytest.mark.parametrize("seed", range(14))
def test_random_trees_and_forests(seed, monkeypatch):
    import deft_b200
    from deft_b200 import TreeMetadata
    from deft_b200.memory_pool import ReqToTokenPool, TokenToKVPool
    rng = random.Random(seed)
    torch.manual_seed(seed)
    dev = torch.device("cuda:0")
    H, HKV, D = GEOMS[seed % len(GEOMS)]
    pool = TokenToKVPool(size=20000, dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=1, device=dev)
    r2t = ReqToTokenPool(size=512, max_context_len=2048, device=dev)
    pool.kv_data[0].normal_()
    K, V = pool.get_key_buffer(0), pool.get_value_buffer(0)
    trees = [random_tree(rng, pool, r2t, 1, HKV, D) for _ in range(rng.choice([1, 1, 2, 3]))]
    paths = [p for t in trees for p in orc.leaf_paths(t)]
    nq = len(paths)
    q = torch.randn(nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)[:, : H * D].view(nq, H, D)
    want = per_leaf(q, K, V, paths)

    def check(o, what):
        assert torch.isfinite(o.float()).all(), (seed, what)
        assert torch.allclose(o.float(), want, atol=ATOL, rtol=RTOL), (seed, what, (o.float() - want).abs().max().item())

    for regroup in ("1", "0"):
        monkeypatch.setenv("DEFT_PLAN_REGROUP", regroup)
        m = TreeMetadata.from_tree_cache(trees[0]) if len(trees) == 1 else TreeMetadata.from_forest(trees)
        assert m.query_num == nq
        o = torch.full((nq, H, D), float("nan"), dtype=torch.float16, device=dev)
        deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, m.block_q, m.block_q_cnts, m.block_q_offset, m.block_bitmasks,
                                             m.block_kv, m.block_lens)
        check(o, f"flatten regroup={regroup}")
        o2 = torch.full_like(o, float("nan"))
        deft_b200.tree_attention_fwd(q, K, V, o2, m.node_kv, m.node_kv_offset, m.node_kv_len, m.node_q, m.node_q_offset, m.node_q_len)
        check(o2, f"node regroup={regroup}")
    monkeypatch.delenv("DEFT_PLAN_REGROUP")
    # the reference's tables without a host plan (device-derived plan) on the same tree
    m = TreeMetadata.from_tree_cache(trees[0]) if len(trees) == 1 else TreeMetadata.from_forest(trees)
    t = {k: getattr(m, k).clone() for k in TABLE_KEYS}
    o3 = torch.full((nq, H, D), float("nan"), dtype=torch.float16, device=dev)
    deft_b200.tree_attention_subtree_fwd(q, K, V, o3, 128, t["block_q"], t["block_q_cnts"], t["block_q_offset"], t["block_bitmasks"],
                                         t["block_kv"], t["block_lens"])
    check(o3, "flatten, device plan")
    # a decode step on top, with the append fused in
    locs = []
    for tr in trees:
        for leaf in tr.leaves.values():
            leaf.append_token(7)
        locs.append(tr.alloc().cache_loc)
    loc_host = torch.cat(locs)
    loc = loc_host.to(dev)
    qkv = torch.randn(nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    q = qkv[:, : H * D].view(nq, H, D)
    k_new, v_new = qkv[:, H * D: (H + HKV) * D].view(nq, HKV, D), qkv[:, (H + HKV) * D:].view(nq, HKV, D)
    expect_pool = pool.kv_data[0].clone()
    expect_pool[loc.long(), 0] = k_new
    expect_pool[loc.long(), 1] = v_new
    pool.kv_data[0][loc.long()] = float("nan")
    fused = D in (64, 128) and H // HKV in (1, 2, 4)      # the fused append lives in the tensor-core kernels
    mf = (TreeMetadata.from_tree_cache(trees[0], fresh_page=loc_host if fused else None) if len(trees) == 1
          else TreeMetadata.from_forest(trees, fresh_page=loc_host if fused else None))
    o4 = torch.full((nq, H, D), float("nan"), dtype=torch.float16, device=dev)
    if fused:
        deft_b200.tree_attention_subtree_fwd(q, K, V, o4, 128, mf.block_q, mf.block_q_cnts, mf.block_q_offset, mf.block_bitmasks,
                                             mf.block_kv, mf.block_lens, append=(k_new, v_new, loc))
    else:
        with pytest.raises(deft_b200._lib.DeftError):    # ... and says so for the other geometries
            deft_b200.tree_attention_subtree_fwd(q, K, V, o4, 128, mf.block_q, mf.block_q_cnts, mf.block_q_offset,
                                                 mf.block_bitmasks, mf.block_kv, mf.block_lens, append=(k_new, v_new, loc))
        deft_b200.kv_append(pool.kv_data[0], k_new, v_new, loc)
        deft_b200.tree_attention_subtree_fwd(q, K, V, o4, 128, mf.block_q, mf.block_q_cnts, mf.block_q_offset, mf.block_bitmasks,
                                             mf.block_kv, mf.block_lens)
    torch.cuda.synchronize()
    assert torch.equal(pool.kv_data[0], expect_pool), (seed, "pool after the fused append")
    paths = [p for t_ in trees for p in orc.leaf_paths(t_)]
    want = per_leaf(q, K, V, paths)
    check(o4, "flatten, fused append")
