"""Parity of the CUDA path (through the C ABI) against the reference outputs and the oracle.  Needs a B200.

Tolerance is the north star's: fp16 ``atol=1e-3, rtol=1e-2`` against the reference DeFT-Flatten
(Triton) outputs stored in tests/golden, plus "at least as close to fp64 as the reference is".
"""
import os

import numpy as np
import pytest
import torch

from oracle import deft_oracle as orc
from oracle.plain_tree import thaw
from oracle.scenarios import SCENARIOS, replay

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-3, 1e-2
TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch.device("cuda:0")


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    return z, thaw({k[5:]: z[k] for k in z.files if k.startswith("tree_")})


def device_inputs(z, dev, strided=True):
    """q as the strided view into a fused qkv buffer (row stride (H+2HKV)*D), KV pool [pool,2,HKV,D]."""
    H, HKV, D = z["geom"][:3].tolist()
    q = torch.from_numpy(z["q"]).to(dev)
    if strided:
        full = torch.randn(q.shape[0], (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
        qv = full[:, : H * D].view(q.shape[0], H, D)
        qv.copy_(q)
        q = qv
    pool = torch.from_numpy(z["kv_pool"]).to(dev)
    return q, pool[:, 0], pool[:, 1]


def tables(z, prefix, dev):
    return {k: torch.from_numpy(z[prefix + k]).to(dev) for k in TABLE_KEYS}


def garbage_like(q):
    return torch.full((q.shape[0], q.shape[1], q.shape[2]), float("nan"), dtype=torch.float16, device=q.device)


def run_flatten(q, K, V, t, plan=None):
    import deft_b200
    o = garbage_like(q)    # the reference needs zeros; we must not
    deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, t["block_q"], t["block_q_cnts"], t["block_q_offset"],
                                         t["block_bitmasks"], t["block_kv"], t["block_lens"], plan=plan)
    return o


def run_node(q, K, V, t, node_kv=None, plan=None):
    import deft_b200
    o = garbage_like(q)
    deft_b200.tree_attention_fwd(q, K, V, o, t["node_kv"] if node_kv is None else node_kv, t["node_kv_offset"],
                                 t["node_kv_len"], t["node_q"], t["node_q_offset"], t["node_q_len"], plan=plan)
    return o


def assert_parity(got: torch.Tensor, want: np.ndarray, exact: np.ndarray = None, what=""):
    g = got.float().cpu().numpy()
    w = want.astype(np.float32)
    assert np.isfinite(g).all(), what
    assert np.allclose(g, w, atol=ATOL, rtol=RTOL), (what, float(np.abs(g - w).max()))
    if exact is not None:
        mine = np.abs(g.astype(np.float64) - exact).max()
        ref = np.abs(w.astype(np.float64) - exact).max()
        assert mine <= ref + 1e-4, (what, mine, ref)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_reference_tables_device_plan(golden_dir, dev, name):
    """Tables exactly as the reference builder made them (no host plan): plan derived on the device."""
    z, tree = load(golden_dir, name)
    q, K, V = device_inputs(z, dev)
    exact = orc.exact_attention(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], orc.leaf_paths(tree)) \
        if name != "spec_merge" else None
    t, tc = tables(z, "t_", dev), tables(z, "tc_", dev)
    assert_parity(run_flatten(q, K, V, t), z["o_flatten"], exact, "flatten")
    assert_parity(run_node(q, K, V, t), z["o_node"], exact, "node")
    assert_parity(run_node(q, K, V, tc), z["o_node_chunk"], exact, "node_chunk")
    # and against the oracle's restatement of the reference arithmetic
    want = orc.flatten_fwd(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], {k: z["t_" + k] for k in TABLE_KEYS},
                           faithful_fp16=False)
    assert_parity(run_flatten(q, K, V, t), want, None, "flatten-vs-oracle")


def build_on_device(cfg, z, dev):
    from deft_b200 import ReqToTokenPool, TokenToKVPool, TreeCache, TreeIndexPool
    H, HKV, D = cfg["H"], cfg["HKV"], cfg["D"]
    r2t = ReqToTokenPool(size=128, max_context_len=cfg["max_ctx"], device=dev)
    kvp = TokenToKVPool(size=cfg["pool"], dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=1, device=dev)
    tix = TreeIndexPool(size=64, max_context_len=cfg["max_ctx"], device=dev) if cfg.get("tree_index") else None
    tree = TreeCache(torch.float16, HKV, D, 1, r2t, kvp, tix, True, tix is not None)
    replay(tree, cfg["script"], lambda n: torch.arange(1, n + 1, dtype=torch.int32))
    kvp.kv_data[0].copy_(torch.from_numpy(z["kv_pool"]))
    return tree, kvp


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_own_metadata_host_plan(golden_dir, dev, name):
    """Tree replayed through deft_b200.TreeCache, tables + plan from the C++ builder, one upload."""
    from deft_b200 import BLOCK_CONFIG, TreeMetadata
    from deft_b200.tree_cache import lookup_plan
    z, ptree = load(golden_dir, name)
    tree, kvp = build_on_device(SCENARIOS[name], z, dev)
    q, _, _ = device_inputs(z, dev)
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    exact = orc.exact_attention(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], orc.leaf_paths(ptree)) \
        if name != "spec_merge" else None
    m = TreeMetadata.from_tree_cache(tree)
    for k in TABLE_KEYS:
        assert torch.equal(getattr(m, k).cpu(), torch.from_numpy(z["t_" + k])), k
    assert lookup_plan(m.block_q) is m and lookup_plan(m.node_q) is m
    t = {k: getattr(m, k) for k in TABLE_KEYS}
    o1 = run_flatten(q, K, V, t)                       # plan found through the registry
    assert_parity(o1, z["o_flatten"], exact, "flatten/host-plan")
    o2 = run_flatten(q, K, V, tables(z, "t_", dev))    # same tables, foreign copies -> device plan
    assert_parity(o2, z["o_flatten"], exact, "flatten/device-plan")
    # the native plan chains tiles and splits differently from the per-block device plan: same maths, other rounding
    assert torch.allclose(o1.float(), o2.float(), atol=5e-4, rtol=5e-3)
    assert torch.equal(o1, run_flatten(q, K, V, t)), "stage 2 must be deterministic"
    assert torch.equal(o2, run_flatten(q, K, V, tables(z, "t_", dev))), "stage 2 must be deterministic (device plan)"
    n1 = run_node(q, K, V, t)
    assert_parity(n1, z["o_node"], exact, "node/host-plan")
    assert torch.equal(n1, o1), "with the native plan both operators run the same work list"
    assert_parity(run_node(q, K, V, tables(z, "t_", dev)), z["o_node"], exact, "node/device-plan")
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        mc = TreeMetadata.from_tree_cache(tree)
    finally:
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    assert_parity(run_node(q, K, V, {k: getattr(mc, k) for k in TABLE_KEYS}), z["o_node_chunk"], exact, "node_chunk/host-plan")


def test_tree_index_mode(golden_dir, dev):
    from deft_b200 import BLOCK_CONFIG, TreeMetadata
    z, ptree = load(golden_dir, "tree_index")
    tree, kvp = build_on_device(SCENARIOS["tree_index"], z, dev)
    q, K, V = device_inputs(z, dev)
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        m = TreeMetadata.from_tree_cache_node(tree)
    finally:
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    assert m.node_kv.dtype == torch.int32
    t = {k: getattr(m, k) for k in ["node_q", "node_q_len", "node_q_offset", "node_kv_offset", "node_kv_len"]}
    exact = orc.exact_attention(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], orc.leaf_paths(ptree))
    assert_parity(run_node(q, K, V, t, node_kv=m.node_kv), z["o_tree_index"], exact, "tree_index/host-plan")
    # the reference's own int32 table + tables, no plan
    tt = {k: torch.from_numpy(z["ti_" + k]).to(dev) for k in t}
    table = torch.from_numpy(z["node_to_kv"]).to(dev).view(-1)
    assert_parity(run_node(q, K, V, tt, node_kv=table), z["o_tree_index"], exact, "tree_index/device-plan")


def test_contiguous_q_and_kv_append(golden_dir, dev):
    import deft_b200
    z, _ = load(golden_dir, "llama_flat8")
    q, K, V = device_inputs(z, dev, strided=False)
    assert_parity(run_flatten(q, K, V, tables(z, "t_", dev)), z["o_flatten"], None, "contiguous q")
    # KV append == the reference's two index_put (tree_cache.py:75-76), bit-exact
    pool = torch.from_numpy(z["kv_pool"]).to(dev)
    want = pool.clone()
    n, HKV, D = 8, pool.shape[2], pool.shape[3]
    qkv = torch.randn(n, (32 + 2 * HKV) * D, dtype=torch.float16, device=dev)
    k_new = qkv[:, 32 * D: (32 + HKV) * D].view(n, HKV, D)      # strided views, like the model's split
    v_new = qkv[:, (32 + HKV) * D:].view(n, HKV, D)
    loc = torch.tensor([401, 5, 77, 402, 0, 459, 13, 200], dtype=torch.int32, device=dev)
    want[:, 0][loc.long()] = k_new
    want[:, 1][loc.long()] = v_new
    deft_b200.kv_append(pool, k_new, v_new, loc)
    assert torch.equal(pool, want)


def per_leaf_reference(q, K, V, paths):
    """fp32 torch restatement of tests/model/test_DeFT_kernel.py:212-276 on the GPU (full-size checks)."""
    nq, H, D = q.shape
    HKV = K.shape[1]
    out = torch.empty(nq, H, D, dtype=torch.float32, device=q.device)
    for i, p in enumerate(paths):
        idx = torch.as_tensor(p, device=q.device)
        k = K[idx].float().repeat_interleave(H // HKV, dim=1)      # [n, H, D]
        v = V[idx].float().repeat_interleave(H // HKV, dim=1)
        s = torch.einsum("hd,nhd->hn", q[i].float(), k) / D ** 0.5
        out[i] = torch.einsum("hn,nhd->hd", torch.softmax(s, dim=-1), v)
    return out


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4"])
def test_baseline_configs_full_size(dev, name):
    """BASELINE.json shapes at full size: all three operator modes agree with per-leaf attention."""
    from deft_b200 import BLOCK_CONFIG, TreeMetadata
    from deft_b200.workloads import build_tree, unique_kv_tokens
    torch.manual_seed(0)
    tree = build_tree(name, layers=1, device=dev)
    kvp = tree.token_to_kv_pool
    kvp.kv_data[0].normal_()
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    nq = len(tree.leaves)
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    m = TreeMetadata.from_tree_cache(tree)
    assert m.total_kv_len == unique_kv_tokens(name) and m.query_num == nq
    want = per_leaf_reference(q, K, V, orc.leaf_paths(tree))
    t = {k: getattr(m, k) for k in TABLE_KEYS}
    outs = {"flatten": run_flatten(q, K, V, t), "node": run_node(q, K, V, t)}
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        mc = TreeMetadata.from_tree_cache(tree)
    finally:
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    outs["node_chunk"] = run_node(q, K, V, {k: getattr(mc, k) for k in TABLE_KEYS})
    for mode, o in outs.items():
        err = (o.float() - want).abs().max().item()
        assert torch.allclose(o.float(), want, atol=ATOL, rtol=RTOL), (name, mode, err)
    # size-independent property: scaling V scales the output (linearity in V), softmax weights unchanged
    kvp.kv_data[0][:, 1].mul_(0.5)
    o_half = run_flatten(q, K, V, t)
    assert torch.allclose(o_half.float() * 2, outs["flatten"].float(), atol=2e-3, rtol=1e-2)


@pytest.mark.parametrize("name", [n for n in SCENARIOS if n != "spec_merge"])
def test_sequence_mode_matches_reference(golden_dir, dev, name):
    """Radix / seq mode (token_attention_fwd, token_attention.py:297-335) through the same kernels: every leaf
    attends its own row of req_to_token; compared with the reference's Triton output and fp64 per-leaf attention."""
    import deft_b200
    z, tree = load(golden_dir, name)
    q, K, V = device_inputs(z, dev)
    r2t = torch.from_numpy(z["req_to_token"]).to(dev)
    req_idx = torch.from_numpy(z["req_idx"]).to(dev)
    seq_lens = torch.from_numpy(z["seq_lens"]).to(dev)
    o = garbage_like(q)
    deft_b200.token_attention_fwd(q, K, V, o, r2t, req_idx, torch.zeros_like(seq_lens), seq_lens,
                                  int(z["seq_lens"].max()), None, int(z["seq_lens"].sum()))
    exact = orc.exact_attention(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], orc.leaf_paths(tree))
    assert_parity(o, z["o_seq"], exact, "seq")


def test_forest_batch_in_one_launch(dev):
    """BASELINE cfg 5 shape, scaled: several independent trees over ONE page pool attended by one call.  Every
    query sees exactly its own tree (per-leaf check), and the result equals the trees run one at a time."""
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_forest
    torch.manual_seed(5)
    trees = build_forest("cfg3", 3, layers=1, device=dev) + []
    kvp = trees[0].token_to_kv_pool
    kvp.kv_data[0].normal_()
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    m = TreeMetadata.from_forest(trees)
    nq = m.query_num
    assert nq == sum(len(t.leaves) for t in trees) == 192
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    t = {k: getattr(m, k) for k in TABLE_KEYS}
    got = run_flatten(q, K, V, t)
    paths = [p for tr in trees for p in orc.leaf_paths(tr)]
    want = per_leaf_reference(q, K, V, paths)
    assert torch.allclose(got.float(), want, atol=ATOL, rtol=RTOL), (got.float() - want).abs().max().item()
    assert torch.equal(got, run_node(q, K, V, t))
    off = 0
    for tr in trees:                                     # one tree at a time, same pool
        mt = TreeMetadata.from_tree_cache(tr)
        n = len(tr.leaves)
        one = run_flatten(q[off: off + n], K, V, {k: getattr(mt, k) for k in TABLE_KEYS})
        assert torch.allclose(one.float(), got[off: off + n].float(), atol=5e-4, rtol=5e-3)
        off += n


def test_forest_of_64_cfg2_trees_matches_per_leaf_attention(dev):
    """The regime bench.py's cfg5 block measures (64 cfg2 trees per GPU: pair-aligned plan, ~76 jobs per CTA, cluster
    multicast), against per-leaf fp32 attention on a sample of 8 of the 64 trees, and against the unpaired plan."""
    from deft_b200 import TreeMetadata, _lib
    from deft_b200.workloads import build_forest
    torch.manual_seed(64)
    trees = build_forest("cfg2", 64, layers=1, device=dev)
    kvp = trees[0].token_to_kv_pool
    kvp.kv_data[0].normal_()
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    m = TreeMetadata.from_forest(trees)
    nq = m.query_num
    assert nq == 64 * 64 and m.flat_plan.paired == 1, "a forest this size is planned as cluster pairs"
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    t = {k: getattr(m, k) for k in TABLE_KEYS}
    got = run_flatten(q, K, V, t)
    assert torch.isfinite(got.float()).all()
    for ti in (0, 7, 13, 21, 34, 42, 55, 63):
        want = per_leaf_reference(q[ti * 64: (ti + 1) * 64], K, V, orc.leaf_paths(trees[ti]))
        g = got[ti * 64: (ti + 1) * 64].float()
        assert torch.allclose(g, want, atol=ATOL, rtol=RTOL), (ti, (g - want).abs().max().item())
    assert torch.equal(got, run_flatten(q, K, V, t)), "deterministic"
    assert torch.allclose(got.float(), run_node(q, K, V, t).float(), atol=5e-4, rtol=5e-3)


def test_two_devices_in_one_process():
    """Per-device launch configuration (cudaFuncSetAttribute, SM count): the same process attends on cuda:0 and on
    cuda:1, from either current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_tree
    outs = []
    for d in (0, 1, 0):
        dv = torch.device("cuda", d)
        torch.manual_seed(3)
        tree = build_tree("cfg3", layers=1, device=dv)
        kvp = tree.token_to_kv_pool
        kvp.kv_data[0].copy_(torch.randn(kvp.kv_data[0].shape, dtype=torch.float16))
        nq = len(tree.leaves)
        q = torch.randn(nq, 32, 128, dtype=torch.float16).to(dv)
        m = TreeMetadata.from_tree_cache(tree)
        torch.cuda.set_device(0)                 # the call finds its device from the tensors, not from the current one
        o = run_flatten(q, kvp.get_key_buffer(0), kvp.get_value_buffer(0), {k: getattr(m, k) for k in TABLE_KEYS})
        want = per_leaf_reference(q, kvp.get_key_buffer(0), kvp.get_value_buffer(0), orc.leaf_paths(tree))
        assert torch.allclose(o.float(), want, atol=ATOL, rtol=RTOL), d
        outs.append(o.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_argument_errors(dev):
    import deft_b200
    q = torch.zeros(2, 8, 48, dtype=torch.float16, device=dev)      # head_dim 48: unsupported, like the reference
    kv = torch.zeros(16, 2, 2, 48, dtype=torch.float16, device=dev)
    i = torch.zeros(1, dtype=torch.int64, device=dev)
    with pytest.raises(AssertionError):
        deft_b200.tree_attention_subtree_fwd(q, kv[:, 0], kv[:, 1], torch.zeros_like(q), 128, i, i, i, i, i, i)
    with pytest.raises(deft_b200._lib.DeftError):
        deft_b200.tree_attention_subtree_fwd(q.cpu(), kv[:, 0].cpu(), kv[:, 1].cpu(), torch.zeros_like(q).cpu(), 128, i, i, i, i, i, i)


def test_decode_step_graph_replays_with_new_table_contents(dev):
    """DecodeStepGraph: the step (kv_append + attention per layer) captured once, replayed while the table layout
    is unchanged -- with DIFFERENT page ids in the tables (another tree of the same shape over other pages) and new
    activations -- and equal to the eager per-layer calls."""
    import deft_b200
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_forest
    torch.manual_seed(5)
    L, H, HKV, D = 3, 32, 8, 128
    trees = build_forest("cfg3", 2, layers=L, device=dev)      # two trees of one shape over one pool
    kvp = trees[0].token_to_kv_pool
    for l in range(L):
        kvp.kv_data[l].normal_()
    nq = len(trees[0].leaves)
    qkv = torch.empty(L, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    out = torch.empty(L, nq, H, D, dtype=torch.float16, device=dev)
    loc = torch.zeros(nq, dtype=torch.int32, device=dev)
    step = deft_b200.DecodeStepGraph(kvp, qkv, out, loc, H, HKV, D, mode="flatten", chunk=2, reference_tables=True)
    layouts = set()
    for it, tree in enumerate([trees[0], trees[1], trees[0]]):
        qkv.normal_()
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        loc.copy_(torch.tensor([leaf.kv_indices[-1] for leaf in leaves], dtype=torch.int32))
        m = step.metadata(tree)
        layouts.add(m.layout)
        step.run(m)
        got = out.clone()
        # eager reference on a copy of the pool state: the appended rows are already in place, attention only
        m2 = TreeMetadata.from_tree_cache(tree)
        # (from the second build on, m shares its views and plans with the previous step's metadata: same layout,
        # same persistent buffer -- they must show THIS step's tables)
        assert m.leaf_to_q == m2.leaf_to_q
        for k in ("block_q", "block_kv", "block_bitmasks", "node_kv", "node_q"):
            assert torch.equal(getattr(m, k), getattr(m2, k)), (it, k)
        want = torch.empty_like(out)
        for l in range(L):
            deft_b200.tree_attention_subtree_fwd(qkv[l, :, : H * D].view(nq, H, D), kvp.get_key_buffer(l), kvp.get_value_buffer(l),
                                                 want[l], 128, m2.block_q, m2.block_q_cnts, m2.block_q_offset, m2.block_bitmasks,
                                                 m2.block_kv, m2.block_lens)
            assert torch.equal(kvp.get_key_buffer(l)[loc.long()], qkv[l, :, H * D: (H + HKV) * D].view(nq, HKV, D))
        assert torch.isfinite(got.float()).all() and torch.equal(got, want), it
    assert len(layouts) == 1 and step.captures == 1, "same shape, other pages: one capture, replayed"


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg3b", "cfg1"])
def test_fused_kv_append_matches_append_then_attend(dev, name):
    """KVCacheUpdater.update + attention (deft_attention.py:120-148) in the attention's own two launches: this step's
    K/V rows are read from the activations and written to their pages by stage 2.  Same outputs as scattering first
    and attending over the pool, and bit-identical pool contents."""
    import deft_b200
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_tree
    torch.manual_seed(21)
    H, HKV, D = 32, 8, 128
    tree = build_tree(name, layers=2, device=dev, headroom=256)
    kvp = tree.token_to_kv_pool
    kvp.kv_data[0].normal_()
    for leaf in tree.leaves.values():
        leaf.append_token(7)
    upd = tree.alloc()                                    # this step's pages
    loc = upd.cache_loc.to(dev)
    kvp.kv_data[0][loc.long()] = float("nan")             # nothing may read them before they are written
    kvp.kv_data[1].copy_(kvp.kv_data[0])
    nq = len(tree.leaves)
    qkv = torch.randn(nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    q = qkv[:, : H * D].view(nq, H, D)
    k_new, v_new = qkv[:, H * D: (H + HKV) * D].view(nq, HKV, D), qkv[:, (H + HKV) * D:].view(nq, HKV, D)
    # reference order: scatter, then attend (pool 1)
    m0 = TreeMetadata.from_tree_cache(tree)
    deft_b200.kv_append(kvp.kv_data[1], k_new, v_new, loc)
    want = run_flatten(q, kvp.get_key_buffer(1), kvp.get_value_buffer(1), {k: getattr(m0, k) for k in TABLE_KEYS})
    exact = per_leaf_reference(q, kvp.get_key_buffer(1), kvp.get_value_buffer(1), orc.leaf_paths(tree))
    # fused (pool 0)
    m = TreeMetadata.from_tree_cache(tree, fresh_page=upd.cache_loc)
    assert m.flat_plan.fresh == 1 and all(torch.equal(getattr(m, k), getattr(m0, k)) for k in TABLE_KEYS)
    for op in ("flatten", "node"):
        kvp.kv_data[0].copy_(kvp.kv_data[1])
        kvp.kv_data[0][loc.long()] = float("nan")
        o = garbage_like(q)
        if op == "flatten":
            deft_b200.tree_attention_subtree_fwd(q, kvp.get_key_buffer(0), kvp.get_value_buffer(0), o, 128, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens,
                                                 append=(k_new, v_new, loc))
        else:
            deft_b200.tree_attention_fwd(q, kvp.get_key_buffer(0), kvp.get_value_buffer(0), o, m.node_kv, m.node_kv_offset,
                                         m.node_kv_len, m.node_q, m.node_q_offset, m.node_q_len, append=(k_new, v_new, loc))
        assert torch.isfinite(o.float()).all(), op
        assert torch.allclose(o.float(), exact, atol=ATOL, rtol=RTOL), (op, (o.float() - exact).abs().max().item())
        assert torch.allclose(o.float(), want.float(), atol=5e-4, rtol=5e-3), op
        assert torch.equal(kvp.kv_data[0], kvp.kv_data[1]), "pool contents after the fused call = after the two index_puts"
    with pytest.raises(deft_b200._lib.DeftError):        # a plan with fresh tokens cannot run without the activations
        deft_b200.tree_attention_subtree_fwd(q, kvp.get_key_buffer(0), kvp.get_value_buffer(0), garbage_like(q), 128, m.block_q,
                                             m.block_q_cnts, m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", ["flatten", "node"])
def test_decode_step_graph_survives_the_tree_growing(dev, mode, fused):
    """A real decode loop (tree_generate.py:109, 128): every step appends a token and a page per leaf, rebuilds the
    tables, appends K/V and attends.  With capacity-padded tables the captured graphs are replayed across the appends
    (a handful of captures over 40 steps instead of one per step), and every step equals the eager per-layer calls on
    a tightly packed build of the same tree."""
    import deft_b200
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_tree
    torch.manual_seed(11)
    L, H, HKV, D, steps = 2, 32, 8, 128, 40
    tree = build_tree("cfg3", layers=L, device=dev, headroom=64 * (steps + 2))
    kvp = tree.token_to_kv_pool
    for l in range(L):
        kvp.kv_data[l].normal_()
    nq = len(tree.leaves)
    qkv = torch.empty(L, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    out = torch.empty(L, nq, H, D, dtype=torch.float16, device=dev)
    loc = torch.zeros(nq, dtype=torch.int32, device=dev)
    # (the unfused variant also keeps the reference's tables in the step's metadata; by default they are left empty)
    step = deft_b200.DecodeStepGraph(kvp, qkv, out, loc, H, HKV, D, mode=mode, chunk=1, reference_tables=not fused)
    for it in range(steps):
        for leaf in tree.leaves.values():
            leaf.append_token(7)
        upd = tree.alloc()
        loc.copy_(upd.cache_loc)
        qkv.normal_()
        m = step.metadata(tree, cache_loc=upd.cache_loc if fused else None)
        step.run(m)
        got = out.clone()
        assert m.flat_plan.fresh == int(fused)
        m2 = TreeMetadata.from_tree_cache(tree)                 # tight packing, fresh buffer, plain plan
        assert m2.total_kv_len == m.total_kv_len == 2048 + 64 * (it + 2)
        for k in ("block_q", "block_kv", "block_bitmasks", "node_kv", "node_q", "node_kv_len"):
            if fused:
                assert getattr(m, k).numel() == 0, (it, k)
            else:
                assert torch.equal(getattr(m, k), getattr(m2, k)), (it, k)
        want = torch.empty_like(out)
        for l in range(L):
            q = qkv[l, :, : H * D].view(nq, H, D)
            K, V = kvp.get_key_buffer(l), kvp.get_value_buffer(l)
            assert torch.equal(K[loc.long()], qkv[l, :, H * D: (H + HKV) * D].view(nq, HKV, D)), "the graph appended this step's K"
            if mode == "flatten":
                deft_b200.tree_attention_subtree_fwd(q, K, V, want[l], 128, m2.block_q, m2.block_q_cnts, m2.block_q_offset,
                                                     m2.block_bitmasks, m2.block_kv, m2.block_lens)
            else:
                deft_b200.tree_attention_fwd(q, K, V, want[l], m2.node_kv, m2.node_kv_offset, m2.node_kv_len, m2.node_q,
                                             m2.node_q_offset, m2.node_q_len)
        assert torch.isfinite(got.float()).all(), (mode, it)
        if fused:   # (other token order inside the tiles: equal up to the rounding of the sums)
            assert torch.allclose(got.float(), want.float(), atol=5e-4, rtol=5e-3), (mode, it)
        else:
            assert torch.equal(got, want), (mode, it)
    assert step.captures <= 8, f"{step.captures} captures over {steps} appends (the subtree grows 40-fold)"


@pytest.mark.parametrize("mode", ["flatten", "node"])
def test_decode_step_pipeline_prepares_the_next_step_under_the_running_one(dev, mode):
    """DecodeStepPipeline: alloc + table build + upload of step t+1 are issued right after step t is enqueued (no
    synchronisation in between: the two halves own their table buffers), and every step still equals the eager calls on
    a tight build of the tree as it was for THAT step."""
    import deft_b200
    from deft_b200 import TreeMetadata
    from deft_b200.workloads import build_tree
    torch.manual_seed(12)
    L, H, HKV, D, steps = 2, 32, 8, 128, 24
    tree = build_tree("cfg3", layers=L, device=dev, headroom=64 * (steps + 3))
    kvp = tree.token_to_kv_pool
    for l in range(L):
        kvp.kv_data[l].normal_()
    nq = len(tree.leaves)
    qkv = torch.empty(L, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    out = torch.empty(L, nq, H, D, dtype=torch.float16, device=dev)
    pipe = deft_b200.DecodeStepPipeline(kvp, qkv, out, H, HKV, D, mode=mode, chunk=1)

    def host_side():
        for leaf in tree.leaves.values():
            leaf.append_token(7)
        upd = tree.alloc()
        m = pipe.prepare(tree, cache_loc=upd.cache_loc)
        assert m.flat_plan.fresh == 1
        return upd.cache_loc.clone(), TreeMetadata.from_tree_cache(tree)

    with pytest.raises(AssertionError):
        pipe.run()                                        # nothing prepared yet
    nxt = host_side()
    for it in range(steps):
        locs, m2 = nxt
        qkv.normal_()
        pipe.run()
        nxt = host_side()                                 # the tree already holds step it+1 while step `it` runs
        torch.cuda.synchronize()
        got = out.clone()
        want = torch.empty_like(out)
        for l in range(L):
            q = qkv[l, :, : H * D].view(nq, H, D)
            K, V = kvp.get_key_buffer(l), kvp.get_value_buffer(l)
            assert torch.equal(K[locs.to(dev).long()], qkv[l, :, H * D: (H + HKV) * D].view(nq, HKV, D)), "this step's K is in the pool"
            if mode == "flatten":
                deft_b200.tree_attention_subtree_fwd(q, K, V, want[l], 128, m2.block_q, m2.block_q_cnts, m2.block_q_offset,
                                                     m2.block_bitmasks, m2.block_kv, m2.block_lens)
            else:
                deft_b200.tree_attention_fwd(q, K, V, want[l], m2.node_kv, m2.node_kv_offset, m2.node_kv_len, m2.node_q,
                                             m2.node_q_offset, m2.node_q_len)
        assert torch.isfinite(got.float()).all(), (mode, it)
        assert torch.allclose(got.float(), want.float(), atol=5e-4, rtol=5e-3), (mode, it)
    assert pipe.captures <= 10, pipe.captures
