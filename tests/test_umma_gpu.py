"""tcgen05 stage-1 kernel: accumulator-level checks (raw S and O of the first work unit) and parity of both
stage-1 implementations.  Needs a B200."""
import os

import numpy as np
import pytest
import torch

from oracle import deft_oracle as orc
from oracle.plain_tree import thaw
from oracle.scenarios import SCENARIOS

pytestmark = pytest.mark.gpu
TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture
def force_impl():
    from deft_b200 import _lib
    def set_impl(i):
        _lib.lib.deft_b200_set_stage1_impl(i)
    yield set_impl
    _lib.lib.deft_b200_set_stage1_impl(_lib.STAGE1_AUTO)
    _lib.lib.deft_b200_set_debug_buffer(None)


def flatten(z, dev, q, K, V):
    import deft_b200
    t = {k: torch.from_numpy(z["t_" + k]).to(dev) for k in TABLE_KEYS}
    o = torch.full_like(q, float("nan"))
    deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, t["block_q"], t["block_q_cnts"], t["block_q_offset"],
                                         t["block_bitmasks"], t["block_kv"], t["block_lens"])
    return o


@pytest.mark.parametrize("name", ["llama_flat8", "toy_binary"])
def test_accumulators_of_first_unit(golden_dir, dev, force_impl, name):
    """S = Q K^T and O = P V of (item 0, kv-head 0, group 0) read back from TMEM, vs fp32 torch."""
    from deft_b200 import _lib
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    H, HKV, D = z["geom"][:3].tolist()
    G = H // HKV
    q = torch.from_numpy(z["q"]).to(dev)
    pool = torch.from_numpy(z["kv_pool"]).to(dev)
    K, V = pool[:, 0], pool[:, 1]
    dbg = torch.zeros(128 * 128 + 128 * D, dtype=torch.float32, device=dev)
    force_impl(_lib.STAGE1_UMMA)
    _lib.lib.deft_b200_set_debug_buffer(dbg.data_ptr())
    o = flatten(z, dev, q, K, V)
    torch.cuda.synchronize()
    _lib.lib.deft_b200_set_debug_buffer(None)
    S = dbg[: 128 * 128].view(128, 128).cpu()
    O = dbg[128 * 128:].view(128, D).cpu()
    cnt, ln = int(z["t_block_q_cnts"][0]), int(z["t_block_lens"][0])
    qids = torch.from_numpy(z["t_block_q"][:cnt])
    pages = torch.from_numpy(z["t_block_kv"][:ln])
    bits = torch.from_numpy(z["t_block_bitmasks"][:ln])
    qq = q.cpu().float()[qids][:, :G]                      # [cnt, G, D] heads of kv-head 0
    kk = K.cpu().float()[pages][:, 0]                      # [ln, D]
    vv = V.cpu().float()[pages][:, 0]
    S_want = torch.einsum("cgd,nd->cgn", qq, kk).reshape(cnt * G, ln)
    err_s = (S[: cnt * G, :ln] - S_want).abs().max().item()
    print(f"{name}: max|S - QK^T| = {err_s:.3e} (|S| max {S_want.abs().max():.2f})")
    assert err_s < 2e-2 * max(1.0, S_want.abs().max().item()), "QK^T through tcgen05 is wrong (descriptor / layout?)"
    assert S[cnt * G:, :].abs().max().item() == 0.0, "rows past the group must be zero (zero-filled Q)"
    assert ln == 128 or S[:, ln:].abs().max().item() == 0.0, "columns past the tile end must be zero (zero-filled K)"
    allow = ((bits[None, :] >> torch.arange(cnt)[:, None]) & 1).bool().repeat_interleave(G, dim=0)
    sc = S_want / D ** 0.5
    sc = torch.where(allow, sc, torch.tensor(float("-inf")))
    m = sc.max(dim=1, keepdim=True).values   # the first tile's exact row maximum (agreed by the row's two threads)
    P = torch.exp(sc - m)
    O_want = P @ vv
    err_o = (O[: cnt * G] - O_want).abs().max().item()
    print(f"{name}: max|O - PV| = {err_o:.3e} (|O| max {O_want.abs().max():.2f})")
    assert err_o < 2e-2 * max(1.0, O_want.abs().max().item()), "PV through tcgen05 is wrong (V / P descriptor?)"
    want = z["o_flatten"].astype(np.float32)
    assert np.allclose(o.float().cpu().numpy(), want, atol=1e-3, rtol=1e-2)


@pytest.mark.parametrize("impl", ["fma", "umma"])
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_both_stage1_kernels_match_reference(golden_dir, dev, force_impl, name, impl):
    import deft_b200
    from deft_b200 import _lib
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    q = torch.from_numpy(z["q"]).to(dev)
    pool = torch.from_numpy(z["kv_pool"]).to(dev)
    K, V = pool[:, 0], pool[:, 1]
    force_impl(_lib.STAGE1_FMA if impl == "fma" else _lib.STAGE1_UMMA)
    o = flatten(z, dev, q, K, V)
    assert np.allclose(o.float().cpu().numpy(), z["o_flatten"].astype(np.float32), atol=1e-3, rtol=1e-2)
    for prefix, key in (("t_", "o_node"), ("tc_", "o_node_chunk")):
        t = {k: torch.from_numpy(z[prefix + k]).to(dev) for k in TABLE_KEYS}
        o = torch.full_like(q, float("nan"))
        deft_b200.tree_attention_fwd(q, K, V, o, t["node_kv"], t["node_kv_offset"], t["node_kv_len"], t["node_q"],
                                     t["node_q_offset"], t["node_q_len"])
        got = o.float().cpu().numpy()
        assert np.isfinite(got).all(), (impl, key)
        assert np.allclose(got, z[key].astype(np.float32), atol=1e-3, rtol=1e-2), (impl, key, np.abs(got - z[key]).max())


def test_kernels_agree_at_full_size(dev, force_impl):
    """cfg2 at full size: tensor-core and warp-FMA stage 1 agree to fp16 rounding of P."""
    from deft_b200 import TreeMetadata, _lib
    import deft_b200
    from deft_b200.workloads import build_tree
    torch.manual_seed(1)
    tree = build_tree("cfg2", layers=1, device=dev)
    kvp = tree.token_to_kv_pool
    kvp.kv_data[0].normal_()
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    nq = len(tree.leaves)
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    m = TreeMetadata.from_tree_cache(tree)
    outs = {}
    for impl in (_lib.STAGE1_FMA, _lib.STAGE1_UMMA):
        force_impl(impl)
        o = torch.empty(nq, 32, 128, dtype=torch.float16, device=dev)
        deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, m.block_q, m.block_q_cnts, m.block_q_offset,
                                             m.block_bitmasks, m.block_kv, m.block_lens)
        outs[impl] = o.float()
    assert torch.allclose(outs[_lib.STAGE1_FMA], outs[_lib.STAGE1_UMMA], atol=1e-3, rtol=1e-2)


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_tma_and_cp_async_paths_agree(dev, name):
    """Runs of consecutive pages / query ids go through TMA boxes, scattered pages through TMA gather4 (four rows
    per instruction) or, with that switched off, cp.async gathers: all stage the same bytes in the same swizzled
    layout, so the outputs are bit-identical."""
    from deft_b200 import TreeMetadata, _lib
    import deft_b200
    from deft_b200.workloads import build_tree
    torch.manual_seed(2)
    tree = build_tree(name, layers=1, device=dev)
    kvp = tree.token_to_kv_pool
    kvp.kv_data[0].normal_()
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    nq = len(tree.leaves)
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    m = TreeMetadata.from_tree_cache(tree)
    outs = []
    try:
        for tma, g4 in ((1, 1), (1, 0), (0, 0)):
            _lib.lib.deft_b200_set_tma(tma)
            _lib.lib.deft_b200_set_gather4(g4)
            o = torch.full((nq, 32, 128), float("nan"), dtype=torch.float16, device=dev)
            deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, m.block_q, m.block_q_cnts, m.block_q_offset,
                                                 m.block_bitmasks, m.block_kv, m.block_lens)
            outs.append(o)
    finally:
        _lib.lib.deft_b200_set_tma(1)
        _lib.lib.deft_b200_set_gather4(1)
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_programmatic_dependent_launch_is_race_free(dev, name):
    """Back-to-back calls sharing one workspace, with and without programmatic dependent launch: the kernels
    overlap only their dependency-free prologues, so every call gives the bit-identical result."""
    from deft_b200 import TreeMetadata, _lib
    import deft_b200
    from deft_b200.workloads import build_tree
    torch.manual_seed(4)
    tree = build_tree(name, layers=4, device=dev)
    kvp = tree.token_to_kv_pool
    for l in range(4):
        kvp.kv_data[l].normal_()
    nq = len(tree.leaves)
    q = torch.randn(4, nq, 48 * 128, dtype=torch.float16, device=dev)
    m = TreeMetadata.from_tree_cache(tree)

    def run(pdl):
        _lib.lib.deft_b200_set_pdl(pdl)
        o = torch.full((12, nq, 32, 128), float("nan"), dtype=torch.float16, device=dev)
        for i in range(12):                     # 12 calls in flight on one stream, 4 layer pools cycled
            l = i % 4
            deft_b200.tree_attention_subtree_fwd(q[l, :, : 32 * 128].view(nq, 32, 128), kvp.get_key_buffer(l),
                                                 kvp.get_value_buffer(l), o[i], 128, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
        torch.cuda.synchronize()
        return o

    try:
        a, b = run(1), run(0)
    finally:
        _lib.lib.deft_b200_set_pdl(1)
    assert torch.isfinite(a.float()).all()
    assert torch.equal(a, b)
    for i in range(4, 12):
        assert torch.equal(a[i], a[i - 4])


@pytest.mark.parametrize("subtree_ramp", [False, True])
def test_reference_maximum_is_raised_mid_chain(dev, subtree_ramp):
    """Scores that grow along the KV chain: the tile exponentials overflow the fp16 range of P against the first
    tile's maximum, so the kernel must raise its reference maximum, rescale the accumulator in TMEM and redo the
    tile (attn_umma.cu, the rare path of the softmax warps).  Compared with fp64 per-leaf attention."""
    from deft_b200 import TreeMetadata
    import deft_b200
    from deft_b200.workloads import build_tree
    torch.manual_seed(7)
    tree = build_tree("cfg2", layers=1, device=dev)
    kvp = tree.token_to_kv_pool
    kv = kvp.kv_data[0]
    kv.normal_()
    # K magnitude grows with the page id: x1 at the start of the prompt, x24 at its end and in the subtree
    pool = kv.shape[0]
    ramp = (1.0 + 23.0 * torch.clamp(torch.arange(pool, device=dev, dtype=torch.float32) / 4096.0, max=1.0)).half()
    if subtree_ramp:   # ... and keeps growing through the subtree pages: the raise then also hits MASKED tiles
        idx = torch.arange(pool, device=dev, dtype=torch.float32)
        ramp = (ramp.float() + 40.0 * torch.clamp(idx - 4096.0, min=0.0) / max(pool - 4096, 1)).half()
    kv[:, 0].mul_(ramp[:, None, None])
    K, V = kvp.get_key_buffer(0), kvp.get_value_buffer(0)
    nq = len(tree.leaves)
    q = torch.randn(nq, 32, 128, dtype=torch.float16, device=dev)
    m = TreeMetadata.from_tree_cache(tree)
    o = torch.full((nq, 32, 128), float("nan"), dtype=torch.float16, device=dev)
    deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, m.block_q, m.block_q_cnts, m.block_q_offset, m.block_bitmasks,
                                         m.block_kv, m.block_lens)
    want = torch.empty(nq, 32, 128, dtype=torch.float64, device=dev)
    for i, path in enumerate(orc.leaf_paths(tree)):
        idx = torch.as_tensor(path, device=dev)
        k = K[idx].double().repeat_interleave(4, dim=1)
        v = V[idx].double().repeat_interleave(4, dim=1)
        sc = torch.einsum("hd,nhd->hn", q[i].double(), k) / 128 ** 0.5
        want[i] = torch.einsum("hn,nhd->hd", torch.softmax(sc, dim=-1), v)
    assert torch.isfinite(o.float()).all()
    # sharply peaked softmax (a few tokens dominate): fp16 rounding of P and of the output, nothing more
    assert torch.allclose(o.double(), want, atol=4e-3, rtol=2e-2), (o.double() - want).abs().max().item()


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_cluster_pairs_share_kv_tiles_by_multicast(dev, name, monkeypatch):
    """Pair-aligned job lists (DEFT_PLAN_PAIR=1: the two slot-jobs of a unit on the two CTAs of a cluster, every
    K/V tile loaded once for both by TMA multicast -- boxes for the prompt, gather4 for the subtree) give the
    bit-identical result of the independent job lists."""
    from deft_b200 import TreeMetadata
    import deft_b200
    from deft_b200.workloads import build_tree
    torch.manual_seed(13)
    tree = build_tree(name, layers=2, device=dev)
    kvp = tree.token_to_kv_pool
    for l in range(2):
        kvp.kv_data[l].normal_()
    nq = len(tree.leaves)
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    outs = []
    for pair in ("1", "0"):
        monkeypatch.setenv("DEFT_PLAN_PAIR", pair)
        m = TreeMetadata.from_tree_cache(tree)
        o = torch.full((6, nq, 32, 128), float("nan"), dtype=torch.float16, device=dev)
        for i in range(6):                  # back to back on one workspace, two layer pools
            deft_b200.tree_attention_subtree_fwd(q, kvp.get_key_buffer(i % 2), kvp.get_value_buffer(i % 2), o[i], 128, m.block_q,
                                                 m.block_q_cnts, m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
        torch.cuda.synchronize()
        outs.append(o)
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0][0], outs[0][2]) and torch.equal(outs[0][1], outs[0][3])
