"""Pin the CPU oracle against outputs of the reference itself (tests/golden, made by oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import deft_oracle as orc
from oracle.plain_tree import thaw
from oracle.scenarios import SCENARIOS, TABLE_SCENARIOS

TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    tree = thaw({k[5:]: z[k] for k in z.files if k.startswith("tree_")})
    return z, tree


@pytest.mark.parametrize("name", list(SCENARIOS) + list(TABLE_SCENARIOS))
def test_tables_bit_exact(golden_dir, name):
    z, tree = load(golden_dir, name)
    for prefix, mbl in (("t_", -1), ("tc_", 128)):
        t = orc.build_tables(tree, max_block_len=mbl)
        for k in TABLE_KEYS:
            assert np.array_equal(t[k], z[prefix + k]), (name, prefix, k)
        q_num, node_num, total, blen = z[prefix + "scalars"].tolist()
        assert (t["query_num"], t["node_num"], t["total_kv_len"], t["block_len"]) == (q_num, node_num, total, blen)
        assert sorted(t["leaf_to_q"].items()) == [tuple(r) for r in z[prefix + "leaf_to_q"].tolist()]


def _inputs(z):
    H, HKV, D = z["geom"][:3].tolist()
    pool = z["kv_pool"]
    return z["q"], pool[:, 0], pool[:, 1]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_flatten_stage1_matches_reference_partials(golden_dir, name):
    z, tree = load(golden_dir, name)
    q, K, V = _inputs(z)
    t = {k: z["t_" + k] for k in TABLE_KEYS}
    po, pl = orc.flatten_stage1(q, K, V, t)
    np.testing.assert_allclose(po, z["flatten_partial_o"], atol=2e-6, rtol=2e-5)
    np.testing.assert_allclose(pl, z["flatten_partial_lse"], atol=5e-6, rtol=1e-6)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_operator_outputs_match_reference(golden_dir, name):
    """Faithful-arithmetic restatement vs the reference Triton operators (interpreter run).

    The reference accumulates the output with fp16 atomics in unspecified order; the restatement
    uses slot order, so agreement is to a few fp16 ulps of the output scale, far inside the
    north-star tolerance (atol 1e-3, rtol 1e-2) asserted here too.
    """
    z, tree = load(golden_dir, name)
    q, K, V = _inputs(z)
    t = {k: z["t_" + k] for k in TABLE_KEYS}
    tc = {k: z["tc_" + k] for k in TABLE_KEYS}
    for got, want in ((orc.flatten_fwd(q, K, V, t), z["o_flatten"]),
                      (orc.node_fwd(q, K, V, t), z["o_node"]),
                      (orc.node_fwd(q, K, V, tc), z["o_node_chunk"])):
        g, w = got.astype(np.float32), want.astype(np.float32)
        assert np.allclose(g, w, atol=1e-3, rtol=1e-2)
        assert np.abs(g - w).max() <= 2.5e-4


def test_tree_index_mode(golden_dir):
    z, tree = load(golden_dir, "tree_index")
    q, K, V = _inputs(z)
    max_ctx = int(z["geom"][4])
    t = orc.build_tables_tree_index(tree, max_ctx, max_block_len=128)
    for k in ["node_q", "node_q_len", "node_q_offset", "node_kv_offset", "node_kv_len"]:
        assert np.array_equal(t[k], z["ti_" + k]), k
    got = orc.node_fwd(q, K, V, t, node_kv=z["node_to_kv"].reshape(-1)).astype(np.float32)
    assert np.abs(got - z["o_tree_index"].astype(np.float32)).max() <= 2.5e-4


@pytest.mark.parametrize("name", [n for n in SCENARIOS if n != "spec_merge"])
def test_exact_and_seq_semantics(golden_dir, name):
    """fp64 per-leaf attention == what every reference operator computes, to the reference's own error."""
    z, tree = load(golden_dir, name)
    q, K, V = _inputs(z)
    paths = orc.leaf_paths(tree)
    # the per-sequence page table the reference's Radix baseline reads gives the same paths
    for i, p in enumerate(paths):
        row = z["req_to_token"][z["req_idx"][i], : z["seq_lens"][i]]
        assert np.array_equal(np.sort(row), np.sort(p)) and len(p) == z["seq_lens"][i]
    exact = orc.exact_attention(q, K, V, paths)
    for key in ("o_flatten", "o_node", "o_node_chunk", "o_seq"):
        assert np.abs(z[key].astype(np.float64) - exact).max() < 6e-4, key
    assert np.abs(orc.seq_attention(q, K, V, paths).astype(np.float64) - exact).max() < 3e-4
    # the fp32-merge variant (what the CUDA stage 2 does) is at least as close to exact as the reference
    t = {k: z["t_" + k] for k in TABLE_KEYS}
    mine = np.abs(orc.flatten_fwd(q, K, V, t, faithful_fp16=False).astype(np.float64) - exact).max()
    assert mine <= np.abs(z["o_flatten"].astype(np.float64) - exact).max() + 1e-4


@pytest.mark.parametrize("name", ["toy_binary", "llama_flat8"])
def test_torch_cpu_baseline_matches_oracle(golden_dir, name):
    """The multi-threaded CPU baseline bench.py times computes what oracle.seq_attention computes."""
    import torch
    from oracle.seq_cpu import seq_attention_torch
    z, tree = load(golden_dir, name)
    paths = orc.leaf_paths(tree)
    got = seq_attention_torch(torch.from_numpy(z["q"]), torch.from_numpy(z["kv_pool"]), paths).float().numpy()
    want = orc.seq_attention(z["q"], z["kv_pool"][:, 0], z["kv_pool"][:, 1], paths).astype(np.float32)
    assert np.abs(got - want).max() <= 1e-3
    assert np.abs(got - z["o_seq"].astype(np.float32)).max() <= 1e-3
