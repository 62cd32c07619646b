"""The drop-in: ``deft_b200.install`` rebinds the reference's names, on the REFERENCE's own modules.  CPU only.

Needs the reference checkout (``/root/reference``, or the copy ``tools/install_reference.py`` makes under
``baseline/_ref``); skipped where neither exists.  The reference hard-codes ``device="cuda"`` in its torch factory
calls: the shim of SURVEY.md Appendix C strips that keyword for the duration of a test (monkeypatch restores it).
The GPU leg of the same check (``DeFTAttention.deft_flatten_forward`` end to end, patched vs. the unpatched Triton
operators) is ``tools/ref_triton_probe.py`` on the B200 box; its result is committed under ``profiles/``.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle.scenarios import SCENARIOS, TABLE_SCENARIOS, replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = ["/root/reference/DeFT", os.path.join(ROOT, "baseline", "_ref", "DeFT")]
TABLE_KEYS = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
              "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


@pytest.fixture()
def reference(monkeypatch):
    """The reference's tree_cache / tree_attention modules, importable on CPU for the length of one test."""
    ref = next((d for d in REF_DIRS if os.path.isdir(os.path.join(d, "deft"))), None)
    if ref is None:
        pytest.skip("no reference checkout here")
    for name in ["tensor", "empty", "ones", "zeros", "full", "arange"]:
        f = getattr(torch, name)
        monkeypatch.setattr(torch, name, (lambda f: lambda *a, **k: f(*a, **{kk: v for kk, v in k.items()
                                                                            if not (kk == "device" and v == "cuda")}))(f))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.syspath_prepend(ref)
    before = set(sys.modules)
    import deft.memory_pool as mp
    import deft.tree_decoding.tree_cache as tc
    import deft.layers.attention.tree_attention as ta
    yield mp, tc, ta
    import deft_b200.install as inst
    inst.uninstall()
    for name in set(sys.modules) - before:
        if name == "deft" or name.startswith("deft."):
            del sys.modules[name]


def _ref_tree(mp, tc, cfg):
    HKV, D = cfg.get("HKV", 1), cfg.get("D", 16)
    r2t = mp.ReqToTokenPool(size=128, max_context_len=cfg["max_ctx"])
    kvp = mp.TokenToKVPool(size=cfg["pool"], dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=1)
    r2t.req_to_token.zero_()
    tree = tc.TreeCache(torch.float16, HKV, D, 1, r2t, kvp, None, True, False)
    replay(tree, cfg["script"], lambda n: torch.arange(1, n + 1, dtype=torch.int32))
    return tree


@pytest.mark.parametrize("name", ["toy_binary", "wide40", "ragged_cut", "spec_merge", "cfg3a_tables"])
def test_install_swaps_the_reference_names_and_tables_stay_bit_exact(reference, golden_dir, name):
    import deft_b200
    import deft_b200.install as inst
    mp, tc, ta = reference
    ref_sub, ref_fwd = ta.tree_attention_subtree_fwd, ta.tree_attention_fwd
    ref_builder = tc.TreeMetadata.__dict__["from_tree_cache"]
    cfg = {**SCENARIOS, **TABLE_SCENARIOS}[name]
    tree = _ref_tree(mp, tc, cfg)                      # grown by the REFERENCE TreeCache
    m_ref = tc.TreeMetadata.from_tree_cache(tree)      # the reference's Python builder
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))

    inst.install(metadata=True)
    assert ta.tree_attention_subtree_fwd is deft_b200.tree_attention_subtree_fwd
    assert ta.tree_attention_fwd is deft_b200.tree_attention_fwd
    m = tc.TreeMetadata.from_tree_cache(tree)          # the patched name: our C++ builder on the reference's tree
    assert isinstance(m, deft_b200.TreeMetadata) and m.flat_plan is not None
    for k in TABLE_KEYS:
        assert np.array_equal(getattr(m, k).numpy(), getattr(m_ref, k).numpy()), (name, k)
        assert np.array_equal(getattr(m, k).numpy(), z["t_" + k]), (name, k, "golden")
    assert (m.query_num, m.node_num, m.total_kv_len, m.block_len) == (m_ref.query_num, m_ref.node_num, m_ref.total_kv_len, m_ref.block_len)
    assert m.leaf_to_q == m_ref.leaf_to_q
    # the CLI selects node_chunk by mutating the REFERENCE module's BLOCK_CONFIG (run_DeFT_llama_paged.py:145-147)
    tc.BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        mc = tc.TreeMetadata.from_tree_cache(tree)
    finally:
        tc.BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
        deft_b200.BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    for k in TABLE_KEYS[:6]:
        assert np.array_equal(mc.__dict__[k].numpy(), z["tc_" + k]), (name, k, "node_chunk")

    inst.uninstall()
    assert ta.tree_attention_subtree_fwd is ref_sub and ta.tree_attention_fwd is ref_fwd
    assert tc.TreeMetadata.__dict__["from_tree_cache"] is ref_builder
    again = tc.TreeMetadata.from_tree_cache(tree)
    assert type(again) is tc.TreeMetadata and torch.equal(again.block_kv, m_ref.block_kv)


def test_install_twice_and_uninstall_restores_the_originals(reference):
    import deft_b200.install as inst
    mp, tc, ta = reference
    orig = ta.tree_attention_subtree_fwd
    inst.install(metadata=False)
    inst.install(metadata=False)
    inst.uninstall()
    assert ta.tree_attention_subtree_fwd is orig
