#!/usr/bin/env python
"""The reference's Triton tree attention and ours, on the SAME B200, on the SAME tensors, in one process.

    python tools/install_reference.py                 # build container: reference -> baseline/_ref (ships with gpurun)
    gpurun -- python tools/ref_triton_probe.py [cfg2 cfg3 ...]

For every workload (trees are grown through the REFERENCE ``TreeCache``, tables by the REFERENCE
``TreeMetadata.from_tree_cache``):

* parity on the GPU: ours (plain drop-in on the reference's tables, and with the native plan of our builder) against
  the reference operator's output (fp16 atol=1e-3 rtol=1e-2, BASELINE.json) and both against an fp64 per-leaf
  oracle; our tables against the reference's, bit for bit; Flatten, Node, Node-Chunk and the sequence-based (Radix)
  operator;
* timing, SURVEY.md 8(d) protocol: CUDA events around back-to-back calls cycling through 8 layer pools (> L2),
  20 warm-ups, median over batches; the reference call includes its ``torch.zeros_like(q)`` (``deft_attention.py:120``:
  its operator needs a zeroed output); ours eager through the same Python signature, and as a CUDA graph;
* host time of ``from_tree_cache`` (reference Python builder vs. our C++ builder on the same tree);
* the drop-in: ``deft_b200.install.install()`` rebinding the names inside the reference's ``deft_attention`` module,
  ``DeFTAttention.deft_flatten_forward`` / ``deft_node_forward`` called end to end, patched vs. unpatched.

Writes ``gpurun_out/ref_probe.json``.
"""
from __future__ import annotations

import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "DeFT")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402

NP = 8          # layer pools cycled (8 x 25 MB > 126 MB L2 at cfg2)
H, HKV, D = 32, 8, 128


def check_manifest():
    man = json.load(open(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")))
    bad = [f for f, h in man["files"].items()
           if hashlib.sha256(open(os.path.join(REF, "deft", f), "rb").read()).hexdigest() != h]
    return {"files": len(man["files"]), "modified": bad}


class Clocks:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                   "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=lambda: [self.rows.append(l.split(",")) for l in self.p.stdout], daemon=True).start()

    def stop(self):
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz_median": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": sorted({r[2].strip() for r in self.rows if len(r) > 2}), "samples": len(sm)}


def timed(fn, warm=20, batches=5, per_batch=40):
    """fn(i) is one call on pool i % NP.  Median over batches of the mean call time (us)."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    out = []
    for b in range(batches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(per_batch):
            fn(b * per_batch + i)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / per_batch)
    return statistics.median(out)


def graphed(fn, replays=25):
    """The same NP calls as ONE CUDA graph (launch overhead of the host removed); us per call."""
    for i in range(NP):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NP):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    out = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / (replays * NP))
    return statistics.median(out)


def oracle64(q, K, V, tree):
    """fp64 per-leaf attention over req_to_token paths (the pattern of tests/model/test_DeFT_kernel.py:212-276)."""
    leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
    out = torch.empty(q.shape, dtype=torch.float64, device=q.device)
    G = H // HKV
    for i, leaf in enumerate(leaves):
        pages = []
        n = leaf
        while n is not None:
            pages = list(n.kv_indices) + pages
            n = n.parent
        idx = torch.tensor(pages, device=q.device)
        k = K[idx].double().repeat_interleave(G, dim=1)      # [T, H, D]
        v = V[idx].double().repeat_interleave(G, dim=1)
        s = torch.einsum("hd,thd->ht", q[i].double(), k) / (D ** 0.5)
        out[i] = torch.einsum("ht,thd->hd", torch.softmax(s, dim=-1), v)
    return out


def main():
    import deft_b200
    from deft_b200 import install as dinstall
    from deft_b200.workloads import WORKLOADS, replay, unique_kv_tokens, n_leaves, max_path_len, n_nodes, algorithmic_bytes
    from deft.memory_pool import ReqToTokenPool, TokenToKVPool
    from deft.tree_decoding import tree_cache as rtc
    from deft.layers.attention import tree_attention as ta
    from deft.layers.attention import token_attention as tk

    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4"]
    dev = torch.device("cuda:0")
    res = {"manifest": check_manifest(), "triton": __import__("triton").__version__, "torch": torch.__version__,
           "gpu": torch.cuda.get_device_name(0), "pools_cycled": NP, "workloads": {}}
    try:
        import deft.layers.attention.deft_attention as rda   # the caller module (needs a GPU at import)
        res["deft_attention_import"] = "ok"
    except Exception as e:   # noqa: BLE001
        rda = None
        res["deft_attention_import"] = f"failed: {type(e).__name__}: {e}"
    ref_sub, ref_node, ref_tok = ta.tree_attention_subtree_fwd, ta.tree_attention_fwd, tk.token_attention_fwd

    for name in names:
        torch.manual_seed(0)
        script = WORKLOADS[name][0]
        nq, uniq = n_leaves(name), unique_kv_tokens(name)
        r2t = ReqToTokenPool(size=max(2 * n_nodes(name), 8), max_context_len=max_path_len(name) + 8)
        kvp = TokenToKVPool(size=uniq + 64, dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=NP)
        tree = rtc.TreeCache(torch.float16, HKV, D, NP, r2t, kvp, None, True, False)
        replay(tree, script)
        for l in range(NP):
            kvp.kv_data[l].normal_()
        qkv = torch.randn(NP, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
        qv = [qkv[l, :, : H * D].view(nq, H, D) for l in range(NP)]           # strided views, row stride 6144
        Ks = [kvp.get_key_buffer(l) for l in range(NP)]
        Vs = [kvp.get_value_buffer(l) for l in range(NP)]
        w = {"nq": nq, "unique_kv": uniq, "algorithmic_bytes": algorithmic_bytes(name)}

        # ---- tables: reference Python builder vs our C++ builder on the same (reference) tree
        def host_ms(fn, reps=5):
            ts = []
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                m = fn()
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            return statistics.median(ts), m
        w["from_tree_cache_ms_reference"], m_ref = host_ms(lambda: rtc.TreeMetadata.from_tree_cache(tree))
        w["from_tree_cache_ms_ours"], m_our = host_ms(lambda: deft_b200.TreeMetadata.from_tree_cache(tree))
        keys = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset", "block_q",
                "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]
        w["tables_bit_exact"] = all(torch.equal(getattr(m_ref, k), getattr(m_our, k)) for k in keys) and \
            (m_ref.query_num, m_ref.node_num, m_ref.total_kv_len) == (m_our.query_num, m_our.node_num, m_our.total_kv_len)
        rtc.BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
        deft_b200.BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
        mc_ref = rtc.TreeMetadata.from_tree_cache(tree)
        mc_our = deft_b200.TreeMetadata.from_tree_cache(tree)
        rtc.BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
        deft_b200.BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
        w["tables_bit_exact_node_chunk"] = all(torch.equal(getattr(mc_ref, k), getattr(mc_our, k)) for k in keys[:6])

        o64 = oracle64(qv[0], Ks[0], Vs[0], tree)

        def flat(fn, m):
            return lambda i, o=None: fn(qv[i % NP], Ks[i % NP], Vs[i % NP], o if o is not None else torch.zeros_like(qv[i % NP]),
                                        m.block_len, m.block_q, m.block_q_cnts, m.block_q_offset, m.block_bitmasks,
                                        m.block_kv, m.block_lens)

        def node(fn, m):
            return lambda i, o=None: fn(qv[i % NP], Ks[i % NP], Vs[i % NP], o if o is not None else torch.zeros_like(qv[i % NP]),
                                        m.node_kv, m.node_kv_offset, m.node_kv_len, m.node_q, m.node_q_offset, m.node_q_len)

        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        req_idx = torch.tensor([tree.leaf_to_req[l.id] for l in leaves], dtype=torch.int32, device=dev)
        seq_host = []
        for leaf in leaves:
            n, s = leaf, 0
            while n is not None:
                s += len(n.kv_indices)
                n = n.parent
            seq_host.append(s)
        seq_lens = torch.tensor(seq_host, dtype=torch.int32, device=dev)
        start_loc = torch.zeros_like(seq_lens)
        start_loc[1:] = torch.cumsum(seq_lens[:-1], dim=0)

        def seq(fn):
            return lambda i, o=None: fn(qv[i % NP], Ks[i % NP], Vs[i % NP], o if o is not None else torch.zeros_like(qv[i % NP]),
                                        r2t.req_to_token, req_idx, start_loc, seq_lens, max(seq_host), None, sum(seq_host))

        ops = {
            "flatten": (flat(ref_sub, m_ref), flat(deft_b200.tree_attention_subtree_fwd, m_ref),
                        flat(deft_b200.tree_attention_subtree_fwd, m_our)),
            "node": (node(ref_node, m_ref), node(deft_b200.tree_attention_fwd, m_ref), node(deft_b200.tree_attention_fwd, m_our)),
            "node_chunk": (node(ref_node, mc_ref), node(deft_b200.tree_attention_fwd, mc_ref),
                           node(deft_b200.tree_attention_fwd, mc_our)),
            "seq": (seq(ref_tok), seq(deft_b200.token_attention_fwd), None),
        }
        if name == "cfg4":
            ops.pop("node")          # the reference's Node kernel walks the 8192-token root serially: minutes at 256 leaves
        clocks = Clocks()
        for mode, (f_ref, f_drop, f_native) in ops.items():
            r = {}
            outs = {}
            for tag, f in (("reference", f_ref), ("ours_dropin", f_drop), ("ours_native", f_native)):
                if f is None:
                    continue
                try:
                    o = torch.zeros_like(qv[0]) if tag == "reference" else torch.full_like(qv[0].contiguous(), float("nan"))
                    f(0, o)
                    torch.cuda.synchronize()
                    outs[tag] = o
                    r[tag + "_max_err_vs_fp64"] = float((o.double() - o64).abs().max())
                    r[tag + "_us_eager"] = timed(f)
                    try:
                        r[tag + "_us_graph"] = graphed(f)
                    except Exception as e:   # noqa: BLE001
                        r[tag + "_us_graph"] = None
                        r[tag + "_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
                        torch.cuda.synchronize()
                except Exception as e:   # noqa: BLE001
                    r[tag + "_error"] = f"{type(e).__name__}: {str(e)[:300]}"
            if "reference" in outs:
                for tag in ("ours_dropin", "ours_native"):
                    if tag in outs:
                        r[tag + "_allclose_reference"] = bool(torch.allclose(outs[tag].float(), outs["reference"].float(),
                                                                              atol=1e-3, rtol=1e-2))
                        r[tag + "_max_diff_reference"] = float((outs[tag].float() - outs["reference"].float()).abs().max())
                        r[tag + "_at_least_as_close_to_fp64"] = \
                            r[tag + "_max_err_vs_fp64"] <= r["reference_max_err_vs_fp64"] + 1e-4
                ours = r.get("ours_native_us_graph") or r.get("ours_dropin_us_graph")
                if ours and r.get("reference_us_eager"):
                    r["speedup_vs_reference_eager"] = r["reference_us_eager"] / ours
                    if r.get("reference_us_graph"):
                        r["speedup_vs_reference_graph"] = r["reference_us_graph"] / ours
            w[mode] = r
            print(name, mode, json.dumps(r), flush=True)
        w["clocks"] = clocks.stop()

        # ---- the drop-in: DeFTAttention.deft_*_forward through the reference's own module, patched vs unpatched
        if rda is not None:
            try:
                att = rda.DeFTAttention(H, D, D ** -0.5, HKV, layer_id=0)
                k_new = qkv[0, :, H * D: (H + HKV) * D]
                v_new = qkv[0, :, (H + HKV) * D:]
                loc = torch.tensor([l.kv_indices[-1] for l in leaves], dtype=torch.int32, device=dev)
                upd = rtc.KVCacheUpdater(use_paged_memory=True, token_to_kv_pool=kvp, cache_loc=loc.long(),
                                         leaf_data=None, is_prompt=False)
                imeta = types.SimpleNamespace(kv_updater=upd, token_to_kv_pool=kvp, req_to_token_pool=r2t)
                q2d = qkv[0, :, : H * D]
                d = {}
                rtc.register_tree_metadata(rtc.TreeMetadata.from_tree_cache(tree))
                o_flat_ref = att.deft_flatten_forward(q2d, k_new, v_new, imeta).clone()
                o_node_ref = att.deft_node_forward(q2d, k_new, v_new, imeta).clone()
                dinstall.install(metadata=True)
                d["names_rebound"] = (rda.tree_attention_subtree_fwd is deft_b200.tree_attention_subtree_fwd
                                      and rda.tree_attention_fwd is deft_b200.tree_attention_fwd
                                      and ta.tree_attention_subtree_fwd is deft_b200.tree_attention_subtree_fwd)
                m_patched = rtc.TreeMetadata.from_tree_cache(tree)
                d["patched_metadata_is_ours"] = isinstance(m_patched, deft_b200.TreeMetadata)
                rtc.register_tree_metadata(m_patched)
                o_flat = att.deft_flatten_forward(q2d, k_new, v_new, imeta)
                o_node = att.deft_node_forward(q2d, k_new, v_new, imeta)
                dinstall.uninstall()
                d["names_restored"] = rda.tree_attention_subtree_fwd is ref_sub and ta.tree_attention_fwd is ref_node
                rtc.unregister_tree_metadata()
                d["flatten_allclose"] = bool(torch.allclose(o_flat.float(), o_flat_ref.float(), atol=1e-3, rtol=1e-2))
                d["node_allclose"] = bool(torch.allclose(o_node.float(), o_node_ref.float(), atol=1e-3, rtol=1e-2))
                d["flatten_max_diff"] = float((o_flat.float() - o_flat_ref.float()).abs().max())
                w["dropin_deft_attention"] = d
            except Exception as e:   # noqa: BLE001
                import traceback
                w["dropin_deft_attention"] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
                dinstall.uninstall()
            print(name, "dropin", json.dumps(w.get("dropin_deft_attention")), flush=True)
        res["workloads"][name] = w
        del kvp, tree, qkv, qv, Ks, Vs
        torch.cuda.empty_cache()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "ref_probe.json")
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
