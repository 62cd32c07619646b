"""A/B on one box: a layer call over tables whose native tiles were laid out incrementally (kept order of the C++ tree
mirror, 60 decode steps) against tables laid out from scratch for the same tree.  Diagnostic, not a bench value."""
import sys

import torch

import deft_b200
from deft_b200 import TreeMetadata
from deft_b200.tree_cache import flatten_tree, mirror_flat
from deft_b200.workloads import build_tree

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
H, HKV, D, L = 32, 8, 128, 8
tree = build_tree(wl, layers=L, device=dev, headroom=300 * (steps + 4))
kvp = tree.token_to_kv_pool
for l in range(L):
    kvp.kv_data[l].normal_()
nq = len(tree.leaves)
for it in range(steps):
    for leaf in tree.leaves.values():
        leaf.append_token(7)
    upd = tree.alloc()
    m_inc = TreeMetadata.from_tree_cache(tree)          # through the mirror: appended tokens only
assert mirror_flat([tree]) is not None and tree.native_tree().syncs == 1
m_new = TreeMetadata._assemble(tree, flatten_tree(tree), 32, -1, tree_index=False)      # flat arrays: laid out from scratch
q = torch.randn(nq, H, D, dtype=torch.float16, device=dev)
outs = []
for name, m in (("kept order", m_inc), ("from scratch", m_new)):
    o = torch.empty(nq, H, D, dtype=torch.float16, device=dev)

    def call(l):
        deft_b200.tree_attention_subtree_fwd(q, kvp.get_key_buffer(l), kvp.get_value_buffer(l), o, 128, m.block_q, m.block_q_cnts,
                                             m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
    for l in range(L):
        call(l)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for rep in range(4):
            for l in range(L):
                call(l)
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (20 * 4 * L) * 1e3
    print(f"{wl} after {steps} steps, {name:13s}: {us:7.2f} us per layer call, {m.flat_plan.n_units} units, "
          f"{m.packed.numel()} table bytes")
    outs.append(o.clone())
print("max |difference| of the outputs:", (outs[0].float() - outs[1].float()).abs().max().item())
