#!/usr/bin/env python
"""Per-CTA timeline of the tcgen05 stage-1 kernel on a BASELINE workload (profiling aid, needs a B200).

    python tools/trace_stage1.py [cfg2] [n_ctas_to_print]

Uses the deft_b200_set_trace_buffer hook: every CTA records SM-cycle timestamps of its first job
(event ids: csrc/attn_umma.cu kTr*).  Prints the events in microseconds at the nominal 1.965 GHz.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import deft_b200
from deft_b200 import TreeMetadata, _lib
from deft_b200.workloads import build_tree

NAMES = {0: "start", 1: "q_ids", 2: "q_issued", 4: "mask0", 5: "k_unit", 6: "mma_q_full", 7: "epi_begin",
         8: "epi_end", 9: "end", 10: "epi_o_done", 12: "k_role_entry"}
TILE = ["k_issued", "s_issued", "sm_s_in_regs", "sm_checked", "sm_p0_handed", "sm_p1_handed", "v_issued", "pv_issued"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    show = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda:0")
    _lib.lib.deft_b200_set_experiment(int(os.environ.get("DEFT_EXPERIMENT", "0")))
    if os.environ.get("DEFT_NO_GATHER4"):
        _lib.lib.deft_b200_set_gather4(0)
    n_trees = int(os.environ.get("TRACE_TREES", "1"))
    if n_trees > 1:
        from deft_b200.workloads import build_forest
        trees = build_forest(wl, n_trees, layers=2, device=dev)
        tree = trees[0]
    else:
        tree = build_tree(wl, layers=2, device=dev)
    kvp = tree.token_to_kv_pool
    for l in range(2):
        kvp.kv_data[l].normal_()
    nq = len(tree.leaves) * n_trees
    q = torch.randn(nq, 48 * 128, dtype=torch.float16, device=dev)[:, : 32 * 128].view(nq, 32, 128)
    o = torch.empty(nq, 32, 128, dtype=torch.float16, device=dev)
    m = TreeMetadata.from_tree_cache(tree) if n_trees == 1 else TreeMetadata.from_forest(trees)
    trace = torch.full((256, 128), -1, dtype=torch.int32, device=dev)

    def call(layer):
        deft_b200.tree_attention_subtree_fwd(q, kvp.get_key_buffer(layer), kvp.get_value_buffer(layer), o, 128, m.block_q,
                                             m.block_q_cnts, m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)

    for _ in range(3):
        call(0)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if not os.environ.get("TRACE_NOFLUSH"):
        flush.zero_()                           # push the KV pool out of L2
    call(1)                                     # plan tables warm in L2 (as inside a decode step: 32 layers share
    torch.cuda.synchronize()                    # them), layer 0's KV stays cold
    _lib.lib.deft_b200_set_trace_buffer(trace.data_ptr())
    call(0)
    torch.cuda.synchronize()
    _lib.lib.deft_b200_set_trace_buffer(None)
    t = trace.cpu().numpy()
    ghz = 1.965
    ends = t[:, 9]
    active = [c for c in range(t.shape[0]) if ends[c] >= 0]
    print(f"{wl}: {len(active)} CTAs traced; end (us): min {ends[active].min() / ghz / 1e3:.2f} "
          f"max {ends[active].max() / ghz / 1e3:.2f} mean {ends[active].mean() / ghz / 1e3:.2f}")
    if os.environ.get("TRACE_JOBS"):       # per job of a few CTAs: duration, tiles, kind -> us per tile by kind
        import collections
        agg = collections.defaultdict(lambda: [0.0, 0, 0])
        for c in active:
            prev = t[c, 16 + 2] if t[c, 16 + 2] >= 0 else t[c, 0]      # first S in registers
            for j in range(32):
                e, info = int(t[c, 64 + 2 * j]), int(t[c, 65 + 2 * j])
                if e < 0:
                    break
                n, box, shared, dense = info & 0xffff, (info >> 16) & 1, (info >> 17) & 1, (info >> 18) & 1
                if j > 0:
                    key = ("box" if box else "gathered/mixed", "shared" if shared else "own", "dense" if dense else "masked")
                    agg[key][0] += (e - prev) / ghz / 1e3
                    agg[key][1] += n
                    agg[key][2] += 1
                if c in active[:3]:
                    print(f"cta {c} job {j}: {n:3d} tiles box={box} shared={shared} dense={dense}  {(e - prev) / ghz / 1e3:7.2f} us  = {(e - prev) / ghz / 1e3 / n:5.2f} us/tile")
                prev = e
        for key, (us, tiles, jobs) in sorted(agg.items()):
            print(f"{key}: {jobs} jobs, {tiles} tiles, {us / tiles:.3f} us per tile (job boundary included), {tiles / jobs:.1f} tiles per job")
        return
    if os.environ.get("TRACE_TABLE"):      # one line per CTA: first S issued, tiles seen (of the first job, <= 6), epilogue, end
        for c in sorted(active, key=lambda c: -ends[c]):
            s0 = t[c, 16 + 1]
            nt = sum(1 for tile in range(6) if t[c, 16 + 8 * tile + 5] >= 0)
            last_p = max([t[c, 16 + 8 * tile + 5] for tile in range(6)] + [-1])
            print(f"cta {c:3d}  S0 {s0 / ghz / 1e3:6.2f}  tiles>={nt}  last_p {last_p / ghz / 1e3:6.2f}  epi {t[c, 7] / ghz / 1e3:6.2f}..{t[c, 8] / ghz / 1e3:6.2f}  end {ends[c] / ghz / 1e3:6.2f}")
        return
    order = sorted(active, key=lambda c: -ends[c])
    for c in order[:show] + [order[len(order) // 2]] + order[-1:]:
        print(f"--- CTA {c}")
        ev = [(int(t[c, i]), NAMES[i]) for i in NAMES if t[c, i] >= 0]
        for tile in range(6):
            for k in range(8):
                v = int(t[c, 16 + 8 * tile + k])
                if v >= 0:
                    ev.append((v, f"t{tile}.{TILE[k]}"))
        for v, name in sorted(ev):
            print(f"   {v / ghz / 1e3:8.2f} us  {name}")


if __name__ == "__main__":
    main()
