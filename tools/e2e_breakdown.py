#!/usr/bin/env python
"""Where one end-to-end decode step (bench.py's e2e leg, cfg2) spends its time: host stamps and CUDA events (profiling aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import deft_b200
from deft_b200.workloads import build_tree

dev = torch.device("cuda:0")
L, H, HKV, D, CH = 32, 32, 8, 128, int(os.environ.get("CH", "8"))
tree = build_tree("cfg2", layers=L, device=dev, headroom=64 * 64)
kvp = tree.token_to_kv_pool
for l in range(L):
    kvp.kv_data[l].normal_()
nq = len(tree.leaves)
host_qkv = torch.randn(L, nq, (H + 2 * HKV) * D, dtype=torch.float16).pin_memory()
host_out = torch.empty(L, nq, H, D, dtype=torch.float16).pin_memory()
dev_qkv = torch.empty(L, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
out = torch.empty(L, nq, H, D, dtype=torch.float16, device=dev)
leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
host_loc = torch.tensor([leaf.kv_indices[-1] for leaf in leaves], dtype=torch.int32).pin_memory()
loc_dev = torch.zeros(nq, dtype=torch.int32, device=dev)
main = torch.cuda.current_stream()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
NC = L // CH
ev_in = [torch.cuda.Event() for _ in range(NC)]
ev_out = [torch.cuda.Event() for _ in range(NC)]
step = deft_b200.DecodeStepGraph(kvp, dev_qkv, out, loc_dev, H, HKV, D, mode="flatten", chunk=CH)
T = lambda: torch.cuda.Event(enable_timing=True)

def one(record):
    stamps = {}
    g = {k: T() for k in ("begin", "tables_up", "c0_start", "compute_end", "end")}
    gc = [T() for _ in range(NC)]
    t0 = time.perf_counter()
    g["begin"].record(main)
    s_in.wait_stream(main)
    def upload(chunks):
        with torch.cuda.stream(s_in):
            for c in chunks:
                dev_qkv[c * CH:(c + 1) * CH].copy_(host_qkv[c * CH:(c + 1) * CH], non_blocking=True)
                ev_in[c].record(s_in)
    upload(range(1) if not os.environ.get("ALL_FIRST") else range(NC))
    stamps["h2d_enqueued"] = time.perf_counter() - t0
    if not os.environ.get("STATIC"):        # a real decode step: a token and a page per leaf first (tree_generate.py:109)
        for leaf in tree.leaves.values():
            leaf.append_token(7)
        host_loc.copy_(tree.alloc().cache_loc)
        stamps["alloc_done"] = time.perf_counter() - t0
    m = step.metadata(tree, cache_loc=None if os.environ.get("STATIC") else host_loc)
    stamps["metadata_done"] = time.perf_counter() - t0
    g["tables_up"].record(main)
    loc_dev.copy_(host_loc, non_blocking=True)

    def before(c):
        main.wait_event(ev_in[c])
        if c == 0:
            g["c0_start"].record(main)

    def after(c):
        if c == 0 and not os.environ.get("ALL_FIRST"):
            upload(range(1, NC))       # behind the tables on the H2D engine, enqueued once chunk 0 is launched
        gc[c].record(main)
        ev_out[c].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_out[c])
            host_out[c * CH:(c + 1) * CH].copy_(out[c * CH:(c + 1) * CH], non_blocking=True)
        stamps[f"chunk{c}_enqueued"] = time.perf_counter() - t0

    step.run(m, before_chunk=before, after_chunk=after)
    g["compute_end"].record(main)
    main.wait_stream(s_out)
    g["end"].record(main)
    main.synchronize()
    stamps["synced"] = time.perf_counter() - t0
    if record:
        print("host (ms since step begin): " + ", ".join(f"{k} {v * 1e3:.3f}" for k, v in stamps.items()))
        b = g["begin"]
        print("gpu  (ms since step begin): tables_up %.3f, chunk0 starts %.3f, " % (b.elapsed_time(g["tables_up"]), b.elapsed_time(g["c0_start"]))
              + ", ".join(f"chunk{c} done {b.elapsed_time(gc[c]):.3f}" for c in range(NC))
              + ", last D2H done %.3f" % b.elapsed_time(g["end"]))

if os.environ.get("META_PARTS"):     # where step.metadata() spends its time inside a step
    from deft_b200 import tree_cache as tc
    acc = {}

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return r
        return w
    tc.flatten_tree = timed("flatten_tree", tc.flatten_tree)
    tc.build_tables_host = timed("build_tables_host (incl. reserve + copy into pinned)", tc.build_tables_host)
    tc._STAGING.reserve = timed("staging.reserve", tc._STAGING.reserve)
    tc._STAGING.send = timed("staging.send (cudaMemcpyAsync + event)", tc._STAGING.send)
    tc.register_plan = timed("register_plan", tc.register_plan)
    step.metadata = timed("step.metadata total", step.metadata)

for i in range(10):
    if i == 5 and os.environ.get("META_PARTS"):
        acc.clear()                   # the first steps allocate the pinned staging ring and capture the graphs
    one(i >= 7)
if os.environ.get("META_PARTS"):
    print("per step (us): " + ", ".join(f"{k} {v / 5 * 1e6:.0f}" for k, v in acc.items()))
