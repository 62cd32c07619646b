#!/usr/bin/env python
"""Decode-step tree-attention benchmark (BASELINE.json metric) for deft_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A *step* is one decode step of the tree-attention path on a Llama-3-8B-geometry synthetic tree:
32 layer-calls of DeFT-Flatten attention (stage 1 + stage 2), each layer over its OWN KV pool, so
that the 32 pools (cfg2: 0.8 GB) cycle far beyond the 126 MB L2 between consecutive calls.

* ``value``      tokens/s with everything resident in HBM; the step is one CUDA graph of the 64
                 launches, timed with CUDA events on the launching stream, max over ranks.
* ``e2e``        a real decode loop through the public Python API with HOST buffers: every step appends one token
                 and one page per leaf (``TreeCache.alloc``, the reference's tree_generate.py:109), rebuilds the
                 tables (C++ builder, one upload), copies the step's fused qkv activations up from pinned memory,
                 appends K/V, attends (32 layers) and reads the outputs back.  The tree GROWS from step to step.
* ``roofline``   stage-1 kernel alone (graph of 32 stage-1 launches): algorithmic bytes (or flops, when the workload's
                 arithmetic intensity is past the machine balance: cfg4) / duration against MEASURED_PEAKS.json.
* ``cfg5``       BASELINE configs[4]: 512 independent cfg2 trees sharded over the N ranks (512/N per GPU, ONE launch
                 per layer per GPU): trees/s, stage-1 roofline fraction and the end-to-end leg of that regime.
* ``cpu_baseline`` / ``--impl reference``: the sequence-based (per-leaf, no prefix reuse) semantics
                 of the reference on the host cores (oracle/seq_cpu.py); a bounded sample.

With N > 1 (torchrun) every rank runs its own tree(s) (trees shard with no collective on the data
path); NCCL carries the barrier and the max-over-ranks reduction only.  scaling = weak.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

LAYERS = 32
H, HKV, D = 32, 8, 128
METRIC = "decode_attention_tokens_per_s"
UNIT = "tokens/s"
CFG5_TREES = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="flatten", choices=["flatten", "node", "node_chunk", "seq"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the 512-tree batch block (BASELINE configs[4])")
    ap.add_argument("--cfg5-trees", type=int, default=CFG5_TREES, help="trees of the cfg5 block over ALL ranks")
    ap.add_argument("--e2e-chunk", type=int, default=4, choices=[1, 2, 4, 8, 16, 32],
                    help="layers per H2D / graph / D2H chunk of the end-to-end leg (4: what is not overlapped -- the first "
                         "chunk up, the last one down -- is an eighth of the step's copies; measured 2 / 4 / 8 / 16: "
                         "cfg2 1.10 / 0.91 / 0.98 / 1.16 ms, cfg4 2.94 / 3.08 / 3.42 / - ms per step)")
    ap.add_argument("--profile-e2e", default=None, metavar="FILE",
                    help="cProfile of 50 end-to-end steps (host side) written to FILE; diagnostic, not a bench value")
    ap.add_argument("--e2e-serial", action="store_true", help="end-to-end leg without the table build of step t+1 under step t")
    ap.add_argument("--e2e-static", action="store_true", help="end-to-end leg over a tree that does NOT grow (r1 behaviour)")
    ap.add_argument("--trees-per-gpu", type=int, default=1,
                    help="independent trees of the workload batched into ONE launch per layer")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return (float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", 1364.4)),
                "measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops_sustained)")
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str, mode: str, trees: int):
    """DRAM bytes per launch of the stage-1 kernel from the newest committed `ncu --set full` capture
    (profiles/*_ncu_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches).  The
    captures are taken on the default workload (cfg2, flatten, one tree); anything else reports null."""
    if workload != "cfg2" or mode != "flatten" or trees != 1:
        return None, None
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_raw.csv")))
    if not files:
        return None, None
    unit_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        vals = [float(r[ir]) * unit_bytes[units[ir]] + float(r[iw]) * unit_bytes[units[iw]]
                for r in rows[2:] if len(r) > max(ir, iw) and "stage1" in r[0]]
        return (sum(vals) / len(vals), os.path.basename(files[-1])) if vals else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU legs: neither imports the product package (the CUDA library is not loaded in the reference arm)
# ---------------------------------------------------------------------------------------------------------
def _scripts():
    """deft_b200/workload_scripts.py by file path: the package itself would load libdeft_b200.so."""
    spec = importlib.util.spec_from_file_location("_deft_workload_scripts", os.path.join(ROOT, "deft_b200", "workload_scripts.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_paths_and_inputs(workload: str, layers: int):
    """Host copies of the synthetic inputs for the CPU baseline (same shapes, seeded)."""
    from oracle import deft_oracle as orc
    from oracle.sim_tree import SimTree
    ws = _scripts()
    tree = SimTree().replay(ws.WORKLOADS[workload][0])
    paths = orc.leaf_paths(tree)
    g = torch.Generator().manual_seed(0)
    size = tree.next_page + 64
    kv_layers = [torch.randn(size, 2, HKV, D, generator=g, dtype=torch.float32).half() for _ in range(layers)]
    q = torch.randn(len(paths), H, D, generator=g, dtype=torch.float32).half()
    return q, kv_layers, paths, ws.WORKLOADS[workload][1]


def cpu_baseline(workload: str):
    """One full step (32 layer-calls over 32 layer pools) of the sequence-based port on the host cores."""
    from oracle.seq_cpu import seq_attention_torch
    q, kv_layers, paths, _ = cpu_paths_and_inputs(workload, layers=LAYERS)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    seq_attention_torch(q, kv_layers[0], paths)          # warm-up
    t0 = time.perf_counter()
    for l in range(LAYERS):
        seq_attention_torch(q, kv_layers[l], paths)
    sec = time.perf_counter() - t0
    return {"value": len(paths) / sec, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": sec * 1e3,
            "sample": f"1 full step of {workload} = 32 layer-calls over 32 layer pools after 1 warm-up call, sequence-based "
                      f"per-leaf attention (oracle/seq_cpu.py, fp32 torch bmm, {threads} threads)"}


def run_reference(args, rank: int):
    """--impl reference: the reference's sequence-based path on the host cores, a bounded sample per step."""
    if rank != 0:
        return
    from oracle.seq_cpu import seq_attention_torch
    q, kv_layers, paths, desc = cpu_paths_and_inputs(args.workload, layers=LAYERS)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    seq_attention_torch(q, kv_layers[0], paths)
    seq_attention_torch(q, kv_layers[1], paths)
    t_layer = (time.perf_counter() - t0) / 2
    warm = min(args.warmup, 2)
    budget_s = 180.0                                     # the whole run stays within a few minutes
    per_step = max(1, min(LAYERS, int(budget_s / ((args.steps + warm) * t_layer))))
    times = []
    for s in range(warm + args.steps):
        t0 = time.perf_counter()
        for l in range(per_step):
            seq_attention_torch(q, kv_layers[(s * per_step + l) % LAYERS], paths)
        if s >= warm:
            times.append((time.perf_counter() - t0) * LAYERS / per_step)
    sec = sum(times) / len(times)
    nq = len(paths)
    value = nq / sec
    sample = (f"each step = {per_step} of the 32 layer-calls of {args.workload}"
              + ("" if per_step == LAYERS else f" (x{LAYERS}/{per_step} extrapolated)")
              + f", 32 layer pools cycled; {args.steps} timed / {warm} warm-up steps (warm-up capped from {args.warmup})")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {desc}", "layers": LAYERS,
                       "path": "sequence-based per-leaf attention on host cores (the reference has no CPU kernel and its "
                               "Flash-Decoding attention is a stub: a port of its seq semantics, oracle/seq_cpu.py)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def _path(leaf):
    nodes = []
    while leaf is not None:
        nodes.append(leaf)
        leaf = leaf.parent
    return nodes[::-1]


def measure(args, dev, rank, world, workload: str, T: int, pools: int, steps: int, warm: int, e2e_steps: int,
            e2e_chunk: int, sample_clocks: bool):
    """Device-resident step, stage-1 / stage-2 alone, and the end-to-end decode loop of ``T`` trees of ``workload`` per rank."""
    import torch.distributed as dist
    import deft_b200
    from deft_b200 import BLOCK_CONFIG, TreeMetadata, _lib
    from deft_b200.sharding import max_over_ranks
    from deft_b200.workloads import build_forest, n_leaves

    torch.manual_seed(1234 + rank)
    grow_steps = 0 if args.e2e_static else 2 * (3 + e2e_steps) + 14 + (50 if args.profile_e2e else 0)     # pipelined + serial legs
    trees = build_forest(workload, T, layers=pools, device=dev, headroom=64 + n_leaves(workload) * grow_steps)
    kvp = trees[0].token_to_kv_pool
    for l in range(pools):
        kvp.kv_data[l].normal_()
    # layer l works on pool l % pools (32 distinct pools whenever they fit: see config.l2)
    kv_view = types.SimpleNamespace(kv_data=[kvp.kv_data[l % pools] for l in range(LAYERS)],
                                    get_key_buffer=lambda l: kvp.get_key_buffer(l % pools),
                                    get_value_buffer=lambda l: kvp.get_value_buffer(l % pools), device=kvp.device)
    nq = sum(len(t.leaves) for t in trees)
    n_act = LAYERS if nq <= 8192 else pools                   # distinct activation buffers (big forests: as many as pools)

    def build_meta():
        return TreeMetadata.from_tree_cache(trees[0]) if T == 1 else TreeMetadata.from_forest(trees)

    qkv = torch.randn(n_act, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)   # fused qkv, row stride 6144
    out = torch.empty(n_act, nq, H, D, dtype=torch.float16, device=dev)
    if args.mode == "node_chunk":
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    meta = build_meta()

    def q_of(buf, l):
        return buf[l % n_act, :, : H * D].view(nq, H, D)

    if args.mode == "seq":     # Radix / sequence-based baseline ON OUR KERNELS: every leaf re-reads its whole path
        r2t = trees[0].req_to_token_pool
        sleaves = [leaf for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)]
        req_idx = torch.tensor([t.leaf_to_req[leaf.id] for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)],
                               dtype=torch.int32, device=dev)
        seq_len_host = [sum(len(n.kv_indices) for n in _path(leaf)) for leaf in sleaves]
        seq_lens = torch.tensor(seq_len_host, dtype=torch.int32, device=dev)
        start_loc = torch.zeros_like(seq_lens)
        r2t_dev = r2t.device_table() if hasattr(r2t, "device_table") else r2t.req_to_token

    def attention(l, qbuf, m, obuf):
        K, V = kv_view.get_key_buffer(l), kv_view.get_value_buffer(l)
        if args.mode == "seq":
            deft_b200.token_attention_fwd(q_of(qbuf, l), K, V, obuf[l % n_act], r2t_dev, req_idx, start_loc, seq_lens,
                                          max(seq_len_host), None, sum(seq_len_host))
        elif args.mode == "flatten":
            deft_b200.tree_attention_subtree_fwd(q_of(qbuf, l), K, V, obuf[l % n_act], m.block_len, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
        else:
            deft_b200.tree_attention_fwd(q_of(qbuf, l), K, V, obuf[l % n_act], m.node_kv, m.node_kv_offset, m.node_kv_len,
                                         m.node_q, m.node_q_offset, m.node_q_len)

    def step_resident():
        for l in range(LAYERS):
            attention(l, qkv, meta, out)

    def capture(stages):
        _lib.lib.deft_b200_set_stages(stages)
        step_resident()                                  # sizes the workspace outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step_resident()
        _lib.lib.deft_b200_set_stages(7)
        return g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_wall = []       # host wall time of every step of the last timed() call (the end-to-end steps end synchronised)

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        barrier()
        del step_wall[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            step_wall.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev) / n

    g_full, g_s1, g_s2 = capture(7), capture(2), capture(4)
    sampler = ClockSampler(dev.index) if (rank == 0 and sample_clocks) else None
    ms_step = timed(g_full.replay, steps, warm)
    clocks = sampler.stop() if sampler else None
    ms_s1 = timed(g_s1.replay, steps, warm)
    ms_s2 = timed(g_s2.replay, steps, warm)
    del g_full, g_s1, g_s2

    # ---- end to end through the public API with host buffers: a decode loop over a GROWING tree ----------
    CH = e2e_chunk
    n_chunks = LAYERS // CH
    n_host = n_chunks if nq <= 256 else 1                     # distinct pinned chunks (forests cycle ONE: the bytes copied are the same)
    host_qkv = torch.randn(n_host * CH, nq, (H + 2 * HKV) * D, dtype=torch.float16).pin_memory()
    host_out = torch.empty(n_host * CH, nq, H, D, dtype=torch.float16).pin_memory()
    dev_qkv = torch.empty(LAYERS if n_act == LAYERS else 4 * CH, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)
    dev_out = out if n_act == LAYERS else torch.empty(4 * CH, nq, H, D, dtype=torch.float16, device=dev)
    n_dev = dev_qkv.shape[0] // CH                             # device chunk buffers (a ring of 4 for big forests)
    leaves = [leaf for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)]
    host_loc = torch.tensor([leaf.kv_indices[-1] for leaf in leaves], dtype=torch.int32).pin_memory()
    table_bytes = [0]
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()      # H2D and D2H ride their own copy engines
    ev_in = [torch.cuda.Event() for _ in range(n_chunks)]
    ev_out = [torch.cuda.Event() for _ in range(n_chunks)]
    ev_free = [torch.cuda.Event() for _ in range(n_chunks)]     # chunk c's device buffers are free again (ring)
    loc_dev = torch.zeros(nq, dtype=torch.int32, device=dev)
    graphed = args.mode != "seq" and n_dev == n_chunks
    if graphed:      # the step's launches (per layer: kv_append + attention) as CUDA graphs, one per chunk of layers
        step = deft_b200.DecodeStepGraph(kv_view, dev_qkv, dev_out, loc_dev, H, HKV, D, mode=args.mode, chunk=CH)
        pipe = deft_b200.DecodeStepPipeline(kv_view, dev_qkv, dev_out, H, HKV, D, mode=args.mode, chunk=CH)

    def grow():
        """One decode step of the reference loop on the host side: a token and a page per leaf (tree_generate.py:109)."""
        locs = []
        for t in trees:
            for leaf in t.leaves.values():
                leaf.append_token(7)
            locs.append(t.alloc().cache_loc)
        host_loc.copy_(locs[0] if len(locs) == 1 else torch.cat(locs))

    def upload(chunks):
        with torch.cuda.stream(s_in):
            for c in chunks:
                if n_dev < n_chunks and c >= n_dev:
                    s_in.wait_event(ev_free[c - n_dev])       # the ring slot's previous tenant has been attended and read back
                dev_qkv[(c % n_dev) * CH: (c % n_dev + 1) * CH].copy_(host_qkv[(c % n_host) * CH: (c % n_host + 1) * CH],
                                                                      non_blocking=True)
                ev_in[c].record(s_in)

    def download(c):
        ev_out[c].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_out[c])
            host_out[(c % n_host) * CH: (c % n_host + 1) * CH].copy_(dev_out[(c % n_dev) * CH: (c % n_dev + 1) * CH],
                                                                     non_blocking=True)
            ev_free[c].record(s_out)

    def run_ring(m):
        """Per-layer eager calls (kv_append + attention) over a ring of device chunk buffers (forests too big to keep 32
        layers of activations on the device)."""
        upload(range(1, min(n_dev, n_chunks)))
        for l in range(LAYERS):
            c = l // CH
            if l % CH == 0:
                main.wait_event(ev_in[c])
            ld = (c % n_dev) * CH + l % CH
            k_new = dev_qkv[ld, :, H * D: (H + HKV) * D].view(nq, HKV, D)
            v_new = dev_qkv[ld, :, (H + HKV) * D:].view(nq, HKV, D)
            deft_b200.kv_append(kv_view.kv_data[l], k_new, v_new, loc_dev)
            K, V = kv_view.get_key_buffer(l), kv_view.get_value_buffer(l)
            qv = dev_qkv[ld, :, : H * D].view(nq, H, D)
            if args.mode == "seq":
                deft_b200.token_attention_fwd(qv, K, V, dev_out[ld], r2t_dev, req_idx, start_loc, seq_lens,
                                              max(seq_len_host), None, sum(seq_len_host))
            elif args.mode == "flatten":
                deft_b200.tree_attention_subtree_fwd(qv, K, V, dev_out[ld], m.block_len, m.block_q, m.block_q_cnts,
                                                     m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
            else:
                deft_b200.tree_attention_fwd(qv, K, V, dev_out[ld], m.node_kv, m.node_kv_offset, m.node_kv_len,
                                             m.node_q, m.node_q_offset, m.node_q_len)
            if l % CH == CH - 1:
                download(c)
                if c + n_dev < n_chunks:
                    upload([c + n_dev])

    def step_e2e():
        """Chunk c+1 of the activations comes up while chunk c attends and chunk c-1's outputs go down."""
        if not args.e2e_static:
            grow()
        s_in.wait_stream(main)                           # the previous step no longer reads dev_qkv
        s_in.wait_stream(s_out)
        upload(range(1))                                 # the first chunk of activations goes up under the table build
        # C++ builder + one H2D copy of tables and plan (into the persistent table buffer of the graphed step)
        m = step.metadata(trees[0] if T == 1 else trees, cache_loc=None if args.e2e_static else host_loc) if graphed else build_meta()
        table_bytes[0] = m.packed.numel()
        loc_dev.copy_(host_loc, non_blocking=True)       # this step's pages (one per leaf)
        if graphed:
            # the other chunks queue BEHIND the tables on the H2D engine, and are enqueued once chunk 0 is launched
            step.run(m, before_chunk=lambda c: main.wait_event(ev_in[c]),
                     after_chunk=lambda c: (upload(range(1, n_chunks)) if c == 0 else None, download(c)))
        else:
            run_ring(m)
        main.wait_stream(s_out)                          # the step ends when the last output is on the host
        main.synchronize()                               # the caller reads the result on the host

    def prepare_next():
        """Host side of the NEXT step (alloc + tables + their upload), enqueued behind the running step."""
        if not args.e2e_static:
            grow()
        if graphed:
            m = pipe.prepare(trees[0] if T == 1 else trees, cache_loc=host_loc, fused=not args.e2e_static)
        else:       # (stream-ordered behind the running step's launches: fresh table tensor, the one cache_loc buffer)
            m = ring_ready[0] = build_meta()
            loc_dev.copy_(host_loc, non_blocking=True)
        table_bytes[0] = m.packed.numel()

    def step_e2e_pipelined():
        """The same step with the host work of step t+1 (a token and a page per leaf, C++ builder, table upload) done
        while the layers of step t run: the tables depend on the tree, not on the tokens step t samples.  The step still
        ends with its outputs on the host, and the next step's activations only go up after that."""
        s_in.wait_stream(main)
        s_in.wait_stream(s_out)
        upload(range(1))
        if graphed:
            pipe.run(before_chunk=lambda c: main.wait_event(ev_in[c]),
                     after_chunk=lambda c: (upload(range(1, n_chunks)) if c == 0 else None, download(c)))
        else:
            run_ring(ring_ready[0])
        prepare_next()                                   # under the GPU time of this step
        main.wait_stream(s_out)
        main.synchronize()                               # the caller reads the result on the host

    ms_serial = None
    wall_serial = None
    ring_ready = [None]
    if args.mode != "seq" and not args.e2e_serial:
        if graphed:                                      # (the ring path of a 512-tree forest: one leg is long enough)
            ms_serial = timed(step_e2e, e2e_steps, 3)
            wall_serial = sorted(step_wall)
        prepare_next()
        if os.environ.get("DEFT_BENCH_STEP_TIMES") and graphed:     # diagnostic: host wall time and captures of every step
            import time as _time
            for i in range(12):
                t0 = _time.perf_counter()
                step_e2e_pipelined()
                print("step %d: %.2f ms, captures %d, table bytes %d" % (i, (_time.perf_counter() - t0) * 1e3, pipe.captures, table_bytes[0]),
                      file=sys.stderr)
        ms_e2e = timed(step_e2e_pipelined, e2e_steps, 3)
        if args.profile_e2e and rank == 0 and graphed and not args.e2e_static:
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(50):
                step_e2e_pipelined()
            pr.disable()
            with open(args.profile_e2e, "w") as f:
                pstats.Stats(pr, stream=f).sort_stats("tottime").print_stats(40)
    else:
        ms_e2e = timed(step_e2e, e2e_steps, 3)
    h2d = LAYERS * nq * (H + 2 * HKV) * D * 2 + table_bytes[0] + host_loc.numel() * 4
    d2h = LAYERS * nq * H * D * 2
    wall = sorted(step_wall)
    res = dict(ms_step=ms_step, ms_s1=ms_s1, ms_s2=ms_s2, ms_e2e=ms_e2e, ms_e2e_serial=ms_serial,
               e2e_wall=dict(median_ms=wall[len(wall) // 2], max_ms=wall[-1], min_ms=wall[0]),
               e2e_wall_serial=None if not wall_serial else dict(median_ms=wall_serial[len(wall_serial) // 2],
                                                                 max_ms=wall_serial[-1], min_ms=wall_serial[0]),
               h2d=h2d, d2h=d2h, nq=nq, clocks=clocks,
               pool_mb=kvp.kv_data[0].numel() * 2 / 1e6, graphed=graphed, pipelined=args.mode != "seq" and not args.e2e_serial,
               captures=(step.captures + pipe.captures) if graphed else None, e2e_chunks=n_chunks, e2e_steps=e2e_steps,
               kv_tokens_end=sum(len(n.kv_indices) for t in trees for n in t.nodes.values()))
    if args.mode == "node_chunk":
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    return res


def roofline(workload: str, T: int, ms_s1: float, mode: str):
    """Stage-1 kernel against the roofline that binds the workload."""
    ws = _scripts()
    alg_b = ws.algorithmic_bytes(workload) * T
    alg_f = ws.algorithmic_flops(workload) * T
    hbm, tflops, src = peaks()
    s = ms_s1 / LAYERS * 1e-3
    tensor_bound = alg_f / alg_b > tflops * 1e12 / (hbm * 1e9)      # arithmetic intensity past the machine balance
    traffic, traffic_src = ncu_traffic(workload, mode, T)
    r = {"bound": "tensor" if tensor_bound else "hbm", "kernel": "stage1_umma_kernel (partial softmax over KV tiles)",
         "algorithmic_bytes_per_launch": alg_b, "algorithmic_flops_per_launch": alg_f, "us_per_launch": s * 1e6,
         "traffic": traffic, "traffic_source": traffic_src, "peak_source": src,
         "hbm_gbs": alg_b / s / 1e9, "hbm_frac": alg_b / s / 1e9 / hbm,
         "tensor_tflops": alg_f / s / 1e12, "tensor_frac": alg_f / s / 1e12 / tflops}
    if tensor_bound:
        r.update(achieved=r["tensor_tflops"], peak=tflops, unit="TFLOP/s", frac=r["tensor_frac"])
    else:
        r.update(achieved=r["hbm_gbs"], peak=hbm, unit="GB/s", frac=r["hbm_frac"])
    return r


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the tree-attention path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_out = sys.stdout
    if world > 1:
        # NCCL writes its banner ("NCCL version ...") to fd 1 when NCCL_DEBUG is set: the JSON line keeps the real
        # stdout, everything else that writes to fd 1 goes to stderr
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    # every rank builds tables and drives copies from Python: keep the ranks off each other's cores
    all_cores = None
    try:
        cores = sorted(os.sched_getaffinity(0))
        all_cores = set(cores)
        if world > 1 and len(cores) >= 2 * world:
            per = len(cores) // world
            os.sched_setaffinity(0, set(cores[local_rank * per: (local_rank + 1) * per]))
        torch.set_num_threads(1 if world > 1 else min(8, len(cores)))
    except (AttributeError, OSError):
        pass

    import deft_b200
    if os.environ.get("DEFT_EXPERIMENT"):            # kernel-variant switch for same-box A/B runs (profiling aid)
        deft_b200._lib.lib.deft_b200_set_experiment(int(os.environ["DEFT_EXPERIMENT"]))
    if os.environ.get("DEFT_PDL"):
        deft_b200._lib.lib.deft_b200_set_pdl(int(os.environ["DEFT_PDL"]))
    if os.environ.get("DEFT_STAGE1_IMPL"):           # 1 = warp-FMA kernel, 2 = tensor-core kernel (A/B runs)
        deft_b200._lib.lib.deft_b200_set_stage1_impl(int(os.environ["DEFT_STAGE1_IMPL"]))
    ws = _scripts()

    warm = max(args.warmup, 3)
    T = max(1, args.trees_per_gpu)
    e2e_steps = max(3, min(3 * args.steps, 60))       # (a decode loop: long enough for the rare table re-layouts to weigh what they weigh)
    r = measure(args, dev, rank, world, args.workload, T, LAYERS, args.steps, warm, e2e_steps, args.e2e_chunk, True)
    torch.cuda.empty_cache()

    cfg5 = None
    if not args.no_cfg5 and args.workload == "cfg2" and T == 1 and args.mode == "flatten" and args.cfg5_trees >= world:
        T5 = args.cfg5_trees // world
        pool_bytes = (ws.unique_kv_tokens("cfg2") + 64 + 64 * 8) * T5 * 2 * HKV * D * 2
        pools5 = int(max(2, min(LAYERS, (56 << 30) // pool_bytes)))
        steps5, e2e5 = max(3, min(args.steps, 5)), 3
        c = measure(args, dev, rank, world, "cfg2", T5, pools5, steps5, 3, e2e5, 2 if T5 > 64 else 4, False)
        torch.cuda.empty_cache()
        rf = roofline("cfg2", T5, c["ms_s1"], "flatten")
        cfg5 = {"workload": f"BASELINE configs[4]: {args.cfg5_trees} independent cfg2 trees sharded over {world} GPU(s)",
                "trees_total": T5 * world, "trees_per_gpu": T5, "layer_pools_per_gpu": pools5,
                "l2": "%d layer pools of %.0f MB cycled by the 32 layer-calls of a step" % (pools5, c["pool_mb"]),
                "steps": steps5, "warmup": 3, "ms_per_step": c["ms_step"],
                "trees_per_s": world * T5 / (c["ms_step"] * 1e-3), "tokens_per_s": world * c["nq"] / (c["ms_step"] * 1e-3),
                "us_per_layer_call": c["ms_step"] / LAYERS * 1e3, "us_stage1": c["ms_s1"] / LAYERS * 1e3,
                "us_stage2": c["ms_s2"] / LAYERS * 1e3,
                "aggregate_hbm_gbs": world * rf["hbm_gbs"], "roofline_frac_stage1": rf["hbm_frac"],
                "roofline_frac_operator": rf["algorithmic_bytes_per_launch"] / (c["ms_step"] / LAYERS * 1e-3) / 1e9 / rf["peak"],
                "e2e": {"trees_per_s": world * T5 / (c["ms_e2e"] * 1e-3), "ms_per_step": c["ms_e2e"], "steps": e2e5,
                        "h2d_bytes_per_step": c["h2d"], "d2h_bytes_per_step": c["d2h"], "graph_captures": c["captures"],
                        "pipelining": "tables of step t+1 (C++ builder over the forest) built while the copies and layers of step t "
                                      "run" if c["pipelined"] else "none"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    nq = r["nq"]
    rf = roofline(args.workload, T, r["ms_s1"], args.mode)
    pools_note = "32 distinct layer KV pools cycled per step (%.0f MB > 126 MB L2)" % (LAYERS * r["pool_mb"])
    line = {
        "metric": METRIC, "value": world * nq / (r["ms_step"] * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {ws.WORKLOADS[args.workload][1]}", "mode": args.mode, "layers": LAYERS,
                   "geometry": "H=32 HKV=8 D=128 fp16", "trees_per_gpu": T, "queries_per_tree": nq // T,
                   "l2": pools_note,
                   "timing": "CUDA graph of one step (64 launches), CUDA events, max over ranks"},
        "trees_per_s": world * T / (r["ms_step"] * 1e-3),
        "us_per_layer_call": r["ms_step"] / LAYERS * 1e3,
        "us_stage1": r["ms_s1"] / LAYERS * 1e3, "us_stage2": r["ms_s2"] / LAYERS * 1e3,
        "clocks": r["clocks"],
        "e2e": {"value": world * nq / (r["ms_e2e"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms_e2e"],
                "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"], "steps": r["e2e_steps"],
                "tree": "fixed shape (--e2e-static)" if args.e2e_static else
                        "GROWS: every step appends one token + one page per leaf (TreeCache.alloc) before the tables are rebuilt; "
                        "%d KV tokens per tree at the end of the run" % (r["kv_tokens_end"] // T),
                "graph_captures": r["captures"],
                "ms_per_step_serial": r["ms_e2e_serial"],
                "host_wall_per_step": r["e2e_wall"],
                "host_wall_per_step_serial": r["e2e_wall_serial"],
                "pipelining": ("DecodeStepPipeline: alloc + table build + table upload of step t+1 run on the host while the layers of "
                               "step t run (two table buffers; the tables depend on the tree, not on the tokens step t samples); every "
                               "step still ends with its outputs on the host before the next step's activations go up; "
                               "ms_per_step_serial = the same loop with the build in line") if r["pipelined"] else
                              "none (table build in line)",
                "path": ("TreeCache.alloc + DecodeStepGraph.metadata (TreeMetadata.from_tree_cache: C++ builder, capacity-padded tables, 1 "
                         "upload into the persistent table buffer; the first chunk of activations goes up under the build, the others "
                         "queue behind the tables) + pinned H2D of the fused qkv in %d-layer chunks on a copy stream + 32 x tree attention "
                         "with the KV append fused in (this step's K/V read from the activations, written to the pool by stage 2) "
                         "replayed as %d CUDA graphs + D2H of the outputs per chunk on a second copy stream; timed until the last "
                         "output is on the host" % (args.e2e_chunk, r["e2e_chunks"])) if r["graphed"] else
                        "per-layer eager calls (kv_append + attention) between chunked pinned H2D / D2H copies"},
        "gpu_launches": args.steps * LAYERS * 2,
        "roofline": rf,
    }
    if cfg5 is not None:
        line["cfg5"] = cfg5
    if not args.no_cpu_baseline and world == 1:      # (the CPU baseline is a rank-0, N = 1 number: it wants every host core)
        line["cpu_baseline"] = cpu_baseline(args.workload)
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
