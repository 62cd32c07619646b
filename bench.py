#!/usr/bin/env python
"""Decode-step tree-attention benchmark (BASELINE.json metric) for deft_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A *step* is one decode step of the tree-attention path on a Llama-3-8B-geometry synthetic tree:
32 layer-calls of DeFT-Flatten attention (stage 1 + stage 2), each layer over its OWN KV pool, so
that the 32 pools (cfg2: 0.8 GB) cycle far beyond the 126 MB L2 between consecutive calls.

* ``value``      tokens/s with everything resident in HBM; the step is one CUDA graph of the 64
                 launches, timed with CUDA events on the launching stream, max over ranks.
* ``e2e``        the same step through the public Python API with HOST buffers: C++ table builder +
                 one-copy upload, host->device copy of the step's fused qkv activations (pinned),
                 KV append, 32 x tree_attention_subtree_fwd, device->host read of the outputs.
* ``roofline``   stage-1 kernel alone (graph of 32 stage-1 launches): algorithmic bytes / duration
                 against MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the sequence-based (per-leaf, no prefix reuse) semantics
                 of the reference on the host cores (oracle/seq_cpu.py); a bounded sample.

With N > 1 (torchrun) every rank runs its own tree (trees shard with no collective on the data
path); NCCL carries the barrier and the max-over-ranks reduction only.  scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

LAYERS = 32
METRIC = "decode_attention_tokens_per_s"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="flatten", choices=["flatten", "node", "node_chunk", "seq"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--e2e-chunk", type=int, default=8, choices=[1, 2, 4, 8, 16, 32],
                    help="layers per H2D / graph / D2H chunk of the end-to-end leg")
    ap.add_argument("--trees-per-gpu", type=int, default=1,
                    help="independent trees of the workload batched into ONE launch per layer (BASELINE cfg 5)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def _path(leaf):
    nodes = []
    while leaf is not None:
        nodes.append(leaf)
        leaf = leaf.parent
    return nodes[::-1]


def cpu_paths_and_inputs(workload: str, layers: int):
    """Host copies of the synthetic inputs for the CPU baseline (same shapes, seeded)."""
    from oracle import deft_oracle as orc
    from deft_b200.workloads import build_tree
    tree = build_tree(workload, layers=0, device="cpu")
    paths = orc.leaf_paths(tree)
    g = torch.Generator().manual_seed(0)
    size = len(tree.token_to_kv_pool.mem_state)
    kv_layers = [torch.randn(size, 2, 8, 128, generator=g, dtype=torch.float32).half() for _ in range(layers)]
    q = torch.randn(len(paths), 32, 128, generator=g, dtype=torch.float32).half()
    return q, kv_layers, paths


def cpu_baseline(workload: str, reps: int):
    from oracle.seq_cpu import time_layer_calls
    q, kv_layers, paths = cpu_paths_and_inputs(workload, layers=2)
    sec, threads = time_layer_calls(q, kv_layers, paths, reps=reps, warmup=1)
    nq = len(paths)
    return {"value": nq / (sec * LAYERS), "unit": UNIT, "cores": threads, "kind": "port",
            "ms_per_layer_call": sec * 1e3,
            "sample": f"{reps} layer-calls of {workload} (of the 32 a step has), sequence-based per-leaf attention "
                      f"(oracle/seq_cpu.py, fp32 torch bmm, {threads} threads); step time = 32 x median layer-call"}


def run_reference(args, rank: int):
    """--impl reference: the reference's sequence-based path on the host cores (bounded sample per step)."""
    if rank != 0:
        return
    from oracle.seq_cpu import time_layer_calls
    from deft_b200.workloads import WORKLOADS
    q, kv_layers, paths = cpu_paths_and_inputs(args.workload, layers=2)
    steps = min(args.steps, 5)
    warm = min(args.warmup, 1)
    sec, threads = time_layer_calls(q, kv_layers, paths, reps=steps, warmup=warm)
    nq = len(paths)
    value = nq / (sec * LAYERS)
    sample = (f"each step = 1 layer-call of {args.workload} sampled from the 32 (x32 extrapolated); "
              f"{steps} timed / {warm} warm-up (capped from --steps {args.steps} --warmup {args.warmup})")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sec * LAYERS * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][2]}", "layers": LAYERS,
                       "path": "sequence-based per-leaf attention on host cores (reference has no CPU kernel; port)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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

    import deft_b200
    if os.environ.get("DEFT_EXPERIMENT"):            # kernel-variant switch for same-box A/B runs (profiling aid)
        deft_b200._lib.lib.deft_b200_set_experiment(int(os.environ["DEFT_EXPERIMENT"]))
    if os.environ.get("DEFT_PDL"):
        deft_b200._lib.lib.deft_b200_set_pdl(int(os.environ["DEFT_PDL"]))
    from deft_b200 import BLOCK_CONFIG, TreeMetadata, _lib
    from deft_b200.sharding import max_over_ranks
    from deft_b200.workloads import WORKLOADS, algorithmic_bytes, build_forest

    # ---- synthetic inputs: T trees per rank over one page pool, 32 layer pools, random-normal fp16 -----
    torch.manual_seed(1234 + rank)
    T = max(1, args.trees_per_gpu)
    trees = build_forest(args.workload, T, layers=LAYERS, device=dev)
    kvp = trees[0].token_to_kv_pool
    for l in range(LAYERS):
        kvp.kv_data[l].normal_()
    nq = sum(len(t.leaves) for t in trees)

    def build_meta():
        return TreeMetadata.from_tree_cache(trees[0]) if T == 1 else TreeMetadata.from_forest(trees)

    H, HKV, D = 32, 8, 128
    qkv = torch.randn(LAYERS, nq, (H + 2 * HKV) * D, dtype=torch.float16, device=dev)   # fused qkv, row stride 6144
    out = torch.empty(LAYERS, nq, H, D, dtype=torch.float16, device=dev)
    if args.mode == "node_chunk":
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    meta = build_meta()

    def q_of(buf, l):
        return buf[l, :, : H * D].view(nq, H, D)

    if args.mode == "seq":     # Radix / sequence-based baseline ON OUR KERNELS: every leaf re-reads its whole path
        r2t = trees[0].req_to_token_pool
        sleaves = [leaf for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)]
        req_idx = torch.tensor([t.leaf_to_req[leaf.id] for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)],
                               dtype=torch.int32, device=dev)
        seq_len_host = [sum(len(n.kv_indices) for n in _path(leaf)) for leaf in sleaves]
        seq_lens = torch.tensor(seq_len_host, dtype=torch.int32, device=dev)
        start_loc = torch.zeros_like(seq_lens)
        r2t_dev = r2t.device_table() if hasattr(r2t, "device_table") else r2t.req_to_token

    def attention(l, qbuf, m):
        K, V = kvp.get_key_buffer(l), kvp.get_value_buffer(l)
        if args.mode == "seq":
            deft_b200.token_attention_fwd(q_of(qbuf, l), K, V, out[l], r2t_dev, req_idx, start_loc, seq_lens,
                                          max(seq_len_host), None, sum(seq_len_host))
        elif args.mode == "flatten":
            deft_b200.tree_attention_subtree_fwd(q_of(qbuf, l), K, V, out[l], m.block_len, m.block_q, m.block_q_cnts,
                                                 m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
        else:
            deft_b200.tree_attention_fwd(q_of(qbuf, l), K, V, out[l], m.node_kv, m.node_kv_offset, m.node_kv_len,
                                         m.node_q, m.node_q_offset, m.node_q_len)

    def step_resident():
        for l in range(LAYERS):
            attention(l, qkv, meta)

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

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev) / steps

    warm = max(args.warmup, 3)
    g_full, g_s1, g_s2 = capture(7), capture(2), capture(4)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(g_full.replay, args.steps, warm)
    clocks = sampler.stop() if sampler else None
    ms_s1 = timed(g_s1.replay, args.steps, warm)
    ms_s2 = timed(g_s2.replay, args.steps, warm)

    # ---- end to end through the public API with host buffers -------------------------------------
    host_qkv = torch.randn(LAYERS, nq, (H + 2 * HKV) * D, dtype=torch.float16).pin_memory()
    host_out = torch.empty(LAYERS, nq, H, D, dtype=torch.float16).pin_memory()
    dev_qkv = torch.empty_like(qkv)
    leaves = [leaf for t in trees for leaf in sorted(t.leaves.values(), key=lambda x: x.id)]
    host_loc = torch.tensor([leaf.kv_indices[-1] for leaf in leaves], dtype=torch.int32).pin_memory()
    table_bytes = [0]
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()      # H2D and D2H ride their own copy engines
    CH = args.e2e_chunk                                          # layers per copy chunk (and per CUDA graph)
    ev_in = [torch.cuda.Event() for _ in range(LAYERS // CH)]
    ev_out = [torch.cuda.Event() for _ in range(LAYERS // CH)]

    loc_dev = torch.zeros(nq, dtype=torch.int32, device=dev)
    graphed = args.mode != "seq"
    if graphed:      # the step's launches (per layer: kv_append + attention) as CUDA graphs, one per chunk of layers
        step = deft_b200.DecodeStepGraph(kvp, dev_qkv, out, loc_dev, H, HKV, D, mode=args.mode, chunk=CH)

    def after_chunk(c):
        ev_out[c].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_out[c])
            host_out[c * CH: (c + 1) * CH].copy_(out[c * CH: (c + 1) * CH], non_blocking=True)

    def step_e2e():
        """Chunk c+1 of the activations comes up while chunk c attends and chunk c-1's outputs go down."""
        s_in.wait_stream(main)                           # the previous step no longer reads dev_qkv

        def upload(chunks):
            with torch.cuda.stream(s_in):
                for c in chunks:
                    dev_qkv[c * CH: (c + 1) * CH].copy_(host_qkv[c * CH: (c + 1) * CH], non_blocking=True)
                    ev_in[c].record(s_in)

        upload(range(1))                                 # the first chunk of activations goes up under the table build
        # C++ builder + one H2D copy of tables and plan (into the persistent table buffer of the graphed step)
        m = step.metadata(trees[0] if T == 1 else trees) if graphed else build_meta()
        table_bytes[0] = m.packed.numel()
        loc_dev.copy_(host_loc, non_blocking=True)       # this step's pages (one per leaf)
        if graphed:
            # the other chunks queue BEHIND the tables on the H2D engine, and are enqueued once chunk 0 is launched
            step.run(m, before_chunk=lambda c: main.wait_event(ev_in[c]),
                     after_chunk=lambda c: (upload(range(1, LAYERS // CH)) if c == 0 else None, after_chunk(c)))
        else:
            upload(range(1, LAYERS // CH))
            for l in range(LAYERS):
                if l % CH == 0:
                    main.wait_event(ev_in[l // CH])
                k_new = dev_qkv[l, :, H * D: (H + HKV) * D].view(nq, HKV, D)
                v_new = dev_qkv[l, :, (H + HKV) * D:].view(nq, HKV, D)
                deft_b200.kv_append(kvp.kv_data[l], k_new, v_new, loc_dev)
                attention(l, dev_qkv, m)
                if l % CH == CH - 1:
                    after_chunk(l // CH)
        main.wait_stream(s_out)                          # the step ends when the last output is on the host
        main.synchronize()                               # the caller reads the result on the host

    e2e_steps = max(3, min(args.steps, 20))
    ms_e2e = timed(step_e2e, e2e_steps, 3)
    h2d = host_qkv.numel() * 2 + table_bytes[0] + host_loc.numel() * 4
    d2h = host_out.numel() * 2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg = algorithmic_bytes(args.workload) * T
    traffic, traffic_src = ncu_traffic(args.workload, args.mode, T)
    peak, peak_src = peaks()
    s1_s = ms_s1 / LAYERS * 1e-3
    achieved = alg / s1_s / 1e9
    line = {
        "metric": METRIC, "value": world * nq / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][2]}", "mode": args.mode, "layers": LAYERS,
                   "geometry": "H=32 HKV=8 D=128 fp16", "trees_per_gpu": T, "queries_per_tree": nq // T,
                   "l2": "32 distinct layer KV pools cycled per step (%.0f MB > 126 MB L2)" % (LAYERS * kvp.kv_data[0].numel() * 2 / 1e6),
                   "timing": "CUDA graph of one step (64 launches), CUDA events, max over ranks"},
        "trees_per_s": world * T / (ms_step * 1e-3),
        "us_per_layer_call": ms_step / LAYERS * 1e3,
        "us_stage1": ms_s1 / LAYERS * 1e3, "us_stage2": ms_s2 / LAYERS * 1e3,
        "clocks": clocks,
        "e2e": {"value": world * nq / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "path": ("DecodeStepGraph.metadata (TreeMetadata.from_tree_cache: C++ builder, 1 upload into the persistent table "
                         "buffer; the first chunk of activations goes up under the build, the others queue behind the tables) + pinned H2D "
                         "of the fused qkv in %d-layer chunks on a copy stream + 32 x (kv_append + "
                         "tree attention) replayed as %d CUDA graphs + D2H of the outputs per chunk on a second copy stream; "
                         "timed until the last output is on the host" % (CH, LAYERS // CH)) if graphed else
                        "per-layer eager calls (kv_append + token_attention_fwd) between chunked pinned H2D / D2H copies"},
        "gpu_launches": args.steps * LAYERS * 2,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "stage1 (partial softmax over KV items)",
                     "algorithmic_bytes_per_launch": alg, "us_per_launch": s1_s * 1e6, "peak_source": peak_src},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_reps)
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
