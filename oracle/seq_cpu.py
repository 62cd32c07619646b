"""CPU timing leg of the sequence-based baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference has no CPU implementation of this path, and its unpaged Flash-Decoding attention is a
stub that returns ``q`` (``/root/reference/DeFT/deft/layers/attention/deft_attention.py:229-266``).
What is timed on the host cores is therefore a port of the *semantics* of its sequence-based modes
(``TreeCache.get_kv_seq`` gather, ``deft/tree_decoding/tree_cache.py:417-439``; paged equivalent
``token_attention_fwd``, ``deft/layers/attention/token_attention.py:297-335``): every leaf attends
independently over its own root->leaf KV, so shared-prefix pages are re-read once per leaf.
Same arithmetic as ``oracle.deft_oracle.seq_attention`` (checked in tests/test_oracle_golden.py),
written with multi-threaded fp32 torch ops so that it can use every host core.
"""
from __future__ import annotations

import math
import os
import time
from typing import List, Sequence

import numpy as np
import torch


def seq_attention_torch(q: torch.Tensor, kv: torch.Tensor, paths: Sequence[np.ndarray]) -> torch.Tensor:
    """q [nq,H,D] fp16 (CPU), kv [pool,2,HKV,D] fp16 (CPU); returns [nq,H,D] fp16."""
    nq, H, D = q.shape
    HKV = kv.shape[2]
    G = H // HKV
    out = torch.empty(nq, H, D, dtype=torch.float16)
    scale = 1.0 / math.sqrt(D)
    for i, pages in enumerate(paths):
        rows = kv[torch.as_tensor(pages, dtype=torch.long)].float()          # gather: [n, 2, HKV, D]
        k = rows[:, 0].permute(1, 2, 0)                                     # [HKV, D, n]
        v = rows[:, 1].permute(1, 0, 2)                                     # [HKV, n, D]
        qq = q[i].float().view(HKV, G, D)
        p = torch.softmax(torch.bmm(qq, k) * scale, dim=-1)                 # [HKV, G, n]
        out[i] = torch.bmm(p, v).view(H, D).half()
    return out


def time_layer_calls(q: torch.Tensor, kv_layers: List[torch.Tensor], paths, reps: int, warmup: int = 1):
    """Seconds per layer-call (median over ``reps``), cycling the given layer pools."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    times = []
    for r in range(warmup + reps):
        t0 = time.perf_counter()
        seq_attention_torch(q, kv_layers[r % len(kv_layers)], paths)
        dt = time.perf_counter() - t0
        if r >= warmup:
            times.append(dt)
    return float(np.median(times)), threads
