#!/usr/bin/env python
"""Host-side cost of TreeMetadata.from_tree_cache on a CUDA pool (profiling aid): python tools/prof_meta.py [cfg2]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from deft_b200 import TreeMetadata
from deft_b200.tree_cache import build_tables_host, flatten_tree
from deft_b200.workloads import build_tree

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
tree = build_tree(wl, layers=1, device="cuda:0")
buf = torch.empty(8 << 20, dtype=torch.uint8, device="cuda:0")


def t(fn, n=100):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return dt


f = flatten_tree(tree)
print("flatten_tree %.3f ms" % t(lambda: flatten_tree(tree)))
print("build_tables_host %.3f ms" % t(lambda: build_tables_host(f, hkv=8, n_ctas=148)))
print("from_tree_cache %.3f ms" % t(lambda: TreeMetadata.from_tree_cache(tree)))
print("from_tree_cache(device_buffer) %.3f ms" % t(lambda: TreeMetadata.from_tree_cache(tree, device_buffer=buf)))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    TreeMetadata.from_tree_cache(tree, device_buffer=buf)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
