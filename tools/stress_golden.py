#!/usr/bin/env python
"""Repeats the plan-less (device-derived plan) Flatten call on every golden scenario and counts results outside
the reference tolerance: python tools/stress_golden.py [reps]   """
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import deft_b200
from deft_b200 import _lib
from oracle.scenarios import SCENARIOS

KEYS = ["block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens"]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    _lib.lib.deft_b200_set_stage1_impl(_lib.STAGE1_UMMA)
    dev = torch.device("cuda:0")
    for name in SCENARIOS:
        z = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        q = torch.from_numpy(z["q"]).to(dev)
        pool = torch.from_numpy(z["kv_pool"]).to(dev)
        K, V = pool[:, 0], pool[:, 1]
        t = {k: torch.from_numpy(z["t_" + k]).to(dev) for k in KEYS}
        want = z["o_flatten"].astype(np.float32)
        bad = 0
        for rep in range(reps):
            o = torch.full_like(q, float("nan"))
            deft_b200.tree_attention_subtree_fwd(q, K, V, o, 128, t["block_q"], t["block_q_cnts"], t["block_q_offset"],
                                                 t["block_bitmasks"], t["block_kv"], t["block_lens"])
            got = o.float().cpu().numpy()
            if not np.allclose(got, want, atol=1e-3, rtol=1e-2):
                bad += 1
                w = np.argwhere(~np.isclose(got, want, atol=1e-3, rtol=1e-2))
                print(f"  {name} rep {rep}: {len(w)} bad elements, first {w[:3].tolist()}, queries {sorted(set(w[:, 0].tolist()))[:8]} "
                      f"heads {sorted(set(w[:, 1].tolist()))}")
        print(f"{name}: {bad}/{reps} bad")


if __name__ == "__main__":
    main()
