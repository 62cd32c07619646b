"""Tree sharding across ranks and the cross-rank reductions of bench.py, on CPU with gloo (world_size 2)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deft_b200.sharding import gather_ranges, max_over_ranks, shard_range, sum_over_ranks


def test_shard_ranges_partition_the_trees():
    for n in (0, 1, 7, 512, 513):
        for world in (1, 2, 4, 8):
            rs = gather_ranges(n, world)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard_range(512, 3, 8) == (192, 256)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    begin, end = shard_range(513, rank, world)
    # each rank "processes" its trees: the job total is the sum, the step time the max over ranks
    ms = max_over_ranks(10.0 + rank, dev)
    trees = sum_over_ranks(end - begin, dev)
    checksum = sum_over_ranks(float(sum(range(begin, end))), dev)
    dist.barrier()
    out.put((rank, ms, trees, checksum))
    dist.destroy_process_group()


def test_reductions_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, trees, checksum in got:
        assert ms == 11.0 and trees == 513 and checksum == float(sum(range(513)))


def test_single_process_is_identity():
    assert max_over_ranks(3.5, torch.device("cpu")) == 3.5 and sum_over_ranks(2.0, torch.device("cpu")) == 2.0


def test_reference_arm_prints_one_contract_line_and_other_ranks_stay_silent():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): one JSON line with the contract's keys
    on rank 0, nothing and exit 0 on any other rank.  Runs the small cfg1 workload on the host cores."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
           "--warmup", "0", "--gpus", "2"]
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 2
    for k in ("metric", "unit", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env["RANK"] = env["LOCAL_RANK"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout, r.stderr[-500:])
