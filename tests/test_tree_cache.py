"""deft_b200.TreeCache replays the scripted scenarios to the SAME pages / refs / tables as the reference.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle.plain_tree import freeze
from oracle.scenarios import SCENARIOS, TABLE_SCENARIOS, replay

from deft_b200 import _lib
from deft_b200.memory_pool import ReqToTokenPool, TokenToKVPool, TreeIndexPool
from deft_b200.tree_cache import BLOCK_CONFIG, TreeCache, TreeMetadata

TABLE_KEYS = _lib.T_NAMES[:12]


def build(cfg):
    H, HKV, D = cfg.get("H", 8), cfg.get("HKV", 1), cfg.get("D", 16)
    r2t = ReqToTokenPool(size=128, max_context_len=cfg["max_ctx"], device="cpu")
    kvp = TokenToKVPool(size=cfg["pool"], dtype=torch.float16, head_num=HKV, head_dim=D, layer_num=1, device="cpu")
    tix = TreeIndexPool(size=64, max_context_len=cfg["max_ctx"], device="cpu") if cfg.get("tree_index") else None
    tree = TreeCache(torch.float16, HKV, D, 1, r2t, kvp, tix, True, tix is not None)
    replay(tree, cfg["script"], lambda n: torch.arange(1, n + 1, dtype=torch.int32))
    return tree, r2t, kvp, tix


@pytest.mark.parametrize("name", list(SCENARIOS) + list(TABLE_SCENARIOS))
def test_replay_is_bit_identical_to_reference(golden_dir, name):
    cfg = {**SCENARIOS, **TABLE_SCENARIOS}[name]
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    tree, r2t, kvp, tix = build(cfg)
    mine = freeze(tree)
    for k, v in mine.items():
        assert np.array_equal(v, z["tree_" + k]), (name, k)
    if "mem_state" in z.files:
        assert np.array_equal(kvp.mem_state.astype(np.int64), z["mem_state"])
    if "req_to_token" in z.files and name != "spec_merge":
        # compare the rows of live leaves (the reference's table is torch.empty elsewhere)
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        assert [tree.leaf_to_req[l.id] for l in leaves] == z["req_idx"].tolist()
        for l, n in zip(leaves, z["seq_lens"].tolist()):
            row = tree.leaf_to_req[l.id]
            assert np.array_equal(r2t.req_to_token[row, :n].numpy(), z["req_to_token"][row, :n])
    if tix is not None:
        for n in tree.nodes.values():
            ln = len(n.kv_indices)
            assert np.array_equal(tix.node_to_kv[n.node_indices_id, :ln].numpy(), z["node_to_kv"][n.node_indices_id, :ln])


@pytest.mark.parametrize("name", list(SCENARIOS) + list(TABLE_SCENARIOS))
def test_metadata_from_tree_cache(golden_dir, name):
    cfg = {**SCENARIOS, **TABLE_SCENARIOS}[name]
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    tree, *_ = build(cfg)
    for prefix, mbl in (("t_", -1), ("tc_", 128)):
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = mbl          # how the reference CLI selects node_chunk
        try:
            m = TreeMetadata.from_tree_cache(tree)
        finally:
            BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
        for k in TABLE_KEYS:
            got = getattr(m, k)
            assert got.dtype == torch.int64
            assert np.array_equal(got.numpy(), z[prefix + k]), (name, prefix, k)
        q_num, node_num, total, blen = z[prefix + "scalars"].tolist()
        assert (m.query_num, m.node_num, m.total_kv_len, m.block_len) == (q_num, node_num, total, blen)
        assert sorted(m.leaf_to_q.items()) == [tuple(r) for r in z[prefix + "leaf_to_q"].tolist()]
        assert m.flat_plan.n_part_rows == len(z[prefix + "block_q"]) and m.node_plan.n_items >= node_num


def test_metadata_tree_index_mode(golden_dir):
    z = np.load(os.path.join(golden_dir, "tree_index.npz"))
    tree, *_ = build(SCENARIOS["tree_index"])
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        m = TreeMetadata.from_tree_cache_node(tree)
    finally:
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
    for k in ["node_q", "node_q_len", "node_q_offset", "node_kv_offset", "node_kv_len"]:
        assert np.array_equal(getattr(m, k).numpy(), z["ti_" + k]), k
    assert m.node_kv.dtype == torch.int32 and m.block_q.numel() == 0 and m.flat_plan is None


def test_pool_allocator_contract():
    """First-free ascending allocation, refcounts, exhaustion -> None (memory_pool.py:74-108)."""
    p = TokenToKVPool(size=8, dtype=torch.float16, head_num=1, head_dim=16, layer_num=2, device="cpu")
    assert p.kv_data[0].shape == (8, 2, 1, 16) and p.get_key_buffer(1).stride() == (32, 16, 1)
    a = p.alloc(3)
    assert a.dtype == torch.int32 and a.tolist() == [0, 1, 2]
    p.add_refs([1])
    assert p.free(torch.tensor([0, 1])) == 1            # page 1 still referenced
    assert p.alloc(2).tolist() == [0, 3] and p.used_size() == 4 and p.available_size() == 4
    assert p.alloc(5) is None and p.alloc(4).tolist() == [4, 5, 6, 7]
    p.clear()
    assert p.available_size() == 8 and p.alloc_ct == 0
    r = ReqToTokenPool(size=3, max_context_len=4, device="cpu")
    assert r.alloc(2).tolist() == [0, 1] and r.alloc(2) is None
    r.free(0)
    assert r.alloc(1).tolist() == [0]


def test_unpaged_mode_is_refused():
    with pytest.raises(NotImplementedError):
        TreeCache(torch.float16, 1, 16, 1, None, None, None, use_paged_memory=False)


def _fresh_flat(tree):
    """flatten_tree with every cache dropped (the walk and the per-node page arrays)."""
    from deft_b200.tree_cache import flatten_tree
    if hasattr(tree, "_flat_topo"):
        del tree._flat_topo
    for n in tree.nodes.values():
        if hasattr(n, "_kv_np"):
            del n._kv_np
    return flatten_tree(tree)


def test_flatten_tree_caches_follow_every_mutation():
    """The kept walk / page arrays of flatten_tree give what a fresh walk gives after alloc, branch, cut and merge."""
    import random
    from deft_b200.tree_cache import flatten_tree
    rng = random.Random(7)
    r2t = ReqToTokenPool(size=256, max_context_len=512, device="cpu")
    kvp = TokenToKVPool(size=8192, dtype=torch.float16, head_num=1, head_dim=16, layer_num=1, device="cpu")
    tree = TreeCache(torch.float16, 1, 16, 1, r2t, kvp, None, True, False)
    tree.init_prompt(torch.arange(1, 201, dtype=torch.int32))

    def check():
        got = flatten_tree(tree)              # cached path
        again = flatten_tree(tree)            # nothing changed in between
        want = _fresh_flat(tree)
        for k in ("parent", "kv_off", "kv", "q_off", "qs", "tix"):
            assert np.array_equal(got[k], want[k]), k
            assert np.array_equal(again[k], want[k]), k
        assert got["leaf_to_q"] == want["leaf_to_q"]

    check()
    for step in range(60):
        op = rng.random()
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        if op < 0.25 and len(leaves) < 24:
            leaf = rng.choice(leaves)
            if leaf.get_len() > 0:
                tree.branch(leaf, rng.choice((2, 3)))
        elif op < 0.35 and len(leaves) > 2:
            tree.cut(rng.choice(leaves))
        elif op < 0.40:
            leaf = rng.choice(leaves)
            leaf.paused = False                # the setter moves the epoch: the walk is redone, same result
        for leaf in tree.leaves.values():      # a decode step: one token and one page per leaf
            leaf.append_token(step)
        tree.alloc()
        check()


def test_plan_search_cache_gives_the_cold_answer(monkeypatch):
    """Two builds in a row (the second takes the remembered piece length) equal a build with the memory off."""
    from deft_b200.tree_cache import build_tables_host, flatten_tree
    from deft_b200.workloads import build_tree
    for name in ("cfg3", "cfg2"):
        tree = build_tree(name, layers=1, device=torch.device("cpu"), H=4, HKV=2, D=16)
        flat = flatten_tree(tree)
        a, da, sa = build_tables_host(flat, hkv=2, n_ctas=148)
        b, db, sb = build_tables_host(flat, hkv=2, n_ctas=148)
        monkeypatch.setenv("DEFT_PLAN_CACHE", "0")
        c, dc, sc = build_tables_host(flat, hkv=2, n_ctas=148)
        monkeypatch.delenv("DEFT_PLAN_CACHE")
        assert np.array_equal(a, b) and np.array_equal(a, c)
        assert np.array_equal(da, db) and np.array_equal(da, dc) and np.array_equal(sa, sc)


def test_flatten_tree_on_a_foreign_tree_object():
    """A TreeCache-shaped object that does not count its structural changes (the reference's own class): the walk is
    redone at every call, the per-node page arrays follow appends and list replacements."""
    from types import SimpleNamespace as NS
    from deft_b200.tree_cache import flatten_tree

    class Node:                                    # hashable by identity, like the reference's TreeNode
        def __init__(self, i, pages):
            self.id, self.children, self.kv_indices, self.parent = i, {}, list(pages), None
            self.refs, self.paused, self.node_indices_id = set(), False, None

    node = Node

    root, a, b = node(0, range(10, 30)), node(1, [40, 42]), node(2, [41, 43])
    for ch in (a, b):
        ch.parent = root
        root.children[ch.id] = ch
        ch.refs.add(ch)
        root.refs.add(ch)
    tree = NS(root=root, leaves={1: a, 2: b}, nodes={0: root, 1: a, 2: b})

    def expect():
        order = [root] + list(root.children.values())
        return np.concatenate([np.asarray(n.kv_indices, dtype=np.int64) for n in order])

    assert np.array_equal(flatten_tree(tree)["kv"], expect())
    a.kv_indices.append(44)                       # append_index
    b.kv_indices.append(45)
    assert np.array_equal(flatten_tree(tree)["kv"], expect())
    a.kv_indices = [7, 8, 9]                      # replaced by a new list (reset_node_KV + re-alloc)
    assert np.array_equal(flatten_tree(tree)["kv"], expect())
    a.kv_indices = a.kv_indices[:2]               # a shorter NEW list
    assert np.array_equal(flatten_tree(tree)["kv"], expect())
    c = node(3, [50])                             # structural change without any counter: seen at the next call
    c.parent = a
    a.children[3] = c
    tree.leaves = {2: b, 3: c}
    tree.nodes[3] = c
    a.refs.discard(a); root.refs.discard(a)
    for n in (c, a, root):
        n.refs.add(c)
    f = flatten_tree(tree)
    assert f["parent"].tolist() == [-1, 0, 1, 0] and f["leaf_to_q"] == {2: 0, 3: 1}
    assert np.array_equal(f["kv"], np.asarray(list(range(10, 30)) + [7, 8, 50, 41, 43, 45], dtype=np.int64))


def test_add_refs_counts_a_page_named_twice_once_like_the_reference():
    """``mem_state[idx] += 1`` (reference memory_pool.py:92-98) is an indexed read-modify-write: duplicates in the
    index list move the refcount ONCE.  ``alloc_ct`` counts list entries."""
    kvp = TokenToKVPool(size=16, dtype=torch.float16, head_num=1, head_dim=16, layer_num=0, device="cpu")
    want = torch.zeros(16, dtype=torch.int16)
    idx = torch.tensor([3, 5, 5, 7, 3, 3])
    kvp.add_refs(idx)
    want[idx] += 1
    assert np.array_equal(kvp.mem_state, want.numpy()) and kvp.alloc_ct == 6
    kvp.add_refs([5, 9])
    want[torch.tensor([5, 9])] += 1
    freed = kvp.decrease_refs(torch.tensor([5, 5, 9]))
    want[torch.tensor([5, 5, 9])] -= 1
    assert np.array_equal(kvp.mem_state, want.numpy())
    assert freed == int((want[torch.tensor([5, 5, 9])] == 0).sum()) and kvp.alloc_ct == 5


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4"])
def test_workload_scripts_and_their_pure_python_replay(name):
    """The scripts of deft_b200/workload_scripts.py: the closed-form counts, the TreeCache replay and the pure-Python
    replay bench.py's reference arm uses (oracle/sim_tree.py) agree on every page list."""
    from oracle import deft_oracle as orc
    from oracle.sim_tree import SimTree
    from deft_b200.workloads import (WORKLOADS, algorithmic_flops, build_tree, max_path_len, n_leaves, n_nodes,
                                     unique_kv_tokens)
    tree = build_tree(name, layers=0, device=torch.device("cpu"), H=4, HKV=2, D=16)
    sim = SimTree().replay(WORKLOADS[name][0])
    assert sorted(tree.nodes) == sorted(sim.nodes) and sorted(tree.leaves) == sorted(sim.leaves)
    for i, n in tree.nodes.items():
        assert n.kv_indices == sim.nodes[i].kv_indices, (name, i)
        assert (n.parent.id if n.parent else -1) == (sim.nodes[i].parent.id if sim.nodes[i].parent else -1)
    paths = orc.leaf_paths(tree)
    assert all(np.array_equal(a, b) for a, b in zip(paths, orc.leaf_paths(sim)))
    assert unique_kv_tokens(name) == sum(len(n.kv_indices) for n in tree.nodes.values()) == sim.next_page
    assert n_leaves(name) == len(tree.leaves) and n_nodes(name) == len(tree.nodes)
    assert max_path_len(name) == max(len(p) for p in paths)
    assert algorithmic_flops(name, H=32, D=128) == sum(len(p) for p in paths) * 32 * 4 * 128


def test_cfg3b_is_a_medusa_style_sparse_tree():
    """BASELINE configs[2] / SURVEY.md 8d cfg 3b: width 6, depth 5, 63 one-token nodes chosen best-first by path score."""
    from deft_b200.workloads import build_tree, medusa_tree
    paths = medusa_tree()
    assert len(paths) == 63 and len(set(paths)) == 63
    assert max(len(p) for p in paths) == 5 and max(max(p) for p in paths) <= 5
    chosen = set(paths)
    for p in paths:                                   # a node's parent and its better-ranked siblings are in the tree
        assert len(p) == 1 or p[:-1] in chosen
        assert p[-1] == 0 or p[:-1] + (p[-1] - 1,) in chosen
    p_r = [0.6 * 0.4 ** r for r in range(6)]
    score = lambda path: float(np.prod([p_r[r] for r in path]))
    worst = min(score(p) for p in paths)
    for p in paths:                                   # nothing left out scores better than the worst node taken
        if len(p) < 5:
            for r in range(6):
                if p + (r,) not in chosen:
                    assert score(p + (r,)) <= worst + 1e-15
    tree = build_tree("cfg3b", layers=0, device=torch.device("cpu"), H=4, HKV=2, D=16)
    assert len(tree.nodes) == 64 and len(tree.root.kv_indices) == 2048
    assert all(len(n.kv_indices) == 1 for n in tree.nodes.values() if n is not tree.root)
    assert tree.root.kv_indices == list(range(2048))
    m = TreeMetadata.from_tree_cache(tree)
    assert m.query_num == len(tree.leaves) == 29 and m.total_kv_len == 2048 + 63


def _tables_equal(a, b):
    """Two (data, directory, scalars) results of build_tables_host hold the same bytes."""
    assert np.array_equal(a[1], b[1]), "directory"
    assert np.array_equal(a[2], b[2]), "scalars"
    assert np.array_equal(a[0], b[0]), "packed tables"


def _reference_tables_equal(a, b):
    """The twelve int64 tables of the reference and the scalars that describe them."""
    for i in range(12):
        assert a[1][i, 1] == b[1][i, 1], _lib.T_NAMES[i]
        x = np.frombuffer(a[0], dtype=np.int64, count=int(a[1][i, 1]), offset=int(a[1][i, 0]))
        y = np.frombuffer(b[0], dtype=np.int64, count=int(b[1][i, 1]), offset=int(b[1][i, 0]))
        assert np.array_equal(x, y), _lib.T_NAMES[i]
    assert np.array_equal(a[2][:6], b[2][:6])


def test_native_tree_mirror_follows_the_tree(monkeypatch):
    """the C++ mirror (deft_tree_t) is fed one deft_b200_tree_append per alloc() and gives, after every
    kind of change -- alloc, branch, cut, merge with and without pruning, reset_node_KV, pause, a page list edited or
    swapped behind the TreeCache's back -- the tables the flat arrays of a fresh walk give, byte for byte."""
    import random
    from deft_b200.tree_cache import build_tables_host, mirror_flat
    rng = random.Random(3)
    r2t = ReqToTokenPool(size=256, max_context_len=1024, device="cpu")
    kvp = TokenToKVPool(size=16384, dtype=torch.float16, head_num=2, head_dim=16, layer_num=1, device="cpu")
    tree = TreeCache(torch.float16, 2, 16, 1, r2t, kvp, None, True, False)
    tree.init_prompt(torch.arange(1, 301, dtype=torch.int32))

    def check(what):
        flat = mirror_flat([tree])
        assert flat is not None and "trees" in flat
        before = tree.native_tree().syncs
        got = build_tables_host(flat, hkv=2)
        want = build_tables_host(_fresh_flat(tree), hkv=2)
        _reference_tables_equal(got, want)       # (the native tiles of a tree that only grew keep their order: test_tables.py)
        assert flat["leaf_to_q"] == {lid: i for i, lid in enumerate(sorted(tree.leaves))}, what
        assert _lib.lib.deft_b200_tree_pages(tree.native_tree().handle) == sum(
            len(n.kv_indices) for n in tree.nodes.values() if not n.paused)

    def step():
        for leaf in tree.leaves.values():
            leaf.append_token(1)
        return tree.alloc()

    check("prompt")
    syncs = tree.native_tree().syncs
    for _ in range(10):                         # a decode loop: alloc only -> no further hand-over of the tree
        step()
        check("alloc")
    assert tree.native_tree().syncs == syncs
    for it in range(50):
        op = rng.random()
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        if op < 0.25 and len(leaves) < 40:
            tree.branch(rng.choice(leaves), rng.choice((2, 3, 5)))
            what = "branch"
        elif op < 0.35 and len(leaves) > 2:
            tree.cut(rng.choice(leaves))
            what = "cut"
        elif op < 0.45 and len(leaves) > 3:
            a, b = rng.sample(leaves, 2)
            tree.merge_nodes(a, b, pruneB_flag=rng.random() < 0.5)
            what = "merge"
        elif op < 0.50:
            rng.choice(leaves).kv_indices.append(kvp.alloc(1).tolist()[0])       # edited in place, behind the tree's back
            what = "foreign append"
        elif op < 0.55:
            leaf = rng.choice(leaves)
            leaf.kv_indices = list(leaf.kv_indices) + kvp.alloc(2).tolist()      # swapped for another list
            what = "foreign swap"
        elif op < 0.65:
            rng.choice(leaves).paused = False    # (the reference never pauses a node; the setter still moves the epoch)
            what = "pause epoch"
        else:
            what = "alloc"
        step()
        check(what)
        step()                                  # (and the append path right after the re-sync)
        check(what + " + alloc")
    leaf = sorted(tree.leaves.values(), key=lambda x: x.id)[0]
    leaf.kv_indices[-1], leaf.kv_indices[-2] = leaf.kv_indices[-2], leaf.kv_indices[-1]      # same length: only an explicit call tells
    syncs = tree.native_tree().syncs
    tree.invalidate_native_tree()
    check("rewritten in place")
    assert tree.native_tree().syncs == syncs + 1
    monkeypatch.setenv("DEFT_NATIVE_TREE", "0")
    assert mirror_flat([tree]) is None          # the switch: flat arrays, as for a foreign tree object


def test_native_tree_mirrors_of_a_forest_and_of_tree_index_mode():
    from deft_b200.tree_cache import build_tables_host, flatten_forest, flatten_tree, mirror_flat
    from deft_b200.workloads import build_forest
    trees = build_forest("cfg3", 5, layers=1, device=torch.device("cpu"), headroom=64 * 8)
    for it in range(4):
        locs = []
        for t in trees:
            for leaf in t.leaves.values():
                leaf.append_token(1)
            locs.append(t.alloc().cache_loc)
        fresh = torch.cat(locs).numpy()
        flat = mirror_flat(trees)
        (_tables_equal if it in (0, 2) else _reference_tables_equal)(        # (0, 2: right after a hand-over of every / one tree)
            build_tables_host(flat, hkv=8, fresh_page=fresh), build_tables_host(flatten_forest(trees), hkv=8, fresh_page=fresh))
        assert flat["leaf_to_q"] == flatten_forest(trees)["leaf_to_q"]
        if it == 1:
            trees[2].branch(sorted(trees[2].leaves.values(), key=lambda x: x.id)[0], 2)   # one tree of the forest changes shape
    assert sum(t.native_tree().syncs for t in trees) == 6        # one hand-over per tree + one for the branch
    # tree-index mode: the index rows travel with the mirror
    r2t = ReqToTokenPool(size=64, max_context_len=512, device="cpu")
    kvp = TokenToKVPool(size=4096, dtype=torch.float16, head_num=2, head_dim=16, layer_num=1, device="cpu")
    tix = TreeIndexPool(size=32, max_context_len=512, device="cpu")
    tree = TreeCache(torch.float16, 2, 16, 1, r2t, kvp, tix, True, True)
    tree.init_prompt(torch.arange(1, 150, dtype=torch.int32))
    tree.branch(tree.root, 3)
    BLOCK_CONFIG["MAX_BLOCK_LEN"] = 128
    try:
        for _ in range(3):
            for leaf in tree.leaves.values():
                leaf.append_token(1)
            tree.alloc()
            _reference_tables_equal(build_tables_host(mirror_flat([tree]), max_block_len=128, tree_index_max_ctx=512, hkv=2),
                                    build_tables_host(flatten_tree(tree), max_block_len=128, tree_index_max_ctx=512, hkv=2))
        assert tree.native_tree().syncs == 1
    finally:
        BLOCK_CONFIG["MAX_BLOCK_LEN"] = -1
