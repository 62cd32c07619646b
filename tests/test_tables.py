"""C++ metadata builder (deft_b200_build_tables) vs the reference's tables and the oracle.  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import deft_oracle as orc
from oracle.plain_tree import PlainNode, PlainTree, thaw
from oracle.scenarios import SCENARIOS, TABLE_SCENARIOS

from deft_b200 import _lib
from deft_b200.tree_cache import build_tables_host, flatten_tree

TABLE_KEYS = _lib.T_NAMES[:12]
ITEM = np.dtype([("kv_off", "<i8"), ("kv_len", "<i4"), ("grp_off", "<i4"), ("n_grp", "<i4"), ("cost", "<i4")])
GROUP = np.dtype([("mask_off", "<i8"), ("q_off", "<i4"), ("q_cnt", "<i4"), ("part_base", "<i4"), ("pad", "<i4")])
UNIT = np.dtype([("kv_off", "<i8"), ("mask_off", "<i8", (2,)), ("kv_tile_stride", "<i4"), ("mask_tile_stride", "<i4"),
                 ("n_tiles", "<i4"), ("last_len", "<i4"), ("q_off", "<i4", (2,)), ("q_cnt", "<i4", (2,)),
                 ("part_base", "<i4", (2,)), ("page0", "<i4"), ("q_id0", "<i4", (2,)), ("dense_tiles", "<i4")])
JOB = np.dtype([("job", "<i4"), ("n_jobs", "<i4"), ("next", "<i4"), ("shared", "<i4"), ("unit", UNIT)])


def unpack(data, directory):
    out = {}
    for i, name in enumerate(_lib.T_NAMES):
        off, cnt = int(directory[i, 0]), int(directory[i, 1])
        dt = np.dtype("<i8") if i < 12 else (ITEM if name.endswith("items") else GROUP if name.endswith("groups")
                                             else UNIT if name == "u_units" else JOB if name == "u_jobs"
                                             else np.dtype("<u4") if name == "u_mask"
                                             else np.dtype("<i4"))
        out[name] = np.frombuffer(data, dtype=dt, count=cnt, offset=off)
    return out


def build(tree, mbl=-1, **kw):
    data, directory, scalars = build_tables_host(flatten_tree(tree), max_block_len=mbl, **kw)
    return unpack(data, directory), scalars


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    return z, thaw({k[5:]: z[k] for k in z.files if k.startswith("tree_")})


def test_struct_sizes():
    assert ITEM.itemsize == _lib.ITEM_BYTES and GROUP.itemsize == _lib.GROUP_BYTES and UNIT.itemsize == _lib.UNIT_BYTES and JOB.itemsize == _lib.JOB_BYTES


def check_unit_plan(t, scalars, tree, hkv, n_ctas):
    """The native unit plan attends exactly the (query, page) pairs of the tree (or of every tree of a
    forest), each once; its CSR, partial-row bases and per-CTA job lists are consistent."""
    units, u_kv, u_mask, u_q = t["u_units"], t["u_kv"], t["u_mask"], t["u_q"]
    nq = int(scalars[0])
    paths = [p for tr in tree for p in orc.leaf_paths(tr)] if isinstance(tree, (list, tuple)) else orc.leaf_paths(tree)
    assert len(paths) == nq
    want = {(q, int(pg)) for q, path in enumerate(paths) for pg in path}
    got = set()
    bases = []
    row_to_q = {}
    attending = set()       # partial rows with at least one attended token
    for u in units:
        assert u["n_tiles"] >= 1 and 1 <= u["last_len"] <= 128 and u["kv_tile_stride"] == 128
        assert 1 <= u["q_cnt"][0] <= 32 and 0 <= u["q_cnt"][1] <= 32
        if u["page0"] >= 0:              # shortcut: full tiles on consecutive pages
            n_tok = int(u["n_tiles"]) * 128
            assert u["last_len"] == 128
            assert u_kv[u["kv_off"]: u["kv_off"] + n_tok].tolist() == list(range(int(u["page0"]), int(u["page0"]) + n_tok))
        for s in range(2):
            cnt = int(u["q_cnt"][s])
            if cnt == 0:
                continue
            assert u["part_base"][s] % 32 == 0
            bases.append(int(u["part_base"][s]))
            qs = u_q[u["q_off"][s]: u["q_off"][s] + cnt]
            if u["q_id0"][s] >= 0:       # shortcut: consecutive query ids
                assert qs.tolist() == list(range(int(u["q_id0"][s]), int(u["q_id0"][s]) + cnt))
            for r, q in enumerate(qs):
                row_to_q[int(u["part_base"][s]) + r] = int(q)
            for tile in range(int(u["n_tiles"])):
                tlen = int(u["last_len"]) if tile == u["n_tiles"] - 1 else 128
                pages = u_kv[u["kv_off"] + tile * 128: u["kv_off"] + tile * 128 + tlen]
                if u["mask_off"][s] < 0:
                    words = np.full(tlen, 0xffffffff, dtype=np.uint32)
                else:
                    m0 = int(u["mask_off"][s]) + tile * int(u["mask_tile_stride"])
                    words = u_mask[m0: m0 + tlen]
                for n in range(tlen):
                    for r in range(cnt):
                        if (int(words[n]) >> r) & 1:
                            pair = (int(qs[r]), int(pages[n]))
                            assert pair not in got, ("attended twice", pair)
                            got.add(pair)
                            attending.add(int(u["part_base"][s]) + r)
    assert got == want
    assert sorted(bases) == [32 * i for i in range(int(scalars[6]))]
    off, rows = t["u_csr_off"], t["u_csr_rows"]
    assert len(off) == nq + 1 and off[0] == 0 and off[-1] == len(rows)
    assert sorted(rows.tolist()) == sorted(attending), "the CSR lists exactly the partial rows that attend something"
    for q in range(nq):
        mine = rows[off[q]: off[q + 1]]
        assert len(mine) >= 1 and np.all(np.diff(mine) > 0) and all(row_to_q[int(r)] == q for r in mine)
    job_off, jobs = t["u_job_off"], t["u_jobs"]
    assert int(scalars[7]) == n_ctas and len(job_off) == n_ctas + 1 and job_off[0] == 0 and np.all(np.diff(job_off) >= 0)
    # records [0, n_ctas): every CTA's first job (with its job count and where its other records start)
    listed = []
    for c in range(n_ctas):
        first = jobs[c]
        n = int(job_off[c + 1] - job_off[c])
        assert (first["job"] >= 0) == (n > 0) and int(first["n_jobs"]) == n
        mine = [first] + [jobs[int(first["next"]) + i] for i in range(n - 1)] if n else []
        for i, r in enumerate(mine):
            listed.append(int(r["job"]))
            assert r["unit"] == units[(int(r["job"]) >> 1) // hkv], "the record carries a copy of its unit"
            if r["shared"]:      # the pair CTA holds the other slot of the same (unit, kv-head) at the same position
                peer = jobs[c ^ 1] if i == 0 else jobs[int(jobs[c ^ 1]["next"]) + i - 1]
                assert peer["shared"] and int(peer["job"]) == int(r["job"]) ^ 1 and n_ctas % 2 == 0
    assert len(jobs) == n_ctas + sum(max(0, int(d) - 1) for d in np.diff(job_off))
    want_jobs = [((ui * hkv + h) << 1) | k for ui, u in enumerate(units) for h in range(hkv) for k in range(2) if u["q_cnt"][k] > 0]
    assert sorted(listed) == sorted(want_jobs), "every (unit, kv-head, live slot) job exactly once"


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_unit_plan_covers_the_tree(golden_dir, name, monkeypatch):
    z, tree = load(golden_dir, name)
    for pair in ("0", "1"):                 # independent job lists / pair-aligned lists (cluster multicast)
        monkeypatch.setenv("DEFT_PLAN_PAIR", pair)
        for mbl, hkv, n_ctas in ((-1, 2, 148), (128, 8, 16)):
            t, scalars = build(tree, mbl, hkv=hkv, n_ctas=n_ctas)
            check_unit_plan(t, scalars, tree, hkv, n_ctas)
            if pair == "1" and (t["u_units"]["q_cnt"][:, 1] > 0).any():
                assert t["u_jobs"]["shared"].any(), "two-slot units exist: their jobs are paired"


def test_forest_plan_covers_every_tree(golden_dir):
    """Several independent trees in one set of tables (batched decoding): queries are numbered tree after tree,
    every tree keeps exactly its own (query, page) pairs, nothing leaks across trees."""
    from deft_b200.tree_cache import flatten_forest
    trees = [load(golden_dir, n)[1] for n in ("toy_binary", "wide40", "single_seq", "ragged_cut")]
    flat = flatten_forest(trees)
    assert int((flat["parent"] == -1).sum()) == len(trees)
    data, directory, scalars = build_tables_host(flat, hkv=2, n_ctas=148)
    t = unpack(data, directory)
    assert int(scalars[0]) == sum(len(tr.leaves) for tr in trees)
    check_unit_plan(t, scalars, trees, 2, 148)
    check_plan(t, int(scalars[0]), "flat")
    check_plan(t, int(scalars[0]), "node")
    # the Node tables of a forest are the trees' Node tables one after the other (queries offset)
    singles = [build(tr)[0] for tr in trees]
    assert t["node_kv"].tolist() == [int(x) for s1 in singles for x in s1["node_kv"]]
    q_base = np.cumsum([0] + [len(tr.leaves) for tr in trees])
    assert t["node_q"].tolist() == [int(x) + int(q_base[i]) for i, s1 in enumerate(singles) for x in s1["node_q"]]


def test_unit_plan_cfg2_shape(golden_dir):
    """cfg2: the 4096-token prompt is one chain of dense tiles cut into balanced pieces; 148 CTAs all get work."""
    z, tree = load(golden_dir, "cfg2_tables")
    t, scalars = build(tree, -1, hkv=8, n_ctas=148)
    units = t["u_units"]
    # 64 queries = one pair of slots: ONE chain of all 48 tiles (prompt + subtree), cut into pieces
    assert np.all(units["q_cnt"][:, 0] == 32) and np.all(np.isin(units["q_cnt"][:, 1], (0, 32)))
    assert int(units["n_tiles"].sum()) == 48
    assert np.array_equal(np.sort(units["kv_off"]), np.cumsum(np.r_[0, units["n_tiles"][np.argsort(units["kv_off"])][:-1]]) * 128)
    root = units[units["kv_off"] + units["n_tiles"] * 128 <= 4096]
    assert len(root) > 0 and np.all(root["mask_off"] == -1), "prompt-only pieces: dense, no mask reads"
    assert np.all(root["page0"] == root["kv_off"]) and np.all(units["q_id0"][:, 0] >= 0), "prompt pages and leaf ids are runs"
    loads = np.diff(t["u_job_off"])
    n_jobs = 8 * int((units["q_cnt"] > 0).sum())           # one job per (unit, kv-head, live slot)
    assert loads.max() <= 2 and int((loads > 0).sum()) >= 140 and loads.sum() == n_jobs, "cfg2: about one job per CTA"
    # the subtree's tokens are regrouped: the deepest level (32 neighbouring leaves x 16 decode steps per slot) sits on
    # runs of 32 consecutive pages -> whole tiles of TMA boxes, the level above on runs of 16, ...; every token still
    # appears exactly once, dummy tokens (page -1) pad a slot's group to whole blocks
    sub = t["u_kv"][4096:]
    assert int(sum((np.diff(b) == 1).all() for b in sub[: len(sub) // 32 * 32].reshape(-1, 32))) >= 32
    assert int(sum((np.diff(b) == 1).all() for b in sub[: len(sub) // 8 * 8].reshape(-1, 8))) >= 32 * 4 + 16 * 2 + 16
    live = sub[: (int(units["n_tiles"].sum()) - 33) * 128 + int(units["last_len"][np.argmax(units["kv_off"])])]
    assert sorted(live[live >= 0].tolist()) == list(range(4096, 4096 + 2016)) and int((live < 0).sum()) <= 31
    # far fewer partial rows than the reference's 2246 (one per (sub-block, query))
    assert len(t["u_csr_rows"]) < 1400


@pytest.mark.parametrize("name", list(SCENARIOS) + list(TABLE_SCENARIOS))
def test_tables_match_reference_bit_exact(golden_dir, name):
    z, tree = load(golden_dir, name)
    for prefix, mbl in (("t_", -1), ("tc_", 128)):
        t, scalars = build(tree, mbl)
        for k in TABLE_KEYS:
            assert np.array_equal(t[k], z[prefix + k]), (name, prefix, k)
        assert scalars[:4].tolist() == z[prefix + "scalars"].tolist()


def test_tree_index_tables_match_reference(golden_dir):
    z, tree = load(golden_dir, "tree_index")
    t, _ = build(tree, 128, tree_index_max_ctx=int(z["geom"][4]))
    for k in ["node_q", "node_q_len", "node_q_offset", "node_kv_offset", "node_kv_len"]:
        assert np.array_equal(t[k], z["ti_" + k]), k


def check_plan(t, nq, kind):
    """Structural invariants of the native plan."""
    items, groups = t[f"{kind}_items"], t[f"{kind}_groups"]
    off, rows = t[f"{kind}_csr_off"], t[f"{kind}_csr_rows"]
    q_list = t["block_q"] if kind == "flat" else t["node_q"]
    assert len(off) == nq + 1 and off[0] == 0 and off[-1] == len(rows)
    # groups are covered by items exactly once, in order
    assert items["grp_off"].tolist() == np.concatenate([[0], np.cumsum(items["n_grp"])[:-1]]).tolist()
    assert int(items["n_grp"].sum()) == len(groups)
    # every partial row appears exactly once in the CSR, under the right query, ascending per query
    row_to_q = {}
    for g in groups:
        assert 1 <= g["q_cnt"] <= 32
        for r in range(g["q_cnt"]):
            row_to_q[int(g["part_base"]) + r] = int(q_list[g["q_off"] + r])
    assert sorted(row_to_q) == sorted(rows.tolist()) and len(set(rows.tolist())) == len(rows)
    for q in range(nq):
        mine = rows[off[q]: off[q + 1]]
        assert np.all(np.diff(mine) > 0)
        assert all(row_to_q[int(r)] == q for r in mine)


@pytest.mark.parametrize("name", list(SCENARIOS) + list(TABLE_SCENARIOS))
def test_plan_invariants(golden_dir, name):
    z, tree = load(golden_dir, name)
    for mbl in (-1, 128):
        t, scalars = build(tree, mbl)
        nq = int(scalars[0])
        check_plan(t, nq, "flat")
        check_plan(t, nq, "node")
        # Flatten items = unique KV blocks: one item per distinct block start, all sub-blocks grouped
        bk = t["block_kv"].reshape(-1, 128)
        firsts = [0] + [b for b in range(1, len(bk)) if bk[b, 0] != bk[b - 1, 0]]
        assert (t["flat_items"]["kv_off"] // 128).tolist() == firsts
        assert t["flat_items"]["kv_len"].tolist() == t["block_lens"][firsts].tolist()
        # Node items tile every entry's KV range with <= 256-token pieces
        covered = sum(int(x) for x in t["node_items"]["kv_len"])
        assert covered == int(t["node_kv_len"].sum())
        assert int(t["node_items"]["kv_len"].max()) <= 256


def random_tree(rng: random.Random) -> PlainTree:
    """Random topology, ragged node lengths around the 128 edge, shuffled page ids, leaf ids != DFS order."""
    t = PlainTree()
    depth = rng.randint(0, 4)
    ids = list(range(1, 4000))
    rng.shuffle(ids)
    root = PlainNode(0)
    t.root = root
    t.nodes[0] = root
    frontier = [root]
    for _ in range(depth):
        nxt = []
        for n in frontier:
            if n is not root and rng.random() < 0.25:
                continue
            for _ in range(rng.choice([1, 2, 2, 3, 7, 40 if len(frontier) < 3 else 2])):
                c = PlainNode(ids.pop())
                c.parent = n
                n.children[c.id] = c
                t.nodes[c.id] = c
                nxt.append(c)
        frontier = nxt or frontier
    pages = list(range(200000))
    rng.shuffle(pages)
    for n in t.nodes.values():
        ln = rng.choice([1, 1, 2, 16, 127, 128, 129, 300, rng.randint(1, 700)])
        n.kv_indices = [pages.pop() for _ in range(ln)]
    for n in t.nodes.values():
        if not n.children:
            t.leaves[n.id] = n
            cur = n
            while cur is not None:
                cur.refs.add(n)
                cur = cur.parent
    return t


@pytest.mark.parametrize("seed", range(40))
def test_random_trees_match_oracle(seed):
    rng = random.Random(seed)
    tree = random_tree(rng)
    for mbl in (-1, 128):
        t, scalars = build(tree, mbl)
        want = orc.build_tables(tree, max_block_len=mbl)
        for k in TABLE_KEYS:
            assert np.array_equal(t[k], want[k]), (seed, mbl, k)
        assert scalars[:4].tolist() == [want["query_num"], want["node_num"], want["total_kv_len"], 128]
        check_plan(t, want["query_num"], "flat")
        check_plan(t, want["query_num"], "node")
    if seed % 4 == 0:
        t, scalars = build(tree, -1, hkv=2, n_ctas=7)
        check_unit_plan(t, scalars, tree, 2, 7)


def test_builder_rejects_malformed_trees():
    tree = PlainTree()
    root = PlainNode(0)
    tree.root = root
    tree.nodes[0] = root
    tree.leaves[0] = root
    root.refs.add(root)
    root.kv_indices = []                 # the reference crashes on this (range step 0); we report
    with pytest.raises(_lib.DeftError, match="no KV pages"):
        build(tree)
    root.kv_indices = [3, 1, 2]
    t, scalars = build(tree)
    assert t["node_kv"].tolist() == [1, 2, 3] and t["block_lens"].tolist() == [3]
    assert t["block_kv"][:4].tolist() == [1, 2, 3, -1] and scalars[2] == 3


def test_capacity_padded_layout_keeps_offsets_while_the_tree_grows():
    """A decode loop appends a page per leaf every step (tree_generate.py:109): every table grows a little.  Built with
    a TableLayout the packed buffer keeps its offsets for many steps in a row (what a captured CUDA graph of the step
    needs), every table holds exactly what the tight packing holds, and a table that outgrows its region moves the
    layout's version on."""
    from deft_b200.tree_cache import TableLayout
    from deft_b200.workloads import build_tree
    tree = build_tree("cfg2", layers=0, device="cpu", H=4, HKV=2, D=16, headroom=64 * 50)
    lay = TableLayout()
    offsets, versions = [], []
    for step in range(48):
        for leaf in tree.leaves.values():
            leaf.append_token(7)
        tree.alloc()
        flat = flatten_tree(tree)
        v0 = lay.version
        d_pad, dir_pad, sc_pad = build_tables_host(flat, hkv=2, n_ctas=148, layout=lay)
        d_tight, dir_tight, sc_tight = build_tables_host(flat, hkv=2, n_ctas=148)
        assert np.array_equal(dir_pad[:, 1], dir_tight[:, 1]) and np.array_equal(sc_pad[:9], sc_tight[:9])
        assert sc_pad[9] >= sc_tight[9] == sc_tight[6], "the slot capacity covers the slots in use"
        a, b = unpack(d_pad, dir_pad), unpack(d_tight, dir_tight)
        for name in _lib.T_NAMES:
            assert np.array_equal(a[name], b[name]), (step, name)
        assert np.all(np.diff(dir_pad[:, 0]) >= 0) and np.all(dir_pad[:, 0] % 256 == 0)
        ends = dir_pad[:-1, 0] + dir_pad[:-1, 1] * np.array([a[n].dtype.itemsize for n in _lib.T_NAMES[:-1]])
        assert np.all(ends <= dir_pad[1:, 0]), "no table runs into the next region"
        if offsets and tuple(dir_pad[:, 0]) != offsets[-1]:
            assert lay.version > v0, "offsets only move with the version"
        offsets.append(tuple(dir_pad[:, 0]))
        versions.append(lay.version)
    assert len(set(offsets)) <= 5, "48 decode steps (the subtree's KV grows 2.5x): a handful of layouts"
    assert len(set(versions)) == len(set(offsets))


@pytest.mark.parametrize("name", ["toy_binary", "wide40", "ragged_cut", "llama_flat8"])
def test_fresh_tokens_of_a_decode_step_are_marked_for_the_fused_append(golden_dir, name):
    """Tables built with ``fresh_page`` (the pages TreeCache.alloc just handed out): in the native tables those tokens
    name their QUERY (bit 30 + query id: the kernel reads the row from the step's activations), they sit in chunks of
    their own (a gather4 instruction takes four rows of ONE tensor), the load descriptors say so (bit 27), and with
    the pages put back the plan covers the tree exactly as the plain build does.  The reference tables do not change."""
    z, tree = load(golden_dir, name)
    leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
    fresh = np.asarray([leaf.kv_indices[-1] for leaf in leaves], dtype=np.int32)      # the last page of every leaf
    flat = flatten_tree(tree)
    plain, dir0, sc0 = build_tables_host(flat, hkv=2, n_ctas=148)
    data, directory, scalars = build_tables_host(flat, hkv=2, n_ctas=148, fresh_page=fresh)
    t0, t = unpack(plain, dir0), unpack(data, directory)
    assert int(scalars[10]) == 1 and int(sc0[10]) == 0
    for k in TABLE_KEYS:
        assert np.array_equal(t[k], t0[k]), k
    FRESH = 1 << 30
    u_kv = t["u_kv"].copy()
    is_fresh = (u_kv >= 0) & ((u_kv & FRESH) != 0)
    assert sorted((u_kv[is_fresh] & ~FRESH).tolist()) == list(range(len(leaves))), "every query's new token exactly once"
    for c4 in np.flatnonzero(is_fresh.reshape(-1, 4).any(axis=1)):       # gather granularity: all fresh, or dummies
        four = u_kv[4 * c4: 4 * c4 + 4]
        assert np.all(((four & FRESH) != 0) | (four < 0)), four
    blk = t["u_blk"]
    for c8 in range(len(blk)):
        eight = u_kv[8 * c8: 8 * c8 + 8]
        kind, fr, first = (int(blk[c8]) >> 28) & 3, (int(blk[c8]) >> 27) & 1, int(blk[c8]) & ((1 << 27) - 1)
        if kind:
            assert fr == int((eight[0] & FRESH) != 0) and first == int(eight[0]) & ~FRESH
            assert np.array_equal(eight, eight[0] + np.arange(8)), "a box chunk is a run"
    decoded = u_kv.copy()
    decoded[is_fresh] = fresh[u_kv[is_fresh] & ~FRESH]
    check_unit_plan({**t, "u_kv": decoded}, scalars, tree, 2, 148)


def _decode_fresh(t, fresh):
    """Native tables with the tokens of this step put back as pages + the layout rules of fresh tokens and box chunks."""
    FRESH = 1 << 30
    u_kv = t["u_kv"].copy()
    is_fresh = (u_kv >= 0) & ((u_kv & FRESH) != 0)
    if fresh is None:
        assert not is_fresh.any()
    else:
        assert sorted((u_kv[is_fresh] & ~FRESH).tolist()) == list(range(len(fresh))), "every query's new token exactly once"
        # past the live tokens of a tree's last tile: zeros (nobody attends them; such a tile can sit in the middle of a
        # chain when small trees share a slot)
        tiles = u_kv.reshape(-1, 128)
        n_live = np.where((tiles != 0).any(axis=1), 128 - np.argmax(tiles[:, ::-1] != 0, axis=1), 0)
        tail = (np.arange(128)[None, :] >= n_live[:, None]).reshape(-1)
        for c4 in np.flatnonzero(is_fresh.reshape(-1, 4).any(axis=1)):       # gather granularity: all fresh, or dummies
            four = u_kv[4 * c4: 4 * c4 + 4]
            assert np.all(((four & FRESH) != 0) | (four < 0) | tail[4 * c4: 4 * c4 + 4]), four
    blk = t["u_blk"]
    kinds = np.zeros(4, dtype=np.int64)
    for c8 in range(len(blk)):
        eight = u_kv[8 * c8: 8 * c8 + 8]
        kind, fr, first = (int(blk[c8]) >> 28) & 3, (int(blk[c8]) >> 27) & 1, int(blk[c8]) & ((1 << 27) - 1)
        kinds[kind] += 1
        if kind:
            assert fr == int((eight[0] & FRESH) != 0) and first == int(eight[0]) & ~FRESH
            assert np.array_equal(eight, eight[0] + np.arange(8)), "a box chunk is a run"
    decoded = u_kv.copy()
    if fresh is not None:
        decoded[is_fresh] = fresh[u_kv[is_fresh] & ~FRESH]
    return {**t, "u_kv": decoded}, kinds


@pytest.mark.parametrize("seed", range(6))
def test_kept_tile_order_of_a_growing_tree(seed):
    """SURVEY 8(f).1, incremental tables: a tree that only grows by alloc() keeps the token order of its native tiles
    in its C++ mirror; a build lays out the appended tokens only.  After every step -- with the step's tokens read from
    the activations or from the pool, on trees of every shape, across branches and cuts (which void the kept order) --
    the reference tables are the flat-array builder's byte for byte and the native plan attends every (query, page)
    pair exactly once; the tiles are as well loadable (TMA boxes against gathers) as the ones laid out from scratch."""
    import torch
    from deft_b200.memory_pool import ReqToTokenPool, TokenToKVPool
    from deft_b200.tree_cache import TreeCache, mirror_flat
    rng = random.Random(seed)
    r2t = ReqToTokenPool(size=512, max_context_len=4096, device="cpu")
    kvp = TokenToKVPool(size=1 << 16, dtype=torch.float16, head_num=2, head_dim=16, layer_num=1, device="cpu")
    tree = TreeCache(torch.float16, 2, 16, 1, r2t, kvp, None, True, False)
    tree.init_prompt(torch.arange(rng.choice([5, 130, 300, 700]), dtype=torch.int32))

    def step():
        for leaf in tree.leaves.values():
            leaf.append_token(1)
        return tree.alloc().cache_loc.numpy().astype(np.int32)

    def branch_some(width):
        for leaf in sorted(tree.leaves.values(), key=lambda x: x.id):
            if len(tree.leaves) < 90 and rng.random() < 0.7:
                tree.branch(leaf, rng.choice(width))

    branch_some((2, 3))
    for _ in range(rng.choice([1, 16])):
        step()
    branch_some((2, 3, 4) if seed % 2 else (2,))
    from deft_b200.tree_cache import TableLayout
    lean_layout = TableLayout(native_only=True)
    kept_steps = 0
    for it in range(45):
        if it in (17, 31):                     # the topology moves: the mirror is handed over again, the order made anew
            leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
            if it == 17:
                tree.branch(rng.choice(leaves), 2)
            elif len(leaves) > 2:
                tree.cut(rng.choice(leaves))
        fresh = step()
        use_fresh = rng.random() < 0.7
        syncs = tree.native_tree().syncs
        got = build_tables_host(mirror_flat([tree]), hkv=2, n_ctas=148, fresh_page=fresh if use_fresh else None)
        want = build_tables_host(flatten_tree(tree), hkv=2, n_ctas=148, fresh_page=fresh if use_fresh else None)
        kept_steps += int(tree.native_tree().syncs == syncs)
        t, t0 = unpack(got[0], got[1]), unpack(want[0], want[1])
        for k in TABLE_KEYS:
            assert np.array_equal(t[k], t0[k]), (it, k)
        # a native-only layout on the same mirror (no token stream is made at all when the kept order serves): same unit plan
        lean = build_tables_host(mirror_flat([tree]), hkv=2, n_ctas=148, fresh_page=fresh if use_fresh else None, layout=lean_layout)
        tl = unpack(lean[0], lean[1])
        for name in _lib.T_NAMES[20:]:
            assert np.array_equal(t[name], tl[name]), (it, name)
        assert all(len(tl[k]) == 0 for k in TABLE_KEYS) and lean[2][2] == got[2][2]
        dec, kinds = _decode_fresh(t, fresh if use_fresh else None)
        check_unit_plan(dec, got[2], tree, 2, 148)
        _, kinds0 = _decode_fresh(t0, fresh if use_fresh else None)
        # load instructions per panel: box32 = 1/4 per chunk of 8, box16 = 1/2, box8 = 1, gathered = 2
        cost, cost0 = (k @ np.array([2.0, 1.0, 0.5, 0.25]) for k in (kinds, kinds0))
        assert cost <= 1.25 * cost0 + 8, (it, kinds, kinds0)
        if rng.random() < 0.3:                 # a second build of the same step, the other way round
            again = build_tables_host(mirror_flat([tree]), hkv=2, n_ctas=148, fresh_page=None if use_fresh else fresh)
            dec2, _ = _decode_fresh(unpack(again[0], again[1]), None if use_fresh else fresh)
            check_unit_plan(dec2, again[2], tree, 2, 148)
    assert kept_steps >= 40


def test_kept_tile_order_of_a_forest():
    """The same for several trees over one pool: every tree keeps its own order (slots are counted over the whole
    forest, so a tree whose first query moves -- a neighbour branched -- starts over)."""
    import torch
    from deft_b200.memory_pool import ReqToTokenPool, TokenToKVPool
    from deft_b200.tree_cache import TreeCache, flatten_forest, mirror_flat
    rng = random.Random(11)
    r2t = ReqToTokenPool(size=512, max_context_len=4096, device="cpu")
    kvp = TokenToKVPool(size=1 << 16, dtype=torch.float16, head_num=2, head_dim=16, layer_num=1, device="cpu")
    trees = []
    for i in range(4):
        tree = TreeCache(torch.float16, 2, 16, 1, r2t, kvp, None, True, False)
        tree.init_prompt(torch.arange((40, 260, 131, 512)[i], dtype=torch.int32))
        for _ in range(2):
            for leaf in sorted(tree.leaves.values(), key=lambda x: x.id):
                if rng.random() < 0.8:
                    tree.branch(leaf, rng.choice((2, 3, 5)))
            for leaf in tree.leaves.values():
                leaf.append_token(1)
            tree.alloc()
        trees.append(tree)
    for it in range(24):
        if it == 9:
            trees[1].branch(sorted(trees[1].leaves.values(), key=lambda x: x.id)[0], 3)   # the queries of trees 2, 3 move up
        if it == 15:       # pages come back to the pool: the next alloc of EVERY tree takes them (lower than what it holds)
            trees[0].cut(sorted(trees[0].leaves.values(), key=lambda x: x.id)[-1])
        locs = []
        for tree in trees:
            for leaf in tree.leaves.values():
                leaf.append_token(1)
            locs.append(tree.alloc().cache_loc.numpy().astype(np.int32))
        fresh = np.concatenate(locs) if it % 3 else None
        got = build_tables_host(mirror_flat(trees), hkv=2, n_ctas=148, fresh_page=fresh)
        want = build_tables_host(flatten_forest(trees), hkv=2, n_ctas=148, fresh_page=fresh)
        t, t0 = unpack(got[0], got[1]), unpack(want[0], want[1])
        for k in TABLE_KEYS:
            assert np.array_equal(t[k], t0[k]), (it, k)
        dec, _ = _decode_fresh(t, fresh)
        check_unit_plan(dec, got[2], trees, 2, 148)
    assert sum(t.native_tree().syncs for t in trees) == 6


def test_native_only_layout_leaves_the_reference_tables_empty():
    """deft_b200_layout_set_native_only: the unit plan (all the tensor-core path reads) is byte for byte the one of the
    full build, the twelve reference tables and the item / group plans are empty, the upload is a third."""
    import torch
    from deft_b200.tree_cache import TableLayout
    from deft_b200.workloads import build_tree
    tree = build_tree("cfg2", layers=1, device=torch.device("cpu"), headroom=64 * 8)
    full_l, lean_l = TableLayout(), TableLayout(native_only=True)
    for it in range(3):
        for leaf in tree.leaves.values():
            leaf.append_token(1)
        fresh = tree.alloc().cache_loc.numpy().astype(np.int32)
        flat = flatten_tree(tree)
        full = build_tables_host(flat, hkv=8, n_ctas=148, layout=full_l, fresh_page=fresh)
        lean = build_tables_host(flat, hkv=8, n_ctas=148, layout=lean_l, fresh_page=fresh)
        t, tl = unpack(full[0], full[1]), unpack(lean[0], lean[1])
        for i, name in enumerate(_lib.T_NAMES):
            if i < 20:
                assert len(tl[name]) == (65 if name.endswith("csr_off") else 0), name      # (an empty CSR still has its offsets)
            else:
                assert np.array_equal(t[name], tl[name]), name
        assert lean[2][0] == 64 and lean[2][2] == full[2][2] and lean[2][4] == 0 and lean[2][5] == 0
        assert np.array_equal(full[2][6:], lean[2][6:])
        dec, _ = _decode_fresh(tl, fresh)
        check_unit_plan(dec, lean[2], tree, 8, 148)
        assert len(lean[0]) < 0.45 * len(full[0])


@pytest.mark.parametrize("seed", range(4))
def test_kept_tile_order_under_random_operations(seed):
    """One to three trees over one pool; every step one of branch / cut / merge (or none), a token and a page per leaf,
    then a build from the mirrors -- with or without the step's tokens marked fresh, tight or native-only, for 148 / 16 /
    7 CTAs: the plan attends every (query, page) pair once, the reference tables are the flat-array builder's."""
    import torch
    from deft_b200.memory_pool import ReqToTokenPool, TokenToKVPool
    from deft_b200.tree_cache import TableLayout, TreeCache, flatten_forest, mirror_flat
    rng = random.Random(1000 + seed)
    r2t = ReqToTokenPool(size=2048, max_context_len=4096, device="cpu")
    kvp = TokenToKVPool(size=1 << 17, dtype=torch.float16, head_num=2, head_dim=16, layer_num=1, device="cpu")
    trees = []
    for _ in range(rng.choice([1, 2, 3])):
        tree = TreeCache(torch.float16, 2, 16, 1, r2t, kvp, None, True, False)
        tree.init_prompt(torch.arange(rng.choice([1, 7, 127, 128, 129, 300, 900]), dtype=torch.int32))
        trees.append(tree)
    lean = TableLayout(native_only=True)
    for it in range(40):
        op, tree = rng.random(), rng.choice(trees)
        leaves = sorted(tree.leaves.values(), key=lambda x: x.id)
        if op < 0.12 and sum(len(t.leaves) for t in trees) < 150:
            tree.branch(rng.choice(leaves), rng.choice((2, 3, 4, 7)))
        elif op < 0.18 and len(leaves) > 2:
            tree.cut(rng.choice(leaves))
        elif op < 0.22 and len(leaves) > 3:
            a, b = rng.sample(leaves, 2)
            tree.merge_nodes(a, b, pruneB_flag=True)
        locs = []
        for t_ in trees:
            for leaf in t_.leaves.values():
                leaf.append_token(1)
            locs.append(t_.alloc().cache_loc.numpy().astype(np.int32))
        fresh = np.concatenate(locs) if rng.random() < 0.6 else None
        layout = lean if rng.random() < 0.5 else None
        got = build_tables_host(mirror_flat(trees), hkv=2, n_ctas=rng.choice([148, 16, 7]), fresh_page=fresh, layout=layout)
        t = unpack(got[0], got[1])
        dec, _ = _decode_fresh(t, fresh)
        check_unit_plan(dec, got[2], trees if len(trees) > 1 else trees[0], 2, int(got[2][7]))
        if layout is None:
            want = build_tables_host(flatten_forest(trees) if len(trees) > 1 else flatten_tree(trees[0]), hkv=2, fresh_page=fresh)
            t0 = unpack(want[0], want[1])
            for k in TABLE_KEYS:
                assert np.array_equal(t[k], t0[k]), (it, k)
