import sys, time, torch
sys.path.insert(0, '.')
import deft_b200
from deft_b200 import TreeMetadata, _lib
from deft_b200.workloads import build_tree
dev = torch.device("cuda:0")
L = 32
tree = build_tree("cfg2", layers=L, device=dev)
kvp = tree.token_to_kv_pool
nq = len(tree.leaves)
qkv = torch.randn(L, nq, 6144, dtype=torch.float16, device=dev)
out = torch.empty(L, nq, 32, 128, dtype=torch.float16, device=dev)
m = TreeMetadata.from_tree_cache(tree)
loc = torch.zeros(nq, dtype=torch.int32, device=dev)
def att(l):
    deft_b200.tree_attention_subtree_fwd(qkv[l, :, :4096].view(nq, 32, 128), kvp.get_key_buffer(l), kvp.get_value_buffer(l), out[l], 128,
        m.block_q, m.block_q_cnts, m.block_q_offset, m.block_bitmasks, m.block_kv, m.block_lens)
def app(l):
    deft_b200.kv_append(kvp.kv_data[l], qkv[l, :, 4096:5120].view(nq, 8, 128), qkv[l, :, 5120:].view(nq, 8, 128), loc)
for _ in range(3):
    for l in range(L): att(l); app(l)
torch.cuda.synchronize()
for name, fn in (("attention", att), ("kv_append", app)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        for l in range(L): fn(l)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name}: host {1e6*(t1-t0)/(10*L):.1f} us per call")
t0 = time.perf_counter()
for _ in range(20): m2 = TreeMetadata.from_tree_cache(tree)
t1 = time.perf_counter()
print(f"from_tree_cache: {1e3*(t1-t0)/20:.3f} ms")
from deft_b200.tree_cache import flatten_tree, build_tables_host
t0 = time.perf_counter()
for _ in range(20): f = flatten_tree(tree)
t1 = time.perf_counter()
for _ in range(20): build_tables_host(f, hkv=8, n_ctas=148)
t2 = time.perf_counter()
print(f"flatten_tree {1e3*(t1-t0)/20:.3f} ms, build_tables_host {1e3*(t2-t1)/20:.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    for l in range(L): att(l)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# ---- the C call alone, fixed arguments
import ctypes as C
q0 = qkv[0, :, :4096].view(nq, 32, 128); K0, V0 = kvp.get_key_buffer(0), kvp.get_value_buffer(0); o0 = out[0]
need = _lib.lib.deft_b200_flatten_workspace_bytes(nq, 32, 8, 128, m.block_q.numel(), m.block_q_cnts.numel(), C.byref(m.flat_plan))
ws = torch.empty(need, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
args = (q0.data_ptr(), q0.stride(0), q0.stride(1), K0.data_ptr(), V0.data_ptr(), K0.stride(0), K0.stride(1), K0.shape[0],
        o0.data_ptr(), o0.stride(0), o0.stride(1), nq, 32, 8, 128, 128, m.block_q.data_ptr(), m.block_q.numel(),
        m.block_q_cnts.data_ptr(), m.block_q_offset.data_ptr(), m.block_lens.data_ptr(), m.block_q_cnts.numel(),
        m.block_bitmasks.data_ptr(), m.block_kv.data_ptr(), C.byref(m.flat_plan), ws.data_ptr(), ws.numel(), stream)
for pdl in (1, 0):
    _lib.lib.deft_b200_set_pdl(pdl)
    for stages in (7, 2, 4, 0):
        _lib.lib.deft_b200_set_stages(stages)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(320): _lib.lib.deft_b200_flatten_fwd(*args)
        t1 = time.perf_counter(); torch.cuda.synchronize()
        print(f"C call alone pdl={pdl} stages={stages}: host {1e6*(t1-t0)/320:.1f} us")
_lib.lib.deft_b200_set_stages(7); _lib.lib.deft_b200_set_pdl(1)
