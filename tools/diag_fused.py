"""One-off diagnosis: fused dsk_lsh_insert_tokens vs signatures + insert, small sizes, every step printed at once."""
import sys, os, time
t_start = time.time()
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datasketch_b200 as dsk
from datasketch_b200.minhash import _make_permutations
print("import %.1fs" % (time.time() - t_start), flush=True)
g = torch.Generator(device="cuda").manual_seed(3)
for k, n in ((128, 1000), (128, 100_000), (256, 100_000), (256, 1_000_000)):
    t = 128
    perms = _make_permutations(k, 1)
    tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * t,), dtype=torch.int32, device="cuda", generator=g)
    off = torch.arange(0, (n + 1) * t, t, dtype=torch.int64, device="cuda")
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.time()
    a = dsk.GpuLSH(threshold=0.8, num_perm=k, capacity=n)
    b = dsk.GpuLSH(threshold=0.8, num_perm=k, capacity=n)
    torch.cuda.synchronize()
    print("k=%d n=%d create %.3fs" % (k, n, time.time() - t0), flush=True)
    e[0].record()
    sig = dsk.engine.bulk_signatures_device(tok, off, n * t, perms)
    e[1].record()
    a.insert(sig)
    e[2].record()
    torch.cuda.synchronize()
    print("  two-step: sig %.3f ms insert %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])), flush=True)
    t0 = time.time()
    e[0].record()
    b.insert_tokens(tok, off, n * t, perms)
    e[1].record()
    print("  fused enqueued after %.3fs" % (time.time() - t0), flush=True)
    torch.cuda.synchronize()
    print("  fused: %.3f ms (wall %.3fs)" % (e[0].elapsed_time(e[1]), time.time() - t0), flush=True)
