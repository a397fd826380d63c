"""Last seconds of the round's GPU budget: the u64 de-duplication by position exchange (parity on a sample + timing on the
500k x 256, 10 % repeats shape) and the lightened fused-insert test, every step printed at once."""
import os, sys, time
t0 = time.time()
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasketch_b200 as dsk
from datasketch_b200.minhash import _make_permutations
from oracle import oracle_clib as oc
print("import %.1fs" % (time.time() - t0), flush=True)
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(1)
n, t, k = 500_000, 256, 128
nt = n * t
perms = _make_permutations(k, 1)
off = torch.arange(0, (n + 1) * t, t, dtype=torch.int64, device=dev)
lo = torch.randint(-2 ** 31, 2 ** 31 - 1, (nt,), dtype=torch.int32, device=dev, generator=g)
hi = torch.randint(0, 1 << 32, (nt,), dtype=torch.int64, device=dev, generator=g)
tok = hi.mul_(1 << 32).add_(lo.to(torch.int64) & 0xFFFFFFFF)
pos = torch.arange(nt, device=dev) % t
for share in (0.0, 0.1, 0.5):
    tk = tok
    if share > 0:
        rep = (torch.rand(nt, device=dev, generator=g) < share) & (pos > 0)
        src = (torch.arange(nt, device=dev) - pos) + (torch.rand(nt, device=dev, generator=g) * pos).long()
        tk = torch.where(rep, tok[src], tok)
    sig = torch.empty((n, k), dtype=torch.int32, device=dev)
    for _ in range(2):
        dsk.engine.bulk_signatures_device(tk, off, nt, perms, d_out=sig)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dsk.engine.bulk_signatures_device(tk, off, nt, perms, d_out=sig)
    e1.record()
    torch.cuda.synchronize()
    idx = np.arange(0, n, 5003)
    sub = tk.view(n, t)[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint64).reshape(-1)
    want = oc.minhash_bulk_u64tok(np.ascontiguousarray(sub), np.arange(len(idx) + 1, dtype=np.int64) * t, perms)
    ok = bool(np.array_equal(sig[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32), want))
    print("u64 tokens 500k x 256, repeat share %.2f: %.3f ms, rows identical: %s" % (share, e0.elapsed_time(e1) / 5, ok), flush=True)
import test_lsh_gpu as tl
t1 = time.time()
tl.test_fused_insert_from_tokens_equals_the_two_step_flow(dsk, 128, (9, 13))
print("fused insert test k=128: passed in %.1fs" % (time.time() - t1), flush=True)
t1 = time.time()
tl.test_fused_insert_from_tokens_equals_the_two_step_flow(dsk, 300, (20, 15))
print("fused insert test k=300: passed in %.1fs" % (time.time() - t1), flush=True)
