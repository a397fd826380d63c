"""Same-box timing of several builds of the library (e.g. from earlier commits) on the C2 shape, through ctypes only --
independent of the Python package, so libraries with older symbol tables load too.

    python tools/ab_libs.py lib1.so lib2.so ...       (prefix a path with v1: to set DSK_TWO_PHASE_V1=1 -- needs its own process)
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as o  # noqa: E402   (permutation parameters only)

n, t, k = 1_000_000, 256, 128
perms = o.init_permutations(k, 1)
a, b = np.ascontiguousarray(perms[0]), np.ascontiguousarray(perms[1])
g = torch.Generator(device="cuda").manual_seed(1)
tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * t + n,), dtype=torch.int32, device="cuda", generator=g)
off_aligned = torch.arange(0, (n + 1) * t, t, dtype=torch.int64, device="cuda")
lens = torch.randint(128, 385, (n,), device="cuda", generator=g)
off_ragged = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
off_ragged[1:] = torch.cumsum(lens, 0)
if int(off_ragged[-1]) > tok.numel():
    off_ragged = torch.clamp(off_ragged, max=tok.numel())
out = torch.empty((n, k), dtype=torch.int32, device="cuda")
ref = {}
vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
for rep in range(2):
    for path in sys.argv[1:]:
        lib = ctypes.CDLL(path)
        lib.dsk_perm_create.argtypes = [vp, vp, ci, ci, ctypes.POINTER(vp)]
        lib.dsk_minhash_bulk.argtypes = [vp, vp, ci, vp, i64, i64, vp, i64, ci, vp, ci, ci, vp]
        h = vp()
        assert lib.dsk_perm_create(a.ctypes.data, b.ctypes.data, k, 0, ctypes.byref(h)) == 0
        st = torch.cuda.current_stream().cuda_stream
        ws = None
        if hasattr(lib, "dsk_minhash_bulk_ws"):
            lib.dsk_minhash_bulk_workspace_size.restype = ctypes.c_size_t
            lib.dsk_minhash_bulk_workspace_size.argtypes = [i64, i64]
            lib.dsk_minhash_bulk_ws.argtypes = [vp, vp, ci, vp, i64, i64, vp, i64, ci, vp, ci, ci, vp, ctypes.c_size_t, vp]
            nb = lib.dsk_minhash_bulk_workspace_size(n, tok.numel())
            ws = torch.empty((nb,), dtype=torch.uint8, device="cuda")

        res = []
        for shape, off in (("aligned", off_aligned), ("ragged", off_ragged)):
            nt = int(off[-1].item())

            def step():
                if ws is not None:
                    rc = lib.dsk_minhash_bulk_ws(h, tok.data_ptr(), 0, off.data_ptr(), n, nt, None, 0, 0, out.data_ptr(), 0, 0,
                                                 ws.data_ptr(), ws.numel(), st)
                else:
                    rc = lib.dsk_minhash_bulk(h, tok.data_ptr(), 0, off.data_ptr(), n, nt, None, 0, 0, out.data_ptr(), 0, 0, st)
                assert rc == 0
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            torch.cuda.synchronize()
            chk = int(out[::997].to(torch.int64).sum().item())
            ref.setdefault(shape, chk)
            res.append("%s %.4f %s" % (shape, e0.elapsed_time(e1) / 10, chk == ref[shape]))
        print(os.path.basename(path), "rep", rep, " ".join(res), flush=True)
