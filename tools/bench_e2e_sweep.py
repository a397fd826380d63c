"""End-to-end (host buffers -> dsk_minhash_bulk_host -> host buffers) step time for the bench workload as a
function of the pipeline slice size, next to the raw PCIe copy rates of this box (the e2e roofline)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datasketch_b200 as dsk  # noqa: E402
from datasketch_b200.minhash import _make_permutations  # noqa: E402

n, t, k = 1_000_000, 256, 128
dev = torch.device("cuda", 0)
perms = _make_permutations(k, 1)
h_tok_t = torch.empty((n, t), dtype=torch.int32, pin_memory=True)
h_tok_t.random_(-2 ** 31, 2 ** 31 - 1)
h_off = torch.arange(0, (n + 1) * t, t, dtype=torch.int64).pin_memory().numpy()
h_out_t = torch.empty((n, k), dtype=torch.int32, pin_memory=True)
h_tok = h_tok_t.numpy().view(np.uint32).reshape(-1)
h_out = h_out_t.numpy().view(np.uint32)

# raw copy rates: H2D alone, D2H alone, both directions at once
d_a = torch.empty_like(h_tok_t, device=dev)
d_b = torch.empty_like(h_out_t, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_tok_t, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out_t.copy_(d_b, non_blocking=True)


t_h2d, t_d2h = timed(h2d), timed(d2h)
t_both = timed(lambda: (h2d(), d2h()))
gb_in, gb_out = h_tok_t.numel() * 4 / 1e9, h_out_t.numel() * 4 / 1e9
print(json.dumps({"pcie": {"h2d_GBps": gb_in / t_h2d, "d2h_GBps": gb_out / t_d2h,
                           "duplex_ms_for_step_bytes": t_both * 1e3, "h2d_ms_alone": t_h2d * 1e3,
                           "d2h_ms_alone": t_d2h * 1e3}}), flush=True)

for slice_tok in (1 << 20, 2 << 20, 4 << 20, 8 << 20, 16 << 20, 32 << 20):
    os.environ["DSK_SLICE_TOKENS"] = str(slice_tok)
    dt = timed(lambda: dsk.engine.bulk_signatures(h_tok, h_off, perms, device=0, out=h_out), reps=8)
    print(json.dumps({"slice_tokens": slice_tok, "e2e_ms": dt * 1e3, "sig_per_s": n / dt,
                      "frac_of_duplex_copy_time": t_both / dt}), flush=True)
