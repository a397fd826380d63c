"""bench.py -- MinHash signatures/sec on the BASELINE.json workload (configs[1]:
1M docs x 256 tokens, num_perm=128 bulk signature build), N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's numpy CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic token hashes.
  value  : whole-job signatures/s with inputs already resident in HBM (CUDA events on the
           launch stream, max over ranks).  Inputs (1 GB/GPU) are larger than L2.
  e2e    : same metric through the host-buffer C-ABI call (dsk_minhash_bulk_host) with pinned
           HOST buffers: H2D of tokens+offsets and D2H of the signatures inside the timed region.
  roofline: algorithmic bytes (4T+4K per document) / kernel time vs the
           measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline: the oracle port of the reference's numpy path on all host cores, bounded sample.
Scaling is weak (each rank builds its own 1M-document shard; documents shard embarrassingly,
no data-path collective); the optional NCCL all-gather of the signature matrix is timed
separately and reported under "allgather".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "minhash_signatures_per_sec"
UNIT = "signatures/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--num-perm", type=int, default=128)
    ap.add_argument("--kernel", default="auto", choices=["auto", "two_phase", "direct", "exact"])
    ap.add_argument("--cpu-sample-docs", type=int, default=0, help="0 = auto (about 10-30 s of CPU work)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
def make_tokens(n_docs: int, t: int, seed: int, out: np.ndarray) -> None:
    """Uniform 32-bit token hashes (i.i.d.), filled chunk-wise into `out`."""
    rng = np.random.default_rng(seed)
    flat = out.reshape(-1)
    step = 1 << 24
    for i in range(0, flat.size, step):
        j = min(i + step, flat.size)
        flat[i:j] = rng.integers(0, 1 << 32, size=j - i, dtype=np.uint32)


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of datasketch's numpy path, all host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    tok, t, k, seed = args
    from oracle import oracle_np as o
    off = np.arange(tok.shape[0] + 1, dtype=np.int64) * t
    sig = o.bulk_signatures_csr(tok.reshape(-1), off, k, seed)
    return int(sig[:, 0].sum() & 0xFFFF)


def cpu_reference_rate(n_docs: int, t: int, k: int, cores: int, repeats: int = 1):
    """signatures/s of MinHash.bulk's numpy arithmetic (oracle/oracle_np.py) over `cores` processes."""
    import multiprocessing as mp
    tok = np.empty((n_docs, t), dtype=np.uint32)
    make_tokens(n_docs, t, 99, tok)
    shards = [s for s in np.array_split(tok, cores) if len(s)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(shards)) as pool:
        pool.map(_cpu_worker, [(s[:8], t, k, 1) for s in shards])  # warm the workers
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(s, t, k, 1) for s in shards])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_docs / best, best


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0: float, t1: float):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = max(mx, cmax)
            if t0 <= ts <= t1 + 0.15:
                sm.append(clk)
                for nm, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (what `numactl` would do), so the
    pinned host buffers of the e2e leg are first-touched on the local node.  Best effort; returns a note."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "numa: unknown"
        if len(bus.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return "numa: node %d (%d cpus)" % (node, len(cpus))
    except Exception as exc:  # noqa: BLE001
        return "numa: not bound (%s)" % type(exc).__name__


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its numpy uint64
    arithmetic, restated in oracle/oracle_np.py because /root/reference is absent on the GPU box),
    on all host cores; each step = a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample_docs or min(args.docs, 2500 * cores)
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_reference_rate(sample, args.tokens, args.num_perm, cores)
        if i >= args.warmup:
            rates.append((r, dt))
    val = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: %d docs x %d tokens, num_perm=%d bulk signature build"
                               % (args.docs, args.tokens, args.num_perm),
                   "sample_docs_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d docs x %d tokens per step, numpy uint64 path over %d processes"
                                   % (sample, args.tokens, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import datasketch_b200 as dsk
    from datasketch_b200 import _native as nv
    from datasketch_b200.minhash import _make_permutations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n, t, k = args.docs, args.tokens, args.num_perm

    # ---- CPU baseline on the same box (rank 0, N=1 only), before CUDA is initialised (fork) --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample_docs or min(n, 2500 * cores)
        rate, dt = cpu_reference_rate(sample, t, k, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d docs x %d tokens once (%.1f s wall), numpy uint64 path of datasketch over %d processes"
                         % (sample, t, dt, cores)}

    numa_note = bind_to_gpu_numa_node(local)   # after the CPU baseline (which uses every core)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG >= VERSION
        if not os.environ.get("DSK_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    perms = _make_permutations(k, 1)

    # ---- synthetic inputs: pinned host copy (for e2e) and resident device copy (for value) ----
    h_tok_t = torch.empty((n, t), dtype=torch.int32, pin_memory=True)
    h_tok = h_tok_t.numpy().view(np.uint32)
    make_tokens(n, t, 1000 + rank, h_tok)
    h_off_t = torch.arange(0, (n + 1) * t, t, dtype=torch.int64).pin_memory()
    h_out_t = torch.empty((n, k), dtype=torch.int32, pin_memory=True)
    d_tok = h_tok_t.to(dev, non_blocking=True)
    d_off = h_off_t.to(dev, non_blocking=True)
    d_out = torch.empty((n, k), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream(dev)

    def step_device():
        dsk.engine.bulk_signatures_device(d_tok, d_off, n * t, perms, d_out=d_out, kernel=args.kernel,
                                          stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- value: device-resident, CUDA events ----------------------------------------------------
    BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_value():
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()
        tw0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
        barrier()
        tw1 = time.perf_counter()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, tw0, tw1

    ms_total, t_wall0, t_wall1 = timed_value()
    remeasured = False
    if rank == 0 and world == 1:
        time.sleep(0.05)
        first = sampler.summary(t_wall0, t_wall1)
        if BAD & set(first["reasons"]):      # thermal / hw slowdown seen: re-measure once (timing rules)
            ms_total, t_wall0, t_wall1 = timed_value()
            remeasured = True
    ms_step = ms_total / args.steps
    value = n * world / (ms_step * 1e-3)

    # parity spot-check on the benchmark's own output (oracle = checker only)
    if rank == 0:
        from oracle import oracle_clib as oc
        idx = np.arange(0, n, max(1, n // 64))[:64]
        got = d_out[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32)
        sub = np.ascontiguousarray(h_tok[idx]).reshape(-1)
        want = oc.minhash_bulk_u32tok(sub, np.arange(len(idx) + 1, dtype=np.int64) * t, perms)
        if not np.array_equal(got, want):
            raise SystemExit("bench: GPU signatures differ from the oracle -- number is invalid")

    # ---- e2e: pinned host buffers through the host C-ABI ------------------------------------------
    e2e = None
    e2e_launches = 0
    if not args.no_e2e:
        h_out = h_out_t.numpy().view(np.uint32)
        h_off = h_off_t.numpy()

        def step_host():
            dsk.engine.bulk_signatures(h_tok.reshape(-1), h_off, perms, kernel=args.kernel, device=local, out=h_out)

        for _ in range(max(2, min(args.warmup, 3))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": n * world * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h_tok.nbytes + h_off.nbytes) * world,
               "d2h_bytes_per_step": int(h_out.nbytes) * world, "ms_per_step": dt / args.steps * 1e3}
        if rank == 0 and not np.array_equal(h_out[idx], want):
            raise SystemExit("bench: host-path signatures differ from the oracle -- number is invalid")
        # the e2e leg's own roofline: the same bytes as bare pinned copies, both directions at once (PCIe duplex)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        d_scr = torch.empty_like(d_tok)

        def copies():
            with torch.cuda.stream(s_in):
                d_scr.copy_(h_tok_t, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out_t.copy_(d_out, non_blocking=True)

        keep = h_out[idx].copy()
        copies()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copies()
        torch.cuda.synchronize()
        floor_ms = (time.perf_counter() - t0) / 3 * 1e3
        if world > 1:
            tt = torch.tensor([floor_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            floor_ms = float(tt.item())
        e2e["copy_floor_ms"] = floor_ms
        e2e["frac_of_copy_floor"] = floor_ms / e2e["ms_per_step"]
        del d_scr
        assert np.array_equal(h_out[idx], keep)  # d_out holds the same signatures the host path produced
        slices = max(-(-n // (128 << 10)), -(-(n * t) // (16 << 20)))
        e2e_launches = slices * args.steps

    # ---- optional: NCCL all-gather of the signature matrix (assembly for LSH bucketing) -----------
    allgather = None
    if world > 1:
        full = torch.empty((world * n, k), dtype=torch.int32, device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(full, d_out)
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
            dist.all_gather_into_tensor(full, d_out)
        ev1.record(stream)
        barrier()
        tt = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_g = float(tt.item()) / args.steps
        allgather = {"value": n * world / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g,
                     "gathered_bytes_per_gpu": int(full.numel() * 4), "how": "kernel, then NCCL all_gather_into_tensor"}
        # fused variant: the kernel stores every row into all ranks' matrices over NVLink (no NCCL pass)
        try:
            from datasketch_b200.distributed import FusedGather
            del full
            fg = FusedGather(world * n, k, device=local)
            for _ in range(2):
                fg.build(d_tok, d_off, n * t, perms, row_offset=rank * n, kernel=args.kernel)
            barrier()
            ev0.record(stream)
            for _ in range(args.steps):
                fg.build(d_tok, d_off, n * t, perms, row_offset=rank * n, kernel=args.kernel, sync=False)
            ev1.record(stream)
            fg.finish()
            barrier()
            tt = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_f = float(tt.item()) / args.steps
            ok = bool(torch.equal(fg.buf[rank * n: rank * n + 4096], d_out[:4096]))
            other = (rank + 1) % world
            chk = torch.empty((4096, k), dtype=torch.int32, device=dev)
            src = d_out[:4096].contiguous()
            gl = [torch.empty_like(src) for _ in range(world)]
            dist.all_gather(gl, src)
            ok = ok and bool(torch.equal(fg.buf[other * n: other * n + 4096], gl[other]))
            allgather["fused"] = {"value": n * world / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f,
                                  "how": "dsk_minhash_bulk_gather: peer stores over NVLink in the kernel epilogue",
                                  "rows_verified": ok}
        except Exception as exc:  # noqa: BLE001  (symmetric memory unavailable on this box)
            allgather["fused"] = {"unavailable": repr(exc)[:200]}

    if rank == 0:
        sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
    if clocks is not None:
        clocks["remeasured"] = remeasured
        if clocks["samples"] < 3:   # the timed region is only tens of ms: add the samples of the whole run under load
            wide = sampler.summary(t_wall0 - 0.2, time.perf_counter())
            clocks["sm_mhz_whole_run"] = wide["sm_mhz"]
            clocks["reasons"] = sorted(set(clocks["reasons"]) | set(wide["reasons"]))
            clocks["samples_whole_run"] = wide["samples"]

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg_bytes = (4 * t + 4 * k) * n  # per launch, per GPU
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this kernel/workload
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        h = nv.perm_handle(perms, local)
        kern = args.kernel if args.kernel != "auto" else ("two_phase" if h.n_unsafe == 0 else "exact")
        if os.path.exists(tpath):
            ent = json.load(open(tpath)).get("minhash_bulk_kernel<%s>|%dx%dxK%d" % (kern, n, t, k))
            if ent:
                traffic = ent["traffic_bytes_per_launch"] / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (mod 2^64 wrap, mod 2^61-1) on u32 lanes", "data": "synthetic",
            "config": {"workload": "configs[1]: %d docs x %d tokens, num_perm=%d bulk signature build per GPU"
                                   % (n, t, k),
                       "kernel": "minhash_bulk_kernel<%s>" % kern, "l2": "inputs (%.2f GB/GPU) larger than L2"
                                   % (h_tok.nbytes / 1e9), "parallelism": "documents sharded x%d, no collective" % world,
                       "host": numa_note},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write; algorithmic %.3f GB)"
                                                             % (alg_bytes / 1e9), "peak_source": peak_src,
                         "note": "binding roof is the integer pipe (T*K evaluations/doc), see DESIGN.md"},
            "e2e": e2e, "gpu_launches": args.steps + e2e_launches, "clocks": clocks,
        }
        # the roof that actually binds this kernel: one 32-bit IMAD per (token, permutation) evaluation is the floor
        # of any exact scheme, and B200 issues IMAD at 16 lanes/clk/SMSP = 64 lanes/clk/SM (profiles/)
        sm_count = nv.device_info(local)["sm_count"]
        clk_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        evals_per_s = float(n) * t * k / (ms_step * 1e-3)
        imad_peak = sm_count * 64 * clk_mhz * 1e6
        line["int_pipe"] = {"bound": "imad (fmaheavy pipe)", "achieved": evals_per_s, "peak": imad_peak,
                            "unit": "evaluations/s per GPU", "frac": evals_per_s / imad_peak,
                            "note": "peak = SMs x 64 IMAD lanes/clk x SM clock under load; 1 IMAD per evaluation minimum"}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if allgather is not None:
            line["allgather"] = allgather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
