#!/usr/bin/env python3
"""bench.py -- Gbp/s scanned (motif 1-50) on the synthetic hg38-sized genome (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload s38|s22] [--impl reference]

One JSON line on stdout (rank 0).  A step = one full scan (all motif sizes 1..50, thresholds,
primitivity, ordering) of the whole workload:
  value   whole-job Gbp/s with the packed planes already resident in HBM (device time, CUDA events,
          max over ranks);
  e2e     the same through the public API with HOST buffers: ASCII bases H2D + pack + scan + results D2H
          every step;
  roofline / cpu_baseline as described in DESIGN.md "Measurement".
Under torchrun (N > 1) every rank takes a contiguous share of (record, chunk) units.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "colab-repeat-finder_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

KMIN, KMAX, MIN_REPEATS, MIN_SPAN = 1, 50, 3, 9       # the reference CLI defaults (prf:86-89)


def metric_name():
    """BASELINE.json's metric; the reads workload (--workload sr, config C5) scans motif 1-20."""
    return f"Gbp/s scanned (motif {KMIN}-{KMAX})"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the JSON result: anything libraries write to fd 1 on the way (NCCL prints its
# version banner there) is sent to stderr, and emit() puts the real stdout back for the one line.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    if _REAL_STDOUT is not None:
        os.dup2(2, 1)


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (oracle port of the reference's tracker loop) ----------------------------------
def cpu_baseline_run(sample_bytes, threads=0):
    from oracle import oracle
    fs = argparse.Namespace(min_motif_size=KMIN, max_motif_size=KMAX, min_repeats=MIN_REPEATS, min_span=MIN_SPAN)
    threads = threads or oracle.max_threads()
    t0 = time.perf_counter()
    start, end, mlen, steps = oracle.detect_repeats_by_k(sample_bytes, fs, threads=threads, arrays=True)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "bp": int(len(sample_bytes)), "rows": (start, end, mlen), "threads": min(threads, KMAX - KMIN + 1)}


def make_workload(name, device, scale):
    from crf_b200 import synth
    if name == "s38":
        return synth.s38(device=device, scale=scale)
    if name == "s22":                                 # configs C1/C2: the real chr22 if this machine has it, else the stand-in
        return synth.chr22(device=device, scale=scale)
    if name == "sr":                                  # config C5: 10 M reads x 150 bp, motif 1-20
        return synth.sr(int(10_000_000 * scale), device=device)
    raise SystemExit(f"unknown workload {name}")


def bench_config(workload_name, total_bp, world):
    """The `config` object of the JSON line: a pure function of the workload and N, so that both arms (ours and
    `--impl reference`) print the same one."""
    chunked = world > 1 and workload_name != "sr"
    return {"workload": workload_name, "motif_sizes": [KMIN, KMAX], "min_repeats": MIN_REPEATS, "min_span": MIN_SPAN,
            "total_bp": int(total_bp),
            "l2": "inputs larger than L2 (no flush): %.0f MB of packed planes over %d GPU(s)" % (total_bp * 0.375 / 1e6, world),
            "partition": ("(record, 2^25-bp chunk) units + 1 Mbp halo" if chunked else
                          "reads split by count" if workload_name.startswith("SR") else "whole records") + f" over {world} rank(s)"}


REF_SLICE_BP = 120_000     # bases per worker process and step for the Python reference (x 50 motif sizes = 6 M tracker steps)


def python_reference_sample(record, sample_name, total_bp, warmup, steps):
    """Time the UNMODIFIED Python reference (oracle/_ref) on `cores` disjoint slices of `record`, one process per host
    core, `steps` timed passes; the C port runs the same bases once for comparison.  Returns the cpu_baseline object."""
    from oracle import oracle, ref
    fs = dict(min_motif_size=KMIN, max_motif_size=KMAX, min_repeats=MIN_REPEATS, min_span=MIN_SPAN)
    n_k = KMAX - KMIN + 1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ref.load()                                # import once; the fork()ed workers inherit it
    slice_bp = max(1000, int(REF_SLICE_BP * 50 / n_k))
    lo0 = min(record.size // 4, 2_000_000)    # past the telomeric N block
    slice_bp = min(slice_bp, max(1, (record.size - lo0) // cores))
    slices = [bytes(record[lo0 + i * slice_bp:lo0 + (i + 1) * slice_bp]).decode("latin-1") for i in range(cores)]
    sample_bp = sum(len(s) for s in slices)
    pool = ref.make_pool(cores)
    times, rows = [], None
    for i in range(warmup + steps):
        wall, _, rows = ref.run_slices(pool, slices, fs)
        if i >= warmup:
            times.append(wall)
    pool.close()
    dt = float(np.mean(times))
    gbps = sample_bp / dt / 1e9
    # the C port on the same bases (one thread per motif size), and its rows against the reference's
    flat = np.frombuffer("".join(slices).encode("latin-1"), dtype=np.uint8)
    port = cpu_baseline_run(flat)
    port_rows = 0
    same = True
    for i, sl in enumerate(slices):
        o = oracle.detect_repeats_by_k(sl, argparse.Namespace(**fs))
        same = same and (o == rows[i])
        port_rows += len(o)
    return {"value": gbps, "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": f"{cores} slices of {slice_bp} bp from {sample_name} (from position {lo0}), {n_k} motif sizes, "
                      f"{dt:.2f} s per pass, {steps} timed pass(es); UNMODIFIED Python reference (oracle/_ref, "
                      f"detect_repeats as shipped, single-threaded), one process per host core; measured, not "
                      f"extrapolated: the whole {total_bp} bp workload at this rate = {total_bp / gbps / 1e9 / 3600:.1f} h "
                      f"on these {cores} cores",
            "seconds_per_pass": dt,
            "per_core_tracker_steps_per_s": sample_bp * n_k / dt / cores,
            "port": {"value": port["bp"] / port["seconds"] / 1e9, "unit": "Gbp/s", "cores": port["threads"],
                     "what": "C port of the tracker loop (oracle/crf_oracle.c) on the same bases, one thread per motif size",
                     "rows_equal_reference": bool(same), "rows": port_rows},
            "python": sys.version.split()[0]}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores, one JSON line.

    With oracle/_ref present (oracle/make_ref.sh; it travels to the GPU box) this is the UNMODIFIED Python
    reference: detect_repeats() (prf:10-81) as shipped, single-threaded, one process per host core over disjoint
    slices of the workload -- the reference's own scale-out shape (1-CPU jobs over intervals,
    hail_batch_pipeline/run_hail_batch_pipeline.py:101).  A step = every worker scans its slice once; value = bases
    scanned per second of wall time, all workers together.  The C port's rate on the same sample is kept beside it.
    Without oracle/_ref the C port alone is timed ("kind": "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle, ref
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = "cuda:0" if torch.cuda.is_available() else None
    if dev:
        bases, offsets, meta = make_workload(args.workload, dev, args.scale)
        total_bp = int(offsets[-1])
        r = min(20, len(offsets) - 2)            # chr21 of S38 (46.7 Mbp); the only record of S22
        record = bases[int(offsets[r]):int(offsets[r + 1])].cpu().numpy()
        sample_name = f"record {r} of the workload"
        del bases
        torch.cuda.empty_cache()
    else:
        from crf_b200 import synth
        record, _, _ = synth.generate_records([int(46_709_983 * args.scale)], 38, device=None)
        meta = {"workload": "S38-like single record (no GPU to generate the full genome)"}
        total_bp = record.size
        sample_name = "standalone 46.7 Mbp record"
    n_k = KMAX - KMIN + 1
    if ref.available():
        cpu = python_reference_sample(record, sample_name, total_bp, args.warmup, args.steps)
        gbps, dt = cpu["value"], cpu["seconds_per_pass"]
    else:
        times = []
        for i in range(args.warmup + args.steps):
            res = cpu_baseline_run(record)
            if i >= args.warmup:
                times.append(res["seconds"])
        dt = float(np.mean(times))
        gbps = record.size / dt / 1e9
        cpu = {"value": gbps, "unit": "Gbp/s", "cores": res["threads"], "kind": "port",
               "sample": f"{sample_name}, {record.size} bp x {n_k} motif sizes per step; oracle/_ref is absent, so this "
                         f"is the C port of the tracker loop (oracle/crf_oracle.c), one thread per motif size; "
                         f"host has {oracle.max_threads()} cores"}
    line = {
        "impl": "reference", "metric": metric_name(), "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(meta["workload"], total_bp, world),
        "cpu_baseline": cpu,
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s38", choices=["s38", "s22", "sr"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--words-per-thread", type=int, default=0)
    ap.add_argument("--trace", action="store_true", help="print per-stage device timings of one extra step (stderr)")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: keep the shares balanced by base pairs")
    ap.add_argument("--phases", type=int, default=1,
                    help="N > 1: phases per step (the rows of one phase travel to rank 0 while the next one is scanned); "
                         "measured slower than one phase on S38 (profiles/r02_scaling_s38.md), kept for larger jobs")
    args = ap.parse_args()
    claim_stdout()
    global KMAX
    if args.workload == "sr":
        KMAX = 20
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from crf_b200 import _cabi, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    if os.environ.get("CRF_BENCH_BLOCKING_SYNC", "0") == "1":     # measured slower at 8 ranks on a 32-core host; off
        # let synchronisation calls sleep instead of spin.  Must precede the creation of the CUDA context.
        import ctypes as _ct
        _rc = _ct.CDLL("libcudart.so.12").cudaInitDevice(local_rank, 4, 1)   # cudaDeviceScheduleBlockingSync, flags valid
        log(f"[rank {rank}] blocking-sync scheduling requested (cudaInitDevice rc={_rc})")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from crf_b200 import multi
    ctx = _cabi.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)          # one stream for torch and the library: the CUDA
    torch.cuda.set_stream(stream)                   # events below see every kernel of the step
    ctx.set_stream(stream.cuda_stream)

    # ---- workload: generated in HBM, deterministic, identical on every rank ----
    t0 = time.time()
    bases, offsets, meta = make_workload(args.workload, dev, args.scale)
    torch.cuda.synchronize()
    lengths = np.diff(offsets.astype(np.int64)).tolist()
    total_bp = int(offsets[-1])
    log(f"[rank {rank}] generated {meta['workload']} in {time.time() - t0:.1f}s")

    # ---- this rank's share: contiguous (record, chunk) units with halo; reads are split by count.  The whole N-rank data
    # ---- path (scan, assembly, rows pushed into rank 0's buffer over NVLink) is crf_b200.multi.RankScan ----
    knobs = {"words_per_thread": args.words_per_thread} if args.words_per_thread else {}
    comm = multi.DistComm(dist, rank, world)
    rs = multi.RankScan(ctx, comm, bases.data_ptr(), offsets[:-1], lengths, KMIN, KMAX, MIN_REPEATS, MIN_SPAN,
                        on_device=True, chunk=args_chunk(world), reads=(args.workload == "sr"), knobs=knobs,
                        phases=args.phases)

    # ---- device-resident timing ----
    rebalanced = []
    for i in range(args.warmup):
        rs.step_async()
        rs.finish()
        if world > 1 and i < min(4, args.warmup - 1) and not args.no_rebalance:
            # shares by measured cost instead of by base pairs (RankScan.rebalance: the slowest rank sets the step); the
            # remaining warm-up steps size the buffers of the new shares
            rebalanced.append(round(rs.rebalance(), 4))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("CRF_BENCH_NO_SAMPLER"):   # one poller only: nvidia-smi queries perturb running kernels
        sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, scan_ms, launches = [], [], 0
    if world > 1:
        dist.barrier()                               # all ranks enter the timed region together
    torch.cuda.synchronize()
    ev0.record(stream)
    if world == 1:
        for _ in range(args.steps):                  # one synchronous crf_scan per step (its counters come back every time)
            rs.step_async()
        total_results = rs.finish()
        st_ = rs.stats()                             # events / counters of the last step (the steps are identical)
        kernel_ms.append(st_["kernel_ms"])
        scan_ms.append(st_["scan_ms"])
        launches = st_["launches"] * args.steps
    else:
        # N ranks: every step is kernel launches only -- counts and rows travel GPU to GPU -- so the K steps are queued back
        # to back and the host waits once; finish() then checks the status word of every one of the K steps
        done = 0
        while done < args.steps:
            burst = min(32, args.steps - done)
            for _ in range(burst):
                rs.step_async()
            done += burst
            if done < args.steps:
                total_results = rs.finish()
                if rs.steps_repeated:
                    raise SystemExit("bench: a timed step had to be repeated (buffers were sized during warm-up?)")
    ev1.record(stream)
    if world > 1:
        total_results = rs.finish()                  # one stream synchronisation; status of every queued step
        if rs.steps_repeated:
            raise SystemExit("bench: a timed step had to be repeated (buffers were sized during warm-up?)")
        st_ = rs.stats()                             # events of the last step, summed over this rank's phases
        kernel_ms.append(st_["kernel_ms"])
        scan_ms.append(st_["scan_ms"])
        launches = st_["launches"] * args.steps
    torch.cuda.synchronize()
    elapsed_ms = ev0.elapsed_time(ev1)
    time.sleep(0.3)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()
    ms_per_step = elapsed_ms / args.steps
    if args.trace:                                   # device time of each stage of one more step, per rank
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        marks[0].record(stream)
        rs.step_async()
        marks[1].record(stream)
        rs.finish()
        st_ = rs.stats()
        xs = torch.cuda.Event(enable_timing=True)
        log(f"[rank {rank}] trace: scan kernels {st_['kernel_ms']:.3f} ms, scan+assembly {st_['scan_ms']:.3f} ms "
            f"({rs.phases} phase(s)), main stream busy {marks[0].elapsed_time(marks[1]):.3f} ms, rows {int(st_['n_results'])}")
        del xs
    value = total_bp / (ms_per_step * 1e-3) / 1e9
    stats = argparse.Namespace(**rs.stats())

    # ---- parity of what the step produced (rank 0 holds the whole job's rows) ----
    parity = None
    if rank == 0:
        rec, st, en, kk = rs.fetch()
        import hashlib
        sha = hashlib.sha256()
        for a_ in (rec, st, en, kk):
            sha.update(np.ascontiguousarray(a_, dtype=np.uint32).tobytes())
        parity = {"rows": int(len(rec)), "rows_sha256": sha.hexdigest()}
        try:                                         # hash of the same workload's rows from a single-GPU run that was
            with open(os.path.join(ROOT, "tests", "golden", "workload_rows_sha256.json")) as f:    # checked against the oracle
                known = json.load(f).get(f"{args.workload}:{args.scale}:{KMIN}-{KMAX}")
            parity["equals_single_gpu_run"] = (known == parity["rows_sha256"]) if known else None
        except OSError:
            parity["equals_single_gpu_run"] = None

    # ---- end to end: host buffers in, results out (on rank 0), every step.  The host buffer is what the native FASTA
    # ---- reader hands over at ingest: 2-bit planes + mask in page-locked memory (crf_pack_ascii; 0.375 B/bp cross PCIe).
    # ---- The same with the raw ASCII text (1 B/bp) is reported beside it as e2e_ascii. ----
    span_lo, span_hi = rs.span()                                         # this rank's contiguous share (+halo)
    span_lo -= span_lo % 32
    n_span = span_hi - span_lo
    host = torch.empty(max(n_span, 1), dtype=torch.uint8, pin_memory=True)
    host[:n_span].copy_(bases[span_lo:span_hi])
    torch.cuda.synchronize()
    host_np = host.numpy()[:n_span]
    nw = (n_span + 31) // 32 + 1
    plane_buf = torch.zeros((3, nw), dtype=torch.int32, pin_memory=True)
    t0 = time.perf_counter()
    planes = _cabi.pack_ascii(host_np, out=tuple(plane_buf[j].numpy().view(np.uint32) for j in range(3))).with_runs()
    pack_s = time.perf_counter() - t0                                    # (the mask goes up as runs of masked positions)
    pinned_out = None

    def e2e_run(source, n_steps):
        nonlocal pinned_out
        times, d2h_ = [], 0
        for i in range(1 + n_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            rs.reload(source, on_device=False, base_offset=span_lo)
            rs.step_async()
            n2 = rs.finish()
            if rank == 0:
                if pinned_out is None or pinned_out.shape[1] < n2:      # result rows land in pinned host memory
                    pinned_out = torch.empty((4, int(n2 * 1.1) + 1024), dtype=torch.int32, pin_memory=True)
                if world == 1:
                    rs.seq.fetch_host(*(pinned_out[j].data_ptr() for j in range(4)), pinned_out.shape[1])
                else:
                    rs.xchg.fetch_to(*(pinned_out[j].data_ptr() for j in range(4)), n2)
                d2h_ = 16 * n2
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()                                          # the step ends when rank 0 holds the rows
            dt_ = time.perf_counter() - t0_
            if i > 0:
                times.append(dt_)
        dt_ = float(np.mean(times))
        if world > 1:
            t_ = torch.tensor([dt_], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            dt_ = float(t_.item())
        return dt_, d2h_

    def total_over_ranks(x):
        if world == 1:
            return int(x)
        t_ = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.SUM)
        return int(t_.item())

    e2e_dt, d2h = e2e_run(planes, args.e2e_steps)
    ascii_dt, _ = e2e_run(host_np, 1)

    # where an end-to-end step goes (one more step, this rank's wall clock with a synchronisation after every stage)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tb0 = time.perf_counter()
    rs.reload(planes, on_device=False, base_offset=span_lo)
    torch.cuda.synchronize()
    tb1 = time.perf_counter()
    rs.step_async()
    n2 = rs.finish()
    torch.cuda.synchronize()
    tb2 = time.perf_counter()
    if rank == 0:
        if world == 1:
            rs.seq.fetch_host(*(pinned_out[j].data_ptr() for j in range(4)), pinned_out.shape[1])
        else:
            rs.xchg.fetch_to(*(pinned_out[j].data_ptr() for j in range(4)), n2)
    torch.cuda.synchronize()
    tb3 = time.perf_counter()
    breakdown = {"upload_and_repack_ms": (tb1 - tb0) * 1e3, "scan_and_gather_ms": (tb2 - tb1) * 1e3,
                 "rows_to_host_ms": (tb3 - tb2) * 1e3,
                 "note": "rank 0, one extra step with a synchronisation after every stage; the upload shrinks with N, the "
                         "rows of the whole job always leave through rank 0's PCIe link"}
    e2e = {"value": total_bp / e2e_dt / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": total_over_ranks(planes.nbytes),
           "d2h_bytes_per_step": total_over_ranks(d2h), "ms_per_step": e2e_dt * 1e3, "breakdown": breakdown,
           "host_buffers": "2-bit planes in page-locked memory + the not-ACGT mask as %d runs, as the native FASTA reader "
                           "packs them at ingest (crf_pack_ascii + crf_mask_runs: %.2f s for this rank's %d bp on the host, "
                           "outside the timed region)" % (planes.runs.shape[0], pack_s, n_span),
           "e2e_ascii": {"value": total_bp / ascii_dt / 1e9, "unit": "Gbp/s", "ms_per_step": ascii_dt * 1e3,
                         "h2d_bytes_per_step": total_over_ranks(n_span),
                         "host_buffers": "the ASCII text itself, page-locked (upper-casing + packing on the GPU)"}}

    if rank != 0:
        rs.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (scan_kernel) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    import ctypes
    tools = ctypes.CDLL(os.path.join(ROOT, "colab-repeat-finder_b200", "crf_b200", "libcrf_tools.so"))
    ops = ctypes.c_double()
    tools.crf_tools_alu_peak(local_rank, ctypes.byref(ops), None)
    alu_peak_tops = ops.value / 1e12
    k_ms = float(np.mean(kernel_ms))
    n_k = KMAX - KMIN + 1
    my_bp = rs.my_bp
    alg_bytes = (my_bp + 3) // 4 + (my_bp + 7) // 8 + 12 * int(stats.n_results)
    alg_ops = 6 * ((my_bp + 31) // 32) * n_k
    hbm_ach = alg_bytes / (k_ms * 1e-3) / 1e9
    int_ach = alg_ops / (k_ms * 1e-3) / 1e12
    t_hbm = alg_bytes / (hbm_peak * 1e9)
    t_int = alg_ops / (alu_peak_tops * 1e12)
    bound_int = t_int >= t_hbm
    traffic = None                 # DRAM bytes of one launch from the committed `ncu --set full` capture (same workload)
    try:
        with open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and args.workload == "s38" and args.scale == 1.0:
            traffic = tj["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "int32-alu" if bound_int else "hbm",
        "achieved": int_ach if bound_int else hbm_ach,
        "peak": alu_peak_tops if bound_int else hbm_peak,
        "unit": "Tops/s" if bound_int else "GB/s",
        "frac": (int_ach / alu_peak_tops) if bound_int else (hbm_ach / hbm_peak),
        "traffic": traffic,
        "kernel": "crf::scan_kernel", "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        "algorithmic_ops_per_launch": alg_ops, "algorithmic_bytes_per_launch": alg_bytes,
        "traffic_unit": "bytes of DRAM read+write per launch; NOT measured in this run: replayed from the committed "
                        "`ncu --set full` capture of the same kernel and workload (profiles/scan_kernel_traffic.json)",
        "peak_source": "INT32 ALU pipe (LOP3+SHF) measured live by crf_tools_alu_peak" if bound_int else hbm_src,
        "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": hbm_src},
        "stated_roofline_ms": max(t_hbm, t_int) * 1e3,
        "note": "stated roofline = slower of one pass of packed bytes at HBM peak and 6 INT32 ops per 32-base word "
                "per motif size at the ALU-pipe peak (SURVEY.md 8d); no tensor-core work on this path",
    }

    # ---- parity on a sample at every N (the gathered rows of one record against the oracle), and at N = 1 the CPU
    # ---- baseline: the Python reference on a bounded sample of the same workload, the C port beside it ----
    from oracle import oracle, ref
    r = 0 if len(lengths) == 1 else 20              # chr21 of S38 (46.7 Mbp x 50 motif sizes)
    if args.workload == "sr":                       # reads: the first 200 000 reads, concatenated with N gaps
        nr = min(200_000, len(lengths))
        gapped = torch.full((nr, 150 + KMAX), ord("N"), dtype=torch.uint8, device=dev)
        gapped[:, :150] = bases[:nr * 150].reshape(nr, 150)
        sample = gapped.flatten().cpu().numpy()
        sample_name = f"the first {nr} reads of the workload"
    else:
        sample = bases[int(offsets[r]):int(offsets[r + 1])].cpu().numpy()
        sample_name = f"record {r} of the workload"
    if not args.no_cpu_baseline:
        res = cpu_baseline_run(sample)
        o_s, o_e, o_m = res["rows"]
        if args.workload == "sr":
            o_rec = o_s // (150 + KMAX)
            sel = rec < nr
            ok = (np.array_equal(rec[sel], o_rec) and np.array_equal(st[sel], o_s - o_rec * (150 + KMAX)) and
                  np.array_equal(en[sel], o_e - o_rec * (150 + KMAX)) and np.array_equal(kk[sel], o_m))
        else:
            sel = rec == r
            ok = np.array_equal(st[sel], o_s) and np.array_equal(en[sel], o_e) and np.array_equal(kk[sel], o_m)
        parity["parity_on_sample"] = bool(ok)
        parity["sample"] = f"{sample_name}: {int(sel.sum())} gathered rows vs the oracle, row by row"
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        port = {"value": res["bp"] / res["seconds"] / 1e9, "unit": "Gbp/s", "cores": res["threads"],
                "what": f"C port of the reference's tracker loop (oracle/crf_oracle.c) on {sample_name}, {res['bp']} bp x "
                        f"{n_k} motif sizes, {res['seconds']:.2f} s, one thread per motif size"}
        if ref.available() and args.workload != "sr":
            cpu = python_reference_sample(sample, sample_name, total_bp, 0, 1)
            cpu["port_on_whole_record"] = port
        else:
            cpu = dict(port, kind="port", sample=port.pop("what"))
        cpu["parity_on_sample"] = parity.get("parity_on_sample")

    line = {
        "metric": metric_name(), "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": bench_config(meta["workload"], total_bp, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "results_per_step": int(total_results), "units": rs.n_units, "rebalanced_gain": rebalanced,
        "scan_stats": {"scan_ms": float(np.mean(scan_ms)), "kernel_ms": k_ms, "candidates": int(stats.n_candidates),
                       "long_runs": int(stats.n_long), "spilled": int(stats.n_spilled), "tiles": int(stats.n_tiles)},
    }
    emit(line)
    rs.close()
    if world > 1:
        dist.destroy_process_group()


def args_chunk(world):
    from crf_b200 import partition
    return partition.DEFAULT_CHUNK if world > 1 else (1 << 62)   # one GPU: whole records, nothing to stitch


if __name__ == "__main__":
    main()
