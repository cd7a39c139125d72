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
    if name == "s22":
        return synth.s22(device=device, scale=scale)
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
    fs = dict(min_motif_size=KMIN, max_motif_size=KMAX, min_repeats=MIN_REPEATS, min_span=MIN_SPAN)
    n_k = KMAX - KMIN + 1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if ref.available():
        ref.load()                                # import once; the fork()ed workers inherit it
        slice_bp = max(1000, int(REF_SLICE_BP * 50 / n_k))
        lo0 = min(record.size // 4, 2_000_000)    # past the telomeric N block
        slice_bp = min(slice_bp, max(1, (record.size - lo0) // cores))
        slices = [bytes(record[lo0 + i * slice_bp:lo0 + (i + 1) * slice_bp]).decode("latin-1") for i in range(cores)]
        sample_bp = sum(len(s) for s in slices)
        pool = ref.make_pool(cores)
        times, rows = [], None
        for i in range(args.warmup + args.steps):
            wall, _, rows = ref.run_slices(pool, slices, fs)
            if i >= args.warmup:
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
        cpu = {"value": gbps, "unit": "Gbp/s", "cores": cores, "kind": "reference",
               "sample": f"{cores} slices of {slice_bp} bp from {sample_name} (from position {lo0}), {n_k} motif sizes, "
                         f"{dt:.2f} s per step; UNMODIFIED Python reference (oracle/_ref, detect_repeats as shipped, "
                         f"single-threaded), one process per host core; measured, not extrapolated: the whole "
                         f"{total_bp} bp workload at this rate = {total_bp / gbps / 1e9 / 3600:.1f} h on these {cores} cores",
               "per_core_tracker_steps_per_s": sample_bp * n_k / dt / cores,
               "port": {"value": port["bp"] / port["seconds"] / 1e9, "unit": "Gbp/s", "cores": port["threads"],
                        "what": "C port of the tracker loop (oracle/crf_oracle.c) on the same bases, one thread per motif size",
                        "rows_equal_reference": bool(same), "rows": port_rows},
               "python": sys.version.split()[0]}
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
    ap.add_argument("--trace", action="store_true", help="print per-stage host timings of one extra step (stderr)")
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

    ctx = _cabi.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)          # one stream for torch, NCCL waits and the library: the CUDA
    torch.cuda.set_stream(stream)                   # events below see every kernel of the step
    ctx.set_stream(stream.cuda_stream)

    # ---- workload: generated in HBM, deterministic, identical on every rank ----
    t0 = time.time()
    bases, offsets, meta = make_workload(args.workload, dev, args.scale)
    torch.cuda.synchronize()
    lengths = np.diff(offsets.astype(np.int64)).tolist()
    total_bp = int(offsets[-1])
    log(f"[rank {rank}] generated {meta['workload']} in {time.time() - t0:.1f}s")

    # ---- this rank's share: contiguous (record, chunk) units with halo; reads are split by count ----
    if args.workload == "sr":
        n_rec = len(lengths)
        lo_r, hi_r = n_rec * rank // world, n_rec * (rank + 1) // world
        plan, mine = None, []
        starts = offsets[lo_r:hi_r].astype(np.uint64)
        lens = np.diff(offsets[lo_r:hi_r + 1]).astype(np.uint64)
        own_lo = own_hi = None
        n_units, my_bp = n_rec, int(lens.sum())
    else:
        plan = partition.Plan(lengths, world, chunk=args_chunk(world), halo=partition.DEFAULT_HALO,
                              kmax=KMAX, min_repeats=MIN_REPEATS, min_span=MIN_SPAN)
        mine = plan.units_of(rank)
        starts, lens, own_lo, own_hi = plan.load_args(rank, offsets)
        if world == 1:
            own_lo = own_hi = None                      # whole records: nothing to own or stitch
        n_units, my_bp = len(plan.units), int(sum(u.d1 - u.d0 for u in mine))
    knobs = {"words_per_thread": args.words_per_thread} if args.words_per_thread else {}

    seq = ctx.load_ranges(bases.data_ptr(), starts, lens, own_lo, own_hi, max_motif_cap=KMAX, on_device=True)
    info = seq.info()

    def one_step():
        return seq.scan(KMIN, KMAX, MIN_REPEATS, MIN_SPAN, **knobs)

    if world > 1 and plan is not None:   # results come out in chromosome coordinates; open-ended results are counted by the library
        seq.set_output_map(out_record=[u.record for u in mine], out_shift=[u.d0 for u in mine],
                           open_ended=[int(u.d1 < u.rec_len) for u in mine])
    gstate = {"cap": 0, "buf": None, "out": None}
    MAX_OPEN = 32

    trace = {"on": False, "t": []}

    def mark(name):
        if trace["on"]:
            torch.cuda.synchronize()
            trace["t"].append((name, time.perf_counter()))

    def gather_to_rank0(n):
        """N > 1: compacted results -> rank 0 (the only collective of the path: one 16-byte all-gather of
        (count, open count), one gather of the rows); runs that left their unit's data (longer than the halo)
        are stitched first.  Returns the whole-job result count."""
        if world == 1:
            return n
        n_open = int(seq.stats().n_open)
        open_rows = seq.fetch_open() if (n_open and plan is not None) else np.zeros((0, 5), np.uint32)

        def fetch_rows():                             # this rank's rows -> its slot of the gather buffer (device to device)
            cap_, buf_ = gstate["cap"], gstate["buf"]
            cols = [buf_[i * cap_:i * cap_ + n] for i in range(4)]
            seq.fetch_device(*(c.data_ptr() for c in cols), n)
            return cols
        cols = fetch_rows() if gstate["cap"] >= n else None   # queued before the header exchange blocks the host
        # one small all-gather carries every rank's row count and its open-ended rows (record, start, end, k)
        hdr = torch.full((2 + 4 * MAX_OPEN,), -1, dtype=torch.int64)
        hdr[0], hdr[1] = n, len(open_rows)
        if 0 < len(open_rows) <= MAX_OPEN:
            hdr[2:2 + 4 * len(open_rows)] = torch.from_numpy(open_rows[:, 1:].astype(np.int64).reshape(-1))
        hdrs = torch.empty(world * hdr.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(hdrs, hdr.to(dev))
        hdrs = hdrs.cpu().view(world, -1)
        mark('hdr all-gather')
        counts = hdrs[:, 0].tolist()
        n_open_all = hdrs[:, 1].tolist()
        if max(counts) > gstate["cap"]:
            cap = int(max(counts) * 1.05) + 1024
            gstate["cap"] = cap
            gstate["buf"] = torch.zeros(4 * cap, dtype=torch.int32, device=dev)
            gstate["out"] = [torch.empty_like(gstate["buf"]) for _ in range(world)] if rank == 0 else None
            cols = None
        if cols is None:
            cols = fetch_rows()
        buf, en = gstate["buf"], cols[2]
        mark('fetch_device')
        if sum(n_open_all) and plan is not None:      # a repeat longer than the halo crossed a unit end
            run_end = lambda unit, lp, k: seq.run_end(unit.index - plan.bounds[rank], lp, k)   # noqa: E731
            if max(n_open_all) <= MAX_OPEN:
                open_all = [tuple(hdrs[r, 2 + 4 * i:6 + 4 * i].tolist()) for r in range(world) for i in range(n_open_all[r])]

                def exchange(ans):                    # one all-reduce per hop
                    t = torch.full((len(open_all),), -1, dtype=torch.int64)
                    for i, v in ans.items():
                        t[i] = v
                    t = t.to(dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    return {i: v for i, v in enumerate(t.cpu().tolist()) if v >= 0}
                fixed = partition.stitch(plan, open_all, run_end, exchange, rank)
            else:
                fixed = partition.stitch_collective(plan, [tuple(int(x) for x in row[1:]) for row in open_rows],
                                                    run_end, rank, world, dist, dev)
            mine_fixed = {(r_, s_, k_): e_ for (r_, s_, e_, k_) in fixed}
            for row in open_rows:
                new_end = mine_fixed[(int(row[1]), int(row[2]), int(row[4]))]
                seq.patch_end(int(row[0]), new_end)
                en[int(row[0])] = new_end
        mark('stitch')
        dist.gather(buf, gstate["out"], dst=0)       # rank order == genome order: concatenation is the sorted result
        torch.cuda.synchronize()                     # keep the NCCL copy kernels out of the next step's scan kernel:
        mark('gather')                               # overlapped, they wait for free SMs and the steps queue up
        return int(sum(counts))

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        n_res = one_step()
        gather_to_rank0(n_res)
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
    step_t = [time.perf_counter()]
    for _ in range(args.steps):
        n_res = one_step()
        st_ = seq.stats()
        kernel_ms.append(st_.kernel_ms)
        scan_ms.append(st_.scan_ms)
        launches += st_.launches
        total_results = gather_to_rank0(n_res)
        step_t.append(time.perf_counter())
    if world > 1:
        dist.barrier()
    ev1.record(stream)
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
    if args.trace:
        log(f"[rank {rank}] host time per timed step (ms): " + " ".join(f"{(b - a) * 1e3:.2f}" for a, b in zip(step_t, step_t[1:])))
        trace["on"] = True
        mark("start")
        n_t = one_step()
        mark("scan")
        gather_to_rank0(n_t)
        trace["on"] = False
        t0_ = trace["t"][0][1]
        log(f"[rank {rank}] trace: " + ", ".join(f"{nm} +{(t - t0_) * 1e3:.3f} ms" for nm, t in trace["t"][1:]))
    value = total_bp / (ms_per_step * 1e-3) / 1e9
    stats = seq.stats()

    # ---- end to end: host buffers in, results out, every step ----
    e2e = None
    span_lo, span_hi = int(starts.min()), int((starts + lens).max())   # this rank's contiguous share (+halo)
    host = torch.empty(span_hi - span_lo, dtype=torch.uint8, pin_memory=True)
    host.copy_(bases[span_lo:span_hi])
    h_starts = starts - np.uint64(span_lo)
    torch.cuda.synchronize()
    host_np = host.numpy()
    e2e_times = []
    pinned_out = None
    d2h = 0
    for i in range(1 + args.e2e_steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with ctx.load_ranges(host_np, h_starts, lens, own_lo, own_hi, max_motif_cap=KMAX, on_device=False) as s2:
            n2 = s2.scan(KMIN, KMAX, MIN_REPEATS, MIN_SPAN, **knobs)
            if pinned_out is None or pinned_out.shape[1] < n2:          # result rows land in pinned host memory
                pinned_out = torch.empty((4, int(n2 * 1.1) + 1024), dtype=torch.int32, pin_memory=True)
            s2.fetch_host(*(pinned_out[j].data_ptr() for j in range(4)), pinned_out.shape[1])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h = 16 * n2
        if i > 0:
            e2e_times.append(dt)
    e2e_dt = float(np.mean(e2e_times))
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        tb = torch.tensor([float(host.numel()), float(d2h)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        h2d_total, d2h_total = int(tb[0].item()), int(tb[1].item())
    else:
        h2d_total, d2h_total = host.numel(), d2h
    e2e = {"value": total_bp / e2e_dt / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": h2d_total,
           "d2h_bytes_per_step": d2h_total, "ms_per_step": e2e_dt * 1e3}

    if rank != 0:
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
        "traffic_unit": "bytes of DRAM read+write per launch (ncu --set full, profiles/r01_scan_kernel_ncu.txt)",
        "peak_source": "INT32 ALU pipe (LOP3+SHF) measured live by crf_tools_alu_peak" if bound_int else hbm_src,
        "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": hbm_src},
        "stated_roofline_ms": max(t_hbm, t_int) * 1e3,
        "note": "stated roofline = slower of one pass of packed bytes at HBM peak and 6 INT32 ops per 32-base word "
                "per motif size at the ALU-pipe peak (SURVEY.md 8d); no tensor-core work on this path",
    }

    # ---- CPU baseline on a bounded sample of the same workload + parity on that sample ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = 0 if len(lengths) == 1 else 20          # chr21 of S38 (46.7 Mbp x 50 motif sizes)
        if args.workload == "sr":                   # reads: the first 200 000 reads, concatenated with N gaps
            nr = min(200_000, len(lengths))
            block = bases[:nr * 150].reshape(nr, 150)
            gapped = torch.full((nr, 150 + KMAX), ord("N"), dtype=torch.uint8, device=dev)
            gapped[:, :150] = block
            sample = gapped.flatten().cpu().numpy()
        else:
            sample = bases[int(offsets[r]):int(offsets[r + 1])].cpu().numpy()
        res = cpu_baseline_run(sample)
        rec, st, en, kk = seq.fetch(int(stats.n_results))
        unit_of_r = [i for i, u in enumerate(mine) if u.record == r]
        parity = None
        if len(unit_of_r) == 1 and args.workload != "sr":  # record scanned as one unit: compare row by row
            sel = rec == unit_of_r[0]
            o_s, o_e, o_m = res["rows"]
            parity = bool(np.array_equal(st[sel], o_s) and np.array_equal(en[sel], o_e) and
                          np.array_equal(kk[sel], o_m))
        from oracle import oracle
        cpu = {"value": res["bp"] / res["seconds"] / 1e9, "unit": "Gbp/s", "cores": res["threads"], "kind": "port",
               "sample": f"record {r} of the workload, {res['bp']} bp x {n_k} motif sizes, {res['seconds']:.2f} s; "
                         f"C port of the reference's tracker loop (oracle/crf_oracle.c), one thread per motif size, "
                         f"host has {oracle.max_threads()} cores",
               "parity_on_sample": parity}

    line = {
        "metric": metric_name(), "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": meta["workload"], "motif_sizes": [KMIN, KMAX], "min_repeats": MIN_REPEATS,
                   "min_span": MIN_SPAN, "total_bp": total_bp, "results_per_step": int(total_results),
                   "l2": "inputs larger than L2 (packed planes %.0f MB per GPU)" % (info.packed_bytes / 1e6),
                   "partition": f"{n_units} (record, chunk) units over {world} rank(s)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "scan_stats": {"scan_ms": float(np.mean(scan_ms)), "kernel_ms": k_ms, "candidates": int(stats.n_candidates),
                       "long_runs": int(stats.n_long), "spilled": int(stats.n_spilled), "tiles": int(stats.n_tiles)},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def args_chunk(world):
    from crf_b200 import partition
    return partition.DEFAULT_CHUNK if world > 1 else (1 << 62)   # one GPU: whole records, nothing to stitch


if __name__ == "__main__":
    main()
