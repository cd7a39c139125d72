#!/usr/bin/env python3
"""Diagnostic: N ranks as threads on ONE device (loopback) -- does the exchange complete, and how fast?
   python profiles/diag/loopback_probe.py [world] [timeout_s]     (env CUDA_DEVICE_MAX_CONNECTIONS is honoured)"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))
import numpy as np  # noqa: E402
from crf_b200 import _cabi, multi, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
timeout = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
bases, offsets, _ = synth.s38(device=None, scale=0.002)
lengths = np.diff(offsets.astype(np.int64))
comms = multi.ThreadComm.split(world)
print(f"world {world} timeout {timeout}s CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')}", flush=True)


def work(rank):
    try:
        rs = multi.RankScan(_cabi.Context(0), comms[rank], bases, offsets[:-1], lengths, 1, 50, 3, 9, chunk=1 << 18,
                            halo=1 << 12, timeout_s=timeout)
        for it in range(3):
            comms[rank].allgather_obj(None)
            t0 = time.perf_counter()
            rs.step_async()
            t1 = time.perf_counter()
            try:
                n = rs.finish()
                print(f"rank {rank} it {it}: enqueue {1e3 * (t1 - t0):.2f} ms, finish {1e3 * (time.perf_counter() - t1):.2f} ms, rows {n} "
                      f"repeated {rs.steps_repeated}", flush=True)
            except Exception as exc:          # noqa: BLE001
                print(f"rank {rank} it {it}: FAILED after {time.perf_counter() - t1:.1f}s: {exc}", flush=True)
        rs.close()
    except threading.BrokenBarrierError:
        pass


threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
for t in threads:
    t.start()
for t in threads:
    t.join()
