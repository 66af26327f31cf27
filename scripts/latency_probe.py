"""Small-call throughput (the reference driver's -b 512 habit, scripts/run-cpu.sh:30, main_banded.cpp:279-291):
  sync      one thread, blocking bsw_extend per call (what every driver thread got before the coalescing queue)
  blocking  T threads, each submit + wait per 512-pair call (what the C++ drop-in class does for the unmodified driver)
  async     T threads, each keeping D calls in flight (a caller written against bsw_extend_async)
python scripts/latency_probe.py > profiles/<tag>_latency.txt"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, genomicsbench_b200 as gb

CALL = 512
cfg = gb.gen_named_config("small")
allp, ref, qer = gb.gen_pairs(cfg, 0, 1 << 17)
want = allp.copy()
with gb.Engine() as e0:
    e0.extend(want, ref, qer, 100)

CPP_ONLY = os.environ.get("LAT_CPP_ONLY", "") not in ("", "0")       # LAT_CPP_ONLY=1: only the C++ probe at the end

print("sync, one thread (pageable buffers):")
sw = gb.BandedPairWiseSW(6, 1, 6, 1, 100, 5, None, 1, 4, 1, devices=[0])
for n in (512, 2048, 4096, 16384, 65536):
    pairs = allp[:n].copy()
    for _ in range(5): sw.getScores16(pairs, ref, qer, n, 1, 100)
    t0 = time.perf_counter(); reps = 30
    for _ in range(reps): sw.getScores16(pairs, ref, qer, n, 1, 100)
    dt = (time.perf_counter() - t0) / reps
    print(f"  n={n:6d} {dt*1e3:7.3f} ms per call  {n/dt/1e6:7.2f} M pairs/s")
sw.close()


def run(T, depth, rounds):
    """T threads, each `rounds` times: submit `depth` calls of CALL pairs, then wait for them."""
    eng = gb.Engine(tiny_batch=1536)
    got = allp.copy()
    ncall = len(got) // CALL
    def worker(tid, nrounds, out):
        k = tid
        for _ in range(nrounds):
            ts = []
            for _ in range(depth):
                v = got[(k % ncall) * CALL:(k % ncall + 1) * CALL]
                ts.append(eng.extend_async(v, ref, qer, 100))
                k += T
            for t in ts: eng.wait(t)
    for nrounds, timed in ((3, False), (rounds, True)):
        ths = [threading.Thread(target=worker, args=(t, nrounds, None)) for t in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in ths]; [t.join() for t in ths]
        dt = time.perf_counter() - t0
    calls, batches = eng.async_stats()
    ok = all(np.array_equal(got[f][:min(len(got), T * depth * CALL)], want[f][:min(len(got), T * depth * CALL)]) for f in gb.RESULT_FIELDS)
    eng.close()
    pairs = T * rounds * depth * CALL
    return pairs / dt / 1e6, calls / max(batches, 1), ok

print("blocking submit + wait per 512-pair call (C++ drop-in route), T threads:")
for T in (() if CPP_ONLY else (1, 2, 4, 8, 16, 32)):
    r, cf, ok = run(T, 1, 60)
    print(f"  T={T:3d}            {r:7.2f} M pairs/s   calls per batch {cf:5.1f}   results ok {ok}")
print("async, T threads x D calls in flight:")
for T, D in (() if CPP_ONLY else ((1, 8), (8, 2), (8, 4), (8, 8), (16, 4))):
    r, cf, ok = run(T, D, 30)
    print(f"  T={T:3d} D={D:2d}       {r:7.2f} M pairs/s   calls per batch {cf:5.1f}   results ok {ok}")

# the same from C++ (no interpreter between the threads and the library): scripts/latency_probe.cpp
import struct, subprocess, tempfile
BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_bin", "latency_probe")
if os.path.exists(BIN):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "pairs.bin")
        with open(path, "wb") as f:
            f.write(struct.pack("<3q", len(allp), len(ref), len(qer)))
            f.write(allp.tobytes()); f.write(ref.tobytes()); f.write(qer.tobytes())
        # LAT_ENVS="A=1 B=2;C=3": one run of the C++ probe per ';'-separated environment (A/B of the shim's switches)
        for envs in os.environ.get("LAT_ENVS", "").split(";"):
            extra = dict(kv.split("=", 1) for kv in envs.split() if "=" in kv)
            print(f"\nC++ (scripts/latency_probe.cpp) {extra if extra else ''}:", flush=True)
            subprocess.run([BIN, path], env=dict(os.environ, **extra))
