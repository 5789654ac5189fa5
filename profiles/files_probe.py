"""Where a file-to-file call spends its time (ssd_run_files on RAM-backed files): run with SSD_IO_TRACE=1 on the GPU box."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "split-seq_demultiplexing_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, fastq_util as fu, ssd_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
wl = fu.load_whitelist()
dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, record_bytes_hint=160)
d1, d2 = ssd_b200.synth_device(dm, ssd_b200.synth_spec(wl, n, seed=0x5EED0001))
tmp = tempfile.mkdtemp(prefix="ssd_probe_", dir="/dev/shm")
f1, f2, fo = [os.path.join(tmp, x) for x in ("r1.fastq", "r2.fastq", "out.fastq")]
d1.cpu().numpy().tofile(f1); d2.cpu().numpy().tofile(f2)
del d1, d2
try:
    for rep in range(4):
        t0 = time.perf_counter()
        wrote, st = dm.run_files(f1, f2, fo, ssd_b200.EMIT_PASSING, append=False)
        dt = time.perf_counter() - t0
        print("rep %d: %.3f s, %.1f M pairs/s, wrote %d" % (rep, dt, n / dt / 1e6, wrote), flush=True)
        sys.stderr.flush()
finally:
    shutil.rmtree(tmp, ignore_errors=True)
