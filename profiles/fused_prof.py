"""Where the fused read-1 pass spends its time: runs the bench workload on a library built with -DSSD_FUSED_PROFILE and prints
the clock totals per stage and wait reason (per tile, in microseconds at the SM clock).
    SSD_B200_LIB=profiles/v_fprof.so python profiles/fused_prof.py [pairs]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "split-seq_demultiplexing_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import ssd_b200  # noqa: E402
import fastq_util as fu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
wl = fu.load_whitelist()
dm = ssd_b200.Demultiplexer(wl, errors=1, min_count=10, record_bytes_hint=160)
d1, d2 = ssd_b200.synth_device(dm, ssd_b200.synth_spec(wl, n))
out = torch.empty(d1.numel() + 64 * n, dtype=torch.uint8, device="cuda")
lib = dm.lib
prof = (C.c_ulonglong * 16)()
names = ["A wait list slot", "A wait data", "A work", "B wait list", "B wait desc slot", "B poll chain1", "B work(after waits)", "C wait desc",
         "C poll chain2", "C work(after desc)", "tiles", "copy wait window", "C describe", "C tasks+chunks (thread 0)", "C fence+bulk wait+sync", "C store"]
for rep in range(3):
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
    dm.filter_single()
    lib.ssd_debug_fused_profile(None, 1)
    dm.set_timing(True)
    dm.read_timings()
    w = dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
    t = dm.read_timings()
    lib.ssd_debug_fused_profile(prof, 0)
tiles = max(1, prof[10])
mhz = 1965.0
print(json.dumps({"emit_ms": t["emit"][0], "written": w, "tiles": tiles,
                  "us_per_tile": {nm: round(prof[k] / tiles / mhz, 3) for k, nm in enumerate(names) if k != 10}}))
