#!/usr/bin/env python
"""bench.py — read pairs/sec demultiplexed (e=1) on B200, device-resident and end-to-end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one pass of the whole hot path over one batch of synthetic read pairs: FASTQ scan of both
mates, barcode/UMI extraction, Hamming<=1 matching, per-cell histogram, distinct-UMI -m 10 filter and
the annotated-FASTQ writer (BASELINE.json configs[1]: synthetic 10 M read pairs, 150 bp, e=1, merged
output).  With N > 1 (torchrun, one rank per GPU) every rank holds its own 10 M pairs (weak scaling), the
per-cell histograms are all-reduced over NCCL and the (cell,UMI) keys are exchanged by cell range so
the -m decision is global and exact.

`value` times the steps with inputs resident in HBM (CUDA events, max over ranks); `e2e` times the same
pass through ssd_run_host/ssd_emit_host with pinned HOST buffers (H2D and D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm (the C oracle port; the reference itself is
Python and does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "split-seq_demultiplexing_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "read pairs/sec demultiplexed (e=1)"
UNIT = "read pairs/s"
SEED = 0x5EED0001
PAIRS_PER_GPU = int(os.environ.get("SSD_BENCH_PAIRS", 10_000_000))
READ_LEN = 150
ERRORS, MIN_COUNT = 1, 10


def load_whitelist():
    with open(os.path.join(ROOT, "tests", "golden", "barcodes_8mer.txt")) as fh:
        return [l.strip() for l in fh if l.strip()]


def workload_config(n_gpus):
    return {"workload": "synthetic %d read pairs per GPU (150 bp both mates, SURVEY 8d shape), e=1, -m 10 (distinct UMIs), "
                        "merged output" % PAIRS_PER_GPU,
            "pairs_per_gpu": PAIRS_PER_GPU, "read_len": READ_LEN, "errors": ERRORS, "min_count": MIN_COUNT,
            "filter": "distinct_umi", "seed": hex(SEED), "sharding": "record ranges, %d rank(s)" % n_gpus,
            "l2": "inputs (6.5 GB per GPU) are far larger than the 126 MB L2; no explicit flush"}


def ncu_traffic(kernel_key, pairs_per_gpu):
    """DRAM bytes (read + written) per launch of the kernel from the committed `ncu --set full` capture of this very
    workload (profiles/r01_dram_traffic.json, made by profiles/extract_traffic.py); None for any other workload."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_dram_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        if t["workload"]["pairs_per_gpu"] != pairs_per_gpu or t["workload"]["read_len"] != READ_LEN:
            return None
        return t["kernels"][kernel_key]["traffic"]
    except (OSError, KeyError, ValueError):
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="ssd_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            # the GPU idles between our launches only outside the timed loops; keep the busy half
            busy = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def oracle_rate(n_pairs, workers, reps=1):
    """pairs/s of the CPU port (CLI3 structure: split -> per-chunk demultiplex on `workers` threads -> filter)."""
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = load_whitelist()
    spec = ssd_b200.synth_spec(wl, n_pairs, seed=SEED, read_len=READ_LEN)
    r1, r2 = ssd_b200.synth_host(spec)
    n_lines = 4 * n_pairs
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = orc.run_cli(r1, r2, wl, ERRORS, MIN_COUNT, n_workers=workers, length_fastq=n_lines)
        times.append(time.perf_counter() - t0)
    return n_pairs / min(times), times, res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # size the per-step sample so that the whole run stays within a couple of minutes
    rate, _, _ = oracle_rate(50_000, cores)
    budget_s = 100.0
    n = int(max(20_000, min(2_000_000, budget_s * rate / max(1, args.steps + args.warmup))))
    import ssd_b200
    from oracle import ssd_oracle_py as orc
    wl = load_whitelist()
    spec = ssd_b200.synth_spec(wl, n, seed=SEED, read_len=READ_LEN)
    r1, r2 = ssd_b200.synth_host(spec)
    for _ in range(args.warmup):
        orc.run_cli(r1, r2, wl, ERRORS, MIN_COUNT, n_workers=cores, length_fastq=4 * n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.run_cli(r1, r2, wl, ERRORS, MIN_COUNT, n_workers=cores, length_fastq=4 * n)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "first %d pairs of the synthetic workload per step, %d worker threads (C port of CLI3: split, demultiplex, filter)" % (n, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python and is not present on the GPU box; this is its algorithm restated in C "
                    "(oracle/ssd_oracle.c), orders of magnitude faster than the Python original (~3e3 pairs/s/core, BASELINE.md)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer loop (default min(steps, 5))")
    ap.add_argument("--e2e-files", action="store_true",
                    help="also time the file-to-file entry point (ssd_run_files) on files written to --files-dir (default /dev/shm)")
    ap.add_argument("--files-dir", default="/dev/shm")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="measure the end-to-end leg one batch at a time only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import ssd_b200
    from ssd_b200.distributed import global_filter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = load_whitelist()
    n = PAIRS_PER_GPU
    dm = ssd_b200.Demultiplexer(wl, errors=ERRORS, min_count=MIN_COUNT, device=local, record_bytes_hint=160)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream: the library, torch and the timing events all use it
    torch.cuda.set_stream(stream)
    dm.set_stream(stream.cuda_stream)
    spec = ssd_b200.synth_spec(wl, n, seed=SEED, first_index=rank * n + 1, read_len=READ_LEN)
    d1, d2 = ssd_b200.synth_device(dm, spec)
    out = torch.empty(d1.numel() + 64 * n + 4096, dtype=torch.uint8, device="cuda")
    last = {}

    def step():
        dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
        st = global_filter(dm, world)  # N > 1: all-reduce of the read histogram, all-to-all of the keys, all-reduce of the distinct table
        st["written"] = dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
        last.update(st)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dm.set_timing(True)
    dm.read_timings()
    launches0 = dm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = dm.launch_count() - launches0
    stage = dm.read_timings()
    dm.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms / 1e3)

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 6)
    e2e_value, e2e_serial, d2h, e2e_note, files_value = None, None, 0, None, None
    if e2e_steps > 0:
        h1 = torch.empty(d1.numel(), dtype=torch.uint8, pin_memory=True)
        h2 = torch.empty(d2.numel(), dtype=torch.uint8, pin_memory=True)
        h1.copy_(d1)
        h2.copy_(d2)
        hout = torch.empty(int(last["written"]) + 4096, dtype=torch.uint8, pin_memory=True)
        a1, a2, aout = h1.numpy(), h2.numpy(), hout.numpy()
    if e2e_steps <= 0:
        pass
    elif world == 1:
        # (a) one batch at a time: H2D -> kernels -> D2H, strictly serial
        res, st = dm.run_host(a1, a2, ssd_b200.EMIT_PASSING, out=aout)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res, st = dm.run_host(a1, a2, ssd_b200.EMIT_PASSING, out=aout)
        torch.cuda.synchronize()
        e2e_serial = n * e2e_steps / (time.perf_counter() - t0)
        d2h = int(res.size)
        e2e_value = e2e_serial
        e2e_note = "one batch at a time: H2D, kernels, D2H"
        if not args.no_e2e_pipeline:
            # (b) two contexts, two host threads, steps handed out alternately: while one batch uploads the previous one
            #     computes and downloads (PCIe is full duplex; the library gives each direction to one context at a time).
            #     Every step still copies its inputs host->device and its result device->host inside the timed region.
            #     The reference runs its chunks through joblib workers the same way (CLI3:56-64).
            pipe = ssd_b200.HostPipeline(wl, depth=2, errors=ERRORS, min_count=MIN_COUNT, device=local, record_bytes_hint=160)
            hout2 = torch.empty(hout.numel(), dtype=torch.uint8, pin_memory=True)
            outs = [aout, hout2.numpy()]
            pipe.map([(a1, a2)] * 2, ssd_b200.EMIT_PASSING, outs=outs)  # warm-up: allocations of the contexts
            total = max(4, 2 * e2e_steps)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            results = pipe.map([(a1, a2)] * total, ssd_b200.EMIT_PASSING, outs=outs)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert all(int(r.size) == d2h for r, _ in results), "pipelined runs returned different output sizes"
            pipe.close()
            piped = n * total / dt
            if piped > e2e_value:
                e2e_value = piped
                e2e_note = ("two contexts in flight on two host threads (each step: H2D of its inputs, kernels, D2H of its result); "
                            "one batch at a time gives serial_value")
        if args.e2e_files:
            # (c) file to file: the two FASTQ files in args.files_dir (RAM-backed by default, so that this measures the
            #     library's ingest/egress path and not a disk), output written next to them
            import tempfile
            tmp = tempfile.mkdtemp(prefix="ssd_bench_", dir=args.files_dir)
            try:
                f1, f2, fo = os.path.join(tmp, "r1.fastq"), os.path.join(tmp, "r2.fastq"), os.path.join(tmp, "out.fastq")
                a1.tofile(f1)
                a2.tofile(f2)
                dm.run_files(f1, f2, fo, ssd_b200.EMIT_PASSING, append=False)
                t0 = time.perf_counter()
                reps = max(1, min(3, e2e_steps))
                for _ in range(reps):
                    wrote, _ = dm.run_files(f1, f2, fo, ssd_b200.EMIT_PASSING, append=False)
                files_value = n * reps / (time.perf_counter() - t0)
                assert wrote == d2h
            finally:
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
    else:
        # every rank moves its own shard host->device, runs the distributed step, and reads its output back
        def e2e_step():
            d1.copy_(h1, non_blocking=True)
            d2.copy_(h2, non_blocking=True)
            step()
            hout[: last["written"]].copy_(out[: last["written"]], non_blocking=True)
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        sync_all()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * n * e2e_steps / float(t.item())
        e2e_serial = e2e_value
        d2h = int(last["written"])
    clocks = sampler.stop() if rank == 0 else {}

    # ---- roofline of the dominant kernel (algorithmic bytes = bytes read once + bytes written once) ----
    peak, peak_src = measured_peak()
    in1, in2, outb = d1.numel(), d2.numel(), int(last["written"])
    kept_frac = last["n_retained"] / max(1, last["n_records_r1"])
    algo = {"scan_r1": in1, "scan_r2": in2, "emit": int(kept_frac * in1) + outb}
    per_kernel = {}
    for k in ("scan_r1", "scan_r2", "emit"):
        t_ms, cnt = stage[k]
        if cnt:
            per_kernel[k] = {"ms": t_ms / cnt, "algorithmic_bytes": algo[k], "gbs": algo[k] / (t_ms / cnt * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"]) if per_kernel else None
    pipeline_bytes = in1 + in2 + outb
    pipeline_gbs = pipeline_bytes * args.steps / (ms * 1e-3) / 1e9
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": {"scan_r1": "k_scan<R1>", "scan_r2": "k_scan<R2>", "emit": "k_emit_stream"}[dom],
                    "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": per_kernel[dom]["gbs"] / peak,
                    "traffic": ncu_traffic(dom, n), "peak_source": peak_src,
                    "per_kernel": per_kernel,
                    "stage_ms_per_step": {k: v[0] / max(1, args.steps) for k, v in stage.items()},
                    "pipeline": {"algorithmic_bytes_per_step": pipeline_bytes, "achieved": pipeline_gbs, "frac": pipeline_gbs / peak}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = int(os.environ.get("SSD_BENCH_CPU_PAIRS", 2_000_000))
        rate, times, _ = oracle_rate(min(n_sample, n), cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "first %d pairs of the same synthetic workload, one pass, %d worker threads; C restatement of the "
                                  "reference (oracle/ssd_oracle.c) — the Python reference itself runs at ~3e3 pairs/s/core" % (min(n_sample, n), cores),
                        "seconds": min(times)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(in1 + in2) * world, "d2h_bytes_per_step": d2h * world,
                        "steps": e2e_steps, "serial_value": e2e_serial, "files_value": files_value, "note": e2e_note},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "result": {k: int(v) for k, v in last.items()}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
