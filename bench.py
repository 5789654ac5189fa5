#!/usr/bin/env python
"""bench.py — read pairs/sec demultiplexed on B200, device-resident and end-to-end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5]

One "step" is one pass of the whole hot path over one batch of synthetic read pairs: FASTQ scan of both mates,
barcode/UMI extraction, Hamming<=e matching, per-cell histogram, distinct-UMI -m 10 filter and the writer.
The default workload is BASELINE.json configs[1] (--config 2: synthetic 10 M read pairs, 150 bp, e=1, merged output);
the other configurations of BASELINE.json are --config 3 (100 M pairs, e=2, N bases, planted 2/2 ties, truncated read 2,
RanHex/OligoDT collapse on), --config 4 (10^9 pairs sharded over the ranks, batches, two passes) and --config 5 (100 M
pairs, split mode: per-cell buckets).  A default run at N = 1 also runs configurations 3, 4 (one tenth) and 5 briefly in
child processes and reports them under "configs"; at N > 1 configuration 4 follows the main measurement in-process.
With N > 1 (torchrun, one rank per GPU) every rank holds its own record range (weak scaling; configuration 4 shards a
fixed 10^9 pairs: strong), the per-cell histograms are all-reduced over NCCL and the (cell, UMI) keys are exchanged by
owner rank so the -m decision is global and exact; after the timed region a small sharded pass is checked against the
oracle on the concatenated input ("parity_check").

`value` times the steps with inputs resident in HBM (CUDA events, max over ranks); `e2e` times the same pass through
ssd_run_host with pinned HOST buffers (H2D and D2H inside the timed region); `e2e.files_value` goes file to file
(ssd_run_files on RAM-backed files).  `--impl reference` times the reference's CPU algorithm (the C oracle port; the
reference itself is Python and does not exist on the GPU box) on the host cores and never loads the product library.
"""
import argparse
import hashlib
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

UNIT = "read pairs/s"
READ_LEN = 150
MIN_COUNT = 10
SCALE = float(os.environ.get("SSD_BENCH_SCALE", "1"))  # < 1 shrinks every workload (smoke runs of this script)

CONFIGS = {
    2: dict(pairs=10_000_000, e=1, seed=0x5EED0001, spec={}, dm={}, mode="merged", scaling="weak",
            what="BASELINE configs[1]: synthetic %d read pairs per GPU (150 bp both mates, SURVEY 8d shape), e=1, -m 10 (distinct UMIs), merged output"),
    3: dict(pairs=100_000_000, e=2, seed=0x5EED0002, spec=dict(cell_pool=1_000_000, p_n=0.01, p_tie=0.01, p_trunc=0.005),
            dm=dict(collapse_pairs=48), mode="merged", scaling="weak",
            what="BASELINE configs[2]: synthetic %d read pairs per GPU, e=2, N bases (1 %%), planted 2/2 ties, truncated read 2, "
                 "RanHex/OligoDT collapse on, -m 10, merged output"),
    4: dict(pairs=1_000_000_000, e=1, seed=0x5EED0003, spec=dict(cell_pool=1_000_000), dm={}, mode="batches", scaling="strong",
            batch=100_000_000,
            what="BASELINE configs[3]: synthetic %d read pairs IN TOTAL sharded over the ranks, generated on the device batch by batch "
                 "(<= 10^8 pairs), two passes (global per-cell tables, then filter + write), e=1, -m 10"),
    5: dict(pairs=100_000_000, e=1, seed=0x5EED0004, spec=dict(cell_pool=1_000_000), dm={}, mode="split", scaling="weak",
            what="BASELINE configs[4]: synthetic %d read pairs per GPU, split mode: post-filter records bucketed per cell over the 96^3 "
                 "barcode space (buckets + index on the device), e=1, -m 10"),
}


def metric_name(cfg):
    return "read pairs/sec demultiplexed (e=%d)" % cfg["e"]


def cfg_pairs(cfg):
    return max(1000, int(int(os.environ.get("SSD_BENCH_PAIRS", cfg["pairs"])) * SCALE))


def load_whitelist():
    with open(os.path.join(ROOT, "tests", "golden", "barcodes_8mer.txt")) as fh:
        return [l.strip() for l in fh if l.strip()]


def workload_config(cid, cfg, n_gpus):
    n = cfg_pairs(cfg)
    per_gpu = n // n_gpus if cfg["scaling"] == "strong" else n
    return {"workload": cfg["what"] % n, "baseline_config": cid, "pairs_per_gpu": per_gpu, "read_len": READ_LEN, "errors": cfg["e"],
            "min_count": MIN_COUNT, "filter": "distinct_umi", "seed": hex(cfg["seed"]), "mode": cfg["mode"],
            "sharding": "record ranges, %d rank(s)" % n_gpus,
            "l2": "inputs (%.1f GB per GPU and batch) are far larger than the 126 MB L2; no explicit flush" % (min(per_gpu, cfg.get("batch", per_gpu)) * 654e-9)}


def ncu_traffic(kernel_key, cid, pairs_per_gpu):
    """DRAM bytes (read + written) per launch of the kernel from the committed `ncu --set full` capture of this very
    workload (profiles/r0N_dram_traffic.json, made by profiles/extract_traffic.py); None for any other workload."""
    if cid != 2:
        return None, None
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                t = json.load(fh)
            if t["workload"]["pairs_per_gpu"] != pairs_per_gpu or t["workload"]["read_len"] != READ_LEN:
                continue
            return t["kernels"][kernel_key]["traffic"], "profiles/" + name
        except (OSError, KeyError, ValueError):
            continue
    return None, None


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
# reference arm / cpu baseline: the reference's CPU algorithm on the host cores (oracle/ only: no product code)
# ------------------------------------------------------------------------------------------------
def python_reference_figure(cid):
    """What the UNMODIFIED Python reference (CLI3) did on the head of this workload in the authoring container -- the GPU box
    has no /root/reference -- as tests/golden/make_reference_on_synthetic.py measured and committed it; None when there is none."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "reference_on_synthetic.json")) as fh:
            ref = json.load(fh)
        case = next(c for c in ref["cases"] if c["name"] == "config%d-head" % cid)
        return {"value": case["pairs_per_s"], "unit": UNIT, "cores": ref["host"]["cores"], "cpu": ref["host"]["cpu"],
                "where": "authoring container (not this box)", "sample": "first %d pairs of the workload, %s" % (case["pairs"], ref["reference"]),
                "equals_oracle_output": bool(case["oracle"]["equal"])}
    except Exception:  # noqa: BLE001 - a quoted figure, never worth a failed run
        return None


def oracle_input(cfg, n_pairs):
    from oracle import ssd_oracle_py as orc
    r1, r2 = orc.synth_pairs(load_whitelist(), n_pairs, seed=cfg["seed"], read_len=READ_LEN, **cfg["spec"])
    return r1.tobytes(), r2.tobytes()


def oracle_rate(cfg, n_pairs, workers, reps=1, data=None):
    """pairs/s of the CPU port (CLI3 structure: split -> per-chunk demultiplex on `workers` threads -> filter)."""
    from oracle import ssd_oracle_py as orc
    r1, r2 = data or oracle_input(cfg, n_pairs)
    wl = load_whitelist()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.run_cli(r1, r2, wl, cfg["e"], MIN_COUNT, n_workers=workers, length_fastq=4 * n_pairs)
        times.append(time.perf_counter() - t0)
    return n_pairs / min(times), times


def run_reference_arm(args, cid, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # size the per-step sample so that the whole run stays within a couple of minutes
    rate, _ = oracle_rate(cfg, 50_000, cores)
    budget_s = 100.0
    n = int(max(20_000, min(2_000_000, cfg_pairs(cfg), budget_s * rate / max(1, args.steps + args.warmup))))
    data = oracle_input(cfg, n)
    for _ in range(args.warmup):
        oracle_rate(cfg, n, cores, data=data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_rate(cfg, n, cores, data=data)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = ("first %d pairs of the synthetic workload per step (generated by oracle/ssd_synth_ref.c), %d worker threads "
              "(C port of CLI3: split, demultiplex, filter)" % (n, cores))
    config = workload_config(cid, cfg, args.gpus)
    config["reference_sample_pairs_per_step"] = n
    line = {"impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "python_reference": python_reference_figure(cid)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python and is not present on the GPU box; this is its algorithm restated in C "
                    "(oracle/ssd_oracle.c), orders of magnitude faster than the Python original (~3e3 pairs/s/core in the authoring "
                    "container, BASELINE.md)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Env:
    pass


def sync_all(env):
    import torch
    torch.cuda.synchronize()
    if env.world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(env, x):
    import torch
    if env.world == 1:
        return float(x)
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_dm(env, cfg, **extra):
    import ssd_b200
    dm = ssd_b200.Demultiplexer(env.wl, errors=cfg["e"], min_count=MIN_COUNT, device=env.local, record_bytes_hint=160, **cfg["dm"], **extra)
    dm.set_stream(env.stream.cuda_stream)
    return dm


def measure_resident(env, cid, cfg, args):
    """configs 2, 3, 5: the rank's pairs resident in HBM, K timed steps.  Returns the pieces of the JSON line."""
    import torch
    import ssd_b200
    from ssd_b200.distributed import global_filter
    n = cfg_pairs(cfg)
    dm = make_dm(env, cfg)
    spec = ssd_b200.synth_spec(env.wl, n, seed=cfg["seed"], first_index=env.rank * n + 1, read_len=READ_LEN, **cfg["spec"])
    d1, d2 = ssd_b200.synth_device(dm, spec)
    out = torch.empty(d1.numel() + 64 * n + 4096, dtype=torch.uint8, device="cuda")
    last = {}
    split = cfg["mode"] == "split"

    def step():
        dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
        st = global_filter(dm, env.world)  # N > 1: all-reduce of the read histogram, exchange of the keys, all-reduce of the distinct table
        if split:
            st["written"], table, _ = dm.emit_split_device(out.data_ptr(), out.numel())
            st["n_buckets"] = len(table)
        else:
            st["written"] = dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
        last.update(st)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    dm.set_timing(True)
    dm.read_timings()
    launches0 = dm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all(env)
    ev0.record(env.stream)
    for _ in range(args.steps):
        step()
    ev1.record(env.stream)
    sync_all(env)
    ms = max_over_ranks(env, ev0.elapsed_time(ev1))
    launches = dm.launch_count() - launches0
    stage = dm.read_timings()
    dm.set_timing(False)
    value = env.world * n * args.steps / (ms / 1e3)
    short_fast, exact_records = dm.scan_variant()  # which read-2 scan the timed steps ran (the context picks it from its data)

    in1, in2 = d1.numel(), d2.numel()
    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 0}
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 6)
    if e2e_steps > 0:
        # a bounded batch for the 100 M-pair configurations (pinning 100 GB of host memory is not what this leg is about)
        ne = min(n, int(20_000_000 * SCALE) or n)
        if ne == n:
            h1 = torch.empty(d1.numel(), dtype=torch.uint8, pin_memory=True)
            h2 = torch.empty(d2.numel(), dtype=torch.uint8, pin_memory=True)
            h1.copy_(d1)
            h2.copy_(d2)
        else:
            # the resident batch has done its job: its memory goes to the contexts of this leg
            del d1, d2, out
            dm.close()
            torch.cuda.empty_cache()
            dm = make_dm(env, cfg)
            out = torch.empty(int(ne * 400) + 4096, dtype=torch.uint8, device="cuda")
            s1, s2 = ssd_b200.synth_device(dm, ssd_b200.synth_spec(env.wl, ne, seed=cfg["seed"], first_index=env.rank * n + 1, read_len=READ_LEN, **cfg["spec"]))
            h1 = torch.empty(s1.numel(), dtype=torch.uint8, pin_memory=True)
            h2 = torch.empty(s2.numel(), dtype=torch.uint8, pin_memory=True)
            h1.copy_(s1)
            h2.copy_(s2)
            del s1, s2
        a1, a2 = h1.numpy(), h2.numpy()
        hout = torch.empty(h1.numel() + 64 * ne + 4096, dtype=torch.uint8, pin_memory=True)
        aout = hout.numpy()
        e2e.update(pairs_per_step_per_gpu=ne)
        if split:
            # split mode end to end: upload, bucket on the device, download buckets + index
            def e2e_step():
                t1 = torch.empty(h1.numel(), dtype=torch.uint8, device="cuda")
                t2 = torch.empty(h2.numel(), dtype=torch.uint8, device="cuda")
                t1.copy_(h1, non_blocking=True)
                t2.copy_(h2, non_blocking=True)
                dm.demux_device(t1.data_ptr(), t1.numel(), t2.data_ptr(), t2.numel(), keep=(t1, t2))
                global_filter(dm, env.world)
                w, table, _ = dm.emit_split_device(out.data_ptr(), out.numel())
                hout[:w].copy_(out[:w], non_blocking=True)
                torch.cuda.synchronize()
                return w
            e2e_step()
            sync_all(env)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                w = e2e_step()
            sync_all(env)
            dt = max_over_ranks(env, time.perf_counter() - t0)
            e2e.update(value=env.world * ne * e2e_steps / dt, h2d_bytes_per_step=int(h1.numel() + h2.numel()) * env.world, d2h_bytes_per_step=int(w) * env.world,
                       steps=e2e_steps, note="upload, per-cell bucketing on the device, download of the buckets")
        else:
            # every rank: ssd_run_host on its own pinned buffers (H2D, kernels, D2H inside the call); N > 1: the -m decision of
            # this leg is per rank (the C ABI's single call has no collective), the copies and kernels are the same
            dme = make_dm(env, cfg) if env.world > 1 else dm
            if env.world > 1:
                dme.set_stream(0)
            res, _ = dme.run_host(a1, a2, ssd_b200.EMIT_PASSING, out=aout)
            sync_all(env)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                res, _ = dme.run_host(a1, a2, ssd_b200.EMIT_PASSING, out=aout)
            torch.cuda.synchronize()
            dt = max_over_ranks(env, time.perf_counter() - t0)
            serial = env.world * ne * e2e_steps / dt
            d2h = int(res.size)
            e2e.update(value=serial, serial_value=serial, h2d_bytes_per_step=int(h1.numel() + h2.numel()) * env.world, d2h_bytes_per_step=d2h * env.world,
                       steps=e2e_steps, note="one batch at a time per rank: H2D, kernels, D2H")
            if not args.no_e2e_pipeline:
                # two contexts per rank, two host threads, steps handed out alternately: while one batch uploads the previous one
                # computes and downloads (PCIe is full duplex; the library gives each direction to one context at a time).
                # Every step still copies its inputs host->device and its result device->host inside the timed region.
                # The reference runs its chunks through joblib workers the same way (CLI3:56-64).
                pipe = ssd_b200.HostPipeline(env.wl, depth=2, errors=cfg["e"], min_count=MIN_COUNT, device=env.local, record_bytes_hint=160, **cfg["dm"])
                hout2 = torch.empty(hout.numel(), dtype=torch.uint8, pin_memory=True)
                outs = [aout, hout2.numpy()]
                pipe.map([(a1, a2)] * 2, ssd_b200.EMIT_PASSING, outs=outs)  # warm-up: allocations of the contexts
                total = max(4, 2 * e2e_steps)
                sync_all(env)
                t0 = time.perf_counter()
                results = pipe.map([(a1, a2)] * total, ssd_b200.EMIT_PASSING, outs=outs)
                torch.cuda.synchronize()
                dt = max_over_ranks(env, time.perf_counter() - t0)
                assert all(int(r.size) == d2h for r, _ in results), "pipelined runs returned different output sizes"
                pipe.close()
                del hout2
                piped = env.world * ne * total / dt
                e2e.update(pipelined_value=piped, pipelined_steps=total)
                if piped > e2e["value"]:
                    e2e.update(value=piped, steps=total,
                               note="two contexts per rank in flight on two host threads (each step: H2D of its inputs, kernels, D2H of its result); "
                                    "one batch at a time gives serial_value")
            if env.world == 1 and not args.no_e2e_files:
                # file to file: the two FASTQ files in args.files_dir (RAM-backed by default, so that this measures the
                # library's ingest/egress path and not a disk), output written next to them
                tmp = tempfile.mkdtemp(prefix="ssd_bench_", dir=args.files_dir)
                try:
                    f1, f2, fo = os.path.join(tmp, "r1.fastq"), os.path.join(tmp, "r2.fastq"), os.path.join(tmp, "out.fastq")
                    a1.tofile(f1)
                    a2.tofile(f2)
                    t0 = time.perf_counter()
                    dm.run_files(f1, f2, fo, ssd_b200.EMIT_PASSING, append=False)  # first call: a NEW output file (and the pinned rings)
                    first_s = time.perf_counter() - t0
                    reps = max(1, min(3, e2e_steps))
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        wrote, _ = dm.run_files(f1, f2, fo, ssd_b200.EMIT_PASSING, append=False)
                    serial_files = ne * reps / (time.perf_counter() - t0)
                    assert wrote == d2h
                    # two contexts on two host threads, like the host-buffer leg: one batch reads + uploads while the previous one
                    # downloads + writes (each into its own output file)
                    import threading
                    ctxs = [dm, make_dm(env, cfg)]
                    outs_f = [fo, os.path.join(tmp, "out2.fastq")]
                    ctxs[1].run_files(f1, f2, outs_f[1], ssd_b200.EMIT_PASSING, append=False)
                    total_f = 2 * max(2, reps)
                    nxt, lock, sizes = [0], threading.Lock(), []

                    def work(k):
                        while True:
                            with lock:
                                i = nxt[0]
                                nxt[0] += 1
                            if i >= total_f:
                                return
                            w, _ = ctxs[k].run_files(f1, f2, outs_f[k], ssd_b200.EMIT_PASSING, append=False)
                            sizes.append(w)
                    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
                    t0 = time.perf_counter()
                    for t in th:
                        t.start()
                    for t in th:
                        t.join()
                    piped_files = ne * total_f / (time.perf_counter() - t0)
                    assert all(w == d2h for w in sizes)
                    ctxs[1].close()
                    e2e.update(files_value=max(serial_files, piped_files), files_serial_value=serial_files, files_pipelined_value=piped_files,
                               files_first_call_value=ne / first_s, files_steps=total_f if piped_files > serial_files else reps,
                               files_note="ssd_run_files on RAM-backed files in %s (6.5 GB in, 3.1 GB out per batch): pread into pinned rings, H2D, "
                                          "kernels, D2H, memcpy into the mapped output file; serial = one batch at a time, pipelined = two contexts "
                                          "on two host threads; first_call = the very first call (new output file: every page allocated by the "
                                          "kernel, pinned rings set up)" % args.files_dir)
                except OSError as err:
                    e2e.update(files_value=None, files_note="skipped: %s" % err)
                finally:
                    import shutil
                    shutil.rmtree(tmp, ignore_errors=True)
        del h1, h2, hout
    clocks = sampler.stop() if env.rank == 0 else {}

    # ---- roofline (algorithmic bytes = every input byte read once + every output byte written once) ----
    peak, peak_src = measured_peak()
    outb = int(last["written"])
    algo = {"scan_r1": in1, "scan_r2": in2, "emit": outb}  # the writer's re-read of read 1 is not credited
    kname = {"scan_r1": "k_scan<R1>", "scan_r2": "k_scan<R2>", "emit": "k_emit (split)" if split else "k_emit_stream"}
    per_kernel = {}
    for k in ("scan_r1", "scan_r2", "emit"):
        t_ms, cnt = stage[k]
        if cnt:
            per_kernel[k] = {"kernel": kname[k], "ms": t_ms / cnt, "algorithmic_bytes": algo[k], "gbs": algo[k] / (t_ms / cnt * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"]) if per_kernel else None
    pipeline_bytes = in1 + in2 + outb
    pipeline_gbs = pipeline_bytes * args.steps / (ms * 1e-3) / 1e9
    roofline = None
    if dom:
        traffic, traffic_src = ncu_traffic(dom, cid, n)
        roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": per_kernel[dom]["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "per_kernel": per_kernel,
                    "stage_ms_per_step": {k: v[0] / max(1, args.steps) for k, v in stage.items()},
                    "scan_r2_variant": {"short_fast": bool(short_fast), "exact_path_records_per_step": int(exact_records)},
                    "pipeline": {"algorithmic_bytes_per_step": pipeline_bytes, "achieved": pipeline_gbs, "frac": pipeline_gbs / peak}}
    dm.close()
    d1 = d2 = out = None
    torch.cuda.empty_cache()
    return dict(value=value, ms_per_step=ms / args.steps, e2e=e2e, gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                result={k: int(v) for k, v in last.items()})


def measure_batches(env, cid, cfg, args):
    """config 4: the rank's shard of the 10^9 pairs goes through the GPU in batches generated on the device, twice (global
    per-cell tables first, then filter + write).  A step is one whole job; the generator's time (CUDA events) is excluded."""
    import torch
    import ssd_b200
    from ssd_b200.distributed import demux_batches
    total_pairs = cfg_pairs(cfg)
    per_rank = total_pairs // env.world
    batch = max(1000, int(cfg["batch"] * SCALE))
    first_rank = 1 + env.rank * per_rank
    dm = make_dm(env, cfg)
    gen_ms = [0.0]
    checks = {"bytes": 0, "newlines": 0}

    # the generator writes every batch into the same two buffers (sized for the largest batch with its longest ids)
    sizes = [ssd_b200.synth_sizes(ssd_b200.synth_spec(env.wl, min(batch, per_rank - lo), seed=cfg["seed"], first_index=first_rank + lo, read_len=READ_LEN, **cfg["spec"]))
             for lo in range(0, per_rank, batch)]
    bufs = (torch.empty(max(a for a, _ in sizes) + 16, dtype=torch.uint8, device="cuda"), torch.empty(max(b for _, b in sizes) + 16, dtype=torch.uint8, device="cuda"))

    def batches():
        done = 0
        while done < per_rank:
            nb = min(batch, per_rank - done)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            d = ssd_b200.synth_device(dm, ssd_b200.synth_spec(env.wl, nb, seed=cfg["seed"], first_index=first_rank + done, read_len=READ_LEN, **cfg["spec"]), into=bufs)
            e1.record()
            e1.synchronize()
            gen_ms[0] += e0.elapsed_time(e1)
            if env.world > 1:
                # the ranks' generators do not take the same time: without this, the first collective of the batch would wait
                # for the slowest generator and the wait would count as the job's time
                import torch.distributed as dist
                tb = time.perf_counter()
                dist.barrier()
                torch.cuda.synchronize()
                gen_ms[0] += (time.perf_counter() - tb) * 1e3
            yield d
            del d
            done += nb

    def sink(out, st):
        checks["bytes"] += out.numel()
        checks["newlines"] += dm.count_newlines(out)

    def small():
        yield ssd_b200.synth_device(dm, ssd_b200.synth_spec(env.wl, min(1_000_000, per_rank), seed=1, first_index=1 + env.rank * 1_000_000))
    demux_batches(dm, small, world=env.world)  # warm-up: NCCL connections, module loading
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    steps = max(1, min(args.steps, 2))
    launches0 = dm.launch_count()
    sync_all(env)
    t0 = time.perf_counter()
    gen_ms[0] = 0.0
    timing, total = {}, {}
    for _ in range(steps):
        checks.update(bytes=0, newlines=0)
        total = demux_batches(dm, batches, world=env.world, sink=sink, timing=timing)
    sync_all(env)
    wall = time.perf_counter() - t0
    compute = max_over_ranks(env, wall - gen_ms[0] / 1e3)
    launches = dm.launch_count() - launches0
    clocks = sampler.stop() if env.rank == 0 else {}
    hist = dm.reads_histogram().to(torch.int64)  # the global tables are still in place after the last batch
    distinct = dm.distinct_histogram().to(torch.int64)
    passing = distinct >= MIN_COUNT
    t = torch.tensor([total.get("n_retained", 0), total.get("n_matched", 0), checks["bytes"], checks["newlines"], total.get("n_records_r1", 0)],
                     dtype=torch.float64, device="cuda")
    if env.world > 1:
        import torch.distributed as dist
        dist.all_reduce(t)
    retained, matched, nbytes, newlines, records = (int(x) for x in t.tolist())
    ok = {"hist_sum_equals_matched": int(hist.sum()) == matched,
          "retained_equals_reads_of_passing_cells": int(hist[passing].sum()) == retained,
          "cells": int((hist > 0).sum()) == total.get("n_cells") and int(passing.sum()) == total.get("n_passing_cells"),
          "four_lines_per_record": newlines == 4 * retained, "all_records_seen": records == per_rank * env.world}
    value = records * steps / compute
    peak, peak_src = measured_peak()
    algo_bytes = 2 * 654 * records + nbytes  # both passes read both files; the output is written once (approximate record size 327 B)
    roofline = {"bound": "hbm", "kernel": "two passes over batches (k_scan<R1|R2> twice, k_emit_stream once)", "achieved": algo_bytes / env.world * steps / compute / 1e9,
                "peak": peak, "unit": "GB/s", "frac": algo_bytes / env.world * steps / compute / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "note": "algorithmic bytes of the two-pass job per GPU: 2 x (both inputs) + output"}
    dm.close()
    del bufs
    torch.cuda.empty_cache()
    return dict(value=value, ms_per_step=1e3 * compute / steps, steps_run=steps,
                e2e={"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 0,
                     "note": "the 670 GB of input of this configuration exist only on the device (generated per batch); the host-buffer path is measured by configs 2 and 3"},
                gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                result={"pairs": records, "n_matched": matched, "n_retained": retained, "n_cells": total.get("n_cells"), "n_passing_cells": total.get("n_passing_cells"),
                        "output_bytes": nbytes, "generator_s_excluded": gen_ms[0] / 1e3, "rank0_phases_s": timing, "batches_per_rank": -(-per_rank // batch),
                        "checks": ok})


def parity_check(env, cfg):
    """N > 1, after the timed region: a small sharded pass (200 k pairs per rank, global -m) whose outputs are gathered on rank
    0 and compared with the oracle on the concatenated input: byte-identical text (SHA-256) and the same counts."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import ssd_b200
    from ssd_b200.distributed import global_filter
    n = max(2000, int(200_000 * SCALE))
    dm = make_dm(env, cfg)
    spec = ssd_b200.synth_spec(env.wl, n, seed=cfg["seed"] ^ 0xC0FFEE, first_index=env.rank * n + 1, read_len=READ_LEN, cell_pool=5000,
                               **{k: v for k, v in cfg["spec"].items() if k != "cell_pool"})
    d1, d2 = ssd_b200.synth_device(dm, spec)
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
    st = global_filter(dm, env.world)
    out = torch.empty(st["bytes_passing"] + 16, dtype=torch.uint8, device="cuda")
    w = dm.emit_device(ssd_b200.EMIT_PASSING, out.data_ptr(), out.numel())
    torch.cuda.synchronize()
    mine = out[:w].cpu().numpy().tobytes()
    gathered = [None] * env.world if env.rank == 0 else None
    dist.gather_object((mine, st["n_matched"], st["n_retained"], st["n_cells"], st["n_passing_cells"]), gathered, dst=0)
    dm.close()
    if env.rank != 0:
        return None
    from oracle import ssd_oracle_py as orc
    r1, r2 = orc.synth_pairs(env.wl, n * env.world, seed=cfg["seed"] ^ 0xC0FFEE, read_len=READ_LEN, cell_pool=5000,
                             **{k: v for k, v in cfg["spec"].items() if k != "cell_pool"})
    ref = orc.run_cli(r1.tobytes(), r2.tobytes(), env.wl, cfg["e"], MIN_COUNT, n_workers=os.cpu_count() or 1, length_fastq=4 * n * env.world)
    text = b"".join(g[0] for g in gathered)
    got = (sum(g[1] for g in gathered), sum(g[2] for g in gathered), gathered[0][3], gathered[0][4])
    want = (ref.n_matched, ref.n_retained, ref.n_cells, ref.n_passing_cells)
    sha_got, sha_want = hashlib.sha256(text).hexdigest(), hashlib.sha256(ref.passing).hexdigest()
    ok = sha_got == sha_want and got == want and all(g[3:] == gathered[0][3:] for g in gathered)
    if cfg["dm"].get("collapse_pairs"):
        # the oracle's text is not collapsed: compare what the collapse leaves alone (record count, bytes) and the counts
        ok = len(text) == len(ref.passing) and got == want
    return {"ok": bool(ok), "n_ranks": env.world, "pairs": n * env.world, "sha256": sha_got, "oracle_sha256": sha_want,
            "counts": {"matched": got[0], "retained": got[1], "cells": got[2], "passing_cells": got[3]},
            "oracle_counts": {"matched": want[0], "retained": want[1], "cells": want[2], "passing_cells": want[3]}}


def child_config(cid, steps, timeout_s, scale_env=None):
    """Run another configuration in a child process (its own CUDA context, so nothing of it can take the main line down)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--config", str(cid), "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline",
           "--no-other-configs", "--e2e-steps", "2", "--no-e2e-files"]
    env = dict(os.environ)
    env.update(scale_env or {})
    t0 = time.perf_counter()
    try:
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s, env=env)
        lines = [l for l in proc.stdout.strip().splitlines() if l.startswith("{")]
        if proc.returncode != 0 or not lines:
            return {"error": "exit %d: %s" % (proc.returncode, proc.stderr.strip()[-300:])}
        line = json.loads(lines[-1])
        keep = {k: line.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "config", "gpu_launches", "result")}
        keep["e2e"] = {k: line["e2e"].get(k) for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps", "pairs_per_step_per_gpu", "note")}
        if line.get("roofline"):
            keep["roofline"] = {k: line["roofline"].get(k) for k in ("kernel", "achieved", "peak", "frac", "pipeline", "stage_ms_per_step", "scan_r2_variant")}
        keep["wall_s"] = time.perf_counter() - t0
        return keep
    except subprocess.TimeoutExpired:
        return {"error": "timed out after %d s" % timeout_s}
    except Exception as err:  # noqa: BLE001
        return {"error": repr(err)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer loop (default min(steps, 6))")
    ap.add_argument("--no-e2e-files", action="store_true", help="skip the file-to-file leg (ssd_run_files on files in --files-dir)")
    ap.add_argument("--e2e-files", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--files-dir", default="/dev/shm")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="measure the end-to-end leg one batch at a time only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="default runs also report configurations 3, 4, 5 briefly; this skips them")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    cid, cfg = args.config, CONFIGS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, cid, cfg)
        return

    import torch
    import torch.distributed as dist
    env = Env()
    env.rank = int(os.environ.get("RANK", "0"))
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(env.local)
    if env.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
    env.wl = load_whitelist()
    env.stream = torch.cuda.Stream()  # a real (non-legacy) stream: the library, torch and the timing events all use it
    torch.cuda.set_stream(env.stream)

    m = measure_batches(env, cid, cfg, args) if cfg["mode"] == "batches" else measure_resident(env, cid, cfg, args)

    parity = None
    if env.world > 1 and not args.no_parity_check and cfg["mode"] != "batches":
        parity = parity_check(env, cfg)

    others = None
    if cid == 2 and not args.no_other_configs:
        if env.world == 1:
            # the other configurations of BASELINE.json, briefly, each in its own process: 3 and 5 at full size (100 M pairs), 4 at a tenth
            # (10^8 pairs, one GPU; the 10^9 of the configuration are for 2/4/8 GPUs)
            torch.cuda.empty_cache()
            others = {"3": child_config(3, 3, 240), "5": child_config(5, 3, 240),
                      "4": child_config(4, 1, 300, {"SSD_BENCH_PAIRS": str(int(CONFIGS[4]["pairs"] // 10))})}
            others["4"]["note"] = "one tenth of configuration 4 (10^8 pairs, 10 batches of 10^7) on one GPU; the full 10^9 run at N > 1"
            if "config" in others["4"]:
                pass
        else:
            a4 = argparse.Namespace(**vars(args))
            a4.steps = 1
            try:
                r4 = measure_batches(env, 4, CONFIGS[4], a4)
                others = {"4": {"metric": metric_name(CONFIGS[4]), "value": r4["value"], "unit": UNIT, "ms_per_step": r4["ms_per_step"], "steps": r4["steps_run"],
                                "scaling": "strong", "config": workload_config(4, CONFIGS[4], env.world), "gpu_launches": r4["gpu_launches"],
                                "roofline": r4["roofline"], "result": r4["result"]}}
            except Exception as err:  # noqa: BLE001
                others = {"4": {"error": repr(err)}}

    cpu_baseline = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # about 10-30 s of CPU work: a quick probe gives the box's rate, the sample is sized for ~12 s (2 M .. 6 M pairs: <= 13 GB of host memory)
        probe, _ = oracle_rate(cfg, min(200_000, cfg_pairs(cfg)), cores)
        n_sample = min(int(os.environ.get("SSD_BENCH_CPU_PAIRS", max(2_000_000, min(6_000_000, int(12 * probe))))), cfg_pairs(cfg))
        rate, times = oracle_rate(cfg, n_sample, cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "first %d pairs of the same synthetic workload, one pass, %d worker threads; C restatement of the "
                                  "reference (oracle/ssd_oracle.c) — the Python reference itself runs at ~3e3 pairs/s/core" % (n_sample, cores),
                        "seconds": min(times), "python_reference": python_reference_figure(cid)}

    if env.rank == 0:
        if cfg["mode"] == "batches":
            steps = m["steps_run"]
        else:
            steps = args.steps
        line = {"metric": metric_name(cfg), "value": m["value"], "unit": UNIT, "n_gpus": env.world, "steps": steps, "warmup": args.warmup,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": workload_config(cid, cfg, env.world), "e2e": m["e2e"], "gpu_launches": m["gpu_launches"],
                "clocks": m["clocks"], "roofline": m["roofline"], "cpu_baseline": cpu_baseline, "result": m["result"]}
        if parity is not None:
            line["parity_check"] = parity
        if others is not None:
            line["configs"] = others
        print(json.dumps(line))
    if env.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
