"""CPU tests of the two command lines the driver and the reference's users call: bench.py's reference arm (must run without
the product library and print the contract's JSON line) and the option surface of the drop-in CLI (CLI3:21-39)."""
import importlib.util
import inspect
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "split-seq_demultiplexing_b200")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_the_contract_line_without_the_product_library():
    """`bench.py --impl reference`: the oracle port on the host cores, one JSON line with impl/cpu_baseline/e2e as the
    contract asks, and libssd_b200.so never mapped (LD_DEBUG=files lists every dlopen of the process)."""
    env = dict(os.environ, SSD_BENCH_SCALE="0.002", LD_DEBUG="files")
    env.pop("LD_DEBUG_OUTPUT", None)  # the trace goes to stderr
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "read pairs/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["metric"] == "read pairs/sec demultiplexed (e=1)" and line["config"]["baseline_config"] == 2
    assert line["config"]["reference_sample_pairs_per_step"] >= 20_000      # the sample is named next to the workload
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    loaded = [l for l in proc.stderr.splitlines() if "file=" in l and ".so" in l]
    assert any("libssd_oracle" in l for l in loaded), "LD_DEBUG produced no trace of the oracle library"
    assert not any("libssd_b200" in l for l in loaded), "the reference arm mapped the product library"


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the others print nothing and exit 0."""
    env = dict(os.environ, SSD_BENCH_SCALE="0.002", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env, cwd=ROOT)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_bench_configs_name_the_baseline_workloads():
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test")
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert sorted(bench.CONFIGS) == [2, 3, 4, 5] and len(base["configs"]) == 5
    assert bench.CONFIGS[2]["pairs"] == 10_000_000 and bench.CONFIGS[2]["e"] == 1
    assert bench.CONFIGS[3]["pairs"] == 100_000_000 and bench.CONFIGS[3]["e"] == 2 and bench.CONFIGS[3]["dm"]["collapse_pairs"] == 48
    assert bench.CONFIGS[4]["pairs"] == 1_000_000_000 and bench.CONFIGS[4]["scaling"] == "strong"
    assert bench.CONFIGS[5]["pairs"] == 100_000_000 and bench.CONFIGS[5]["mode"] == "split"
    for cid, cfg in bench.CONFIGS.items():
        c = bench.workload_config(cid, cfg, 4)
        assert "model" not in c and c["baseline_config"] == cid and c["workload"].startswith("BASELINE configs[%d]" % (cid - 1))
        assert c["pairs_per_gpu"] == (cfg["pairs"] // 4 if cfg["scaling"] == "strong" else cfg["pairs"])


# the reference's options, CLI3:22-38: (short, long, required)
CLI3_OPTIONS = [("-n", "--numCores", True), ("-e", "--errors", True), ("-m", "--minReads", True), ("-1", "--round1Barcodes", True),
                ("-2", "--round2Barcodes", True), ("-3", "--round3Barcodes", True), ("-f", "--fastqF", True), ("-r", "--fastqR", True),
                ("-o", "--outputDir", True), ("-a", "--align", True), ("-x", "--starGenome", False), ("-s", "--geneAnnotation", False),
                ("-i", "--indexType", False), ("-b", "--numReadsBin", True), ("-p", "--performanceMetrics", True), ("-l", "--lengthFastq", True),
                ("-k", "--positionDetection", False)]


@pytest.mark.parametrize("script", ["splitseqdemultiplex_0.2.3.py"])
def test_cli_accepts_exactly_the_reference_options(script):
    cli = _load(os.path.join(PKG, script), "cli_under_test")
    parser = cli.build_parser()
    by_long = {}
    for act in parser._actions:
        for s in act.option_strings:
            if s.startswith("--"):
                by_long[s] = act
    for short, long_, required in CLI3_OPTIONS:
        act = by_long[long_]
        assert short in act.option_strings and act.required == required, long_
    # -k is store_false with dest positionDetection: absent means detection ON (CLI3:38, 63)
    base = ["-n", "4", "-e", "1", "-m", "10", "-1", "a", "-2", "b", "-3", "c", "-f", "F", "-r", "R", "-o", "out", "-a", "", "-b", "100000",
            "-p", "True", "-l", "20000"]
    assert parser.parse_args(base).positionDetection is True
    assert parser.parse_args(base + ["-k"]).positionDetection is False
    # the bash wrappers' vocabulary (splitseqdemultiplex_0.2.2.sh) is accepted on top
    ns = parser.parse_args(base + ["-v", "fast", "-c", "true", "-t", "8000", "-g", "100000"])
    assert ns.version == "fast" and ns.collapseRandomHexamers == "true"
    # a required option missing is an error like in the reference
    with pytest.raises(SystemExit):
        parser.parse_args(base[2:])


def test_reference_named_modules_keep_the_reference_signatures():
    """V2:32, V2:384, LIB:6, 53, 113, 117: names and argument names a caller of the reference passes by keyword."""
    sys.path.insert(0, PKG)
    try:
        v2 = _load(os.path.join(PKG, "DemultiplexUsingBarcodes_New_V2.py"), "v2_under_test")
        lib = _load(os.path.join(PKG, "Splitseq_fun_lib.py"), "lib_under_test")
    finally:
        sys.path.remove(PKG)
    assert list(inspect.signature(v2.demultiplex_fun).parameters) == ["inputFastqF", "inputFastqR", "outputDir", "numReadsBin", "errorThreshold",
                                                                     "performanceMetrics", "readsPerCellThreshold", "positionDetection", "verbose"]
    assert inspect.signature(v2.demultiplex_fun).parameters["verbose"].default is False
    assert list(inspect.signature(v2.calc_demux_results).parameters) == ["outputDir", "performanceMetrics", "readsPerCellThreshold"]
    for name in ("split_fastqF_fun", "split_fastqR_fun"):
        assert list(inspect.signature(getattr(lib, name)).parameters) == ["split_num", "fastq_file", "lengthFastq"]
    for name in ("remove_file_fun", "remove_dir_fun"):
        assert list(inspect.signature(getattr(lib, name)).parameters) == ["filename"]
    for name in ("run_star_alignment_fun", "run_featureCounts_SAF_fun", "run_samtools_fun", "run_umi_tools_fun", "run_minimap2_alignment_fun"):
        with pytest.raises(NotImplementedError):                             # out of scope, but named
            getattr(lib, name)(resultsDir="results")
