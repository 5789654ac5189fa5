"""ctypes binding of libssd_b200.so (C ABI in include/ssd_b200.h).

The shared library is built in-tree from csrc/ with nvcc for sm_100a.  Nothing here computes: every
function forwards to the library, and loading fails loudly when the library is missing — there is no
CPU fallback for the demultiplexing path."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("SSD_B200_LIB") or os.path.join(CSRC, "libssd_b200.so")  # override: A/B builds while tuning

def sources():
    """Everything the library is compiled from: every .cu / .cuh under csrc/ (ssd_kernels.cuh includes them all)."""
    import glob
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))

HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "ssd_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "186"]  # 186: comparisons that are constant in some template instances

OK, EINVAL, ECUDA, EKEY, EMALFORMED, ENOMEM, ESTATE, ERANGE, EINDEX, EIO = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9
FILTER_DISTINCT_UMI, FILTER_READS = 0, 1
EMIT_MATCHED, EMIT_PASSING = 0, 1


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_whitelist", C.c_int32),
                ("whitelist", C.c_char_p), ("whitelist_len", C.POINTER(C.c_uint8)),
                ("umi_start", C.c_int32), ("umi_end", C.c_int32), ("bc1_start", C.c_int32), ("bc1_end", C.c_int32),
                ("bc2_start", C.c_int32), ("bc2_end", C.c_int32), ("bc3_start", C.c_int32), ("bc3_end", C.c_int32),
                ("errors", C.c_int32), ("min_count", C.c_int64), ("filter_mode", C.c_int32),
                ("collapse_pairs", C.c_int32), ("record_bytes_hint", C.c_int32), ("header_style", C.c_int32),
                ("n_whitelist2", C.c_int32), ("n_whitelist3", C.c_int32),
                ("whitelist2", C.c_char_p), ("whitelist2_len", C.POINTER(C.c_uint8)),
                ("whitelist3", C.c_char_p), ("whitelist3_len", C.POINTER(C.c_uint8))]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_records_r1", "n_records_r2", "n_matched", "n_cells", "n_passing_cells",
                                         "n_retained", "bytes_matched", "bytes_passing", "n_regular_keys",
                                         "n_irregular_keys")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SynthSpec(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("seed", C.c_uint64), ("first_index", C.c_uint64), ("n_pairs", C.c_uint64),
                ("read_len", C.c_int32), ("n_whitelist", C.c_int32), ("whitelist", C.c_char_p),
                ("cell_pool", C.c_uint32), ("p_sub_ppm", C.c_uint32), ("p_n_ppm", C.c_uint32), ("p_junk_ppm", C.c_uint32),
                ("p_dup_umi_ppm", C.c_uint32), ("p_tie_ppm", C.c_uint32), ("p_trunc_ppm", C.c_uint32),
                ("p_atqual_ppm", C.c_uint32)]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ssd_b200.cu into csrc/libssd_b200.so with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libssd_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH + ".tmp", os.path.join(CSRC, "ssd_b200.cu")]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


_lib = None

_SIGS = {
    "ssd_abi_version": (C.c_int, []),
    "ssd_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "ssd_destroy": (None, [C.c_void_p]),
    "ssd_last_error": (C.c_char_p, [C.c_void_p]),
    "ssd_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssd_demux_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ssd_annotated_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ssd_annotated_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "ssd_cell_space": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ssd_reads_histogram_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ssd_distinct_histogram_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ssd_local_unique_keys": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ssd_count_distinct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "ssd_count_distinct_owned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int]),
    "ssd_run_files": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "ssd_annotated_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "ssd_load_file_range": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_uint64, C.c_uint64]),
    "ssd_demux_loaded": (C.c_int, [C.c_void_p]),
    "ssd_annotated_loaded": (C.c_int, [C.c_void_p]),
    "ssd_emit_file": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_size_t)]),
    "ssd_fastq_lines_are_stripped": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "ssd_fastq_record_cuts": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                        C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ssd_key_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssd_scan_repeats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ssd_scan_variant": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "ssd_owner_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssd_partition_keys": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint64)]),
    "ssd_build_filter": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "ssd_filter_single": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "ssd_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "ssd_distinct_local": (C.c_int, [C.c_void_p]),
    "ssd_cell_umi_table": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ssd_anchor_positions": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "ssd_count_newlines_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "ssd_loaded_newlines": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "ssd_pack_decision": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ssd_undecided_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "ssd_adopt_decision": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssd_emit_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ssd_emit_split_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "ssd_split_index": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "ssd_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(Stats)]),
    "ssd_emit_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ssd_output_bound": (C.c_size_t, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "ssd_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ssd_host_free": (C.c_int, [C.c_void_p]),
    "ssd_record_cells": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "ssd_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "ssd_read_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_int]),
    "ssd_launch_count": (C.c_uint64, [C.c_void_p]),
    "ssd_synth_sizes": (C.c_int, [C.POINTER(SynthSpec), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ssd_synth_generate_device": (C.c_int, [C.c_void_p, C.POINTER(SynthSpec), C.c_void_p, C.c_void_p]),
    "ssd_synth_generate_host": (C.c_int, [C.POINTER(SynthSpec), C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGS)


def load():
    """dlopen the in-tree library.  Raises when it has not been built: the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libssd_b200.so is missing (%s). Build it with `python __graft_entry__.py` or "
                               "ssd_b200.build(); there is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.ssd_abi_version() != 1:
            raise RuntimeError("libssd_b200.so ABI version mismatch")
        _lib = lib
    return _lib
