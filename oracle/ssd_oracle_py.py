"""ctypes binding of the CPU oracle (oracle/libssd_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libssd_oracle.so")

OK, EKEY, EMALFORMED, EINDEX, ENOMEM = 0, -1, -2, -3, -4


class _Cfg(C.Structure):
    _fields_ = [("whitelist", C.c_char_p), ("wl_len", C.POINTER(C.c_uint8)), ("n_whitelist", C.c_int),
                ("umi_start", C.c_int), ("umi_end", C.c_int), ("bc1_start", C.c_int), ("bc1_end", C.c_int),
                ("bc2_start", C.c_int), ("bc2_end", C.c_int), ("bc3_start", C.c_int), ("bc3_end", C.c_int),
                ("errors", C.c_int), ("bin_reads", C.c_long),
                ("whitelist2", C.c_char_p), ("wl_len2", C.POINTER(C.c_uint8)), ("n_whitelist2", C.c_int),
                ("whitelist3", C.c_char_p), ("wl_len3", C.POINTER(C.c_uint8)), ("n_whitelist3", C.c_int)]


class _Cell(C.Structure):
    _fields_ = [("cell", C.c_char * 64), ("reads", C.c_uint64), ("distinct_umis", C.c_uint64), ("passing", C.c_int)]


class _Out(C.Structure):
    _fields_ = [("merged1", C.POINTER(C.c_char)), ("merged1_len", C.c_size_t),
                ("passing", C.POINTER(C.c_char)), ("passing_len", C.c_size_t),
                ("rec_cell", C.POINTER(C.c_int32)), ("n_records", C.c_uint64),
                ("n_matched", C.c_uint64), ("n_retained", C.c_uint64),
                ("cells", C.POINTER(_Cell)), ("n_cells", C.c_uint64), ("n_passing_cells", C.c_uint64),
                ("error", C.c_int), ("error_msg", C.c_char * 256)]


class _SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_index", C.c_uint64), ("n_pairs", C.c_uint64), ("read_len", C.c_int32),
                ("n_whitelist", C.c_int32), ("whitelist", C.c_char_p), ("cell_pool", C.c_uint32), ("p_sub_ppm", C.c_uint32),
                ("p_n_ppm", C.c_uint32), ("p_junk_ppm", C.c_uint32), ("p_dup_umi_ppm", C.c_uint32), ("p_tie_ppm", C.c_uint32),
                ("p_trunc_ppm", C.c_uint32), ("p_atqual_ppm", C.c_uint32)]


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, "ssd_oracle.c"), os.path.join(HERE, "ssd_oracle.h"), os.path.join(HERE, "ssd_synth_ref.c")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libssd_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ssd_oracle_demultiplex.argtypes = [C.POINTER(_Cfg), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(_Out)]
        _lib.ssd_oracle_calc_demux_results.argtypes = [C.POINTER(_Out), C.c_long]
        _lib.ssd_oracle_run_cli.argtypes = [C.POINTER(_Cfg), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_long, C.c_int, C.c_long, C.POINTER(_Out)]
        _lib.ssd_oracle_learn_positions.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        _lib.ssd_oracle_collapse_map.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        _lib.ssd_oracle_collapse_map.restype = None
        _lib.ssd_oracle_free.argtypes = [C.POINTER(_Out)]
        _lib.ssd_oracle_free.restype = None
        _lib.ssd_oracle_synth_sizes.argtypes = [C.POINTER(_SynthSpec), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib.ssd_oracle_synth_generate.argtypes = [C.POINTER(_SynthSpec), C.c_void_p, C.c_void_p, C.c_int]
    return _lib


class OracleError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__("%d: %s" % (code, msg))
        self.code = code
        self.msg = msg


@dataclass
class OracleResult:
    merged1: bytes = b""
    passing: bytes = b""
    rec_cell: List[int] = field(default_factory=list)
    n_records: int = 0
    n_matched: int = 0
    n_retained: int = 0
    cells: Dict[bytes, Tuple[int, int, bool]] = field(default_factory=dict)  # cell -> (reads, distinct, passing)
    n_cells: int = 0
    n_passing_cells: int = 0


DEFAULT_POSITIONS = dict(umi=(0, 10), bc3=(10, 18), bc2=(48, 56), bc1=(86, 94))  # V2:155-162


def _cfg(whitelist: List[str], e: int, bin_reads: int, positions: Optional[dict], whitelist2=None, whitelist3=None):
    pos = dict(DEFAULT_POSITIONS)
    if positions:
        pos.update(positions)

    def pack(ws):
        return b"".join(w.encode().ljust(8, b"\0")[:8] for w in ws), (C.c_uint8 * len(ws))(*[min(len(w), 8) for w in ws])
    raw, lens = pack(whitelist)
    cfg = _Cfg(raw, lens, len(whitelist), pos["umi"][0], pos["umi"][1], pos["bc1"][0], pos["bc1"][1],
               pos["bc2"][0], pos["bc2"][1], pos["bc3"][0], pos["bc3"][1], e, bin_reads)
    cfg._keep = [raw, lens]
    if whitelist2 is not None:  # extension: the Round2 / Round3 windows against their own lists
        raw2, lens2 = pack(whitelist2)
        cfg.whitelist2, cfg.wl_len2, cfg.n_whitelist2 = raw2, lens2, len(whitelist2)
        cfg._keep += [raw2, lens2]
    if whitelist3 is not None:
        raw3, lens3 = pack(whitelist3)
        cfg.whitelist3, cfg.wl_len3, cfg.n_whitelist3 = raw3, lens3, len(whitelist3)
        cfg._keep += [raw3, lens3]
    return cfg


def _bytes_at(ptr, n: int) -> bytes:
    """n bytes at a C pointer (ctypes.string_at stops at 2 GiB; the full-size comparisons are larger)."""
    if not n:
        return b""
    return bytes((C.c_char * n).from_address(C.cast(ptr, C.c_void_p).value))


def _collect(out: _Out, want_rec_cell: bool) -> OracleResult:
    res = OracleResult()
    res.merged1 = _bytes_at(out.merged1, out.merged1_len)
    res.passing = _bytes_at(out.passing, out.passing_len)
    res.n_records, res.n_matched, res.n_retained = out.n_records, out.n_matched, out.n_retained
    res.n_cells, res.n_passing_cells = out.n_cells, out.n_passing_cells
    if want_rec_cell and out.n_records:
        res.rec_cell = list(out.rec_cell[: out.n_records])
    for i in range(out.n_cells):
        c = out.cells[i]
        res.cells[c.cell] = (c.reads, c.distinct_umis, bool(c.passing))
    return res


def demultiplex(r1: bytes, r2: bytes, whitelist: List[str], e: int, m: int, *, bin_reads: int = 100000,
                positions: Optional[dict] = None, want_rec_cell: bool = True, whitelist2=None, whitelist3=None) -> OracleResult:
    """demultiplex_fun followed by calc_demux_results (V2:32-462).  whitelist2 / whitelist3: the per-round extension."""
    L = lib()
    cfg = _cfg(whitelist, e, bin_reads, positions, whitelist2, whitelist3)
    out = _Out()
    try:
        rc = L.ssd_oracle_demultiplex(C.byref(cfg), r1, len(r1), r2, len(r2), C.byref(out))
        if rc == OK:
            rc = L.ssd_oracle_calc_demux_results(C.byref(out), m)
        if rc != OK:
            raise OracleError(rc, out.error_msg.decode(errors="replace"))
        return _collect(out, want_rec_cell)
    finally:
        L.ssd_oracle_free(C.byref(out))


def run_cli(r1: bytes, r2: bytes, whitelist: List[str], e: int, m: int, *, n_workers: int, length_fastq: int,
            bin_reads: int = 100000, positions: Optional[dict] = None) -> OracleResult:
    """splitseqdemultiplex_0.2.3.py:52-70 with OpenMP threads as the joblib workers."""
    L = lib()
    cfg = _cfg(whitelist, e, bin_reads, positions)
    out = _Out()
    try:
        rc = L.ssd_oracle_run_cli(C.byref(cfg), r1, len(r1), r2, len(r2), length_fastq, n_workers, m, C.byref(out))
        if rc != OK:
            raise OracleError(rc, out.error_msg.decode(errors="replace"))
        return _collect(out, False)
    finally:
        L.ssd_oracle_free(C.byref(out))


def learn_positions(sample: bytes) -> Tuple[int, int, int]:
    pos = (C.c_int * 3)()
    rc = lib().ssd_oracle_learn_positions(sample, len(sample), pos)
    if rc != OK:
        raise OracleError(rc, "ValueError: max() arg is an empty sequence")
    return pos[0], pos[1], pos[2]


def collapse_map(present: bytes, n: int, n_pairs: int) -> List[int]:
    remap = (C.c_uint32 * (n * n * n))()
    lib().ssd_oracle_collapse_map(present, n, n_pairs, remap)
    return list(remap)


def synth_pairs(whitelist: List[str], n_pairs: int, seed: int = 0x5EED0001, first_index: int = 1, read_len: int = 150,
                cell_pool: int = 100000, p_sub: float = 0.01, p_n: float = 0.002, p_junk: float = 0.10, p_dup_umi: float = 0.05,
                p_tie: float = 0.0, p_trunc: float = 0.0, p_atqual: float = 0.02, n_threads: Optional[int] = None):
    """The synthetic read pairs of SURVEY 8d (same arguments and same bytes as ssd_b200.synth_spec + synth_host / synth_device)
    made by oracle/ssd_synth_ref.c, without the product library.  Returns two numpy uint8 arrays (read 1, read 2)."""
    import numpy as np
    raw = b"".join(w.encode()[:8] for w in whitelist)
    ppm = lambda p: int(round(p * 1e6))  # noqa: E731
    spec = _SynthSpec(seed, first_index, n_pairs, read_len, len(whitelist), raw, cell_pool, ppm(p_sub), ppm(p_n), ppm(p_junk),
                      ppm(p_dup_umi), ppm(p_tie), ppm(p_trunc), ppm(p_atqual))
    a, b = C.c_uint64(), C.c_uint64()
    if lib().ssd_oracle_synth_sizes(C.byref(spec), C.byref(a), C.byref(b)) != 0:
        raise ValueError("bad synthetic spec")
    r1, r2 = np.empty(a.value, dtype=np.uint8), np.empty(b.value, dtype=np.uint8)
    if lib().ssd_oracle_synth_generate(C.byref(spec), r1.ctypes.data_as(C.c_void_p), r2.ctypes.data_as(C.c_void_p),
                                       n_threads or os.cpu_count() or 1) != 0:
        raise ValueError("bad synthetic spec")
    return r1, r2
