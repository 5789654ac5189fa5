"""Host-side driver of the B200 demultiplexer: owns a library context, moves buffers, maps library
error codes onto the exception types the reference raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib as L

DEFAULT_POSITIONS = {"umi": (0, 10), "bc3": (10, 18), "bc2": (48, 56), "bc1": (86, 94)}  # V2:155-162


class MalformedFastq(ValueError):
    pass


def _raise(code: int, msg: str):
    if code == L.EKEY:
        raise KeyError(msg)                       # V2:350
    if code == L.EINDEX:
        raise IndexError(msg)                     # V2:397-398
    if code == L.EMALFORMED:
        raise MalformedFastq(msg)
    if code == L.EINVAL:
        raise ValueError(msg)
    if code == L.ENOMEM:
        raise MemoryError(msg)
    if code == L.EIO:
        raise (FileNotFoundError if "No such file" in msg else OSError)(msg)
    raise RuntimeError("libssd_b200 error %d: %s" % (code, msg))


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor()."""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        self._owner = owner


class _SplitTable:
    """The bucket index of split mode as a read-only sequence of (cell string, first byte, None, n bytes).  The arrays stay in
    numpy (hundreds of thousands of buckets at full size: the tuples and cell strings are made when somebody asks)."""

    def __init__(self, owner, cells, offs, sizes):
        self._owner, self.cells, self.offsets, self.sizes = owner, cells, offs, sizes

    def __len__(self):
        return int(self.cells.size)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        return (self._owner.cell_string(int(self.cells[k])), int(self.offsets[k]), None, int(self.sizes[k]))

    def __iter__(self):
        return (self[k] for k in range(len(self)))


class Demultiplexer:
    """One library context = one GPU.  Mirrors the knobs of V2.demultiplex_fun / calc_demux_results."""

    def __init__(self, whitelist: Sequence[str], errors: int = 1, min_count: int = 10,
                 positions: Optional[Dict[str, Tuple[int, int]]] = None, filter_mode: str = "distinct_umi",
                 collapse_pairs: int = 0, device: Optional[int] = None, record_bytes_hint: int = 0,
                 whitelist2: Optional[Sequence[str]] = None, whitelist3: Optional[Sequence[str]] = None, header_style: str = "fast"):
        """whitelist: matched against the Round1 window (bc1) -- and against the other two as well (V2:60, 302-304: the reference's
        fast path has ONE list) unless whitelist2 / whitelist3 give the Round2 (bc2) / Round3 (bc3) windows their own lists
        (extension).  header_style: "fast" = V2:134-141 (`id_cell_umi`, blanks of the name removed, cut at the first
        "/"), "extract" = Extract_BC_UMI.py:78-85 (`id_cell_umi comment`; IndexError for names of one token; exact but slower path)."""
        self.lib = L.load()
        pos = dict(DEFAULT_POSITIONS)
        if positions:
            pos.update(positions)
        for k, (a, b) in pos.items():
            if a < 0 or b < a:
                raise ValueError("slice %s=%d:%d is outside the supported contract (0 <= start <= end)" % (k, a, b))
        def as_list(ws):
            return [w if isinstance(w, str) else w.decode() for w in ws]

        def pack(ws):
            return b"".join(w.encode().ljust(8, b"\0")[:8] for w in ws), (C.c_uint8 * len(ws))(*[min(len(w), 8) for w in ws])
        self.whitelist = as_list(whitelist)
        # the list of every round (rounds[k] is matched against window bc<k+1>); the same object when a round has no list of its own
        self.rounds = [self.whitelist, as_list(whitelist2) if whitelist2 is not None else self.whitelist,
                       as_list(whitelist3) if whitelist3 is not None else self.whitelist]
        self.positions = pos
        raw, lens = pack(self.whitelist)
        cfg = L.Config()
        cfg.struct_size = C.sizeof(L.Config)
        cfg.device = -1 if device is None else int(device)
        cfg.n_whitelist = len(self.whitelist)
        cfg.whitelist = raw
        cfg.whitelist_len = lens
        if whitelist2 is not None:
            raw2, lens2 = pack(self.rounds[1])
            cfg.n_whitelist2, cfg.whitelist2, cfg.whitelist2_len = len(self.rounds[1]), raw2, lens2
        if whitelist3 is not None:
            raw3, lens3 = pack(self.rounds[2])
            cfg.n_whitelist3, cfg.whitelist3, cfg.whitelist3_len = len(self.rounds[2]), raw3, lens3
        cfg.umi_start, cfg.umi_end = pos["umi"]
        cfg.bc1_start, cfg.bc1_end = pos["bc1"]
        cfg.bc2_start, cfg.bc2_end = pos["bc2"]
        cfg.bc3_start, cfg.bc3_end = pos["bc3"]
        cfg.errors = int(errors)
        cfg.min_count = int(min_count)
        cfg.filter_mode = {"distinct_umi": L.FILTER_DISTINCT_UMI, "reads": L.FILTER_READS}[filter_mode]
        cfg.collapse_pairs = int(collapse_pairs)
        cfg.record_bytes_hint = int(record_bytes_hint)
        cfg.header_style = {"fast": 0, "extract": 1}[header_style]
        self.filter_mode = filter_mode
        self.min_count = int(min_count)
        self._ctx = C.c_void_p()
        rc = self.lib.ssd_create(C.byref(cfg), C.byref(self._ctx))
        if rc != L.OK:
            self._ctx = C.c_void_p()
            _raise(rc, self.lib.ssd_last_error(None).decode(errors="replace"))
        n = C.c_uint64()
        self.lib.ssd_cell_space(self._ctx, C.byref(n))
        self.n_cells_total = n.value
        ub, cb = C.c_int(), C.c_int()
        self.lib.ssd_key_layout(self._ctx, C.byref(ub), C.byref(cb))
        self.umi_bits, self.cell_bits = ub.value, cb.value
        self._keep = []  # device tensors the current batch reads from

    # -- plumbing ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.ssd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _check(self, rc: int):
        if rc != L.OK:
            _raise(rc, self.lib.ssd_last_error(self._ctx).decode(errors="replace"))

    def _no_records_check(self, st, heads):
        """A call that found no record at all in non-empty input: if that is because the lines end in a lone '\r' (which the
        reference's universal-newline reading takes for line ends and the scans do not), say so instead of returning
        an empty result.  heads() -> the first bytes of the two inputs; only looked at in that case."""
        if st.n_records_r1 == 0 and st.n_records_r2 == 0:
            from . import hostlogic
            hostlogic.reject_lone_cr(*heads())

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.ssd_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def cell_indices(self, cell: int) -> Tuple[int, int, int]:
        """(i1, i2, i3) of a cell id = (i1 * n2 + i2) * n3 + i3."""
        n2, n3 = len(self.rounds[1]), len(self.rounds[2])
        return cell // (n2 * n3), (cell // n3) % n2, cell % n3

    def cell_string(self, cell: int) -> str:
        i1, i2, i3 = self.cell_indices(cell)
        return self.rounds[0][i1] + self.rounds[1][i2] + self.rounds[2][i3]

    # -- phases on device-resident inputs ---------------------------------------------------------
    def demux_device(self, r1_ptr: int, r1_len: int, r2_ptr: int, r2_len: int, keep=()):
        """Phase 1 (V2:207-358).  `keep` holds whatever owns the device memory, until the next batch."""
        self._keep = list(keep)
        self._check(self.lib.ssd_demux_device(self._ctx, C.c_void_p(r1_ptr), r1_len, C.c_void_p(r2_ptr), r2_len))

    def reads_histogram(self):
        import torch
        p = C.c_void_p()
        self._check(self.lib.ssd_reads_histogram_device(self._ctx, C.byref(p)))
        return torch.as_tensor(_DevArray(p.value, self.n_cells_total, "<i4", self), device="cuda")

    def distinct_histogram(self):
        import torch
        p = C.c_void_p()
        self._check(self.lib.ssd_distinct_histogram_device(self._ctx, C.byref(p)))
        return torch.as_tensor(_DevArray(p.value, self.n_cells_total, "<i4", self), device="cuda")

    def local_unique_keys(self):
        """Sorted unique (cell,UMI) keys of this batch as torch views: int64[n] and int64[n_irr, 2]."""
        import torch
        pk, pi, nk, ni = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.ssd_local_unique_keys(self._ctx, C.byref(pk), C.byref(nk), C.byref(pi), C.byref(ni)))
        keys = torch.as_tensor(_DevArray(pk.value, nk.value, "<i8", self), device="cuda") if nk.value else torch.empty(0, dtype=torch.int64, device="cuda")
        irr = (torch.as_tensor(_DevArray(pi.value, 2 * ni.value, "<i8", self), device="cuda").view(-1, 2)
               if ni.value else torch.empty((0, 2), dtype=torch.int64, device="cuda"))
        return keys, irr

    def owner_layout(self):
        """(shift, n_buckets): the rank that owns a cell in a world of W ranks is ((cell >> shift) * W) // n_buckets."""
        sh, nb = C.c_int(), C.c_int()
        self._check(self.lib.ssd_owner_layout(self._ctx, C.byref(sh), C.byref(nb)))
        return sh.value, nb.value

    def partition_keys(self, world: int):
        """This batch's (cell,UMI) keys with duplicates, as torch views: int64[n] grouped by owner rank, the world + 1
        offsets of the groups (list), and the irregular keys int64[n_irr, 2] in no particular order."""
        import torch
        pk, pi, ni = C.c_void_p(), C.c_void_p(), C.c_uint64()
        offs = (C.c_uint64 * (world + 1))()
        self._check(self.lib.ssd_partition_keys(self._ctx, world, C.byref(pk), offs, C.byref(pi), C.byref(ni)))
        offsets = list(offs)
        nk = offsets[-1]
        keys = torch.as_tensor(_DevArray(pk.value, nk, "<i8", self), device="cuda") if nk else torch.empty(0, dtype=torch.int64, device="cuda")
        irr = (torch.as_tensor(_DevArray(pi.value, 2 * ni.value, "<i8", self), device="cuda").view(-1, 2)
               if ni.value else torch.empty((0, 2), dtype=torch.int64, device="cuda"))
        return keys, offsets, irr

    def count_distinct(self, keys, irr):
        """Per-cell distinct counts from int64 key tensors (any order, duplicates allowed; reordered in place)."""
        self._check(self.lib.ssd_count_distinct(self._ctx, C.c_void_p(keys.data_ptr() if keys.numel() else 0), keys.numel(),
                                                C.c_void_p(irr.data_ptr() if irr.numel() else 0), irr.numel() // 2))

    def count_distinct_owned(self, keys, irr, rank: int, world: int):
        """count_distinct for rank `rank` of `world`: keys = what the peers sent for this rank's cells, irr = everybody's
        irregular keys (all-gathered; rows of -1 are padding), of which only this rank's cells are counted."""
        self._check(self.lib.ssd_count_distinct_owned(self._ctx, C.c_void_p(keys.data_ptr() if keys.numel() else 0), keys.numel(),
                                                      C.c_void_p(irr.data_ptr() if irr.numel() else 0), irr.numel() // 2, rank, world))

    def distinct_local(self):
        """Per-cell distinct-UMI counts of THIS batch into the distinct table (the read histogram must still be the batch's own)."""
        self._check(self.lib.ssd_distinct_local(self._ctx))

    def pack_decision(self):
        """One int64 per cell (CUDA tensor over the cell space, owned by the context): bits 0..39 = reads of the cell in this batch,
        bit 40 = this batch alone counts >= m distinct UMIs (needs distinct_local first).  All-reduce it (SUM) in place."""
        import torch
        p = C.c_void_p()
        self._check(self.lib.ssd_pack_decision(self._ctx, C.byref(p)))
        return torch.as_tensor(_DevArray(p.value, self.n_cells_total, "<i8", self), device="cuda")

    def undecided_keys(self, packed_global, capacity: int):
        """This batch's (cell, UMI) keys (duplicates kept) of the cells no rank decides alone (all-reduced pack_decision words: no
        rank passes it, global reads >= m), as one packed list: int64[capacity + 1, 2] on the device, row 0 = (-1, number of keys
        the batch had), then the keys in the irregular-key format, unused rows all ones.  Nothing is read back."""
        import torch
        p = C.c_void_p()
        self._check(self.lib.ssd_undecided_keys(self._ctx, C.c_void_p(packed_global.data_ptr()), int(capacity), C.byref(p)))
        return torch.as_tensor(_DevArray(p.value, 2 * (int(capacity) + 1), "<i8", self), device="cuda").view(-1, 2)

    def adopt_decision(self, packed_global):
        """After count_distinct on the gathered lists: read histogram := global reads, distinct := max(distinct, m) where some rank
        decided "pass" alone."""
        self._check(self.lib.ssd_adopt_decision(self._ctx, C.c_void_p(packed_global.data_ptr())))

    def build_filter(self) -> dict:
        st = L.Stats()
        self._check(self.lib.ssd_build_filter(self._ctx, C.byref(st)))
        return st.as_dict()

    def filter_single(self) -> dict:
        st = L.Stats()
        self._check(self.lib.ssd_filter_single(self._ctx, C.byref(st)))
        return st.as_dict()

    def stats(self) -> dict:
        """The statistics of the current batch as the last filter left them."""
        st = L.Stats()
        self._check(self.lib.ssd_get_stats(self._ctx, C.byref(st)))
        return st.as_dict()

    def emit_device(self, which: int, out_ptr: int, cap: int) -> int:
        n = C.c_size_t()
        self._check(self.lib.ssd_emit_device(self._ctx, which, C.c_void_p(out_ptr), cap, C.byref(n)))
        return n.value

    def emit_split_device(self, out_ptr: int, cap: int):
        """Per-cell buckets of the post-filter records into the device buffer; returns (bytes, table) where table is
        a list of (cell string, first byte, n records, n bytes)."""
        import numpy as np
        n, nb = C.c_size_t(), C.c_uint64()
        self._check(self.lib.ssd_emit_split_device(self._ctx, C.c_void_p(out_ptr), cap, C.byref(n), C.byref(nb)))
        cells = np.empty(nb.value, dtype=np.uint32)
        offs = np.empty(nb.value, dtype=np.uint64)
        first = np.empty(nb.value, dtype=np.uint64)
        if nb.value:
            self._check(self.lib.ssd_split_index(self._ctx, cells.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                                                 first.ctypes.data_as(C.c_void_p), nb.value))
        self._split_ids = cells  # cell ids of the buckets, in table order (numpy)
        sizes = np.diff(np.append(offs, np.uint64(n.value))).astype(np.uint64) if nb.value else offs
        return n.value, _SplitTable(self, cells, offs, sizes), first

    def write_split_files(self, out_dir: str, round_names=None, stats: Optional[dict] = None) -> List[Tuple[str, int]]:
        """Split mode on disk: after a run, one FASTQ file per passing cell holding that cell's annotated
        records in input order, named like the legacy pipeline's per-cell files, round1-round2-round3.fastq with the
        full-length barcodes of the three Round files (demultiplex_using_barcodes.py:364; round_names =
        hostlogic.load_round_names(dir)), or with the 8-mers when no names are given.  Files are opened in append mode like
        splitseq_utilities.flushBuffers (utilities:15-21).  `stats` = what the run returned (else the filter is rebuilt).
        Returns [(file name, bytes)] in cell order."""
        import torch
        from . import hostlogic
        stats = stats or self.build_filter()
        out = torch.empty(int(stats["bytes_passing"]) + 16, dtype=torch.uint8, device="cuda")
        n, table, _ = self.emit_split_device(out.data_ptr(), out.numel())
        host = out[:n].cpu().numpy()
        del out
        os.makedirs(out_dir, exist_ok=True)
        jobs = []
        for off, nbytes, cid in zip(table.offsets.tolist(), table.sizes.tolist(), table.cells.tolist()):
            i1, i2, i3 = self.cell_indices(cid)
            name = (hostlogic.split_file_name(round_names, i1, i2, i3) if round_names is not None
                    else "%s-%s-%s.fastq" % (self.rounds[0][i1], self.rounds[1][i2], self.rounds[2][i3]))
            jobs.append((name, off, nbytes))

        def put(job):  # one open / write / close per cell: system calls, which release the interpreter lock
            name, off, nbytes = job
            with open(os.path.join(out_dir, name), "ab") as fh:
                fh.write(memoryview(host)[off:off + nbytes])
            return name, nbytes
        if len(jobs) < 64:
            return [put(j) for j in jobs]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:  # every cell has its own file: no two threads share one
            return list(pool.map(put, jobs, chunksize=256))

    def count_newlines(self, text) -> int:
        """Number of line feeds in a uint8 CUDA tensor (16-byte aligned): the framing check of an output, at memory speed."""
        n = C.c_uint64()
        self._check(self.lib.ssd_count_newlines_device(self._ctx, C.c_void_p(text.data_ptr() if text.numel() else 0), text.numel(), C.byref(n)))
        return n.value

    def record_cells(self):
        import numpy as np
        n = C.c_uint64()
        self._check(self.lib.ssd_record_cells(self._ctx, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.int32)
        if n.value:
            self._check(self.lib.ssd_record_cells(self._ctx, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    # -- instrumentation -------------------------------------------------------------------------
    STAGES = ("scan_r1", "scan_r2", "distinct", "filter", "plan", "emit")

    def set_timing(self, on: bool):
        self._check(self.lib.ssd_set_timing(self._ctx, 1 if on else 0))

    def read_timings(self) -> Dict[str, Tuple[float, int]]:
        ms = (C.c_double * len(self.STAGES))()
        cnt = (C.c_uint32 * len(self.STAGES))()
        self._check(self.lib.ssd_read_timings(self._ctx, ms, cnt, len(self.STAGES)))
        return {s: (ms[i], cnt[i]) for i, s in enumerate(self.STAGES)}

    def scan_repeats(self) -> int:
        """How many times a scan was repeated without guessing the record phase (diagnostics)."""
        n = C.c_uint64()
        self._check(self.lib.ssd_scan_repeats(self._ctx, C.byref(n)))
        return n.value

    def scan_variant(self) -> Tuple[bool, int]:
        """(short_fast, exact_records): whether the NEXT batch's read-2 scan keeps short sequences and windows with several
        non-ACGT bytes on its fast path (the context turns that on after a batch that sent more than 1 record in 1 000 down the
        exact record-by-record path; SSD_SCAN_SHORT=0/1 pins it), and how many read-2 records of the LAST batch went that way."""
        sf, n = C.c_int(), C.c_uint64()
        self._check(self.lib.ssd_scan_variant(self._ctx, C.byref(sf), C.byref(n)))
        return bool(sf.value), n.value

    def launch_count(self) -> int:
        return int(self.lib.ssd_launch_count(self._ctx))

    # -- whole pipeline ---------------------------------------------------------------------------
    def run_device(self, r1, r2, which: int = L.EMIT_PASSING):
        """r1/r2: torch uint8 CUDA tensors holding the two files.  Returns (uint8 CUDA tensor, stats)."""
        import torch
        handle = torch.cuda.current_stream().cuda_stream
        if handle == 0:
            torch.cuda.current_stream().synchronize()  # the library then runs on its own stream
        self.set_stream(handle)
        self.demux_device(r1.data_ptr(), r1.numel(), r2.data_ptr(), r2.numel(), keep=(r1, r2))
        stats = self.filter_single()
        need = stats["bytes_matched"] if which == L.EMIT_MATCHED else stats["bytes_passing"]
        out = torch.empty(need + 16, dtype=torch.uint8, device=r1.device)
        n = self.emit_device(which, out.data_ptr(), out.numel())
        return out[:n], self.stats()

    def run_host(self, r1, r2, which: int = L.EMIT_PASSING, out=None):
        """r1/r2: objects exposing the buffer protocol (bytes, numpy, pinned torch .numpy()).
        Returns (bytes-like numpy uint8 array, stats).  H2D and D2H copies happen inside the library."""
        import numpy as np
        a1 = np.frombuffer(r1, dtype=np.uint8) if not isinstance(r1, np.ndarray) else r1
        a2 = np.frombuffer(r2, dtype=np.uint8) if not isinstance(r2, np.ndarray) else r2
        st = L.Stats()
        n = C.c_size_t()
        args = (self._ctx, a1.ctypes.data_as(C.c_void_p), a1.size, a2.ctypes.data_as(C.c_void_p), a2.size, which)
        if out is not None and out.size:
            # the caller's buffer (pinned, reused): everything in one call
            rc = self.lib.ssd_run_host(*args, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n), C.byref(st))
            if rc != L.ERANGE:
                self._check(rc)
                self._no_records_check(st, lambda: (bytes(a1[:65536]), bytes(a2[:65536])))
                return out[: n.value], self.stats()
        else:
            self._check(self.lib.ssd_run_host(*args, None, 0, C.byref(n), C.byref(st)))
        self._no_records_check(st, lambda: (bytes(a1[:65536]), bytes(a2[:65536])))
        out = np.empty(n.value, dtype=np.uint8)
        if n.value:
            self._check(self.lib.ssd_emit_host(self._ctx, which, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out[: n.value], self.stats()

    def run_files(self, path_r1: str, path_r2: str, out_path: str, which: int = L.EMIT_PASSING, append: bool = True):
        """File to file (what V2.demultiplex_fun does with one pair of chunk files): reads, uploads, demultiplexes and
        appends (or writes) the selected output to out_path, all inside the library with pinned rings and overlapped
        copies.  Returns (bytes written, stats).  Files above SSD_BATCH_BYTES (16 GiB) go through the GPU in batches of whole
        records (ssd_b200.filebatch) with the same bytes as result."""
        from . import filebatch
        if filebatch.needs_batches(path_r1, path_r2):
            return filebatch.run_files_batched(self, path_r1, path_r2, out_path, which, append)
        st = L.Stats()
        n = C.c_size_t()
        self._check(self.lib.ssd_run_files(self._ctx, os.fsencode(path_r1), os.fsencode(path_r2), which, os.fsencode(out_path), 1 if append else 0,
                                           C.byref(n), C.byref(st)))

        def heads():
            with open(path_r1, "rb") as f1, open(path_r2, "rb") as f2:
                return f1.read(65536), f2.read(65536)
        self._no_records_check(st, heads)
        return n.value, st.as_dict()

    def filter_annotated_file(self, path_in: str, out_path: str, append: bool = True):
        """V2.calc_demux_results file to file: returns (bytes written, stats).  Files above SSD_BATCH_BYTES: two passes over
        batches (ssd_b200.filebatch)."""
        from . import filebatch
        if filebatch.needs_batches(path_in):
            return filebatch.filter_annotated_file_batched(self, path_in, out_path, append)
        st = L.Stats()
        n = C.c_size_t()
        self._check(self.lib.ssd_annotated_file(self._ctx, os.fsencode(path_in), os.fsencode(out_path), 1 if append else 0, C.byref(n), C.byref(st)))
        return n.value, st.as_dict()

    def filter_annotated_host(self, text):
        """V2.calc_demux_results on the text of MergedCells_1.fastq (any buffer-protocol object): returns
        (uint8 array with the records of passing cells, stats)."""
        import numpy as np
        a = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        st = L.Stats()
        n = C.c_size_t()
        self._check(self.lib.ssd_annotated_host(self._ctx, a.ctypes.data_as(C.c_void_p), a.size, None, 0, C.byref(n), C.byref(st)))
        out = np.empty(n.value, dtype=np.uint8)
        if n.value:
            self._check(self.lib.ssd_emit_host(self._ctx, L.EMIT_PASSING, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out[: n.value], st.as_dict()

    def write_cell_table(self, path: str) -> int:
        """The per-cell (reads, distinct UMIs) table of the last run as tab-separated text (hostlogic.format_cell_table);
        returns the number of cells."""
        from . import hostlogic
        table = self.cell_table()
        with open(path, "w") as fh:
            fh.write(hostlogic.format_cell_table(table, self.min_count))
        return len(table)

    def cell_umi_table(self):
        """Per-cell x UMI table of the last batch: numpy arrays (cells uint32[n], umis bytes-array S10[n], reads
        uint32[n]) with one row per distinct (cell, UMI) pair and the number of reads it was seen in, sorted by cell id and,
        inside a cell, regular UMIs (by 2-bit code) before irregular ones.  Sort + run-length count on the device
        (ssd_cell_umi_table); only the unique rows come back."""
        import numpy as np
        import torch
        pk, pr, pi, pir = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nk, ni = C.c_uint64(), C.c_uint64()
        self._check(self.lib.ssd_cell_umi_table(self._ctx, C.byref(pk), C.byref(pr), C.byref(nk), C.byref(pi), C.byref(pir), C.byref(ni)))
        nk, ni = nk.value, ni.value

        def host(ptr, n, typestr):
            return torch.as_tensor(_DevArray(ptr, n, typestr, self), device="cuda").cpu().numpy() if n else np.zeros(0, dtype=np.dtype(typestr))

        keys, kreads = host(pk.value, nk, "<i8").astype(np.uint64), host(pr.value, nk, "<i4").astype(np.uint32)
        irr, ireads = host(pi.value, 2 * ni, "<i8").astype(np.uint64).reshape(-1, 2), host(pir.value, ni, "<i4").astype(np.uint32)
        ulen = self.umi_bits // 2
        cells_k = (keys >> np.uint64(self.umi_bits)).astype(np.uint32)
        umis_k = np.zeros((nk, 10), dtype=np.uint8)
        lut = np.frombuffer(b"ACTG", dtype=np.uint8)  # 2-bit code -> base (csrc/ssd_scan.cuh pack4)
        for j in range(ulen):
            umis_k[:, j] = lut[((keys >> np.uint64(2 * j)) & np.uint64(3)).astype(np.int64)]
        cells_i = ((irr[:, 0] >> np.uint64(36)) & np.uint64((1 << 21) - 1)).astype(np.uint32)
        umis_i = np.zeros((ni, 10), dtype=np.uint8)
        for j in range(10):  # bytes 0..3 in hi (bits 24, 16, 8, 0), bytes 4..9 in lo (bits 40 .. 0); unused bytes are zero
            word, shift = (irr[:, 0], 8 * (3 - j)) if j < 4 else (irr[:, 1], 8 * (9 - j))
            umis_i[:, j] = ((word >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.uint8)
        cells = np.concatenate([cells_k, cells_i])
        umis = np.concatenate([umis_k, umis_i]).view("S10").reshape(-1)
        reads = np.concatenate([kreads, ireads])
        order = np.argsort(cells, kind="stable")  # both halves are sorted by cell already: regular rows stay ahead of irregular ones
        return cells[order], umis[order], reads[order]

    def write_cell_umi_table(self, path: str, passing_only: bool = True) -> int:
        """cell <tab> UMI <tab> reads, one line per distinct (cell, UMI) pair of the last batch (of the passing cells only by
        default: what a per-cell UMI counter downstream consumes); returns the number of rows."""
        import numpy as np
        from . import hostlogic
        cells, umis, reads = self.cell_umi_table()
        if passing_only:
            thr = self.min_count
            ok = (self.reads_histogram() if self.filter_mode == "reads" else self.distinct_histogram()).cpu().numpy() >= thr
            keep = ok[cells]
            cells, umis, reads = cells[keep], umis[keep], reads[keep]
        with open(path, "wb") as fh:
            fh.write(hostlogic.format_cell_umi_table(self.rounds, cells, umis, reads))
        return int(cells.size)

    def cell_table(self) -> Dict[str, Tuple[int, int]]:
        """{cell string: (reads, distinct UMIs)} for every detected cell (small outputs / tests)."""
        hist = self.reads_histogram().cpu().numpy()
        dist = self.distinct_histogram().cpu().numpy()
        import numpy as np
        return {self.cell_string(int(c)): (int(hist[c]), int(dist[c])) for c in np.nonzero(hist)[0]}


class HostPipeline:
    """Several contexts on one GPU fed from host threads: while one batch uploads, the previous one computes and
    downloads (PCIe is full duplex; the library hands each direction to one context at a time).  The reference runs
    its chunks through joblib workers the same way (CLI3:56-64).  `map` keeps the order of the batches."""

    def __init__(self, whitelist: Sequence[str], depth: int = 2, **config):
        self.contexts = [Demultiplexer(whitelist, **config) for _ in range(max(1, depth))]

    def map(self, batches, which: int = L.EMIT_PASSING, outs=None):
        """batches: sequence of (read-1 bytes-like, read-2 bytes-like) host buffers (pinned for full PCIe speed).
        outs: optional list of one reusable host output array per context.  Returns [(output array, stats), ...]."""
        import threading
        batches = list(batches)
        results = [None] * len(batches)
        errors = []
        nxt = {"i": 0}
        lock = threading.Lock()

        def work(k):
            dm = self.contexts[k]  # the library selects the context's device on every call
            while True:
                with lock:
                    i = nxt["i"]
                    if i >= len(batches) or errors:
                        return
                    nxt["i"] = i + 1
                try:
                    results[i] = dm.run_host(batches[i][0], batches[i][1], which, out=None if outs is None else outs[k])
                except Exception as e:  # noqa: BLE001 - re-raised below, in the caller's thread
                    errors.append(e)
                    return

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(self.contexts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def close(self):
        for dm in self.contexts:
            dm.close()


def synth_spec(whitelist: Sequence[str], n_pairs: int, seed: int = 0x5EED0001, first_index: int = 1, read_len: int = 150,
               cell_pool: int = 100000, p_sub: float = 0.01, p_n: float = 0.002, p_junk: float = 0.10, p_dup_umi: float = 0.05,
               p_tie: float = 0.0, p_trunc: float = 0.0, p_atqual: float = 0.02):
    """ssd_synth_spec for the synthetic read pairs of SURVEY.md 8d."""
    raw = b"".join(w.encode()[:8] for w in whitelist)
    s = L.SynthSpec()
    s.struct_size = C.sizeof(L.SynthSpec)
    s.seed, s.first_index, s.n_pairs, s.read_len = seed, first_index, n_pairs, read_len
    s.n_whitelist, s.whitelist, s.cell_pool = len(whitelist), raw, cell_pool
    ppm = lambda p: int(round(p * 1e6))  # noqa: E731
    s.p_sub_ppm, s.p_n_ppm, s.p_junk_ppm, s.p_dup_umi_ppm = ppm(p_sub), ppm(p_n), ppm(p_junk), ppm(p_dup_umi)
    s.p_tie_ppm, s.p_trunc_ppm, s.p_atqual_ppm = ppm(p_tie), ppm(p_trunc), ppm(p_atqual)
    s._keep = raw
    return s


def synth_sizes(spec) -> Tuple[int, int]:
    a, b = C.c_uint64(), C.c_uint64()
    rc = L.load().ssd_synth_sizes(C.byref(spec), C.byref(a), C.byref(b))
    if rc != L.OK:
        raise ValueError("bad synthetic spec")
    return a.value, b.value


def synth_host(spec) -> Tuple[bytes, bytes]:
    """CPU twin of the device generator (same bytes)."""
    n1, n2 = synth_sizes(spec)
    b1, b2 = C.create_string_buffer(n1 + 1), C.create_string_buffer(n2 + 1)
    rc = L.load().ssd_synth_generate_host(C.byref(spec), b1, b2)
    if rc != L.OK:
        raise ValueError("bad synthetic spec")
    return b1.raw[:n1], b2.raw[:n2]


def synth_device(dm: Demultiplexer, spec, into=None):
    """Generate on dm's GPU; returns two uint8 CUDA tensors.  into = (buffer 1, buffer 2): uint8 CUDA tensors of at least
    synth_sizes(spec) + 16 bytes to generate into (batch after batch into the same memory), else two new tensors."""
    import torch
    n1, n2 = synth_sizes(spec)
    if into is not None:
        t1, t2 = into
        if t1.numel() < n1 + 16 or t2.numel() < n2 + 16:
            raise ValueError("synth_device: the buffers are too small for this batch")
    else:
        t1 = torch.empty(n1 + 16, dtype=torch.uint8, device="cuda")
        t2 = torch.empty(n2 + 16, dtype=torch.uint8, device="cuda")
    rc = dm.lib.ssd_synth_generate_device(dm._ctx, C.byref(spec), C.c_void_p(t1.data_ptr()), C.c_void_p(t2.data_ptr()))
    dm._check(rc)
    return t1[:n1], t2[:n2]
