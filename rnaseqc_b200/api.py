"""ctypes binding of the C-ABI in include/rsq_b200.h plus the host helper library (librsq_host.so).

`Backend(path)` loads ONE library implementing the ABI.  The product is `librsq_b200.so` (CUDA, sm_100a); tests
may also load the CPU oracle `oracle/_build/librsq_oracle.so` through the same class to compare results — the
product code never does (PRODUCT_LIB below is the only path `default_backend()` will load, and it raises if the
library is missing or no GPU is present).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.environ.get("RSQ_B200_LIB") or os.path.join(ROOT, "rnaseqc_b200", "lib", "librsq_b200.so")  # env override: A/B builds of the CUDA library
HOST_LIB = os.path.join(ROOT, "rnaseqc_b200", "lib", "librsq_host.so")

RSQ_N_COUNTERS = 64
RSQ_OK, RSQ_ERR_INVALID, RSQ_ERR_NO_DEVICE, RSQ_ERR_CUDA, RSQ_ERR_NOMEM, RSQ_ERR_CAPACITY = range(6)

COUNTER_NAMES = [
    "Alternative Alignments", "Supplementary Alignments", "Failed Vendor QC", "Low Mapping Quality",
    "Chimeric Fragments_auto", "Chimeric Fragments_tag", "Unique Mapping, Vendor QC Passed Reads", "Unpaired Reads",
    "Mapped Reads", "Mapped Duplicate Reads", "Mapped Unique Reads", "Total Mapped Pairs", "End 1 Mapped Reads",
    "End 1 Mismatches", "End 1 Bases", "Duplicate Pairs", "Unique Fragments", "End 2 Mapped Reads", "End 2 Mismatches",
    "End 2 Bases", "Mismatched Bases", "Total Bases", "High Quality Reads", "Low Quality Reads",
    "Reads used for Intron/Exon counts", "Alignment Blocks", "Non-Globin Reads", "Non-Globin Duplicate Reads",
    "Intronic Reads", "Intragenic Reads", "HQ Intronic Reads", "HQ Intragenic Reads", "Intergenic Reads",
    "HQ Intergenic Reads", "Exonic Reads", "HQ Exonic Reads", "Ambiguous Reads", "HQ Ambiguous Reads", "rRNA Reads",
    "End 1 Sense", "End 1 Antisense", "End 2 Sense", "End 2 Antisense", "Total Alignments",
] + [f"Filtered by tag {k}" for k in range(8)] + ["Split Reads", "Fragment Candidates", "Fragment Set Entries"]
COUNTER_INDEX = {n: i for i, n in enumerate(COUNTER_NAMES)}

READ_DTYPE = np.dtype([
    ("pos", "<i4"), ("mpos", "<i4"), ("endpos", "<i4"), ("isize", "<i4"), ("l_qseq", "<i4"), ("nm", "<i4"),
    ("flag", "<u2"), ("mapq", "u1"), ("aux", "u1"), ("contig", "<i2"), ("filter_mask", "u1"), ("reserved0", "u1"),
    ("name_hash", "<u8"), ("block_off", "<u4"), ("n_blocks", "<u2"), ("reserved1", "<u2")])
BLOCK_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4")])
# compact wire format (rsq_read_packed, 32 bytes)
PACKED_DTYPE = np.dtype([
    ("pos", "<i4"), ("mpos", "<i4"), ("isize", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("aux", "u1"), ("name_hash", "<u8"),
    ("contig", "<i2"), ("l_qseq", "<u2"), ("nm", "<u2"), ("n_blocks", "u1"), ("filter_mask", "u1")])
RSQ_AUX_IMPLICIT_BLOCK = 0x40


class RsqConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("mapq_threshold", C.c_uint32),
                ("base_mismatch_threshold", C.c_uint32), ("chimeric_distance", C.c_int32), ("strand_mode", C.c_uint32),
                ("unpaired", C.c_uint32), ("exclude_chimeric", C.c_uint32), ("n_filter_tags", C.c_uint32),
                ("fragment_samples", C.c_uint32), ("max_records", C.c_uint64), ("collect_coverage", C.c_uint32),
                ("legacy", C.c_uint32), ("user_stream", C.c_uint64), ("reserved", C.c_uint32 * 4)]


class RsqAnnotation(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_contigs", C.c_uint32), ("n_genes", C.c_uint32), ("n_exons", C.c_uint32),
                ("n_segs", C.c_uint32), ("n_exon_slots", C.c_uint32),
                ("exon_off", C.POINTER(C.c_uint32)), ("exon_start", C.POINTER(C.c_int32)), ("exon_end", C.POINTER(C.c_int32)),
                ("exon_gene", C.POINTER(C.c_uint32)), ("exon_slot", C.POINTER(C.c_uint32)), ("exon_flags", C.POINTER(C.c_uint8)),
                ("exon_cov_off", C.POINTER(C.c_uint64)), ("seg_off", C.POINTER(C.c_uint32)), ("seg_start", C.POINTER(C.c_int32)),
                ("seg_flags", C.POINTER(C.c_uint8)), ("gene_flags", C.POINTER(C.c_uint8)), ("cov_len", C.c_uint64),
                ("n_gene_lines", C.c_uint32), ("reserved1", C.c_uint32), ("gline_off", C.POINTER(C.c_uint32)),
                ("gline_start", C.POINTER(C.c_int32)), ("gline_end", C.POINTER(C.c_int32)), ("gline_gene", C.POINTER(C.c_uint32)),
                ("gline_flags", C.POINTER(C.c_uint8)), ("gline_order", C.POINTER(C.c_uint32)), ("exon_order", C.POINTER(C.c_uint32)),
                ("gline_fid", C.POINTER(C.c_uint32))]


class RsqBed(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_contigs", C.c_uint32), ("n_intervals", C.c_uint32), ("reserved", C.c_uint32),
                ("off", C.POINTER(C.c_uint32)), ("start", C.POINTER(C.c_int32)), ("end", C.POINTER(C.c_int32)),
                ("name_id", C.POINTER(C.c_uint32)), ("contig_present", C.POINTER(C.c_uint8))]


class RsqResults(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("counters", C.c_uint64 * RSQ_N_COUNTERS),
                ("gene_reads", C.POINTER(C.c_uint64)), ("gene_unique", C.POINTER(C.c_uint64)),
                ("gene_fragments", C.POINTER(C.c_uint64)), ("exon_whole", C.POINTER(C.c_uint64)),
                ("exon_frac", C.POINTER(C.c_uint64)), ("coverage", C.POINTER(C.c_uint32)),
                ("fragment_hist", C.POINTER(C.c_uint64)), ("fragment_hist_len", C.c_uint64),
                ("fragment_samples_taken", C.c_uint64), ("gc_hist", C.POINTER(C.c_uint64))]


class RsqReference(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_contigs", C.c_uint32), ("present", C.POINTER(C.c_uint8)),
                ("length", C.POINTER(C.c_uint64)), ("word_off", C.POINTER(C.c_uint64)), ("gc_bits", C.POINTER(C.c_uint32))]


class RsqTranscripts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_genes", C.c_uint32), ("n_exon_refs", C.c_uint32), ("coverage_mask", C.c_uint32),
                ("bias_offset", C.c_int32), ("bias_window", C.c_int32), ("bias_gene_length", C.c_uint32), ("reserved", C.c_uint32),
                ("gene_cov_off", C.POINTER(C.c_uint64)), ("gene_exon_off", C.POINTER(C.c_uint32)), ("gene_strand", C.POINTER(C.c_uint8)),
                ("exon_len", C.POINTER(C.c_uint32)), ("exon_slot", C.POINTER(C.c_uint32)), ("gene_owned", C.POINTER(C.c_uint8))]


class RsqCoverageStats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("gene_mean", C.POINTER(C.c_double)),
                ("gene_std", C.POINTER(C.c_double)), ("gene_cv", C.POINTER(C.c_double)), ("five_end", C.POINTER(C.c_uint64)),
                ("three_end", C.POINTER(C.c_uint64)), ("exon_cv", C.POINTER(C.c_double))]


class RunOptionsC(C.Structure):
    _fields_ = [("sample_name", C.c_char_p), ("sample_given", C.c_int), ("output_dir", C.c_char_p), ("use_rpkm", C.c_int),
                ("write_coverage", C.c_int), ("coverage_mask", C.c_uint), ("bias_offset", C.c_int), ("bias_window", C.c_int),
                ("bias_gene_length", C.c_ulong), ("detection_threshold", C.c_uint), ("filter_tags", C.POINTER(C.c_char_p)),
                ("n_filter_tags", C.c_int), ("exact_complexity_scan", C.c_int)]


class StreamStateC(C.Structure):
    _fields_ = [("read_length", C.c_int), ("alignment_count", C.c_ulonglong), ("final_contig", C.c_longlong),
                ("final_position", C.c_int), ("n_dropped", C.c_uint), ("dropped", C.c_uint * 1024)]


ABI_SYMBOLS = ["rsq_backend", "rsq_abi_version", "rsq_host_alloc", "rsq_host_free", "rsq_create", "rsq_destroy",
               "rsq_last_error", "rsq_set_annotation", "rsq_set_bed", "rsq_set_reference", "rsq_submit", "rsq_submit_packed", "rsq_submit_compact", "rsq_submit_device", "rsq_sync", "rsq_sync_inputs",
               "rsq_reset", "rsq_finalize", "rsq_coverage_statistics", "rsq_set_partition", "rsq_export_sample_counts", "rsq_export_counts",
               "rsq_import_counts", "rsq_comm_unique_id", "rsq_comm_init", "rsq_exchange", "rsq_algorithmic_bytes",
               "rsq_launch_count", "rsq_kernel_time_ms"]
RSQ_COMM_ID_BYTES = 128


def make_config(device: int = 0, mapq: int = 255, mismatches: int = 6, chimeric_distance: int = 2000000,
                strand_mode: int = 0, unpaired: bool = False, exclude_chimeric: bool = False, n_filter_tags: int = 0,
                fragment_samples: int = 0, max_records: int = 0, collect_coverage: bool = True, user_stream: int = 0,
                legacy: bool = False) -> RsqConfig:
    cfg = RsqConfig()
    cfg.struct_size = C.sizeof(RsqConfig)
    cfg.device = device
    cfg.mapq_threshold = mapq
    cfg.base_mismatch_threshold = mismatches
    cfg.chimeric_distance = chimeric_distance
    cfg.strand_mode = strand_mode
    cfg.unpaired = int(unpaired)
    cfg.exclude_chimeric = int(exclude_chimeric)
    cfg.n_filter_tags = n_filter_tags
    cfg.fragment_samples = fragment_samples
    cfg.max_records = max_records
    cfg.collect_coverage = int(collect_coverage)
    cfg.user_stream = user_stream
    cfg.legacy = int(legacy)
    return cfg


def pack_records(reads: np.ndarray, blocks: np.ndarray, out_reads: Optional[np.ndarray] = None,
                 out_blocks: Optional[np.ndarray] = None) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """rsq_pack_records: one submit batch (block_off relative to `blocks`) -> (packed records, remaining blocks), or None when
    the batch does not fit the wire format.  out_* may be preallocated (e.g. pinned) arrays of at least the input sizes."""
    H = host_lib()
    n, nb = reads.nbytes // 48, blocks.nbytes // 8
    pr = out_reads if out_reads is not None else np.empty(n, PACKED_DTYPE)
    pb = out_blocks if out_blocks is not None else np.empty(max(nb, 1), BLOCK_DTYPE)
    w = C.c_size_t(0)
    rc = H.rsqh_pack_records(reads.ctypes.data, n, blocks.ctypes.data, nb, pr.ctypes.data, pb.ctypes.data, C.byref(w))
    if rc != 0:
        return None
    return pr[:n], pb[:w.value]


def compact_bound(n_reads: int, n_blocks: int) -> int:
    return int(host_lib().rsqh_compact_bound(n_reads, n_blocks))


def pack_compact(reads: np.ndarray, blocks: np.ndarray, out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
    """rsq_pack_compact: one submit batch (block_off relative to `blocks`) -> the compact batch as a uint8 array (a view of
    `out` when given: a 16-byte aligned uint8 buffer of at least compact_bound bytes, e.g. pinned), or None when the batch does
    not suit the format."""
    H = host_lib()
    n, nb = reads.nbytes // 48, blocks.nbytes // 8
    cap = compact_bound(n, nb)
    if out is None:
        raw = np.empty(cap + 16, np.uint8)
        a = (-raw.ctypes.data) % 16
        out = raw[a:a + cap]
    assert out.nbytes >= cap and out.ctypes.data % 16 == 0
    w = C.c_size_t(0)
    rc = H.rsqh_pack_compact(reads.ctypes.data, n, blocks.ctypes.data, nb, out.ctypes.data, out.nbytes, C.byref(w))
    if rc != 0:
        return None
    return out[:w.value]


class RsqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rsq error {code}: {msg}")
        self.code = code


class Results:
    """Host copy of everything rsq_finalize returns."""

    def __init__(self, n_genes: int, n_exons: int, cov_len: int, hist_len: int = 1 << 16, want_coverage: bool = True, alloc=None):
        """alloc(n, dtype) -> zeroed array; default numpy (pageable).  Context.finalize(pinned=True) passes an allocator over
        rsq_host_alloc, so the device->host copies of the result tables run at PCIe speed instead of through a bounce buffer."""
        z = alloc or (lambda n, dt: np.zeros(n, dt))
        self.counters = np.zeros(RSQ_N_COUNTERS, np.uint64)
        self.gene_reads = z(n_genes, np.uint64)
        self.gene_unique = z(n_genes, np.uint64)
        self.gene_fragments = z(n_genes, np.uint64)
        self.exon_whole = z(n_exons, np.uint64)
        self.exon_frac = z(n_exons, np.uint64)
        self.coverage = z(cov_len + 1, np.uint32) if want_coverage else None
        self.fragment_hist = z(hist_len, np.uint64)
        self.fragment_samples_taken = 0
        self.gc_hist = z(100, np.uint64)

    def struct(self) -> RsqResults:
        r = RsqResults()
        r.struct_size = C.sizeof(RsqResults)
        p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
        r.gene_reads, r.gene_unique, r.gene_fragments = p64(self.gene_reads), p64(self.gene_unique), p64(self.gene_fragments)
        r.exon_whole, r.exon_frac = p64(self.exon_whole), p64(self.exon_frac)
        if self.coverage is not None:
            r.coverage = self.coverage.ctypes.data_as(C.POINTER(C.c_uint32))
        r.fragment_hist = p64(self.fragment_hist)
        r.fragment_hist_len = len(self.fragment_hist)
        r.gc_hist = p64(self.gc_hist)
        return r

    def counter(self, name: str) -> int:
        return int(self.counters[COUNTER_INDEX[name]])

    def exon_counts(self) -> np.ndarray:
        """exact Q64.64 -> float64 (what exon_reads.gct prints)"""
        return self.exon_whole.astype(np.float64) + self.exon_frac.astype(np.float64) * 2.0 ** -64


class Backend:
    """One loaded library that implements include/rsq_b200.h."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is not built (run `make` / __graft_entry__.build())")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.rsq_backend.restype = C.c_char_p
        L.rsq_abi_version.restype = C.c_uint32
        L.rsq_host_alloc.restype = C.c_void_p
        L.rsq_host_alloc.argtypes = [C.c_size_t]
        L.rsq_host_free.argtypes = [C.c_void_p]
        L.rsq_create.argtypes = [C.POINTER(RsqConfig), C.POINTER(C.c_void_p)]
        L.rsq_destroy.argtypes = [C.c_void_p]
        L.rsq_last_error.restype = C.c_char_p
        L.rsq_last_error.argtypes = [C.c_void_p]
        L.rsq_set_annotation.argtypes = [C.c_void_p, C.POINTER(RsqAnnotation)]
        L.rsq_set_bed.argtypes = [C.c_void_p, C.POINTER(RsqBed)]
        L.rsq_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.rsq_submit_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.rsq_submit_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.rsq_submit_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.rsq_set_reference.argtypes = [C.c_void_p, C.POINTER(RsqReference)]
        L.rsq_sync.argtypes = [C.c_void_p]
        L.rsq_sync_inputs.argtypes = [C.c_void_p]
        L.rsq_reset.argtypes = [C.c_void_p]
        L.rsq_finalize.argtypes = [C.c_void_p, C.POINTER(RsqResults)]
        L.rsq_coverage_statistics.argtypes = [C.c_void_p, C.POINTER(RsqTranscripts), C.POINTER(RsqCoverageStats)]
        L.rsq_export_counts.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rsq_import_counts.argtypes = [C.c_void_p]
        L.rsq_set_partition.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.rsq_export_sample_counts.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rsq_comm_unique_id.argtypes = [C.c_void_p]
        L.rsq_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.rsq_exchange.argtypes = [C.c_void_p]
        L.rsq_algorithmic_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rsq_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rsq_kernel_time_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]

    @property
    def name(self) -> str:
        return self.lib.rsq_backend().decode()

    def comm_unique_id(self) -> bytes:
        """rsq_comm_unique_id: created by ONE rank and handed to all (bench.py broadcasts it through torch.distributed)"""
        buf = C.create_string_buffer(RSQ_COMM_ID_BYTES)
        rc = self.lib.rsq_comm_unique_id(buf)
        if rc != RSQ_OK:
            raise RsqError(rc, self.lib.rsq_last_error(None).decode())
        return buf.raw

    def create(self, cfg: RsqConfig) -> "Context":
        h = C.c_void_p()
        rc = self.lib.rsq_create(C.byref(cfg), C.byref(h))
        if rc != RSQ_OK:
            raise RsqError(rc, self.lib.rsq_last_error(None).decode())
        return Context(self, h, cfg)


class Context:
    def __init__(self, be: Backend, handle: C.c_void_p, cfg: RsqConfig):
        self.be, self.h, self.cfg = be, handle, cfg
        self.n_genes = self.n_exons = self.cov_len = 0
        self._pinned = {}          # name -> (pointer, numpy view): page-locked result buffers, reused by every pinned finalize

    def _check(self, rc: int) -> None:
        if rc != RSQ_OK:
            raise RsqError(rc, self.be.lib.rsq_last_error(self.h).decode())

    def close(self) -> None:
        if self.h:
            for ptr, _ in self._pinned.values():
                self.be.lib.rsq_host_free(C.c_void_p(ptr))
            self._pinned = {}
            self.be.lib.rsq_destroy(self.h)
            self.h = None

    def _pinned_array(self, name: str, n: int, dtype) -> np.ndarray:
        """Page-locked host array owned by this context (rsq_host_alloc), zeroed when first allocated; the SAME memory (with
        the previous call's content, which rsq_finalize / rsq_coverage_statistics overwrite in full) is handed out again by
        the next call with this name, so results that must survive the next finalize have to be copied."""
        nbytes = max(int(n) * np.dtype(dtype).itemsize, 8)
        have = self._pinned.get(name)
        if have is None or have[1].nbytes < nbytes:
            if have is not None:
                self.be.lib.rsq_host_free(C.c_void_p(have[0]))
            ptr = self.be.lib.rsq_host_alloc(nbytes)
            if not ptr:
                raise MemoryError("rsq_host_alloc failed")
            have = (ptr, np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)))
            have[1][:] = 0
            self._pinned[name] = have
        return have[1][:int(n) * np.dtype(dtype).itemsize].view(dtype)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_annotation(self, ann: "Annotation") -> None:
        v = ann.view
        self.n_genes, self.n_exons, self.cov_len = v.n_genes, (v.n_exon_slots or v.n_exons), v.cov_len   # n_exons: output slots
        self._check(self.be.lib.rsq_set_annotation(self.h, C.byref(v)))
        if ann.bed_view is not None and self.cfg.fragment_samples:
            self._check(self.be.lib.rsq_set_bed(self.h, ann.bed_view))
        if ann.ref_view is not None:
            self._check(self.be.lib.rsq_set_reference(self.h, ann.ref_view))

    def submit(self, reads: np.ndarray, blocks: np.ndarray) -> None:
        """host arrays (structured READ_DTYPE/BLOCK_DTYPE, or raw uint8/int32 views of the same bytes)"""
        assert reads.flags["C_CONTIGUOUS"] and blocks.flags["C_CONTIGUOUS"]
        n_reads = reads.nbytes // 48
        n_blocks = blocks.nbytes // 8
        self._check(self.be.lib.rsq_submit(self.h, reads.ctypes.data, n_reads, blocks.ctypes.data, n_blocks))

    def submit_ptr(self, reads_ptr: int, n_reads: int, blocks_ptr: int, n_blocks: int) -> None:
        self._check(self.be.lib.rsq_submit(self.h, reads_ptr, n_reads, blocks_ptr, n_blocks))

    def submit_packed(self, reads: np.ndarray, blocks: np.ndarray) -> None:
        """host arrays in the compact wire format (PACKED_DTYPE + the blocks of the records without an implicit block)"""
        assert reads.flags["C_CONTIGUOUS"] and blocks.flags["C_CONTIGUOUS"]
        self._check(self.be.lib.rsq_submit_packed(self.h, reads.ctypes.data, reads.nbytes // 32, blocks.ctypes.data, blocks.nbytes // 8))

    def submit_packed_ptr(self, reads_ptr: int, n_reads: int, blocks_ptr: int, n_blocks: int) -> None:
        self._check(self.be.lib.rsq_submit_packed(self.h, reads_ptr, n_reads, blocks_ptr, n_blocks))

    def submit_compact(self, batch: np.ndarray) -> None:
        """one compact batch (pack_compact): the 16-byte wire format"""
        assert batch.flags["C_CONTIGUOUS"]
        self._check(self.be.lib.rsq_submit_compact(self.h, batch.ctypes.data, batch.nbytes))

    def submit_compact_ptr(self, ptr: int, nbytes: int) -> None:
        self._check(self.be.lib.rsq_submit_compact(self.h, ptr, nbytes))

    def submit_device(self, reads_ptr: int, n_reads: int, blocks_ptr: int, n_blocks: int) -> None:
        self._check(self.be.lib.rsq_submit_device(self.h, reads_ptr, n_reads, blocks_ptr, n_blocks))

    def sync(self) -> None:
        self._check(self.be.lib.rsq_sync(self.h))

    def reset(self) -> None:
        self._check(self.be.lib.rsq_reset(self.h))

    def finalize(self, want_coverage: bool = True, hist_len: int = 1 << 16, pinned: bool = False) -> Results:
        """pinned=True: the result tables live in page-locked buffers of this context and are overwritten by the next
        pinned finalize (bench.py; a caller that keeps results copies them)."""
        if pinned:
            names = iter(range(1 << 30))
            res = Results(self.n_genes, self.n_exons, self.cov_len, hist_len, want_coverage,
                          alloc=lambda n, dt: self._pinned_array(f"fin{next(names)}", n, dt))
        else:
            res = Results(self.n_genes, self.n_exons, self.cov_len, hist_len, want_coverage)
        st = res.struct()
        self._check(self.be.lib.rsq_finalize(self.h, C.byref(st)))
        res.counters[:] = np.frombuffer(st.counters, dtype=np.uint64)
        res.fragment_samples_taken = int(st.fragment_samples_taken)
        return res

    def coverage_statistics(self, ann: "Annotation", mask: int = 500, bias_offset: int = 0, bias_window: int = 100,
                            bias_gene_length: int = 200, pinned: bool = False, owned: Optional[np.ndarray] = None) -> dict:
        """per-gene-line coverage statistics (rsq_coverage_statistics); arrays indexed by gene line (contig, start order).
        owned: uint8 per gene line (rsq_transcripts.gene_owned) in coordinate-partitioned runs."""
        q = RsqTranscripts()
        host_lib().rsqh_transcripts(ann._h, mask, bias_offset, bias_window, bias_gene_length, C.byref(q))
        if owned is not None:
            owned = np.ascontiguousarray(owned, np.uint8)
            assert len(owned) == q.n_genes
            q.gene_owned = owned.ctypes.data_as(C.POINTER(C.c_uint8))
        n = q.n_genes
        z = (lambda name, k, dt: self._pinned_array("cs_" + name, k, dt)) if pinned else (lambda name, k, dt: np.zeros(k, dt))
        out = {"gene_mean": z("gene_mean", n, np.float64), "gene_std": z("gene_std", n, np.float64), "gene_cv": z("gene_cv", n, np.float64),
               "five_end": z("five_end", n, np.uint64), "three_end": z("three_end", n, np.uint64), "exon_cv": z("exon_cv", self.n_exons, np.float64)}
        st = RsqCoverageStats()
        st.struct_size = C.sizeof(RsqCoverageStats)
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        st.gene_mean, st.gene_std, st.gene_cv, st.exon_cv = pd(out["gene_mean"]), pd(out["gene_std"]), pd(out["gene_cv"]), pd(out["exon_cv"])
        st.five_end = out["five_end"].ctypes.data_as(C.POINTER(C.c_uint64))
        st.three_end = out["three_end"].ctypes.data_as(C.POINTER(C.c_uint64))
        self._check(self.be.lib.rsq_coverage_statistics(self.h, C.byref(q), C.byref(st)))
        return out

    def set_partition(self, rank: int, world: int) -> None:
        self._check(self.be.lib.rsq_set_partition(self.h, rank, world))

    def export_sample_counts(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.be.lib.rsq_export_sample_counts(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def comm_init(self, comm_id: bytes, rank: int, world: int) -> None:
        """collective over all ranks (NCCL inside the library)"""
        assert len(comm_id) == RSQ_COMM_ID_BYTES
        self._check(self.be.lib.rsq_comm_init(self.h, C.c_char_p(comm_id), rank, world))

    def exchange(self) -> None:
        """collective: the end-of-run exchange of a partitioned run, on the context's stream"""
        self._check(self.be.lib.rsq_exchange(self.h))

    def export_counts(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.be.lib.rsq_export_counts(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def import_counts(self) -> None:
        self._check(self.be.lib.rsq_import_counts(self.h))

    def algorithmic_bytes(self) -> int:
        v = C.c_uint64()
        self._check(self.be.lib.rsq_algorithmic_bytes(self.h, C.byref(v)))
        return int(v.value)

    def launch_count(self) -> int:
        v = C.c_uint64()
        self._check(self.be.lib.rsq_launch_count(self.h, C.byref(v)))
        return int(v.value)

    def kernel_time_ms(self) -> float:
        v = C.c_double()
        self._check(self.be.lib.rsq_kernel_time_ms(self.h, C.byref(v)))
        return float(v.value)


def default_backend() -> Backend:
    """The product library.  Raises if it is not built; Context creation raises RsqError(NO_DEVICE) without a GPU."""
    return Backend(PRODUCT_LIB)


# ------------------------------------------------------------------------------------------------------------------
# host helper library: GTF/BED parsing, SAM/BAM decode, report writing (C++; librsq_host.so)
# ------------------------------------------------------------------------------------------------------------------
_host = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise FileNotFoundError(f"{HOST_LIB} is not built (run `make`)")
        H = C.CDLL(HOST_LIB)
        H.rsqh_load_annotation.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        H.rsqh_free_annotation.argtypes = [C.c_void_p]
        H.rsqh_annotation_view.restype = C.POINTER(RsqAnnotation)
        H.rsqh_annotation_view.argtypes = [C.c_void_p]
        H.rsqh_bed_view.restype = C.POINTER(RsqBed)
        H.rsqh_bed_view.argtypes = [C.c_void_p]
        H.rsqh_reference_view.restype = C.POINTER(RsqReference)
        H.rsqh_reference_view.argtypes = [C.c_void_p]
        H.rsqh_set_fasta.argtypes = [C.c_char_p]
        H.rsqh_transcripts.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_ulong, C.POINTER(RsqTranscripts)]
        H.rsqh_gene_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        H.rsqh_islands.restype = C.c_size_t
        H.rsqh_islands.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("rsqh_n_gene_list", "rsqh_n_contigs"):
            getattr(H, f).restype = C.c_uint
            getattr(H, f).argtypes = [C.c_void_p]
        H.rsqh_gene_list_index.restype = C.c_uint
        H.rsqh_gene_list_index.argtypes = [C.c_void_p, C.c_uint]
        for f in ("rsqh_gene_id", "rsqh_exon_id", "rsqh_contig_name"):
            getattr(H, f).restype = C.c_char_p
            getattr(H, f).argtypes = [C.c_void_p, C.c_uint]
        H.rsqh_contig_index.restype = C.c_int
        H.rsqh_contig_index.argtypes = [C.c_void_p, C.c_char_p]
        H.rsqh_decode_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                       C.POINTER(StreamStateC), C.c_char_p, C.c_size_t]
        H.rsqh_free.argtypes = [C.c_void_p]
        H.rsqh_pack_records.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        H.rsqh_compact_bound.restype = C.c_size_t
        H.rsqh_compact_bound.argtypes = [C.c_size_t, C.c_size_t]
        H.rsqh_pack_compact.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        H.rsqh_hash_qname.restype = C.c_uint64
        H.rsqh_hash_qname.argtypes = [C.c_char_p, C.c_size_t]
        H.rsqh_write_reports.argtypes = [C.c_void_p, C.POINTER(RsqResults), C.c_void_p, C.POINTER(RsqCoverageStats), C.POINTER(RunOptionsC),
                                         C.POINTER(StreamStateC), C.c_char_p, C.c_size_t]
        _host = H
    return _host


class Annotation:
    """Flattened annotation parsed from GTF (+BED) text by the C++ host library."""

    def __init__(self, gtf: str, bed: Optional[str] = None, legacy: bool = False, fasta: Optional[str] = None):
        H = host_lib()
        self.legacy = legacy
        H.rsqh_set_legacy(int(legacy))
        H.rsqh_set_fasta(fasta.encode() if fasta else None)
        self._h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = H.rsqh_load_annotation(gtf.encode(), bed.encode() if bed else None, C.byref(self._h), err, 1024)
        if rc != 0:
            raise RuntimeError(f"annotation load failed (exit code {rc}): {err.value.decode()}")
        self.view = H.rsqh_annotation_view(self._h).contents
        bv = H.rsqh_bed_view(self._h)
        self.bed_view = bv if bv else None
        rv = H.rsqh_reference_view(self._h)
        self.ref_view = rv if rv else None
        H.rsqh_set_fasta(None)

    def __del__(self):
        try:
            if self._h:
                host_lib().rsqh_free_annotation(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def n_genes(self) -> int:
        return self.view.n_genes

    @property
    def n_exons(self) -> int:
        """Output slots (exonList); under --legacy larger than the number of exon intervals kept."""
        return self.view.n_exon_slots or self.view.n_exons

    @property
    def n_exon_intervals(self) -> int:
        return self.view.n_exons

    @property
    def cov_len(self) -> int:
        return self.view.cov_len

    def contig_index(self, name: str) -> int:
        return host_lib().rsqh_contig_index(self._h, name.encode())

    def contig_names(self) -> List[str]:
        H = host_lib()
        return [H.rsqh_contig_name(self._h, c).decode() for c in range(H.rsqh_n_contigs(self._h))]

    def gene_lines(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(annotation contig id, start, end) of the kept gene lines, in the order of the coverage-statistics request"""
        q = RsqTranscripts()
        host_lib().rsqh_transcripts(self._h, 500, 0, 100, 200, C.byref(q))
        c, s, e = np.zeros(q.n_genes, np.uint32), np.zeros(q.n_genes, np.int64), np.zeros(q.n_genes, np.int64)
        host_lib().rsqh_gene_lines(self._h, c.ctypes.data, s.ctypes.data, e.ctypes.data)
        return c, s, e

    def islands(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(annotation contig id, start, end) of the merged gene hulls + BED intervals: a partitioned stream may be cut
        between two islands and nowhere else (csrc/host/annotation.h)"""
        H = host_lib()
        n = H.rsqh_islands(self._h, None, None, None)
        c, s, e = np.zeros(n, np.uint32), np.zeros(n, np.int64), np.zeros(n, np.int64)
        H.rsqh_islands(self._h, c.ctypes.data, s.ctypes.data, e.ctypes.data)
        return c, s, e

    def gene_ids(self) -> List[str]:
        H = host_lib()
        return [H.rsqh_gene_id(self._h, g).decode() for g in range(self.n_genes)]

    def array(self, field: str, n: int, dtype) -> np.ndarray:
        ptr = getattr(self.view, field)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    def decode(self, path: str, chimeric_tag: str = "ch", filter_tags: Sequence[str] = (), exclude_chimeric: bool = False,
               chimeric_distance: int = 2000000) -> Tuple[np.ndarray, np.ndarray, StreamStateC]:
        """SAM/BAM -> (reads, blocks, stream state) through the C++ decoder."""
        H = host_lib()
        tags = (C.c_char_p * max(len(filter_tags), 1))(*[t.encode() for t in filter_tags])
        pr, pb = C.c_void_p(), C.c_void_p()
        nr, nbk = C.c_size_t(), C.c_size_t()
        st = StreamStateC()
        err = C.create_string_buffer(1024)
        rc = H.rsqh_decode_file(self._h, path.encode(), chimeric_tag.encode(), tags, len(filter_tags), int(exclude_chimeric),
                                chimeric_distance, C.byref(pr), C.byref(nr), C.byref(pb), C.byref(nbk), C.byref(st), err, 1024)
        if rc != 0:
            raise RuntimeError(f"decode failed (exit code {rc}): {err.value.decode()}")
        reads = np.frombuffer(C.string_at(pr.value, nr.value * 48), dtype=READ_DTYPE).copy() if nr.value else np.zeros(0, READ_DTYPE)
        blocks = np.frombuffer(C.string_at(pb.value, nbk.value * 8), dtype=BLOCK_DTYPE).copy() if nbk.value else np.zeros(0, BLOCK_DTYPE)
        H.rsqh_free(pr)
        H.rsqh_free(pb)
        return reads, blocks, st

    def write_reports(self, res: Results, cov_stats: dict, out_dir: str, sample: str, state: StreamStateC, sample_given: bool = False,
                      write_coverage: bool = False, use_rpkm: bool = False, filter_tags: Sequence[str] = (),
                      exon_seq: Optional[np.ndarray] = None, coverage_mask: int = 500, detection_threshold: int = 5) -> None:
        H = host_lib()
        o = RunOptionsC()
        o.sample_name = sample.encode(); o.sample_given = int(sample_given); o.output_dir = out_dir.encode()
        o.use_rpkm = int(use_rpkm); o.write_coverage = int(write_coverage); o.coverage_mask = coverage_mask
        o.bias_offset = 0; o.bias_window = 100; o.bias_gene_length = 200; o.detection_threshold = detection_threshold
        tags = (C.c_char_p * max(len(filter_tags), 1))(*[t.encode() for t in filter_tags])
        o.filter_tags = tags; o.n_filter_tags = len(filter_tags); o.exact_complexity_scan = 0
        st = res.struct()
        C.memmove(st.counters, res.counters.ctypes.data, 8 * RSQ_N_COUNTERS)
        st.fragment_samples_taken = res.fragment_samples_taken
        cs = RsqCoverageStats()
        cs.struct_size = C.sizeof(RsqCoverageStats)
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        cs.gene_mean, cs.gene_std, cs.gene_cv, cs.exon_cv = pd(cov_stats["gene_mean"]), pd(cov_stats["gene_std"]), pd(cov_stats["gene_cv"]), pd(cov_stats["exon_cv"])
        cs.five_end = cov_stats["five_end"].ctypes.data_as(C.POINTER(C.c_uint64))
        cs.three_end = cov_stats["three_end"].ctypes.data_as(C.POINTER(C.c_uint64))
        err = C.create_string_buffer(1024)
        rc = H.rsqh_write_reports(self._h, C.byref(st), exon_seq.ctypes.data if exon_seq is not None else None, C.byref(cs), C.byref(o),
                                  C.byref(state), err, 1024)
        if rc != 0:
            raise RuntimeError(f"write_reports failed (exit code {rc}): {err.value.decode()}")
