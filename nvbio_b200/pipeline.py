"""Seed + extend composition (examples/fmmap/fmmap.cu:255-400 shaped) over the C ABI."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional
import torch
from ._lib import lib, check, SeedExtendParamsStruct, BestAlignmentOutStruct, PairParamsStruct, PairOutStruct
from .strings import PackedStringSet
from .fmindex import FMIndexDevice
from . import aln


@dataclass
class SeedExtendParams:
    seed_len: int = 20
    seed_interval: int = 10          # int(1 + 0.75*sqrtf(150)): nvBowtie's SimpleFunc evaluated on the host (params.cpp:157-158)
    band_len: int = 31
    type: int = aln.LOCAL
    both_strands: bool = True
    max_seed_hits: int = 100         # nvBowtie max_hits
    dedup_jobs: bool = True          # score identical (strand, window) jobs of a read once
    scheme: object = field(default_factory=lambda: aln.SimpleGotohScheme(2, -2, -5, -3))
    read_quals: Optional[torch.Tensor] = None   # uint8 base qualities indexed like the read symbols (with a QualityGotohScheme)

    def struct(self) -> SeedExtendParamsStruct:
        p = SeedExtendParamsStruct()
        p.seed_len, p.seed_interval, p.band_len, p.type = self.seed_len, self.seed_interval, self.band_len, self.type
        p.both_strands = 1 if self.both_strands else 0
        p.max_seed_hits = self.max_seed_hits
        p.dedup_jobs = 1 if self.dedup_jobs else 0
        p.scheme = self.scheme.struct()
        p.d_read_quals = self.read_quals.data_ptr() if self.read_quals is not None else None
        return p


class SeedExtendWorkspace:
    """pre-allocated outputs + temp storage for repeated calls on equally-shaped batches"""

    def __init__(self, fmi: FMIndexDevice, genome: torch.Tensor, reads: PackedStringSet, params: SeedExtendParams,
                 hit_capacity: int, keep_hits: bool = False, traceback: bool = False):
        dev = fmi.device
        n = reads.count
        self.best_score = torch.empty(n, dtype=torch.int32, device=dev)
        self.best_pos = torch.empty(n, dtype=torch.int32, device=dev)
        self.n_hits = torch.zeros(3, dtype=torch.int32, device=dev)      # kept, found, distinct alignment jobs
        self.hit_capacity = hit_capacity
        self.hit_read = self.hit_window = self.hit_score = self.hit_sink = None
        if keep_hits:
            self.hit_read = torch.empty(hit_capacity, dtype=torch.int32, device=dev)
            self.hit_window = torch.empty((hit_capacity, 2), dtype=torch.int32, device=dev)
            self.hit_score = torch.empty(hit_capacity, dtype=torch.int32, device=dev)
            self.hit_sink = torch.empty((hit_capacity, 2), dtype=torch.int32, device=dev)
        # optional alignment (CIGAR ops, begin, strand) of every read's best hit
        self.best_ops = self.best_n_ops = self.best_begin = self.best_strand = None
        self.max_ops = reads.length + params.band_len + 1
        if traceback:
            self.best_ops = torch.zeros((n, self.max_ops), dtype=torch.uint8, device=dev)
            self.best_n_ops = torch.zeros(n, dtype=torch.int32, device=dev)
            self.best_begin = torch.empty((n, 2), dtype=torch.int32, device=dev)
            self.best_strand = torch.empty(n, dtype=torch.uint8, device=dev)
        tb = C.c_size_t(0)
        r = _call(fmi, genome, reads, params, self, None, tb)
        if r != -2:
            check(r, "nvb_seed_extend(size query)")
        self.temp = torch.empty(tb.value, dtype=torch.uint8, device=dev)
        self.temp_bytes = tb.value


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _call(fmi, genome, reads, params, ws, temp, tb):
    s, rd, ps = fmi.struct(), reads.struct(), params.struct()
    if ws.best_ops is not None:
        ba = BestAlignmentOutStruct()
        ba.d_ops, ba.max_ops, ba.d_n_ops = ws.best_ops.data_ptr(), ws.max_ops, ws.best_n_ops.data_ptr()
        ba.d_begin, ba.d_strand = ws.best_begin.data_ptr(), ws.best_strand.data_ptr()
        return lib().nvb_seed_extend_traceback(C.byref(s), _p(genome), C.byref(rd), C.c_uint32(reads.count), C.byref(ps),
                                               C.c_uint32(ws.hit_capacity), _p(ws.best_score), _p(ws.best_pos), _p(ws.n_hits),
                                               _p(ws.hit_read), _p(ws.hit_window), _p(ws.hit_score), _p(ws.hit_sink), C.byref(ba),
                                               _p(temp), C.byref(tb), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return lib().nvb_seed_extend(C.byref(s), _p(genome), C.byref(rd), C.c_uint32(reads.count), C.byref(ps),
                                 C.c_uint32(ws.hit_capacity), _p(ws.best_score), _p(ws.best_pos), _p(ws.n_hits),
                                 _p(ws.hit_read), _p(ws.hit_window), _p(ws.hit_score), _p(ws.hit_sink),
                                 _p(temp), C.byref(tb), C.c_void_p(torch.cuda.current_stream().cuda_stream))


def seed_extend(fmi: FMIndexDevice, genome: torch.Tensor, reads: PackedStringSet, params: SeedExtendParams,
                workspace: Optional[SeedExtendWorkspace] = None, hit_capacity: Optional[int] = None, keep_hits: bool = False,
                traceback: bool = False):
    """returns the workspace: .best_score[n], .best_pos[n], .n_hits[3] = (kept, total, distinct jobs), optional per-hit
    arrays, and with traceback=True the alignment of every read's best hit (.best_ops END->START, .best_n_ops,
    .best_begin = (genome start, read start), .best_strand)"""
    if workspace is None:
        if hit_capacity is None:
            hit_capacity = 32 * reads.count + 1024
        workspace = SeedExtendWorkspace(fmi, genome, reads, params, hit_capacity, keep_hits, traceback)
    tb = C.c_size_t(workspace.temp_bytes)
    check(_call(fmi, genome, reads, params, workspace, workspace.temp, tb), "nvb_seed_extend")
    return workspace


PAIR_UNPAIRED, PAIR_CONCORDANT, PAIR_RESCUED_MATE1, PAIR_RESCUED_MATE2 = 0, 1, 2, 4


@dataclass
class PairParams:
    """fragment constraints of the paired-end stage (nvBowtie --minins / --maxins, FR orientation)"""
    min_frag: int = 0
    max_frag: int = 500
    min_mate_score: int = 60          # a mate's alignment (anchor or rescued) must reach this score to take part in a pair
    rescue_capacity: Optional[int] = None

    def struct(self, n_pairs) -> PairParamsStruct:
        p = PairParamsStruct()
        p.min_frag, p.max_frag, p.min_mate_score = self.min_frag, self.max_frag, self.min_mate_score
        p.rescue_capacity = 2 * n_pairs if self.rescue_capacity is None else self.rescue_capacity
        return p


class PairedWorkspace:
    """outputs + temp storage of seed_extend_paired for repeated calls on equally-shaped batches"""

    def __init__(self, fmi, genome, reads: PackedStringSet, params: SeedExtendParams, pair: PairParams, hit_capacity: int):
        dev = fmi.device
        assert reads.count % 2 == 0
        self.n_pairs = n = reads.count // 2
        self.hit_capacity = hit_capacity
        self.pair_score = torch.empty(n, dtype=torch.int32, device=dev)
        self.pair_flags = torch.empty(n, dtype=torch.int32, device=dev)
        self.mate_score = torch.empty((2, n), dtype=torch.int32, device=dev)
        self.mate_pos = torch.empty((2, n), dtype=torch.int32, device=dev)
        self.mate_strand = torch.empty((2, n), dtype=torch.uint8, device=dev)
        self.n_rescue = torch.zeros(2, dtype=torch.int32, device=dev)
        self.n_hits = torch.zeros(3, dtype=torch.int32, device=dev)
        tb = C.c_size_t(0)
        r = _call_paired(fmi, genome, reads, params, pair, self, None, tb)
        if r != -2:
            check(r, "nvb_seed_extend_paired(size query)")
        self.temp = torch.empty(tb.value, dtype=torch.uint8, device=dev)
        self.temp_bytes = tb.value


def _call_paired(fmi, genome, reads, params, pair, ws, temp, tb):
    s, rd, ps, pp = fmi.struct(), reads.struct(), params.struct(), pair.struct(ws.n_pairs)
    po = PairOutStruct()
    po.d_pair_score, po.d_pair_flags = ws.pair_score.data_ptr(), ws.pair_flags.data_ptr()
    po.d_mate_score, po.d_mate_pos, po.d_mate_strand = ws.mate_score.data_ptr(), ws.mate_pos.data_ptr(), ws.mate_strand.data_ptr()
    po.d_n_rescue = ws.n_rescue.data_ptr()
    return lib().nvb_seed_extend_paired(C.byref(s), _p(genome), C.byref(rd), C.c_uint32(ws.n_pairs), C.byref(ps), C.c_uint32(ws.hit_capacity),
                                        C.byref(pp), C.byref(po), _p(ws.n_hits), _p(temp), C.byref(tb),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))


def seed_extend_paired(fmi: FMIndexDevice, genome: torch.Tensor, reads: PackedStringSet, params: SeedExtendParams, pair: PairParams,
                       workspace: Optional[PairedWorkspace] = None, hit_capacity: Optional[int] = None):
    """paired-end seed + extend (reads = mate 1 of every pair, then mate 2 of every pair): concordant pairs straight from the two
    independent alignments, opposite-mate full-DP rescue for the rest (nvBowtie's best-approx paired flow,
    score_opposite_inl.h:90-266).  Returns the workspace: .pair_score[n], .pair_flags[n], .mate_score/.mate_pos/.mate_strand[2,n],
    .n_rescue[2] = (full-DP jobs run, wanted)"""
    if workspace is None:
        if hit_capacity is None:
            hit_capacity = 32 * reads.count + 1024
        workspace = PairedWorkspace(fmi, genome, reads, params, pair, hit_capacity)
    tb = C.c_size_t(workspace.temp_bytes)
    check(_call_paired(fmi, genome, reads, params, pair, workspace, workspace.temp, tb), "nvb_seed_extend_paired")
    return workspace


STAGES = ("strings", "seed_match", "hit_slots", "locate_windows", "dedup", "extend", "reduce")


def last_stage_ms():
    """device time (ms) of the stages of the most recent seed_extend call"""
    ms = (C.c_float * 7)()
    check(lib().nvb_seed_extend_stage_ms(ms), "nvb_seed_extend_stage_ms")
    return dict(zip(STAGES, [float(v) for v in ms]))


def _host_view(ptr, count, dtype):
    """a torch tensor over `count` elements of pinned host memory owned by the C library (no copy)"""
    nbytes = count * torch.tensor([], dtype=dtype).element_size()
    buf = (C.c_char * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype, count=count)


class StreamingSeedExtend:
    """Host-to-host batches through the C ABI's nvb_pipeline (include/nvbio_b200.h): the production entry point that replaces
    nvBowtie's input thread -> compute thread hand-off (nvBowtie/bowtie2/cuda/compute_thread.cu:213-243) and its per-stage
    cudaDeviceSynchronize.

    `submit(host_words)` enqueues H2D copy -> seed_extend[_paired] -> D2H copy of the per-read results and returns a ticket;
    `result(ticket)` waits for that batch only.  `depth` batches are in flight: copies overlap kernels, and consecutive batches
    run on different compute streams.  host_words: int32 tensor [n_reads, words_per_read] in host memory (pinned = async copy).
    pair: PairParams for paired-end batches (reads = mate 1 of every pair, then mate 2)."""

    def __init__(self, fmi: FMIndexDevice, genome: torch.Tensor, params: SeedExtendParams, n_reads: int, read_len: int,
                 words_per_read: int, hit_capacity: Optional[int] = None, depth: int = 2, bits: int = 2, pair: Optional["PairParams"] = None):
        self.fmi, self.genome, self.params, self.pair = fmi, genome, params, pair          # keep the device buffers alive
        self.n_reads, self.read_len, self.wpr, self.bits, self.depth = n_reads, read_len, words_per_read, bits, depth
        if hit_capacity is None:
            hit_capacity = 32 * n_reads + 1024
        self._params_struct = params.struct()
        s = fmi.struct()
        pp = pair.struct(n_reads // 2) if pair is not None else None
        self._h = C.c_void_p()
        with torch.cuda.device(fmi.device):
            check(lib().nvb_pipeline_create(C.byref(s), C.c_void_p(genome.data_ptr()), C.byref(self._params_struct),
                                            C.byref(pp) if pp is not None else None,
                                            C.c_uint32(n_reads), C.c_uint32(read_len), C.c_uint32(words_per_read), C.c_uint32(bits),
                                            C.c_uint32(hit_capacity), C.c_uint32(depth), C.byref(self._h)), "nvb_pipeline_create")
        h2d, d2h = C.c_size_t(0), C.c_size_t(0)
        lib().nvb_pipeline_traffic(self._h, C.byref(h2d), C.byref(d2h))
        self.h2d_bytes, self.d2h_bytes = int(h2d.value), int(d2h.value)
        self._keep = {}
        self.last_device_ms = None

    def submit(self, host_words: torch.Tensor) -> int:
        assert not host_words.is_cuda and host_words.is_contiguous() and host_words.numel() == self.n_reads * self.wpr
        t = C.c_uint32(0)
        check(lib().nvb_pipeline_submit(self._h, C.c_void_p(host_words.data_ptr()), C.byref(t)), "nvb_pipeline_submit")
        self._keep[int(t.value)] = host_words                # the copy is asynchronous: keep the source alive until result()
        return int(t.value)

    def result(self, ticket: int):
        """single end: (best_score[n], best_pos[n], n_hits[3]); paired: dict of the nvb_pair_out arrays -- views of the pipeline's
        pinned host buffers, valid until `depth` further batches have been submitted"""
        from ._lib import PipelineResultStruct
        r = PipelineResultStruct()
        check(lib().nvb_pipeline_wait(self._h, C.c_uint32(ticket), C.byref(r)), "nvb_pipeline_wait")
        self._keep.pop(ticket, None)
        self.last_device_ms = float(r.device_ms)
        n = self.n_reads
        if self.pair is None:
            return _host_view(r.best_score, n, torch.int32), _host_view(r.best_pos, n, torch.int32), _host_view(r.n_hits, 3, torch.int32)
        return dict(pair_score=_host_view(r.pair_score, n // 2, torch.int32), pair_flags=_host_view(r.pair_flags, n // 2, torch.int32),
                    mate_score=_host_view(r.mate_score, n, torch.int32).view(2, n // 2), mate_pos=_host_view(r.mate_pos, n, torch.int32).view(2, n // 2),
                    mate_strand=_host_view(r.mate_strand, n, torch.uint8).view(2, n // 2), n_rescue=_host_view(r.n_rescue, 2, torch.int32),
                    n_hits=_host_view(r.n_hits, 3, torch.int32))

    def close(self):
        if self._h:
            lib().nvb_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
