"""Seed + extend composition (examples/fmmap/fmmap.cu:255-400 shaped) over the C ABI."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional
import torch
from ._lib import lib, check, SeedExtendParamsStruct
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

    def struct(self) -> SeedExtendParamsStruct:
        p = SeedExtendParamsStruct()
        p.seed_len, p.seed_interval, p.band_len, p.type = self.seed_len, self.seed_interval, self.band_len, self.type
        p.both_strands = 1 if self.both_strands else 0
        p.max_seed_hits = self.max_seed_hits
        p.dedup_jobs = 1 if self.dedup_jobs else 0
        p.scheme = self.scheme.struct()
        return p


class SeedExtendWorkspace:
    """pre-allocated outputs + temp storage for repeated calls on equally-shaped batches"""

    def __init__(self, fmi: FMIndexDevice, genome: torch.Tensor, reads: PackedStringSet, params: SeedExtendParams,
                 hit_capacity: int, keep_hits: bool = False):
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
    return lib().nvb_seed_extend(C.byref(s), _p(genome), C.byref(rd), C.c_uint32(reads.count), C.byref(ps),
                                 C.c_uint32(ws.hit_capacity), _p(ws.best_score), _p(ws.best_pos), _p(ws.n_hits),
                                 _p(ws.hit_read), _p(ws.hit_window), _p(ws.hit_score), _p(ws.hit_sink),
                                 _p(temp), C.byref(tb), C.c_void_p(torch.cuda.current_stream().cuda_stream))


def seed_extend(fmi: FMIndexDevice, genome: torch.Tensor, reads: PackedStringSet, params: SeedExtendParams,
                workspace: Optional[SeedExtendWorkspace] = None, hit_capacity: Optional[int] = None, keep_hits: bool = False):
    """returns the workspace: .best_score[n], .best_pos[n], .n_hits[2] = (kept, total), optional per-hit arrays"""
    if workspace is None:
        if hit_capacity is None:
            hit_capacity = 32 * reads.count + 1024
        workspace = SeedExtendWorkspace(fmi, genome, reads, params, hit_capacity, keep_hits)
    tb = C.c_size_t(workspace.temp_bytes)
    check(_call(fmi, genome, reads, params, workspace, workspace.temp, tb), "nvb_seed_extend")
    return workspace


STAGES = ("strings", "seed_match", "hit_slots", "locate_windows", "dedup", "extend", "reduce")


def last_stage_ms():
    """device time (ms) of the stages of the most recent seed_extend call"""
    ms = (C.c_float * 7)()
    check(lib().nvb_seed_extend_stage_ms(ms), "nvb_seed_extend_stage_ms")
    return dict(zip(STAGES, [float(v) for v in ms]))
