#!/usr/bin/env python
"""Full-matrix Gotoh throughput sweep (GCUPS = n*M*N / t): packed pair kernel at its occupancy variants vs the int32 kernel."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nvbio_b200 as nb
from nvbio_b200 import aln
from nvbio_b200.strings import PackedStringSet

def time_ms(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

CONFIGS = ((200_000, 150, 500), (1_000_000, 150, 300), (10_000, 100, 1000), (100_000, 100, 1000), (20_000, 150, 500), (50_000, 150, 500))
if "--one" in sys.argv:
    CONFIGS = CONFIGS[:1]
for (n, M, N) in CONFIGS:
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    pw = torch.randint(-2**31, 2**31, (n, (M + 15) // 16), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    tw = torch.randint(-2**31, 2**31, (n, (N + 15) // 16), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    P = PackedStringSet.fixed(pw.reshape(-1), n, M, stride=pw.shape[1] * 16)
    T = PackedStringSet.fixed(tw.reshape(-1), n, N, stride=tw.shape[1] * 16)
    for typ in (1, 0, 2):
        al = aln.make_gotoh_aligner(typ, aln.SimpleGotohScheme(2, -2, -5, -3))
        out = {"n": n, "M": M, "N": N, "type": typ}
        nb.lib().nvb_debug_full_warp(C.c_int(2))
        for minb in (2, 3, 4):
            nb.lib().nvb_debug_full_minb(C.c_int(minb))
            ms = time_ms(lambda: aln.batch_alignment_score(al, P, T))
            out["packed_minb%d_gcups" % minb] = round(n * M * N / ms / 1e6, 1)
        nb.lib().nvb_debug_full_warp(C.c_int(1))
        ms = time_ms(lambda: aln.batch_alignment_score(al, P, T))
        out["warp_gcups"] = round(n * M * N / ms / 1e6, 1)
        nb.lib().nvb_debug_full_warp(C.c_int(0))
        nb.lib().nvb_debug_full_minb(C.c_int(0))
        nb.lib().nvb_debug_force_gotoh_path(C.c_int(1))
        ms = time_ms(lambda: aln.batch_alignment_score(al, P, T))
        nb.lib().nvb_debug_force_gotoh_path(C.c_int(0))
        out["int32_gcups"] = round(n * M * N / ms / 1e6, 1)
        if (n, M, N) == CONFIGS[0]:
            # nvBowtie's quality-dependent scheme (256 x 2 table + one base quality per pattern symbol): packed per-column-profile kernel vs int32
            qs = aln.QualityGotohScheme(match_bonus=2, mm_min=2, mm_max=6, read_gap_const=5, read_gap_coeff=3, ref_gap_const=5, ref_gap_coeff=3)
            alq = aln.make_gotoh_aligner(typ, qs)
            quals = torch.randint(0, 41, (n * pw.shape[1] * 16,), dtype=torch.uint8, device="cuda", generator=g)
            ms = time_ms(lambda: aln.batch_alignment_score(alq, P, T, quals=quals))
            out["quality_packed_gcups"] = round(n * M * N / ms / 1e6, 1)
            nb.lib().nvb_debug_force_gotoh_path(C.c_int(1))
            ms = time_ms(lambda: aln.batch_alignment_score(alq, P, T, quals=quals))
            nb.lib().nvb_debug_force_gotoh_path(C.c_int(0))
            out["quality_int32_gcups"] = round(n * M * N / ms / 1e6, 1)
        print(json.dumps(out), flush=True)
