#!/usr/bin/env python
"""banded Gotoh traceback throughput (GCUPS = n * M * BAND / t) next to the score-only kernels on the same batch"""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nvbio_b200 as nb
from nvbio_b200 import aln, synth
from nvbio_b200.strings import PackedStringSet

n = 50_000_000
gw = synth.random_genome_words(n)
n_al, M, W = 1_000_000, 150, 181
rw, pos, _ = synth.sample_reads(gw, n, n_al, M, rc_half=False)
P = PackedStringSet.fixed(rw.reshape(-1), n_al, M, stride=rw.shape[1] * 16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for typ in (1, 2):
    al = aln.make_gotoh_aligner(typ, aln.SimpleGotohScheme(2, -2, -5, -3))
    for band in (31, 15):
        begin = (pos - band // 2).clamp_(0)                  # the read's own diagonal sits mid-band
        T = PackedStringSet(words=gw, bits=2, big_endian=True, offsets=begin.to(torch.int32), lengths=None, stride=0, length=M + band - 1, count=n_al)
        def score(): return aln.batch_banded_alignment_score(band, al, P, T)
        def trace(): return aln.batch_banded_alignment_traceback(band, al, P, T)
        out = {}
        for name, fn in (("score", score), ("traceback", trace)):
            fn(); torch.cuda.synchronize(); best = 1e30
            for _ in range(3):
                e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            out[name + "_ms"] = round(best, 3); out[name + "_gcups"] = round(n_al * M * band / best / 1e6, 1)
        print(json.dumps(dict(type=typ, band=band, **out)), flush=True)
