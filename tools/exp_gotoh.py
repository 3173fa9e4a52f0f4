#!/usr/bin/env python
"""experiment: banded pair kernel, compile-time pattern format (PFMT 2/4) vs the run-time-format kernel, per band"""
import sys, os, json, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvbio_b200 as nb
from nvbio_b200 import aln, synth
from nvbio_b200.strings import PackedStringSet

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--n-al", type=int, default=4_000_000)
ap.add_argument("--bands", type=int, nargs="*", default=[7, 15, 31])
ap.add_argument("--fmts", type=int, nargs="*", default=[1, 0], help="1 = compile-time pattern format kernels, 0 = run-time format")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--types", type=int, nargs="*", default=[1], help="0 GLOBAL, 1 LOCAL, 2 SEMI_GLOBAL")
ap.add_argument("--rows2", type=int, nargs="*", default=[1, 0], help="1 = two pattern rows in flight per thread (LOCAL), 0 = one")
ap.add_argument("--extra-smem", type=int, nargs="*", default=[0], help="unused dynamic shared memory per CTA (occupancy experiment)")
ap.add_argument("--check-file", default=None, help="JSON of result checksums shared between runs (e.g. of different library builds)")
args = ap.parse_args()
n = 100_000_000
gw = synth.random_genome_words(n)
n_al, M, W = args.n_al, 151, 300
rw, pos, _ = synth.sample_reads(gw, n, n_al, M, rc_half=False)
begin = synth.windows_for_reads(n, pos, M, W)
P = PackedStringSet.fixed(rw.reshape(-1), n_al, M, stride=rw.shape[1] * 16)
T = PackedStringSet(words=gw, bits=2, big_endian=True, offsets=begin.to(torch.int32), lengths=None, stride=0, length=W, count=n_al)
res = (torch.empty(n_al, dtype=torch.int32, device="cuda"), torch.empty((n_al, 2), dtype=torch.int32, device="cuda"))
L = nb.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = {}
if args.check_file and os.path.exists(args.check_file):
    ref = {tuple(int(v) for v in k.split(',')): c for k, c in json.load(open(args.check_file)).items()}
for fmt, rows2, xs in [(f, r2, x) for f in args.fmts for r2 in args.rows2 for x in args.extra_smem]:
    L.nvb_debug_pair_format(C.c_int(fmt)); L.nvb_debug_pair_rows2(C.c_int(rows2)); L.nvb_debug_pair_extra_smem(C.c_int(xs))
    for band, typ in [(b, t) for t in args.types for b in args.bands]:
        al = aln.make_gotoh_aligner(typ, aln.SimpleGotohScheme(2, -2, -5, -3))
        temp = torch.empty(aln.banded_temp_bytes(band, al, P, T) + 256, dtype=torch.uint8, device="cuda")
        aln.batch_banded_alignment_score(band, al, P, T, out=res, temp=temp); torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            e0.record(); aln.batch_banded_alignment_score(band, al, P, T, out=res, temp=temp); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        key = (band, typ)
        chk = int(res[0].to(torch.int64).sum().item())
        if key in ref: assert ref[key] == chk, "results differ between the two kernels"
        ref[key] = chk
        print(json.dumps({"pfmt_kernels": bool(fmt), "rows2": bool(rows2), "extra_smem": xs, "type": typ, "band": band, "ms": best, "GCUPS": n_al * M * band / best / 1e6}))
if args.check_file:
    json.dump({"%d,%d" % k: c for k, c in ref.items()}, open(args.check_file, "w"))
