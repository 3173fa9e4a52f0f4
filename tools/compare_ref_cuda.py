#!/usr/bin/env python
"""BASELINE TOOL (not product): run the reference's own CUDA kernels recompiled for sm_100a
(oracle/_ref/ref_cuda_bench, built by oracle/Makefile from /root/reference) and the nvbio_b200 kernels on the
SAME inputs, compare the results bit for bit, and print both throughputs.

    python tools/compare_ref_cuda.py [--n 1000000] [--genome-mbp 100] [--seeds 1000000]

Outputs one JSON line per path (banded Gotoh LOCAL band 31 on n x 150 bp vs 181 bp windows; FM-index exact
match + locate of 22-mers)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvbio_b200 as nb                      # noqa: E402
from nvbio_b200 import aln, synth            # noqa: E402
from nvbio_b200.strings import PackedStringSet  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_bench")


def time_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def banded_compare(gw, n_g, n, M=150, reps=3):
    """banded Gotoh LOCAL band 31 (2,-2,-5,-3), n x M bp reads vs their M+31 bp genome windows: nvb_banded_gotoh_score vs the reference's
    batched_banded_alignment_score_kernel<128,1,31> (DeviceThreadScheduler) recompiled for sm_100a, same inputs, bit-compared"""
    rw, pos, _ = synth.sample_reads(gw, n_g, n, M, rc_half=False)
    begin = (pos - 15).clamp_(0)
    end = (begin + M + 31).clamp_(max=n_g)
    wpr = rw.shape[1]
    P = PackedStringSet.fixed(rw.reshape(-1), n, M, stride=wpr * 16)
    T = PackedStringSet(words=gw, bits=2, big_endian=True, offsets=begin.to(torch.int32), lengths=(end - begin).to(torch.int32),
                        stride=0, length=M + 31, count=n)
    al = aln.make_gotoh_aligner(aln.LOCAL, aln.SimpleGotohScheme(2, -2, -5, -3))
    out = (torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty((n, 2), dtype=torch.int32, device="cuda"))
    temp = torch.empty(aln.banded_temp_bytes(31, al, P, T) + 256, dtype=torch.uint8, device="cuda")
    ours_ms = time_ms(lambda: aln.batch_banded_alignment_score(31, al, P, T, out=out, temp=temp))
    with tempfile.TemporaryDirectory() as d:
        np.array([n, M, wpr * 16, 2, -2 & 0xFFFFFFFF, -5 & 0xFFFFFFFF, -3 & 0xFFFFFFFF, reps], dtype=np.uint32).tofile(d + "/meta.bin")
        rw.cpu().numpy().tofile(d + "/pat_words.bin")
        gw.cpu().numpy().tofile(d + "/genome_words.bin")
        torch.stack([begin, end], dim=1).to(torch.int32).cpu().numpy().tofile(d + "/windows.bin")
        r = subprocess.run([BIN, "banded", d], capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        ref = json.loads(r.stdout.strip().splitlines()[-1])
        ref_scores = np.fromfile(d + "/ref_scores.bin", dtype=np.int32)
        ref_sinks = np.fromfile(d + "/ref_sinks.bin", dtype=np.uint32).reshape(-1, 2)
    same = bool(np.array_equal(out[0].cpu().numpy(), ref_scores) and np.array_equal(out[1].cpu().numpy().view(np.uint32), ref_sinks))
    gcups = n * M * 31 / (ours_ms * 1e-3) / 1e9
    return {"path": "banded Gotoh LOCAL band 31, %d x %d bp, (2,-2,-5,-3)" % (n, M), "nvbio_b200_ms": ours_ms, "nvbio_b200_gcups": gcups,
            "reference_cuda_sm100a_ms": ref["ms"], "reference_cuda_sm100a_gcups": ref["gcups"], "speedup": ref["ms"] / ours_ms,
            "bit_identical_scores_and_sinks": same}


def fm_compare(fmi, gw, n_g, nq, L=22):
    """FM-index exact match + locate of nq L-mers: nvb_fm_match / filter vs the reference's FMIndexFilterDevice::rank / locate
    (thrust::transform of nvbio::match / locate) recompiled for sm_100a, same index and seeds, bit-compared"""
    sw, spos = synth.sample_seeds(gw, n_g, nq, L)
    q = PackedStringSet.fixed(sw.reshape(-1), nq, L, stride=32)
    ranges = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    ours_ms = time_ms(lambda: nb.match(fmi, q, out=ranges))
    flt = nb.FMIndexFilterDevice()
    n_hits = flt.rank(fmi, q)
    hits = torch.empty((n_hits, 2), dtype=torch.int32, device="cuda")
    loc_ms = time_ms(lambda: flt.locate(0, n_hits, hits))
    with tempfile.TemporaryDirectory() as d:
        meta = [fmi.length, fmi.primary] + list(fmi.L2) + [nq, L, 32, 3]
        np.array(meta, dtype=np.uint32).tofile(d + "/meta.bin")
        fmi.bwt_occ.cpu().numpy().tofile(d + "/bwt_occ.bin")
        fmi.ssa.cpu().numpy().tofile(d + "/ssa.bin")
        sw.cpu().numpy().tofile(d + "/seed_words.bin")
        r = subprocess.run([BIN, "fm", d], capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        ref = json.loads(r.stdout.strip().splitlines()[-1])
        ref_ranges = np.fromfile(d + "/ref_ranges.bin", dtype=np.uint32).reshape(-1, 2)
        ref_hits = np.fromfile(d + "/ref_hits.bin", dtype=np.uint32).reshape(-1, 2)
    same_r = bool(np.array_equal(ranges.cpu().numpy().view(np.uint32), ref_ranges))
    same_h = bool(np.array_equal(hits.cpu().numpy().view(np.uint32)[:len(ref_hits)], ref_hits))
    return {"path": "FM-index exact match, %d x %d bp seeds, %.0f Mbp genome" % (nq, L, n_g / 1e6), "nvbio_b200_match_ms": ours_ms,
            "nvbio_b200_mseeds_s": nq / (ours_ms * 1e-3) / 1e6, "reference_cuda_sm100a_rank_ms": ref["rank_ms"],
            "reference_cuda_sm100a_mseeds_s": ref["mseeds_per_s"], "speedup_match": ref["rank_ms"] / ours_ms,
            "nvbio_b200_locate_ms": loc_ms, "reference_cuda_sm100a_locate_ms": ref["locate_ms"], "speedup_locate": ref["locate_ms"] / loc_ms,
            "n_hits": n_hits, "bit_identical_ranges": same_r, "bit_identical_hits": same_h}


def full_compare(nf, Mf=150, Nf=500):
    """full-matrix Gotoh LOCAL (sw-benchmark shape): nvb_gotoh_score vs the reference's batched_alignment_score_kernel"""
    from nvbio_b200.strings import unpack_symbols, pack_symbols
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    tsw = (Nf + 15) // 16
    tw = torch.randint(-2**31, 2**31, (nf, tsw), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    tsym = unpack_symbols(tw.cpu().numpy().reshape(-1).view(np.uint32), nf * tsw * 16, 2, True).reshape(nf, tsw * 16)
    rng = np.random.default_rng(3)
    st = rng.integers(0, Nf - Mf, nf)
    psym = np.stack([tsym[i, st[i]:st[i] + Mf] for i in range(nf)])
    mut = rng.random(psym.shape) < 0.03
    psym = np.where(mut, rng.integers(0, 4, psym.shape), psym).astype(np.uint8)
    psw = (Mf + 15) // 16
    pfull = np.zeros((nf, psw * 16), np.uint8); pfull[:, :Mf] = psym
    pw = torch.from_numpy(pack_symbols(pfull.reshape(-1), 2, True).view(np.int32)).cuda()
    Pf = PackedStringSet.fixed(pw, nf, Mf, stride=psw * 16)
    Tf = PackedStringSet.fixed(tw.reshape(-1), nf, Nf, stride=tsw * 16)
    al = aln.make_gotoh_aligner(aln.LOCAL, aln.SimpleGotohScheme(2, -2, -5, -3))
    res = [None]

    def go():
        res[0] = aln.batch_alignment_score(al, Pf, Tf)
    ours_ms = time_ms(go, reps=3)
    with tempfile.TemporaryDirectory() as d:
        np.array([nf, Mf, psw * 16, Nf, tsw * 16, 2, -2 & 0xFFFFFFFF, -5 & 0xFFFFFFFF, -3 & 0xFFFFFFFF, 2], dtype=np.uint32).tofile(d + "/meta.bin")
        pw.cpu().numpy().tofile(d + "/pat_words.bin")
        tw.cpu().numpy().tofile(d + "/txt_words.bin")
        r = subprocess.run([BIN, "full", d], capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        ref = json.loads(r.stdout.strip().splitlines()[-1])
        ref_scores = np.fromfile(d + "/ref_scores.bin", dtype=np.int32)
        ref_sinks = np.fromfile(d + "/ref_sinks.bin", dtype=np.uint32).reshape(-1, 2)
    same = bool(np.array_equal(res[0][0].cpu().numpy(), ref_scores) and np.array_equal(res[0][1].cpu().numpy().view(np.uint32), ref_sinks))
    return {"path": "full-matrix Gotoh LOCAL, %d x (%d bp vs %d bp), (2,-2,-5,-3)" % (nf, Mf, Nf), "nvbio_b200_ms": ours_ms,
            "nvbio_b200_gcups": nf * Mf * Nf / (ours_ms * 1e-3) / 1e9, "reference_cuda_sm100a_ms": ref["ms"],
            "reference_cuda_sm100a_gcups": ref["gcups"], "speedup": ref["ms"] / ours_ms, "bit_identical_scores_and_sinks": same}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--genome-mbp", type=float, default=100.0)
    ap.add_argument("--seeds", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--full-n", type=int, default=200_000)
    ap.add_argument("--approx", action="store_true", help="also compare nvb_fm_match_approx with nvBowtie's map<true> (device harness)")
    args = ap.parse_args()
    if not os.path.exists(BIN):
        print(json.dumps({"unavailable": "oracle/_ref/ref_cuda_bench not built"})); return
    n_g = int(args.genome_mbp * 1e6)
    gw = synth.random_genome_words(n_g)
    fmi, _ = nb.FMIndexDevice.from_text(gw, n_g)
    print(json.dumps(banded_compare(gw, n_g, args.n, args.read_len)), flush=True)
    if args.full_n:
        print(json.dumps(full_compare(args.full_n)), flush=True)
    if args.approx:
        print(json.dumps(approx_check(fmi, gw, n_g)), flush=True)
    print(json.dumps(fm_compare(fmi, gw, n_g, args.seeds)), flush=True)


def approx_check(fmi, gw, n_g, nq=20000, L=22, len1=11, max_out=64, with_n=True, seed=11):
    """nvb_fm_match_approx vs nvBowtie's own detail::map<true> (a __device__ function; oracle/_ref/ref_cuda_bench approx):
    every pushed range in push order, push counts and range sums, on seeds with 0 / 1 / 2 substitutions and (with_n) seeds
    carrying one N in the inexact region, one N in the exact region, or two N's.  Returns the comparison as a dict."""
    from nvbio_b200.strings import unpack_symbols
    sw, spos = synth.sample_seeds(gw, n_g, nq, L)
    sym = np.stack([unpack_symbols(sw[i].cpu().numpy().view(np.uint32), L) for i in range(nq)])
    rng = np.random.default_rng(seed)
    for k in (1, 2):                                   # a third of the seeds with one, a third with two substitutions
        rows = np.arange(k - 1, nq, 3)
        cols = rng.integers(0, L, len(rows))
        sym[rows, cols] = (sym[rows, cols] + 1 + rng.integers(0, 3, len(rows))) % 4
    if with_n:
        rows = np.arange(5, nq, 17); sym[rows, rng.integers(min(len1, L - 1), L, len(rows))] = 4     # an N in the inexact region (or the last symbol)
        if len1 > 0:
            rows = np.arange(7, nq, 41); sym[rows, rng.integers(0, len1, len(rows))] = 4             # an N in the exact region: no hits
        rows = np.arange(11, nq, 53); sym[rows, L - 1] = 4; sym[rows, L - 2] = 4                      # two N's: no hits
    # the mapper CONSUMES the seed front to back while the backward search prepends: like nvBowtie (whose reads are stored
    # reversed, mapping_inl.h:292-309) hand it the seed reversed, so that the consumed order spells the text right to left
    sym = np.ascontiguousarray(sym[:, ::-1])
    q = PackedStringSet.from_symbols(sym.reshape(-1), np.arange(nq, dtype=np.uint32) * L, np.full(nq, L, np.uint32), bits=4, big_endian=True)
    ranges, counts, sums = nb.match_approx(fmi, q, exact_len=len1, find_exact=True, max_out=max_out, flags=1)     # NVB_MATCH_FORWARD_ORDER
    torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as d:
        meta = [fmi.length, fmi.primary] + list(fmi.L2) + [nq, L, len1, max_out]
        np.array(meta, dtype=np.uint32).tofile(d + "/meta.bin")
        fmi.bwt_occ.cpu().numpy().tofile(d + "/bwt_occ.bin")
        fmi.ssa.cpu().numpy().tofile(d + "/ssa.bin")
        sym.astype(np.uint8).tofile(d + "/seed_bytes.bin")
        r = subprocess.run([BIN, "approx", d], capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        ref_ranges = np.fromfile(d + "/ref_ranges.bin", dtype=np.uint32).reshape(nq, max_out, 2)
        ref_counts = np.fromfile(d + "/ref_counts.bin", dtype=np.uint32)
        ref_sums = np.fromfile(d + "/ref_sums.bin", dtype=np.uint32)
    oc = counts.cpu().numpy().view(np.uint32); osum = sums.cpu().numpy().view(np.uint32); orng = ranges.cpu().numpy().view(np.uint32)
    same_counts = bool(np.array_equal(oc, ref_counts)); same_sums = bool(np.array_equal(osum, ref_sums))
    k = np.minimum(ref_counts, max_out)
    mask = np.arange(max_out)[None, :] < k[:, None]
    same_ranges = bool(np.array_equal(orng.reshape(nq, max_out, 2)[mask], ref_ranges[mask]))
    return {"path": "one-mismatch seed search, %d x %d bp (exact region %d), nvBowtie detail::map<true>" % (nq, L, len1),
            "seeds_with_hits": int((ref_counts > 0).sum()), "pushes": int(ref_counts.sum()), "max_pushes_per_seed": int(ref_counts.max()),
            "bit_identical_counts": same_counts, "bit_identical_range_sums": same_sums, "bit_identical_ranges_in_push_order": same_ranges}


if __name__ == "__main__":
    main()
