#!/usr/bin/env python
"""experiment: the seed-match stage of the C3 step (1M x 150 bp, 3 Gbp, full SA) under different k-mer table formats -- one index,
one read batch, the table rebuilt per variant; per-stage device times from nvb_seed_extend_stage_ms"""
import sys, os, json, argparse, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvbio_b200 as nb
from nvbio_b200 import aln, synth
from nvbio_b200.strings import PackedStringSet
from nvbio_b200.pipeline import SeedExtendWorkspace, last_stage_ms

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=3000.0)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--variants", nargs="*", default=["16:0", "16:1", "15:1", "15:0", "14:1", "16:1"], help="k:located[:split[:perfect]] (located 0 = {x, y} entries, 1 = + SA values, 2 = + text context; split 0 = seed match in one pass; "
                "perfect 0 = every alignment job through the DP kernels)")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

n = int(args.genome_mbp * 1e6)
genome = synth.random_genome_words(n)
fmi, _ = nb.FMIndexDevice.from_text(genome, n, sa_interval=1)
torch.cuda.empty_cache()
rw, _, _ = synth.sample_reads(genome, n, args.reads, 150, sub_rate=0.01, indel_rate=0.001, seed=7, mut_seed=8)
rs = PackedStringSet.fixed(rw.reshape(-1), args.reads, 150, stride=rw.shape[1] * 16)
params = nb.SeedExtendParams(seed_len=20, seed_interval=10, band_len=31, type=aln.LOCAL, both_strands=True, max_seed_hits=100,
                             scheme=aln.SimpleGotohScheme(2, -2, -5, -3))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ref = None
for v in args.variants:
    k, located, split, perfect = (tuple(int(x) for x in v.split(":")) + (1, 1))[:4]
    nb.lib().nvb_debug_seed_split(ctypes.c_int(split)); nb.lib().nvb_debug_perfect_shortcut(ctypes.c_int(perfect))
    fmi.ktab = None; torch.cuda.empty_cache()
    if k:
        fmi.build_ktab(k, located=bool(located), text=genome if located == 2 else None)
    ws = SeedExtendWorkspace(fmi, genome, rs, params, 24 * args.reads, keep_hits=False)
    for _ in range(3):
        flush.zero_(); nb.seed_extend(fmi, genome, rs, params, workspace=ws)
    torch.cuda.synchronize()
    acc = None
    for _ in range(args.steps):
        flush.zero_(); nb.seed_extend(fmi, genome, rs, params, workspace=ws); torch.cuda.synchronize()
        st = last_stage_ms()
        acc = st if acc is None else {kk: acc[kk] + st[kk] for kk in st}
    chk = (int(ws.best_score.to(torch.int64).sum().item()), int(ws.best_pos.to(torch.int64).sum().item()), [int(x) for x in ws.n_hits.cpu()])
    if ref is None:
        ref = chk
    assert chk == ref, "results differ between table formats"
    print(json.dumps({"ktab_k": k, "located": located, "split": split, "perfect": perfect, "table_GB": round((4 ** k) * (16 if located else 8) / 1e9, 1) if k else 0,
                      "stage_ms": {kk: round(vv / args.steps, 4) for kk, vv in acc.items()}}), flush=True)
    del ws
