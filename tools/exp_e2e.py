#!/usr/bin/env python
"""experiment: where does the host-to-host pipeline lose time against the device-resident step?"""
import sys, os, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvbio_b200 as nb
from nvbio_b200 import aln, synth
from nvbio_b200.strings import PackedStringSet

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 3_000_000_000
n_reads = 1_000_000
dev = torch.device("cuda", 0)
genome = synth.random_genome_words(n, device=dev)
fmi, _ = nb.FMIndexDevice.from_text(genome, n, sa_interval=1)
torch.cuda.empty_cache()
fmi.build_ktab(16 if n > 1e9 else 12, located=True, text=genome)
params = nb.SeedExtendParams()
batches = [synth.sample_reads(genome, n, n_reads, 150, sub_rate=0.01, indel_rate=0.001, device=dev, seed=5 + b, mut_seed=9 + b)[0].contiguous() for b in range(2)]
wpr = batches[0].shape[1]
host = [b.cpu().pin_memory() for b in batches]
# raw copies
d = torch.empty_like(batches[0]); o = torch.empty(2_000_064, dtype=torch.int32, device=dev); ho = torch.empty(2_000_064, dtype=torch.int32).pin_memory()
for _ in range(3):
    d.copy_(host[0], non_blocking=True); ho.copy_(o, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(host[0], non_blocking=True)
torch.cuda.synchronize(); h2d_ms = (time.perf_counter() - t0) * 50
t0 = time.perf_counter()
for _ in range(20):
    ho.copy_(o, non_blocking=True)
torch.cuda.synchronize(); d2h_ms = (time.perf_counter() - t0) * 50
print(json.dumps({"h2d_40MB_ms": h2d_ms, "d2h_8MB_ms": d2h_ms}))
# device-resident step
rs = PackedStringSet.fixed(batches[0].reshape(-1), n_reads, 150, stride=wpr * 16)
from nvbio_b200.pipeline import SeedExtendWorkspace
ws = SeedExtendWorkspace(fmi, genome, rs, params, 24 * n_reads)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    nb.seed_extend(fmi, genome, rs, params, workspace=ws)
e0.record()
for _ in range(20):
    nb.seed_extend(fmi, genome, rs, params, workspace=ws)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"device_step_ms_back_to_back_no_flush": e0.elapsed_time(e1) / 20}))
del ws
for depth, streams in ((1, 1), (2, 1), (3, 1), (4, 1), (2, 2)):
    os.environ["NVB_PIPELINE_COMPUTE_STREAMS"] = str(streams)
    st = nb.StreamingSeedExtend(fmi, genome, params, n_reads, 150, wpr, hit_capacity=24 * n_reads, depth=depth)
    def run(k):
        q, dm = [], []
        for i in range(k):
            q.append(st.submit(host[i & 1]))
            if len(q) == depth:
                st.result(q.pop(0)); dm.append(st.last_device_ms)
        while q:
            st.result(q.pop(0)); dm.append(st.last_device_ms)
        return dm
    run(6)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dm = run(60)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3 / 60
    print(json.dumps({"depth": depth, "compute_streams": streams, "wall_ms_per_step": wall, "Mreads_s": n_reads / wall / 1e3,
                      "device_ms_per_batch_mean": sum(dm) / len(dm), "device_ms_min": min(dm), "device_ms_max": max(dm)}))
    st.close()
