"""-m gpu: the B200 kernels against the reference's OWN CUDA code recompiled for sm_100a (oracle/_ref/ref_cuda_bench, built from
/root/reference by oracle/Makefile; the prebuilt binary travels with the snapshot):
  * nvb_fm_match_approx == nvBowtie's device function detail::map<true> (nvBowtie/bowtie2/cuda/mapping_inl.h:128-220), which has no
    host build: every pushed SA range in push order, the push counts and the range sums, incl. seeds with N's;
  * nvb_banded_gotoh_score == batched_banded_alignment_score_kernel, nvb_fm_match / filter == FMIndexFilterDevice::rank / locate."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import nvbio_b200 as nb
from nvbio_b200 import synth
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool():
    require_gpu()
    spec = importlib.util.spec_from_file_location("compare_ref_cuda", os.path.join(ROOT, "tools", "compare_ref_cuda.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if not os.path.exists(m.BIN):
        pytest.skip("oracle/_ref/ref_cuda_bench not built (needs /root/reference at build time)")
    return m


@pytest.fixture(scope="module")
def index2m():
    require_gpu()
    n = 2_000_000
    gw = synth.random_genome_words(n, seed=4242)
    fmi, _ = nb.FMIndexDevice.from_text(gw, n)
    return n, gw, fmi


@pytest.mark.parametrize("L,len1", [(22, 11), (22, 0), (20, 10), (22, 22), (16, 5)])
def test_match_approx_equals_nvbowtie_map(tool, index2m, L, len1):
    n, gw, fmi = index2m
    r = tool.approx_check(fmi, gw, n, nq=20000, L=L, len1=len1, max_out=80, with_n=True, seed=100 + L + len1)
    assert "error" not in r, r
    assert r["bit_identical_counts"] and r["bit_identical_range_sums"] and r["bit_identical_ranges_in_push_order"], r
    assert r["pushes"] > 20000 * (0.3 if len1 < L else 0.2), r
    assert r["max_pushes_per_seed"] <= 80, r


def test_match_approx_equals_nvbowtie_map_repetitive(tool):
    """a genome with long repeats: wide ranges, many substitutions survive"""
    require_gpu()
    n = 600_000
    rng = np.random.default_rng(5)
    unit = rng.integers(0, 4, 3000)
    sym = np.concatenate([unit if (i % 3) else rng.integers(0, 4, 3000) for i in range(n // 3000)]).astype(np.uint8)
    from nvbio_b200.strings import pack_symbols
    gw = torch.from_numpy(pack_symbols(sym, 2, True).view(np.int32)).cuda()
    fmi, _ = nb.FMIndexDevice.from_text(gw, n)
    r = tool.approx_check(fmi, gw, n, nq=10000, L=14, len1=7, max_out=96, with_n=True, seed=9)
    assert "error" not in r, r
    assert r["bit_identical_counts"] and r["bit_identical_range_sums"] and r["bit_identical_ranges_in_push_order"], r


def _reads_4bit(gw, n, n_reads, rng, min_len=30, max_len=150, with_n=True):
    """variable-length DNA_N reads sampled from the genome (half reverse-complemented, 2% substitutions, a few N's), packed back to
    back 4 bits per symbol big-endian, plus the symbol index (n_reads + 1 offsets)"""
    from nvbio_b200.strings import unpack_symbols, pack_symbols
    gsym = unpack_symbols(gw.cpu().numpy().view(np.uint32), n)
    syms, index = [], [0]
    for r in range(n_reads):
        L = int(rng.integers(min_len, max_len + 1)) if r % 5 else int(rng.integers(8, 20))       # some reads shorter than min_read_len / seed_len
        p = int(rng.integers(0, n - L))
        s = gsym[p:p + L].copy()
        m = rng.random(L) < 0.02
        s[m] = (s[m] + 1 + rng.integers(0, 3, int(m.sum()))) % 4
        if r & 1:
            s = (3 - s)[::-1].copy()
        if with_n and r % 7 == 0:
            s[rng.integers(0, L, 2)] = 4
        syms.append(s[::-1].copy()); index.append(index[-1] + L)      # nvBowtie stores reads reversed (mapping_inl.h:263-270)
    allsym = np.concatenate(syms).astype(np.uint8)
    return allsym, np.array(index, np.uint32), pack_symbols(np.concatenate([allsym, np.zeros(16, np.uint8)]), 4, True)


@pytest.mark.parametrize("algo,seed_len,seed_freq,max_hits,subseed,retry,fw,rc",
                         [(0, 22, 10, 100, 0, 0, 1, 1), (0, 20, 7, 100, 0, 1, 1, 1), (0, 16, 5, 3, 0, 0, 1, 0),
                          (1, 22, 10, 100, 11, 0, 1, 1), (1, 16, 8, 100, 0, 2, 0, 1), (1, 20, 10, 4, 10, 0, 1, 1)])
def test_map_seeds_equals_nvbowtie_map_queues_kernel(tool, algo, seed_len, seed_freq, max_hits, subseed, retry, fw, rc):
    """nvb_map_seeds == nvBowtie's own map_queues_kernel<EXACT_MAPPING|APPROX_MAPPING> (mapping_inl.h:539-591) run on the device over the
    same DNA_N reads, index and input queue: deque sizes, reseed flags and -- per read -- the SeedHits themselves as multisets (the
    reference stores each deque in interval-heap order).  With a small max_hits (deques overflow) the multiset of (range size) and
    the sizes must still agree; which of several equally large ranges was dropped is the heap's private choice."""
    import subprocess, tempfile, json
    from nvbio_b200.strings import PackedStringSet
    require_gpu()
    n = 300_000
    rng = np.random.default_rng(1000 + algo * 100 + seed_len + max_hits)
    # a genome with repeats so that ranges of many sizes occur
    from nvbio_b200.strings import pack_symbols
    unit = rng.integers(0, 4, 500)
    sym = np.concatenate([unit if (i % 4 == 1) else rng.integers(0, 4, 500) for i in range(n // 500)]).astype(np.uint8)
    gw = torch.from_numpy(pack_symbols(np.concatenate([sym, np.zeros(64, np.uint8)]), 2, True).view(np.int32)).cuda()
    fmi, _ = nb.FMIndexDevice.from_text(gw, n)
    n_reads = 3000
    allsym, index, words = _reads_4bit(gw, n, n_reads, rng)
    queue = np.array([i for i in range(n_reads) if i % 11 != 3], np.uint32)
    rng.shuffle(queue)
    reads = PackedStringSet(words=torch.from_numpy(words.view(np.int32)).cuda(), bits=4, big_endian=True,
                            offsets=torch.from_numpy(index[:-1].astype(np.int32)).cuda(), lengths=torch.from_numpy(np.diff(index).astype(np.int32)).cuda(),
                            stride=0, length=150, count=n_reads)
    hits, counts, reseed, stats = nb.map_seeds(fmi, reads, algorithm=algo, seed_len=seed_len, seed_freq=seed_freq, max_hits=max_hits, max_reseed=2,
                                               rep_seeds=5, subseed_len=subseed, min_read_len=20, fw=bool(fw), rc=bool(rc),
                                               queue=torch.from_numpy(queue.astype(np.int32)).cuda(), retry=retry)
    torch.cuda.synchronize()
    arena = n_reads * max_hits
    with tempfile.TemporaryDirectory() as d:
        meta = [fmi.length, fmi.primary] + list(fmi.L2) + [n_reads, len(queue), algo, seed_len, seed_freq, max_hits, 2, 5, subseed, 20, retry, fw, rc, arena]
        np.array(meta, dtype=np.uint32).tofile(d + "/meta.bin")
        fmi.bwt_occ.cpu().numpy().tofile(d + "/bwt_occ.bin"); fmi.ssa.cpu().numpy().tofile(d + "/ssa.bin")
        words.tofile(d + "/read_words.bin"); index.tofile(d + "/read_index.bin"); queue.tofile(d + "/queue.bin")
        r = subprocess.run([tool.BIN, "mapq", d], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-600:]
        info = json.loads(r.stdout.strip().splitlines()[-1])
        assert info["sizeof_SeedHit"] == 8
        ref_counts = np.fromfile(d + "/ref_counts.bin", np.uint32); ref_index = np.fromfile(d + "/ref_index.bin", np.uint32)
        ref_reseed = np.fromfile(d + "/ref_reseed.bin", np.uint8); ref_hits = np.fromfile(d + "/ref_hits.bin", np.uint32).reshape(-1, 2)
    oc = counts.cpu().numpy().view(np.uint32); oh = hits.cpu().numpy().view(np.uint32)
    queued = np.zeros(n_reads, bool); queued[queue] = True
    assert np.array_equal(oc[queued], ref_counts[queued])
    lens = np.diff(index)
    # reseed flags: set for every queue entry whose read is long enough (the reference leaves the others untouched)
    long_enough = lens[queue] >= 20
    assert np.array_equal(reseed.cpu().numpy()[long_enough], ref_reseed[long_enough])
    n_exact = n_over = 0
    for rid in queue:
        c = int(ref_counts[rid])
        ref = ref_hits[ref_index[rid]:ref_index[rid] + c]
        ours = oh[rid, :c]
        sizes_ours = ours[:, 1] & 0xFFFFF
        assert np.all(np.diff(sizes_ours.astype(np.int64)) >= 0)                      # sorted by range size
        pushes = int(stats.cpu().numpy()[np.nonzero(queue == rid)[0][0], 1]) if c == max_hits else c
        if c < max_hits or pushes <= max_hits:
            assert sorted(map(tuple, ours.tolist())) == sorted(map(tuple, ref.tolist())), rid
            n_exact += 1
        else:
            assert sorted((ref[:, 1] & 0xFFFFF).tolist()) == sorted(sizes_ours.tolist()), rid
            n_over += 1
    assert n_exact > 1000
    if max_hits <= 12:
        assert n_over > 20
