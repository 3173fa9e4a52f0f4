"""-m gpu: banded Gotoh traceback through the C ABI: the CIGARs asserted by the reference's own test, the oracle on
seeded batches (incl. 4-bit reads with N's and quality tables), and re-scoring of every alignment."""
import numpy as np
import pytest
import torch
from oracle import orc
import nvbio_b200 as nb
from nvbio_b200 import aln
from nvbio_b200.strings import PackedStringSet
from tests.gpu_util import require_gpu, host_u32
from tests.golden.make_golden import G1_P, G1_T, G2_P, G2_T
from tests.test_host_core import fixed_problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    require_gpu()
    return orc.Oracle()


def run(band, typ, scheme, pr, pbits=4, tbits=2):
    pat, p_off, p_len, txt, t_off, t_len = pr
    P = PackedStringSet.from_symbols(pat, p_off, p_len, bits=pbits, big_endian=True)
    T = PackedStringSet.from_symbols(txt, t_off, t_len, bits=tbits, big_endian=True)
    o = aln.batch_banded_alignment_traceback(band, aln.make_gotoh_aligner(typ, aln.SimpleGotohScheme(*scheme)), P, T, max_ops=512)
    torch.cuda.synchronize()
    return dict(score=o["score"].cpu().numpy(), sink=host_u32(o["sink"]), source=host_u32(o["source"]),
                ops=o["ops"].cpu().numpy(), n_ops=host_u32(o["n_ops"]))


def test_reference_asserted_cigars(O):
    for P, T, scheme, band, want in ((G1_P, G1_T, (2, -1, -1, -1), 7, "4M1D3M"), (G2_P, G2_T, (0, -5, -8, -3), 31, "147M2D3M")):
        p, t = orc.dna(P), orc.dna(T)
        o = run(band, aln.SEMI_GLOBAL, scheme, (p, [0], [len(p)], t, [0], [len(t)]), pbits=2)
        assert orc.rle(o["ops"][0][:o["n_ops"][0]]) == want          # the reference's END->START order
        assert aln.cigar(o["ops"][0], int(o["n_ops"][0])) == orc.rle(o["ops"][0][:o["n_ops"][0]][::-1])


@pytest.mark.parametrize("band", [7, 15, 31])
@pytest.mark.parametrize("typ", [0, 1, 2])
def test_traceback_vs_oracle(O, band, typ):
    rng = np.random.default_rng(band + 10 * typ)
    for scheme in ((2, -2, -5, -3), (2, -1, -1, -1), (0, -5, -8, -3)):
        pr = list(fixed_problems(rng, 777, band, 150, extra_text=int(rng.integers(0, 3)), ragged=True))
        pr[0] = pr[0].copy(); pr[0][rng.integers(0, len(pr[0]), 40)] = 4            # N's in the reads
        want = O.banded_traceback(band, typ, scheme, *pr)
        got = run(band, typ, scheme, pr)
        for k in ("score", "sink", "source", "n_ops", "ops"):
            assert np.array_equal(got[k], want[k]), (band, typ, scheme, k)
        # every alignment re-scores to its reported score (TestBacktracker::score, alignment_test_utils.h:645-700)
        m, x, go, ge = scheme
        pat, p_off, p_len, txt, t_off, t_len = pr
        for a in range(0, 777, 37):
            ops = got["ops"][a][:got["n_ops"][a]][::-1]
            i, j = int(got["source"][a][1]), int(got["source"][a][0])
            # GLOBAL starts from row 0 of the band, whose cell j carries the text-gap cost of skipping j symbols
            # (init_row_zero, gotoh_banded_inl.h:57-59)
            s, prev = ((go + (j - 1) * ge) if (typ == 0 and j > 0) else 0), -1
            for op in ops:
                if op == 0:
                    s += m if pat[p_off[a] + i] == txt[t_off[a] + j] else x; i += 1; j += 1
                elif op == 1:
                    s += ge if prev == 1 else go; i += 1
                else:
                    s += ge if prev == 2 else go; j += 1
                prev = op
            assert (j, i) == (int(got["sink"][a][0]), int(got["sink"][a][1]))
            assert s == int(got["score"][a]), (a, s, got["score"][a])


@pytest.mark.parametrize("typ", [1, 2])
def test_gapless_fast_path_equals_direction_matrix_traceback(O, typ):
    """reads with few indels (most alignments are resolved by the gapless fast path from the score kernels' sink): every output equals
    the oracle's traceback, and equals the direction-matrix traceback run for every alignment (nvb_debug_traceback_fast(0))"""
    import ctypes as C
    rng = np.random.default_rng(77 + typ)
    for band, scheme in ((31, (2, -2, -5, -3)), (15, (2, -6, -8, -3)), (7, (1, -1, -1, -1))):
        n, m = 1500, 120
        pats, txts, p_off, p_len, t_off, t_len = [], [], [], [], [], []
        po = to = 0
        for a in range(n):
            N = m + band
            t = rng.integers(0, 4, N).astype(np.uint8)
            j = int(rng.integers(0, band)); p = []
            indel = 0.0 if a % 3 else 0.01
            while len(p) < m:
                r = rng.random()
                if r < indel and j < N: p.append(int(rng.integers(0, 4)))
                elif r < 2 * indel: j += 1
                elif j < N:
                    c = int(t[j]); c = int(rng.integers(0, 4)) if rng.random() < 0.03 else c
                    p.append(c); j += 1
                else: p.append(int(rng.integers(0, 4)))
            pats.append(np.array(p[:m], np.uint8)); txts.append(t)
            p_off.append(po); p_len.append(m); po += m; t_off.append(to); t_len.append(N); to += N
        pr = (np.concatenate(pats), np.array(p_off, np.uint32), np.array(p_len, np.uint32), np.concatenate(txts), np.array(t_off, np.uint32), np.array(t_len, np.uint32))
        want = O.banded_traceback(band, typ, scheme, *pr)
        assert int((want["ops"].max(axis=1) == 0).sum()) > n // 3          # plenty of gap-free alignments
        fast = run(band, typ, scheme, pr)
        nb.lib().nvb_debug_traceback_fast(C.c_int(0))
        try:
            full = run(band, typ, scheme, pr)
        finally:
            nb.lib().nvb_debug_traceback_fast(C.c_int(1))
        for k in ("score", "sink", "source", "n_ops", "ops"):
            assert np.array_equal(fast[k], want[k]), (band, typ, k)
            assert np.array_equal(full[k], want[k]), (band, typ, k, "full")
