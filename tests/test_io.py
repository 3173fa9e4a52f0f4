"""Index files in the reference's .bwt/.sa format: CPU round trip of the writer against a hand-parsed file, and
(GPU) loading + searching."""
import numpy as np
import pytest
from oracle import orc


def test_file_layout(tmp_path):
    from nvbio_b200 import io as nio
    O = orc.Oracle()
    text = np.random.default_rng(3).integers(0, 4, 1000).astype(np.uint8)
    idx = O.build_index(text)
    prefix = str(tmp_path / "g")
    nio.save_index(prefix, idx.bwt_occ, idx.ssa, idx.L2, idx.n, idx.primary)
    raw = np.fromfile(prefix + ".bwt", dtype=np.uint32)
    assert raw[0] == idx.primary and list(raw[1:5]) == list(idx.L2[1:]) and raw[4] == 1000
    assert np.array_equal(raw[5:], idx.bwt[:len(raw) - 5])
    sa = np.fromfile(prefix + ".sa", dtype=np.uint32)
    assert sa[0] == idx.primary and sa[5] == 16 and sa[6] == 1000
    assert np.array_equal(sa[7:], idx.ssa[1:]) and len(sa) - 7 == (1000 + 16) // 16 - 1


@pytest.mark.gpu
def test_load_and_search(tmp_path):
    import torch
    import nvbio_b200 as nb
    from nvbio_b200 import io as nio
    from nvbio_b200.strings import PackedStringSet
    from tests.gpu_util import require_gpu, host_u32, dev_u32, mask_pad
    require_gpu()
    O = orc.Oracle()
    rng = np.random.default_rng(4)
    text = rng.integers(0, 4, 77777).astype(np.uint8)
    idx = O.build_index(text)
    prefix = str(tmp_path / "g")
    nio.save_index(prefix, idx.bwt_occ, idx.ssa, idx.L2, idx.n, idx.primary)
    fmi = nio.load_index(prefix)
    assert fmi.length == idx.n and fmi.primary == idx.primary and fmi.sa_interval == 16
    assert np.array_equal(mask_pad(host_u32(fmi.bwt_occ), idx.n), mask_pad(idx.bwt_occ, idx.n))
    assert np.array_equal(host_u32(fmi.ssa), idx.ssa)
    lens = rng.integers(5, 25, 2000).astype(np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
    q = rng.integers(0, 4, int(lens.sum())).astype(np.uint8)
    for i in range(0, 2000, 2):
        st = int(rng.integers(0, idx.n - 30)); q[offs[i]:offs[i] + lens[i]] = text[st:st + lens[i]]
    want, _ = O.match(idx, q, offs, lens)
    got = host_u32(nb.match(fmi, PackedStringSet.from_symbols(q, offs, lens)))
    assert np.array_equal(got, want)
    rows = rng.integers(0, idx.n + 1, 1000).astype(np.uint32)
    assert np.array_equal(host_u32(nb.locate(fmi, dev_u32(rows))), O.locate(idx, rows))
    # a mismatching .sa is rejected like the reference's file_mismatch
    bad = np.fromfile(prefix + ".sa", dtype=np.uint32); bad[0] += 1; bad.tofile(prefix + ".sa")
    with pytest.raises(IOError):
        nio.load_index(prefix)
