import numpy as np
import pytest
import torch


def require_gpu():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product has no CPU fallback")


def dev_u32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host_u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


def mask_pad(bwt_occ, n):
    b = np.array(bwt_occ, dtype=np.uint32).copy().reshape(-1, 8)
    w = b[:, :4].reshape(-1).copy()
    full, rem = n // 16, n % 16
    if rem:
        w[full] &= np.uint32((0xFFFFFFFF << (32 - 2 * rem)) & 0xFFFFFFFF)
        full += 1
    w[full:] = 0
    b[:, :4] = w.reshape(-1, 4)
    return b.reshape(-1)
