"""Packed string sets (nvbio/basic/packedstream.h, nvbio/strings/string_set.h) as plain tensors."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional
import numpy as np
import torch
from ._lib import StringSetStruct


def pack_symbols(sym: np.ndarray, bits: int = 2, big_endian: bool = True, pad_words: int = 4) -> np.ndarray:
    """symbols (uint8) -> uint32 words with PackedStream semantics (packedstream_inl.h:336-372):
    word i/SPW, shift = BE ? 32-bits-bits*(i%SPW) : bits*(i%SPW).  `pad_words` zero words are appended so
    that kernels may over-read a couple of words past the end."""
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    if bits == 8:
        nw = (len(sym) + 3) // 4 + pad_words
        out = np.zeros(nw * 4, dtype=np.uint8)
        out[:len(sym)] = sym
        return out.view(np.uint32)
    spw = 32 // bits
    n = len(sym)
    nw = (n + spw - 1) // spw
    buf = np.zeros(nw * spw, dtype=np.uint32)
    buf[:n] = sym & ((1 << bits) - 1)
    buf = buf.reshape(nw, spw)
    k = np.arange(spw, dtype=np.uint32)
    sh = (32 - bits - bits * k) if big_endian else (bits * k)
    words = np.bitwise_or.reduce(buf << sh, axis=1).astype(np.uint32)
    return np.concatenate([words, np.zeros(pad_words, dtype=np.uint32)])


def unpack_symbols(words: np.ndarray, n: int, bits: int = 2, big_endian: bool = True) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.uint32)
    if bits == 8:
        return words.view(np.uint8)[:n].copy()
    spw = 32 // bits
    i = np.arange(n, dtype=np.int64)
    k = (i % spw).astype(np.uint32)
    sh = (32 - bits - bits * k) if big_endian else (bits * k)
    return ((words[i // spw] >> sh) & ((1 << bits) - 1)).astype(np.uint8)


def _i32(t):
    return None if t is None else t


@dataclass
class PackedStringSet:
    """String i = symbols [off_i, off_i+len_i) of `words`;  off_i = offsets[i] or i*stride,
    len_i = lengths[i] or length.  With `lengths`, `length` must hold the maximum length."""
    words: torch.Tensor                      # int32 view of the uint32 words (device)
    bits: int = 2
    big_endian: bool = True
    offsets: Optional[torch.Tensor] = None   # int32 (uint32 bit pattern)
    lengths: Optional[torch.Tensor] = None
    stride: int = 0
    length: int = 0
    count: int = 0

    def struct(self) -> StringSetStruct:
        s = StringSetStruct()
        s.d_words = self.words.data_ptr()
        s.bits = self.bits
        s.big_endian = 1 if self.big_endian else 0
        s.d_offsets = self.offsets.data_ptr() if self.offsets is not None else None
        s.d_lengths = self.lengths.data_ptr() if self.lengths is not None else None
        s.stride = self.stride
        s.length = self.length
        return s

    @staticmethod
    def from_symbols(sym: np.ndarray, offsets, lengths, bits=2, big_endian=True, device="cuda"):
        """concatenated (ragged) set from unpacked symbols"""
        words = torch.from_numpy(pack_symbols(sym, bits, big_endian).view(np.int32)).to(device)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        return PackedStringSet(words=words, bits=bits, big_endian=big_endian,
                               offsets=torch.from_numpy(offsets.view(np.int32)).to(device),
                               lengths=torch.from_numpy(lengths.view(np.int32)).to(device),
                               stride=0, length=int(lengths.max()) if len(lengths) else 0, count=len(offsets))

    @staticmethod
    def fixed(words: torch.Tensor, count: int, length: int, stride: Optional[int] = None, bits=2, big_endian=True):
        return PackedStringSet(words=words, bits=bits, big_endian=big_endian, stride=length if stride is None else stride,
                               length=length, count=count)
