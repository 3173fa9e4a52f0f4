"""nvbio_b200 -- B200-native (sm_100a) implementation of nvbio's two data-parallel hot paths:
FM-index rank/match/locate (nvbio/fmindex) and batched banded Gotoh scoring (nvbio/alignment),
behind a C ABI (include/nvbio_b200.h).  This package is the host-side mirror of the reference's
interface for those paths; torch is used only for device memory, streams and torch.distributed."""
from ._lib import NvbError, lib, LIB_PATH                                        # noqa: F401
from .strings import PackedStringSet, pack_symbols, unpack_symbols               # noqa: F401
from .fmindex import (FMIndexDevice, FMIndexFilterDevice, rank, rank4, match, match_approx, locate, map_seeds, locate_init, locate_lookup, locate_sorted,  # noqa: F401
                      MAP_EXACT, MAP_APPROX, dict_rank, dict_build_occ,
                      MATCH_FORWARD_ORDER, MATCH_COMPLEMENT)
from . import aln                                                                # noqa: F401
from .pipeline import SeedExtendParams, seed_extend, StreamingSeedExtend, PairParams, seed_extend_paired     # noqa: F401
