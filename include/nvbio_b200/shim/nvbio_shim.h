// nvbio_b200/shim/nvbio_shim.h -- the drop-in boundary on the reference's side: include this header right after nvbio's own
// (or force-include it with `nvcc -include nvbio_b200/shim/nvbio_shim.h`), add -I<nvbio_b200>/include and link
// libnvbio_b200.so.  Code written against nvbio's templates -- FMIndexFilterDevice<fm_index_type>::rank / locate,
// aln::batch_banded_alignment_score<BAND_LEN>(), aln::batch_alignment_score(), BatchedBandedAlignmentScore<...>::enact -- then
// runs on the B200 kernels without a source change; see INTEGRATION.md and tests/shim/shim_harness.cu.
#pragma once
#include <nvbio_b200/shim/views.h>
#include <nvbio_b200/shim/fmindex_filter.h>
#include <nvbio_b200/shim/batched_alignment.h>
