// oracle/ref_nvbwt_writer.cu -- TEST INFRASTRUCTURE: the reference's own index-file writers, nvBWT's save_bwt() / save_ssa()
// (nvBWT/nvBWT.cu:314-351), compiled from the source where it lies so that a .bwt / .sa fixture is written by the reference's code
// (tests/golden/make_golden.py --only-nvbwt -> tests/golden/nvbwt_files.npz).  nvBWT.cu is a whole program: its main() is renamed, its
// body is placed in a namespace (its global `using namespace nvbio;` otherwise makes `cuda::` ambiguous under CUDA 12), and every
// header it includes is included first, outside that namespace.  Only the two writers are ever called; the program's other
// functions reference parts of libnvbio that are not built here, so the library is loaded with lazy binding (oracle/orc.py).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <crc/crc.h>
#include <nvbio/basic/console.h>
#include <nvbio/basic/exceptions.h>
#include <nvbio/basic/bnt.h>
#include <nvbio/basic/numbers.h>
#include <nvbio/basic/timer.h>
#include <nvbio/basic/packedstream.h>
#include <nvbio/basic/thrust_view.h>
#include <nvbio/basic/dna.h>
#include <nvbio/basic/cuda/arch.h>
#include <nvbio/fmindex/bwt.h>
#include <nvbio/fasta/fasta.h>
#include <nvbio/io/fmindex/fmindex.h>
#include <nvbio/sufsort/sufsort.h>
#include "filelist.h"
#define main nvbwt_program_main
namespace nvbwt_prog {
namespace cuda = ::nvbio::cuda;
#include <nvBWT/nvBWT.cu>
}
#undef main
// arguments exactly as nvBWT's build() passes them (nvBWT.cu:394-405, 514-515): seq_words = ceil(n / 16), ssa_len = (n + SA_INT) / SA_INT,
// cumFreq[c] = number of symbols <= c
extern "C" void ref_nvbwt_save_bwt(unsigned seq_length, unsigned seq_words, unsigned primary, const unsigned* cumFreq, const unsigned* bwt, const char* name)
{ nvbwt_prog::save_bwt(seq_length, seq_words, primary, cumFreq, bwt, name); }
extern "C" void ref_nvbwt_save_ssa(unsigned seq_length, unsigned sa_intv, unsigned ssa_len, unsigned primary, const unsigned* cumFreq, const unsigned* ssa, const char* name)
{ nvbwt_prog::save_ssa(seq_length, sa_intv, ssa_len, primary, cumFreq, ssa, name); }
