// C++ smoke of the header mirror: the reference-asserted banded problem (alignment_test.cu:799-825) through
// nvbio_b200::aln::batch_banded_alignment_score<31>.  Built by tests/test_cpp_mirror.py with g++ (no nvcc needed).
#include <nvbio_b200/nvbio_b200.hpp>
#include <cstdio>
#include <cstring>
using namespace nvbio_b200;

static std::vector<uint32_t> pack2(const char* s) {
    const size_t n = strlen(s);
    std::vector<uint32_t> w((n + 15) / 16 + 4, 0u);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
        w[i >> 4] |= c << (30 - 2 * (i & 15));
    }
    return w;
}

int main() {
    const char* P = "TTATGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTAT";
    const char* T = "ATCGGATTCTTTCTTACTTGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTATCTCTCTCTCCATCTAT";
    device_buffer<uint32_t> dp, dt; dp.upload(pack2(P)); dt.upload(pack2(T));
    nvb_string_set ps = { dp.ptr, 2, 1, nullptr, nullptr, 0, (uint32_t)strlen(P) };
    nvb_string_set ts = { dt.ptr, 2, 1, nullptr, nullptr, 0, (uint32_t)strlen(T) };
    device_buffer<int32_t> score(1); device_buffer<nvb_uint2> sink(1); device_buffer<char> temp;
    aln::batch_banded_alignment_score<31u>(aln::make_gotoh_aligner<aln::SEMI_GLOBAL>(aln::SimpleGotohScheme(0, -5, -8, -3)), ps, ts, 1u, score.ptr, sink.ptr, temp);
    cudaDeviceSynchronize();
    const int32_t s = score.download(1)[0]; const nvb_uint2 k = sink.download(1)[0];
    printf("score %d sink (%u,%u)\n", s, k.x, k.y);
    if (!(s == -11 && k.x == 165 && k.y == 150)) return 1;      // values of the reference on this problem
    // full-matrix DP of the reference's 7 x 20 strings (alignment_test.cu:761-793), LOCAL: score 13 ending at (18,7)
    const char* P2 = "ACAACTA"; const char* T2 = "AAACACCCTAACACACTAAA";
    device_buffer<uint32_t> dp2, dt2; dp2.upload(pack2(P2)); dt2.upload(pack2(T2));
    nvb_string_set ps2 = { dp2.ptr, 2, 1, nullptr, nullptr, 0, (uint32_t)strlen(P2) };
    nvb_string_set ts2 = { dt2.ptr, 2, 1, nullptr, nullptr, 0, (uint32_t)strlen(T2) };
    aln::batch_alignment_score(aln::make_gotoh_aligner<aln::LOCAL>(aln::SimpleGotohScheme(2, -1, -1, -1)), ps2, ts2, 1u, score.ptr, sink.ptr, temp);
    cudaDeviceSynchronize();
    const int32_t s2 = score.download(1)[0]; const nvb_uint2 k2 = sink.download(1)[0];
    printf("full score %d sink (%u,%u)\n", s2, k2.x, k2.y);
    return (s2 == 13 && k2.x == 18 && k2.y == 7) ? 0 : 1;
}
