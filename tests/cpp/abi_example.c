/* Plain C program over the C ABI (include/nvbio_b200.h): the reference-asserted banded problem (alignment_test.cu:799-825) and the
 * reference's 7 x 20 full-matrix strings (alignment_test.cu:761-793).  Built by tests/test_cabi_example.py with gcc -- no nvcc, no
 * C++, no torch: what a foreign-language binding (cgo, JNI, ctypes) sees. */
#include <nvbio_b200.h>
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(e) do { int e_ = (int)(e); if (e_ != 0) { fprintf(stderr, "%s failed: %s\n", #e, nvb_error_string(e_)); return 2; } } while (0)

/* 2-bit big-endian packing (nvbio PackedStream<uint32*,uint8,2,true>) on the device */
static uint32_t* upload_dna(const char* s)
{
    const size_t n = strlen(s), nw = (n + 15) / 16 + 4;
    uint32_t* h = (uint32_t*)calloc(nw, sizeof(uint32_t));
    uint32_t* d = NULL;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = s[i] == 'A' ? 0u : s[i] == 'C' ? 1u : s[i] == 'G' ? 2u : 3u;
        h[i >> 4] |= c << (30 - 2 * (i & 15));
    }
    if (cudaMalloc((void**)&d, nw * sizeof(uint32_t)) != cudaSuccess) return NULL;
    cudaMemcpy(d, h, nw * sizeof(uint32_t), cudaMemcpyHostToDevice);
    free(h);
    return d;
}

static int run(int band, int type, const nvb_gotoh_scheme* scheme, const char* P, const char* T, int32_t* score, nvb_uint2* sink)
{
    nvb_string_set ps, ts;
    int32_t* d_score = NULL; nvb_uint2* d_sink = NULL; void* d_temp = NULL; size_t bytes = 0; int r;
    memset(&ps, 0, sizeof(ps)); memset(&ts, 0, sizeof(ts));
    ps.d_words = upload_dna(P); ps.bits = 2; ps.big_endian = 1; ps.length = (uint32_t)strlen(P);
    ts.d_words = upload_dna(T); ts.bits = 2; ts.big_endian = 1; ts.length = (uint32_t)strlen(T);
    if (!ps.d_words || !ts.d_words) return 3;
    cudaMalloc((void**)&d_score, sizeof(int32_t)); cudaMalloc((void**)&d_sink, sizeof(nvb_uint2));
    /* first call: the scratch size (NVB_E_TEMP_SIZE); second call: the work */
    r = band ? nvb_banded_gotoh_score(band, type, scheme, &ps, NULL, &ts, 1u, d_score, d_sink, NULL, &bytes, NULL)
             : nvb_gotoh_score(type, scheme, &ps, NULL, &ts, 1u, d_score, d_sink, NULL, &bytes, NULL);
    if (r != NVB_E_TEMP_SIZE) return r ? r : 4;
    cudaMalloc(&d_temp, bytes);
    CHECK(band ? nvb_banded_gotoh_score(band, type, scheme, &ps, NULL, &ts, 1u, d_score, d_sink, d_temp, &bytes, NULL)
               : nvb_gotoh_score(type, scheme, &ps, NULL, &ts, 1u, d_score, d_sink, d_temp, &bytes, NULL));
    cudaMemcpy(score, d_score, sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaMemcpy(sink, d_sink, sizeof(nvb_uint2), cudaMemcpyDeviceToHost);
    cudaFree(d_temp); cudaFree(d_sink); cudaFree(d_score); cudaFree((void*)ps.d_words); cudaFree((void*)ts.d_words);
    return 0;
}

int main(void)
{
    const char* P = "TTATGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTAT";
    const char* T = "ATCGGATTCTTTCTTACTTGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTATCTCTCTCTCCATCTAT";
    nvb_gotoh_scheme s1, s2;
    int32_t score = 0; nvb_uint2 sink = { 0u, 0u };
    memset(&s1, 0, sizeof(s1)); memset(&s2, 0, sizeof(s2));
    s1.match = 0; s1.mismatch = -5; s1.pattern_gap_open = s1.text_gap_open = -8; s1.pattern_gap_ext = s1.text_gap_ext = -3;   /* SimpleGotohScheme(0,-5,-8,-3) */
    CHECK(run(31, NVB_SEMI_GLOBAL, &s1, P, T, &score, &sink));
    printf("score %d sink (%u,%u)\n", score, sink.x, sink.y);
    if (!(score == -11 && sink.x == 165 && sink.y == 150)) return 1;          /* the reference's values on this problem */
    s2.match = 2; s2.mismatch = -1; s2.pattern_gap_open = s2.text_gap_open = -1; s2.pattern_gap_ext = s2.text_gap_ext = -1;   /* SimpleGotohScheme(2,-1,-1,-1) */
    CHECK(run(0, NVB_LOCAL, &s2, "ACAACTA", "AAACACCCTAACACACTAAA", &score, &sink));
    printf("full score %d sink (%u,%u)\n", score, sink.x, sink.y);
    return (score == 13 && sink.x == 18 && sink.y == 7) ? 0 : 1;
}
