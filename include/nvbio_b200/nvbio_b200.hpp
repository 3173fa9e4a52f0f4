// nvbio_b200.hpp -- thin C++ mirror of the nvbio interfaces this library replaces, over the C ABI
// (include/nvbio_b200.h).  Header-only; needs only the CUDA runtime for device memory.  Names, argument
// meaning and results follow the reference:
//   nvbio::FMIndexFilter<device_tag,...>::rank / locate / n_hits / ranges / ranks   (nvbio/fmindex/filter.h:145-214)
//   nvbio::aln::SimpleGotohScheme, make_gotoh_aligner<TYPE>, BestSink<int32>         (nvbio/alignment/utils.h:114-135,
//                                                                                     alignment_base.h:255-298, sink.h:69-93)
//   nvbio::aln::batch_banded_alignment_score<BAND_LEN>                               (nvbio/alignment/batched_inl.h:1067-1101)
//   nvbio::aln::batch_alignment_score (full DP)                                      (nvbio/alignment/batched_inl.h:984-1040)
// Errors are thrown as std::runtime_error carrying nvb_error_string(), the analogue of the reference's
// cuda::check_error exceptions.
#pragma once
#include "../nvbio_b200.h"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace nvbio_b200 {

inline void check(int err, const char* what) {
    if (err != NVB_OK) throw std::runtime_error(std::string(what) + ": " + nvb_error_string(err));
}

// minimal owning device buffer
template <typename T>
struct device_buffer {
    T* ptr = nullptr; size_t count = 0;
    device_buffer() = default;
    explicit device_buffer(size_t n) { resize(n); }
    device_buffer(const device_buffer&) = delete;
    device_buffer& operator=(const device_buffer&) = delete;
    ~device_buffer() { if (ptr) cudaFree(ptr); }
    void resize(size_t n) {
        if (n <= count) return;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; count = 0;
        if (n) { check((int)cudaMalloc((void**)&ptr, n * sizeof(T)), "cudaMalloc"); count = n; }
    }
    void upload(const std::vector<T>& h) { resize(h.size()); check((int)cudaMemcpy(ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "H2D"); }
    std::vector<T> download(size_t n) const { std::vector<T> h(n); check((int)cudaMemcpy(h.data(), ptr, n * sizeof(T), cudaMemcpyDeviceToHost), "D2H"); return h; }
};

/// FMIndexFilter<device_tag, fm_index_type>
struct FMIndexFilterDevice {
    typedef nvb_uint2 range_type;
    typedef nvb_uint2 hit_type;          // (text position, query id)

    /// rank(index, string_set) -> total number of hits
    uint64_t rank(const nvb_fm_index& index, const nvb_string_set& strings, uint32_t n_queries, cudaStream_t stream = 0) {
        m_index = index; m_n_queries = n_queries;
        m_ranges.resize(n_queries); m_slots.resize(n_queries);
        size_t tb = 0; uint64_t n_hits = 0;
        int r = nvb_fm_filter_rank(&index, &strings, n_queries, 0u, m_ranges.ptr, m_slots.ptr, &n_hits, nullptr, &tb, stream);
        if (r != NVB_E_TEMP_SIZE) check(r, "nvb_fm_filter_rank");
        m_temp.resize(tb ? tb : 1);
        check(nvb_fm_filter_rank(&index, &strings, n_queries, 0u, m_ranges.ptr, m_slots.ptr, &n_hits, m_temp.ptr, &tb, stream), "nvb_fm_filter_rank");
        return m_n_occurrences = n_hits;
    }
    /// locate(begin, end, hits)
    void locate(uint64_t begin, uint64_t end, hit_type* d_hits, cudaStream_t stream = 0) {
        check(nvb_fm_filter_locate(&m_index, m_ranges.ptr, m_slots.ptr, m_n_queries, begin, end, d_hits, stream), "nvb_fm_filter_locate");
    }
    uint64_t          n_hits() const { return m_n_occurrences; }
    const range_type* ranges() const { return m_ranges.ptr; }
    const uint64_t*   ranks()  const { return m_slots.ptr; }

    nvb_fm_index m_index{}; uint32_t m_n_queries = 0; uint64_t m_n_occurrences = 0;
    device_buffer<nvb_uint2> m_ranges; device_buffer<uint64_t> m_slots; device_buffer<char> m_temp;
};

namespace aln {

enum AlignmentType { GLOBAL = NVB_GLOBAL, LOCAL = NVB_LOCAL, SEMI_GLOBAL = NVB_SEMI_GLOBAL };

struct SimpleGotohScheme {
    SimpleGotohScheme(int32_t match, int32_t mm, int32_t gap_open, int32_t gap_ext) : m_match(match), m_mismatch(mm), m_gap_open(gap_open), m_gap_ext(gap_ext) {}
    int32_t m_match, m_mismatch, m_gap_open, m_gap_ext;
    nvb_gotoh_scheme abi() const { nvb_gotoh_scheme s = { m_match, m_mismatch, m_gap_open, m_gap_ext, m_gap_open, m_gap_ext, nullptr, 0, 0 }; return s; }
};
template <AlignmentType TYPE, typename scheme_type> struct GotohAligner { scheme_type scheme; };
template <AlignmentType TYPE, typename scheme_type>
GotohAligner<TYPE, scheme_type> make_gotoh_aligner(const scheme_type& scheme) { return GotohAligner<TYPE, scheme_type>{ scheme }; }

/// batch_banded_alignment_score<BAND_LEN>(aligner, patterns, texts, scores, sinks): BestSink<int32> as SoA
template <uint32_t BAND_LEN, AlignmentType TYPE, typename scheme_type>
void batch_banded_alignment_score(const GotohAligner<TYPE, scheme_type> aligner, const nvb_string_set& patterns, const nvb_string_set& texts,
                                  uint32_t n, int32_t* d_scores, nvb_uint2* d_sinks, device_buffer<char>& temp, cudaStream_t stream = 0)
{
    const nvb_gotoh_scheme s = aligner.scheme.abi();
    size_t tb = 0;
    int r = nvb_banded_gotoh_score(BAND_LEN, TYPE, &s, &patterns, nullptr, &texts, n, d_scores, d_sinks, nullptr, &tb, stream);   // min_temp_storage
    if (r != NVB_E_TEMP_SIZE) check(r, "nvb_banded_gotoh_score");
    temp.resize(tb ? tb : 1);
    check(nvb_banded_gotoh_score(BAND_LEN, TYPE, &s, &patterns, nullptr, &texts, n, d_scores, d_sinks, temp.ptr, &tb, stream), "nvb_banded_gotoh_score");
}

/// batch_alignment_score(aligner, patterns, texts, scores, sinks): the full-matrix DP (nvbio/alignment/batched_inl.h:984-1040)
template <AlignmentType TYPE, typename scheme_type>
void batch_alignment_score(const GotohAligner<TYPE, scheme_type> aligner, const nvb_string_set& patterns, const nvb_string_set& texts,
                           uint32_t n, int32_t* d_scores, nvb_uint2* d_sinks, device_buffer<char>& temp, cudaStream_t stream = 0)
{
    const nvb_gotoh_scheme s = aligner.scheme.abi();
    size_t tb = 0;
    int r = nvb_gotoh_score(TYPE, &s, &patterns, nullptr, &texts, n, d_scores, d_sinks, nullptr, &tb, stream);
    if (r != NVB_E_TEMP_SIZE) check(r, "nvb_gotoh_score");
    temp.resize(tb ? tb : 1);
    check(nvb_gotoh_score(TYPE, &s, &patterns, nullptr, &texts, n, d_scores, d_sinks, temp.ptr, &tb, stream), "nvb_gotoh_score");
}

/// the alignment of a traceback call, SoA over the batch: BestSink + Alignment<int32>::source + the backtracer's op stream
struct TracebackArrays { int32_t* d_scores; nvb_uint2* d_sinks; nvb_uint2* d_sources; uint8_t* d_ops; uint32_t max_ops; uint32_t* d_n_ops; };

/// batch form of aln::banded_alignment_traceback<BAND_LEN,...> (nvbio/alignment/banded_inl.h:352-489)
template <uint32_t BAND_LEN, AlignmentType TYPE, typename scheme_type>
void batch_banded_alignment_traceback(const GotohAligner<TYPE, scheme_type> aligner, const nvb_string_set& patterns, const nvb_string_set& texts,
                                      uint32_t n, const TracebackArrays& out, device_buffer<char>& temp, cudaStream_t stream = 0)
{
    const nvb_gotoh_scheme s = aligner.scheme.abi();
    size_t tb = 0;
    int r = nvb_banded_gotoh_traceback(BAND_LEN, TYPE, &s, &patterns, nullptr, &texts, n, out.d_scores, out.d_sinks, out.d_sources, out.d_ops, out.max_ops,
                                       out.d_n_ops, nullptr, &tb, stream);
    if (r != NVB_E_TEMP_SIZE) check(r, "nvb_banded_gotoh_traceback");
    temp.resize(tb ? tb : 1);
    check(nvb_banded_gotoh_traceback(BAND_LEN, TYPE, &s, &patterns, nullptr, &texts, n, out.d_scores, out.d_sinks, out.d_sources, out.d_ops, out.max_ops,
                                     out.d_n_ops, temp.ptr, &tb, stream), "nvb_banded_gotoh_traceback");
}

/// batch form of aln::alignment_traceback (full matrix; nvbio/alignment/alignment_inl.h:365-530)
template <AlignmentType TYPE, typename scheme_type>
void batch_alignment_traceback(const GotohAligner<TYPE, scheme_type> aligner, const nvb_string_set& patterns, const nvb_string_set& texts,
                               uint32_t n, const TracebackArrays& out, device_buffer<char>& temp, cudaStream_t stream = 0)
{
    const nvb_gotoh_scheme s = aligner.scheme.abi();
    size_t tb = 0;
    int r = nvb_gotoh_traceback(TYPE, &s, &patterns, nullptr, &texts, n, out.d_scores, out.d_sinks, out.d_sources, out.d_ops, out.max_ops, out.d_n_ops, nullptr, &tb, stream);
    if (r != NVB_E_TEMP_SIZE) check(r, "nvb_gotoh_traceback");
    temp.resize(tb ? tb : 1);
    check(nvb_gotoh_traceback(TYPE, &s, &patterns, nullptr, &texts, n, out.d_scores, out.d_sinks, out.d_sources, out.d_ops, out.max_ops, out.d_n_ops, temp.ptr, &tb, stream),
          "nvb_gotoh_traceback");
}

/// one [window_begin, window_end) pass of aln::banded_alignment_score<BAND_LEN>(..., window_begin, window_end, sink, checkpoint)
/// over a batch (nvbio/alignment/banded_inl.h:178-218); d_checkpoints holds BAND_LEN short2 per alignment
template <uint32_t BAND_LEN, AlignmentType TYPE, typename scheme_type>
void batch_banded_alignment_score_window(const GotohAligner<TYPE, scheme_type> aligner, const nvb_string_set& patterns, const nvb_string_set& texts,
                                         uint32_t n, uint32_t window_begin, uint32_t window_end, const int32_t* d_min_scores, int16_t* d_checkpoints,
                                         int32_t* d_scores, nvb_uint2* d_sinks, uint8_t* d_alive, cudaStream_t stream = 0)
{
    const nvb_gotoh_scheme s = aligner.scheme.abi();
    check(nvb_banded_gotoh_score_window(BAND_LEN, TYPE, &s, &patterns, nullptr, &texts, n, window_begin, window_end, d_min_scores, d_checkpoints,
                                        d_scores, d_sinks, d_alive, stream), "nvb_banded_gotoh_score_window");
}

} // namespace aln
} // namespace nvbio_b200
