// Device-side ingest of the text formats the reference reads before a fit:
//   * MatrixMarket coordinate files -- `scipy.io.mmread` at bin/scHPF:373-374, written by its
//     `prep` command with field='integer' (bin/scHPF:327,361);
//   * the reference's own tab-separated triples -- `load_coo`, schpf/preprocessing.py:11-29
//     (`np.loadtxt(filename, delimiter='\t', dtype=int)`: one "row<TAB>col<TAB>count" line per nonzero).
// The file's bytes are copied to the device as they are and parsed there: a first pass counts the
// data lines of every 8 KB block, an exclusive scan turns the counts into output offsets, a second
// pass parses each line where it starts.  Output order = file order (what mmread / loadtxt return),
// so the result is compared with theirs by np.array_equal.  The header (comments and the size line
// of a .mtx) is at most a few hundred bytes and is read by the host, which passes the offset of the
// first data byte.
//
// Integer work throughout: indices and counts are exact or the line is rejected (position reported).
#include <cub/cub.cuh>

#include "common.cuh"

namespace schpf {

namespace {

constexpr int INGEST_THREADS = 256;
constexpr int INGEST_BYTES_PER_THREAD = 32;
constexpr int64_t INGEST_BLOCK_BYTES = (int64_t)INGEST_THREADS * INGEST_BYTES_PER_THREAD;

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || c == '\t'; }
__device__ __forceinline__ bool is_digit(char c) { return c >= '0' && c <= '9'; }

// a data line starts at byte i: first byte of the data region or the byte after a newline, and the
// line is neither empty nor a comment
__device__ __forceinline__ bool line_starts_at(const char *__restrict__ t, int64_t i, int64_t begin)
{
    if (i != begin && t[i - 1] != '\n') return false;
    const char c = t[i];
    return c != '\n' && c != '\r' && c != '%';
}

__global__ void __launch_bounds__(INGEST_THREADS)
count_lines_kernel(const char *__restrict__ text, int64_t begin, int64_t end, int *__restrict__ block_counts)
{
    typedef cub::BlockReduce<int, INGEST_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const int64_t b = begin + (int64_t)blockIdx.x * INGEST_BLOCK_BYTES + (int64_t)threadIdx.x * INGEST_BYTES_PER_THREAD;
    int n = 0;
    for (int64_t i = b; i < b + INGEST_BYTES_PER_THREAD && i < end; ++i) n += line_starts_at(text, i, begin);
    const int total = Reduce(tmp).Sum(n);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// unsigned decimal integer at t[p..); advances p; false when there is no digit or it exceeds `limit`
__device__ __forceinline__ bool parse_index(const char *__restrict__ t, int64_t &p, int64_t end, long long limit,
                                            long long &v)
{
    while (p < end && is_space(t[p])) ++p;
    if (p < end && t[p] == '+') ++p;
    bool any = false;
    v = 0;
    while (p < end && is_digit(t[p])) {
        v = v * 10 + (t[p] - '0');
        if (v > limit) return false;
        any = true;
        ++p;
    }
    return any;
}

// a count: an integer, possibly written as a real ("3", "3.0", "3.000000000000000e+00", "30e-1");
// false when it is negative, not an integer or does not fit int32
__device__ __forceinline__ bool parse_count(const char *__restrict__ t, int64_t &p, int64_t end, long long &v)
{
    while (p < end && is_space(t[p])) ++p;
    if (p < end && t[p] == '+') ++p;
    unsigned long long mant = 0;
    int ndig = 0, e10 = 0;
    bool any = false;
    while (p < end && is_digit(t[p])) {
        if (ndig < 18) {
            mant = mant * 10 + (unsigned)(t[p] - '0');
            ndig += (mant != 0);
        } else ++e10;                                   // digits beyond 18 only scale the value
        any = true;
        ++p;
    }
    if (p < end && t[p] == '.') {
        ++p;
        while (p < end && is_digit(t[p])) {
            if (ndig < 18) {
                mant = mant * 10 + (unsigned)(t[p] - '0');
                ndig += (mant != 0);
                --e10;
            } else if (t[p] != '0') return false;       // a non-zero digit we cannot represent
            any = true;
            ++p;
        }
    }
    if (!any) return false;
    if (p < end && (t[p] == 'e' || t[p] == 'E')) {
        ++p;
        bool neg = false;
        if (p < end && (t[p] == '+' || t[p] == '-')) neg = t[p++] == '-';
        int ex = 0;
        bool anye = false;
        while (p < end && is_digit(t[p])) {
            if (ex < 10000) ex = ex * 10 + (t[p] - '0');
            anye = true;
            ++p;
        }
        if (!anye) return false;
        e10 += neg ? -ex : ex;
    }
    if (mant == 0) {
        v = 0;
        return true;
    }
    while (e10 < 0) {
        if (mant % 10) return false;                    // fractional part: not a count
        mant /= 10;
        ++e10;
    }
    while (e10 > 0) {
        mant *= 10;
        if (mant > 0x7fffffffULL) return false;
        --e10;
    }
    if (mant > 0x7fffffffULL) return false;
    v = (long long)mant;
    return true;
}

__global__ void __launch_bounds__(INGEST_THREADS)
parse_lines_kernel(const char *__restrict__ text, int64_t begin, int64_t end, int nfields, int index_base,
                   const int64_t *__restrict__ block_first, int64_t capacity, int32_t *__restrict__ row,
                   int32_t *__restrict__ col, int32_t *__restrict__ val, unsigned long long *__restrict__ err_pos)
{
    typedef cub::BlockScan<int, INGEST_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t b = begin + (int64_t)blockIdx.x * INGEST_BLOCK_BYTES + (int64_t)threadIdx.x * INGEST_BYTES_PER_THREAD;
    int n = 0;
    for (int64_t i = b; i < b + INGEST_BYTES_PER_THREAD && i < end; ++i) n += line_starts_at(text, i, begin);
    int first = 0;
    Scan(tmp).ExclusiveSum(n, first);
    int64_t out = block_first[blockIdx.x] + first;
    for (int64_t i = b; i < b + INGEST_BYTES_PER_THREAD && i < end; ++i) {
        if (!line_starts_at(text, i, begin)) continue;
        int64_t p = i;
        long long r = 0, c = 0, y = 1;
        bool ok = parse_index(text, p, end, 0x7fffffffLL + index_base, r) &&
                  parse_index(text, p, end, 0x7fffffffLL + index_base, c);
        if (ok && nfields == 3) ok = parse_count(text, p, end, y);
        while (ok && p < end && (is_space(text[p]) || text[p] == '\r')) ++p;
        ok = ok && (p >= end || text[p] == '\n') && r >= index_base && c >= index_base && out < capacity;
        if (!ok) {
            atomicMin(err_pos, (unsigned long long)i);
        } else {
            row[out] = (int32_t)(r - index_base);
            col[out] = (int32_t)(c - index_base);
            val[out] = (int32_t)y;
        }
        ++out;
    }
}

// device temporaries of one schpf_parse_triples call
struct IngestPlan {
    int nblocks = 0;
    int *counts = nullptr;              // data lines per block
    int64_t *first = nullptr;           // their exclusive prefix sum = first output index of a block
    void *tmp = nullptr;                // CUB work area
    size_t tmp_bytes = 0;
    unsigned long long *err = nullptr;  // byte offset of the first malformed line (all ones = none)
};

}  // namespace

}  // namespace schpf

using namespace schpf;

extern "C" {

// Number of data lines in d_text[begin, nbytes): every non-empty line that does not start with '%'.
int schpf_count_lines(int device, void *stream_v, const char *d_text, int64_t nbytes, int64_t begin, int64_t *n_lines)
{
    if (!d_text || !n_lines || begin < 0 || begin > nbytes) {
        set_error("schpf_count_lines: bad arguments (nbytes=%lld begin=%lld)", (long long)nbytes, (long long)begin);
        return SCHPF_ERR_ARG;
    }
    *n_lines = 0;
    if (begin == nbytes) return SCHPF_OK;
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const int64_t nb64 = (nbytes - begin + INGEST_BLOCK_BYTES - 1) / INGEST_BLOCK_BYTES;
    if (nb64 > 0x7fffffffLL) {
        set_error("schpf_count_lines: %lld bytes are more than one call handles", (long long)(nbytes - begin));
        return SCHPF_ERR_ARG;
    }
    const int nblocks = (int)nb64;
    int *counts = nullptr;
    int64_t *total = nullptr;
    CUDA_TRY(pool_malloc(reinterpret_cast<void **>(&counts), sizeof(int) * (size_t)nblocks, stream));
    CUDA_TRY(pool_malloc(reinterpret_cast<void **>(&total), sizeof(int64_t), stream));
    count_lines_kernel<<<nblocks, INGEST_THREADS, 0, stream>>>(d_text, begin, nbytes, counts);
    cudaError_t e = cudaGetLastError();
    size_t tmp_bytes = 0;
    cub::DeviceReduce::Sum(nullptr, tmp_bytes, counts, total, nblocks, stream);
    void *tmp = nullptr;
    if (e == cudaSuccess) e = pool_malloc(&tmp, tmp_bytes ? tmp_bytes : 1, stream);
    if (e == cudaSuccess) e = cub::DeviceReduce::Sum(tmp, tmp_bytes, counts, total, nblocks, stream);
    int64_t h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, total, sizeof(int64_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (tmp) pool_free(tmp, stream);
    pool_free(total, stream);
    pool_free(counts, stream);
    if (e != cudaSuccess) {
        set_error("schpf_count_lines -> %s", cudaGetErrorString(e));
        return SCHPF_ERR_CUDA;
    }
    *n_lines = h;
    return SCHPF_OK;
}

// Parse the data lines of d_text[begin, nbytes) into COO triples on the device, in file order.
//   nfields 3: "row col count" (count may be written as a real with zero fraction); 2: "row col" (count 1,
//   MatrixMarket `pattern`); fields are separated by blanks or tabs; CR LF line ends are accepted.
//   index_base: 1 for MatrixMarket, 0 for the reference's tsv (preprocessing.py:13-15).
// On a malformed line nothing is promised about the outputs; the byte offset of the first one is
// reported through *err_offset (and in schpf_last_error).
int schpf_parse_triples(int device, void *stream_v, const char *d_text, int64_t nbytes, int64_t begin, int nfields,
                        int index_base, int32_t *d_row, int32_t *d_col, int32_t *d_val, int64_t capacity,
                        int64_t *n_out, int64_t *err_offset)
{
    if (!d_text || !d_row || !d_col || !d_val || !n_out || begin < 0 || begin > nbytes || capacity < 0 ||
        (nfields != 2 && nfields != 3) || (index_base != 0 && index_base != 1)) {
        set_error("schpf_parse_triples: bad arguments");
        return SCHPF_ERR_ARG;
    }
    *n_out = 0;
    if (err_offset) *err_offset = -1;
    if (begin == nbytes) return SCHPF_OK;
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const int64_t nb64 = (nbytes - begin + INGEST_BLOCK_BYTES - 1) / INGEST_BLOCK_BYTES;
    if (nb64 > 0x7fffffffLL) {
        set_error("schpf_parse_triples: %lld bytes are more than one call handles", (long long)(nbytes - begin));
        return SCHPF_ERR_ARG;
    }
    IngestPlan P;
    P.nblocks = (int)nb64;
    cudaError_t e = pool_malloc(reinterpret_cast<void **>(&P.counts), sizeof(int) * (size_t)P.nblocks, stream);
    if (e == cudaSuccess) e = pool_malloc(reinterpret_cast<void **>(&P.first), sizeof(int64_t) * ((size_t)P.nblocks + 1), stream);
    if (e == cudaSuccess) e = pool_malloc(reinterpret_cast<void **>(&P.err), sizeof(unsigned long long), stream);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, P.tmp_bytes, P.counts, P.first, P.nblocks, stream);
    if (e == cudaSuccess) e = pool_malloc(&P.tmp, P.tmp_bytes ? P.tmp_bytes : 1, stream);
    int64_t first_last = 0;
    int count_last = 0;
    unsigned long long err = ~0ULL;
    if (e == cudaSuccess) {
        e = cudaMemsetAsync(P.err, 0xff, sizeof(unsigned long long), stream);
        count_lines_kernel<<<P.nblocks, INGEST_THREADS, 0, stream>>>(d_text, begin, nbytes, P.counts);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cub::DeviceScan::ExclusiveSum(P.tmp, P.tmp_bytes, P.counts, P.first, P.nblocks, stream);
        parse_lines_kernel<<<P.nblocks, INGEST_THREADS, 0, stream>>>(d_text, begin, nbytes, nfields, index_base, P.first,
                                                                   capacity, d_row, d_col, d_val, P.err);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&first_last, P.first + (P.nblocks - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&count_last, P.counts + (P.nblocks - 1), sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&err, P.err, sizeof(err), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    }
    if (P.tmp) pool_free(P.tmp, stream);
    if (P.err) pool_free(P.err, stream);
    if (P.first) pool_free(P.first, stream);
    if (P.counts) pool_free(P.counts, stream);
    if (e != cudaSuccess) {
        set_error("schpf_parse_triples -> %s", cudaGetErrorString(e));
        return SCHPF_ERR_CUDA;
    }
    const int64_t n = first_last + count_last;
    *n_out = n;
    if (n > capacity) {
        set_error("schpf_parse_triples: %lld data lines, the output buffers hold %lld", (long long)n, (long long)capacity);
        return SCHPF_ERR_ARG;
    }
    if (err != ~0ULL) {
        if (err_offset) *err_offset = (int64_t)err;
        set_error("malformed data line at byte %llu (expected %d integer fields, indices >= %d, counts that are "
                  "non-negative integers below 2^31)", err, nfields, index_base);
        return SCHPF_ERR_ARG;
    }
    return SCHPF_OK;
}

}  // extern "C"
