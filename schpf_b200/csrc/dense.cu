// Dense (cells x K / genes x K) kernels and the simple per-nonzero kernels.
//
// Reference lines each kernel replaces are cited at the kernel.  None of these
// is the hot kernel (that is sweep.cu); they touch (C+G)*K doubles per
// iteration, or are debugging / t==0 / function-level paths.
#include "common.cuh"

namespace schpf {

namespace {

constexpr int DENSE_THREADS = 256;
constexpr int DENSE_WARPS = DENSE_THREADS / 32;

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// one warp per row, rows strided over a grid of at most 148 SMs x 8 CTAs
inline int finalize_grid(int64_t n)
{
    const int64_t want = (n + (256 / 32) - 1) / (256 / 32);
    return (int)(want < 1 ? 1 : want > 148 * 8 ? 148 * 8 : want);
}

// column sums of per-thread values over the block, then one atomicAdd per k
// per block (K <= 64).  `val(k)` returns this thread's contribution.
template <typename F>
__device__ __forceinline__ void block_colsum_atomic(int K, F val, double *colsum)
{
    __shared__ double sh[DENSE_WARPS][SCHPF_MAX_FACTORS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < K; ++k) {
        const double v = warp_sum(val(k));
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < DENSE_WARPS; ++w) t += sh[w][threadIdx.x];
        atomicAdd(colsum + threadIdx.x, t);
    }
}

// One WARP per row of a (n x K) Gamma family, lanes over the factors (K <= 64: two per lane).
//   UPDATE: shp = prior_shape + v,  rte = cap_shp/cap_rte(old) + other_colsum[k],
//           cap_rte = prior_rate + sum_k shp/rte
//     beta : scHPF_.py:699-704 (hpf_numba.py:129-156, :160-177)   v = sum over the gene's nonzeros
//     theta: scHPF_.py:709-714
//   always: elog = psi(shp) - log(rte)  (hpf_numba.py:83-94, scHPF_.py:108-111),
//           E = exp(elog - max_k elog)  (factored softmax table),
//           colsum_out[k] += shp/rte    (the sum at hpf_numba.py:167-170 for the NEXT rate update)
// Row accesses are coalesced, the digamma / log / exp chains of a row run in parallel, the
// row-wise sum and max are warp shuffles, and each lane carries its own column's partial sum
// over the rows its warp visits (summed over the CTA's warps, then one atomicAdd per column and CTA).
template <bool UPDATE>
__global__ void __launch_bounds__(DENSE_THREADS)
finalize_kernel(int64_t n, int K, TabGeom tg, double prior_shape, double prior_rate,
                const double *__restrict__ folded, const double *__restrict__ acc,
                const double *__restrict__ direct, const double *__restrict__ other_colsum,
                const double *__restrict__ cap_shp, double *__restrict__ cap_rte,
                double *__restrict__ shp, double *__restrict__ rte, double *__restrict__ elog,
                double *__restrict__ Etab, double *__restrict__ colsum_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * DENSE_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * DENSE_THREADS) >> 5;
    double col[2] = {0.0, 0.0};
    double oc[2] = {0.0, 0.0};
    if (UPDATE) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (lane + 32 * t < K) oc[t] = other_colsum[lane + 32 * t];
    }
    for (int64_t i = warp0; i < n; i += nwarps) {
        double sk[2] = {1.0, 1.0}, rk[2] = {1.0, 1.0}, el[2];
        double sum_ex = 0.0, m = -INFINITY;
        const double cap_ex = UPDATE ? cap_shp[i] / cap_rte[i] : 0.0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int k = lane + 32 * t;
            el[t] = -INFINITY;
            if (k < K) {
                if (UPDATE) {
                    double v;
                    if (folded) v = folded[i * K + k];
                    else v = fma(Etab[tab_index(tg, i, k)], acc[i * K + k], direct[i * K + k]);
                    sk[t] = prior_shape + v;
                    rk[t] = cap_ex + oc[t];
                    shp[i * K + k] = sk[t];
                    rte[i * K + k] = rk[t];
                } else {
                    sk[t] = shp[i * K + k];
                    rk[t] = rte[i * K + k];
                }
                const double ex = sk[t] / rk[t];
                sum_ex += ex;
                col[t] += ex;
                el[t] = digamma_pos(sk[t]) - log(rk[t]);
                elog[i * K + k] = el[t];
                m = fmax(m, el[t]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (UPDATE) {
            sum_ex = warp_sum(sum_ex);
            if (lane == 0) cap_rte[i] = prior_rate + sum_ex;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int k = lane + 32 * t;
            if (k < K) tab_store(tg, Etab, i, k, exp(el[t] - m));
        }
    }
    if (colsum_out) {
        // one atomic per column and CTA: the K addresses are shared by the whole grid, and
        // same-address fp64 atomics serialise in L2 (per-warp atomics were ~50 us of a 60 us kernel)
        __shared__ double part[DENSE_WARPS][64];
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int t = 0; t < 2; ++t) part[warp][lane + 32 * t] = col[t];
        __syncthreads();
        if (threadIdx.x < K) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < DENSE_WARPS; ++w) t += part[w][threadIdx.x];
            atomicAdd(colsum_out + threadIdx.x, t);
        }
    }
}

// e_x table in the padded sweep layout (hpf_numba.py:33-41)
__global__ void ex_table_kernel(int64_t n, int K, TabGeom tg, const double *__restrict__ shp,
                                const double *__restrict__ rte, double *__restrict__ X)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * K) return;
    const int64_t i = idx / K;
    const int k = (int)(idx - i * K);
    tab_store(tg, X, i, k, shp[idx] / rte[idx]);
}

// out[i,k] = E[i,k] * acc[i,k] + direct[i,k] : this shard's part of
// sum_{nonzeros of gene i} Xphi[:,k]  (the loop at hpf_numba.py:152-155)
__global__ void fold_kernel(int64_t n, int K, TabGeom tg, const double *__restrict__ E,
                            const double *__restrict__ acc, const double *__restrict__ direct,
                            double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * K) return;
    const int64_t i = idx / K;
    const int k = (int)(idx - i * K);
    out[idx] = fma(E[tab_index(tg, i, k)], acc[idx], direct[idx]);
}

// The reference's E-step written literally (hpf_numba.py:98-112): one thread
// per nonzero, log-space softmax with the max subtracted, (y*rho)/sum; the
// result is optionally materialised (debug / function-level shim) and
// optionally scatter-added with fp64 atomics (the "variant 1" engine path and
// the cross-check of the tiled sweep).
__global__ void literal_kernel(int64_t nnz, int K, const int32_t *__restrict__ row,
                               const int32_t *__restrict__ col, const int32_t *__restrict__ data,
                               const double *__restrict__ elog_t, const double *__restrict__ elog_b,
                               double *__restrict__ xphi_out, double *__restrict__ direct_t,
                               double *__restrict__ direct_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = row[i], c = col[i];
    const double y = (double)data[i];
    const double *et = elog_t + r * K, *eb = elog_b + c * K;
    double largest = -INFINITY;
    for (int k = 0; k < K; ++k) largest = fmax(largest, et[k] + eb[k]);
    double normalizer = 0.0;
    for (int k = 0; k < K; ++k) normalizer += exp(et[k] + eb[k] - largest);
    for (int k = 0; k < K; ++k) {
        const double v = y * exp(et[k] + eb[k] - largest) / normalizer;
        if (xphi_out) xphi_out[i * K + k] = v;
        if (direct_t) atomicAdd(direct_t + r * K + k, v);
        if (direct_b) atomicAdd(direct_b + c * K + k, v);
    }
}

// The queued nonzeros of one sweep whose factored normaliser underflowed, redone exactly as the
// reference does every nonzero (hpf_numba.py:98-112: log space, max subtracted) and added to the
// owner's `direct` row.  Launched after every shape sweep with a small fixed grid; the queue is
// empty unless priors are far below 1e-2, and the kernel then costs one read.
struct FixupSide {
    const int4 *queue;
    const unsigned long long *count;
    const double *own_elog, *oth_elog;
    double *direct;
};

// blockIdx.y selects the sweep direction (one launch redoes both queues)
__global__ void slow_fixup_kernel(FixupSide s0, FixupSide s1, unsigned int cap, int K,
                                  unsigned long long *__restrict__ slow_hits, int *__restrict__ overflow)
{
    const FixupSide S = blockIdx.y == 0 ? s0 : s1;
    const int4 *__restrict__ queue = S.queue;
    const double *__restrict__ own_elog = S.own_elog, *__restrict__ oth_elog = S.oth_elog;
    double *__restrict__ direct = S.direct;
    unsigned long long n = *S.count;
    if (n == 0) return;
    if (n > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(overflow, 1);
        n = cap;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(slow_hits, n);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int4 q = queue[i];
        const double y = __hiloint2double(q.z, q.w);
        const double *eo = own_elog + (int64_t)q.x * K, *et = oth_elog + (int64_t)q.y * K;
        double largest = -INFINITY, normalizer = 0.0;
        for (int k = 0; k < K; ++k) largest = fmax(largest, eo[k] + et[k]);
        for (int k = 0; k < K; ++k) normalizer += exp(eo[k] + et[k] - largest);
        for (int k = 0; k < K; ++k)
            atomicAdd(direct + (int64_t)q.x * K + k, y * exp(eo[k] + et[k] - largest) / normalizer);
    }
}

// Philox4x32-10 (Salmon et al., SC'11), written from the published round function.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double exp1_draw(uint32_t bits)
{
    // -log(U), U uniform on (0,1) with 32-bit resolution
    return -log(((double)bits + 0.5) * 2.3283064365386963e-10);
}

// t == 0 of _fit with reinit (scHPF_.py:652-655): Xphi_i = y_i * Dirichlet(1_K),
// drawn as normalised Exp(1) variates from a counter-based generator keyed by
// (seed; global row, col), so the draw does not depend on sharding or order.
__global__ void random_phi_kernel(int64_t nnz, int K, const int32_t *__restrict__ row,
                                  const int32_t *__restrict__ col, const int32_t *__restrict__ data,
                                  uint64_t seed, int64_t row_offset, double *__restrict__ direct_t,
                                  double *__restrict__ direct_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = row[i], c = col[i];
    const uint64_t gr = (uint64_t)(r + row_offset);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int nblk = (K + 3) / 4;
    double total = 0.0;
    for (int j = 0; j < nblk; ++j) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)c, (uint32_t)j), key);
        const uint32_t b[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (4 * j + d < K) total += exp1_draw(b[d]);
    }
    const double scale = (double)data[i] / total;
    for (int j = 0; j < nblk; ++j) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)c, (uint32_t)j), key);
        const uint32_t b[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int k = 4 * j + d;
            if (k < K) {
                const double v = exp1_draw(b[d]) * scale;
                atomicAdd(direct_t + r * K + k, v);
                if (direct_b) atomicAdd(direct_b + c * K + k, v);
            }
        }
    }
}

// hpf_numba.py:152-155 with fp64 atomics: out[keep[i], k] += xphi[i, k]
__global__ void scatter_xphi_kernel(int64_t nnz, int K, const double *__restrict__ xphi,
                                    const int32_t *__restrict__ keep, double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnz * K) return;
    const int64_t i = idx / K;
    const int k = (int)(idx - i * K);
    atomicAdd(out + (int64_t)keep[i] * K + k, xphi[idx]);
}

// hpf_numba.py:45-50 per nonzero, straight from shape / rate
__global__ void llh_pointwise_kernel(int64_t nnz, int K, const int32_t *__restrict__ row,
                                     const int32_t *__restrict__ col, const int32_t *__restrict__ data,
                                     const double *__restrict__ ts, const double *__restrict__ tr,
                                     const double *__restrict__ bs, const double *__restrict__ br,
                                     double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = (int64_t)row[i] * K, c = (int64_t)col[i] * K;
    double e_rate = 0.0;
    for (int k = 0; k < K; ++k) e_rate += (ts[r + k] / tr[r + k]) * (bs[c + k] / br[c + k]);
    const double y = (double)data[i];
    out[i] = y * log(e_rate) - e_rate - lgamma(y + 1.0);
}

// sum_i lgamma(y_i + 1): a constant of the data (hpf_numba.py:50), taken once
__global__ void __launch_bounds__(DENSE_THREADS)
lgamma_partial_kernel(int64_t nnz, const int32_t *__restrict__ data, double *__restrict__ partials)
{
    // counts are small integers almost always: ln(y!) for y < 256 comes from a table the block fills once with
    // the same lgamma (so the sum is bit-identical to evaluating it per nonzero, at a fraction of the cost)
    __shared__ double table[DENSE_THREADS];
    table[threadIdx.x] = lgamma((double)threadIdx.x + 1.0);
    __syncthreads();
    double t = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * DENSE_THREADS + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * DENSE_THREADS) {
        const int y = data[i];
        t += (unsigned)y < (unsigned)DENSE_THREADS ? table[y] : lgamma((double)y + 1.0);
    }
    __shared__ double sh[DENSE_WARPS];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DENSE_WARPS; ++w) s += sh[w];
        partials[blockIdx.x] = s;
    }
}

// fixed-order final reduction (single block): deterministic
__global__ void __launch_bounds__(DENSE_THREADS)
sum_partials_kernel(const double *__restrict__ partials, int n, double *__restrict__ out)
{
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += DENSE_THREADS) t += partials[i];
    __shared__ double sh[DENSE_WARPS];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DENSE_WARPS; ++w) s += sh[w];
        *out = s;
    }
}

// out = [sum_i llh_i of this shard, its nnz]: the two scalars shards all-reduce at a loss check
__global__ void pack_loss_kernel(const double *__restrict__ llh_sum, double lgamma_sum, double nnz,
                                 double *__restrict__ out2)
{
    out2[0] = *llh_sum - lgamma_sum;
    out2[1] = nnz;
}

__global__ void validate_coo_kernel(int64_t nnz, const int32_t *__restrict__ row,
                                    const int32_t *__restrict__ col, const int32_t *__restrict__ data,
                                    int64_t C, int64_t G, int *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int f = 0;
    if (row[i] < 0 || row[i] >= C) f |= 1;
    if (col[i] < 0 || col[i] >= G) f |= 2;
    if (data[i] < 0) f |= 4;
    if (data[i] >= (1 << PACKED_COUNT_BITS)) f |= 8;      // too large for the packed stream format
    if (data[i] >= (1 << YHI_MAX_COUNT_BITS)) f |= 16;    // low word of (double)count is not zero
    if (f) atomicOr(flag, f);
}

__global__ void fill_kernel(double *p, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void psi_kernel(int64_t n, const double *__restrict__ x, double *__restrict__ out, int which)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = which == 0 ? digamma_pos(x[i]) : lgamma(x[i]);
}

// hpf_numba.py:172-176: result[i,k] = prior_shp[i]/prior_rte[i] + colsum[k]
__global__ void rate_update_kernel(int64_t n, int K, const double *__restrict__ pshp,
                                   const double *__restrict__ prte, const double *__restrict__ colsum,
                                   double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * K) return;
    const int64_t i = idx / K;
    const int k = (int)(idx - i * K);
    out[idx] = pshp[i] / prte[i] + colsum[k];
}

// hpf_numba.py:167-170: colsum[k] = sum_i shp[i,k]/rte[i,k]
__global__ void __launch_bounds__(DENSE_THREADS)
colsum_ex_kernel(int64_t m, int K, const double *__restrict__ shp, const double *__restrict__ rte,
                 double *__restrict__ colsum)
{
    const int64_t i = (int64_t)blockIdx.x * DENSE_THREADS + threadIdx.x;
    const bool live = i < m;
    block_colsum_atomic(
        K, [&](int k) { return live ? shp[i * K + k] / rte[i * K + k] : 0.0; }, colsum);
}

// hpf_numba.py:181-188: result[i] = prior + sum_k shp/rte (k ascending)
__global__ void capacity_rate_kernel(int64_t n, int K, const double *__restrict__ shp,
                                     const double *__restrict__ rte, double prior,
                                     double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = prior;
    for (int k = 0; k < K; ++k) t += shp[i * K + k] / rte[i * K + k];
    out[i] = t;
}

}  // namespace

#define LAUNCH_CHECK()                                                                      \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return SCHPF_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

int launch_slow_fixup(cudaStream_t s, const SweepArgs &A, const SweepArgs *B, int *overflow_flag)
{
    const FixupSide a{A.slow_queue, A.slow_count, A.own_elog, A.oth_elog, A.direct};
    const FixupSide b = B ? FixupSide{B->slow_queue, B->slow_count, B->own_elog, B->oth_elog, B->direct} : a;
    slow_fixup_kernel<<<dim3(64, B ? 2 : 1), 128, 0, s>>>(a, b, A.slow_cap, A.K, A.slow_hits, overflow_flag);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_prep_side(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *shp, const double *rte,
                     double *elog, double *E, double *colsum)
{
    if (n <= 0) return SCHPF_OK;
    finalize_kernel<false><<<finalize_grid(n), DENSE_THREADS, 0, s>>>(
        n, K, tg, 0.0, 0.0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
        const_cast<double *>(shp), const_cast<double *>(rte), elog, E, colsum);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_finalize(cudaStream_t s, int64_t n, int K, const TabGeom &tg, double prior_shape, double prior_rate,
                    const double *folded, const double *E, const double *acc, const double *direct,
                    const double *other_colsum, const double *cap_shp, double *cap_rte, double *shp,
                    double *rte, double *elog, double *Etab, double *colsum_out)
{
    if (n <= 0) return SCHPF_OK;
    (void)E;  // the factored table read for the fold is the one being rewritten (Etab)
    finalize_kernel<true><<<finalize_grid(n), DENSE_THREADS, 0, s>>>(
        n, K, tg, prior_shape, prior_rate, folded, acc, direct, other_colsum, cap_shp, cap_rte, shp,
        rte, elog, Etab, colsum_out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_ex_table(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *shp, const double *rte, double *X)
{
    if (n <= 0) return SCHPF_OK;
    ex_table_kernel<<<blocks_for(n * K, 256), 256, 0, s>>>(n, K, tg, shp, rte, X);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_fold(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *E, const double *acc,
                const double *direct, double *out)
{
    if (n <= 0) return SCHPF_OK;
    fold_kernel<<<blocks_for(n * K, 256), 256, 0, s>>>(n, K, tg, E, acc, direct, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_literal(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                   const int32_t *data, const double *elog_t, const double *elog_b, double *xphi_out,
                   double *direct_t, double *direct_b)
{
    if (nnz <= 0) return SCHPF_OK;
    literal_kernel<<<blocks_for(nnz, 256), 256, 0, s>>>(nnz, K, row, col, data, elog_t, elog_b, xphi_out,
                                                       direct_t, direct_b);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_random_phi(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                      const int32_t *data, uint64_t seed, int64_t row_offset, double *direct_t,
                      double *direct_b)
{
    if (nnz <= 0) return SCHPF_OK;
    random_phi_kernel<<<blocks_for(nnz, 256), 256, 0, s>>>(nnz, K, row, col, data, seed, row_offset,
                                                          direct_t, direct_b);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_scatter_xphi(cudaStream_t s, int64_t nnz, int K, const double *xphi, const int32_t *keep,
                        double *out)
{
    if (nnz <= 0) return SCHPF_OK;
    scatter_xphi_kernel<<<blocks_for(nnz * K, 256), 256, 0, s>>>(nnz, K, xphi, keep, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_llh_pointwise(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                         const int32_t *data, const double *ts, const double *tr, const double *bs,
                         const double *br, double *out)
{
    if (nnz <= 0) return SCHPF_OK;
    llh_pointwise_kernel<<<blocks_for(nnz, 256), 256, 0, s>>>(nnz, K, row, col, data, ts, tr, bs, br, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_lgamma_sum(cudaStream_t s, int64_t nnz, const int32_t *data, double *partials, int nblk,
                      double *out)
{
    lgamma_partial_kernel<<<nblk, DENSE_THREADS, 0, s>>>(nnz, data, partials);
    LAUNCH_CHECK();
    return launch_sum_partials(s, partials, nblk, out);
}

int launch_sum_partials(cudaStream_t s, const double *partials, int n, double *out)
{
    sum_partials_kernel<<<1, DENSE_THREADS, 0, s>>>(partials, n, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_pack_loss(cudaStream_t s, const double *llh_sum, double lgamma_sum, double nnz, double *out2)
{
    pack_loss_kernel<<<1, 1, 0, s>>>(llh_sum, lgamma_sum, nnz, out2);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_validate_coo(cudaStream_t s, int64_t nnz, const int32_t *row, const int32_t *col,
                        const int32_t *data, int64_t C, int64_t G, int *flag)
{
    if (nnz <= 0) return SCHPF_OK;
    validate_coo_kernel<<<blocks_for(nnz, 256), 256, 0, s>>>(nnz, row, col, data, C, G, flag);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_fill(cudaStream_t s, double *p, int64_t n, double v)
{
    if (n <= 0) return SCHPF_OK;
    fill_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, n, v);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_psi(cudaStream_t s, int64_t n, const double *x, double *out, int which)
{
    if (n <= 0) return SCHPF_OK;
    psi_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, x, out, which);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_rate_update(cudaStream_t s, int64_t n, int K, const double *pshp, const double *prte,
                       const double *colsum, double *out)
{
    if (n <= 0) return SCHPF_OK;
    rate_update_kernel<<<blocks_for(n * K, 256), 256, 0, s>>>(n, K, pshp, prte, colsum, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_colsum_ex(cudaStream_t s, int64_t m, int K, const double *shp, const double *rte,
                     double *colsum)
{
    if (m <= 0) return SCHPF_OK;
    colsum_ex_kernel<<<blocks_for(m, DENSE_THREADS), DENSE_THREADS, 0, s>>>(m, K, shp, rte, colsum);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

int launch_capacity_rate(cudaStream_t s, int64_t n, int K, const double *shp, const double *rte,
                         double prior, double *out)
{
    if (n <= 0) return SCHPF_OK;
    capacity_rate_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, K, shp, rte, prior, out);
    LAUNCH_CHECK();
    return SCHPF_OK;
}

}  // namespace schpf
