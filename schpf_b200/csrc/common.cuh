// Shared declarations for the schpf_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/schpf_b200.h"

namespace schpf {

// ---------------------------------------------------------------- errors ----
void set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                               \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::schpf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                               cudaGetErrorString(e__));                             \
            return SCHPF_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)

#define RC_TRY(expr)                                                                 \
    do {                                                                             \
        int rc__ = (expr);                                                           \
        if (rc__ != SCHPF_OK) return rc__;                                           \
    } while (0)

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the
// release threshold raised, so the multi-GB layout temporaries and per-fit buffers are
// recycled between calls / fits instead of being mapped and unmapped every time.
cudaError_t pool_malloc(void **p, size_t bytes, cudaStream_t stream);
void pool_free(void *p, cudaStream_t stream);

// SCHPF_TRACE=1: print host-side phase timings (synchronises the stream at each mark)
void trace_mark(cudaStream_t stream, const char *what);

// ------------------------------------------------------- table geometry ----
// The sweep kernels give every nonzero to a PAIR of lanes; lane h of the pair
// owns the 16-byte units {2j+h} of a K-row, so rows are padded to KP = a
// multiple of 4 doubles.  Rows are stored with a stride ST chosen so that
// ST*8/32 is odd: then the 32-byte piece a lane pair reads in one LDS.128 lands
// in shared-memory bank group (row + j) mod 4, and four pairs reading rows of
// four different residues mod 4 never collide.
__host__ __device__ constexpr int kp_of(int K) { return (K + 3) & ~3; }
__host__ __device__ constexpr int stride_of_kp(int KP) { return ((KP / 4) & 1) ? KP : KP + 4; }

constexpr int GROUPS_PER_WARP = 16;          // lane pairs (sweep.cu); the one-lane kernel has 32 owners per warp
constexpr double TINY_NORMALIZER = 1e-280;   // below this the factored softmax is redone in log space

// ------------------------------------------------------- device helpers ----
#define SCHPF_EULER 0.57721566490153286061

// digamma for x > 0: upward recurrence to x >= 10, then the Bernoulli asymptotic series
// (Abramowitz & Stegun 6.3.18); exact harmonic numbers for integers <= 10.  The recurrence
// sum_j 1/(x+j) (up to 10 terms) is accumulated as ONE fraction N/D, so it costs a single
// fp64 division instead of ten (the finalize kernels are bound by the fp64 pipe).  Checked
// against scipy.special.digamma in tests/test_gpu_kernels.py: <= 5e-15 * max(1, |psi|).
__device__ __forceinline__ double digamma_pos(double x)
{
    if (!(x > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
    if (x <= 10.0 && x == floor(x)) {
        double y = 0.0;
        for (int i = (int)x - 1; i >= 1; --i) y += 1.0 / (double)i;
        return y - SCHPF_EULER;
    }
    double s = x, num = 0.0, den = 1.0;
    while (s < 10.0) {
        num = fma(num, s, den);      // N/D + 1/s = (N s + D) / (D s)
        den *= s;
        s += 1.0;
    }
    const double w = num / den;
    double y = 0.0;
    if (s < 1.0e17) {
        const double z = 1.0 / (s * s);
        double p = 8.33333333333333333333e-2;
        p = p * z - 2.10927960927960927961e-2;
        p = p * z + 7.57575757575757575758e-3;
        p = p * z - 4.16666666666666666667e-3;
        p = p * z + 3.96825396825396825397e-3;
        p = p * z - 8.33333333333333333333e-3;
        p = p * z + 8.33333333333333333333e-2;
        y = z * p;
    }
    return log(s) - 0.5 / s - y - w;
}

// y / s for a normal positive s: hardware seed (MUFU.RCP64H, relative error e ~ 2^-20),
// one cubic Newton step r(1 + e + e^2) (error e^3 ~ 2^-60), then the product: < 2 ulp.
__device__ __forceinline__ double div_pos(double y, double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    const double e = fma(-s, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    return y * r;
}

// The same quotient with the product folded into the correction: q = y r, w = q (1 + e + e^2).
// The dependent chain is seed -> {e, q} -> t -> w, one operation shorter; same error bound.
__device__ __forceinline__ double div_pos_folded(double y, double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    const double e = fma(-s, r, 1.0);
    const double q = y * r;
    const double t = fma(e, e, e);
    return fma(q, t, q);
}

// p(lane) + p(lane ^ 1) for every lane, without the shuffle unit: one fp64 m8n8k4 MMA.  A[r][c] is
// the value of lane 4r + c; with the 0/1 matrix B of the caller (B[c][n] = 1 iff lanes 4r+c and
// 4r+n/2 form a pair), D[r][2m] -- the first accumulator of lane 4r + m -- is the sum over that
// lane's own pair.  Adding the two exact zeros changes nothing, so the result equals p + p' bit
// for bit.  (Finite inputs only: 0 * inf would poison the other pair of the row.)
__device__ __forceinline__ double pair_sum_mma(double p, double sel)
{
    double d0, d1;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(p), "d"(sel), "d"(0.0), "d"(0.0));
    return d0;
}

// Tables the sweeps read (Et/Eb, Xt/Xb) come in two geometries, chosen per K:
//   pairs (sweep.cu):       one plane, rows of ST doubles        -> KA = ST, KB = 0
//   lanes (sweep_lanes.cu): plane A = first KA = 16*NA doubles of every row (row stride KA), then
//                           plane B = the last KB = 4 doubles (row stride KB) starting at offB = n_pad*KA
struct TabGeom {
    int KA, KB;
    int64_t offB;
    // fp32 sweep (sweep_f32.cu): the table is ALSO written as floats, rows of `kf` floats, and the fp64
    // copy then holds the float-rounded value (the finalisation multiplies by what the sweep read)
    float *f32 = nullptr;
    int kf = 0;
};
// store one table element in every copy the geometry asks for; returns the value the fp64 copy holds
__device__ __forceinline__ double tab_store(const TabGeom &g, double *tab, int64_t i, int k, double v)
{
    if (g.f32) {
        const float f = (float)v;
        g.f32[i * g.kf + k] = f;
        v = (double)f;
    }
    tab[k < g.KA ? i * g.KA + k : g.offB + i * g.KB + (k - g.KA)] = v;
    return v;
}
__host__ __device__ __forceinline__ int64_t tab_index(const TabGeom &g, int64_t i, int k)
{
    return k < g.KA ? i * g.KA + k : g.offB + i * g.KB + (k - g.KA);
}
// padded row length of the one-lane kernel's classes: K <= 16 -> 16, 17..20 -> 20, 29..32 -> 32 (else unsupported).
// K <= 12 runs the 16-wide kernel on zero columns: measured on cfg-3 (profiles/r2e_bench_K*.json) the lane-pair
// kernel takes 2.10 (K=7) and 2.31 ms (K=10) per sweep pair, the 16-wide one-lane kernel 2.02 whatever K is.
__host__ __device__ constexpr int lanes_kp_of(int K)
{
    return (K >= 1 && K <= 16) ? 16 : (K >= 17 && K <= 20) ? 20 : (K >= 29 && K <= 32) ? 32 : 0;
}

// y as the HIGH WORD of its double: counts below 2^21 have an all-zero low word, so the sweep
// builds the double from the stream word with no conversion instruction (I2F.F64 runs on the
// fp64 pipe the kernel is short of)
constexpr int YHI_MAX_COUNT_BITS = 21;
constexpr int TINY_NORMALIZER_HI = 0x05cd0b15;   // high word of TINY_NORMALIZER (1e-280)

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) -------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// 128-bit shared-memory load from a 32-bit shared address (no generic-address arithmetic)
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// streaming 128-bit load that does not allocate in L1 (entry stream is read once)
__device__ __forceinline__ int4 ld_stream_int4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ int2 ld_stream_int2(const int2 *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ---- entry encodings of the sweep stream ------------------------------------------
// wide   (8 B): {row | pad<<31, count}
// packed (4 B): pad<<31 | count<<12 | row       (count < 2^19, row < 2^12); used whenever the
//               largest count of the matrix allows it: half the stream bytes per nonzero
constexpr int PACKED_ROW_BITS = 12;
constexpr int PACKED_COUNT_BITS = 19;
__host__ __device__ inline uint32_t pack_entry(uint32_t row, uint32_t count, bool pad)
{
    return (pad ? 0x80000000u : 0u) | (count << PACKED_ROW_BITS) | row;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------ layouts ------
// One "side" of the two-pass sweep: the OWNER axis keeps its K-row and its
// accumulators in registers, the OTHER axis is streamed through shared memory
// in panels of `panel_rows` rows.  See DESIGN.md §3.
struct SideLayout {
    int64_t n_own = 0, n_oth = 0;
    int panel_rows = 0;        // Po
    int npanel = 0;
    int warps = 0;             // W warps per CTA; W*16 owners per CTA block
    int nblocks = 0;           // owner blocks
    int panels_per_range = 0;  // PR
    int nranges = 0;
    int64_t total_pairs = 0;   // warp-steps / 2
    int64_t padded_entries = 0;
    int32_t *own_id = nullptr;   // [nblocks * W * 16] slot -> owner id, -1 = empty slot
    int64_t *seg_ptr = nullptr;  // [(nblocks * W) * (npanel + 1)] in step pairs
    void *entries = nullptr;     // [total_pairs * opw] x (int4 wide | int2 packed): two steps per element
    bool packed = false;
    int opw = GROUPS_PER_WARP;   // owners per warp: 16 (one per lane pair) or 32 (one per lane)
    bool free_mode = false;      // no bank-class schedule (tables whose rows all start at bank group 0)
    bool ranked_per_range = false;   // owners re-ranked by their count inside every panel range
    bool yhi = false;            // count field holds the high word of (double)count
    int64_t n_slots = 0;         // nblocks * warps * opw
    size_t bytes = 0;
    cudaStream_t stream = nullptr;   // stream the buffers were allocated on (pool_free)
    void release();
};

enum LayoutFlags { LAYOUT_FREE = 1, LAYOUT_RANK_PER_RANGE = 2, LAYOUT_YHI = 4, LAYOUT_SINGLE_PANEL_RANGES = 8,
                   LAYOUT_ROTATE_CLASS = 16 };
int build_side_layout(SideLayout &L, cudaStream_t stream, int64_t nnz, const int32_t *d_own,
                      const int32_t *d_oth, const int32_t *d_val, int64_t n_own, int64_t n_oth,
                      int panel_rows, int warps, int target_ctas, bool packed, int owners_per_warp = GROUPS_PER_WARP,
                      int flags = 0);
// debug / tests: the layout's entries decoded back to (owner, other, count) triples, host arrays of
// L.padded_entries elements each (pads: count 0 and other = -1); returns the number written
int64_t dump_side_layout(const SideLayout &L, cudaStream_t stream, int32_t *own, int32_t *oth, int32_t *cnt);

// frees the per-device scratch block the layout builds keep between calls
int release_layout_scratch(int device);

// ------------------------------------------------------------- sweeps ------
enum SweepMode { SWEEP_SHAPE = 0, SWEEP_LLH = 1 };

struct SweepArgs {
    const int32_t *own_id;
    const int64_t *seg_ptr;
    const void *entries;     // int4 (wide) or int2 (packed) per lane pair and two steps
    const double *own_tab;   // [n_own x ST]   owner-side table (factored exp / e_x)
    const double *oth_tab;   // [npanel*Po x ST]
    double *acc;             // SHAPE: [n_own x K] sum_i w_i * oth_tab[oth_i, k]   (atomicAdd)
    double *partial;         // LLH:   [gridDim.x] sum (y log r - r)
    const double *own_elog;  // [n_own x K]  log-space fallback
    const double *oth_elog;  // [n_oth x K]
    double *direct;          // [n_own x K]  fallback contributions, already y*phi
    unsigned long long *slow_hits;
    // nonzeros whose factored normaliser underflowed are queued here and redone in log space by
    // slow_fixup_kernel after the sweep (the exp() code stays out of the hot loop's register budget)
    int4 *slow_queue;                 // {owner, other (global), y as double hi, lo}
    unsigned long long *slow_count;
    unsigned int slow_cap;
    int K;
    int npanel, panel_rows, warps, panels_per_range, nranges;
    int64_t own_range_stride;   // 0, or n_slots when owners are ranked per range (own_id is [nranges][n_slots])
    int64_t own_offB, oth_offB; // lanes geometry: start of plane B in the two tables
};

// append one underflowed nonzero to the sweep's queue (dropped beyond slow_cap: the fix-up kernel
// then raises the engine's sticky overflow flag)
__device__ __forceinline__ void slow_enqueue(const SweepArgs &A, int own, int oth_global, double y)
{
    const unsigned long long at = atomicAdd(A.slow_count, 1ULL);
    if (at < A.slow_cap) A.slow_queue[at] = make_int4(own, oth_global, __double2hiint(y), __double2loint(y));
}
// redo the queued nonzeros of one sweep direction, or of both (B != nullptr) with one launch
int launch_slow_fixup(cudaStream_t s, const SweepArgs &A, const SweepArgs *B, int *overflow_flag);

int launch_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream);
// one-lane-per-owner kernels (sweep_lanes.cu)
bool lanes_supported(int K);
int lanes_default_warps(int K);
int launch_lane_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream);
inline size_t lane_sweep_smem_bytes(int KP, int panel_rows) { return (size_t)panel_rows * KP * 8 + 16 + 16 * 8; }
int lanes_max_panel_rows(int K);
// fp32 one-lane-per-owner kernels (sweep_f32.cu): tables of 32 (K <= 32) or 64 floats per row
int f32_row_floats(int K);
int f32_default_warps(int K);
int f32_max_panel_rows(int K);
size_t f32_sweep_smem_bytes(int K, int panel_rows, int warps);
int launch_f32_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream);
size_t sweep_smem_bytes(int K, int panel_rows);
int max_panel_rows(int K, int ctas_per_sm);

// ------------------------------------------------ dense / per-nnz kernels ---
int launch_prep_side(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *shp, const double *rte,
                     double *elog, double *E, double *colsum /* K, accumulated; may be null */);
int launch_ex_table(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *shp, const double *rte, double *X);
int launch_fold(cudaStream_t s, int64_t n, int K, const TabGeom &tg, const double *E, const double *acc,
                const double *direct, double *out);
int launch_finalize(cudaStream_t s, int64_t n, int K, const TabGeom &tg, double prior_shape, double prior_rate,
                    const double *folded /* n x K, or null */, const double *E, const double *acc,
                    const double *direct, const double *other_colsum, const double *cap_shp,
                    double *cap_rte, double *shp, double *rte, double *elog, double *Etab,
                    double *colsum_out);
int launch_literal(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                   const int32_t *data, const double *elog_t, const double *elog_b,
                   double *xphi_out, double *direct_t, double *direct_b);
int launch_random_phi(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                      const int32_t *data, uint64_t seed, int64_t row_offset, double *direct_t,
                      double *direct_b);
int launch_scatter_xphi(cudaStream_t s, int64_t nnz, int K, const double *xphi, const int32_t *keep,
                        double *out);
int launch_llh_pointwise(cudaStream_t s, int64_t nnz, int K, const int32_t *row, const int32_t *col,
                         const int32_t *data, const double *ts, const double *tr, const double *bs,
                         const double *br, double *out);
int launch_lgamma_sum(cudaStream_t s, int64_t nnz, const int32_t *data, double *partials, int nblk,
                      double *out);
int launch_sum_partials(cudaStream_t s, const double *partials, int n, double *out);
int launch_pack_loss(cudaStream_t s, const double *llh_sum, double lgamma_sum, double nnz, double *out2);
int launch_validate_coo(cudaStream_t s, int64_t nnz, const int32_t *row, const int32_t *col,
                        const int32_t *data, int64_t C, int64_t G, int *flag);
int launch_fill(cudaStream_t s, double *p, int64_t n, double v);
int launch_psi(cudaStream_t s, int64_t n, const double *x, double *out, int which);
int launch_rate_update(cudaStream_t s, int64_t n, int K, const double *pshp, const double *prte,
                       const double *colsum, double *out);
int launch_colsum_ex(cudaStream_t s, int64_t m, int K, const double *shp, const double *rte,
                     double *colsum);
int launch_capacity_rate(cudaStream_t s, int64_t n, int K, const double *shp, const double *rte,
                         double prior, double *out);

}  // namespace schpf
