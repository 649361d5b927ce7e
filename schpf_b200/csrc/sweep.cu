// The hot kernel: one "sweep" over all nonzeros of the count matrix.
//
// Replaces, per CAVI iteration, the reference's compute_Xphi_data
// (hpf_numba.py:55-114) fused with ONE of its two compute_loading_shape_update
// scatters (hpf_numba.py:129-156): the cells-own sweep produces the theta shape
// sums, the genes-own sweep the beta shape sums.  The (nnz x K) Xphi array of
// the reference is never materialised.  In SWEEP_LLH mode the same traversal
// evaluates compute_pois_llh (hpf_numba.py:25-51) summed over nonzeros.
//
// Maths: with Elog = psi(shape) - log(rate) and the factored
// tables Et[c,k] = exp(Elog_t[c,k] - max_k), Eb[g,k] likewise,
//     phi_ik = Et[c,k] Eb[g,k] / s_i,   s_i = sum_k Et[c,k] Eb[g,k]
// which is the reference's softmax (:98-112) with the max subtraction moved
// out of the per-nonzero loop.  For an OWNER row o and its nonzeros i
//     sum_i y_i phi_ik = E_own[o,k] * sum_i (y_i / s_i) E_oth[oth_i, k]
// so a lane keeps E_own[o,:] and the running sum in registers, streams the
// OTHER side's rows out of a shared-memory panel, and needs no atomics per
// nonzero.  If s_i underflows (rows whose maxima sit on different factors by
// > ~640 nats) the nonzero is redone in log space exactly as the reference
// does and added to a separate `direct` accumulator.
//
// Mapping (B200): a CTA owns W*16 owners (one per lane PAIR; lane h of a pair
// holds the 16-byte units {2j+h} of the K-row), walks a range of panels of the
// other axis; each panel is brought into shared memory with one 1-D bulk
// async copy (cp.async.bulk -> UBLKCP) completing on an mbarrier; the entry
// stream is a per-(warp, panel) sliced-ELL block, two steps per 128-bit load.  A loop trip takes
// two stream elements (four steps) from two register sets that are reloaded a whole trip ahead
// of their use; at K=20 the four steps are one straight-line block (12-warp CTAs), otherwise two
// blocks of two (as many warps as the registers allow).  The two lanes' partial dot products
// are summed by an fp64 MMA against a 0/1 matrix (DMMA.8x8x4) rather than by shuffles, which
// keeps that traffic off the shared-memory pipe the row loads saturate.
#include "common.cuh"

namespace schpf {

namespace {

constexpr int SWEEP_MAX_WARPS = 16;

// build-time experiment knobs (python -m schpf_b200.build reads SCHPF_NVCC_FLAGS)
#ifndef SWEEP_INTERLEAVE
#define SWEEP_INTERLEAVE 1
#endif
#ifndef SWEEP_W20
#define SWEEP_W20 12
#endif
#ifndef SWEEP_WMID
#define SWEEP_WMID 12     // warps per CTA for 20 < KP <= 32
#endif
#ifndef SWEEP_MINCTA20
#define SWEEP_MINCTA20 1
#endif
#ifndef SWEEP_UNROLL2
#define SWEEP_UNROLL2 1
#endif
#ifndef SWEEP_STEPS4
#define SWEEP_STEPS4 1    // four steps (both stream elements of a trip) as one block, for SWEEP_S4_LO <= KP <= SWEEP_S4_HI
#endif
#ifndef SWEEP_S4_LO
#define SWEEP_S4_LO 20    // measured at KP=20 only (12 warps x 168 registers); the other K are next round's sweep
#endif
#ifndef SWEEP_S4_HI
#define SWEEP_S4_HI 20
#endif
#ifndef SWEEP_WSMALL
#define SWEEP_WSMALL 16   // warps per CTA for KP <= 16
#endif
#ifndef SWEEP_WBIG
#define SWEEP_WBIG 8      // warps per CTA for KP > 32
#endif
#ifndef SWEEP_DMMA
#define SWEEP_DMMA 1      // 1: the lane-pair sum of the partial dot products is an fp64 MMA, not shuffles
#endif
#ifndef SWEEP_RCP4
#define SWEEP_RCP4 1      // 1: y/s with the product folded into the Newton step (one dependent op fewer)
#endif

// A lane keeps 4*KP/2 doubles live (owner row, accumulators, the streamed rows of the two
// steps in flight) plus ~45 registers of addressing.  One CTA per SM owns the whole shared
// memory (the longer the per-owner lists of a panel, the less SELL padding: measured 24 % with
// 720-row panels, 17 % with 1448), and the CTA is as many warps as the register file allows
// without spilling (ptxas -v; registers are granted per SM sub-partition, so warps come in 4s).
// K=20 is the exception that was measured: 12 warps x 168 registers with FOUR steps in flight
// per lane pair (3.18 ms per sweep pair on cfg-3) beat 16 warps x 128 registers with two (3.29):
// the kernel is bound by latencies, and 12 x 4 independent steps hide more than 16 x 2.
__host__ __device__ constexpr int sweep_min_ctas(int KP) { return KP == 20 ? SWEEP_MINCTA20 : 1; }
__host__ __device__ constexpr int sweep_max_warps(int KP)
{
    return KP == 20 ? SWEEP_W20 : KP <= 16 ? SWEEP_WSMALL : KP <= 32 ? SWEEP_WMID : SWEEP_WBIG;
}

template <int KP>
struct SweepCfg {
    static constexpr int ST = stride_of_kp(KP);
    static constexpr int U = KP / 4;      // 16-byte units per lane
    static constexpr int D = KP / 2;      // doubles per lane
    static constexpr int MIN_CTAS = sweep_min_ctas(KP);
    static constexpr int MAX_WARPS = sweep_max_warps(KP);
    // measured on cfg-3 (K=20): 3.36 ms per sweep pair unrolled against 3.50 rolled; above KP=52
    // the unrolled body spills inside the loop (ptxas -v), so those stay rolled
    static constexpr bool UNROLL2 = SWEEP_UNROLL2 && KP <= 52;
    static constexpr bool STEPS4 = SWEEP_STEPS4 && KP >= SWEEP_S4_LO && KP <= SWEEP_S4_HI;
};

template <int N> struct Steps { static constexpr int value = N; };

// two steps of a lane pair as fetched from the entry stream, and their decoded form
template <bool PACKED> struct EntryPair;
template <> struct EntryPair<false> {
    int4 v;
    __device__ __forceinline__ static EntryPair zero() { return {make_int4(0, 0, 0, 0)}; }
    __device__ __forceinline__ static EntryPair load(const void *base, int64_t idx)
    {
        return {ld_stream_int4(reinterpret_cast<const int4 *>(base) + idx)};
    }
    __device__ __forceinline__ int row(int e) const { return (e ? v.z : v.x) & 0x7fffffff; }
    __device__ __forceinline__ int count(int e) const { return e ? v.w : v.y; }
    __device__ __forceinline__ bool pad(int e) const { return (e ? v.z : v.x) < 0; }
};
template <> struct EntryPair<true> {
    int2 v;
    __device__ __forceinline__ static EntryPair zero() { return {make_int2(0, 0)}; }
    __device__ __forceinline__ static EntryPair load(const void *base, int64_t idx)
    {
        return {ld_stream_int2(reinterpret_cast<const int2 *>(base) + idx)};
    }
    __device__ __forceinline__ int row(int e) const { return (e ? v.y : v.x) & ((1 << PACKED_ROW_BITS) - 1); }
    __device__ __forceinline__ int count(int e) const
    {
        return ((e ? v.y : v.x) >> PACKED_ROW_BITS) & ((1 << PACKED_COUNT_BITS) - 1);
    }
    __device__ __forceinline__ bool pad(int e) const { return (e ? v.y : v.x) < 0; }
};

template <int KP, int MODE, bool PACKED>
__global__ void __launch_bounds__(SweepCfg<KP>::MAX_WARPS * 32, SweepCfg<KP>::MIN_CTAS)
sweep_kernel(const SweepArgs A)
{
    using Ent = EntryPair<PACKED>;
    using Cfg = SweepCfg<KP>;
    constexpr int ST = Cfg::ST, U = Cfg::U, D = Cfg::D;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *panel = reinterpret_cast<double *>(smem_raw);
    const uint32_t panel_bytes = (uint32_t)A.panel_rows * ST * 8u;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + panel_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 1, h = lane & 1;
    const int b = blockIdx.x / A.nranges, r = blockIdx.x - b * A.nranges;
    const int p0 = r * A.panels_per_range;
    const int p1 = min(p0 + A.panels_per_range, A.npanel);
    const int wg = b * A.warps + warp;                       // global warp index
    const int own = A.own_id[(int64_t)wg * GROUPS_PER_WARP + q];

    double a[D], acc[D];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        double2 v = make_double2(0.0, 0.0);
        if (own >= 0) v = reinterpret_cast<const double2 *>(A.own_tab + (int64_t)own * ST)[2 * j + h];
        a[2 * j] = v.x;
        a[2 * j + 1] = v.y;
        acc[2 * j] = 0.0;
        acc[2 * j + 1] = 0.0;
    }
    double llh = 0.0;
#if SWEEP_DMMA
    // B operand of pair_sum_mma: B[c][n] = 1 where column c of A (lane 4r+c) is in the lane pair
    // that receives D[r][n]; this lane holds B[lane % 4][lane / 4]
    const double pair_sel = (((lane >> 1) & 1) == (lane >> 4)) ? 1.0 : 0.0;
#endif

    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
    const uint32_t panel_s = smem_u32(panel);
    const int64_t *sp = A.seg_ptr + (int64_t)wg * (A.npanel + 1);

    // this warp's first panel: segment bounds and the first two entry pairs
    int64_t i0 = 0, i1 = 0;
    Ent cur = Ent::zero(), nxt = cur;
    int64_t ep = 0;                     // element index of this lane pair's current int4 / int2
    if (p0 < p1) {
        i0 = sp[p0];
        i1 = sp[p0 + 1];
        ep = i0 * GROUPS_PER_WARP + q;
        if (i0 < i1) cur = Ent::load(A.entries, ep);
        if (i0 + 1 < i1) nxt = Ent::load(A.entries, ep + GROUPS_PER_WARP);
    }

    for (int p = p0; p < p1; ++p) {
        if (p > p0) __syncthreads();     // every warp is done with the previous panel
        if (tid == 0) {
            mbar_expect_tx(mbar, panel_bytes);
            bulk_g2s(panel, A.oth_tab + (int64_t)p * A.panel_rows * ST, panel_bytes, mbar);
        }
        // bounds of the NEXT panel's segment, fetched while this panel is in flight
        int64_t n0 = 0, n1 = 0;
        if (p + 1 < p1) {
            n0 = i1;                     // segments of consecutive panels are contiguous
            n1 = sp[p + 2];
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;

        // the NS steps one lane pair takes from NS/2 stream elements, as one straight-line block
        auto process = [&](auto ns_tag, const int *ex, const int *ey, const bool *epad) {
            constexpr int NS = decltype(ns_tag)::value;
            double s[NS];
            bool slow = false;
#if SWEEP_INTERLEAVE
            // two steps per iteration as one straight-line block: the two independent dot
            // products / divisions interleave (more ILP, 2*D more live registers)
            double bv[NS][D];
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                // pad entries (bit 31, count 0) point at a row of a free bank group
                const uint32_t addr = panel_s + (uint32_t)ex[e] * (ST * 8) + h * 16;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const double2 v = lds_f64x2(addr + j * 32);
                    bv[e][2 * j] = v.x;
                    bv[e][2 * j + 1] = v.y;
                }
            }
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                double s0 = a[0] * bv[e][0], s1 = a[1] * bv[e][1];
#pragma unroll
                for (int k = 2; k < D; k += 2) {
                    s0 = fma(a[k], bv[e][k], s0);
                    s1 = fma(a[k + 1], bv[e][k + 1], s1);
                }
                s[e] = s0 + s1;
            }
#if SWEEP_DMMA
#pragma unroll
            for (int e = 0; e < NS; ++e) s[e] = pair_sum_mma(s[e], pair_sel);
#else
#pragma unroll
            for (int e = 0; e < NS; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 1);
#endif
            if (MODE == SWEEP_SHAPE) {
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    const double y = (double)ey[e];
                    const bool ok = s[e] > TINY_NORMALIZER;
#if SWEEP_RCP4
                    const double w = ok ? div_pos_folded(y, s[e]) : 0.0;
#else
                    const double w = ok ? div_pos(y, s[e]) : 0.0;
#endif
#pragma unroll
                    for (int k = 0; k < D; ++k) acc[k] = fma(w, bv[e][k], acc[k]);
                    slow |= (!ok && ey[e] != 0);
                }
            }
#else
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                double b1[D];
                const uint32_t addr = panel_s + (uint32_t)ex[e] * (ST * 8) + h * 16;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const double2 v = lds_f64x2(addr + j * 32);
                    b1[2 * j] = v.x;
                    b1[2 * j + 1] = v.y;
                }
                double s0 = a[0] * b1[0], s1 = a[1] * b1[1];
#pragma unroll
                for (int k = 2; k < D; k += 2) {
                    s0 = fma(a[k], b1[k], s0);
                    s1 = fma(a[k + 1], b1[k + 1], s1);
                }
                s[e] = s0 + s1;
                s[e] += __shfl_xor_sync(0xffffffffu, s[e], 1);
                if (MODE == SWEEP_SHAPE) {
                    const double y = (double)ey[e];
                    const bool ok = s[e] > TINY_NORMALIZER;
                    const double w = ok ? div_pos(y, s[e]) : 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) acc[k] = fma(w, b1[k], acc[k]);
                    slow |= (!ok && ey[e] != 0);
                }
            }
#endif
            if (MODE == SWEEP_SHAPE) {
                if (slow && own >= 0 && h == 0) {
                    // underflowed nonzeros are redone in log space (hpf_numba.py:98-112) by
                    // slow_fixup_kernel after the sweep; one lane of the pair queues them
#pragma unroll
                    for (int e = 0; e < NS; ++e)
                        if (!(s[e] > TINY_NORMALIZER) && ey[e] != 0)
                            slow_enqueue(A, own, p * A.panel_rows + ex[e], (double)ey[e]);
                }
            } else {
                // hpf_numba.py:49-50 without the lgamma term (a constant of the data).  Both lanes of
                // a pair hold both normalisers: lane 0 finishes step 0, lane 1 step 1 (one log each).
#pragma unroll
                for (int e2 = 0; e2 < NS; e2 += 2) {
                    const double sm = h ? s[e2 + 1] : s[e2];
                    const bool padm = h ? epad[e2 + 1] : epad[e2];
                    const double ym = (double)(h ? ey[e2 + 1] : ey[e2]);
                    const double v = fma(ym, log(sm), -sm);
                    if (!padm) llh += v;
                }
            }
        };
        if constexpr (Cfg::STEPS4) {
            // both stream elements of a trip as ONE block of four steps: twice the independent
            // work per warp (fewer warps fit: 4 * D more live registers)
            const int64_t eq = q;
            int64_t i = i0;
            for (; i + 1 < i1; i += 2) {
                const int ex[4] = {cur.row(0), cur.row(1), nxt.row(0), nxt.row(1)};
                const int ey[4] = {cur.count(0), cur.count(1), nxt.count(0), nxt.count(1)};
                const bool epad[4] = {cur.pad(0), cur.pad(1), nxt.pad(0), nxt.pad(1)};
                if (i + 2 < i1) cur = Ent::load(A.entries, (i + 2) * GROUPS_PER_WARP + eq);
                if (i + 3 < i1) nxt = Ent::load(A.entries, (i + 3) * GROUPS_PER_WARP + eq);
                process(Steps<4>{}, ex, ey, epad);
            }
            if (i < i1) {
                const int ex[2] = {cur.row(0), cur.row(1)}, ey[2] = {cur.count(0), cur.count(1)};
                const bool epad[2] = {cur.pad(0), cur.pad(1)};
                process(Steps<2>{}, ex, ey, epad);
            }
        } else if constexpr (Cfg::UNROLL2) {
            // two stream elements per trip, each in its own registers: no rotation moves, and a
            // reload is two iterations ahead of its use (cur holds element i, nxt element i + 1)
            const int64_t eq = q;
            int64_t i = i0;
            for (; i + 1 < i1; i += 2) {
                {
                    const int ex[2] = {cur.row(0), cur.row(1)}, ey[2] = {cur.count(0), cur.count(1)};
                    const bool epad[2] = {cur.pad(0), cur.pad(1)};
                    if (i + 2 < i1) cur = Ent::load(A.entries, (i + 2) * GROUPS_PER_WARP + eq);
                    process(Steps<2>{}, ex, ey, epad);
                }
                {
                    const int ex[2] = {nxt.row(0), nxt.row(1)}, ey[2] = {nxt.count(0), nxt.count(1)};
                    const bool epad[2] = {nxt.pad(0), nxt.pad(1)};
                    if (i + 3 < i1) nxt = Ent::load(A.entries, (i + 3) * GROUPS_PER_WARP + eq);
                    process(Steps<2>{}, ex, ey, epad);
                }
            }
            if (i < i1) {
                const int ex[2] = {cur.row(0), cur.row(1)}, ey[2] = {cur.count(0), cur.count(1)};
                const bool epad[2] = {cur.pad(0), cur.pad(1)};
                process(Steps<2>{}, ex, ey, epad);
            }
        } else {
            for (int64_t i = i0; i < i1; ++i) {
                ep += GROUPS_PER_WARP;
                Ent nxt2 = nxt;
                if (i + 2 < i1) nxt2 = Ent::load(A.entries, ep + GROUPS_PER_WARP);
                const int ex[2] = {cur.row(0), cur.row(1)}, ey[2] = {cur.count(0), cur.count(1)};
                const bool epad[2] = {cur.pad(0), cur.pad(1)};
                process(Steps<2>{}, ex, ey, epad);
                cur = nxt;
                nxt = nxt2;
            }
        }
        // first entries of the next panel: issued before the barrier so their latency overlaps it
        i0 = n0;
        i1 = n1;
        ep = i0 * GROUPS_PER_WARP + q;
        if (i0 < i1) cur = Ent::load(A.entries, ep);
        if (i0 + 1 < i1) nxt = Ent::load(A.entries, ep + GROUPS_PER_WARP);
    }

    if (MODE == SWEEP_SHAPE) {
        if (own >= 0) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const int k = 4 * j + 2 * h + d;
                    if (k < A.K) atomicAdd(A.acc + (int64_t)own * A.K + k, acc[2 * j + d]);
                }
            }
        }
    } else {
        __shared__ double red[SWEEP_MAX_WARPS];
        llh = warp_sum(llh);
        if (lane == 0) red[warp] = llh;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < A.warps; ++w) t += red[w];
            A.partial[blockIdx.x] = t;
        }
    }
}

template <int KP, int MODE, bool PACKED>
int launch_packed(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    const size_t smem = (size_t)L.panel_rows * SweepCfg<KP>::ST * 8 + 16;
    static bool configured = false;   // per instantiation
    static size_t configured_smem = 0;
    if (!configured || smem > configured_smem) {
        CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<KP, MODE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<KP, MODE, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured = true;
        configured_smem = smem;
    }
    const int grid = L.nblocks * L.nranges;
    if (grid <= 0) return SCHPF_OK;
    sweep_kernel<KP, MODE, PACKED><<<grid, L.warps * 32, smem, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("sweep_kernel<KP=%d,mode=%d> launch (grid %d, block %d, smem %zu) -> %s", KP, MODE, grid,
                  L.warps * 32, smem, cudaGetErrorString(e));
        return SCHPF_ERR_CUDA;
    }
    return SCHPF_OK;
}

template <int KP, int MODE>
int launch_one(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    return L.packed ? launch_packed<KP, MODE, true>(L, args, stream) : launch_packed<KP, MODE, false>(L, args, stream);
}

template <int MODE>
int dispatch_kp(int KP, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    switch (KP) {
#define CASE_KP(v) \
    case v: return launch_one<v, MODE>(L, args, stream);
        CASE_KP(4) CASE_KP(8) CASE_KP(12) CASE_KP(16) CASE_KP(20) CASE_KP(24) CASE_KP(28) CASE_KP(32)
        CASE_KP(36) CASE_KP(40) CASE_KP(44) CASE_KP(48) CASE_KP(52) CASE_KP(56) CASE_KP(60) CASE_KP(64)
#undef CASE_KP
        default:
            set_error("nfactors padded to %d is outside the instantiated range (K <= %d)", KP,
                      SCHPF_MAX_FACTORS);
            return SCHPF_ERR_ARG;
    }
}

}  // namespace

size_t sweep_smem_bytes(int K, int panel_rows)
{
    return (size_t)panel_rows * stride_of_kp(kp_of(K)) * 8 + 16;
}

// Largest panel (rows of the other axis) such that `ctas_per_sm` CTAs fit in the
// 227 KB of shared memory an SM can give (1 KB per CTA is reserved by the system).
int max_panel_rows(int K, int ctas_per_sm)
{
    const int ST = stride_of_kp(kp_of(K));
    // 1 KB per CTA is reserved by the system; 256 B cover the mbarrier and the static reduction scratch
    const size_t per_cta = (size_t)(228 * 1024) / ctas_per_sm - 1024 - 256;
    int rows = (int)(per_cta / ((size_t)ST * 8));
    rows &= ~3;
    if (rows > 4096) rows = 4096;   // 12-bit local index in the sort key
    return rows;
}

int sweep_ctas_per_sm(int K) { return sweep_min_ctas(kp_of(K)); }
int sweep_default_warps(int K) { return sweep_max_warps(kp_of(K)); }

int launch_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    const int KP = kp_of(K);
    if (L.warps > sweep_max_warps(KP) || L.warps < 1) {
        set_error("warps_per_cta must be in [1, %d] for K=%d", sweep_max_warps(KP), K);
        return SCHPF_ERR_ARG;
    }
    if (mode == SWEEP_SHAPE) return dispatch_kp<SWEEP_SHAPE>(KP, L, args, stream);
    return dispatch_kp<SWEEP_LLH>(KP, L, args, stream);
}

}  // namespace schpf
