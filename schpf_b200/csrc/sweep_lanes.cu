// One-lane-per-owner sweep (round 2): the same two-pass factored-softmax sweep as sweep.cu
// (compute_Xphi_data fused with one compute_loading_shape_update, hpf_numba.py:55-114,129-156;
// SWEEP_LLH: compute_pois_llh, hpf_numba.py:25-51), re-mapped so that ONE lane owns one owner
// row.  Why: the lane-pair kernel spends ~54 warp instructions per 16 nonzeros at K=20 and its
// issue slots (81 % of the shared-memory time), fp64 pipe and LSU are all within 1.5x of each
// other, so none can be saturated.  One lane per owner covers 32 nonzeros per warp instruction:
// no pair sum (DMMA / shuffles), one reciprocal and one decode per nonzero instead of two, and
// the shared-memory pipe is left as the only busy unit.
//
// Shared-memory layout that makes this conflict-free.  A 128-bit LDS is served one quarter warp
// (8 lanes x 16 B) per wavefront, conflict-free iff the 8 lanes hit 8 different 16-byte bank
// groups.  The K-row is therefore SPLIT into two planes:
//   plane A: the first 16*NA doubles, row stride 128*NA bytes.  Every row starts at bank group 0,
//            so lane i (of its quarter warp) reads the 16-byte unit (t ^ i) at time t: 8 lanes,
//            8 different units, whatever rows they read.  No schedule, no constraint.  The owner
//            row and the accumulators sit in registers in the same per-lane rotated order.
//   plane B: the last 4 doubles (REM = 1, e.g. K = 20 = 16 + 4), row stride 32 bytes: row r
//            covers bank groups 2*(r mod 4) + {0,1}.  Even lanes read its halves in order
//            (0,1), odd lanes (1,0); the layout's 4x4 edge colouring (layout.cu) guarantees that
//            the 4 even lanes of a quarter warp read rows of 4 different residues mod 4 at every
//            step, and likewise the 4 odd lanes -> 8 different bank groups.
// K <= 16 is NA=1, REM=0 (zero columns beyond K); {17..20} NA=1, REM=1; {29..32} NA=2, REM=0: with REM = 0 the
// stream needs no bank schedule at all.
//
// Epilogue: the accumulators are staged in shared memory in natural order and added to the
// global accumulator rows by the TMA engine (cp.reduce.async.bulk ... add.f64 -> UBLKRED), one
// bulk reduction per owner row instead of K `RED.ADD.F64` per owner and lane.
#include "common.cuh"

namespace schpf {

namespace {

#ifndef LANES_W16
#define LANES_W16 12      // warps per CTA, KP = 16 (168 registers, no spill; 16 warps x 128 registers spill 40 bytes: 1.98 vs 1.84 ms per pair)
#endif
#ifndef LANES_W20
#define LANES_W20 12      // KP = 20
#endif
#ifndef LANES_W32
#define LANES_W32 8       // KP = 32
#endif
#ifndef LANES_NS16
#define LANES_NS16 2      // steps processed as one straight-line block
#endif
#ifndef LANES_NS20
#define LANES_NS20 2
#endif
#ifndef LANES_NS32
#define LANES_NS32 2      // measured K=30, cfg-3: two steps per block 4.00 ms per sweep pair, one 4.43 (254 registers, 8 warps)
#endif
#ifndef LANES_BULK_RED
#define LANES_BULK_RED 1
#endif
#ifndef LANES_CTAS
#define LANES_CTAS 1      // CTAs per SM: 2 = half the warps and half the panel each (fills and epilogues of one hide behind the other)
#endif
// measurement-only builds (never shipped): 1 = row loads without the arithmetic, 2 = arithmetic
// without the row loads -- how much of the sweep is LSU time, how much fp64 time, how much overlaps
#ifndef LANES_EXP
#define LANES_EXP 0
#endif
#ifndef LANES_PREFETCH
#define LANES_PREFETCH 1  // 1: next trip's stream elements requested at the top of a trip; 0: reloaded after last use
#endif
#ifndef LANES_L2_AHEAD
#define LANES_L2_AHEAD 8  // trips ahead of which the stream lines are asked into L2
#endif

struct Steps1 { static constexpr int value = 1; };
struct Steps2 { static constexpr int value = 2; };

template <int NA, int REM>
struct LaneCfg {
    static constexpr int KA = 16 * NA, KB = 4 * REM, KP = KA + KB;
    static constexpr int WARPS = (KP <= 16 ? LANES_W16 : KP <= 20 ? LANES_W20 : LANES_W32) / LANES_CTAS;
    static constexpr int NS = KP <= 16 ? LANES_NS16 : KP <= 20 ? LANES_NS20 : LANES_NS32;
};

__device__ __forceinline__ void bulk_red_add_f64(double *dst_gmem, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(src_smem), "r"(bytes)
                 : "memory");
}

// Stream entry of the one-lane kernels: {row, field}.  `row` is the plain panel-local row (pads
// point at a real row, so it needs no masking); `field` carries the count and, in bit 31, the pad
// flag.  YHI: the low 31 bits are the HIGH WORD of (double)count -- counts below 2^21 have an
// all-zero low word, so y is built with no conversion instruction (a pad becomes -0.0, which adds
// nothing); otherwise the field is the integer count.
template <bool YHI>
__device__ __forceinline__ double entry_count(int field)
{
    // (the 2^52 + count trick -- one DADD instead of I2F.F64 -- was measured: it costs two more registers
    // per step and the 168-register K=20 kernel then spills 176 bytes in the loop: 5.6 ms per sweep pair)
    return YHI ? __hiloint2double(field, 0) : (double)(field & 0x7fffffff);
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// One stream element (two steps of a lane) decoded to {row, count field, row, count field}.
// ENC 0: wide, integer count; 1: wide, count as the high word of its double; 2: packed 4-byte
// entries pad<<31 | count<<12 | row (common.cuh pack_entry) -- half the stream bytes; the count
// field comes out as pad<<31 | count like the wide integer encoding.
constexpr int ENC_WIDE = 0, ENC_YHI = 1, ENC_PACKED = 2;
// Packed elements stay RAW in their two registers ({w0, w0, w1, w1}: the compiler keeps one copy) until
// process() decodes them at the point of use, so the loads issued a trip ahead hold two registers, not four.
template <int ENC>
__device__ __forceinline__ int4 ld_elem(const char *p)
{
    if (ENC == ENC_PACKED) {
        const int2 w = ld_stream_int2(reinterpret_cast<const int2 *>(p));
        return make_int4(w.x, w.x, w.y, w.y);
    }
    return ld_stream_int4(reinterpret_cast<const int4 *>(p));
}
template <int ENC>
__device__ __forceinline__ int elem_row(int x)
{
    return ENC == ENC_PACKED ? (x & ((1 << PACKED_ROW_BITS) - 1)) : x;
}
template <int ENC>
__device__ __forceinline__ int elem_field(int y)       // pad<<31 | count
{
    return ENC == ENC_PACKED ? ((y & (int)0x80000000) | ((y >> PACKED_ROW_BITS) & ((1 << PACKED_COUNT_BITS) - 1))) : y;
}

template <int NA, int REM, int MODE, int ENC>
__global__ void __launch_bounds__(LaneCfg<NA, REM>::WARPS * 32, LANES_CTAS)
lane_sweep_kernel(const SweepArgs A)
{
    constexpr bool YHI = ENC == ENC_YHI;
    constexpr int ES = ENC == ENC_PACKED ? 8 : 16;      // bytes of one stream element
    using Cfg = LaneCfg<NA, REM>;
    constexpr int KA = Cfg::KA, KB = Cfg::KB, KP = Cfg::KP;
    constexpr int ROWA = KA * 8;                  // bytes per plane-A row (128 * NA)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t bytesA = (uint32_t)A.panel_rows * ROWA;
    const uint32_t bytesB = (uint32_t)A.panel_rows * (KB * 8);
    // the panel region also stages the accumulators of the epilogue (one KP-row per thread)
    const uint32_t region = (uint32_t)max(A.panel_rows, A.warps * 32) * (KP * 8);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + region);
    double *red = reinterpret_cast<double *>(smem_raw + region + 16);   // LLH: per-warp sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i8 = lane & 7, par = lane & 1;
    const int b = blockIdx.x / A.nranges, r = blockIdx.x - b * A.nranges;
    const int p0 = r * A.panels_per_range;
    const int p1 = min(p0 + A.panels_per_range, A.npanel);
    const int wg = b * A.warps + warp;                        // global warp index
    // lane -> slot of its warp (inverse of layout.cu's stream_pos): the even / odd lanes of quarter
    // warp qw are the scheduling groups 2*qw / 2*qw + 1, four owners each
    const int slot = (((lane >> 3) * 2 + (lane & 1)) << 2) + ((lane & 7) >> 1);
    const int own = A.own_id[(int64_t)r * A.own_range_stride + (int64_t)wg * 32 + slot];

    // owner row in the per-lane rotated order: slot t of block n holds unit (t ^ i8)
    double a[KP], acc[KP];
#pragma unroll
    for (int n = 0; n < NA; ++n)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            double2 v = make_double2(0.0, 0.0);
            if (own >= 0)
                v = *reinterpret_cast<const double2 *>(A.own_tab + (int64_t)own * KA + 16 * n + 2 * (t ^ i8));
            a[16 * n + 2 * t] = v.x;
            a[16 * n + 2 * t + 1] = v.y;
        }
    if (REM) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            double2 v = make_double2(0.0, 0.0);
            if (own >= 0)
                v = *reinterpret_cast<const double2 *>(A.own_tab + A.own_offB + (int64_t)own * KB + 2 * (t ^ par));
            a[KA + 2 * t] = v.x;
            a[KA + 2 * t + 1] = v.y;
        }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.0;
    double llh = 0.0;

    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
    const uint32_t panelA_s = smem_u32(smem_raw) + ((uint32_t)i8 << 4);          // + this lane's unit offset
    const uint32_t panelB_s = smem_u32(smem_raw) + bytesA + ((uint32_t)par << 4);
    const int64_t *sp = A.seg_ptr + (int64_t)wg * (A.npanel + 1);

    // NS steps of this lane as one straight-line block.  `oth0` = global row of the panel's row 0.
    auto process = [&](auto ns_tag, const int *ex_raw, const int *ey_raw, int oth0) {
        constexpr int NS = decltype(ns_tag)::value;
        double bv[NS][KP], s[NS];
        bool slow = false;
        int ex[NS], ey[NS];
#pragma unroll
        for (int e = 0; e < NS; ++e) {
            ex[e] = elem_row<ENC>(ex_raw[e]);
            ey[e] = elem_field<ENC>(ey_raw[e]);
        }
#if LANES_EXP == 2
#pragma unroll
        for (int e = 0; e < NS; ++e)
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                bv[e][k] = a[(k + 1) % KP];
                asm volatile("" : "+d"(bv[e][k]));
            }
#else
#pragma unroll
        for (int e = 0; e < NS; ++e) {
            const uint32_t addrA = panelA_s + (uint32_t)ex[e] * ROWA;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t at = addrA ^ ((uint32_t)t << 4);
#pragma unroll
                for (int n = 0; n < NA; ++n) {
                    const double2 v = lds_f64x2(at + n * 128);
                    bv[e][16 * n + 2 * t] = v.x;
                    bv[e][16 * n + 2 * t + 1] = v.y;
                }
            }
            if (REM) {
                const uint32_t addrB = panelB_s + (uint32_t)ex[e] * 32;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const double2 v = lds_f64x2(addrB ^ ((uint32_t)t << 4));
                    bv[e][KA + 2 * t] = v.x;
                    bv[e][KA + 2 * t + 1] = v.y;
                }
            }
        }
#endif
#if LANES_EXP == 1
        {
            int x = 0;
#pragma unroll
            for (int e = 0; e < NS; ++e)
#pragma unroll
                for (int k = 0; k < KP; ++k) x ^= __double2hiint(bv[e][k]) ^ __double2loint(bv[e][k]);
            if (x == 0x12345677) llh += 1.0;
            return;
        }
#endif
#pragma unroll
        for (int e = 0; e < NS; ++e) {
            double s0 = a[0] * bv[e][0], s1 = a[1] * bv[e][1], s2 = a[2] * bv[e][2], s3 = a[3] * bv[e][3];
#pragma unroll
            for (int k = 4; k < KP; k += 4) {
                s0 = fma(a[k], bv[e][k], s0);
                s1 = fma(a[k + 1], bv[e][k + 1], s1);
                s2 = fma(a[k + 2], bv[e][k + 2], s2);
                s3 = fma(a[k + 3], bv[e][k + 3], s3);
            }
            s[e] = (s0 + s1) + (s2 + s3);
        }
        if (MODE == SWEEP_SHAPE) {
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                const double y = entry_count<YHI>(ey[e]);
                // s >= 0: comparing the high words is the test s > TINY_NORMALIZER up to the low
                // word of the threshold, and runs on the integer pipe
                const bool ok = __double2hiint(s[e]) > TINY_NORMALIZER_HI;
                const double w = ok ? div_pos_folded(y, s[e]) : 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] = fma(w, bv[e][k], acc[k]);
                slow |= !ok;
            }
            if (slow) {
                // cold: a nonzero whose factored normaliser underflowed is queued for the log-space
                // redo after the sweep (slow_fixup_kernel); pads and empty slots fall out here
#pragma unroll
                for (int e = 0; e < NS; ++e)
                    if (__double2hiint(s[e]) <= TINY_NORMALIZER_HI && (ey[e] & 0x7fffffff) != 0 && own >= 0)
                        slow_enqueue(A, own, oth0 + ex[e], entry_count<YHI>(ey[e] & 0x7fffffff));
            }
        } else {
            // hpf_numba.py:49-50 without the lgamma term (a constant of the data)
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                const double v = fma(entry_count<YHI>(ey[e] & 0x7fffffff), log(s[e]), -s[e]);
                if (ey[e] >= 0) llh += v;          // bit 31 of the count field = pad
            }
        }
    };

    for (int p = p0; p < p1; ++p) {
        if (p > p0) __syncthreads();     // every warp is done with the previous panel
        if (tid == 0) {
            mbar_expect_tx(mbar, bytesA + bytesB);
            bulk_g2s(smem_raw, A.oth_tab + (int64_t)p * A.panel_rows * KA, bytesA, mbar);
            if (REM)
                bulk_g2s(smem_raw + bytesA, A.oth_tab + A.oth_offB + (int64_t)p * A.panel_rows * KB, bytesB, mbar);
        }
        // this warp's segment of the stream: n elements (two steps each) of 32 lanes
        const int64_t i0 = sp[p];
        const int n = (int)(sp[p + 1] - i0);
        const char *ptr = reinterpret_cast<const char *>(A.entries) + (i0 * 32 + lane) * ES;
        int4 cur = make_int4(0, 0, 0, 0), nxt = cur;
        if (n > 0) cur = ld_elem<ENC>(ptr);
        if (n > 1) nxt = ld_elem<ENC>(ptr + 32 * ES);
        const int oth0 = p * A.panel_rows;
        mbar_wait(mbar, parity);
        parity ^= 1u;

        // Two stream elements per trip.  The elements of the NEXT trip are requested before this
        // trip's arithmetic starts (a whole trip = four steps of latency cover), and the lines of the
        // trip LANES_L2_AHEAD ahead are asked into L2.
        int j = 0;
        for (; j + 1 < n; j += 2, ptr += 64 * ES) {
            prefetch_l2(ptr + LANES_L2_AHEAD * 64 * ES);             // immediate offsets from the lane's own pointer:
            prefetch_l2(ptr + (LANES_L2_AHEAD * 64 + 32) * ES);      // no predicate, no extra registers
#if LANES_PREFETCH == 0
            // variant: an element is reloaded right after its last use (no copies, two steps of
            // cover: enough for an L2 hit, which the prefetch above is there to make it)
            if constexpr (Cfg::NS == 2) {
                {
                    const int ex[2] = {cur.x, cur.z}, ey[2] = {cur.y, cur.w};
                    process(Steps2{}, ex, ey, oth0);
                }
                if (j + 2 < n) cur = ld_elem<ENC>(ptr + 64 * ES);
                {
                    const int ex[2] = {nxt.x, nxt.z}, ey[2] = {nxt.y, nxt.w};
                    process(Steps2{}, ex, ey, oth0);
                }
                if (j + 3 < n) nxt = ld_elem<ENC>(ptr + 96 * ES);
                continue;
            }
#endif
            const int4 c0 = cur, c1 = nxt;
            if (j + 2 < n) cur = ld_elem<ENC>(ptr + 64 * ES);
            if (j + 3 < n) nxt = ld_elem<ENC>(ptr + 96 * ES);
            if constexpr (Cfg::NS == 2) {
                {
                    const int ex[2] = {c0.x, c0.z}, ey[2] = {c0.y, c0.w};
                    process(Steps2{}, ex, ey, oth0);
                }
                {
                    const int ex[2] = {c1.x, c1.z}, ey[2] = {c1.y, c1.w};
                    process(Steps2{}, ex, ey, oth0);
                }
            } else {
                {
                    const int ex[1] = {c0.x}, ey[1] = {c0.y};
                    process(Steps1{}, ex, ey, oth0);
                }
                {
                    const int ex[1] = {c0.z}, ey[1] = {c0.w};
                    process(Steps1{}, ex, ey, oth0);
                }
                {
                    const int ex[1] = {c1.x}, ey[1] = {c1.y};
                    process(Steps1{}, ex, ey, oth0);
                }
                {
                    const int ex[1] = {c1.z}, ey[1] = {c1.w};
                    process(Steps1{}, ex, ey, oth0);
                }
            }
        }
        if (j < n) {
            if constexpr (Cfg::NS == 2) {
                const int ex[2] = {cur.x, cur.z}, ey[2] = {cur.y, cur.w};
                process(Steps2{}, ex, ey, oth0);
            } else {
                {
                    const int ex[1] = {cur.x}, ey[1] = {cur.y};
                    process(Steps1{}, ex, ey, oth0);
                }
                {
                    const int ex[1] = {cur.z}, ey[1] = {cur.w};
                    process(Steps1{}, ex, ey, oth0);
                }
            }
        }
    }

    if (MODE == SWEEP_SHAPE) {
        const int K = A.K;
        if (LANES_BULK_RED && !(K & 1)) {
            // stage the accumulators in natural k order where the panel was, then one TMA bulk
            // reduction (add.f64) per owner row into the global accumulator
            __syncthreads();             // all warps are done reading the panel
            const uint32_t stage = smem_u32(smem_raw) + (uint32_t)tid * (KP * 8);
#pragma unroll
            for (int n = 0; n < NA; ++n)
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(stage + n * 128 + ((t ^ i8) << 4)),
                                 "d"(acc[16 * n + 2 * t]), "d"(acc[16 * n + 2 * t + 1])
                                 : "memory");
            if (REM) {
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(stage + KA * 8 + ((t ^ par) << 4)),
                                 "d"(acc[KA + 2 * t]), "d"(acc[KA + 2 * t + 1])
                                 : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (own >= 0) bulk_red_add_f64(A.acc + (int64_t)own * K, stage, (uint32_t)K * 8u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else if (own >= 0) {
#pragma unroll
            for (int n = 0; n < NA; ++n)
#pragma unroll
                for (int t = 0; t < 8; ++t)
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        const int k = 16 * n + 2 * (t ^ i8) + d;
                        if (k < K) atomicAdd(A.acc + (int64_t)own * K + k, acc[16 * n + 2 * t + d]);
                    }
            if (REM) {
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        const int k = KA + 2 * (t ^ par) + d;
                        if (k < K) atomicAdd(A.acc + (int64_t)own * K + k, acc[KA + 2 * t + d]);
                    }
            }
        }
    } else {
        llh = warp_sum(llh);
        __syncthreads();
        if (lane == 0) red[warp] = llh;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < A.warps; ++w) t += red[w];
            A.partial[blockIdx.x] = t;
        }
    }
}

template <int NA, int REM, int MODE, int YHI>
int launch_lanes_y(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    using Cfg = LaneCfg<NA, REM>;
    const size_t smem = lane_sweep_smem_bytes(Cfg::KP, L.panel_rows > L.warps * 32 ? L.panel_rows : L.warps * 32);
    static bool configured = false;   // per instantiation
    static size_t configured_smem = 0;
    if (!configured || smem > configured_smem) {
        CUDA_TRY(cudaFuncSetAttribute(lane_sweep_kernel<NA, REM, MODE, YHI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(lane_sweep_kernel<NA, REM, MODE, YHI>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
        configured_smem = smem;
    }
    const int grid = L.nblocks * L.nranges;
    if (grid <= 0) return SCHPF_OK;
    if (L.warps > Cfg::WARPS || L.warps < 1) {
        set_error("lane sweep: warps_per_cta must be in [1, %d] for KP=%d", Cfg::WARPS, Cfg::KP);
        return SCHPF_ERR_ARG;
    }
    lane_sweep_kernel<NA, REM, MODE, YHI><<<grid, L.warps * 32, smem, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("lane_sweep_kernel<KP=%d,mode=%d> launch (grid %d, block %d, smem %zu) -> %s", Cfg::KP, MODE, grid,
                  L.warps * 32, smem, cudaGetErrorString(e));
        return SCHPF_ERR_CUDA;
    }
    return SCHPF_OK;
}

template <int NA, int REM, int MODE>
int launch_lanes(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    return L.packed ? launch_lanes_y<NA, REM, MODE, ENC_PACKED>(L, args, stream)
           : L.yhi  ? launch_lanes_y<NA, REM, MODE, ENC_YHI>(L, args, stream)
                    : launch_lanes_y<NA, REM, MODE, ENC_WIDE>(L, args, stream);
}

template <int MODE>
int dispatch_lanes(int KP, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    switch (KP) {
        case 16: return launch_lanes<1, 0, MODE>(L, args, stream);
        case 20: return launch_lanes<1, 1, MODE>(L, args, stream);
        case 32: return launch_lanes<2, 0, MODE>(L, args, stream);
        default:
            set_error("lane sweep is not instantiated for KP=%d", KP);
            return SCHPF_ERR_ARG;
    }
}

}  // namespace

// K classes served by the one-lane-per-owner kernel: KP = 16 (K <= 16), 20 (17..20), 32 (29..32)
bool lanes_supported(int K)
{
    const int kp = lanes_kp_of(K);
    return kp == 16 || kp == 20 || kp == 32;
}

int lanes_default_warps(int K)
{
    const int kp = lanes_kp_of(K);
    return (kp <= 16 ? LANES_W16 : kp <= 20 ? LANES_W20 : LANES_W32) / LANES_CTAS;
}

// rows of the other axis per panel: the whole shared memory of an SM (one CTA per SM)
int lanes_max_panel_rows(int K)
{
    const int kp = lanes_kp_of(K);
    const size_t budget = (size_t)(228 * 1024) / LANES_CTAS - 1024 - 256;
    int rows = (int)(budget / ((size_t)kp * 8));
    rows &= ~3;
    if (rows > 4096) rows = 4096;   // 12-bit local index in the sort key
    return rows;
}

int launch_lane_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    const int KP = lanes_kp_of(K);
    if (mode == SWEEP_SHAPE) return dispatch_lanes<SWEEP_SHAPE>(KP, L, args, stream);
    return dispatch_lanes<SWEEP_LLH>(KP, L, args, stream);
}

}  // namespace schpf
