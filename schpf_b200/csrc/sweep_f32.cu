// fp32 sweep for `dtype=np.float32` models (reference: scHPF(dtype=...) scHPF_.py:225-246; with
// float32 arrays its numba kernels run compute_Xphi_data / compute_loading_shape_update /
// compute_pois_llh entirely in float32, hpf_numba.py:25-51,55-114,129-156).
//
// Same two-pass factored-softmax sweep and the same one-lane-per-owner mapping as sweep_lanes.cu,
// with the streamed tables held in shared memory as fp32: a K-row is one (K <= 32) or two
// (K <= 64) 128-byte planes, so every row starts at bank group 0 and lane i of a quarter warp
// reads the 16-byte unit (t ^ i) at time t -- conflict-free for ANY set of rows, no bank schedule.
// Per nonzero that is 128 (256) bytes of shared-memory traffic against 160 (256 / 512) in fp64,
// and the arithmetic runs on the fp32 pipe.
//
// Precision contract: table entries, the dot product s_i, the quotient y_i / s_i and
// the per-panel accumulators are fp32 -- what the reference computes in float32 -- while
// everything that crosses panels is fp64: the per-panel sums are converted and added to the fp64
// global accumulators (one TMA bulk reduction per owner row), and the state, the rate / digamma
// updates and the log-likelihood sum stay fp64 (the reference's own float32 run is mixed too:
// beta shape and eta rate come back float64).  The factored tables the finalisation
// multiplies by are the SAME fp32-rounded values the sweep read, so phi sums to one.
#include <type_traits>

#include "common.cuh"

namespace schpf {

namespace {

#ifndef F32_W1
#define F32_W1 12         // warps per CTA, one plane (K <= 32)
#endif
#ifndef F32_W2
#define F32_W2 8          // two planes (K <= 64)
#endif
#ifndef F32_NS1
#define F32_NS1 2         // steps processed as one straight-line block
#endif
#ifndef F32_NS2
#define F32_NS2 1
#endif

constexpr float TINY_NORMALIZER_F32 = 1e-30f;   // below this the nonzero is redone in log space (fp64)

template <int NP>
struct F32Cfg {
    static constexpr int KF = 32 * NP;
    static constexpr int WARPS = NP == 1 ? F32_W1 : F32_W2;
    static constexpr int NS = NP == 1 ? F32_NS1 : F32_NS2;
};

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void bulk_red_add_f64(double *dst_gmem, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(src_smem), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// one stream element (two steps of a lane) as {row, pad<<31 | count, row, pad<<31 | count}; PACKED:
// 4-byte entries pad<<31 | count<<12 | row (common.cuh pack_entry)
template <bool PACKED>
__device__ __forceinline__ int4 ld_elem(const char *p)
{
    if (PACKED) {
        const int2 w = ld_stream_int2(reinterpret_cast<const int2 *>(p));
        constexpr int RM = (1 << PACKED_ROW_BITS) - 1, CM = (1 << PACKED_COUNT_BITS) - 1;
        return make_int4(w.x & RM, (w.x & (int)0x80000000) | ((w.x >> PACKED_ROW_BITS) & CM), w.y & RM,
                         (w.y & (int)0x80000000) | ((w.y >> PACKED_ROW_BITS) & CM));
    }
    return ld_stream_int4(reinterpret_cast<const int4 *>(p));
}

template <int NP, int MODE, bool PACKED>
__global__ void __launch_bounds__(F32Cfg<NP>::WARPS * 32, 1)
lane_sweep_f32_kernel(const SweepArgs A)
{
    using Cfg = F32Cfg<NP>;
    constexpr int KF = Cfg::KF, NS = Cfg::NS;
    constexpr int ES = PACKED ? 8 : 16;           // bytes of one stream element
    constexpr int ROWB = KF * 4;                  // bytes per table row (128 * NP)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float *own_tab = reinterpret_cast<const float *>(A.own_tab);
    const float *oth_tab = reinterpret_cast<const float *>(A.oth_tab);
    const uint32_t panel_bytes = (uint32_t)A.panel_rows * ROWB;
    // the panel region also stages the fp64 accumulators of the epilogue (K doubles per thread)
    const uint32_t stage_bytes = (uint32_t)(A.warps * 32) * (uint32_t)A.K * 8u;
    const uint32_t region = (max(panel_bytes, stage_bytes) + 127u) & ~127u;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + region);
    double *red = reinterpret_cast<double *>(smem_raw + region + 16);   // LLH: per-warp sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i8 = lane & 7;
    const int b = blockIdx.x / A.nranges, r = blockIdx.x - b * A.nranges;
    const int p0 = r * A.panels_per_range;
    const int p1 = min(p0 + A.panels_per_range, A.npanel);
    const int wg = b * A.warps + warp;
    // lane -> slot of its warp: the inverse of layout.cu's stream_pos for 32 owners per warp
    const int slot = (((lane >> 3) * 2 + (lane & 1)) << 2) + ((lane & 7) >> 1);
    const int own = A.own_id[(int64_t)r * A.own_range_stride + (int64_t)wg * 32 + slot];

    // owner row in the per-lane rotated order: slot t of plane n holds unit (t ^ i8)
    float a[KF], acc[KF];
#pragma unroll
    for (int n = 0; n < NP; ++n)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (own >= 0) v = *reinterpret_cast<const float4 *>(own_tab + (int64_t)own * KF + 32 * n + 4 * (t ^ i8));
            a[32 * n + 4 * t] = v.x;
            a[32 * n + 4 * t + 1] = v.y;
            a[32 * n + 4 * t + 2] = v.z;
            a[32 * n + 4 * t + 3] = v.w;
        }
#pragma unroll
    for (int k = 0; k < KF; ++k) acc[k] = 0.f;
    double llh = 0.0;

    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
    const uint32_t panel_s = smem_u32(smem_raw) + ((uint32_t)i8 << 4);          // + this lane's unit offset
    const int64_t *sp = A.seg_ptr + (int64_t)wg * (A.npanel + 1);

    // NSTEP steps of this lane as one straight-line block.  `oth0` = global row of the panel's row 0.
    auto process = [&](auto ns_tag, const int *ex, const int *ey, int oth0) {
        constexpr int NSTEP = decltype(ns_tag)::value;
        float bv[NSTEP][KF], s[NSTEP];
#pragma unroll
        for (int e = 0; e < NSTEP; ++e) {
            const uint32_t addr = panel_s + (uint32_t)ex[e] * ROWB;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t at = addr ^ ((uint32_t)t << 4);
#pragma unroll
                for (int n = 0; n < NP; ++n) {
                    const float4 v = lds_f32x4(at + n * 128);
                    bv[e][32 * n + 4 * t] = v.x;
                    bv[e][32 * n + 4 * t + 1] = v.y;
                    bv[e][32 * n + 4 * t + 2] = v.z;
                    bv[e][32 * n + 4 * t + 3] = v.w;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < NSTEP; ++e) {
            float s0 = a[0] * bv[e][0], s1 = a[1] * bv[e][1], s2 = a[2] * bv[e][2], s3 = a[3] * bv[e][3];
#pragma unroll
            for (int k = 4; k < KF; k += 4) {
                s0 = fmaf(a[k], bv[e][k], s0);
                s1 = fmaf(a[k + 1], bv[e][k + 1], s1);
                s2 = fmaf(a[k + 2], bv[e][k + 2], s2);
                s3 = fmaf(a[k + 3], bv[e][k + 3], s3);
            }
            s[e] = (s0 + s1) + (s2 + s3);
        }
        if (MODE == SWEEP_SHAPE) {
            bool slow = false;
#pragma unroll
            for (int e = 0; e < NSTEP; ++e) {
                const float y = __int2float_rn(ey[e] & 0x7fffffff);      // pads carry count 0
                const bool ok = s[e] > TINY_NORMALIZER_F32;
                // y / s for a normal positive s without the division subroutine: hardware seed
                // (MUFU.RCP, ~1 ulp), one Newton step, the product, one residual correction (< 1 ulp)
                float rc;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(s[e]));
                rc = fmaf(fmaf(-s[e], rc, 1.f), rc, rc);
                const float q = y * rc;
                const float wr = ok ? fmaf(fmaf(-q, s[e], y), rc, q) : 0.f;
#pragma unroll
                for (int k = 0; k < KF; ++k) acc[k] = fmaf(wr, bv[e][k], acc[k]);
                slow |= !ok;
            }
            if (slow) {
                // cold: a nonzero whose factored normaliser underflowed fp32 is queued for the log-space
                // redo after the sweep (slow_fixup_kernel, fp64); pads and empty slots fall out here
#pragma unroll
                for (int e = 0; e < NSTEP; ++e)
                    if (!(s[e] > TINY_NORMALIZER_F32) && (ey[e] & 0x7fffffff) != 0 && own >= 0)
                        slow_enqueue(A, own, oth0 + ex[e], (double)(ey[e] & 0x7fffffff));
            }
        } else {
            // hpf_numba.py:49-50 without the lgamma term (a constant of the data); the terms are
            // fp32 like the reference's float32 run, their sum is fp64
#pragma unroll
            for (int e = 0; e < NSTEP; ++e) {
                const float v = fmaf(__int2float_rn(ey[e] & 0x7fffffff), logf(s[e]), -s[e]);
                if (ey[e] >= 0) llh += (double)v;          // bit 31 of the count field = pad
            }
        }
    };

    for (int p = p0; p < p1; ++p) {
        if (p > p0) __syncthreads();     // every warp is done with the previous panel
        if (tid == 0) {
            mbar_expect_tx(mbar, panel_bytes);
            bulk_g2s(smem_raw, oth_tab + (int64_t)p * A.panel_rows * KF, panel_bytes, mbar);
        }
        // this warp's segment of the stream: n elements (two steps each) of 32 lanes
        const int64_t i0 = sp[p];
        const int n = (int)(sp[p + 1] - i0);
        const char *ptr = reinterpret_cast<const char *>(A.entries) + (i0 * 32 + lane) * ES;
        int4 cur = make_int4(0, 0, 0, 0), nxt = cur;
        if (n > 0) cur = ld_elem<PACKED>(ptr);
        if (n > 1) nxt = ld_elem<PACKED>(ptr + 32 * ES);
        const int oth0 = p * A.panel_rows;
        mbar_wait(mbar, parity);
        parity ^= 1u;

        int j = 0;
        for (; j + 1 < n; j += 2, ptr += 64 * ES) {
            prefetch_l2(ptr + 8 * 64 * ES);
            prefetch_l2(ptr + (8 * 64 + 32) * ES);
            const int4 c0 = cur, c1 = nxt;
            if (j + 2 < n) cur = ld_elem<PACKED>(ptr + 64 * ES);
            if (j + 3 < n) nxt = ld_elem<PACKED>(ptr + 96 * ES);
            if constexpr (NS == 2) {
                {
                    const int ex[2] = {c0.x, c0.z}, ey[2] = {c0.y, c0.w};
                    process(std::integral_constant<int, 2>{}, ex, ey, oth0);
                }
                {
                    const int ex[2] = {c1.x, c1.z}, ey[2] = {c1.y, c1.w};
                    process(std::integral_constant<int, 2>{}, ex, ey, oth0);
                }
            } else {
                const int xs[4] = {c0.x, c0.z, c1.x, c1.z}, ys[4] = {c0.y, c0.w, c1.y, c1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ex[1] = {xs[q]}, ey[1] = {ys[q]};
                    process(std::integral_constant<int, 1>{}, ex, ey, oth0);
                }
            }
        }
        if (j < n) {
            if constexpr (NS == 2) {
                const int ex[2] = {cur.x, cur.z}, ey[2] = {cur.y, cur.w};
                process(std::integral_constant<int, 2>{}, ex, ey, oth0);
            } else {
                const int xs[2] = {cur.x, cur.z}, ys[2] = {cur.y, cur.w};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ex[1] = {xs[q]}, ey[1] = {ys[q]};
                    process(std::integral_constant<int, 1>{}, ex, ey, oth0);
                }
            }
        }
    }

    if (MODE == SWEEP_SHAPE) {
        const int K = A.K;
        __syncthreads();             // all warps are done reading the panel
        if (!(K & 1)) {
            // per-panel fp32 sums -> fp64, staged in natural k order where the panel was, then one
            // TMA bulk reduction (add.f64) per owner row into the global fp64 accumulator
            const uint32_t stage = smem_u32(smem_raw) + (uint32_t)tid * ((uint32_t)K * 8u);
#pragma unroll
            for (int n = 0; n < NP; ++n)
#pragma unroll
                for (int t = 0; t < 8; ++t)
#pragma unroll
                    for (int d = 0; d < 4; d += 2) {
                        const int k = 32 * n + 4 * (t ^ i8) + d;
                        if (k < K)
                            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(stage + (uint32_t)k * 8u),
                                         "d"((double)acc[32 * n + 4 * t + d]), "d"((double)acc[32 * n + 4 * t + d + 1])
                                         : "memory");
                    }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (own >= 0) bulk_red_add_f64(A.acc + (int64_t)own * K, stage, (uint32_t)K * 8u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else if (own >= 0) {
#pragma unroll
            for (int n = 0; n < NP; ++n)
#pragma unroll
                for (int t = 0; t < 8; ++t)
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const int k = 32 * n + 4 * (t ^ i8) + d;
                        if (k < K) atomicAdd(A.acc + (int64_t)own * K + k, (double)acc[32 * n + 4 * t + d]);
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

template <int NP, int MODE, bool PACKED>
int launch_f32_enc(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    using Cfg = F32Cfg<NP>;
    const size_t smem = f32_sweep_smem_bytes(args.K, L.panel_rows, L.warps);
    static bool configured = false;   // per instantiation
    static size_t configured_smem = 0;
    if (!configured || smem > configured_smem) {
        CUDA_TRY(cudaFuncSetAttribute(lane_sweep_f32_kernel<NP, MODE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(lane_sweep_f32_kernel<NP, MODE, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured = true;
        configured_smem = smem;
    }
    const int grid = L.nblocks * L.nranges;
    if (grid <= 0) return SCHPF_OK;
    if (L.warps > Cfg::WARPS || L.warps < 1) {
        set_error("fp32 sweep: warps_per_cta must be in [1, %d] for K=%d", Cfg::WARPS, args.K);
        return SCHPF_ERR_ARG;
    }
    lane_sweep_f32_kernel<NP, MODE, PACKED><<<grid, L.warps * 32, smem, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("lane_sweep_f32_kernel<planes=%d,mode=%d> launch (grid %d, block %d, smem %zu) -> %s", NP, MODE, grid,
                  L.warps * 32, smem, cudaGetErrorString(e));
        return SCHPF_ERR_CUDA;
    }
    return SCHPF_OK;
}

template <int NP, int MODE>
int launch_f32(const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    return L.packed ? launch_f32_enc<NP, MODE, true>(L, args, stream) : launch_f32_enc<NP, MODE, false>(L, args, stream);
}

}  // namespace

int f32_row_floats(int K) { return K <= 32 ? 32 : 64; }

int f32_default_warps(int K) { return K <= 32 ? F32_W1 : F32_W2; }

size_t f32_sweep_smem_bytes(int K, int panel_rows, int warps)
{
    const size_t panel = (size_t)panel_rows * f32_row_floats(K) * 4, stage = (size_t)warps * 32 * K * 8;
    return (((panel > stage ? panel : stage) + 127) & ~size_t(127)) + 16 + 16 * 8;
}

// rows of the other axis per panel: the whole shared memory of an SM (one CTA per SM)
int f32_max_panel_rows(int K)
{
    const size_t budget = (size_t)(228 * 1024) - 1024 - 256 - 128;
    int rows = (int)(budget / ((size_t)f32_row_floats(K) * 4));
    rows &= ~3;
    if (rows > 4096) rows = 4096;   // 12-bit local index in the sort key
    return rows;
}

int launch_f32_sweep(int mode, int K, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    if (K <= 32) return mode == SWEEP_SHAPE ? launch_f32<1, SWEEP_SHAPE>(L, args, stream) : launch_f32<1, SWEEP_LLH>(L, args, stream);
    return mode == SWEEP_SHAPE ? launch_f32<2, SWEEP_SHAPE>(L, args, stream) : launch_f32<2, SWEEP_LLH>(L, args, stream);
}

}  // namespace schpf
