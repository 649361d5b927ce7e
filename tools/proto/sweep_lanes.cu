// PROTOTYPE, not part of the library (schpf_b200/build.py does not compile it) and never run:
// the sweep with ONE LANE per owner (32 owners per warp) for small K, DESIGN.md §7-3.  It exists
// for the compile-time evidence quoted there:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v \
//        -I schpf_b200/csrc -c tools/proto/sweep_lanes.cu -o /tmp/sweep_lanes.o
//   KP=8: 126 registers at 16 warps, no spills; KP=12: 128 (8-byte spill) at 16 warps, 166 at 12;
//   executed loop instructions per entry (SASS, rare-path branches not taken): ~1.4 against ~2.2 for
//   the lane-pair kernel at KP=8, and no DMMA / SHFL.
// Missing on purpose: the log-space fallback (only counted), the layout with 8 bank classes per
// quarter warp (tools/sim_schedule.py:greedy_place is its reference), the table re-stride (ST/2 odd).
#include "common.cuh"
namespace schpf {
namespace {
template <int KP, int MODE, int WARPS, int NS>
__global__ void __launch_bounds__(WARPS * 32, 1) sweep_lane_kernel(const SweepArgs A, int ST1)
{
    constexpr int U = KP / 2;      // 16-byte units per row
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *panel = reinterpret_cast<double *>(smem_raw);
    const uint32_t panel_bytes = (uint32_t)A.panel_rows * ST1 * 8u;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + panel_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / A.nranges, r = blockIdx.x - b * A.nranges;
    const int p0 = r * A.panels_per_range, p1 = min(p0 + A.panels_per_range, A.npanel);
    const int wg = b * A.warps + warp;
    const int own = A.own_id[(int64_t)wg * 32 + lane];
    double a[KP], acc[KP];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        double2 v = make_double2(0.0, 0.0);
        if (own >= 0) v = reinterpret_cast<const double2 *>(A.own_tab + (int64_t)own * ST1)[j];
        a[2 * j] = v.x; a[2 * j + 1] = v.y; acc[2 * j] = 0.0; acc[2 * j + 1] = 0.0;
    }
    double llh = 0.0;
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t panel_s = smem_u32(panel);
    const int64_t *sp = A.seg_ptr + (int64_t)wg * (A.npanel + 1);
    const int4 *ent = reinterpret_cast<const int4 *>(A.entries);
    for (int p = p0; p < p1; ++p) {
        if (p > p0) __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(mbar, panel_bytes);
            bulk_g2s(panel, A.oth_tab + (int64_t)p * A.panel_rows * ST1, panel_bytes, mbar);
        }
        const int64_t i0 = sp[p], i1 = sp[p + 1];
        int4 cur = make_int4(0, 0, 0, 0), nxt = cur;
        if (i0 < i1) cur = ld_stream_int4(ent + i0 * 32 + lane);
        if (i0 + 1 < i1) nxt = ld_stream_int4(ent + (i0 + 1) * 32 + lane);
        mbar_wait(mbar, parity);
        parity ^= 1u;
        auto process = [&](const int *ex, const int *ey, const bool *epad) {
            double bv[NS][KP], s[NS];
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                const uint32_t addr = panel_s + (uint32_t)ex[e] * (uint32_t)(ST1 * 8);
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const double2 v = lds_f64x2(addr + j * 16);
                    bv[e][2 * j] = v.x; bv[e][2 * j + 1] = v.y;
                }
            }
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                double c0 = a[0] * bv[e][0], c1 = a[1] * bv[e][1], c2 = a[2] * bv[e][2], c3 = a[3] * bv[e][3];
#pragma unroll
                for (int k = 4; k < KP; k += 4) {
                    c0 = fma(a[k], bv[e][k], c0); c1 = fma(a[k + 1], bv[e][k + 1], c1);
                    c2 = fma(a[k + 2], bv[e][k + 2], c2); c3 = fma(a[k + 3], bv[e][k + 3], c3);
                }
                s[e] = (c0 + c1) + (c2 + c3);
            }
            if (MODE == SWEEP_SHAPE) {
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    const bool ok = s[e] > TINY_NORMALIZER;
                    const double w = ok ? div_pos_folded((double)ey[e], s[e]) : 0.0;
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] = fma(w, bv[e][k], acc[k]);
                    if (!ok && ey[e] != 0 && own >= 0) atomicAdd(A.slow_hits, 1ULL);   // (fallback omitted in the prototype)
                }
            } else {
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    const double v = fma((double)ey[e], log(s[e]), -s[e]);
                    if (!epad[e]) llh += v;
                }
            }
        };
        int64_t i = i0;
        for (; i + 1 < i1; i += 2) {
            {
                const int ex[2] = {cur.x & 0x7fffffff, cur.z & 0x7fffffff}, ey[2] = {cur.y, cur.w};
                const bool epad[2] = {cur.x < 0, cur.z < 0};
                if (i + 2 < i1) cur = ld_stream_int4(ent + (i + 2) * 32 + lane);
                process(ex, ey, epad);
            }
            {
                const int ex[2] = {nxt.x & 0x7fffffff, nxt.z & 0x7fffffff}, ey[2] = {nxt.y, nxt.w};
                const bool epad[2] = {nxt.x < 0, nxt.z < 0};
                if (i + 3 < i1) nxt = ld_stream_int4(ent + (i + 3) * 32 + lane);
                process(ex, ey, epad);
            }
        }
        if (i < i1) {
            const int ex[2] = {cur.x & 0x7fffffff, cur.z & 0x7fffffff}, ey[2] = {cur.y, cur.w};
            const bool epad[2] = {cur.x < 0, cur.z < 0};
            process(ex, ey, epad);
        }
    }
    if (MODE == SWEEP_SHAPE) {
        if (own >= 0)
#pragma unroll
            for (int k = 0; k < KP; ++k) if (k < A.K) atomicAdd(A.acc + (int64_t)own * A.K + k, acc[k]);
    } else {
        llh = warp_sum(llh);
        if (lane == 0) atomicAdd(A.partial + blockIdx.x, llh);
    }
}
}  // namespace
void proto_launch(const SweepArgs &A, int ST1, cudaStream_t s)
{
    sweep_lane_kernel<8, 0, 16, 2><<<1, 512, 1024, s>>>(A, ST1);
    sweep_lane_kernel<12, 0, 16, 2><<<1, 512, 1024, s>>>(A, ST1);
    sweep_lane_kernel<12, 0, 12, 2><<<1, 384, 1024, s>>>(A, ST1);
    sweep_lane_kernel<16, 0, 12, 2><<<1, 384, 1024, s>>>(A, ST1);
    sweep_lane_kernel<20, 0, 12, 1><<<1, 384, 1024, s>>>(A, ST1);
    sweep_lane_kernel<20, 0, 8, 2><<<1, 256, 1024, s>>>(A, ST1);
}
}  // namespace schpf
