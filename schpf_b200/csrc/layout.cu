// Builds the panelled sliced-ELL layout one sweep direction reads.
//
// The reference hands its kernels plain COO arrays (X.data, X.row, X.col,
// scHPF_.py:661-664) and re-reads them in whatever order scipy left them.
// Here the triples are re-ordered once per fit, on the device, into the order
// the sweep kernel consumes:
//
//   owners (cells for the theta sweep, genes for the beta sweep) are ranked by
//   their nonzero count (descending) and dealt into slots: 16 consecutive
//   slots = one warp (one owner per lane pair), `warps` warps = one CTA block;
//   the other axis is cut into panels of `panel_rows` rows;
//   for every (warp, panel) the 16 owners' nonzeros inside that panel are
//   stored step-interleaved (step i of all 16 pairs is contiguous), two steps
//   per element: int4 {other_local | pad<<31, y, other_local | pad<<31, y}, or -- when every
//   count of the matrix is below 2^19 -- int2 of packed words pad<<31 | y<<12 | other_local;
//   WHICH nonzero a pair handles at which step is a conflict-free schedule: a
//   shared-memory row of the panel falls into bank group (other_local mod 4),
//   and the four lane pairs of a quarter warp are served by one wavefront only
//   if they read four different groups.  For each (quarter warp, panel) the
//   4 owners x 4 groups count matrix is a bipartite multigraph; it is edge-
//   coloured with Delta = max(row sums, column sums) colours (Koenig), one
//   colour = one step = a partial permutation owners -> groups.  Pairs with no
//   edge of a colour get a pad entry (count 0) pointing at a row of the group
//   the permutation leaves them, so pads never conflict either.  The
//   (warp, panel) block has max over its 4 quarter warps of Delta steps,
//   rounded up to even.
//
// Integer work only; results are bit-exact and independent of the input order
// of the triples (ties are broken by the other-axis index, duplicates by the
// radix sort's stability).
#include <cub/cub.cuh>

#include <mutex>

#include "common.cuh"

namespace schpf {

namespace {

constexpr int KEY_LOCAL_BITS = 12;   // panel_rows <= 4096
constexpr int KEY_ROT_BITS = 2;

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// histogram of owner ids.  COO rows usually arrive sorted, i.e. whole warps hit one counter:
// equal keys are combined inside the warp first (one atomic per distinct key and warp).
__global__ void count_owners_kernel(int64_t nnz, const int32_t *__restrict__ own, int32_t *__restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int32_t key = own[i];
    const unsigned peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(cnt + key, __popc(peers));
}

// per-range variant: cnt[range * n_own + owner], range = panel of the other index / panels_per_range
__global__ void count_owners_ranged_kernel(int64_t nnz, const int32_t *__restrict__ own,
                                           const int32_t *__restrict__ oth, int panel_rows, int panels_per_range,
                                           int64_t n_own, int32_t *__restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t key = (int64_t)((oth[i] / panel_rows) / panels_per_range) * n_own + own[i];
    const unsigned peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(cnt + key, __popc(peers));
}

// sort keys of the per-range ranking: range ascending, count descending (stable: ties keep owner order)
__global__ void ranged_keys_kernel(int64_t n, int64_t n_own, const int32_t *__restrict__ cnt,
                                   uint64_t *__restrict__ keys, int32_t *__restrict__ ids)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t range = j / n_own;
    keys[j] = ((uint64_t)range << 32) | (uint64_t)(0xffffffffu - (uint32_t)cnt[j]);
    ids[j] = (int32_t)(j - range * n_own);
}

// slot_of[range][owner], own_id[range][slot] from the per-range sorted order
__global__ void assign_slots_ranged_kernel(int64_t n_own, int64_t n_slots, int nranges,
                                           const int32_t *__restrict__ sorted_ids, int32_t *__restrict__ slot_of,
                                           int32_t *__restrict__ own_id)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_slots * nranges) return;
    const int64_t range = j / n_slots, s = j - range * n_slots;
    if (s < n_own) {
        const int32_t o = sorted_ids[range * n_own + s];
        own_id[j] = o;
        slot_of[range * n_own + o] = (int32_t)s;
    } else {
        own_id[j] = -1;
    }
}

__global__ void iota_kernel(int64_t n, int32_t *p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// slot_of[owner] and own_id[slot] from the count-sorted owner order
__global__ void assign_slots_kernel(int64_t n_own, int64_t n_slots, const int32_t *__restrict__ sorted_ids,
                                    int32_t *__restrict__ slot_of, int32_t *__restrict__ own_id)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    if (s < n_own) {
        const int32_t o = sorted_ids[s];
        own_id[s] = o;
        slot_of[o] = (int32_t)s;
    } else {
        own_id[s] = -1;
    }
}

__global__ void make_keys_kernel(int64_t nnz, const int32_t *__restrict__ own, const int32_t *__restrict__ oth,
                                 const int32_t *__restrict__ val, const int32_t *__restrict__ slot_of,
                                 int panel_rows, int npanel, int64_t ranged_n_own, int panels_per_range, bool free_mode,
                                 bool rotate, uint64_t *__restrict__ keys, uint64_t *__restrict__ vals, int64_t *__restrict__ seg_cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int32_t t = oth[i];
    const int32_t p = t / panel_rows;
    const int32_t tl = t - p * panel_rows;
    // ranged_n_own > 0: owners are ranked inside every panel range, slot_of is [nranges][n_own]
    const int32_t slot = slot_of[(int64_t)(p / panels_per_range) * ranged_n_own + own[i]];
    const int64_t seg = (int64_t)slot * npanel + p;
    // scheduled streams: the bank class of the row.  Free streams keep the owner's nonzeros of a panel
    // contiguous; `rotate` orders them by bank class, starting at a class that differs between the
    // four owners of a scheduling group, so that plane-B reads of the group collide less often.
    const uint32_t cls = !free_mode ? ((uint32_t)tl & 3u) : rotate ? (((uint32_t)tl - (uint32_t)slot) & 3u) : 0u;
    keys[i] = ((uint64_t)seg << (KEY_LOCAL_BITS + KEY_ROT_BITS)) | ((uint64_t)cls << KEY_LOCAL_BITS) |
              (uint64_t)tl;
    vals[i] = ((uint64_t)(uint32_t)val[i] << 32) | (uint64_t)(uint32_t)tl;
    const int64_t list = seg * 4 + cls;
    const unsigned peers = __match_any_sync(__activemask(), list);
    if ((threadIdx.x & 31) == __ffs(peers) - 1)
        atomicAdd(reinterpret_cast<unsigned long long *>(seg_cnt + list), (unsigned long long)__popc(peers));
}

// ---- conflict-free schedule of one (quarter warp, panel) ----------------------
// n[o][c]: nonzeros of owner o (0..3, one per lane pair) whose panel-local row is in
// bank group c.  Delta = max(row sums, column sums) steps suffice and are necessary.
__device__ __forceinline__ int qw_delta(const int64_t *__restrict__ cnt4 /* 4 owners x [npanel*4] */,
                                        int64_t base0, int64_t owner_stride, int n[4][4], bool free_mode = false)
{
    int rs[4] = {0, 0, 0, 0}, cs[4] = {0, 0, 0, 0};
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            n[o][c] = (int)cnt4[base0 + o * owner_stride + c];
            rs[o] += n[o][c];
            cs[c] += n[o][c];
        }
    int d = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) d = max(d, free_mode ? rs[i] : max(rs[i], cs[i]));   // free: no bank classes
    return d;
}

// step pairs per (warp, panel) = ceil(max over the warp's 4 quarter warps of Delta / 2)
__global__ void warp_steps_kernel(int64_t n_warps, int npanel, int opw, bool free_mode,
                                  const int64_t *__restrict__ cnt4, int64_t *__restrict__ pairs)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_warps * (npanel + 1)) return;
    const int64_t wg = idx / (npanel + 1);
    const int p = (int)(idx - wg * (npanel + 1));
    if (p == npanel) {
        pairs[idx] = 0;
        return;
    }
    int m = 0;
    for (int qw = 0; qw < opw / 4; ++qw) {       // groups of 4 owners scheduled together
        int n[4][4];
        const int64_t slot0 = wg * opw + qw * 4;
        m = max(m, qw_delta(cnt4, (slot0 * npanel + p) * 4, (int64_t)npanel * 4, n, free_mode));
    }
    pairs[idx] = (m + 1) >> 1;
}

// position of slot s (0..opw-1 within its warp) in a step row of the stream.  16 owners per warp: the
// lane pair s.  32 owners per warp: group g = s / 4 of 4 owners scheduled together is the even
// (g even) or odd lanes of quarter warp g / 2, member m = s % 4 -> lane (g/2)*8 + 2m + (g&1).
__host__ __device__ __forceinline__ int stream_pos(int s, int opw)
{
    if (opw == GROUPS_PER_WARP) return s;
    const int g = s >> 2, m = s & 3;
    return (g >> 1) * 8 + 2 * m + (g & 1);
}
// and the member index (0..3 inside its scheduling group) of the owner at stream position `pos`
__host__ __device__ __forceinline__ int pos_member(int pos, int opw)
{
    return opw == GROUPS_PER_WARP ? (pos & 3) : ((pos & 7) >> 1);
}

template <bool PACKED>
__global__ void fill_pad_kernel(int64_t n_elems, int opw, bool free_mode, void *entries)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // pads read a real row of the bank group their lane (pair) would own in an all-pad step
    // (rows 0..3 are in groups 0..3), so they never conflict with each other
    const int o = free_mode ? 0 : pos_member((int)(i % opw), opw);
    if (i >= n_elems) return;
    if (PACKED) {
        const int w = (int)pack_entry((uint32_t)o, 0u, true);
        reinterpret_cast<int2 *>(entries)[i] = make_int2(w, w);
    } else if (opw == 32) {
        // one-lane stream: the pad flag is bit 31 of the COUNT field, the row is plain (sweep_lanes.cu)
        reinterpret_cast<int4 *>(entries)[i] = make_int4(o, (int)0x80000000, o, (int)0x80000000);
    } else {
        reinterpret_cast<int4 *>(entries)[i] = make_int4((int)0x80000000 | o, 0, (int)0x80000000 | o, 0);
    }
}

// the 24 permutations of {0,1,2,3}, 2 bits per element
__constant__ unsigned char PERM4[24] = {
    0xE4, 0xB4, 0xD8, 0x78, 0x9C, 0x6C, 0xE1, 0xB1, 0xC9, 0x39, 0x8D, 0x2D,
    0xD2, 0x72, 0xC6, 0x36, 0x4E, 0x1E, 0x93, 0x63, 0x87, 0x27, 0x4B, 0x1B};

// One thread per (quarter warp, panel): colour the 4x4 multigraph and copy each nonzero
// to its (step, pair) slot.  At a step with R steps left, every row / column whose
// remaining degree equals R must be matched (then the maximum degree drops to R-1); a
// matching doing so always exists in a bipartite multigraph, and with 4+4 vertices the
// 24 permutations are simply tried.
__device__ __forceinline__ int count_field(uint32_t count, bool yhi)
{
    return yhi ? __double2hiint((double)count) : (int)count;
}

template <bool PACKED>
__global__ void place_entries_kernel(int64_t n_qw, int npanel, int opw, bool free_mode, bool yhi,
                                     const int64_t *__restrict__ cnt4,
                                     const int64_t *__restrict__ first4, const uint64_t *__restrict__ vals,
                                     const int64_t *__restrict__ seg_ptr, void *__restrict__ entries_v,
                                     int *__restrict__ unplaced)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_qw * npanel) return;
    const int64_t qwg = idx / npanel;              // global quarter-warp index
    const int p = (int)(idx - qwg * npanel);
    const int64_t slot0 = qwg * 4;
    const int64_t wg = slot0 / opw;
    const int q0 = (int)(slot0 - wg * opw);
    const int64_t base0 = (slot0 * npanel + p) * 4, ostride = (int64_t)npanel * 4;
    int n[4][4], used[4][4];
    const int delta = qw_delta(cnt4, base0, ostride, n, free_mode);
    if (delta == 0) return;
    if (free_mode) {
        // no bank classes: owner o's nonzeros of this panel in ascending order, one per step
        const int64_t pair0 = seg_ptr[wg * (npanel + 1) + p];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int64_t src0 = first4[base0 + o * ostride];      // the owner's four class lists are contiguous
            const int pos = stream_pos(q0 + o, opw);
            const int cnt_o = n[o][0] + n[o][1] + n[o][2] + n[o][3];
            // Rotated class lists (K 17..20: list c of the owner in slot s holds the rows of bank class c + s):
            // taken ROUND-ROBIN, list (step mod 4) at every step while it lasts, the four owners of a
            // scheduling group read four different bank classes at the same step -- in phase at every step
            // instead of only while four class-major runs happen to have the same length.  Streams without
            // classes have everything in list 0 and come out in the sorted order as before.
            int l0 = 0, l1 = n[o][0], l2 = l1 + n[o][1], l3 = l2 + n[o][2];      // cursors of the four lists
            const int h0 = l1, h1 = l2, h2 = l3, h3 = cnt_o;                       // (registers: no indexed arrays)
            for (int step = 0; step < cnt_o; ++step) {
                const int64_t dst = ((pair0 + (step >> 1)) * opw + pos) * 2 + (step & 1);
                const unsigned avail = (l0 < h0 ? 1u : 0u) | (l1 < h1 ? 2u : 0u) | (l2 < h2 ? 4u : 0u) | (l3 < h3 ? 8u : 0u);
                const int want = step & 3;
                const int c = (want + __ffs(((avail | (avail << 4)) >> want) & 0xfu) - 1) & 3;   // first list with entries from `want` on
                const int at = c == 0 ? l0 : c == 1 ? l1 : c == 2 ? l2 : l3;
                l0 += c == 0;
                l1 += c == 1;
                l2 += c == 2;
                l3 += c == 3;
                const uint64_t v = vals[src0 + at];
                if (PACKED)
                    reinterpret_cast<uint32_t *>(entries_v)[dst] = pack_entry((uint32_t)v, (uint32_t)(v >> 32), false);
                else
                    reinterpret_cast<int2 *>(entries_v)[dst] =
                        make_int2((int)(uint32_t)v, count_field((uint32_t)(v >> 32), yhi));
            }
        }
        return;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) used[o][c] = 0;
    int rs[4], cs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rs[i] = n[i][0] + n[i][1] + n[i][2] + n[i][3];
        cs[i] = n[0][i] + n[1][i] + n[2][i] + n[3][i];
    }
    const int64_t pair0 = seg_ptr[wg * (npanel + 1) + p];
    for (int step = 0; step < delta; ++step) {
        const int R = delta - step;
        int best = -1, best_score = -1;
        for (int k = 0; k < 24; ++k) {
            const unsigned pm = PERM4[k];
            bool ok = true;
            int score = 0;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int c = (pm >> (2 * o)) & 3;
                const bool has = n[o][c] > 0;
                score += has;
                if (!has && (rs[o] == R || cs[c] == R)) ok = false;
            }
            if (ok && score > best_score) {
                best = k;
                best_score = score;
            }
        }
        if (best < 0) best = 0;   // cannot happen (Koenig); keeps the loop well defined
        const unsigned pm = PERM4[best];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int c = (pm >> (2 * o)) & 3;
            const int64_t dst = ((pair0 + (step >> 1)) * opw + stream_pos(q0 + o, opw)) * 2 + (step & 1);
            if (n[o][c] > 0) {
                const int64_t src = first4[base0 + o * ostride + c] + used[o][c];
                const uint64_t v = vals[src];
                if (PACKED)
                    reinterpret_cast<uint32_t *>(entries_v)[dst] = pack_entry((uint32_t)v, (uint32_t)(v >> 32), false);
                else
                    reinterpret_cast<int2 *>(entries_v)[dst] =
                        make_int2((int)(uint32_t)v, count_field((uint32_t)(v >> 32), yhi));
                ++used[o][c];
                --n[o][c];
                --rs[o];
                --cs[c];
            } else {
                // idle pair: a pad that reads row c, the bank group this permutation leaves it
                if (PACKED)
                    reinterpret_cast<uint32_t *>(entries_v)[dst] = pack_entry((uint32_t)c, 0u, true);
                else if (opw == 32)
                    reinterpret_cast<int2 *>(entries_v)[dst] = make_int2(c, (int)0x80000000);
                else
                    reinterpret_cast<int2 *>(entries_v)[dst] = make_int2((int)(0x80000000u | (unsigned)c), 0);
            }
        }
    }
    if (rs[0] | rs[1] | rs[2] | rs[3]) atomicAdd(unplaced, 1);   // schedule incomplete: never expected
}

// ---- scratch for the build's temporaries ----------------------------------------
// One block per device, kept between builds (grow-only; schpf_release_scratch frees it): the
// temporaries are ~30 B per nonzero, and growing the stream-ordered pool by gigabytes costs
// 50-500 ms per build (measured, tools/diag_layout.py) against ~50 ms of actual work.
struct ScratchSlot {
    void *p = nullptr;
    size_t cap = 0;
    bool busy = false;
};
constexpr int MAX_DEVICES = 64;
ScratchSlot g_scratch[MAX_DEVICES];
std::mutex g_scratch_mutex;

struct Scratch {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    int slot = -1;               // >= 0: the device's cached block; -1: a one-off pool allocation
    cudaStream_t stream = nullptr;

    cudaError_t acquire(size_t bytes, cudaStream_t s)
    {
        stream = s;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        {
            std::lock_guard<std::mutex> lock(g_scratch_mutex);
            if (dev < MAX_DEVICES && !g_scratch[dev].busy) {
                ScratchSlot &S = g_scratch[dev];
                if (S.cap < bytes) {
                    if (S.p) cudaFree(S.p);      // idle: its last user synchronised before releasing
                    S.p = nullptr;
                    S.cap = 0;
                    e = cudaMalloc(&S.p, bytes);
                    if (e != cudaSuccess) return e;
                    S.cap = bytes;
                }
                S.busy = true;
                slot = dev;
                base = static_cast<char *>(S.p);
                cap = S.cap;
                return cudaSuccess;
            }
        }
        // another build is using the block (trials on several host threads): one-off allocation
        void *p = nullptr;
        e = pool_malloc(&p, bytes, stream);
        if (e != cudaSuccess) return e;
        base = static_cast<char *>(p);
        cap = bytes;
        return cudaSuccess;
    }
    ~Scratch()
    {
        if (!base) return;
        cudaStreamSynchronize(stream);   // also on error paths: nothing may still be using the block
        if (slot >= 0) {
            std::lock_guard<std::mutex> lock(g_scratch_mutex);
            g_scratch[slot].busy = false;
        } else {
            pool_free(base, stream);
        }
    }
    static size_t rounded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
    template <typename T> T *take(size_t n)
    {
        T *p = reinterpret_cast<T *>(base + used);
        used += rounded(sizeof(T) * (n ? n : 1));
        return p;
    }
};

int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

}  // namespace

int release_layout_scratch(int device)
{
    if (device < 0 || device >= MAX_DEVICES) return SCHPF_OK;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    ScratchSlot &S = g_scratch[device];
    if (S.busy) {
        set_error("layout scratch of device %d is in use", device);
        return SCHPF_ERR_STATE;
    }
    if (S.p) {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaFree(S.p));
    }
    S.p = nullptr;
    S.cap = 0;
    return SCHPF_OK;
}

void SideLayout::release()
{
    if (own_id) pool_free(own_id, stream);
    if (seg_ptr) pool_free(seg_ptr, stream);
    if (entries) pool_free(entries, stream);
    own_id = nullptr;
    seg_ptr = nullptr;
    entries = nullptr;
    bytes = 0;
}

int build_side_layout(SideLayout &L, cudaStream_t stream, int64_t nnz, const int32_t *d_own,
                      const int32_t *d_oth, const int32_t *d_val, int64_t n_own, int64_t n_oth,
                      int panel_rows, int warps, int target_ctas, bool packed, int owners_per_warp, int flags)
{
    L.release();
    const int opw = owners_per_warp;
    const bool free_mode = flags & LAYOUT_FREE, ranked = flags & LAYOUT_RANK_PER_RANGE, yhi = flags & LAYOUT_YHI;
    if (opw != GROUPS_PER_WARP && opw != 32) {
        set_error("layout: owners per warp must be 16 or 32, got %d", opw);
        return SCHPF_ERR_ARG;
    }
    if (packed && (yhi || (opw == 32 && !free_mode))) {
        set_error("layout: packed entries exist for the lane-pair stream and for the schedule-free one-lane stream");
        return SCHPF_ERR_ARG;
    }
    L.packed = packed;
    L.opw = opw;
    L.free_mode = free_mode;
    L.ranked_per_range = ranked;
    L.yhi = yhi;
    if (panel_rows < 4 || panel_rows > (1 << KEY_LOCAL_BITS) || (panel_rows & 3)) {
        set_error("panel_rows must be a multiple of 4 in [4, %d], got %d", 1 << KEY_LOCAL_BITS, panel_rows);
        return SCHPF_ERR_ARG;
    }
    L.n_own = n_own;
    L.n_oth = n_oth;
    L.panel_rows = panel_rows;
    L.npanel = (int)((n_oth + panel_rows - 1) / panel_rows);
    if (L.npanel < 1) L.npanel = 1;
    L.warps = warps;
    const int64_t owners_per_block = (int64_t)warps * opw;
    L.nblocks = (int)((n_own + owners_per_block - 1) / owners_per_block);
    if (L.nblocks < 1) L.nblocks = 1;
    int nranges = (target_ctas + L.nblocks - 1) / L.nblocks;
    if (nranges < 1) nranges = 1;
    if (nranges > L.npanel || (flags & LAYOUT_SINGLE_PANEL_RANGES)) nranges = L.npanel;
    L.panels_per_range = (L.npanel + nranges - 1) / nranges;
    L.nranges = (L.npanel + L.panels_per_range - 1) / L.panels_per_range;

    const int64_t n_slots = (int64_t)L.nblocks * owners_per_block;
    const int64_t n_warps = (int64_t)L.nblocks * warps;
    const int64_t n_seg = n_slots * L.npanel * 4;   // (slot, panel, bank group) lists
    const int64_t n_ptr = n_warps * (L.npanel + 1);
    const int64_t n_rank = ranked ? (int64_t)L.nranges * n_own : n_own;      // ranking problems x owners
    const int64_t n_ownid = ranked ? (int64_t)L.nranges * n_slots : n_slots;
    L.n_slots = n_slots;
    if (n_rank > 0x7fffffffLL) {
        set_error("layout: %lld (range, owner) pairs exceed the ranking sort", (long long)n_rank);
        return SCHPF_ERR_ARG;
    }

    L.stream = stream;
    CUDA_TRY(pool_malloc(reinterpret_cast<void **>(&L.own_id), sizeof(int32_t) * n_ownid, stream));
    CUDA_TRY(pool_malloc(reinterpret_cast<void **>(&L.seg_ptr), sizeof(int64_t) * n_ptr, stream));

    // sizes of the sort / scan work areas first (no device work), then ONE scratch block
    const int key_bits = KEY_LOCAL_BITS + KEY_ROT_BITS + bits_for((uint64_t)(n_seg > 4 ? n_seg / 4 - 1 : 0));
    const int rank_bits = 32 + bits_for((uint64_t)(L.nranges > 1 ? L.nranges - 1 : 0));
    size_t tmp_bytes = 0, need = 0;
    {
        cub::DoubleBuffer<int32_t> k32(nullptr, nullptr), v32(nullptr, nullptr);
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, k32, v32, (int)n_own, 0, 32, stream));
        tmp_bytes = need;
        cub::DoubleBuffer<uint64_t> k64(nullptr, nullptr), v64(nullptr, nullptr);
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, k64, v64, nnz, 0, key_bits, stream));
        if (need > tmp_bytes) tmp_bytes = need;
        if (ranked) {
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, k64, v32, (int)n_rank, 0, rank_bits, stream));
            if (need > tmp_bytes) tmp_bytes = need;
        }
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, (int64_t *)nullptr, (int64_t *)nullptr, n_seg, stream));
        if (need > tmp_bytes) tmp_bytes = need;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, (int64_t *)nullptr, (int64_t *)nullptr, n_ptr, stream));
        if (need > tmp_bytes) tmp_bytes = need;
    }
    const size_t R = 256;   // Scratch::take rounds every piece up to this
    const size_t total = 5 * (sizeof(int32_t) * n_rank + R) + 2 * (sizeof(uint64_t) * (ranked ? n_rank : 0) + R) +
                         2 * (sizeof(int64_t) * n_seg + R) + (sizeof(int64_t) * n_ptr + R) +
                         4 * (sizeof(uint64_t) * nnz + R) + (tmp_bytes + R) + 2 * R;
    Scratch scratch;
    CUDA_TRY(scratch.acquire(total, stream));
    int *unplaced = scratch.take<int>(1);
    int32_t *cnt = scratch.take<int32_t>(n_rank), *ids = scratch.take<int32_t>(n_rank);
    int32_t *cnt_alt = scratch.take<int32_t>(n_rank), *ids_alt = scratch.take<int32_t>(n_rank);
    int32_t *slot_of = scratch.take<int32_t>(n_rank);
    uint64_t *rkeys = scratch.take<uint64_t>(ranked ? n_rank : 0), *rkeys_alt = scratch.take<uint64_t>(ranked ? n_rank : 0);
    int64_t *seg_cnt = scratch.take<int64_t>(n_seg), *seg_first = scratch.take<int64_t>(n_seg);
    int64_t *pairs = scratch.take<int64_t>(n_ptr);
    uint64_t *keys = scratch.take<uint64_t>(nnz), *vals = scratch.take<uint64_t>(nnz);
    uint64_t *keys_alt = scratch.take<uint64_t>(nnz), *vals_alt = scratch.take<uint64_t>(nnz);
    void *tmp = scratch.take<char>(tmp_bytes);
    if (scratch.used > scratch.cap) {
        set_error("layout: scratch accounting is off (%zu > %zu)", scratch.used, scratch.cap);
        return SCHPF_ERR_STATE;
    }
    CUDA_TRY(cudaMemsetAsync(unplaced, 0, sizeof(int), stream));

    trace_mark(stream, "  alloc");
    // 1. owners ranked by nonzero count, descending (stable: ties keep index order) -- over the whole
    //    other axis, or (ranked) separately inside every panel range
    CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n_rank, stream));
    if (!ranked) {
        if (nnz > 0) count_owners_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(nnz, d_own, cnt);
        iota_kernel<<<blocks_for(n_own, 256), 256, 0, stream>>>(n_own, ids);
        cub::DoubleBuffer<int32_t> cnt_db(cnt, cnt_alt), ids_db(ids, ids_alt);
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, cnt_db, ids_db, (int)n_own, 0, 32, stream));
        assign_slots_kernel<<<blocks_for(n_slots, 256), 256, 0, stream>>>(n_own, n_slots, ids_db.Current(), slot_of,
                                                                         L.own_id);
    } else {
        if (nnz > 0)
            count_owners_ranged_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(nnz, d_own, d_oth, panel_rows,
                                                                                L.panels_per_range, n_own, cnt);
        ranged_keys_kernel<<<blocks_for(n_rank, 256), 256, 0, stream>>>(n_rank, n_own, cnt, rkeys, ids);
        cub::DoubleBuffer<uint64_t> rk_db(rkeys, rkeys_alt);
        cub::DoubleBuffer<int32_t> ids_db(ids, ids_alt);
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, rk_db, ids_db, (int)n_rank, 0, rank_bits, stream));
        assign_slots_ranged_kernel<<<blocks_for(n_ownid, 256), 256, 0, stream>>>(n_own, n_slots, L.nranges,
                                                                                ids_db.Current(), slot_of, L.own_id);
    }
    CUDA_TRY(cudaGetLastError());

    trace_mark(stream, "  owner ranking");
    // 2. sort keys (slot, panel, bank group, local index) and per-list counts
    CUDA_TRY(cudaMemsetAsync(seg_cnt, 0, sizeof(int64_t) * n_seg, stream));
    const uint64_t *vals_sorted = vals;
    if (nnz > 0) {
        make_keys_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(nnz, d_own, d_oth, d_val, slot_of, panel_rows,
                                                                  L.npanel, ranked ? n_own : 0, L.panels_per_range,
                                                                  free_mode, (flags & LAYOUT_ROTATE_CLASS) != 0, keys, vals,
                                                                  seg_cnt);
        trace_mark(stream, "  keys + counts");
        // both buffers of a pair are ours, so the sort ping-pongs between them (O(1) extra storage)
        cub::DoubleBuffer<uint64_t> keys_db(keys, keys_alt), vals_db(vals, vals_alt);
        // Sorted on (list, class) only: the radix sort is stable, so the entries of a list keep their input
        // order (ascending for the usual row-major COO) and the 12 bits of the panel-local row need not be
        // sorted -- 3 radix passes instead of 5 over 16 bytes per nonzero.  Nothing depends on the order inside a
        // list (sums are order-insensitive to rounding only; the round trip is a multiset).
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_db, vals_db, nnz, KEY_LOCAL_BITS, key_bits, stream));
        vals_sorted = vals_db.Current();
        trace_mark(stream, "  radix sort");
    }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, seg_cnt, seg_first, n_seg, stream));

    // 3. padded lengths per (warp, panel) and their prefix sum
    warp_steps_kernel<<<blocks_for(n_ptr, 256), 256, 0, stream>>>(n_warps, L.npanel, opw, free_mode, seg_cnt, pairs);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pairs, L.seg_ptr, n_ptr, stream));
    int64_t total_pairs = 0;
    CUDA_TRY(cudaMemcpyAsync(&total_pairs, L.seg_ptr + (n_ptr - 1), sizeof(int64_t), cudaMemcpyDeviceToHost,
                             stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    L.total_pairs = total_pairs;   // last column of every warp row is 0, so the last prefix is the total
    L.padded_entries = total_pairs * 2 * opw;

    trace_mark(stream, "  scans + lengths");
    // 4. entry stream
    const int64_t n_elems = total_pairs * opw;      // one element = two steps of a lane (pair)
    const size_t elem_bytes = packed ? sizeof(int2) : sizeof(int4);
    CUDA_TRY(pool_malloc(&L.entries, elem_bytes * (n_elems > 0 ? n_elems : 1), stream));
    if (n_elems > 0) {
        if (packed) fill_pad_kernel<true><<<blocks_for(n_elems, 256), 256, 0, stream>>>(n_elems, opw, free_mode, L.entries);
        else fill_pad_kernel<false><<<blocks_for(n_elems, 256), 256, 0, stream>>>(n_elems, opw, free_mode, L.entries);
    }
    if (nnz > 0) {
        const int64_t n_qw = n_slots / 4;
        if (packed)
            place_entries_kernel<true><<<blocks_for(n_qw * L.npanel, 128), 128, 0, stream>>>(
                n_qw, L.npanel, opw, free_mode, yhi, seg_cnt, seg_first, vals_sorted, L.seg_ptr, L.entries, unplaced);
        else
            place_entries_kernel<false><<<blocks_for(n_qw * L.npanel, 128), 128, 0, stream>>>(
                n_qw, L.npanel, opw, free_mode, yhi, seg_cnt, seg_first, vals_sorted, L.seg_ptr, L.entries, unplaced);
    }
    CUDA_TRY(cudaGetLastError());
    trace_mark(stream, "  schedule + place");
    int n_unplaced = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_unplaced, unplaced, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));   // the scratch block is handed back on return
    if (n_unplaced) {
        set_error("layout: %d (quarter warp, panel) schedules left nonzeros unplaced", n_unplaced);
        L.release();
        return SCHPF_ERR_STATE;
    }
    L.bytes = sizeof(int32_t) * n_ownid + sizeof(int64_t) * n_ptr + elem_bytes * n_elems;
    return SCHPF_OK;
}

// ---- debug dump: the stream decoded back to triples (tests: exact integer round trip) -------------
namespace {
template <bool PACKED>
__global__ void dump_entries_kernel(int64_t n_warps, int npanel, int opw, int panel_rows, int panels_per_range,
                                    int64_t own_range_stride, bool yhi, const int32_t *__restrict__ own_id,
                                    const int64_t *__restrict__ seg_ptr, const void *__restrict__ entries_v,
                                    int32_t *__restrict__ own, int32_t *__restrict__ oth, int32_t *__restrict__ cnt)
{
    // one thread per (warp, panel)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_warps * npanel) return;
    const int64_t wg = idx / npanel;
    const int p = (int)(idx - wg * npanel);
    const int64_t i0 = seg_ptr[wg * (npanel + 1) + p], i1 = seg_ptr[wg * (npanel + 1) + p + 1];
    const int64_t range = p / panels_per_range;
    for (int64_t i = i0; i < i1; ++i)
        for (int s = 0; s < opw; ++s) {
            const int pos = stream_pos(s, opw);
            const int32_t o = own_id[range * own_range_stride + wg * opw + s];
            for (int e = 0; e < 2; ++e) {
                const int64_t at = (i * opw + pos) * 2 + e;
                int rowf, field;
                bool pad;
                if (PACKED) {
                    const uint32_t w = reinterpret_cast<const uint32_t *>(entries_v)[at];
                    pad = w >> 31;
                    rowf = (int)(w & ((1u << PACKED_ROW_BITS) - 1));
                    field = (int)((w >> PACKED_ROW_BITS) & ((1u << PACKED_COUNT_BITS) - 1));
                } else {
                    const int2 v = reinterpret_cast<const int2 *>(entries_v)[at];
                    pad = opw == 32 ? v.y < 0 : v.x < 0;
                    rowf = v.x & 0x7fffffff;
                    field = v.y & 0x7fffffff;
                }
                const int c = yhi ? (int)__hiloint2double(field, 0) : field;
                own[at] = pad ? -1 : o;
                oth[at] = pad ? -1 : p * panel_rows + rowf;
                cnt[at] = pad ? 0 : c;
            }
        }
}
}  // namespace

int64_t dump_side_layout(const SideLayout &L, cudaStream_t stream, int32_t *own, int32_t *oth, int32_t *cnt)
{
    const int64_t n = L.padded_entries;
    if (n <= 0) return 0;
    int32_t *d = nullptr;
    if (pool_malloc(reinterpret_cast<void **>(&d), sizeof(int32_t) * 3 * n, stream) != cudaSuccess) return -1;
    const int64_t n_warps = (int64_t)L.nblocks * L.warps;
    const int64_t stride = L.ranked_per_range ? L.n_slots : 0;
    const int blocks = (int)((n_warps * L.npanel + 127) / 128);
    if (L.packed)
        dump_entries_kernel<true><<<blocks, 128, 0, stream>>>(n_warps, L.npanel, L.opw, L.panel_rows, L.panels_per_range,
                                                              stride, L.yhi, L.own_id, L.seg_ptr, L.entries, d, d + n, d + 2 * n);
    else
        dump_entries_kernel<false><<<blocks, 128, 0, stream>>>(n_warps, L.npanel, L.opw, L.panel_rows, L.panels_per_range,
                                                               stride, L.yhi, L.own_id, L.seg_ptr, L.entries, d, d + n, d + 2 * n);
    cudaMemcpyAsync(own, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(oth, d + n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(cnt, d + 2 * n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream);
    const cudaError_t e = cudaStreamSynchronize(stream);
    pool_free(d, stream);
    return e == cudaSuccess ? n : -1;
}

}  // namespace schpf
