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
//   stored step-interleaved (step i of all 16 pairs is contiguous), padded to
//   the longest of the 16 lists rounded up to an even number of steps, two
//   steps per int4 {other_local | pad<<31, y, other_local | pad<<31, y};
//   within a list, nonzeros are ordered by ((other_local - slot) mod 4, other_local)
//   so that the four lane pairs of a quarter warp tend to read shared-memory rows
//   of four different bank groups in the same step.
//
// Integer work only; results are bit-exact and independent of the input order
// of the triples (ties are broken by the other-axis index, duplicates by the
// radix sort's stability).
#include <cub/cub.cuh>

#include "common.cuh"

namespace schpf {

namespace {

constexpr int KEY_LOCAL_BITS = 12;   // panel_rows <= 4096
constexpr int KEY_ROT_BITS = 2;

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

__global__ void count_owners_kernel(int64_t nnz, const int32_t *__restrict__ own, int32_t *__restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(cnt + own[i], 1);
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
                                 int panel_rows, int npanel, uint64_t *__restrict__ keys,
                                 uint64_t *__restrict__ vals, int64_t *__restrict__ seg_cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int32_t slot = slot_of[own[i]];
    const int32_t t = oth[i];
    const int32_t p = t / panel_rows;
    const int32_t tl = t - p * panel_rows;
    const int64_t seg = (int64_t)slot * npanel + p;
    const uint32_t rot = (uint32_t)(tl - slot) & 3u;
    keys[i] = ((uint64_t)seg << (KEY_LOCAL_BITS + KEY_ROT_BITS)) | ((uint64_t)rot << KEY_LOCAL_BITS) |
              (uint64_t)tl;
    vals[i] = ((uint64_t)(uint32_t)val[i] << 32) | (uint64_t)(uint32_t)tl;
    atomicAdd(reinterpret_cast<unsigned long long *>(seg_cnt + seg), 1ULL);
}

// step pairs per (warp, panel) = ceil(max over the warp's 16 lists / 2)
__global__ void warp_steps_kernel(int64_t n_warps, int npanel, const int64_t *__restrict__ seg_cnt,
                                  int64_t *__restrict__ pairs)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_warps * (npanel + 1)) return;
    const int64_t wg = idx / (npanel + 1);
    const int p = (int)(idx - wg * (npanel + 1));
    if (p == npanel) {
        pairs[idx] = 0;
        return;
    }
    int64_t m = 0;
    for (int q = 0; q < GROUPS_PER_WARP; ++q) {
        const int64_t c = seg_cnt[(wg * GROUPS_PER_WARP + q) * npanel + p];
        m = c > m ? c : m;
    }
    pairs[idx] = (m + 1) >> 1;
}

__global__ void fill_pad_kernel(int64_t n_int4, int4 *entries)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_int4) entries[i] = make_int4((int)0x80000000, 0, (int)0x80000000, 0);
}

__global__ void place_entries_kernel(int64_t nnz, const uint64_t *__restrict__ keys,
                                     const uint64_t *__restrict__ vals, const int64_t *__restrict__ seg_first,
                                     const int64_t *__restrict__ seg_ptr, int npanel, int2 *__restrict__ entries)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int64_t seg = (int64_t)(keys[j] >> (KEY_LOCAL_BITS + KEY_ROT_BITS));
    const int64_t slot = seg / npanel;
    const int p = (int)(seg - slot * npanel);
    const int64_t rank = j - seg_first[seg];
    const int64_t wg = slot / GROUPS_PER_WARP;
    const int q = (int)(slot - wg * GROUPS_PER_WARP);
    const int64_t pair = seg_ptr[wg * (npanel + 1) + p] + (rank >> 1);
    const uint64_t v = vals[j];
    entries[(pair * GROUPS_PER_WARP + q) * 2 + (rank & 1)] = make_int2((int)(uint32_t)v, (int)(uint32_t)(v >> 32));
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

}  // namespace

void SideLayout::release()
{
    if (own_id) cudaFree(own_id);
    if (seg_ptr) cudaFree(seg_ptr);
    if (entries) cudaFree(entries);
    own_id = nullptr;
    seg_ptr = nullptr;
    entries = nullptr;
    bytes = 0;
}

int build_side_layout(SideLayout &L, cudaStream_t stream, int64_t nnz, const int32_t *d_own,
                      const int32_t *d_oth, const int32_t *d_val, int64_t n_own, int64_t n_oth,
                      int panel_rows, int warps, int target_ctas)
{
    L.release();
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
    const int64_t owners_per_block = (int64_t)warps * GROUPS_PER_WARP;
    L.nblocks = (int)((n_own + owners_per_block - 1) / owners_per_block);
    if (L.nblocks < 1) L.nblocks = 1;
    int nranges = (target_ctas + L.nblocks - 1) / L.nblocks;
    if (nranges < 1) nranges = 1;
    if (nranges > L.npanel) nranges = L.npanel;
    L.panels_per_range = (L.npanel + nranges - 1) / nranges;
    L.nranges = (L.npanel + L.panels_per_range - 1) / L.panels_per_range;

    const int64_t n_slots = (int64_t)L.nblocks * owners_per_block;
    const int64_t n_warps = (int64_t)L.nblocks * warps;
    const int64_t n_seg = n_slots * L.npanel;
    const int64_t n_ptr = n_warps * (L.npanel + 1);

    CUDA_TRY(cudaMalloc(&L.own_id, sizeof(int32_t) * n_slots));
    CUDA_TRY(cudaMalloc(&L.seg_ptr, sizeof(int64_t) * n_ptr));

    DevBuf cnt, ids, cnt_sorted, ids_sorted, slot_of, seg_cnt, seg_first, pairs, keys, vals, keys2, vals2, tmp;
    CUDA_TRY(cnt.alloc(sizeof(int32_t) * n_own));
    CUDA_TRY(ids.alloc(sizeof(int32_t) * n_own));
    CUDA_TRY(cnt_sorted.alloc(sizeof(int32_t) * n_own));
    CUDA_TRY(ids_sorted.alloc(sizeof(int32_t) * n_own));
    CUDA_TRY(slot_of.alloc(sizeof(int32_t) * n_own));
    CUDA_TRY(seg_cnt.alloc(sizeof(int64_t) * n_seg));
    CUDA_TRY(seg_first.alloc(sizeof(int64_t) * n_seg));
    CUDA_TRY(pairs.alloc(sizeof(int64_t) * n_ptr));
    CUDA_TRY(keys.alloc(sizeof(uint64_t) * nnz));
    CUDA_TRY(vals.alloc(sizeof(uint64_t) * nnz));
    CUDA_TRY(keys2.alloc(sizeof(uint64_t) * nnz));
    CUDA_TRY(vals2.alloc(sizeof(uint64_t) * nnz));

    // 1. owners ranked by nonzero count, descending (stable: ties keep index order)
    CUDA_TRY(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * n_own, stream));
    if (nnz > 0) count_owners_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(nnz, d_own, cnt.as<int32_t>());
    iota_kernel<<<blocks_for(n_own, 256), 256, 0, stream>>>(n_own, ids.as<int32_t>());
    size_t tmp_bytes = 0, need = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, cnt.as<int32_t>(), cnt_sorted.as<int32_t>(),
                                                       ids.as<int32_t>(), ids_sorted.as<int32_t>(), (int)n_own,
                                                       0, 32, stream));
    tmp_bytes = need;
    const int key_bits = KEY_LOCAL_BITS + KEY_ROT_BITS + bits_for((uint64_t)(n_seg > 0 ? n_seg - 1 : 0));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, keys.as<uint64_t>(), keys2.as<uint64_t>(),
                                             vals.as<uint64_t>(), vals2.as<uint64_t>(), nnz, 0, key_bits, stream));
    if (need > tmp_bytes) tmp_bytes = need;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, seg_cnt.as<int64_t>(), seg_first.as<int64_t>(), n_seg,
                                           stream));
    if (need > tmp_bytes) tmp_bytes = need;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, pairs.as<int64_t>(), L.seg_ptr, n_ptr, stream));
    if (need > tmp_bytes) tmp_bytes = need;
    CUDA_TRY(tmp.alloc(tmp_bytes));

    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, cnt.as<int32_t>(),
                                                       cnt_sorted.as<int32_t>(), ids.as<int32_t>(),
                                                       ids_sorted.as<int32_t>(), (int)n_own, 0, 32, stream));
    assign_slots_kernel<<<blocks_for(n_slots, 256), 256, 0, stream>>>(n_own, n_slots, ids_sorted.as<int32_t>(),
                                                                     slot_of.as<int32_t>(), L.own_id);

    // 2. sort keys (slot, panel, rotated bank class, local index) and per-list counts
    CUDA_TRY(cudaMemsetAsync(seg_cnt.p, 0, sizeof(int64_t) * n_seg, stream));
    if (nnz > 0) {
        make_keys_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(
            nnz, d_own, d_oth, d_val, slot_of.as<int32_t>(), panel_rows, L.npanel, keys.as<uint64_t>(),
            vals.as<uint64_t>(), seg_cnt.as<int64_t>());
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.as<uint64_t>(), keys2.as<uint64_t>(),
                                                 vals.as<uint64_t>(), vals2.as<uint64_t>(), nnz, 0, key_bits,
                                                 stream));
    }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, seg_cnt.as<int64_t>(), seg_first.as<int64_t>(), n_seg,
                                           stream));

    // 3. padded lengths per (warp, panel) and their prefix sum
    warp_steps_kernel<<<blocks_for(n_ptr, 256), 256, 0, stream>>>(n_warps, L.npanel, seg_cnt.as<int64_t>(),
                                                                 pairs.as<int64_t>());
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, pairs.as<int64_t>(), L.seg_ptr, n_ptr, stream));
    int64_t total_pairs = 0;
    CUDA_TRY(cudaMemcpyAsync(&total_pairs, L.seg_ptr + (n_ptr - 1), sizeof(int64_t), cudaMemcpyDeviceToHost,
                             stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    L.total_pairs = total_pairs;   // last column of every warp row is 0, so the last prefix is the total
    L.padded_entries = total_pairs * 2 * GROUPS_PER_WARP;

    // 4. entry stream
    const int64_t n_int4 = total_pairs * GROUPS_PER_WARP;
    CUDA_TRY(cudaMalloc(&L.entries, sizeof(int4) * (n_int4 > 0 ? n_int4 : 1)));
    if (n_int4 > 0) fill_pad_kernel<<<blocks_for(n_int4, 256), 256, 0, stream>>>(n_int4, L.entries);
    if (nnz > 0)
        place_entries_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(
            nnz, keys2.as<uint64_t>(), vals2.as<uint64_t>(), seg_first.as<int64_t>(), L.seg_ptr, L.npanel,
            reinterpret_cast<int2 *>(L.entries));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));   // temporaries are freed on return
    L.bytes = sizeof(int32_t) * n_slots + sizeof(int64_t) * n_ptr + sizeof(int4) * n_int4;
    return SCHPF_OK;
}

}  // namespace schpf
