// Engine-level C ABI: the loop body of the reference's scHPF._fit
// (schpf/scHPF_.py:642-715) with the count matrix and the eight variational
// arrays resident in HBM.  See include/schpf_b200.h for the contract.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <string>
#include <vector>

#include "common.cuh"

namespace schpf {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int sweep_ctas_per_sm(int K);
int sweep_default_warps(int K);

cudaError_t pool_malloc(void **p, size_t bytes, cudaStream_t stream)
{
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != configured_device) {
        cudaMemPool_t pool;
        e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        uint64_t keep = UINT64_MAX;      // never hand memory back to the OS between fits
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    return cudaMallocAsync(p, bytes ? bytes : 1, stream);
}

void pool_free(void *p, cudaStream_t stream)
{
    if (p) cudaFreeAsync(p, stream);
}

static thread_local cudaStream_t g_alloc_stream = nullptr;   // stream of the handle being served

void trace_mark(cudaStream_t stream, const char *what)
{
    static const bool on = getenv("SCHPF_TRACE") != nullptr;
    static auto last = std::chrono::steady_clock::now();
    if (!on) return;
    cudaStreamSynchronize(stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[schpf trace] %-28s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

}  // namespace schpf

using namespace schpf;

// ---- NCCL, bound at run time (no link-time dependency) -----------------------
namespace {
struct NcclId128 {          // layout of ncclUniqueId (nccl.h: char internal[128]), passed by value
    char b[128];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *id) = nullptr;
    int (*CommInitRank)(void **comm, int nranks, NcclId128 id, int rank) = nullptr;
    int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, void *comm,
                     cudaStream_t stream) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
typedef decltype(NcclApi::CommInitRank) comm_init_fn;
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // nccl.h: ncclFloat64, ncclSum

int load_nccl()
{
    if (g_nccl.lib) return SCHPF_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
        return SCHPF_ERR_STATE;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<comm_init_fn>(dlsym(lib, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("libnccl.so.2 lacks an expected symbol");
        return SCHPF_ERR_STATE;
    }
    g_nccl.lib = lib;
    return SCHPF_OK;
}

int nccl_check(int rc, const char *what)
{
    if (rc == 0) return SCHPF_OK;
    set_error("%s -> NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return SCHPF_ERR_CUDA;
}
}  // namespace

struct schpf_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t C = 0, G = 0, nnz = 0;
    int K = 0, ST = 0;
    int64_t C_pad = 0, G_pad = 0;
    // sweep family and table geometry (common.cuh TabGeom): lanes = one-lane-per-owner kernels
    bool lanes = false;
    int KA = 0, KB = 0;
    // fp32 sweep (option "precision" = 32, sweep_f32.cu): float copies of the four streamed tables
    bool f32 = false;
    int KF = 0;                                      // floats per row of the fp32 tables
    float *Et32 = nullptr, *Eb32 = nullptr, *Xt32 = nullptr, *Xb32 = nullptr;
    TabGeom geom_t() const { return TabGeom{KA, KB, C_pad * KA, Et32, KF}; }
    TabGeom geom_b() const { return TabGeom{KA, KB, G_pad * KA, Eb32, KF}; }
    TabGeom geom_xt() const { return TabGeom{KA, KB, C_pad * KA, Xt32, KF}; }
    TabGeom geom_xb() const { return TabGeom{KA, KB, G_pad * KA, Xb32, KF}; }

    // options
    int opt_panel_rows = 0;     // 0 = largest that fits
    int opt_warps = 0;          // 0 = the most the register budget allows for this K
    int opt_target_ctas = 4736; // 148 SMs x 2 CTAs x 16 waves
    int opt_variant = 0;
    int opt_timing = 0;
    int opt_packed_entries = -1; // 4-byte stream entries when every count is < 2^19: -1 = one-lane streams only, 0 = never, 1 = lane pairs too
    int opt_lanes = 1;          // 0 = lane-pair kernels for every K (sweep.cu)
    int opt_precision = 64;     // 32 = fp32 sweep for float32 models (sweep_f32.cu); state and updates stay fp64
    int opt_rank_per_range = -1; // -1 = automatic (on for streams without a bank schedule); 0 / 1
    int opt_free_schedule = -1;  // K 17..20 (plane B): 1 = no bank schedule (rotated class order instead), 0 = 4x4 colouring
    int64_t row_offset = 0;     // global index of local cell 0 (random-phi stream)

    bool have_coo = false, have_hyper = false, have_state = false;
    bool tables_t_valid = false, tables_b_valid = false;
    double a = 0, ap = 0, bp = 0, c = 0, cp = 0, dp = 0;

    // state (unpadded, row-major)
    double *theta_shp = nullptr, *theta_rte = nullptr, *beta_shp = nullptr, *beta_rte = nullptr;
    double *xi_shp = nullptr, *xi_rte = nullptr, *eta_shp = nullptr, *eta_rte = nullptr;
    // tables
    double *Et = nullptr, *Eb = nullptr;            // [C_pad x ST], [G_pad x ST]
    double *Xt = nullptr, *Xb = nullptr;            // e_x tables in the same layout (llh sweep)
    double *elog_t = nullptr, *elog_b = nullptr;    // [C x K], [G x K]
    // accumulators: one allocation [acc_t | direct_t | acc_b | direct_b]
    double *accum = nullptr;
    double *acc_t = nullptr, *direct_t = nullptr, *acc_b = nullptr, *direct_b = nullptr;
    double *exch = nullptr;                          // [G*K | K]
    double *colsum_b = nullptr, *colsum_t_next = nullptr;   // [K]
    double *partials = nullptr;                      // llh / lgamma block partials
    int partials_cap = 0;
    double *scalars = nullptr;                       // [0] llh sum, [1] lgamma sum
    unsigned long long *slow_hits = nullptr;
    // queues of underflowed nonzeros, one per sweep direction (common.cuh slow_enqueue); the two
    // counters are the first 16 doubles of `accum`, so the per-iteration memset clears them too
    int4 *slow_q_t = nullptr, *slow_q_b = nullptr;
    unsigned int slow_cap = 0;
    int *overflow = nullptr;     // sticky: a queue overflowed (reported at the next read-out)
    int *flag = nullptr;
    double lgamma_sum = 0.0;

    int64_t tab_t_elems = 0, tab_b_elems = 0;        // allocated sizes of Et/Xt and Eb/Xb
    int32_t *row = nullptr, *col = nullptr, *data = nullptr;
    SideLayout cells, genes;

    // cell sharding with the exchange done by the engine (schpf_comm_attach; not owned)
    void *comm = nullptr;
    // the all-reduce runs on its own stream under the cells-own sweep (option "overlap_exchange")
    int opt_overlap_exchange = 1;
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_folded = nullptr, ev_reduced = nullptr;
    // the two shape sweeps of an iteration are independent: the cells-own one runs on a second stream, so the
    // tail of one grid (a last partial wave of CTAs) is filled by the other's CTAs (option "overlap_sweeps")
    int opt_overlap_sweeps = 1;
    cudaStream_t sstream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // counters
    double n_iterations = 0, n_sweeps = 0, n_shape_sweeps = 0, n_launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<int> ev_mode;            // sweep mode of every recorded event pair
    size_t ev_used = 0;
    double sweep_ms_accum = 0.0, sweep_ms_shape = 0.0, sweep_ms_llh = 0.0;
};

namespace {

#ifndef LANES_PACKED_DEFAULT
#define LANES_PACKED_DEFAULT 1
#endif
#ifndef LANES_FREE20_DEFAULT
#define LANES_FREE20_DEFAULT 1
#endif
constexpr int64_t ACCUM_HEADER = 16;   // doubles in front of the accumulators: the two queue counters

template <typename T>
int dev_alloc(T **p, int64_t n)
{
    CUDA_TRY(pool_malloc(reinterpret_cast<void **>(p), sizeof(T) * (size_t)(n > 0 ? n : 1), g_alloc_stream));
    return SCHPF_OK;
}

template <typename T>
void dev_free(T *&p)
{
    if (p) pool_free(p, g_alloc_stream);
    p = nullptr;
}

int check_handle(schpf_engine *h)
{
    if (!h) {
        set_error("null engine handle");
        return SCHPF_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    g_alloc_stream = h->stream;
    return SCHPF_OK;
}

void free_coo(schpf_engine *h)
{
    dev_free(h->slow_q_t);
    dev_free(h->slow_q_b);
    dev_free(h->row);
    dev_free(h->col);
    dev_free(h->data);
    h->cells.release();
    h->genes.release();
    h->have_coo = false;
}

int launch_one_sweep(schpf_engine *h, int mode, const SideLayout &L, const SweepArgs &args, cudaStream_t stream)
{
    if (h->f32) RC_TRY(launch_f32_sweep(mode, h->K, L, args, stream));
    else if (L.opw == 32) RC_TRY(launch_lane_sweep(mode, h->K, L, args, stream));
    else RC_TRY(launch_sweep(mode, h->K, L, args, stream));
    h->n_sweeps += 1;
    if (mode == SWEEP_SHAPE) h->n_shape_sweeps += 1;
    h->n_launches += 1;
    return SCHPF_OK;
}

// event pair of the timing option, recorded on the engine's stream around what the caller launches
int timing_begin(schpf_engine *h, int mode, cudaEvent_t *e1_out)
{
    *e1_out = nullptr;
    if (!h->opt_timing) return SCHPF_OK;
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        h->ev_pool.emplace_back(a, b);
        h->ev_mode.push_back(0);
    }
    h->ev_mode[h->ev_used] = mode;
    CUDA_TRY(cudaEventRecord(h->ev_pool[h->ev_used].first, h->stream));
    *e1_out = h->ev_pool[h->ev_used].second;
    ++h->ev_used;
    return SCHPF_OK;
}

int ensure_sweep_stream(schpf_engine *h)
{
    if (h->sstream) return SCHPF_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->sstream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    return SCHPF_OK;
}

int timed_sweep(schpf_engine *h, int mode, const SideLayout &L, const SweepArgs &args)
{
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->opt_timing) {
        if (h->ev_used == h->ev_pool.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            h->ev_pool.emplace_back(a, b);
            h->ev_mode.push_back(0);
        }
        h->ev_mode[h->ev_used] = mode;
        e0 = h->ev_pool[h->ev_used].first;
        e1 = h->ev_pool[h->ev_used].second;
        ++h->ev_used;
        CUDA_TRY(cudaEventRecord(e0, h->stream));
    }
    if (h->f32) RC_TRY(launch_f32_sweep(mode, h->K, L, args, h->stream));
    else if (L.opw == 32) RC_TRY(launch_lane_sweep(mode, h->K, L, args, h->stream));
    else RC_TRY(launch_sweep(mode, h->K, L, args, h->stream));
    if (h->opt_timing) CUDA_TRY(cudaEventRecord(e1, h->stream));
    h->n_sweeps += 1;
    if (mode == SWEEP_SHAPE) h->n_shape_sweeps += 1;
    h->n_launches += 1;
    return SCHPF_OK;
}

int collect_timing(schpf_engine *h)
{
    if (h->ev_used == 0) return SCHPF_OK;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->ev_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_pool[i].first, h->ev_pool[i].second));
        h->sweep_ms_accum += ms;
        (h->ev_mode[i] == SWEEP_SHAPE ? h->sweep_ms_shape : h->sweep_ms_llh) += ms;
    }
    h->ev_used = 0;
    return SCHPF_OK;
}

SweepArgs side_args(schpf_engine *h, const SideLayout &L)
{
    SweepArgs A;
    memset(&A, 0, sizeof(A));
    A.own_id = L.own_id;
    A.seg_ptr = L.seg_ptr;
    A.entries = L.entries;
    A.slow_hits = h->slow_hits;
    A.slow_queue = &L == &h->cells ? h->slow_q_t : h->slow_q_b;
    A.slow_count = reinterpret_cast<unsigned long long *>(h->accum) + (&L == &h->cells ? 0 : 1);
    A.slow_cap = h->slow_cap;
    A.K = h->K;
    A.npanel = L.npanel;
    A.panel_rows = L.panel_rows;
    A.warps = L.warps;
    A.panels_per_range = L.panels_per_range;
    A.nranges = L.nranges;
    A.own_range_stride = L.ranked_per_range ? L.n_slots : 0;
    // plane B of the owner's table and of the streamed table (lanes geometry; 0-sized otherwise)
    const bool cells_own = &L == &h->cells;
    A.own_offB = (cells_own ? h->C_pad : h->G_pad) * h->KA;
    A.oth_offB = (cells_own ? h->G_pad : h->C_pad) * h->KA;
    return A;
}

// Elog / factored tables and column sums from the current state
// (hpf_numba.py:83-94; first half of compute_loading_rate_update :167-170)
int ensure_tables(schpf_engine *h)
{
    const int K = h->K;
    if (!h->tables_t_valid) {
        CUDA_TRY(cudaMemsetAsync(h->colsum_t_next, 0, sizeof(double) * K, h->stream));
        RC_TRY(launch_prep_side(h->stream, h->C, K, h->geom_t(), h->theta_shp, h->theta_rte, h->elog_t, h->Et,
                                h->colsum_t_next));
        h->n_launches += 1;
        h->tables_t_valid = true;
    }
    if (!h->tables_b_valid) {
        CUDA_TRY(cudaMemsetAsync(h->colsum_b, 0, sizeof(double) * K, h->stream));
        RC_TRY(launch_prep_side(h->stream, h->G, K, h->geom_b(), h->beta_shp, h->beta_rte, h->elog_b, h->Eb,
                                h->colsum_b));
        h->n_launches += 1;
        h->tables_b_valid = true;
    }
    return SCHPF_OK;
}

int require_ready(schpf_engine *h)
{
    if (!h->have_coo || !h->have_hyper || !h->have_state) {
        set_error("engine not ready: set_coo=%d set_hyper=%d set_state=%d", (int)h->have_coo,
                  (int)h->have_hyper, (int)h->have_state);
        return SCHPF_ERR_STATE;
    }
    return SCHPF_OK;
}

int zero_accumulators(schpf_engine *h, bool genes_too)
{
    const size_t nt = (size_t)h->C * h->K * 2, nb = (size_t)h->G * h->K * 2;
    CUDA_TRY(cudaMemsetAsync(h->accum, 0, sizeof(double) * (ACCUM_HEADER + nt + (genes_too ? nb : 0)), h->stream));
    return SCHPF_OK;
}

int cells_sweep(schpf_engine *h, SweepArgs *defer_fixup = nullptr, cudaStream_t untimed_on = nullptr)
{
    // theta side: cells own, gene panels stream through shared memory
    SweepArgs A = side_args(h, h->cells);
    A.own_tab = h->f32 ? reinterpret_cast<const double *>(h->Et32) : h->Et;
    A.oth_tab = h->f32 ? reinterpret_cast<const double *>(h->Eb32) : h->Eb;
    A.acc = h->acc_t;
    A.own_elog = h->elog_t;
    A.oth_elog = h->elog_b;
    A.direct = h->direct_t;
    // untimed_on: launched on that stream with no events of its own (the caller times the concurrent pair)
    if (untimed_on) RC_TRY(launch_one_sweep(h, SWEEP_SHAPE, h->cells, A, untimed_on));
    else RC_TRY(timed_sweep(h, SWEEP_SHAPE, h->cells, A));
    if (defer_fixup) {           // the caller redoes both directions' queues with one launch
        *defer_fixup = A;
        return SCHPF_OK;
    }
    h->n_launches += 1;
    return launch_slow_fixup(h->stream, A, nullptr, h->overflow);
}

int genes_sweep(schpf_engine *h, SweepArgs *defer_fixup = nullptr, bool untimed = false)
{
    // beta side: genes own, cell panels stream
    SweepArgs B = side_args(h, h->genes);
    B.own_tab = h->f32 ? reinterpret_cast<const double *>(h->Eb32) : h->Eb;
    B.oth_tab = h->f32 ? reinterpret_cast<const double *>(h->Et32) : h->Et;
    B.acc = h->acc_b;
    B.own_elog = h->elog_b;
    B.oth_elog = h->elog_t;
    B.direct = h->direct_b;
    if (untimed) RC_TRY(launch_one_sweep(h, SWEEP_SHAPE, h->genes, B, h->stream));
    else RC_TRY(timed_sweep(h, SWEEP_SHAPE, h->genes, B));
    if (defer_fixup) {
        *defer_fixup = B;
        return SCHPF_OK;
    }
    h->n_launches += 1;
    return launch_slow_fixup(h->stream, B, nullptr, h->overflow);
}

// this shard's beta shape sums + column sums of theta.e_x (theta BEFORE its update,
// scHPF_.py:701-703) into the exchange buffer
int fold_exchange_buffer(schpf_engine *h, bool fold = true)
{
    const int K = h->K;
    // fold == false (one GPU, schpf_step): the beta finalisation reads the accumulators itself, only the
    // K column sums of theta.e_x are parked in the buffer's tail (theta's finalisation overwrites its own copy)
    if (fold) RC_TRY(launch_fold(h->stream, h->G, K, h->geom_b(), h->Eb, h->acc_b, h->direct_b, h->exch));
    CUDA_TRY(cudaMemcpyAsync(h->exch + (size_t)h->G * K, h->colsum_t_next, sizeof(double) * K,
                             cudaMemcpyDeviceToDevice, h->stream));
    if (fold) h->n_launches += 1;
    return SCHPF_OK;
}

// mode 0: E-step from the resident state; 1: random phi; 2: Xphi supplied (already scattered by caller)
int step_begin_impl(schpf_engine *h, int flags, int mode, uint64_t seed, bool fold = true)
{
    const bool freeze = flags & SCHPF_FREEZE_GENES;
    const int K = h->K;
    RC_TRY(ensure_tables(h));
    if (mode != 2) RC_TRY(zero_accumulators(h, !freeze));
    if (mode == 1) {
        RC_TRY(launch_random_phi(h->stream, h->nnz, K, h->row, h->col, h->data, seed, h->row_offset,
                                 h->direct_t, freeze ? nullptr : h->direct_b));
        h->n_launches += 1;
    } else if (mode == 0) {
        if (h->opt_variant == 1) {
            RC_TRY(launch_literal(h->stream, h->nnz, K, h->row, h->col, h->data, h->elog_t, h->elog_b,
                                  nullptr, h->direct_t, freeze ? nullptr : h->direct_b));
            h->n_launches += 1;
        } else {
            // both directions' underflow queues are redone by ONE launch (empty queues cost one read each)
            SweepArgs A, B;
            if (!freeze && h->opt_overlap_sweeps) {
                // cells-own sweep on the second stream, genes-own on the engine's: one train of CTAs, one tail
                RC_TRY(ensure_sweep_stream(h));
                cudaEvent_t e1 = nullptr;
                RC_TRY(timing_begin(h, SWEEP_SHAPE, &e1));
                CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
                CUDA_TRY(cudaStreamWaitEvent(h->sstream, h->ev_fork, 0));
                RC_TRY(cells_sweep(h, &A, h->sstream));
                RC_TRY(genes_sweep(h, &B, true));
                CUDA_TRY(cudaEventRecord(h->ev_join, h->sstream));
                CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
                if (e1) CUDA_TRY(cudaEventRecord(e1, h->stream));
            } else {
                RC_TRY(cells_sweep(h, &A));
                if (!freeze) RC_TRY(genes_sweep(h, &B));
            }
            RC_TRY(launch_slow_fixup(h->stream, A, freeze ? nullptr : &B, h->overflow));
            h->n_launches += 1;
        }
    }
    if (!freeze) RC_TRY(fold_exchange_buffer(h, fold));
    return SCHPF_OK;
}

// the one exchange step of an iteration, when the engine owns the communicator
int allreduce_exchange(schpf_engine *h, cudaStream_t stream)
{
    return nccl_check(g_nccl.AllReduce(h->exch, h->exch, (size_t)(h->G * h->K + h->K), NCCL_FLOAT64, NCCL_SUM,
                                       h->comm, stream),
                      "ncclAllReduce(exchange buffer)");
}

// the minibatch order exchanges AFTER the cell update (inside step_end_impl), not between the phases
inline bool exchange_is_late(int flags) { return (flags & SCHPF_CELLS_FIRST) && !(flags & SCHPF_SIMULTANEOUS); }

int step_end_impl(schpf_engine *h, int flags, bool folded = true)
{
    const bool freeze = flags & SCHPF_FREEZE_GENES;
    const bool simultaneous = flags & SCHPF_SIMULTANEOUS;
    const bool cells_first = flags & SCHPF_CELLS_FIRST;
    const int K = h->K;
    // split-phase callers of the minibatch order on a sharded engine: PHASE_CELLS = theta/xi and the
    // new theta's column sums into the exchange buffer, (all-reduce by the caller,) PHASE_GENES = beta/eta
    const bool only_cells = flags & SCHPF_PHASE_CELLS, only_genes = flags & SCHPF_PHASE_GENES;
    if ((only_cells || only_genes) && (!cells_first || simultaneous || (only_cells && only_genes))) {
        set_error("SCHPF_PHASE_CELLS / SCHPF_PHASE_GENES split the SCHPF_CELLS_FIRST order only (one of them per call)");
        return SCHPF_ERR_ARG;
    }
    auto theta_update = [&]() -> int {
        // scHPF_.py:709-714: theta shape from the row sums, rate from xi (old) + column sums of
        // beta.e_x, then xi rate; also next iteration's tables and theta.e_x column sums
        CUDA_TRY(cudaMemsetAsync(h->colsum_t_next, 0, sizeof(double) * K, h->stream));
        RC_TRY(launch_finalize(h->stream, h->C, K, h->geom_t(), h->a, h->bp, nullptr, h->Et, h->acc_t, h->direct_t,
                               h->colsum_b, h->xi_shp, h->xi_rte, h->theta_shp, h->theta_rte, h->elog_t,
                               h->Et, h->colsum_t_next));
        h->n_launches += 1;
        return SCHPF_OK;
    };
    auto beta_update = [&]() -> int {
        // scHPF_.py:699-704
        CUDA_TRY(cudaMemsetAsync(h->colsum_b, 0, sizeof(double) * K, h->stream));
        // folded: the (all-reduced) exchange buffer holds the genes' sums; otherwise they are still Eb * acc_b +
        // direct_b -- the same fma the fold kernel would have done, so the result is bit-identical
        RC_TRY(launch_finalize(h->stream, h->G, K, h->geom_b(), h->c, h->dp, folded ? h->exch : nullptr, h->Eb,
                               folded ? nullptr : h->acc_b, folded ? nullptr : h->direct_b,
                               h->exch + (size_t)h->G * K, h->eta_shp, h->eta_rte, h->beta_shp, h->beta_rte,
                               h->elog_b, h->Eb, h->colsum_b));
        h->n_launches += 1;
        return SCHPF_OK;
    };
    if (simultaneous) {
        // scHPF_.py:666-684: cell updates see the OLD beta, gene updates the OLD theta
        RC_TRY(theta_update());
        if (!freeze) RC_TRY(beta_update());
    } else if (cells_first) {
        // scHPF_.py:686-704 (`batched`): cell updates first, from the OLD beta; beta's rate then sums
        // theta.e_x of the NEW theta, which theta_update has just left in colsum_t_next
        if (!only_genes) {
            RC_TRY(theta_update());
            if (!freeze)
                CUDA_TRY(cudaMemcpyAsync(h->exch + (size_t)h->G * K, h->colsum_t_next, sizeof(double) * K,
                                         cudaMemcpyDeviceToDevice, h->stream));
        }
        if (only_cells) return SCHPF_OK;
        if (!freeze) {
            // sharded minibatch: the batch's cells are spread over the ranks; beta needs the sums of all of them
            if (h->comm && !only_genes) RC_TRY(allreduce_exchange(h, h->stream));
            RC_TRY(beta_update());
        }
    } else {
        if (!freeze) RC_TRY(beta_update());
        RC_TRY(theta_update());
    }
    h->n_iterations += 1;
    return SCHPF_OK;
}

// after a synchronisation: did a queue of underflowed nonzeros overflow since the last check?
int check_overflow(schpf_engine *h)
{
    int f = 0;
    CUDA_TRY(cudaMemcpyAsync(&f, h->overflow, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (f) {
        CUDA_TRY(cudaMemsetAsync(h->overflow, 0, sizeof(int), h->stream));
        set_error("more than %u nonzeros per sweep underflowed the factored softmax (priors far below 1e-2?): "
                  "results since the last check are incomplete; use option variant=1 (literal kernel) for this model",
                  h->slow_cap);
        return SCHPF_ERR_NUMERIC;
    }
    return SCHPF_OK;
}

int upload(double *dst, const double *src, int64_t n, cudaStream_t s)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    return SCHPF_OK;
}

int download(double *dst, const double *src, int64_t n, cudaStream_t s)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    return SCHPF_OK;
}

int finish_coo(schpf_engine *h)
{
    // validate indices, take the data constant sum lgamma(y+1), build both layouts
    trace_mark(h->stream, "coo copy");
    CUDA_TRY(cudaMemsetAsync(h->flag, 0, sizeof(int), h->stream));
    RC_TRY(launch_validate_coo(h->stream, h->nnz, h->row, h->col, h->data, h->C, h->G, h->flag));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, h->flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (flag & 7) {
        set_error("COO triples out of range:%s%s%s (matrix is %lld x %lld)", (flag & 1) ? " row" : "",
                  (flag & 2) ? " col" : "", (flag & 4) ? " negative count" : "", (long long)h->C,
                  (long long)h->G);
        free_coo(h);
        return SCHPF_ERR_ARG;
    }
    // sweep family for this K: one lane per owner where it is instantiated, lane pairs otherwise
    h->f32 = h->opt_precision == 32 && h->opt_variant == 0;
    h->KF = h->f32 ? f32_row_floats(h->K) : 0;
    h->lanes = !h->f32 && h->opt_lanes && h->opt_variant == 0 && lanes_supported(h->K);
    if (h->lanes) {
        const int kp = lanes_kp_of(h->K);
        h->KB = kp == 20 ? 4 : 0;
        h->KA = kp - h->KB;
    } else {
        h->KA = h->ST;
        h->KB = 0;
    }
    const int TW = h->KA + h->KB;     // doubles per table row
    const int ctas = sweep_ctas_per_sm(h->K);
    const int Pmax = h->f32 ? f32_max_panel_rows(h->K) : h->lanes ? lanes_max_panel_rows(h->K) : max_panel_rows(h->K, 1);
    int Po = h->opt_panel_rows > 0 ? h->opt_panel_rows : (h->lanes || h->f32) ? Pmax : max_panel_rows(h->K, ctas);
    if (Po > Pmax) Po = Pmax;
    Po &= ~3;
    if (Po < 4) Po = 4;
    // tables are streamed panel-wise: pad their row counts to whole panels (zero rows)
    const int64_t C_pad = ((h->C + Po - 1) / Po) * Po, G_pad = ((h->G + Po - 1) / Po) * Po;
    if (C_pad * TW != h->tab_t_elems || !h->Et) {
        dev_free(h->Et);
        dev_free(h->Xt);
        RC_TRY(dev_alloc(&h->Et, C_pad * TW));
        RC_TRY(dev_alloc(&h->Xt, C_pad * TW));
        h->tab_t_elems = C_pad * TW;
    }
    if (G_pad * TW != h->tab_b_elems || !h->Eb) {
        dev_free(h->Eb);
        dev_free(h->Xb);
        RC_TRY(dev_alloc(&h->Eb, G_pad * TW));
        RC_TRY(dev_alloc(&h->Xb, G_pad * TW));
        h->tab_b_elems = G_pad * TW;
    }
    h->C_pad = C_pad;
    h->G_pad = G_pad;
    dev_free(h->Et32);
    dev_free(h->Eb32);
    dev_free(h->Xt32);
    dev_free(h->Xb32);
    if (h->f32) {
        const size_t nt = (size_t)C_pad * h->KF, nb = (size_t)G_pad * h->KF;
        RC_TRY(dev_alloc(&h->Et32, (int64_t)nt));
        RC_TRY(dev_alloc(&h->Xt32, (int64_t)nt));
        RC_TRY(dev_alloc(&h->Eb32, (int64_t)nb));
        RC_TRY(dev_alloc(&h->Xb32, (int64_t)nb));
        CUDA_TRY(cudaMemsetAsync(h->Et32, 0, sizeof(float) * nt, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->Xt32, 0, sizeof(float) * nt, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->Eb32, 0, sizeof(float) * nb, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->Xb32, 0, sizeof(float) * nb, h->stream));
    }
    // pad rows / pad columns of the streamed tables stay zero for the lifetime of the layout
    CUDA_TRY(cudaMemsetAsync(h->Et, 0, sizeof(double) * (size_t)C_pad * TW, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->Eb, 0, sizeof(double) * (size_t)G_pad * TW, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->Xt, 0, sizeof(double) * (size_t)C_pad * TW, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->Xb, 0, sizeof(double) * (size_t)G_pad * TW, h->stream));
    h->tables_t_valid = h->tables_b_valid = false;

    int warps = h->f32 ? f32_default_warps(h->K) : h->lanes ? lanes_default_warps(h->K) : sweep_default_warps(h->K);
    if (h->opt_warps > 0 && h->opt_warps < warps) warps = h->opt_warps;
    int opw = GROUPS_PER_WARP, lflags = 0;
    if (h->lanes) {
        opw = 32;
        // plane A rows all start at bank group 0: no schedule needed.  Plane B (K 17..20) is either
        // scheduled (4x4 colouring, pads) or left free with its reads ordered by rotated bank class
        // (a few conflicts on 2 of the 10 row loads, no pads)
        const bool free_mode = h->KB == 0 || (h->opt_free_schedule < 0 ? LANES_FREE20_DEFAULT : h->opt_free_schedule != 0);
        if (free_mode) lflags |= LAYOUT_FREE | (h->KB ? LAYOUT_ROTATE_CLASS : 0);
        if (!(flag & 16)) lflags |= LAYOUT_YHI;            // every count < 2^21
        // without a bank schedule the only padding left is the spread of the owners' counts inside a
        // panel: rank the owners panel by panel (one panel per CTA) and it all but disappears
        // (measured at K=20 with the plane-B schedule too: 19 % -> 15 % pads, 3.25 -> 3.15 ms per pair)
        const bool rank = h->opt_rank_per_range < 0 ? true : h->opt_rank_per_range != 0;
        if (rank) lflags |= LAYOUT_RANK_PER_RANGE | LAYOUT_SINGLE_PANEL_RANGES;
    } else if (h->f32) {
        // every fp32 row starts at bank group 0: no bank schedule; counts stay plain integers
        opw = 32;
        lflags |= LAYOUT_FREE;
        if (h->opt_rank_per_range != 0) lflags |= LAYOUT_RANK_PER_RANGE | LAYOUT_SINGLE_PANEL_RANGES;
    }
    trace_mark(h->stream, "validate + tables");
    // 4-byte entries (common.cuh pack_entry) are possible when every count fits 19 bits (flag bit 8 = some
    // count >= 2^19) and a panel has at most 2^12 rows; they halve the entry stream and the resident layout.
    // Measured on cfg-3, sweep pair packed / wide (gpurun_out/r2o_*, r2m_*; decode at the point of use):
    //   one-lane K <= 16   1.98 / 2.02 ms   fp32 (any K)  1.93 / 1.94, 1.90 / 1.92   -> packed by default
    //   one-lane K 17..20  2.47 / 2.41      K 29..32      4.40 / 4.01 (the I2F.F64 of the count lands on the
    //   fp64 pipe these kernels are short of; wide entries carry the count as the high word of its double)
    //   lane pairs (round 1) 3.61 / 3.43                                                 -> wide by default
    // "packed_entries": -1 = these defaults, 0 = always wide, 1 = packed wherever the stream allows it.
    const bool lane_stream = (h->lanes || h->f32) && (lflags & LAYOUT_FREE);
    const bool packed_by_default = LANES_PACKED_DEFAULT && lane_stream && (h->f32 || lanes_kp_of(h->K) == 16);
    const bool want_packed = h->opt_packed_entries == 1 ? (lane_stream || !(h->lanes || h->f32))
                                                        : h->opt_packed_entries < 0 && packed_by_default;
    const bool packed = want_packed && !(flag & 8) && Po <= (1 << PACKED_ROW_BITS);
    if (packed) lflags &= ~LAYOUT_YHI;
    RC_TRY(build_side_layout(h->cells, h->stream, h->nnz, h->row, h->col, h->data, h->C, h->G, Po, warps,
                             h->opt_target_ctas, packed, opw, lflags));
    trace_mark(h->stream, "layout cells total");
    RC_TRY(build_side_layout(h->genes, h->stream, h->nnz, h->col, h->row, h->data, h->G, h->C, Po, warps,
                             h->opt_target_ctas, packed, opw, lflags));

    trace_mark(h->stream, "layout genes total");
    // room for every nonzero of small matrices, 4 M entries (64 MB per direction) beyond
    h->slow_cap = (unsigned int)(h->nnz < 1 ? 1 : h->nnz > (1 << 22) ? (1 << 22) : h->nnz);
    RC_TRY(dev_alloc(&h->slow_q_t, (int64_t)h->slow_cap));
    RC_TRY(dev_alloc(&h->slow_q_b, (int64_t)h->slow_cap));
    int need = h->cells.nblocks * h->cells.nranges;
    if (need < 1024) need = 1024;
    if (need > h->partials_cap) {
        dev_free(h->partials);
        RC_TRY(dev_alloc(&h->partials, need));
        h->partials_cap = need;
    }
    RC_TRY(launch_lgamma_sum(h->stream, h->nnz, h->data, h->partials, 1024, h->scalars + 1));
    CUDA_TRY(cudaMemcpyAsync(&h->lgamma_sum, h->scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->have_coo = true;
    trace_mark(h->stream, "lgamma sum");
    return SCHPF_OK;
}

}  // namespace

extern "C" {

int schpf_version(void) { return 210; }

const char *schpf_last_error(void) { return g_last_error.c_str(); }

int schpf_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
        return -SCHPF_ERR_CUDA;
    }
    return n;
}

int schpf_release_scratch(int device) { return release_layout_scratch(device); }

int schpf_create(schpf_engine_t **out, int device, int64_t ncells, int64_t ngenes, int nfactors, void *stream)
{
    if (!out) {
        set_error("null output handle");
        return SCHPF_ERR_ARG;
    }
    *out = nullptr;
    if (ncells <= 0 || ngenes <= 0 || nfactors <= 0 || nfactors > SCHPF_MAX_FACTORS) {
        set_error("bad dimensions: ncells=%lld ngenes=%lld nfactors=%d (1 <= K <= %d)", (long long)ncells,
                  (long long)ngenes, nfactors, SCHPF_MAX_FACTORS);
        return SCHPF_ERR_ARG;
    }
    if (ncells > 0x7fffffffLL || ngenes > 0x7fffffffLL) {
        set_error("indices are int32: ncells and ngenes must be < 2^31");
        return SCHPF_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    schpf_engine *h = new schpf_engine();
    h->device = device;
    h->stream = reinterpret_cast<cudaStream_t>(stream);
    g_alloc_stream = h->stream;
    h->C = ncells;
    h->G = ngenes;
    h->K = nfactors;
    h->ST = stride_of_kp(kp_of(nfactors));
    const int64_t CK = ncells * nfactors, GK = ngenes * nfactors;
    int rc = SCHPF_OK;
    auto A = [&](int r) { if (rc == SCHPF_OK) rc = r; };
    A(dev_alloc(&h->theta_shp, CK));
    A(dev_alloc(&h->theta_rte, CK));
    A(dev_alloc(&h->beta_shp, GK));
    A(dev_alloc(&h->beta_rte, GK));
    A(dev_alloc(&h->xi_shp, ncells));
    A(dev_alloc(&h->xi_rte, ncells));
    A(dev_alloc(&h->eta_shp, ngenes));
    A(dev_alloc(&h->eta_rte, ngenes));
    A(dev_alloc(&h->elog_t, CK));
    A(dev_alloc(&h->elog_b, GK));
    A(dev_alloc(&h->accum, ACCUM_HEADER + 2 * CK + 2 * GK));
    A(dev_alloc(&h->overflow, (int64_t)1));
    A(dev_alloc(&h->exch, GK + nfactors));
    A(dev_alloc(&h->colsum_b, (int64_t)nfactors));
    A(dev_alloc(&h->colsum_t_next, (int64_t)nfactors));
    A(dev_alloc(&h->scalars, (int64_t)4));
    A(dev_alloc(&h->slow_hits, (int64_t)1));
    A(dev_alloc(&h->flag, (int64_t)1));
    if (rc != SCHPF_OK) {
        schpf_destroy(h);
        return rc;
    }
    h->acc_t = h->accum + ACCUM_HEADER;
    h->direct_t = h->acc_t + CK;
    h->acc_b = h->acc_t + 2 * CK;
    h->direct_b = h->acc_t + 2 * CK + GK;
    cudaMemsetAsync(h->slow_hits, 0, sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->overflow, 0, sizeof(int), h->stream);
    cudaMemsetAsync(h->exch, 0, sizeof(double) * (size_t)(GK + nfactors), h->stream);
    *out = h;
    return SCHPF_OK;
}

int schpf_destroy(schpf_engine_t *h)
{
    if (!h) return SCHPF_OK;
    cudaSetDevice(h->device);
    g_alloc_stream = h->stream;
    cudaStreamSynchronize(h->stream);
    h->comm = nullptr;       // not owned
    if (h->xstream) {
        cudaStreamSynchronize(h->xstream);
        cudaStreamDestroy(h->xstream);
        cudaEventDestroy(h->ev_folded);
        cudaEventDestroy(h->ev_reduced);
        h->xstream = nullptr;
    }
    if (h->sstream) {
        cudaStreamSynchronize(h->sstream);
        cudaStreamDestroy(h->sstream);
        cudaEventDestroy(h->ev_fork);
        cudaEventDestroy(h->ev_join);
        h->sstream = nullptr;
    }
    free_coo(h);
    dev_free(h->theta_shp); dev_free(h->theta_rte); dev_free(h->beta_shp); dev_free(h->beta_rte);
    dev_free(h->xi_shp); dev_free(h->xi_rte); dev_free(h->eta_shp); dev_free(h->eta_rte);
    dev_free(h->Et); dev_free(h->Eb); dev_free(h->Xt); dev_free(h->Xb); dev_free(h->elog_t); dev_free(h->elog_b);
    dev_free(h->Et32); dev_free(h->Eb32); dev_free(h->Xt32); dev_free(h->Xb32);
    dev_free(h->accum); dev_free(h->exch); dev_free(h->colsum_b); dev_free(h->colsum_t_next);
    dev_free(h->partials); dev_free(h->scalars); dev_free(h->slow_hits); dev_free(h->flag);
    dev_free(h->overflow);
    for (auto &e : h->ev_pool) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    delete h;
    return SCHPF_OK;
}

int schpf_set_option(schpf_engine_t *h, const char *key, int64_t value)
{
    RC_TRY(check_handle(h));
    if (!key) {
        set_error("null option key");
        return SCHPF_ERR_ARG;
    }
    if (!strcmp(key, "panel_rows")) h->opt_panel_rows = (int)value;
    else if (!strcmp(key, "warps_per_cta")) h->opt_warps = (int)value;
    else if (!strcmp(key, "target_ctas")) h->opt_target_ctas = (int)value;
    else if (!strcmp(key, "variant")) h->opt_variant = (int)value;
    else if (!strcmp(key, "timing")) h->opt_timing = (int)value;
    else if (!strcmp(key, "packed_entries")) h->opt_packed_entries = (int)value;
    else if (!strcmp(key, "row_offset")) h->row_offset = value;
    else if (!strcmp(key, "overlap_exchange")) h->opt_overlap_exchange = (int)value;
    else if (!strcmp(key, "overlap_sweeps")) h->opt_overlap_sweeps = (int)value;
    else if (!strcmp(key, "lanes")) h->opt_lanes = (int)value;
    else if (!strcmp(key, "precision")) {
        if (value != 32 && value != 64) {
            set_error("precision must be 32 or 64");
            return SCHPF_ERR_ARG;
        }
        if (h->have_coo && (int)value != h->opt_precision) {
            set_error("precision must be set before the matrix (it selects the layout)");
            return SCHPF_ERR_STATE;
        }
        h->opt_precision = (int)value;
    }
    else if (!strcmp(key, "rank_per_range")) h->opt_rank_per_range = (int)value;
    else if (!strcmp(key, "free_schedule")) h->opt_free_schedule = (int)value;
    else {
        set_error("unknown option '%s'", key);
        return SCHPF_ERR_ARG;
    }
    if (h->opt_warps < 0 || h->opt_warps > 16) {
        set_error("warps_per_cta must be in [0, 16] (0 = automatic)");
        h->opt_warps = 0;
        return SCHPF_ERR_ARG;
    }
    return SCHPF_OK;
}

static int set_coo_common(schpf_engine_t *h, const int32_t *row, const int32_t *col, const int32_t *data,
                          int64_t nnz, cudaMemcpyKind kind)
{
    RC_TRY(check_handle(h));
    if (nnz < 0 || (nnz > 0 && (!row || !col || !data))) {
        set_error("bad COO arguments (nnz=%lld)", (long long)nnz);
        return SCHPF_ERR_ARG;
    }
    if (nnz > 0x7fffffffLL) {
        set_error("nnz must be < 2^31 per engine (shard the cells)");
        return SCHPF_ERR_ARG;
    }
    free_coo(h);
    h->nnz = nnz;
    RC_TRY(dev_alloc(&h->row, nnz));
    RC_TRY(dev_alloc(&h->col, nnz));
    RC_TRY(dev_alloc(&h->data, nnz));
    if (nnz > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->row, row, sizeof(int32_t) * nnz, kind, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->col, col, sizeof(int32_t) * nnz, kind, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->data, data, sizeof(int32_t) * nnz, kind, h->stream));
    }
    return finish_coo(h);
}

int schpf_set_coo(schpf_engine_t *h, const int32_t *row, const int32_t *col, const int32_t *data, int64_t nnz)
{
    return set_coo_common(h, row, col, data, nnz, cudaMemcpyHostToDevice);
}

int schpf_set_coo_device(schpf_engine_t *h, const int32_t *d_row, const int32_t *d_col, const int32_t *d_data,
                         int64_t nnz)
{
    return set_coo_common(h, d_row, d_col, d_data, nnz, cudaMemcpyDeviceToDevice);
}

int schpf_set_hyper(schpf_engine_t *h, double a, double ap, double bp, double c, double cp, double dp)
{
    RC_TRY(check_handle(h));
    if (!(a > 0) || !(ap > 0) || !(bp > 0) || !(c > 0) || !(cp > 0) || !(dp > 0)) {
        set_error("hyperparameters must be positive: a=%g ap=%g bp=%g c=%g cp=%g dp=%g", a, ap, bp, c, cp, dp);
        return SCHPF_ERR_ARG;
    }
    h->a = a; h->ap = ap; h->bp = bp; h->c = c; h->cp = cp; h->dp = dp;
    h->have_hyper = true;
    return SCHPF_OK;
}

int schpf_set_state(schpf_engine_t *h, const double *theta_shp, const double *theta_rte, const double *beta_shp,
                    const double *beta_rte, const double *xi_shp, const double *xi_rte, const double *eta_shp,
                    const double *eta_rte)
{
    RC_TRY(check_handle(h));
    const int64_t CK = h->C * h->K, GK = h->G * h->K;
    if ((!theta_shp) != (!theta_rte) || (!beta_shp) != (!beta_rte) || (!xi_shp) != (!xi_rte) ||
        (!eta_shp) != (!eta_rte)) {
        set_error("shape and rate of a distribution must be given together");
        return SCHPF_ERR_ARG;
    }
    if (theta_shp) {
        RC_TRY(upload(h->theta_shp, theta_shp, CK, h->stream));
        RC_TRY(upload(h->theta_rte, theta_rte, CK, h->stream));
        h->tables_t_valid = false;
    }
    if (beta_shp) {
        RC_TRY(upload(h->beta_shp, beta_shp, GK, h->stream));
        RC_TRY(upload(h->beta_rte, beta_rte, GK, h->stream));
        h->tables_b_valid = false;
    }
    if (xi_shp) {
        RC_TRY(upload(h->xi_shp, xi_shp, h->C, h->stream));
        RC_TRY(upload(h->xi_rte, xi_rte, h->C, h->stream));
    }
    if (eta_shp) {
        RC_TRY(upload(h->eta_shp, eta_shp, h->G, h->stream));
        RC_TRY(upload(h->eta_rte, eta_rte, h->G, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // host buffers may be reused on return
    h->have_state = true;
    return SCHPF_OK;
}

int schpf_get_state(schpf_engine_t *h, double *theta_shp, double *theta_rte, double *beta_shp, double *beta_rte,
                    double *xi_shp, double *xi_rte, double *eta_shp, double *eta_rte)
{
    RC_TRY(check_handle(h));
    const int64_t CK = h->C * h->K, GK = h->G * h->K;
    if (theta_shp) RC_TRY(download(theta_shp, h->theta_shp, CK, h->stream));
    if (theta_rte) RC_TRY(download(theta_rte, h->theta_rte, CK, h->stream));
    if (beta_shp) RC_TRY(download(beta_shp, h->beta_shp, GK, h->stream));
    if (beta_rte) RC_TRY(download(beta_rte, h->beta_rte, GK, h->stream));
    if (xi_shp) RC_TRY(download(xi_shp, h->xi_shp, h->C, h->stream));
    if (xi_rte) RC_TRY(download(xi_rte, h->xi_rte, h->C, h->stream));
    if (eta_shp) RC_TRY(download(eta_shp, h->eta_shp, h->G, h->stream));
    if (eta_rte) RC_TRY(download(eta_rte, h->eta_rte, h->G, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_overflow(h);
}

int schpf_copy_gene_state(schpf_engine_t *dst, schpf_engine_t *src)
{
    RC_TRY(check_handle(src));
    RC_TRY(check_handle(dst));
    if (dst == src) return SCHPF_OK;
    if (dst->device != src->device || dst->G != src->G || dst->K != src->K) {
        set_error("schpf_copy_gene_state: handles differ (device %d/%d, ngenes %lld/%lld, nfactors %d/%d)",
                  dst->device, src->device, (long long)dst->G, (long long)src->G, dst->K, src->K);
        return SCHPF_ERR_ARG;
    }
    if (!src->have_state) {
        set_error("schpf_copy_gene_state: source engine has no state");
        return SCHPF_ERR_STATE;
    }
    // the two handles may use different streams: finish the source's work first
    CUDA_TRY(cudaStreamSynchronize(src->stream));
    const size_t GK = sizeof(double) * (size_t)dst->G * dst->K, Gb = sizeof(double) * (size_t)dst->G;
    CUDA_TRY(cudaMemcpyAsync(dst->beta_shp, src->beta_shp, GK, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->beta_rte, src->beta_rte, GK, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->eta_shp, src->eta_shp, Gb, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->eta_rte, src->eta_rte, Gb, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaStreamSynchronize(dst->stream));   // src may be stepped again right away
    dst->tables_b_valid = false;
    return SCHPF_OK;
}

int schpf_copy_cell_state(schpf_engine_t *dst, int64_t dst_row0, schpf_engine_t *src, int64_t src_row0,
                          int64_t nrows)
{
    RC_TRY(check_handle(src));
    RC_TRY(check_handle(dst));
    if (dst->device != src->device || dst->K != src->K) {
        set_error("schpf_copy_cell_state: handles differ (device %d/%d, nfactors %d/%d)", dst->device,
                  src->device, dst->K, src->K);
        return SCHPF_ERR_ARG;
    }
    if (nrows < 0 || dst_row0 < 0 || src_row0 < 0 || dst_row0 + nrows > dst->C || src_row0 + nrows > src->C) {
        set_error("schpf_copy_cell_state: rows [%lld, +%lld) of %lld -> [%lld, +%lld) of %lld out of range",
                  (long long)src_row0, (long long)nrows, (long long)src->C, (long long)dst_row0,
                  (long long)nrows, (long long)dst->C);
        return SCHPF_ERR_ARG;
    }
    if (dst == src && dst_row0 == src_row0) return SCHPF_OK;
    if (dst == src && dst_row0 < src_row0 + nrows && src_row0 < dst_row0 + nrows) {
        set_error("schpf_copy_cell_state: overlapping ranges within one handle");
        return SCHPF_ERR_ARG;
    }
    if (!src->have_state) {
        set_error("schpf_copy_cell_state: source engine has no state");
        return SCHPF_ERR_STATE;
    }
    if (nrows == 0) return SCHPF_OK;
    const bool two_streams = dst->stream != src->stream;
    if (two_streams) CUDA_TRY(cudaStreamSynchronize(src->stream));     // the source's work first
    const size_t K = (size_t)dst->K;
    const size_t rowsK = sizeof(double) * (size_t)nrows * K, rows1 = sizeof(double) * (size_t)nrows;
    const size_t d0 = (size_t)dst_row0, s0 = (size_t)src_row0;
    CUDA_TRY(cudaMemcpyAsync(dst->theta_shp + d0 * K, src->theta_shp + s0 * K, rowsK, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->theta_rte + d0 * K, src->theta_rte + s0 * K, rowsK, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->xi_shp + d0, src->xi_shp + s0, rows1, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->xi_rte + d0, src->xi_rte + s0, rows1, cudaMemcpyDeviceToDevice, dst->stream));
    if (two_streams) CUDA_TRY(cudaStreamSynchronize(dst->stream));     // src may be written again right away
    dst->tables_t_valid = false;
    dst->have_state = true;
    return SCHPF_OK;
}

static int exchange_if_sharded(schpf_engine_t *h, int flags)
{
    if (!h->comm || (flags & SCHPF_FREEZE_GENES) || exchange_is_late(flags)) return SCHPF_OK;
    return allreduce_exchange(h, h->stream);
}

// One regular iteration on a sharded engine with the all-reduce hidden: the exchange buffer
// needs only the genes-own sweep, so that sweep goes first and the all-reduce then runs on a
// second stream underneath the cells-own sweep (1.6 ms of work against a latency-bound 3 MB
// message).  The two sweeps write different accumulators, so their order does not matter.
static int step_overlapped(schpf_engine_t *h, int flags)
{
    RC_TRY(ensure_tables(h));
    RC_TRY(zero_accumulators(h, true));
    if (h->opt_overlap_sweeps) {
        // both sweeps start together (cells-own on its own stream); the fold and the all-reduce follow the
        // genes-own sweep alone and hide under what is left of the cells-own one
        RC_TRY(ensure_sweep_stream(h));
        SweepArgs A;
        cudaEvent_t e1 = nullptr;
        RC_TRY(timing_begin(h, SWEEP_SHAPE, &e1));
        CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->sstream, h->ev_fork, 0));
        // genes-own FIRST: CTAs are dispatched in launch order, so it also finishes first and its fold and
        // all-reduce run while the cells-own sweep still has CTAs to go
        RC_TRY(genes_sweep(h, nullptr, true));
        RC_TRY(cells_sweep(h, &A, h->sstream));
        CUDA_TRY(cudaEventRecord(h->ev_join, h->sstream));
        RC_TRY(fold_exchange_buffer(h));
        CUDA_TRY(cudaEventRecord(h->ev_folded, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->xstream, h->ev_folded, 0));
        RC_TRY(allreduce_exchange(h, h->xstream));
        CUDA_TRY(cudaEventRecord(h->ev_reduced, h->xstream));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        if (e1) CUDA_TRY(cudaEventRecord(e1, h->stream));
        RC_TRY(launch_slow_fixup(h->stream, A, nullptr, h->overflow));
        h->n_launches += 1;
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_reduced, 0));
        return step_end_impl(h, flags);
    }
    RC_TRY(genes_sweep(h));
    RC_TRY(fold_exchange_buffer(h));
    CUDA_TRY(cudaEventRecord(h->ev_folded, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->xstream, h->ev_folded, 0));
    RC_TRY(allreduce_exchange(h, h->xstream));
    CUDA_TRY(cudaEventRecord(h->ev_reduced, h->xstream));
    RC_TRY(cells_sweep(h));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_reduced, 0));
    return step_end_impl(h, flags);
}

int schpf_step(schpf_engine_t *h, int n_iters, int flags)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    const bool overlap = h->comm && h->xstream && h->opt_overlap_exchange && h->opt_variant == 0 &&
                         !(flags & SCHPF_FREEZE_GENES) && !exchange_is_late(flags);
    for (int t = 0; t < n_iters; ++t) {
        if (overlap) {
            RC_TRY(step_overlapped(h, flags));
            continue;
        }
        const bool fold = h->comm != nullptr;      // one GPU: no exchange, beta is finalised from the accumulators
        RC_TRY(step_begin_impl(h, flags, 0, 0, fold));
        RC_TRY(exchange_if_sharded(h, flags));
        RC_TRY(step_end_impl(h, flags, fold));
    }
    return SCHPF_OK;
}

int schpf_step_with_xphi(schpf_engine_t *h, const double *xphi_host, int flags)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    if (!xphi_host && h->nnz > 0) {
        set_error("null Xphi");
        return SCHPF_ERR_ARG;
    }
    const bool freeze = flags & SCHPF_FREEZE_GENES;
    double *d_xphi = nullptr;
    RC_TRY(dev_alloc(&d_xphi, h->nnz * h->K));
    int rc = upload(d_xphi, xphi_host, h->nnz * h->K, h->stream);
    if (rc == SCHPF_OK) rc = ensure_tables(h);
    if (rc == SCHPF_OK) rc = zero_accumulators(h, !freeze);
    // hpf_numba.py:152-155 for both axes
    if (rc == SCHPF_OK) rc = launch_scatter_xphi(h->stream, h->nnz, h->K, d_xphi, h->row, h->direct_t);
    if (rc == SCHPF_OK && !freeze)
        rc = launch_scatter_xphi(h->stream, h->nnz, h->K, d_xphi, h->col, h->direct_b);
    if (rc == SCHPF_OK) rc = step_begin_impl(h, flags, 2, 0);
    if (rc == SCHPF_OK) rc = exchange_if_sharded(h, flags);
    if (rc == SCHPF_OK) rc = step_end_impl(h, flags);
    cudaStreamSynchronize(h->stream);
    pool_free(d_xphi, h->stream);
    return rc;
}

int schpf_step_random_phi(schpf_engine_t *h, uint64_t seed, int flags)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    RC_TRY(step_begin_impl(h, flags, 1, seed));
    RC_TRY(exchange_if_sharded(h, flags));
    return step_end_impl(h, flags);
}

int schpf_step_begin(schpf_engine_t *h, int flags, int mode, uint64_t seed)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    if (mode != 0 && mode != 1) {
        set_error("schpf_step_begin: mode must be 0 (E-step) or 1 (random phi)");
        return SCHPF_ERR_ARG;
    }
    return step_begin_impl(h, flags, mode, seed);
}

int schpf_exchange_buffer(schpf_engine_t *h, void **device_ptr, int64_t *n_doubles)
{
    RC_TRY(check_handle(h));
    if (device_ptr) *device_ptr = h->exch;
    if (n_doubles) *n_doubles = h->G * h->K + h->K;
    return SCHPF_OK;
}

int schpf_step_end(schpf_engine_t *h, int flags)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    return step_end_impl(h, flags);
}

int schpf_loss_parts(schpf_engine_t *h, double *sum_llh, int64_t *count)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    const int K = h->K;
    // e_x tables in the sweep layout (hpf_numba.py:33-41)
    RC_TRY(launch_ex_table(h->stream, h->C, K, h->geom_xt(), h->theta_shp, h->theta_rte, h->Xt));
    RC_TRY(launch_ex_table(h->stream, h->G, K, h->geom_xb(), h->beta_shp, h->beta_rte, h->Xb));
    SweepArgs A = side_args(h, h->cells);
    A.own_tab = h->f32 ? reinterpret_cast<const double *>(h->Xt32) : h->Xt;
    A.oth_tab = h->f32 ? reinterpret_cast<const double *>(h->Xb32) : h->Xb;
    A.partial = h->partials;
    RC_TRY(timed_sweep(h, SWEEP_LLH, h->cells, A));
    RC_TRY(launch_sum_partials(h->stream, h->partials, h->cells.nblocks * h->cells.nranges, h->scalars));
    h->n_launches += 3;
    double pair[2] = {0.0, 0.0};
    if (h->comm) {
        // [sum_i llh_i of this shard, nnz of this shard] summed over shards (2 doubles)
        RC_TRY(launch_pack_loss(h->stream, h->scalars, h->lgamma_sum, (double)h->nnz, h->scalars + 2));
        RC_TRY(nccl_check(g_nccl.AllReduce(h->scalars + 2, h->scalars + 2, 2, NCCL_FLOAT64, NCCL_SUM, h->comm,
                                           h->stream),
                          "ncclAllReduce(loss)"));
        CUDA_TRY(cudaMemcpyAsync(pair, h->scalars + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(pair, h->scalars, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        pair[0] -= h->lgamma_sum;
        pair[1] = (double)h->nnz;
    }
    if (sum_llh) *sum_llh = pair[0];
    if (count) *count = (int64_t)(pair[1] + 0.5);
    return check_overflow(h);
}

int schpf_loss(schpf_engine_t *h, double *mean_negative_llh)
{
    double s = 0.0;
    int64_t n = 0;
    RC_TRY(schpf_loss_parts(h, &s, &n));
    if (mean_negative_llh) *mean_negative_llh = n > 0 ? -s / (double)n : 0.0;
    if (n > 0 && !(s == s)) {
        set_error("non-finite log-likelihood");
        return SCHPF_ERR_NUMERIC;
    }
    return SCHPF_OK;
}

int schpf_llh_pointwise(schpf_engine_t *h, double *out_host_nnz)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    double *d = nullptr;
    RC_TRY(dev_alloc(&d, h->nnz));
    int rc = launch_llh_pointwise(h->stream, h->nnz, h->K, h->row, h->col, h->data, h->theta_shp, h->theta_rte,
                                  h->beta_shp, h->beta_rte, d);
    if (rc == SCHPF_OK) rc = download(out_host_nnz, d, h->nnz, h->stream);
    cudaStreamSynchronize(h->stream);
    pool_free(d, h->stream);
    return rc;
}

int schpf_xphi_debug(schpf_engine_t *h, double *out)
{
    RC_TRY(check_handle(h));
    RC_TRY(require_ready(h));
    RC_TRY(ensure_tables(h));
    double *d = nullptr;
    RC_TRY(dev_alloc(&d, h->nnz * h->K));
    int rc = launch_literal(h->stream, h->nnz, h->K, h->row, h->col, h->data, h->elog_t, h->elog_b, d, nullptr,
                            nullptr);
    if (rc == SCHPF_OK) rc = download(out, d, h->nnz * h->K, h->stream);
    cudaStreamSynchronize(h->stream);
    pool_free(d, h->stream);
    return rc;
}

int schpf_layout_dump(schpf_engine_t *h, int side, int64_t capacity, int32_t *own, int32_t *oth, int32_t *count,
                      int64_t *n_out)
{
    RC_TRY(check_handle(h));
    if (!h->have_coo) {
        set_error("schpf_layout_dump: no matrix");
        return SCHPF_ERR_STATE;
    }
    const SideLayout &L = side == 0 ? h->cells : h->genes;
    if (n_out) *n_out = L.padded_entries;
    if (!own || !oth || !count) return SCHPF_OK;       // size query
    if (capacity < L.padded_entries) {
        set_error("schpf_layout_dump: buffers hold %lld entries, the layout has %lld", (long long)capacity,
                  (long long)L.padded_entries);
        return SCHPF_ERR_ARG;
    }
    if (dump_side_layout(L, h->stream, own, oth, count) < 0) {
        set_error("schpf_layout_dump: device error %s", cudaGetErrorString(cudaGetLastError()));
        return SCHPF_ERR_CUDA;
    }
    return SCHPF_OK;
}

int schpf_comm_unique_id(char *id128_out)
{
    if (!id128_out) {
        set_error("null id buffer");
        return SCHPF_ERR_ARG;
    }
    RC_TRY(load_nccl());
    return nccl_check(g_nccl.GetUniqueId(id128_out), "ncclGetUniqueId");
}

int schpf_comm_create(void **comm_out, int device, const char *id128, int rank, int world_size)
{
    if (!comm_out || !id128 || world_size < 1 || rank < 0 || rank >= world_size) {
        set_error("bad communicator arguments (rank %d of %d)", rank, world_size);
        return SCHPF_ERR_ARG;
    }
    *comm_out = nullptr;
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(load_nccl());
    NcclId128 id;
    memcpy(id.b, id128, 128);
    return nccl_check(g_nccl.CommInitRank(comm_out, world_size, id, rank), "ncclCommInitRank");
}

int schpf_comm_destroy(void *comm)
{
    if (!comm) return SCHPF_OK;
    RC_TRY(load_nccl());
    return nccl_check(g_nccl.CommDestroy(comm), "ncclCommDestroy");
}

int schpf_comm_attach(schpf_engine_t *h, void *comm)
{
    RC_TRY(check_handle(h));
    if (comm) RC_TRY(load_nccl());
    if (comm && !h->xstream) {
        // highest priority: the all-reduce's few CTAs are placed as soon as a sweep CTA retires (the
        // sweep fills every SM; at default priority the collective waited for the sweep's tail)
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->xstream, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_folded, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_reduced, cudaEventDisableTiming));
    }
    h->comm = comm;
    return SCHPF_OK;
}

int schpf_synchronize(schpf_engine_t *h)
{
    RC_TRY(check_handle(h));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SCHPF_OK;
}

int schpf_counter(schpf_engine_t *h, const char *what, double *value)
{
    RC_TRY(check_handle(h));
    if (!what || !value) {
        set_error("null counter argument");
        return SCHPF_ERR_ARG;
    }
    if (!strcmp(what, "nnz")) *value = (double)h->nnz;
    else if (!strcmp(what, "padded_nnz_cells")) *value = (double)h->cells.padded_entries;
    else if (!strcmp(what, "padded_nnz_genes")) *value = (double)h->genes.padded_entries;
    else if (!strcmp(what, "sweep_launches")) *value = h->n_sweeps;
    else if (!strcmp(what, "kernel_launches")) *value = h->n_launches;
    else if (!strcmp(what, "iterations")) *value = h->n_iterations;
    else if (!strcmp(what, "layout_bytes")) *value = (double)(h->cells.bytes + h->genes.bytes);
    else if (!strcmp(what, "panel_rows")) *value = (double)h->cells.panel_rows;
    else if (!strcmp(what, "lanes")) *value = h->cells.opw == 32 ? 1.0 : 0.0;
    else if (!strcmp(what, "precision")) *value = h->f32 ? 32.0 : 64.0;
    else if (!strcmp(what, "warps_per_cta")) *value = (double)h->cells.warps;
    else if (!strcmp(what, "packed_entries")) *value = h->cells.packed ? 1.0 : 0.0;
    else if (!strcmp(what, "grid_cells")) *value = (double)h->cells.nblocks * h->cells.nranges;
    else if (!strcmp(what, "grid_genes")) *value = (double)h->genes.nblocks * h->genes.nranges;
    else if (!strcmp(what, "shape_sweep_launches")) *value = h->n_shape_sweeps;
    else if (!strcmp(what, "sweep_ms") || !strcmp(what, "sweep_ms_shape") || !strcmp(what, "sweep_ms_llh")) {
        RC_TRY(collect_timing(h));
        *value = !strcmp(what, "sweep_ms") ? h->sweep_ms_accum : !strcmp(what, "sweep_ms_shape") ? h->sweep_ms_shape
                                                                                               : h->sweep_ms_llh;
    } else if (!strcmp(what, "reset")) {
        RC_TRY(collect_timing(h));
        h->sweep_ms_accum = h->sweep_ms_shape = h->sweep_ms_llh = 0.0;
        h->n_sweeps = h->n_shape_sweeps = h->n_launches = h->n_iterations = 0;
        *value = 0.0;
    } else if (!strcmp(what, "slow_path_hits")) {
        unsigned long long v = 0;
        CUDA_TRY(cudaMemcpyAsync(&v, h->slow_hits, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        *value = (double)v;
    } else {
        set_error("unknown counter '%s'", what);
        return SCHPF_ERR_ARG;
    }
    return SCHPF_OK;
}

}  // extern "C"
