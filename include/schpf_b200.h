/*
 * schpf_b200 -- C ABI of the B200-native CAVI engine for single-cell
 * Hierarchical Poisson Factorization (drop-in for the hot path of
 * simslab/scHPF: schpf/hpf_numba.py, the loop body of schpf/scHPF_.py:_fit,
 * schpf/loss.py).
 *
 * The reference has no FFI seam of its own: its estimator calls six numba
 * functions on flat ndarrays (schpf/scHPF_.py:21 `from schpf.hpf_numba import *`,
 * schpf/loss.py:13).  This header is the seam a maintainer would bind with
 * ctypes instead.  Two levels are exported:
 *
 *   1. function level  -- one entry point per reference kernel, same argument
 *      meaning, host pointers in / host pointers out (upload -> kernel ->
 *      download).  Used by the parity tests and by callers that want a 1:1
 *      replacement of a single numba function.
 *   2. engine level    -- an opaque handle that keeps the COO matrix and the
 *      eight variational arrays resident in HBM and runs whole CAVI
 *      iterations (the replacement for the loop body of _fit), with a
 *      split-phase step for one-process-per-GPU cell sharding.
 *
 * Conventions: plain C types only; all matrices are C-contiguous row-major
 * fp64; indices and counts are int32; every function returns 0 on success or
 * a SCHPF_ERR_* code, and schpf_last_error() then describes the failure.
 * Nothing here falls back to a CPU implementation: without a CUDA device the
 * calls fail with SCHPF_ERR_CUDA.
 */
#ifndef SCHPF_B200_H
#define SCHPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCHPF_OK            0
#define SCHPF_ERR_CUDA      1   /* CUDA runtime error (no device, OOM, launch failure) */
#define SCHPF_ERR_ARG       2   /* bad argument (null pointer, K out of range, index out of range) */
#define SCHPF_ERR_STATE     3   /* call order (step before set_coo / set_state / set_hyper) */
#define SCHPF_ERR_NUMERIC   4   /* non-finite value produced */

#define SCHPF_MAX_FACTORS   64  /* largest K the sweep kernels are instantiated for */

/* flags for schpf_step* */
#define SCHPF_FREEZE_GENES  1   /* skip every beta/eta update (scHPF_.py:617,668,682,697) */
#define SCHPF_SIMULTANEOUS  2   /* beta_theta_simultaneous=True ordering (scHPF_.py:666-684) */
#define SCHPF_CELLS_FIRST   4   /* minibatch ordering (scHPF_.py:686-704, `batched`): theta/xi are updated
                                   first (from the old beta), and beta's rate then sees the NEW theta.
                                   On a sharded engine (a minibatch whose cells are spread over the ranks)
                                   the exchange therefore happens AFTER the cell update: with an attached
                                   communicator inside schpf_step / schpf_step_end; split-phase callers use */
#define SCHPF_PHASE_CELLS   8   /* schpf_step_end(CELLS_FIRST | PHASE_CELLS): theta/xi + the new theta's column
                                   sums into the exchange buffer; all-reduce the buffer; then ...            */
#define SCHPF_PHASE_GENES  16   /* ... schpf_step_end(CELLS_FIRST | PHASE_GENES): beta/eta                  */

typedef struct schpf_engine schpf_engine_t;

int         schpf_version(void);
const char *schpf_last_error(void);
/* number of visible CUDA devices, or a negative SCHPF_ERR_* */
int         schpf_device_count(void);
/* schpf_set_coo keeps one work area per device between calls (about 32 bytes per nonzero of the
 * largest matrix seen; growing device allocations by gigabytes costs more than the re-layout
 * itself).  This frees it; SCHPF_ERR_STATE if a schpf_set_coo is running on that device. */
int         schpf_release_scratch(int device);

/* ------------------------------------------------------------------------
 * Function level: one entry per reference kernel.  All pointers are HOST.
 * --------------------------------------------------------------------- */

/* hpf_numba.py:16-18 `psi` and :20-22 `cgammaln`, elementwise over n values */
int schpf_psi(int device, int64_t n, const double *x, double *out);
int schpf_gammaln(int device, int64_t n, const double *x, double *out);

/* hpf_numba.py:55-114 compute_Xphi_data -> Xphi (nnz x K) */
int schpf_compute_Xphi_data(int device, int64_t nnz, int64_t ncells, int64_t ngenes, int nfactors,
                            const int32_t *X_data, const int32_t *X_row, const int32_t *X_col,
                            const double *theta_vi_shape, const double *theta_vi_rate,
                            const double *beta_vi_shape, const double *beta_vi_rate,
                            double *Xphi_out);

/* hpf_numba.py:129-156 compute_loading_shape_update -> (nkeep x K) */
int schpf_compute_loading_shape_update(int device, int64_t nnz, int nfactors,
                                       const double *Xphi_data, const int32_t *X_keep,
                                       int64_t nkeep, double shape_prior, double *result);

/* hpf_numba.py:160-177 compute_loading_rate_update -> (n x K);
 * prior_* have n entries, other_loading_* are (m x K) */
int schpf_compute_loading_rate_update(int device, int64_t n, int64_t m, int nfactors,
                                      const double *prior_vi_shape, const double *prior_vi_rate,
                                      const double *other_loading_vi_shape,
                                      const double *other_loading_vi_rate, double *result);

/* hpf_numba.py:181-188 compute_capacity_rate_update -> (n) */
int schpf_compute_capacity_rate_update(int device, int64_t n, int nfactors,
                                       const double *loading_vi_shape,
                                       const double *loading_vi_rate, double prior_rate,
                                       double *result);

/* hpf_numba.py:25-51 compute_pois_llh -> (nnz) */
int schpf_compute_pois_llh(int device, int64_t nnz, int64_t ncells, int64_t ngenes, int nfactors,
                           const int32_t *X_data, const int32_t *X_row, const int32_t *X_col,
                           const double *theta_vi_shape, const double *theta_vi_rate,
                           const double *beta_vi_shape, const double *beta_vi_rate,
                           double *llh_out);

/* ------------------------------------------------------------------------
 * Engine level: the loop body of scHPF_.py:_fit (lines 642-715) with state
 * resident on one GPU.
 * --------------------------------------------------------------------- */

/* `stream` is a cudaStream_t (or NULL for the default stream) that all work
 * of this handle is enqueued on; it lets a host framework order its own
 * work (e.g. an NCCL all-reduce of the exchange buffer) with the engine's. */
int schpf_create(schpf_engine_t **out, int device, int64_t ncells, int64_t ngenes,
                 int nfactors, void *stream);
int schpf_destroy(schpf_engine_t *h);

/* tuning knobs, before schpf_set_coo: "panel_rows", "warps_per_cta",
 * "target_ctas", "variant" (0 = tiled two-pass sweep, 1 = literal per-nnz
 * kernel with atomics), "timing" (1 = record CUDA events around the sweeps),
 * "packed_entries" (4-byte stream entries pad<<31 | count<<12 | row when every count is < 2^19: half the
 * entry stream and resident layout; -1 default = where it is not slower (one-lane K <= 16 and the fp32
 * sweep), 0 = never, 1 = wherever the stream allows it), "overlap_exchange" (default 1: with an attached
 * communicator schpf_step runs the all-reduce on a second stream underneath the cells-own sweep;
 * 0 = in order on the engine's stream), "overlap_sweeps" (default 1: the two shape sweeps of an iteration start
 * together, the cells-own one on a second stream, so that one grid's last partial wave is filled by the other's CTAs), "lanes" (default 1: K <= 20 and 29..32 run the
 * one-lane-per-owner sweep; 0 = lane-pair sweep for every K), "rank_per_range" (-1 automatic),
 * "precision" (64 default; 32 = the fp32 sweep used for dtype=np.float32 models, scHPF_.py:225-246:
 * table entries, dot product, quotient and per-panel sums in fp32, everything else fp64) */
int schpf_set_option(schpf_engine_t *h, const char *key, int64_t value);

/* The sparse count matrix as COO triples (X.row, X.col, X.data of a
 * scipy coo_matrix; any order, duplicates allowed -- each triple is one
 * "nonzero" exactly as in the reference).  Host or device pointers. */
int schpf_set_coo(schpf_engine_t *h, const int32_t *row, const int32_t *col,
                  const int32_t *data, int64_t nnz);
int schpf_set_coo_device(schpf_engine_t *h, const int32_t *d_row, const int32_t *d_col,
                         const int32_t *d_data, int64_t nnz);

/* a, a', b', c, c', d' (scHPF_.py:225-246, :847-879) */
int schpf_set_hyper(schpf_engine_t *h, double a, double ap, double bp,
                    double c, double cp, double dp);

/* theta/beta: (n x K); xi/eta: (n).  Host pointers.  A NULL pair leaves that
 * distribution untouched (set) / is skipped (get). */
int schpf_set_state(schpf_engine_t *h,
                    const double *theta_shp, const double *theta_rte,
                    const double *beta_shp, const double *beta_rte,
                    const double *xi_shp, const double *xi_rte,
                    const double *eta_shp, const double *eta_rte);
int schpf_get_state(schpf_engine_t *h,
                    double *theta_shp, double *theta_rte,
                    double *beta_shp, double *beta_rte,
                    double *xi_shp, double *xi_rte,
                    double *eta_shp, double *eta_rte);

/* Copy beta and eta (shape and rate) from `src` to `dst`, device to device.  Both handles must be on
 * the same device with the same ngenes / nfactors.  This is how the minibatch loop (scHPF_.py:642-650:
 * a different row subset of X every iteration) hands the gene side from one batch's engine to the next. */
int schpf_copy_gene_state(schpf_engine_t *dst, schpf_engine_t *src);
/* Copy theta (shape, rate rows) and xi (shape, rate) of `nrows` cells, src rows [src_row0, +nrows) ->
 * dst rows [dst_row0, +nrows), device to device.  Same device and nfactors.  With the cells of a
 * matrix permuted once by the minibatch shuffle (util.py:220-221) every batch is a contiguous row
 * range, and this moves a batch between the engine that holds all cells and the batch's engine. */
int schpf_copy_cell_state(schpf_engine_t *dst, int64_t dst_row0, schpf_engine_t *src, int64_t src_row0,
                          int64_t nrows);

/* n full CAVI iterations (Xphi -> beta -> eta -> theta -> xi), single GPU. */
int schpf_step(schpf_engine_t *h, int n_iters, int flags);

/* t == 0 branch of _fit (scHPF_.py:652-655): the caller supplies Xphi
 * (nnz x K, host, in the order the triples were given to schpf_set_coo) ... */
int schpf_step_with_xphi(schpf_engine_t *h, const double *xphi_host, int flags);
/* ... or the engine draws y * Dirichlet(1_K) per nonzero on the device
 * (counter-based generator keyed by (seed, row, col); not numpy's stream). */
int schpf_step_random_phi(schpf_engine_t *h, uint64_t seed, int flags);

/* Split-phase iteration for cell sharding (one process per GPU):
 *   begin  : both sweeps; writes this shard's partial sums into the exchange
 *            buffer  [ G*K beta-shape partials | K column sums of theta.e_x ]
 *   (host framework all-reduces the buffer in place, sum, fp64)
 *   end    : beta/eta then theta/xi finalisation from the reduced buffer.
 * `mode`: 0 = regular E-step, 1 = random phi (uses `seed`).               */
int schpf_step_begin(schpf_engine_t *h, int flags, int mode, uint64_t seed);
int schpf_exchange_buffer(schpf_engine_t *h, void **device_ptr, int64_t *n_doubles);
int schpf_step_end(schpf_engine_t *h, int flags);

/* Optional: let the engine run the exchange itself.  A communicator (one per process and
 * group of ranks; creating one takes seconds, so the host framework keeps it and shares it
 * between successive engines) is created from the 128-byte ncclUniqueId that
 * schpf_comm_unique_id produced on rank 0 and the host framework distributed.  After
 * schpf_comm_attach every schpf_step* call of that handle performs the all-reduce
 * (ncclAllReduce, sum, fp64, in place on the exchange buffer) between its two phases -- for
 * schpf_step on a second stream, ordered with events, so that it hides under the cells-own sweep;
 * for the t == 0 variants in order on the engine's own stream -- and schpf_loss* returns the loss
 * over all shards.  NCCL is
 * loaded with dlopen("libnccl.so.2") at first use.  Collective calls: every rank must make
 * the same sequence of create / step / loss calls.  The engine does not own the communicator. */
int schpf_comm_unique_id(char *id128_out);
int schpf_comm_create(void **comm_out, int device, const char *id128, int rank, int world_size);
int schpf_comm_destroy(void *comm);
int schpf_comm_attach(schpf_engine_t *h, void *comm);    /* comm == NULL detaches */

/* loss.py:142-168 mean_negative_pois_llh of the resident matrix under the
 * resident state.  `sum_llh` receives sum_i llh_i and `count` the number of
 * nonzeros, so that shards can be combined: loss = -sum(sum_llh)/sum(count). */
int schpf_loss(schpf_engine_t *h, double *mean_negative_llh);
int schpf_loss_parts(schpf_engine_t *h, double *sum_llh, int64_t *count);
/* loss.py:107-139 pois_llh_pointwise, in the order given to schpf_set_coo */
int schpf_llh_pointwise(schpf_engine_t *h, double *out_host_nnz);
/* Xphi of the resident state (nnz x K, host) -- small inputs, debugging */
int schpf_xphi_debug(schpf_engine_t *h, double *out_host_nnz_x_K);

/* The device layout of one sweep direction (side 0: cells own, 1: genes own) decoded back to COO
 * triples, in stream order: `own` / `oth` are the owner's and the other axis' GLOBAL indices, pad
 * entries have own = oth = -1 and count 0.  *n_out = entries in the stream (pads included); call
 * with null buffers to query it.  Test hook for the exact integer round trip
 * multiset{(row, col, y)} in == out; small matrices. */
int schpf_layout_dump(schpf_engine_t *h, int side, int64_t capacity, int32_t *own, int32_t *oth,
                      int32_t *count, int64_t *n_out);

/* ---- device-side ingest of the text formats the reference loads before a fit ----------------------
 * Replaces `scipy.io.mmread` (bin/scHPF:373-374; files written by `scHPF prep`, bin/scHPF:327,361) and
 * `load_coo` (schpf/preprocessing.py:11-29: np.loadtxt of "row<TAB>col<TAB>count" lines) for the
 * data lines; the host reads the few header lines and passes the offset of the first data byte.
 * `d_text` is the file's bytes in device memory.  Data lines = non-empty lines not starting with '%'.
 * Output order = file order (what mmread / loadtxt return).
 *   schpf_count_lines   how many data lines d_text[begin, nbytes) holds.
 *   schpf_parse_triples nfields 3: "row col count" (a count may be written as a real with zero
 *                       fraction, e.g. 3.000e+00), 2: "row col" with count 1 (MatrixMarket `pattern`);
 *                       index_base 1 (MatrixMarket) or 0 (the tsv); outputs are device int32 arrays of
 *                       `capacity` elements; *n_out = lines parsed.  A malformed line (non-integer or
 *                       negative field, index below the base, value >= 2^31, extra fields) returns
 *                       SCHPF_ERR_ARG with its byte offset in *err_offset. */
int schpf_count_lines(int device, void *stream, const char *d_text, int64_t nbytes, int64_t begin, int64_t *n_lines);
int schpf_parse_triples(int device, void *stream, const char *d_text, int64_t nbytes, int64_t begin, int nfields,
                        int index_base, int32_t *d_row, int32_t *d_col, int32_t *d_val, int64_t capacity,
                        int64_t *n_out, int64_t *err_offset);

/* wait for all enqueued work of this handle */
int schpf_synchronize(schpf_engine_t *h);

/* counters: what = "nnz", "padded_nnz_cells", "padded_nnz_genes",
 * "sweep_launches", "shape_sweep_launches", "kernel_launches", "sweep_ms" / "sweep_ms_shape" /
 * "sweep_ms_llh" (need option timing=1; synchronise), "iterations", "slow_path_hits",
 * "layout_bytes", "panel_rows", "warps_per_cta", "lanes" (1 = one-lane-per-owner sweep),
 * "precision" (32 = fp32 sweep, option "precision" before schpf_set_coo; 64 otherwise) */
int schpf_counter(schpf_engine_t *h, const char *what, double *value);

#ifdef __cplusplus
}
#endif
#endif /* SCHPF_B200_H */
