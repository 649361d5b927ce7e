// Function-level C ABI: one entry point per reference numba kernel
// (schpf/hpf_numba.py), host pointers in / host pointers out.  These are the
// 1:1 replacements a maintainer can bind in place of a single numba function
// and what the parity tests call; the engine (engine.cu) is the fast path.
#include <vector>

#include "common.cuh"

using namespace schpf;

namespace {

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch()
    {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    int alloc(T **p, int64_t n)
    {
        void *q = nullptr;
        CUDA_TRY(cudaMalloc(&q, sizeof(T) * (size_t)(n > 0 ? n : 1)));
        ptrs.push_back(q);
        *p = reinterpret_cast<T *>(q);
        return SCHPF_OK;
    }
    template <typename T>
    int put(T **p, const T *host, int64_t n)
    {
        RC_TRY(alloc(p, n));
        if (n > 0) CUDA_TRY(cudaMemcpy(*p, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
        return SCHPF_OK;
    }
};

int get(double *host, const double *dev, int64_t n)
{
    if (n > 0) CUDA_TRY(cudaMemcpy(host, dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return SCHPF_OK;
}

int check_k(int K)
{
    if (K <= 0 || K > SCHPF_MAX_FACTORS) {
        set_error("nfactors must be in [1, %d], got %d", SCHPF_MAX_FACTORS, K);
        return SCHPF_ERR_ARG;
    }
    return SCHPF_OK;
}

}  // namespace

extern "C" {

int schpf_psi(int device, int64_t n, const double *x, double *out)
{
    CUDA_TRY(cudaSetDevice(device));
    Scratch S;
    double *dx, *dy;
    RC_TRY(S.put(&dx, x, n));
    RC_TRY(S.alloc(&dy, n));
    RC_TRY(launch_psi(nullptr, n, dx, dy, 0));
    return get(out, dy, n);
}

int schpf_gammaln(int device, int64_t n, const double *x, double *out)
{
    CUDA_TRY(cudaSetDevice(device));
    Scratch S;
    double *dx, *dy;
    RC_TRY(S.put(&dx, x, n));
    RC_TRY(S.alloc(&dy, n));
    RC_TRY(launch_psi(nullptr, n, dx, dy, 1));
    return get(out, dy, n);
}

int schpf_compute_Xphi_data(int device, int64_t nnz, int64_t ncells, int64_t ngenes, int nfactors,
                            const int32_t *X_data, const int32_t *X_row, const int32_t *X_col,
                            const double *theta_vi_shape, const double *theta_vi_rate,
                            const double *beta_vi_shape, const double *beta_vi_rate, double *Xphi_out)
{
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(check_k(nfactors));
    const int K = nfactors, ST = stride_of_kp(kp_of(K));
    Scratch S;
    int32_t *d_data, *d_row, *d_col;
    double *ts, *tr, *bs, *br, *elt, *elb, *Et, *Eb, *xphi;
    int *flag;
    RC_TRY(S.put(&d_data, X_data, nnz));
    RC_TRY(S.put(&d_row, X_row, nnz));
    RC_TRY(S.put(&d_col, X_col, nnz));
    RC_TRY(S.alloc(&flag, 1));
    CUDA_TRY(cudaMemset(flag, 0, sizeof(int)));
    RC_TRY(launch_validate_coo(nullptr, nnz, d_row, d_col, d_data, ncells, ngenes, flag));
    int f = 0;
    CUDA_TRY(cudaMemcpy(&f, flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (f & 3) {
        set_error("COO index out of range");
        return SCHPF_ERR_ARG;
    }
    RC_TRY(S.put(&ts, theta_vi_shape, ncells * K));
    RC_TRY(S.put(&tr, theta_vi_rate, ncells * K));
    RC_TRY(S.put(&bs, beta_vi_shape, ngenes * K));
    RC_TRY(S.put(&br, beta_vi_rate, ngenes * K));
    RC_TRY(S.alloc(&elt, ncells * K));
    RC_TRY(S.alloc(&elb, ngenes * K));
    RC_TRY(S.alloc(&Et, ncells * ST));
    RC_TRY(S.alloc(&Eb, ngenes * ST));
    RC_TRY(S.alloc(&xphi, nnz * K));
    const TabGeom tg = {ST, 0, 0};
    RC_TRY(launch_prep_side(nullptr, ncells, K, tg, ts, tr, elt, Et, nullptr));
    RC_TRY(launch_prep_side(nullptr, ngenes, K, tg, bs, br, elb, Eb, nullptr));
    RC_TRY(launch_literal(nullptr, nnz, K, d_row, d_col, d_data, elt, elb, xphi, nullptr, nullptr));
    return get(Xphi_out, xphi, nnz * K);
}

int schpf_compute_loading_shape_update(int device, int64_t nnz, int nfactors, const double *Xphi_data,
                                       const int32_t *X_keep, int64_t nkeep, double shape_prior,
                                       double *result)
{
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(check_k(nfactors));
    for (int64_t i = 0; i < nnz; ++i)
        if (X_keep[i] < 0 || X_keep[i] >= nkeep) {
            set_error("X_keep[%lld]=%d out of range [0,%lld)", (long long)i, X_keep[i], (long long)nkeep);
            return SCHPF_ERR_ARG;
        }
    Scratch S;
    double *xphi, *out;
    int32_t *keep;
    RC_TRY(S.put(&xphi, Xphi_data, nnz * nfactors));
    RC_TRY(S.put(&keep, X_keep, nnz));
    RC_TRY(S.alloc(&out, nkeep * nfactors));
    RC_TRY(launch_fill(nullptr, out, nkeep * nfactors, shape_prior));
    RC_TRY(launch_scatter_xphi(nullptr, nnz, nfactors, xphi, keep, out));
    return get(result, out, nkeep * nfactors);
}

int schpf_compute_loading_rate_update(int device, int64_t n, int64_t m, int nfactors,
                                      const double *prior_vi_shape, const double *prior_vi_rate,
                                      const double *other_loading_vi_shape,
                                      const double *other_loading_vi_rate, double *result)
{
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(check_k(nfactors));
    Scratch S;
    double *ps, *pr, *os, *orr, *colsum, *out;
    RC_TRY(S.put(&ps, prior_vi_shape, n));
    RC_TRY(S.put(&pr, prior_vi_rate, n));
    RC_TRY(S.put(&os, other_loading_vi_shape, m * nfactors));
    RC_TRY(S.put(&orr, other_loading_vi_rate, m * nfactors));
    RC_TRY(S.alloc(&colsum, nfactors));
    RC_TRY(S.alloc(&out, n * nfactors));
    CUDA_TRY(cudaMemset(colsum, 0, sizeof(double) * nfactors));
    RC_TRY(launch_colsum_ex(nullptr, m, nfactors, os, orr, colsum));
    RC_TRY(launch_rate_update(nullptr, n, nfactors, ps, pr, colsum, out));
    return get(result, out, n * nfactors);
}

int schpf_compute_capacity_rate_update(int device, int64_t n, int nfactors, const double *loading_vi_shape,
                                       const double *loading_vi_rate, double prior_rate, double *result)
{
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(check_k(nfactors));
    Scratch S;
    double *ls, *lr, *out;
    RC_TRY(S.put(&ls, loading_vi_shape, n * nfactors));
    RC_TRY(S.put(&lr, loading_vi_rate, n * nfactors));
    RC_TRY(S.alloc(&out, n));
    RC_TRY(launch_capacity_rate(nullptr, n, nfactors, ls, lr, prior_rate, out));
    return get(result, out, n);
}

int schpf_compute_pois_llh(int device, int64_t nnz, int64_t ncells, int64_t ngenes, int nfactors,
                           const int32_t *X_data, const int32_t *X_row, const int32_t *X_col,
                           const double *theta_vi_shape, const double *theta_vi_rate,
                           const double *beta_vi_shape, const double *beta_vi_rate, double *llh_out)
{
    CUDA_TRY(cudaSetDevice(device));
    RC_TRY(check_k(nfactors));
    const int K = nfactors;
    Scratch S;
    int32_t *d_data, *d_row, *d_col;
    double *ts, *tr, *bs, *br, *out;
    int *flag;
    RC_TRY(S.put(&d_data, X_data, nnz));
    RC_TRY(S.put(&d_row, X_row, nnz));
    RC_TRY(S.put(&d_col, X_col, nnz));
    RC_TRY(S.alloc(&flag, 1));
    CUDA_TRY(cudaMemset(flag, 0, sizeof(int)));
    RC_TRY(launch_validate_coo(nullptr, nnz, d_row, d_col, d_data, ncells, ngenes, flag));
    int f = 0;
    CUDA_TRY(cudaMemcpy(&f, flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (f & 3) {
        set_error("COO index out of range");
        return SCHPF_ERR_ARG;
    }
    RC_TRY(S.put(&ts, theta_vi_shape, ncells * K));
    RC_TRY(S.put(&tr, theta_vi_rate, ncells * K));
    RC_TRY(S.put(&bs, beta_vi_shape, ngenes * K));
    RC_TRY(S.put(&br, beta_vi_rate, ngenes * K));
    RC_TRY(S.alloc(&out, nnz));
    RC_TRY(launch_llh_pointwise(nullptr, nnz, K, d_row, d_col, d_data, ts, tr, bs, br, out));
    return get(llh_out, out, nnz);
}

}  // extern "C"
