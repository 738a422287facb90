// K11 — logistic regression (pds.logistic_reg): the per-row pass of Newton / IRLS, and the probability pass.
//
// Reference: faer_logistic_reg (/root/reference/src/linear/logistic/logistic_solver.rs:107-146) minimises
//     cost(w) = mean_i stable_log_loss(y_i, x_i.w) + (l2 / 2) |w_features|^2           (:42-74)
// with argmin's L-BFGS (OWL-QN when l1 > 0) from a seeded random start; the expression bodies are pl_logistic_coeffs /
// pl_logistic_pred (src/num_ext/logistic_regression.rs:10-99).  The minimiser of that convex cost does not depend on the
// optimiser, so this path runs Newton's method in its IRLS form, which is built from the kernels of the linear path:
// at the current w every row gets
//     eta = x.w,  mu = sigmoid(eta),  weight = mu (1 - mu),  working response z = eta + (y - mu) / weight
// (this kernel, one coalesced pass that also returns the summed log-loss), then the WEIGHTED moments
// [X | z | 1]' diag(weight) [X | z | 1] (K2a) and the ridge solve on them (K3) give the next w; the gradient norm the
// reference stops on comes out of the same moments:  X' (mu - y) = (X' W X) w - X' W z.
// HBM-bound: (p + 1) s read + 2 s written per row and iteration here, (p + 2) s read by the moments pass.
#include "../common.h"
#include "kernels.h"

namespace pdsb {

namespace {

constexpr int IRLS_THREADS = 256;
constexpr double IRLS_MIN_WEIGHT = 1e-12;     // keeps z finite when mu saturates (the row then carries no information)

// logistic_solver.rs:10-18
__device__ __forceinline__ double stable_sigmoid(double x) {
  const double r = 1.0 / (1.0 + exp(-fabs(x)));
  return x >= 0.0 ? r : 1.0 - r;
}
// logistic_solver.rs:22-25
__device__ __forceinline__ double stable_log_loss(double y, double z) { return fmax(z, 0.0) - y * z + log1p(exp(-fabs(z))); }

template <typename T>
__global__ void __launch_bounds__(IRLS_THREADS)
irls_rows_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, const T* __restrict__ mask, int64_t n, int p,
                 int add_bias, const double* __restrict__ beta, T* __restrict__ w_out, T* __restrict__ z_out,
                 double* __restrict__ loss_parts /* [gridDim.x] */) {
  __shared__ double sb[65];
  __shared__ double red[IRLS_THREADS / 32];
  for (int i = threadIdx.x; i < p + add_bias; i += blockDim.x) sb[i] = beta[i];
  __syncthreads();
  double loss = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double eta = add_bias ? sb[p] : 0.0;
    for (int c = 0; c < p; ++c) eta = fma((double)X[(int64_t)c * ldx + r], sb[c], eta);
    const bool use = !mask || mask[r] != T(0);
    const double yv = (double)y[r];
    const double mu = stable_sigmoid(eta);
    const double wt = fmax(mu * (1.0 - mu), IRLS_MIN_WEIGHT);
    w_out[r] = use ? (T)wt : T(0);
    z_out[r] = use ? (T)(eta + (yv - mu) / wt) : T(0);
    if (use) loss += stable_log_loss(yv, eta);
  }
  // fixed-shape reduction -> one partial per block (summed in order by the host: reproducible)
  for (int off = 16; off; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int i = 0; i < IRLS_THREADS / 32; ++i) v += red[i];
    loss_parts[blockIdx.x] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(IRLS_THREADS)
sigmoid_predict_kernel(const T* __restrict__ X, int64_t ldx, int64_t n, int p, int add_bias, const double* __restrict__ beta,
                       T* __restrict__ out) {
  __shared__ double sb[65];
  for (int i = threadIdx.x; i < p + add_bias; i += blockDim.x) sb[i] = beta[i];
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double eta = add_bias ? sb[p] : 0.0;
    for (int c = 0; c < p; ++c) eta = fma((double)X[(int64_t)c * ldx + r], sb[c], eta);
    out[r] = (T)stable_sigmoid(eta);
  }
}

inline int irls_grid(int64_t n) {
  const int64_t g = ceil_div(n > 0 ? n : 1, (int64_t)IRLS_THREADS);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < cap ? g : cap);
}

}  // namespace

int irls_max_parts() { return sm_count() * 8; }

template <typename T>
int irls_rows(const T* X, int64_t ldx, const T* y, const T* mask, int64_t n, int p, int add_bias, const double* beta,
              T* w, T* z, double* loss_parts, int* n_parts, cudaStream_t s) {
  if (p + add_bias > 65) { set_error("logistic_reg: more than 64 features are not supported"); return 1; }
  const int grid = irls_grid(n);
  irls_rows_kernel<T><<<grid, IRLS_THREADS, 0, s>>>(X, ldx, y, mask, n, p, add_bias, beta, w, z, loss_parts);
  PDSB_LAUNCH_OK();
  count_launch();
  *n_parts = grid;
  return 0;
}

template <typename T>
int sigmoid_predict(const T* X, int64_t ldx, int64_t n, int p, int add_bias, const double* beta, T* out, cudaStream_t s) {
  if (p + add_bias > 65) { set_error("logistic_reg: more than 64 features are not supported"); return 1; }
  sigmoid_predict_kernel<T><<<irls_grid(n), IRLS_THREADS, 0, s>>>(X, ldx, n, p, add_bias, beta, out);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template int irls_rows<double>(const double*, int64_t, const double*, const double*, int64_t, int, int, const double*, double*,
                               double*, double*, int*, cudaStream_t);
template int sigmoid_predict<double>(const double*, int64_t, int64_t, int, int, const double*, double*, cudaStream_t);

}  // namespace pdsb
