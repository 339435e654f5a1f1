// a14 + f4: the log-likelihood functions of models/loss.py:10-45 and the self-critical (SCST) loss tail of
// OSIE/train.py:242-258, forward and analytic backward.
//
//   LogAction(p, mask)[n]              = sum_t log(p[n,t] + eps) mask[n,t] / sum(mask)            loss.py:34-37
//   LogDuration(x, mu, s2, mask)[n]    = sum_t logpdf(x; mu, s2)[n,t] mask[n,t] / sum(mask)       loss.py:39-45
//   CrossEntropyLoss(z, gt, mask)      = -sum gt log(softmax(z) + eps) mask / sum(mask)           loss.py:10-14
//   MLPLogNormalDistribution           = -sum_{mask == 1} logpdf(gt; mu, s2) / sum(mask)          loss.py:27-32
//   logpdf(x; mu, s2) = log(1/(x+eps) * 1/sqrt(2 pi s2)) - (log(x+eps) - mu)^2 / (2 s2)           (s2 is the variance)
// Quirk kept: the row sums of LogAction / LogDuration are divided by the WHOLE batch's mask.sum().
//
// SCST tail (one launch pair instead of ~40 eager ops per trial): for the K sampled trials of N images
//   accepted trial : every image kept at least one scored pair (train.py:237-238 rejects a trial whose
//                    pairs_eval table has a NaN); the first k_use accepted trials are used
//   advantage[k,n] = reward[k,n] - mean over the used trials of reward[.,n]                     train.py:252-254
//   loss           = sum_{k,n} (-LogAction[k,n] - LogDuration[k,n]) advantage[k,n]               train.py:256-258
// and the backward pass writes d loss / d (all_actions_prob, log_normal_mu, log_normal_sigma2) directly: the
// gradient of the selected-probability gather (sampling.py:24) is a scatter into [N,T,A], done here by one thread
// per (image, step) looping over the trials in order -- deterministic, no atomics.
// All of it is tiny (N*T*K elements): latency-bound, so the point is the launch count, not bandwidth.
#include <math.h>

#include "common.cuh"

namespace spb {

constexpr float kEps = 1e-7f;            // loss.py:8
constexpr float kTwoPi = 6.283185307179586f;

__device__ __forceinline__ float lognormal_logpdf(float x, float mu, float s2) {
    const float lx = logf(x + kEps);
    return logf(1.0f / (x + kEps) * 1.0f / sqrtf(kTwoPi * s2)) + (-(lx - mu) * (lx - mu) / (2.0f * s2));
}
// d logpdf / d (mu, s2, x)
__device__ __forceinline__ void lognormal_grad(float x, float mu, float s2, float &dmu, float &ds2, float &dx) {
    const float lx = logf(x + kEps), d = lx - mu;
    dmu = d / s2;
    ds2 = -0.5f / s2 + d * d / (2.0f * s2 * s2);
    dx = -1.0f / (x + kEps) - d / (s2 * (x + kEps));
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
    return r;
}

// ---------------------------------------------------------------------------
// Masked row log-likelihoods.  One block per trial k (or one block in total for the plain functions):
// phase 1 the trial's mask sums, phase 2 one warp per row.
//   p      [K,N,T] selected probabilities, or NULL: gathered from probs[N,T,A] at actions[K,N,T]
//   x      [K,N,T] durations; mu, s2 [N,T] shared by the trials
//   out_a / out_d [K,N] = LogAction / LogDuration (sign as in loss.py); msum [K,2] = (sum action mask, sum duration mask)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loglik_rows_kernel(const float *__restrict__ p, const float *__restrict__ probs, const int32_t *__restrict__ actions,
                   const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ s2,
                   const float *__restrict__ am, const float *__restrict__ dm, int N, int T, int A,
                   float *__restrict__ out_a, float *__restrict__ out_d, float *__restrict__ msum) {
    __shared__ double sh[8];
    const int k = blockIdx.x;
    const int64_t base = (int64_t)k * N * T;
    double sa = 0.0, sd = 0.0;
    for (int i = threadIdx.x; i < N * T; i += blockDim.x) {
        if (am) sa += am[base + i];
        if (dm) sd += dm[base + i];
    }
    sa = block_sum(sa, sh);
    sd = block_sum(sd, sh);
    if (threadIdx.x == 0 && msum) { msum[2 * k] = (float)sa; msum[2 * k + 1] = (float)sd; }
    const float ma = (float)sa, md = (float)sd;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int n = w; n < N; n += nw) {
        float la = 0.0f, ld = 0.0f;
        for (int t = lane; t < T; t += 32) {
            const int64_t i = base + (int64_t)n * T + t;
            if (am && out_a) {
                const float pv = p ? p[i] : probs[((int64_t)n * T + t) * A + actions[i]];
                la += logf(pv + kEps) * am[i];
            }
            if (dm && out_d) ld += lognormal_logpdf(x[i], mu[n * T + t], s2[n * T + t]) * dm[i];
        }
        for (int o = 16; o > 0; o >>= 1) {
            la += __shfl_xor_sync(0xffffffffu, la, o);
            ld += __shfl_xor_sync(0xffffffffu, ld, o);
        }
        if (lane == 0) {
            if (out_a) out_a[(int64_t)k * N + n] = la / ma;
            if (out_d) out_d[(int64_t)k * N + n] = ld / md;
        }
    }
}

// backward of the row log-likelihoods for given upstream row gradients ga, gd [K,N]:
//   grad_p [K,N,T]; grad_mu, grad_s2 [N,T] summed over the trials in order; grad_x [K,N,T]
__global__ void __launch_bounds__(256)
loglik_rows_backward_kernel(const float *__restrict__ p, const float *__restrict__ x, const float *__restrict__ mu,
                            const float *__restrict__ s2, const float *__restrict__ am, const float *__restrict__ dm,
                            const float *__restrict__ msum, const float *__restrict__ ga, const float *__restrict__ gd,
                            int K, int N, int T, float *__restrict__ grad_p, float *__restrict__ grad_mu,
                            float *__restrict__ grad_s2, float *__restrict__ grad_x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (n, t)
    if (i >= N * T) return;
    const int n = i / T;
    float gmu = 0.0f, gs2 = 0.0f;
    for (int k = 0; k < K; ++k) {
        const int64_t j = (int64_t)k * N * T + i;
        if (grad_p) grad_p[j] = ga[(int64_t)k * N + n] * am[j] / (msum[2 * k] * (p[j] + kEps));
        if (gd) {
            float dmu, ds2, dx;
            lognormal_grad(x[j], mu[i], s2[i], dmu, ds2, dx);
            const float c = gd[(int64_t)k * N + n] * dm[j] / msum[2 * k + 1];
            gmu += c * dmu; gs2 += c * ds2;
            if (grad_x) grad_x[j] = c * dx;
        }
    }
    if (grad_mu) grad_mu[i] = gmu;
    if (grad_s2) grad_s2[i] = gs2;
}

// ---------------------------------------------------------------------------
// CrossEntropyLoss (loss.py:10-14): one block per (n, t) row of logits [R, A]; gt dense [R, A].
//   row_loss[r] = -mask[r] sum_a gt[a] log(softmax(z)[a] + eps);  with grad_z != NULL also the gradient
//   d(sum_r row_loss)/dz scaled by *scale_num / *scale_den (= upstream gradient / mask.sum()).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cross_entropy_rows_kernel(const float *__restrict__ z, const float *__restrict__ gt, const float *__restrict__ mask,
                          int A, float *__restrict__ row_loss, float *__restrict__ grad_z,
                          const float *__restrict__ scale_num, const float *__restrict__ scale_den) {
    __shared__ double sh[8];
    __shared__ float shf[8];
    const int64_t r = blockIdx.x;
    const float *zr = z + r * A, *gr = gt + r * A;
    float mx = -INFINITY;
    for (int a = threadIdx.x; a < A; a += blockDim.x) mx = fmaxf(mx, zr[a]);
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) shf[w] = mx;
        __syncthreads();
        mx = shf[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, shf[i]);
    }
    double se = 0.0;
    for (int a = threadIdx.x; a < A; a += blockDim.x) se += expf(zr[a] - mx);
    const float inv = (float)(1.0 / block_sum(se, sh));
    double acc = 0.0, wsum = 0.0;                       // sum gt log(p+eps),  sum gt p/(p+eps)
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const float pa = expf(zr[a] - mx) * inv, g = gr[a];
        if (g != 0.0f) { acc += g * logf(pa + kEps); wsum += g * pa / (pa + kEps); }
    }
    acc = block_sum(acc, sh);
    const float m = mask[r];
    if (row_loss && threadIdx.x == 0) row_loss[r] = (float)(-acc * m);
    if (grad_z) {
        wsum = block_sum(wsum, sh);
        const float sc = -m * scale_num[0] / scale_den[0];
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
            const float pa = expf(zr[a] - mx) * inv;
            grad_z[r * A + a] = sc * (gr[a] * pa / (pa + kEps) - pa * (float)wsum);
        }
    }
}

// MLPLogNormalDistribution (loss.py:27-32) elements: item[i] = -logpdf where mask == 1 else 0; backward likewise.
__global__ void __launch_bounds__(256)
lognormal_nll_kernel(const float *__restrict__ mu, const float *__restrict__ s2, const float *__restrict__ gt,
                     const float *__restrict__ mask, int64_t n, float *__restrict__ item, float *__restrict__ grad_mu,
                     float *__restrict__ grad_s2, const float *__restrict__ scale_num,
                     const float *__restrict__ scale_den) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool on = mask[i] == 1.0f;
    if (item) item[i] = on ? -lognormal_logpdf(gt[i], mu[i], s2[i]) : 0.0f;
    if (grad_mu) {
        float dmu = 0.0f, ds2 = 0.0f, dx;
        if (on) lognormal_grad(gt[i], mu[i], s2[i], dmu, ds2, dx);
        const float sc = -scale_num[0] / scale_den[0];
        grad_mu[i] = sc * dmu; grad_s2[i] = sc * ds2;
    }
}

// out[0] = sum(x[0..n)) (double accumulation, fixed order), one block
__global__ void __launch_bounds__(256)
sum_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ out, const float *__restrict__ den) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = den ? (float)(s / (double)den[0]) : (float)s;
}

// ---------------------------------------------------------------------------
// SCST finish: trial selection, self-critical baseline, advantage, loss.  One block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scst_finish_kernel(spb_scst_args a) {
    __shared__ double sh[8];
    __shared__ int n_used;
    const int K = a.K, N = a.N;
    if (threadIdx.x == 0) {
        int used = 0;
        for (int k = 0; k < K; ++k) {
            bool ok = true;
            if (a.d_group_valid)
                for (int n = 0; n < N; ++n) ok = ok && a.d_group_valid[(int64_t)k * N + n] != 0;
            for (int n = 0; n < N && ok; ++n) ok = !isnan(a.d_reward[(int64_t)k * N + n]);
            const bool take = ok && used < a.k_use;
            a.d_trial_used[k] = take ? 1 : 0;
            used += take ? 1 : 0;
        }
        a.d_trial_used[K] = used;                       // slot K: number of trials used (k_use if enough were accepted)
        n_used = used;
    }
    __syncthreads();
    double la = 0.0, ld = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double mean = 0.0;
        for (int k = 0; k < K; ++k)
            if (a.d_trial_used[k]) mean += (double)(float)a.d_reward[(int64_t)k * N + n];   // rewards are float32 (train.py:241)
        mean = n_used > 0 ? mean / n_used : 0.0;
        for (int k = 0; k < K; ++k) {
            const int64_t g = (int64_t)k * N + n;
            float adv = 0.0f;
            if (a.d_trial_used[k]) {
                adv = (float)((double)(float)a.d_reward[g] - mean);
                if (a.d_extra_adv) adv += a.d_extra_adv[g];
            }
            a.d_adv[g] = adv;
            la += (double)(-a.d_log_actions[g]) * adv;
            ld += (double)(-a.d_log_durations[g]) * adv;
        }
    }
    la = block_sum(la, sh);
    ld = block_sum(ld, sh);
    if (threadIdx.x == 0) {
        a.d_loss[0] = (float)(la + ld); a.d_loss[1] = (float)la; a.d_loss[2] = (float)ld;
    }
}

// SCST backward: thread = (image, step); trials in order.
__global__ void __launch_bounds__(256)
scst_backward_kernel(spb_scst_args a, const float *__restrict__ grad_loss, float *__restrict__ grad_probs,
                     float *__restrict__ grad_mu, float *__restrict__ grad_s2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N * a.T) return;
    const int n = i / a.T;
    // upstream gradient on (loss, loss_actions, loss_duration); loss = loss_actions + loss_duration
    const float gla = grad_loss ? grad_loss[0] + grad_loss[1] : 1.0f, gld = grad_loss ? grad_loss[0] + grad_loss[2] : 1.0f;
    const float mu = a.d_mu[i], s2 = a.d_sigma2[i];
    float gmu = 0.0f, gs2 = 0.0f;
    for (int k = 0; k < a.K; ++k) {
        if (!a.d_trial_used[k]) continue;
        const int64_t j = (int64_t)k * a.N * a.T + i;
        const float adv = -a.d_adv[(int64_t)k * a.N + n];          // d loss_x / d Log_x[k,n]
        const float am = a.d_action_mask[j], dm = a.d_duration_mask[j];
        if (am != 0.0f) {
            const int64_t o = (int64_t)i * a.A + a.d_actions[j];
            grad_probs[o] += gla * adv * am / (a.d_mask_sums[2 * k] * (a.d_probs[o] + kEps));
        }
        if (dm != 0.0f) {
            float dmu, ds2, dx;
            lognormal_grad(a.d_dur[j], mu, s2, dmu, ds2, dx);
            const float c = gld * adv * dm / a.d_mask_sums[2 * k + 1];
            gmu += c * dmu; gs2 += c * ds2;
        }
    }
    grad_mu[i] = gmu; grad_s2[i] = gs2;
}

}  // namespace spb

using namespace spb;

extern "C" int spb_loglik_rows(const float *d_p, const float *d_probs, const int32_t *d_actions, const float *d_x,
                               const float *d_mu, const float *d_sigma2, const float *d_action_mask,
                               const float *d_duration_mask, int32_t K, int32_t N, int32_t T, int32_t A,
                               float *d_log_actions, float *d_log_durations, float *d_mask_sums, spb_stream stream) {
    SPB_CHECK_ARG(K > 0 && N > 0 && T > 0, "bad sizes");
    SPB_CHECK_ARG(d_log_actions || d_log_durations, "nothing to compute");
    SPB_CHECK_ARG(!d_log_actions || (d_action_mask && (d_p || (d_probs && d_actions && A > 0))), "LogAction inputs missing");
    SPB_CHECK_ARG(!d_log_durations || (d_duration_mask && d_x && d_mu && d_sigma2), "LogDuration inputs missing");
    loglik_rows_kernel<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(d_p, d_probs, d_actions, d_x, d_mu, d_sigma2,
                                                                      d_action_mask, d_duration_mask, N, T, A,
                                                                      d_log_actions, d_log_durations, d_mask_sums);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

extern "C" int spb_loglik_rows_backward(const float *d_p, const float *d_x, const float *d_mu, const float *d_sigma2,
                                        const float *d_action_mask, const float *d_duration_mask,
                                        const float *d_mask_sums, const float *d_grad_log_actions,
                                        const float *d_grad_log_durations, int32_t K, int32_t N, int32_t T,
                                        float *d_grad_p, float *d_grad_mu, float *d_grad_sigma2, float *d_grad_x,
                                        spb_stream stream) {
    SPB_CHECK_ARG(K > 0 && N > 0 && T > 0 && d_mask_sums, "bad sizes");
    SPB_CHECK_ARG(!d_grad_p || (d_p && d_action_mask && d_grad_log_actions), "LogAction backward inputs missing");
    SPB_CHECK_ARG(!d_grad_log_durations || (d_x && d_mu && d_sigma2 && d_duration_mask), "LogDuration backward inputs missing");
    loglik_rows_backward_kernel<<<(unsigned)((N * T + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_p, d_x, d_mu, d_sigma2, d_action_mask, d_duration_mask, d_mask_sums, d_grad_log_actions, d_grad_log_durations, K,
        N, T, d_grad_p, d_grad_mu, d_grad_sigma2, d_grad_x);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

extern "C" int spb_cross_entropy(const float *d_logits, const float *d_gt, const float *d_mask, int64_t rows, int32_t A,
                                 float *d_row_loss, float *d_loss, float *d_mask_sum, const float *d_grad_out,
                                 float *d_grad_logits, spb_stream stream) {
    SPB_CHECK_ARG(d_logits && d_gt && d_mask && rows > 0 && A > 0 && d_mask_sum, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (d_grad_logits == nullptr) {
        SPB_CHECK_ARG(d_row_loss && d_loss, "forward needs d_row_loss and d_loss");
        sum_kernel<<<1, 256, 0, s>>>(d_mask, rows, d_mask_sum, nullptr);
        SPB_LAUNCH_CHECK();
        cross_entropy_rows_kernel<<<(unsigned)rows, 256, 0, s>>>(d_logits, d_gt, d_mask, A, d_row_loss, nullptr, nullptr, nullptr);
        SPB_LAUNCH_CHECK();
        sum_kernel<<<1, 256, 0, s>>>(d_row_loss, rows, d_loss, d_mask_sum);
        SPB_LAUNCH_CHECK();
    } else {
        SPB_CHECK_ARG(d_grad_out, "backward needs the upstream gradient");
        cross_entropy_rows_kernel<<<(unsigned)rows, 256, 0, s>>>(d_logits, d_gt, d_mask, A, nullptr, d_grad_logits, d_grad_out,
                                                                d_mask_sum);
        SPB_LAUNCH_CHECK();
    }
    return SPB_OK;
}

extern "C" int spb_lognormal_nll(const float *d_mu, const float *d_sigma2, const float *d_gt, const float *d_mask,
                                 int64_t n, float *d_item, float *d_loss, float *d_mask_sum, const float *d_grad_out,
                                 float *d_grad_mu, float *d_grad_sigma2, spb_stream stream) {
    SPB_CHECK_ARG(d_mu && d_sigma2 && d_gt && d_mask && n > 0 && d_mask_sum, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (d_grad_mu == nullptr) {
        SPB_CHECK_ARG(d_item && d_loss, "forward needs d_item and d_loss");
        sum_kernel<<<1, 256, 0, s>>>(d_mask, n, d_mask_sum, nullptr);
        SPB_LAUNCH_CHECK();
        lognormal_nll_kernel<<<blocks, 256, 0, s>>>(d_mu, d_sigma2, d_gt, d_mask, n, d_item, nullptr, nullptr, nullptr, nullptr);
        SPB_LAUNCH_CHECK();
        sum_kernel<<<1, 256, 0, s>>>(d_item, n, d_loss, d_mask_sum);
        SPB_LAUNCH_CHECK();
    } else {
        SPB_CHECK_ARG(d_grad_out && d_grad_sigma2, "backward needs the upstream gradient and both outputs");
        lognormal_nll_kernel<<<blocks, 256, 0, s>>>(d_mu, d_sigma2, d_gt, d_mask, n, nullptr, d_grad_mu, d_grad_sigma2,
                                                    d_grad_out, d_mask_sum);
        SPB_LAUNCH_CHECK();
    }
    return SPB_OK;
}

static int scst_check(const spb_scst_args *a) {
    SPB_CHECK_ARG(a != nullptr, "null struct pointer");
    SPB_CHECK_ARG(a->N > 0 && a->T > 0 && a->A > 0 && a->K > 0 && a->k_use > 0 && a->k_use <= a->K, "bad sizes");
    SPB_CHECK_ARG(a->d_probs && a->d_mu && a->d_sigma2 && a->d_actions && a->d_dur && a->d_action_mask &&
                      a->d_duration_mask && a->d_reward,
                  "null input pointer");
    SPB_CHECK_ARG(a->d_loss && a->d_adv && a->d_log_actions && a->d_log_durations && a->d_trial_used && a->d_mask_sums,
                  "null output pointer");
    return SPB_OK;
}

extern "C" int spb_scst_loss(const spb_scst_args *a, spb_stream stream) {
    int rc = scst_check(a);
    if (rc != SPB_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    loglik_rows_kernel<<<(unsigned)a->K, 256, 0, s>>>(nullptr, a->d_probs, a->d_actions, a->d_dur, a->d_mu, a->d_sigma2,
                                                     a->d_action_mask, a->d_duration_mask, a->N, a->T, a->A,
                                                     a->d_log_actions, a->d_log_durations, a->d_mask_sums);
    SPB_LAUNCH_CHECK();
    scst_finish_kernel<<<1, 256, 0, s>>>(*a);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

extern "C" int spb_scst_loss_backward(const spb_scst_args *a, const float *d_grad_loss, float *d_grad_probs,
                                      float *d_grad_mu, float *d_grad_sigma2, spb_stream stream) {
    int rc = scst_check(a);
    if (rc != SPB_OK) return rc;
    SPB_CHECK_ARG(d_grad_probs && d_grad_mu && d_grad_sigma2, "null gradient pointer");
    cudaStream_t s = (cudaStream_t)stream;
    SPB_CUDA(cudaMemsetAsync(d_grad_probs, 0, (size_t)a->N * a->T * a->A * sizeof(float), s));
    scst_backward_kernel<<<(unsigned)((a->N * a->T + 255) / 256), 256, 0, s>>>(*a, d_grad_loss, d_grad_probs, d_grad_mu,
                                                                              d_grad_sigma2);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}
