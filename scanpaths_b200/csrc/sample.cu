// K5 sample_paths: K post-hoc samples per image from the decoder's per-step
// distributions.  Replaces Sampling.random_sample + Sampling.generate_scanpath
// (models/sampling.py:16-77), including their quirks:
//   * the stop action is masked for the first min_length steps, then the row is
//     renormalised in float32 (Categorical(probs=...));
//   * torch.multinomial(p, 1) is the exponential race argmax(p_norm / q), q ~ Exp(1),
//     first index on ties -- with the reference's q injected the draw is identical;
//   * the selected probability is gathered from the UNMASKED probs;
//   * duration = exp(z * sigma2 + mu): the variance is used as the scale (:27);
//   * scanpath_length ignores a stop at step 0 (:29-33).
// One warp per (image, step) row: the 1201 probabilities are read from HBM once,
// normalised in registers and reused for all K samples (the reference re-reads
// them K times and syncs with the host 2N times per sample).
#include <math.h>

#include "common.cuh"

namespace spb {

constexpr int kMaxPerLane = 38;   // A <= 1216

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t x) {        // (0, 1), 24 bits
    return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256)
sample_actions_kernel(const float *__restrict__ probs, const float *__restrict__ mu,
                      const float *__restrict__ sigma2, const float *__restrict__ q_in,
                      const float *__restrict__ z_in, uint2 key, int N, int T, int A, int K, int min_len,
                      int32_t *__restrict__ actions, float *__restrict__ sel_prob, float *__restrict__ dur) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t rows = (int64_t)N * T;
    const int nper = (A + 31) / 32;
    for (int64_t row = gwarp; row < rows; row += nwarps) {
        const int t = (int)(row % T);
        const float *prow = probs + row * A;
        float pn[kMaxPerLane];
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxPerLane; ++r) {
            const int a = lane + 32 * r;
            float p = (r < nper && a < A) ? prow[a] : 0.0f;
            if (a == 0 && t < min_len) p = 0.0f;                 // probs[:, :min_length, 0] = 0
            pn[r] = p;
            s += (double)p;
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float sumf = (float)s;                             // float32 row sum
#pragma unroll
        for (int r = 0; r < kMaxPerLane; ++r) pn[r] = __fdiv_rn(pn[r], sumf);
        const float mu_r = mu[row], s2_r = sigma2[row];

        for (int k = 0; k < K; ++k) {
            const int64_t orow = (int64_t)k * rows + row;        // [K, N, T]
            float best_v = -INFINITY;
            int best_a = 0x7fffffff;
            if (q_in != nullptr) {
                const float *qrow = q_in + orow * A;
#pragma unroll
                for (int r = 0; r < kMaxPerLane; ++r) {
                    const int a = lane + 32 * r;
                    if (r < nper && a < A) {
                        const float v = __fdiv_rn(pn[r], qrow[a]);
                        if (v > best_v) { best_v = v; best_a = a; }
                    }
                }
            } else {
#pragma unroll
                for (int r4 = 0; r4 < (kMaxPerLane + 3) / 4; ++r4) {
                    const uint4 x = philox4x32_10(
                        make_uint4((uint32_t)orow, (uint32_t)(orow >> 32), (uint32_t)(lane | (r4 << 5)), 0u), key);
                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = r4 * 4 + e;
                        const int a = lane + 32 * r;
                        if (r < kMaxPerLane && r < nper && a < A) {
                            const float v = __fdiv_rn(pn[r], -__logf(u01(xs[e])));
                            if (v > best_v) { best_v = v; best_a = a; }
                        }
                    }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
                const int oa = __shfl_xor_sync(0xffffffffu, best_a, o);
                if (ov > best_v || (ov == best_v && oa < best_a)) { best_v = ov; best_a = oa; }
            }
            if (lane == 0) {
                if (best_a >= A) best_a = 0;
                actions[orow] = best_a;
                sel_prob[orow] = prow[best_a];                   // gather from the unmasked probs (:23-24)
                float z;
                if (z_in != nullptr) z = z_in[orow];
                else {
                    const uint4 x = philox4x32_10(make_uint4((uint32_t)orow, (uint32_t)(orow >> 32), 0u, 1u), key);
                    z = sqrtf(-2.0f * logf(u01(x.x))) * cospif(2.0f * u01(x.y));   // Box-Muller
                }
                dur[orow] = expf(__fadd_rn(__fmul_rn(z, s2_r), mu_r));
            }
        }
    }
}

// generate_scanpath (sampling.py:48-77) + scanpath_length (:29-34): one thread per sample.
__global__ void __launch_bounds__(256)
finalize_paths_kernel(const int32_t *__restrict__ actions, const float *__restrict__ dur, int64_t n_samples, int T,
                      spb_sample_geom g, float *__restrict__ action_mask, float *__restrict__ duration_mask,
                      float *__restrict__ length, double *__restrict__ xyd, int32_t *__restrict__ len) {
    const double xg = (double)g.width / (double)g.map_width, yg = (double)g.height / (double)g.map_height;
    for (int64_t sidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; sidx < n_samples;
         sidx += (int64_t)gridDim.x * blockDim.x) {
        const int32_t *a = actions + sidx * T;
        float lq = 0.0f;
        for (int t = 0; t < T; ++t)
            if (lq == 0.0f && a[t] == 0) lq = (float)t;
        if (lq == 0.0f) lq = (float)T;
        length[sidx] = lq;
        int n = 0;
        bool stopped = false;
        for (int t = 0; t < T; ++t) {
            float am = 0.0f, dm = 0.0f;
            double x = 0.0, y = 0.0, d = 0.0;
            if (!stopped) {
                am = 1.0f;
                if (a[t] == 0) stopped = true;
                else {
                    const int cell = a[t] - 1;
                    x = (double)(cell % g.map_width) * xg + xg / 2;
                    y = (double)(cell / g.map_width) * yg + yg / 2;
                    d = (double)dur[sidx * T + t];
                    dm = 1.0f;
                    double *o = xyd + (sidx * T + n) * 3;
                    o[0] = x; o[1] = y; o[2] = d;
                    ++n;
                }
            }
            action_mask[sidx * T + t] = am;
            duration_mask[sidx * T + t] = dm;
        }
        for (int t = n; t < T; ++t) {
            double *o = xyd + (sidx * T + t) * 3;
            o[0] = 0.0; o[1] = 0.0; o[2] = 0.0;
        }
        len[sidx] = n;
    }
}

}  // namespace spb

extern "C" int spb_sample_paths(const float *d_probs, const float *d_mu, const float *d_sigma2, const float *d_q,
                                const float *d_z, uint64_t seed, int32_t N, int32_t T, int32_t A, int32_t K,
                                const spb_sample_geom *geom, int32_t *d_actions, float *d_sel_prob, float *d_dur,
                                float *d_action_mask, float *d_duration_mask, float *d_length, double *d_xyd,
                                int32_t *d_len, spb_stream stream) {
    SPB_CHECK_ARG(geom != nullptr, "geom is null");
    SPB_CHECK_ARG(N >= 0 && T > 0 && A > 0 && K > 0, "bad sizes");
    SPB_CHECK_ARG(A <= 32 * spb::kMaxPerLane, "A > 1216 actions not supported");
    SPB_CHECK_ARG(A == geom->map_width * geom->map_height + 1, "A must be map_width*map_height + 1");
    if (N == 0) return SPB_OK;
    SPB_CHECK_ARG(d_probs && d_mu && d_sigma2 && d_actions && d_sel_prob && d_dur && d_action_mask &&
                      d_duration_mask && d_length && d_xyd && d_len, "null device pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rows = (int64_t)N * T;
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)spb::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    spb::prof_begin(spb::kTagSample, s);
    spb::sample_actions_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_probs, d_mu, d_sigma2, d_q, d_z, key, N, T, A, K,
                                                                 geom->min_length, d_actions, d_sel_prob, d_dur);
    SPB_LAUNCH_CHECK();
    spb::prof_end(s);
    const int64_t ns = (int64_t)K * N;
    int64_t fb = (ns + 255) / 256;
    if (fb > cap) fb = cap;
    spb::finalize_paths_kernel<<<(unsigned)fb, 256, 0, s>>>(d_actions, d_dur, ns, T, *geom, d_action_mask,
                                                             d_duration_mask, d_length, d_xyd, d_len);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

extern "C" int spb_generate_scanpaths(const int32_t *d_actions, const float *d_dur, int64_t n_samples, int32_t T,
                                      const spb_sample_geom *geom, float *d_action_mask, float *d_duration_mask,
                                      float *d_length, double *d_xyd, int32_t *d_len, spb_stream stream) {
    SPB_CHECK_ARG(geom != nullptr, "geom is null");
    SPB_CHECK_ARG(n_samples >= 0 && T > 0, "bad sizes");
    if (n_samples == 0) return SPB_OK;
    SPB_CHECK_ARG(d_actions && d_dur && d_action_mask && d_duration_mask && d_length && d_xyd && d_len,
                  "null device pointer");
    int64_t fb = (n_samples + 255) / 256;
    const int64_t cap = (int64_t)spb::num_sms() * 8;
    if (fb > cap) fb = cap;
    spb::finalize_paths_kernel<<<(unsigned)fb, 256, 0, (cudaStream_t)stream>>>(
        d_actions, d_dur, n_samples, T, *geom, d_action_mask, d_duration_mask, d_length, d_xyd, d_len);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}
