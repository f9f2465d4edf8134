// Samplers for the distributions the reference draws with np.random (reference
// digicamtoy/tracegenerator/__init__.py):
//   Poisson process of NSB photo-electrons  (:244-250  Poisson count + iid uniform times)
//   Borel cluster size of optical crosstalk (:269-277  Galton-Watson recursion, closed form in
//                                            reference digicamtoy/utils/pdf.py:26-32)
//   Poisson(n_photon) signal size + one Galton-Watson tree per primary (:236, :319-324)
//   Gaussian gain smearing / electronic noise (:337, :342, :369)
#pragma once
#include "philox.cuh"

namespace dct {

// Box-Muller; both outputs are independent N(0,1).
__device__ __forceinline__ void normal_pair(uint32_t r0, uint32_t r1, float& z0, float& z1) {
    const float u = u01_open0(r0);
    // angle in [-pi, pi): best accuracy range of the MUFU sin/cos; one FFMA, no I2F
    const float ang = fmaf(__uint_as_float(0x3f800000u | (r1 >> 9)), 6.28318530718f, -9.42477796077f);
    const float rad = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(ang, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

__device__ __forceinline__ float normal_one(uint32_t r0, uint32_t r1) {
    const float u = u01_open0(r0);
    const float ang = fmaf(__uint_as_float(0x3f800000u | (r1 >> 9)), 6.28318530718f, -9.42477796077f);
    return sqrtf(-2.0f * __logf(u)) * __cosf(ang);
}

// Exponential inter-arrival time of a Poisson process with the given 1/rate.
__device__ __forceinline__ float exp_gap(uint32_t r, float inv_rate) {
    return -__logf(u01_open0(r)) * inv_rate;
}

// Borel(mu) by inversion: P(1) = e^-mu, P(n+1)/P(n) = mu e^-mu (1+1/n)^(n-1).
// exp_neg_mu = e^-mu is precomputed per pixel.  92 % of the draws (mu = 0.08) leave at the first test.
__device__ __forceinline__ int borel(uint32_t r, float mu, float exp_neg_mu) {
    const float u = u01_open1(r);
    if (u < exp_neg_mu) return 1;
    float p = exp_neg_mu, cdf = exp_neg_mu;
    const float step = mu * exp_neg_mu;
    int n = 1;
    while (u >= cdf && n < 256) {
        const float fn = (float)n;
        p *= step * __powf(1.0f + 1.0f / fn, fn - 1.0f);
        cdf += p;
        ++n;
    }
    return n;
}

// Poisson(lambda).  Small means: sequential inversion.  Large means: Hormann's PTRS transformed
// rejection (1993), acceptance test in double so that k ~ 1e4 keeps a sharp log-factorial.
__device__ inline int poisson(BlockReader& rd, float lambda) {
    if (!(lambda > 0.0f)) return 0;
    if (lambda < 12.0f) {
        const float u = u01_open1(rd.next());
        float p = __expf(-lambda), cdf = p;
        int k = 0;
        while (u >= cdf && k < 200) {
            ++k;
            p *= lambda / (float)k;
            cdf += p;
        }
        return k;
    }
    const double lam = (double)lambda;
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
        const double U = (double)u01_open1(rd.next()) - 0.5 + 2.98e-8;
        const double V = (double)u01_open0(rd.next());
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (int)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0))
            return (int)kf;
    }
    return (int)lam;
}

// Total progeny of n_primary independent Galton-Watson trees with Poisson(mu) offspring, generation
// by generation: the sum of X Poisson(mu) variables is Poisson(mu X), so a pulse of thousands of pe
// costs a handful of draws.  Law: generalised Poisson (reference utils/pdf.py:37-42).
__device__ inline int branching_total(BlockReader& rd, int n_primary, float mu) {
    int total = n_primary, gen = n_primary;
    if (!(mu > 0.0f)) return total;
    for (int g = 0; g < 4096 && gen > 0; ++g) {
        gen = poisson(rd, mu * (float)gen);
        total += gen;
    }
    return total;
}

}  // namespace dct
