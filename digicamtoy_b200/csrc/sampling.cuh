// Samplers for the distributions the reference draws with np.random (reference
// digicamtoy/tracegenerator/__init__.py):
//   Poisson process of NSB photo-electrons  (:244-250  Poisson count + iid uniform times)
//   Borel cluster size of optical crosstalk (:269-277  Galton-Watson recursion, closed form in
//                                            reference digicamtoy/utils/pdf.py:26-32)
//   Poisson(n_photon) signal size + one Galton-Watson tree per primary (:236, :319-324)
//   Gaussian gain smearing / electronic noise (:337, :342, :369)
// Transcendentals are the single-instruction MUFU approximations (lg2 / sqrt / sin / cos .approx):
// absolute errors ~1e-7..1e-6, far below the 2^-24 granularity of the uniforms feeding them.
#pragma once
#include "philox.cuh"

// -DDCT_EXP_NO_TAIL_STOP=1 builds the inversion loops without their tail stop (to show what
// tests/test_generate_gpu.py::test_inversion_samplers_stop_in_the_tail_for_the_largest_uniforms guards).
#if DCT_EXP_NO_TAIL_STOP
#define DCT_TAIL_STOP(next, cur)
#else
#define DCT_TAIL_STOP(next, cur) if ((next) == (cur)) break;
#endif

namespace dct {

__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// angle in [-pi, pi) from 23 random bits: one FFMA, no I2F; best range of MUFU.SIN/COS
__device__ __forceinline__ float random_angle(uint32_t r) {
    return fmaf(__uint_as_float(0x3f800000u | (r >> 9)), 6.28318530718f, -9.42477796077f);
}

constexpr float NEG_2LN2 = -1.38629436112f;     // -2 ln 2 : -2 ln(u) = NEG_2LN2 * lg2(u)
constexpr float NEG_LN2 = -0.69314718056f;

// Box-Muller; both outputs are independent N(0, s2) where the variance scale s2 multiplies -2 ln u
// under the root (lets callers fold sqrt(n) of the gain smearing into the same MUFU.SQRT).
__device__ __forceinline__ void normal_pair(uint32_t r0, uint32_t r1, float& z0, float& z1) {
    const float rad = fast_sqrt(fast_lg2(u01_tail(r0)) * NEG_2LN2);
    float s, c;
    __sincosf(random_angle(r1), &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// One N(0, var_scale) draw: cos(angle) * sqrt(-2 var_scale ln u).
__device__ __forceinline__ float normal_scaled(uint32_t r0, uint32_t r1, float var_scale) {
    return fast_sqrt(fast_lg2(u01_tail(r0)) * (NEG_2LN2 * var_scale)) * __cosf(random_angle(r1));
}
__device__ __forceinline__ float normal_one(uint32_t r0, uint32_t r1) { return normal_scaled(r0, r1, 1.0f); }

// Exponential inter-arrival time of a Poisson process; neg_ln2_over_rate = -ln2 / rate.
__device__ __forceinline__ float exp_gap(uint32_t r, float neg_ln2_over_rate) {
    return fast_lg2(u01_open0(r)) * neg_ln2_over_rate;
}

// Borel(mu) by inversion: P(1) = e^-mu, P(n+1)/P(n) = mu e^-mu (1+1/n)^(n-1).
// Sequential search, used for the tail only.
__device__ __noinline__ int borel_search(float u, float mu) {
    const float e = __expf(-mu);
    float p = e, cdf = e;
    const float step = mu * e;
    int n = 1;
    while (u >= cdf && n < 256) {
        const float fn = (float)n;
        p *= step * __powf(1.0f + 1.0f / fn, fn - 1.0f);
        ++n;
        // The float sum of the masses can stop a few 1e-8 short of 1: once a term no longer moves
        // it, the largest uniforms (1 - 2^-23) would otherwise run to the cap -- a 256-pe cluster
        // once per ~1e7 draws.  The value reached here is where the tail has that probability.
        const float next = cdf + p;
        DCT_TAIL_STOP(next, cdf)
        cdf = next;
    }
    return n;
}

// Borel(mu) with the first four cumulative probabilities precomputed per pixel (cdf.x = P(n<=1) ..
// cdf.w = P(n<=4)): branch-free for all but P(n>4) of the draws (1e-4 at mu = 0.08).
__device__ __forceinline__ int borel(uint32_t r, float mu, const float4 cdf) {
    const float u = u01_open1(r);
    int n = 1 + (u >= cdf.x) + (u >= cdf.y) + (u >= cdf.z) + (u >= cdf.w);
    if (u >= cdf.w) n = max(5, borel_search(u, mu));
    return n;
}

// Poisson(lambda), large means: Hormann's PTRS transformed rejection (1993), acceptance test in
// double so that k ~ 1e4 keeps a sharp log-factorial.  Out of line: it is rare and register-hungry.
__device__ __noinline__ int poisson_ptrs(BlockReader& rd, float lambda) {
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

// Poisson(lambda).  Small means: sequential inversion.
__device__ __forceinline__ int poisson(BlockReader& rd, float lambda) {
    if (!(lambda > 0.0f)) return 0;
    if (lambda >= 12.0f) return poisson_ptrs(rd, lambda);
    const float u = u01_open1(rd.next());
    float p = __expf(-lambda), cdf = p;
    int k = 0;
    while (u >= cdf && k < 200) {
        ++k;
        p *= lambda / (float)k;
        const float next = cdf + p;
        DCT_TAIL_STOP(next, cdf)                  // tail below float resolution (see borel_search)
        cdf = next;
    }
    return k;
}

// Total progeny of n_primary independent Galton-Watson trees with Poisson(mu) offspring, generation
// by generation: the sum of X Poisson(mu) variables is Poisson(mu X), so a pulse of thousands of pe
// costs a handful of draws.  Law: generalised Poisson (reference utils/pdf.py:37-42).
__device__ __noinline__ int branching_total(BlockReader& rd, int n_primary, float mu) {
    int total = n_primary, gen = n_primary;
    if (!(mu > 0.0f)) return total;
    for (int g = 0; g < 4096 && gen > 0; ++g) {
        gen = poisson(rd, mu * (float)gen);
        total += gen;
    }
    return total;
}

// Poisson(lambda) by sequential inversion of one given uniform; p0 = exp(-lambda) (callers that
// draw many counts at one mean keep it in a register).  lambda < 12.
__device__ __forceinline__ int poisson_inv(float u, float lambda, float p0) {
    float p = p0, cdf = p0;
    int k = 0;
    while (u >= cdf && k < 200) {
        ++k;
        p *= lambda * __frcp_rn((float)k);
        const float next = cdf + p;
        DCT_TAIL_STOP(next, cdf)                  // tail below float resolution (see borel_search)
        cdf = next;
    }
    return k;
}

// The same inversion when many counts are drawn at one mean: the first POISSON_TABLE cumulative
// probabilities are tabulated once (same float recurrence, so the result is the one poisson_inv
// gives) and searched without a data-dependent loop.  cdf[i] = P(k <= i); stride in floats.
constexpr int POISSON_TABLE = 16;

__device__ __forceinline__ void poisson_table_fill(float* cdf, int stride, float lambda, float p0) {
    float p = p0, c = p0;
    bool live = true;                        // poisson_inv stops at the first term that no longer moves the sum
    cdf[0] = c;
    for (int k = 1; k < POISSON_TABLE; ++k) {
        p *= lambda * __frcp_rn((float)k);
        const float next = c + p;
        live = live && next != c;
        c = next;
        cdf[k * stride] = live ? c : __int_as_float(0x7f800000);   // +inf: no uniform gets past it
    }
}

__device__ __forceinline__ int poisson_table_draw(float u, const float* cdf, int stride, float lambda, float p0) {
    // number of entries among cdf[0..14] that are <= u (the table is non-decreasing)
    int k = (u >= cdf[7 * stride]) ? 8 : 0;
    k += (u >= cdf[(k + 3) * stride]) ? 4 : 0;
    k += (u >= cdf[(k + 1) * stride]) ? 2 : 0;
    k += (u >= cdf[k * stride]) ? 1 : 0;
    if (k == POISSON_TABLE - 1 && u >= cdf[(POISSON_TABLE - 1) * stride]) {
        // beyond the table: carry the recurrence on from its last entry
        float p = p0;
        for (int i = 1; i < POISSON_TABLE; ++i) p *= lambda * __frcp_rn((float)i);
        float c = cdf[(POISSON_TABLE - 1) * stride];
        k = POISSON_TABLE - 1;
        while (u >= c && k < 200) {
            ++k;
            p *= lambda * __frcp_rn((float)k);
            const float next = c + p;
            DCT_TAIL_STOP(next, c)
            c = next;
        }
    }
    return k;
}

constexpr int GP_TABLE_ENTRIES = 32;
__device__ int gp_tail_entry(float u, float lam, float mu, float cdf_last);

// The same look-up for tables that stay in global memory (one per pixel and pulse): a linear scan
// in 16-byte pieces, most draws end in the first one or two.  Same result as gp_table_draw.
__device__ __forceinline__ int gp_table_scan(float u, const float* cdf, float lam, float mu) {
    const float4* c4 = reinterpret_cast<const float4*>(cdf);
    for (int i = 0; i < GP_TABLE_ENTRIES / 4; ++i) {
        const float4 v = c4[i];
        if (u < v.w) return 4 * i + (u >= v.x) + (u >= v.y) + (u >= v.z);
    }
    return gp_tail_entry(u, lam, mu, cdf[GP_TABLE_ENTRIES - 1]);
}

// ---- On-grid NSB mode: the total number of pe on one sample = Poisson(lam) primaries plus their
// Galton-Watson crosstalk progeny (Poisson(mu) offspring), i.e. the generalised Poisson law
//     P(T = k) = lam (lam + k mu)^(k-1) e^(-lam - k mu) / k!          (reference utils/pdf.py:37-42;
// what generate_nsb :259-267 followed by generate_crosstalk :306-312 leaves on a sample).
// Sampled by inversion of ONE uniform in a per-pixel table of the first GP_TABLE cumulative
// probabilities, computed in double by dct_create; beyond the table (P ~ 1e-9 at 1 GHz, xt 0.08)
// the recurrence carries on in float.
constexpr int GP_TABLE = GP_TABLE_ENTRIES;
constexpr float GP_LAM_MAX = 12.0f;      // pixels with rate * dt at or above it keep the generation-by-generation sampler

__device__ __noinline__ int gp_tail(float u, float lam, float mu, float cdf_last) {
    float cdf = cdf_last;
    const float ln_lam = logf(lam);
    for (int k = GP_TABLE; k < 4096; ++k) {
        const float fk = (float)k, a = fmaf(mu, fk, lam);
        const float p = expf(ln_lam + (fk - 1.0f) * logf(a) - a - lgammaf(fk + 1.0f));
        const float next = cdf + p;
        if (next == cdf) return k;                 // tail below float resolution (see borel_search)
        cdf = next;
        if (u < cdf) return k;
    }
    return 4096;
}

__device__ __forceinline__ int gp_tail_entry(float u, float lam, float mu, float cdf_last) { return gp_tail(u, lam, mu, cdf_last); }

// T = min { k : u < cdf[k] }; cdf[i * stride] = P(T <= i), non-decreasing, i < GP_TABLE.
__device__ __forceinline__ int gp_table_draw(float u, const float* cdf, int stride, float lam, float mu) {
    int k = (u >= cdf[15 * stride]) ? 16 : 0;
    k += (u >= cdf[(k + 7) * stride]) ? 8 : 0;
    k += (u >= cdf[(k + 3) * stride]) ? 4 : 0;
    k += (u >= cdf[(k + 1) * stride]) ? 2 : 0;
    k += (u >= cdf[k * stride]) ? 1 : 0;
    if (k == GP_TABLE - 1 && u >= cdf[(GP_TABLE - 1) * stride]) k = gp_tail(u, lam, mu, cdf[(GP_TABLE - 1) * stride]);
    return k;
}

// On-grid NSB mode, one simulated bin: the part of the count that the bin's own Philox block could
// not provide (see subbin_cluster in generate.cuh) -- a Poisson mean >= 12 (PTRS needs a variable
// number of uniforms) or a crosstalk chain that outlived the draws set aside for it.  Reads from
// `first_block` of STREAM_SUBBIN on: blocks (bs + 1) << 16 ..., which no bin uses as its own block
// (bins < 65536).  gen < 0: the primaries are still to be drawn; else `gen` pe of the last
// generation still need their offspring.
__device__ __noinline__ int subbin_overflow(uint32_t k0, uint32_t k1, uint32_t ev_lo, uint32_t ev_hi, uint32_t pixel,
                                            uint32_t first_block, float lambda, float mu, int total, int gen) {
    StreamKey key;
    key.k0 = k0; key.k1 = k1; key.ev_lo = ev_lo; key.ev_hi = ev_hi; key.pixel = pixel;
    BlockReader rd(key, STREAM_SUBBIN, first_block);
    if (gen < 0) total = gen = poisson(rd, lambda);
    if (mu > 0.0f)
        for (int g = 0; g < 4096 && gen > 0; ++g) {
            gen = poisson(rd, mu * (float)gen);
            total += gen;
        }
    return total;
}

}  // namespace dct
