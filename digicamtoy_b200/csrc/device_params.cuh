// Device-side view of one generator configuration.  Built once by dct_create (api.cu) from the
// host dct_config; passed to every kernel by value.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace dct {

struct DevParams {
    // geometry
    int P, K, B, n_drop;          // pixels, pulses per pixel, kept bins, dropped leading bins
    int n_iv;                     // template intervals (= knots - 1)
    int poisson, sub_binning;
    int adc_min, adc_max;
    // time grid, in the "onset" coordinate s = (t - t0) + x_lo, where bins sit at b*dt and the
    // template argument is y = b*dt - s in [0, L]   (x_lo = first knot, L = template length)
    float dt, inv_dt;
    float span;                   // t_end - t0, length of the NSB window (reference :244)
    float x_lo, L;
    float knot_dx, inv_knot_dx;
    float t0f;
    int   cell0;                  // floor(x_lo / dt): cell of the first possible NSB onset
    float off0;                   // x_lo - cell0*dt  in [0, dt)
    // polyphase float32 table: tap j, phase f -> interval f + n_phase*j   (valid if n_phase > 0)
    int n_phase, n_taps;
    float phase_cap;              // 2^23 + n_phase - 1 (see phase_of)
    // on-grid taps for sub_binning > 0: h(m*dt), m = grid_m_lo .. grid_m_lo + grid_n - 1
    int grid_m_lo, grid_n;
    // double precision copies for the replay kernels
    double t0_d, dt_d, x_lo_d, x_hi_d, inv_knot_dx_d;

    // per-pixel float32 SoA, [P]
    const float* rate;            // GHz
    const float* neg_ln2_over_rate;  // -ln2 / rate [ns] (0 where rate == 0): gap = lg2(u) * this
    const float* xt;
    const float4* borel_cdf;      // P(n<=1), P(n<=2), P(n<=3), P(n<=4) of Borel(xt)
    const float* gain;
    const float* sigma1;          // relative
    const float* sigma_e;
    const float* baseline;
    // per-pixel float64 copies for replay, [P]
    const double* gain_d;
    const double* baseline_d;
    // per-pulse tables, [P, K]
    const float* npe;
    const float* tsig;
    const float* jitter;
    const float* sig_cdf;         // [P, K][32] generalised-Poisson tables of the small pulses (poisson, 0 < n_photon < 12), else null
    // template tables
    const float4* coef32;         // [n_iv]  (c0, c1, c2, c3)
    const double* coef64;         // [n_iv][4]
    const double* knots;          // [n_iv + 1]
    const float4* poly;           // [n_taps][n_phase]
    const float*  grid_tap;       // [grid_n]
    const float*  gp_cdf;         // [P][32] cumulative generalised-Poisson probabilities of the on-grid mode (sub_binning > 0), else null
};

}  // namespace dct
