// K2: deterministic replay of recorded photo-electron lists through shaping + digitisation.
//
// Reproduces compute_analog_signal / generate_electronic_smearing / convert_to_digital
// (reference digicamtoy/tracegenerator/__init__.py:350-381) for pe times and amplitudes captured
// from the reference itself, so GPU arithmetic can be compared with the reference sample by
// sample.  One warp per trace, lanes over bins, the pe list streamed once.
#pragma once
#include "device_params.cuh"
#include "generate.cuh"

namespace dct {

constexpr int REPLAY_WARPS = 4;
constexpr int REPLAY_BINS_PER_LANE = 4;       // 128 bins per pass

struct ReplayArgs {
    DevParams p;
    uint64_t n_traces;
    const int64_t* nsb_offsets;
    const double* nsb_time;
    const double* nsb_amp;
    const double* sig_time;
    const double* sig_amp;
    const double* e_noise;
    int16_t* adc;
    unsigned long long* n_clamped;
};

template <typename T> struct Coef4 { T c0, c1, c2, c3; };

template <typename T>
__device__ __forceinline__ T replay_template(const DevParams& p, const Coef4<T>* tab, double x) {
    // inclusive support [x_lo, x_hi], zero outside (interp1d fill_value = 0, :141-143)
    if (!(x >= p.x_lo_d && x <= p.x_hi_d)) return (T)0;
    int m = (int)floor((x - p.x_lo_d) * p.inv_knot_dx_d);
    m = max(0, min(m, p.n_iv - 1));
    // guard the floor against the product rounding across a knot
    if (m > 0 && x < p.knots[m]) --m;
    else if (m < p.n_iv - 1 && x >= p.knots[m + 1]) ++m;
    const T u = (T)(x - p.knots[m]);
    const Coef4<T> c = tab[m];
    return ((c.c3 * u + c.c2) * u + c.c1) * u + c.c0;
}

template <typename T>
__global__ void __launch_bounds__(REPLAY_WARPS * 32)
k_replay(const ReplayArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& p = a.p;
    Coef4<T>* tab = reinterpret_cast<Coef4<T>*>(smem_raw);
    for (int i = threadIdx.x; i < p.n_iv; i += blockDim.x) {
        Coef4<T> c;
        c.c0 = (T)p.coef64[4 * i + 0];
        c.c1 = (T)p.coef64[4 * i + 1];
        c.c2 = (T)p.coef64[4 * i + 2];
        c.c3 = (T)p.coef64[4 * i + 3];
        tab[i] = c;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5);
    if (g >= a.n_traces) return;
    const int pixel = (int)(g % (uint64_t)p.P);
    const int64_t lo = a.nsb_offsets[g], hi = a.nsb_offsets[g + 1];
    const T gain = (T)p.gain_d[pixel];
    const T baseline = (T)p.baseline_d[pixel];
    unsigned int clamped = 0;

    for (int b_base = 0; b_base < p.B; b_base += 32 * REPLAY_BINS_PER_LANE) {
        double tb[REPLAY_BINS_PER_LANE];
        T acc_nsb[REPLAY_BINS_PER_LANE], acc_sig[REPLAY_BINS_PER_LANE];
#pragma unroll
        for (int r = 0; r < REPLAY_BINS_PER_LANE; ++r) {
            const int b = b_base + lane + 32 * r;
            // np.arange(time_start, time_end, dt)[n_drop + b]  (:197-198, :377-378)
            tb[r] = p.t0_d + (double)(p.n_drop + b) * p.dt_d;
            acc_nsb[r] = (T)0;
            acc_sig[r] = (T)0;
        }
        // NSB list, in the reference's order (np.sum over the pe axis, :356)
        for (int64_t i = lo; i < hi; ++i) {
            const double t = a.nsb_time[i];
            const T amp = (T)a.nsb_amp[i];
#pragma unroll
            for (int r = 0; r < REPLAY_BINS_PER_LANE; ++r)
                acc_nsb[r] += replay_template<T>(p, tab, tb[r] - t) * amp;
        }
        // signal pulses, summed separately and then added (:357-361)
        for (int k = 0; k < p.K; ++k) {
            const double t = a.sig_time[g * p.K + k];
            const T amp = (T)a.sig_amp[g * p.K + k];
#pragma unroll
            for (int r = 0; r < REPLAY_BINS_PER_LANE; ++r)
                acc_sig[r] += replay_template<T>(p, tab, tb[r] - t) * amp;
        }
#pragma unroll
        for (int r = 0; r < REPLAY_BINS_PER_LANE; ++r) {
            const int b = b_base + lane + 32 * r;
            if (b < p.B) {
                const T noise = a.e_noise ? (T)a.e_noise[g * p.B + b] : (T)0;
                T v = (acc_nsb[r] + acc_sig[r]) * gain;     // :361-362
                v = v + noise;                               // :371
                v = v + baseline;                            // :379
                const double rounded = rint((double)v);      // :380 round half to even
                double cl = rounded;
                if (cl < (double)p.adc_min) { cl = (double)p.adc_min; ++clamped; }
                else if (cl > (double)p.adc_max) { cl = (double)p.adc_max; ++clamped; }
                a.adc[g * p.B + b] = (int16_t)(int)cl;
            }
        }
    }
    if (a.n_clamped) {
        for (int o = 16; o > 0; o >>= 1) clamped += __shfl_xor_sync(0xffffffffu, clamped, o);
        if (lane == 0 && clamped) atomicAdd(a.n_clamped, (unsigned long long)clamped);
    }
}

}  // namespace dct
