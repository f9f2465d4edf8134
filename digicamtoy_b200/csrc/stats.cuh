// K3: validation statistics of an ADC cube, and K0: the write-only calibration kernel.
//
// K3 accumulates what the reference's analysis scripts compute on the host from the full cube
// (reference nsb_baseline.py:21-32: mean / std / 4th moment over waveforms x bins; ADC histogram
// of digicamtoy/test/test.py:40-47) so that a sharded run only has to all-reduce a few KB.
#pragma once
#include "device_params.cuh"

namespace dct {

constexpr int STATS_THREADS = 256;
constexpr int STATS_BINS_PER_LANE = 4;

struct StatsArgs {
    int P, B, adc_min, adc_max;
    uint64_t n_traces;
    const int16_t* adc;
    const float* baseline;          // [P]
    unsigned long long* adc_hist;
    double* bin_sum;
    double* bin_sumsq;
    double* pixel_sum;
    double* pixel_sumsq;
    unsigned long long* charge_hist;
    int charge_lo;
    int n_charge_bins;
    double* moments;
};

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per trace (lanes over bins: coalesced 2-byte loads), grid-stride over traces.
__global__ void __launch_bounds__(STATS_THREADS)
k_stats(const StatsArgs a) {
    extern __shared__ unsigned int hist_s[];        // [adc_max - adc_min + 1] if adc_hist
    const int n_hist = a.adc_max - a.adc_min + 1;
    if (a.adc_hist) {
        for (int i = threadIdx.x; i < n_hist; i += blockDim.x) hist_s[i] = 0u;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (STATS_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t n_warps = (uint64_t)gridDim.x * (STATS_THREADS / 32);

    double m1 = 0, m2 = 0, m3 = 0, m4 = 0;
    for (int b_base = 0; b_base < a.B; b_base += 32 * STATS_BINS_PER_LANE) {
        double bs[STATS_BINS_PER_LANE] = {0, 0, 0, 0}, bq[STATS_BINS_PER_LANE] = {0, 0, 0, 0};
        const bool first_pass = (b_base == 0);
        const bool single_pass = a.B <= 32 * STATS_BINS_PER_LANE;
        for (uint64_t g = warp; g < a.n_traces; g += n_warps) {
            const int16_t* row = a.adc + g * (uint64_t)a.B;
            double ts = 0, tq = 0;
#pragma unroll
            for (int r = 0; r < STATS_BINS_PER_LANE; ++r) {
                const int b = b_base + lane + 32 * r;
                if (b < a.B) {
                    const int v = row[b];
                    const double x = (double)v;
                    bs[r] += x; bq[r] += x * x;
                    ts += x; tq += x * x;
                    m1 += x; m2 += x * x; m3 += x * x * x; m4 += x * x * x * x;
                    if (a.adc_hist) atomicAdd(&hist_s[v - a.adc_min], 1u);
                }
            }
            if (a.pixel_sum || a.charge_hist) {
                // per-trace quantities; a trace wider than one pass is completed with atomics
                ts = warp_sum(ts);
                tq = warp_sum(tq);
                const int pixel = (int)(g % (uint64_t)a.P);
                if (lane == 0) {
                    if (a.pixel_sum) atomicAdd(a.pixel_sum + pixel, ts);
                    if (a.pixel_sumsq) atomicAdd(a.pixel_sumsq + pixel, tq);
                    if (a.charge_hist && single_pass) {
                        const long long q = llrint(ts) - (long long)a.B * llrintf(a.baseline[pixel]);
                        long long bin = q - a.charge_lo;
                        bin = bin < 0 ? 0 : (bin >= a.n_charge_bins ? a.n_charge_bins - 1 : bin);
                        atomicAdd(a.charge_hist + bin, 1ull);
                    }
                }
            }
        }
        (void)first_pass;
#pragma unroll
        for (int r = 0; r < STATS_BINS_PER_LANE; ++r) {
            const int b = b_base + lane + 32 * r;
            if (b < a.B) {
                if (a.bin_sum) atomicAdd(a.bin_sum + b, bs[r]);
                if (a.bin_sumsq) atomicAdd(a.bin_sumsq + b, bq[r]);
            }
        }
    }
    if (a.moments) {
        m1 = warp_sum(m1); m2 = warp_sum(m2); m3 = warp_sum(m3); m4 = warp_sum(m4);
        if (lane == 0) {
            atomicAdd(a.moments + 0, m1); atomicAdd(a.moments + 1, m2);
            atomicAdd(a.moments + 2, m3); atomicAdd(a.moments + 3, m4);
        }
    }
    if (a.adc_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_hist; i += blockDim.x) {
            const unsigned int c = hist_s[i];
            if (c) atomicAdd(a.adc_hist + i, (unsigned long long)c);
        }
    }
}

// K0: nothing but the stores K1 issues.  16 bytes per thread per iteration, grid-stride.
__global__ void __launch_bounds__(256)
k_write_calibrate(int4* __restrict__ dst, size_t n_vec, int16_t* __restrict__ tail, size_t n_tail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < n_vec; i += stride) {
        const int v = (int)(i & 0x0fff);
        __stcs(dst + i, make_int4(v, v, v, v));
    }
    if (i0 < n_tail) tail[i0] = (int16_t)i0;
}

}  // namespace dct
