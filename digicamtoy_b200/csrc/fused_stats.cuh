// K3 fused into the generation kernels: the validation statistics of a CTA's traces, taken from the
// packed int16 rows in shared memory between digitisation and the store -- the cube is never read
// back from HBM.  Same quantities, bit for bit, as k_stats_tile / k_stats (stats.cuh), which stay
// as the checker and for cubes that already exist:
//   ADC histogram, per-bin and per-pixel sum and sum of squares, integrated-charge histogram, and
//   the raw power sums behind compute_moments_nsb (reference nsb_baseline.py:21-32).
//
// Per CTA (NT traces): thread-private 16-bit histogram columns for the FS_WIN values around the
// pedestal (no atomics), a CTA histogram of FS_HIST values for the rest (shared-memory atomics),
// global atomics beyond that (bright pulses; the two saturation rails are counted in registers);
// per-bin sums by a second pass with lane = word column; everything is flushed once per CTA with
// fire-and-forget global reductions (~500 per 128 traces).
#pragma once
#include "device_params.cuh"

#ifndef DCT_FUSED_STATS
#define DCT_FUSED_STATS 1      // 0: compile the fused statistics out of the generation kernels (timing experiments)
#endif

namespace dct {

constexpr int FS_WIN = 16;         // values per thread-private column set
constexpr int FS_HIST = 256;       // values of the CTA histogram

struct FusedStats {
    unsigned long long* adc_hist = nullptr;     // nullptr: statistics off
    double* bin_sum = nullptr;
    double* bin_sumsq = nullptr;
    double* pixel_sum = nullptr;
    double* pixel_sumsq = nullptr;
    unsigned long long* charge_hist = nullptr;  // optional
    double* moments = nullptr;                  // optional: raw power sums 1..4
    int charge_lo = 0, n_charge_bins = 0;
    int win_lo = 0;                             // first value of the private window
    int hist_lo = 0;                            // first value of the CTA histogram
};

// bytes of dynamic shared memory behind the rows: private columns, CTA histogram, per-bin sums
__host__ __device__ inline size_t fused_stats_smem(int nt, int B) {
    return (size_t)nt * FS_WIN * 2 + (size_t)FS_HIST * 4 + (size_t)B * 16;
}

struct FsOutside {
    double m[4];
    unsigned int rail_hi, rail_lo;
};

// a sample beyond the CTA histogram: global histogram (rails: registers) and its power sums
__device__ __noinline__ void fs_outside(FsOutside& o, unsigned long long* adc_hist, int adc_min, int adc_max,
                                        int v, unsigned int c, bool defer_rails) {
    if (defer_rails && v == adc_max) { o.rail_hi += c; return; }
    if (defer_rails && v == adc_min) { o.rail_lo += c; return; }
    const unsigned int hg = (unsigned int)(v - adc_min);
    if (hg <= (unsigned int)(adc_max - adc_min)) atomicAdd(adc_hist + hg, (unsigned long long)c);
    const double x = (double)v, cd = (double)c;
    o.m[0] += cd * x; o.m[1] += cd * x * x; o.m[2] += cd * x * x * x; o.m[3] += cd * x * x * x * x;
}

// Called by NT threads (tid = threadIdx.x < NT, thread = trace) after every row of the CTA holds
// its packed int16 samples (sample b -> half-word b of the row) and a barrier has made them
// visible.  BAR = named barrier of exactly these NT threads.  `smem` = fused_stats_smem(NT, B) bytes.
// Out of line on purpose: inlined, its registers and code size cost the generation kernels 1-8 % even
// with the statistics switched off (measured on the grid and sweep kernels).
template <int NT, int BAR>
__device__ __noinline__ void cta_fused_stats(const FusedStats s, const int B, const int P, const int adc_min,
                                             const int adc_max, const float* baseline, uint64_t trace0,
                                             uint64_t n_traces, const float* rows, unsigned char* smem) {
    // (everything by value: a reference to the kernel's parameter block would move it to local memory)
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int SB = B | 1;
    unsigned short* priv_all = reinterpret_cast<unsigned short*>(smem);                   // [NT/32][FS_WIN][32]
    unsigned int* hist_s = reinterpret_cast<unsigned int*>(smem + (size_t)NT * FS_WIN * 2);   // [FS_HIST]
    unsigned long long* bin_s = reinterpret_cast<unsigned long long*>(hist_s + FS_HIST);   // [B] sums, [B] sums of squares
    auto barrier = [] { asm volatile("bar.sync %0, %1;" :: "n"(BAR), "n"(NT) : "memory"); };

    {   // clear
        unsigned int* w = reinterpret_cast<unsigned int*>(smem);
        const int n_words = (int)(fused_stats_smem(NT, B) / 4);
        for (int i = tid; i < n_words; i += NT) w[i] = 0u;
    }
    barrier();

    const uint64_t left = n_traces - trace0;
    const int n_live = (int)(left < (uint64_t)NT ? left : (uint64_t)NT);
    const bool mine = tid < n_live;
    unsigned short* priv = priv_all + wib * (FS_WIN * 32) + lane;
    const unsigned int hist_addr = (unsigned int)__cvta_generic_to_shared(hist_s);
    FsOutside outside = {{0, 0, 0, 0}, 0u, 0u};
    unsigned int n_charge_over = 0, n_charge_under = 0;

    if (mine) {
        const unsigned short* hw = reinterpret_cast<const unsigned short*>(rows + (size_t)tid * SB);
        long long ts = 0;
        unsigned long long tq = 0;
        bool beyond = false;
        for (int b = 0; b < B; ++b) {
            const int v = (int)(short)hw[b];
            ts += v;
            tq += (unsigned int)(v * v);
            const unsigned int h = (unsigned int)(v - s.win_lo);
            if (h < (unsigned int)FS_WIN) {
                priv[h * 32] += 1;
            } else {
                const unsigned int hs = (unsigned int)(v - s.hist_lo);
                if (hs < (unsigned int)FS_HIST) asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(hist_addr + 4u * hs) : "memory");
                else beyond = true;
            }
        }
        if (beyond)                                   // bright traces only: second scan for those samples
            for (int b = 0; b < B; ++b) {
                const int v = (int)(short)hw[b];
                if ((unsigned int)(v - s.win_lo) >= (unsigned int)FS_WIN && (unsigned int)(v - s.hist_lo) >= (unsigned int)FS_HIST)
                    fs_outside(outside, s.adc_hist, adc_min, adc_max, v, 1u, true);
            }
        const uint64_t g = trace0 + tid;
        const int pixel = (int)(g % (uint64_t)P);
        atomicAdd(s.pixel_sum + pixel, (double)ts);
        atomicAdd(s.pixel_sumsq + pixel, (double)tq);
        if (s.charge_hist) {
            const long long bin = ts - (long long)B * llrintf(baseline[pixel]) - s.charge_lo;
            if (bin >= s.n_charge_bins - 1) ++n_charge_over;
            else if (bin <= 0) ++n_charge_under;
            else atomicAdd(s.charge_hist + bin, 1ull);
        }
    }

    // per-bin sums: lane = bin column, this warp's 32 rows; then one shared-memory add per (warp, bin)
    {
        const int r_begin = 32 * wib, r_end = min(n_live, 32 * (wib + 1));
        const unsigned short* hw_all = reinterpret_cast<const unsigned short*>(rows);
        for (int b = lane; b < B; b += 32) {
            long long cs = 0;
            unsigned long long cq = 0;
            for (int r = r_begin; r < r_end; ++r) {
                const int v = (int)(short)hw_all[(size_t)r * SB * 2 + b];
                cs += v;
                cq += (unsigned int)(v * v);
            }
            if (r_begin < r_end) {
                atomicAdd(bin_s + b, (unsigned long long)cs);            // two's complement: negative sums add up correctly
                atomicAdd(bin_s + B + b, cq);
            }
        }
    }
    // fold this warp's private columns into the CTA histogram / beyond it
    __syncwarp();
    for (int w = 0; w < FS_WIN; ++w) {
        const unsigned int c = __reduce_add_sync(0xffffffffu, (unsigned int)priv[w * 32]);
        if (lane == 0 && c) {
            const int v = s.win_lo + w;
            const unsigned int hs = (unsigned int)(v - s.hist_lo);
            if (hs < (unsigned int)FS_HIST) atomicAdd(&hist_s[hs], c);
            else fs_outside(outside, s.adc_hist, adc_min, adc_max, v, c, false);
        }
    }
    if (s.charge_hist) {
        n_charge_over = __reduce_add_sync(0xffffffffu, n_charge_over);
        n_charge_under = __reduce_add_sync(0xffffffffu, n_charge_under);
        if (lane == 0 && n_charge_over) atomicAdd(s.charge_hist + (s.n_charge_bins - 1), (unsigned long long)n_charge_over);
        if (lane == 0 && n_charge_under) atomicAdd(s.charge_hist, (unsigned long long)n_charge_under);
    }
    {
        const unsigned int hi = __reduce_add_sync(0xffffffffu, outside.rail_hi);
        const unsigned int lo = __reduce_add_sync(0xffffffffu, outside.rail_lo);
        if (lane == 0 && hi) fs_outside(outside, s.adc_hist, adc_min, adc_max, adc_max, hi, false);
        if (lane == 0 && lo) fs_outside(outside, s.adc_hist, adc_min, adc_max, adc_min, lo, false);
    }
    barrier();

    // flush: CTA histogram (+ its power sums), per-bin sums
    double m[4] = {outside.m[0], outside.m[1], outside.m[2], outside.m[3]};
    const int n_hist = adc_max - adc_min + 1;
    for (int i = tid; i < FS_HIST; i += NT) {
        const unsigned int c = hist_s[i];
        const int gidx = i + s.hist_lo - adc_min;
        if (c && gidx >= 0 && gidx < n_hist) {
            atomicAdd(s.adc_hist + gidx, (unsigned long long)c);
            const double x = (double)(i + s.hist_lo), cd = (double)c;
            m[0] += cd * x; m[1] += cd * x * x; m[2] += cd * x * x * x; m[3] += cd * x * x * x * x;
        }
    }
    for (int b = tid; b < B; b += NT) {
        atomicAdd(s.bin_sum + b, (double)(long long)bin_s[b]);
        atomicAdd(s.bin_sumsq + b, (double)bin_s[B + b]);
    }
    if (s.moments) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = m[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v != 0.0) atomicAdd(s.moments + k, v);
        }
    }
}

}  // namespace dct
