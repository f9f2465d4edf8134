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

// Fast path (even B <= 128, 16-byte aligned tiles, |adc| <= 11585).
//
// History (profiles/r01): v1 (warp = trace, shared-memory atomics + float64 power sums per sample)
// took longer than K1 on the dark run; v2/v3 (lane-private histogram columns, per-pixel work items)
// removed the atomics but still spent ~225 warp-instructions per trace because a warp only holds
// 2-4 samples of a trace per lane.  This version turns the mapping around:
//   * thread = trace.  A CTA owns a block of 128 consecutive pixels and walks a slice of events;
//     each event's 128 rows are one contiguous run of the cube, brought into shared memory with
//     16-byte cp.async (one buffer in the compact layout, else two) and then read row-wise.
//   * per sample: a plain 16-bit read-modify-write of the thread-private histogram column (window of
//     [wide cubes -- NSB pedestal spread far beyond the window -- get win_lo out of reach from the host:
//      every sample then takes the CTA-histogram path and the warp does not split, ncu_k3_c4_r02.txt]
//     STATS_WIN (16) values around the pedestal; the rare sample outside goes to the CTA histogram with
//     an atomic), and integer sums for the per-trace / per-pixel quantities.
//   * per bin: a second pass over the tile with lane = word column, accumulated in 64-bit registers.
//   * per-pixel sums live in registers for the whole slice; raw power sums are derived from the
//     CTA histogram when it is flushed.
#ifndef DCT_STATS_WIN
#define DCT_STATS_WIN 16      // measured (profiles/r01/k3_compact_variants.txt): 64 -> 16 values frees 12 KB per CTA;
#endif                        // the extra resident CTAs pay more than the private columns lose (dark 3.36 -> 2.45 ms)
constexpr int STATS_WIN = DCT_STATS_WIN;
constexpr int STATS_ROWS = 128;               // traces per tile = threads per CTA
#ifndef DCT_STATS_CHARGE_WIN
#define DCT_STATS_CHARGE_WIN 1024
#endif
constexpr int STATS_CHARGE_WIN = DCT_STATS_CHARGE_WIN;   // CTA-private window of the charge histogram
#ifndef DCT_STATS_SETS
#define DCT_STATS_SETS 1      // 2: separate column sets for the low and the high half-word of a cube word (two independent
#endif                        //    read-modify-write chains per thread)
constexpr int STATS_SETS = DCT_STATS_SETS;
constexpr int STATS_PRIV_COLS = (STATS_ROWS / 32) * STATS_WIN * STATS_SETS;    // (set, warp, value) columns
constexpr int STATS_PRIV_WORDS = STATS_PRIV_COLS * 16;                         // 32-bit words of the columns
// (two independent column sets -- one per half-word -- were measured: the shorter dependency chain
//  did not pay for the occupancy lost to the extra 16 KB: 3.2 ms instead of 2.4 ms per C4 cube)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void red_shared_inc(unsigned int smem_addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_addr) : "memory");
}

// Samples beyond the CTA histogram's slice (compact layout only, bright pulses only).  Out of line and
// with its accumulators in local memory: the per-sample loop must not pay registers for it.
struct HistOutside {
    double m[4];                    // power sums of those samples
    unsigned int rail_hi, rail_lo;  // those at adc_max / adc_min: one global address each, so counted here
};
__device__ __noinline__ void hist_outside(HistOutside& o, unsigned long long* adc_hist, int adc_min, int adc_max,
                                          int v, unsigned int c, bool defer_rails) {
    if (defer_rails && v == adc_max) { o.rail_hi += c; return; }
    if (defer_rails && v == adc_min) { o.rail_lo += c; return; }
    const unsigned int hg = (unsigned int)(v - adc_min);
    if (hg <= (unsigned int)(adc_max - adc_min)) atomicAdd(adc_hist + hg, (unsigned long long)c);
    const double x = (double)v, cd = (double)c;
    o.m[0] += cd * x; o.m[1] += cd * x * x; o.m[2] += cd * x * x * x; o.m[3] += cd * x * x * x * x;
}

__global__ void __launch_bounds__(STATS_ROWS)
k_stats_tile(const StatsArgs a, int win_lo, int n_slices, int charge_win_lo,
             int hist_lo, int n_hist_s, int single_tile, int n_charge_s) {
    // n_charge_s: size of the CTA window of the charge histogram (<= STATS_CHARGE_WIN; the host picks a
    // small one for quiet cubes: less shared memory, one more CTA per SM)
    // hist_lo / n_hist_s: the slice [hist_lo, hist_lo + n_hist_s) of the ADC range that the CTA
    // histogram covers (the whole range for short traces; for long ones a window, so that more CTAs
    // fit on an SM -- samples outside it go straight to the global histogram).  single_tile: one
    // tile buffer instead of two, the other resident CTAs hide the load.
    extern __shared__ __align__(16) unsigned int smem_u[];
    const int n_hist = a.adc_max - a.adc_min + 1;
    const int wpr = a.B >> 1;                                   // 32-bit words per trace
    const int tile_words = (STATS_ROWS * wpr + 3) & ~3;
    unsigned int* tile0 = smem_u;
    unsigned int* tile1 = single_tile ? tile0 : smem_u + tile_words;
    unsigned int* hist_s = tile1 + tile_words;                  // [n_hist_s]
    unsigned int* charge_s = hist_s + n_hist_s;                 // [STATS_CHARGE_WIN]
    unsigned short* priv_all = reinterpret_cast<unsigned short*>(charge_s + n_charge_s);
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    unsigned short* priv = priv_all + wib * (STATS_WIN * 32) + lane;      // my column: priv[w * 32]
    unsigned short* priv_hi = priv + (STATS_SETS - 1) * (STATS_ROWS / 32) * STATS_WIN * 32;   // second set (or the same)
    for (int i = tid; i < n_hist_s + n_charge_s + STATS_PRIV_WORDS; i += STATS_ROWS)
        hist_s[i] = 0u;
    __syncthreads();

    const uint64_t n_events = a.n_traces / (uint64_t)a.P;
    const int n_pb = (a.P + STATS_ROWS - 1) / STATS_ROWS;
    const uint64_t n_items = (uint64_t)n_pb * n_slices;
    const uint64_t event_words = (uint64_t)a.P * wpr;
    const uint32_t* cube = reinterpret_cast<const uint32_t*>(a.adc);
    long long bs[4] = {0, 0, 0, 0}, bq[4] = {0, 0, 0, 0};       // bins 2*lane, 2*lane+1, 2*(lane+32), +1
    const unsigned int hist_addr = (unsigned int)__cvta_generic_to_shared(hist_s);
    HistOutside outside = {{0, 0, 0, 0}, 0u, 0u};
    // a sample outside the thread-private window: CTA histogram, or the global one beyond its slice
    // (beyond the slice the per-sample loop only raises a flag; the trace is then scanned a second
    //  time for those samples -- a call in the per-sample loop cost 10% even when never taken)
    bool beyond = false;
    auto hist_rare = [&](int v) {
        const unsigned int hs = (unsigned int)(v - hist_lo);
        if (hs < (unsigned int)n_hist_s) red_shared_inc(hist_addr + 4u * hs);
        else beyond = true;
    };
    auto hist_add = [&](int v, unsigned int c) {                 // c samples of value v (fold of the columns)
        const unsigned int hs = (unsigned int)(v - hist_lo);
        if (hs < (unsigned int)n_hist_s) atomicAdd(&hist_s[hs], c);
        else hist_outside(outside, a.adc_hist, a.adc_min, a.adc_max, v, c, false);
    };
    int my_samples = 0;
    // Every clamped charge would hit the same end bin of the global histogram: count those in
    // registers (ablation, profiles/r01/cost_attribution.txt: as atomics they cost 2.5 ms of 3.8).
    unsigned int n_charge_over = 0, n_charge_under = 0;

    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pb = (int)(item % (uint64_t)n_pb);
        const uint64_t slice = item / (uint64_t)n_pb;
        const uint64_t e_begin = slice * n_events / n_slices, e_end = (slice + 1) * n_events / n_slices;
        const int p0 = pb * STATS_ROWS;
        const int np = min(STATS_ROWS, a.P - p0);                // rows in this block
        const int n_vec = (np * wpr + 3) >> 2;                   // 16-byte pieces per tile (tail padded)
        const bool mine = tid < np;
        const int pixel = p0 + tid;
        const long long charge_ref = mine ? (long long)a.B * llrintf(a.baseline[pixel]) : 0;
        long long pix_s = 0, pix_q = 0;
        const uint32_t* src0 = cube + (uint64_t)p0 * wpr;

        if (e_begin < e_end) {
            const uint32_t* src = src0 + e_begin * event_words;
            for (int v = tid; v < n_vec; v += STATS_ROWS) cp_async16(tile0 + 4 * v, src + 4 * v);
            cp_async_commit();
        }
        for (uint64_t e = e_begin; e < e_end; ++e) {
            unsigned int* cur = ((e - e_begin) & 1) ? tile1 : tile0;
            unsigned int* nxt = ((e - e_begin) & 1) ? tile0 : tile1;
            cp_async_wait_all();
            __syncthreads();                                     // tile e landed; tile e-1 fully consumed
            if (!single_tile && e + 1 < e_end) {
                const uint32_t* src = src0 + (e + 1) * event_words;
                for (int v = tid; v < n_vec; v += STATS_ROWS) cp_async16(nxt + 4 * v, src + 4 * v);
                cp_async_commit();
            }
            const unsigned int* row = cur + tid * wpr;
            int ts = 0;
            unsigned int tq = 0;
            for (int w = 0; w < wpr; ++w) {
                const unsigned int word = mine ? row[w] : 0u;
                const int v0 = (int)(short)word, v1 = (int)word >> 16;
                const unsigned int q0 = (unsigned int)(v0 * v0), q1 = (unsigned int)(v1 * v1);
                ts += v0 + v1;
                tq += q0 + q1;
                if (a.adc_hist && mine) {
                    const unsigned h0 = (unsigned)(v0 - win_lo), h1 = (unsigned)(v1 - win_lo);
                    if (h0 < (unsigned)STATS_WIN) priv[h0 * 32] += 1; else hist_rare(v0);
                    if (h1 < (unsigned)STATS_WIN) priv_hi[h1 * 32] += 1; else hist_rare(v1);
                }
            }
            if (beyond) {                                        // compact layout, bright traces only
                for (int w = 0; w < wpr; ++w) {
                    const unsigned int word = row[w];
                    const int v[2] = {(int)(short)word, (int)word >> 16};
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if ((unsigned)(v[k] - win_lo) >= (unsigned)STATS_WIN && (unsigned)(v[k] - hist_lo) >= (unsigned)n_hist_s)
                            hist_outside(outside, a.adc_hist, a.adc_min, a.adc_max, v[k], 1u, true);
                }
                beyond = false;
            }
            if (a.bin_sum) {
                // column sums: lane = word column, this warp's 32 rows of the tile (consecutive lanes read
                // consecutive words: conflict-free); replaces 4 REDUX per word (~38 cycles each)
                const int r_end = min(np, 32 * (wib + 1));
                for (int slot = 0; slot < 2; ++slot) {
                    const int w = lane + 32 * slot;
                    if (w < wpr) {
                        int s0 = 0, s1 = 0;
                        unsigned int r0 = 0u, r1 = 0u;
                        for (int r = 32 * wib; r < r_end; ++r) {
                            const unsigned int word = cur[r * wpr + w];
                            const int v0 = (int)(short)word, v1 = (int)word >> 16;
                            s0 += v0; s1 += v1;
                            r0 += (unsigned int)(v0 * v0); r1 += (unsigned int)(v1 * v1);
                        }
                        bs[2 * slot] += s0; bs[2 * slot + 1] += s1;
                        bq[2 * slot] += r0; bq[2 * slot + 1] += r1;
                    }
                }
            }
            if (single_tile && e + 1 < e_end) {
                __syncthreads();                                 // every row and column of the tile was read
                const uint32_t* src = src0 + (e + 1) * event_words;
                for (int v = tid; v < n_vec; v += STATS_ROWS) cp_async16(tile0 + 4 * v, src + 4 * v);
                cp_async_commit();
            }
            if (mine) {
                pix_s += ts;
                pix_q += tq;
                if (a.charge_hist) {
                    // clamp first (as k_stats does), then the CTA window: with fewer than
                    // STATS_CHARGE_WIN bins the window would otherwise swallow charges past the end
                    const long long bin = (long long)ts - charge_ref - a.charge_lo;
                    const long long cw = bin - charge_win_lo;
                    if (bin >= a.n_charge_bins - 1) ++n_charge_over;
                    else if (bin <= 0) ++n_charge_under;
                    else if (cw >= 0 && cw < n_charge_s) atomicAdd(&charge_s[cw], 1u);
                    else atomicAdd(a.charge_hist + bin, 1ull);
                }
            }
            my_samples += a.B;
            if (a.adc_hist && my_samples >= 60000) {             // 16-bit columns: fold this warp's window
                __syncwarp();
                for (int w = 0; w < STATS_WIN; ++w) {
                    unsigned int mine_c = priv[w * 32];
                    priv[w * 32] = 0;
                    if (STATS_SETS > 1) { mine_c += priv_hi[w * 32]; priv_hi[w * 32] = 0; }
                    const unsigned int c = __reduce_add_sync(0xffffffffu, mine_c);
                    if (lane == 0 && c) hist_add(win_lo + w, c);
                }
                my_samples = 0;
            }
        }
        __syncthreads();                                         // last tile consumed before the next item loads
        if (mine && a.pixel_sum) {
            atomicAdd(a.pixel_sum + pixel, (double)pix_s);
            if (a.pixel_sumsq) atomicAdd(a.pixel_sumsq + pixel, (double)pix_q);
        }
    }
    if (a.charge_hist) {
        n_charge_over = __reduce_add_sync(0xffffffffu, n_charge_over);
        n_charge_under = __reduce_add_sync(0xffffffffu, n_charge_under);
        if (lane == 0 && n_charge_over) atomicAdd(a.charge_hist + (a.n_charge_bins - 1), (unsigned long long)n_charge_over);
        if (lane == 0 && n_charge_under) atomicAdd(a.charge_hist, (unsigned long long)n_charge_under);
    }
    const int bins[4] = {2 * lane, 2 * lane + 1, 2 * (lane + 32), 2 * (lane + 32) + 1};
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (bins[r] < a.B && (bs[r] | bq[r])) {
            if (a.bin_sum) atomicAdd(a.bin_sum + bins[r], (double)bs[r]);
            if (a.bin_sumsq) atomicAdd(a.bin_sumsq + bins[r], (double)bq[r]);
        }
    __syncthreads();
    if (a.adc_hist) {
        // fold every thread-private column into the CTA histogram
        for (int i = tid; i < STATS_PRIV_COLS; i += STATS_ROWS) {
            const unsigned short* col = priv_all + (size_t)i * 32;
            unsigned int c = 0;
            for (int l = 0; l < 32; ++l) c += col[(l + tid) & 31];
            if (c) hist_add(win_lo + (i % STATS_WIN), c);
        }
        __syncthreads();
    }
    if (a.charge_hist)
        for (int i = tid; i < n_charge_s; i += STATS_ROWS) {
            const unsigned int c = charge_s[i];
            if (c && charge_win_lo + i < a.n_charge_bins) atomicAdd(a.charge_hist + charge_win_lo + i, (unsigned long long)c);
        }
    // flush the CTA histogram; the raw power sums come from it (exact: integers < 2^53 per CTA)
    if (a.adc_hist) {                                            // rails seen beyond the slice
        const unsigned int hi = __reduce_add_sync(0xffffffffu, outside.rail_hi);
        const unsigned int lo = __reduce_add_sync(0xffffffffu, outside.rail_lo);
        if (lane == 0 && hi) hist_outside(outside, a.adc_hist, a.adc_min, a.adc_max, a.adc_max, hi, false);
        if (lane == 0 && lo) hist_outside(outside, a.adc_hist, a.adc_min, a.adc_max, a.adc_min, lo, false);
    }
    double m[4] = {outside.m[0], outside.m[1], outside.m[2], outside.m[3]};
    if (a.adc_hist) {
        for (int i = tid; i < n_hist_s; i += STATS_ROWS) {
            const unsigned int c = hist_s[i];
            const int g = i + hist_lo - a.adc_min;
            if (c && g >= 0 && g < n_hist) {
                atomicAdd(a.adc_hist + g, (unsigned long long)c);
                const double x = (double)(i + hist_lo), cd = (double)c;
                m[0] += cd * x; m[1] += cd * x * x; m[2] += cd * x * x * x; m[3] += cd * x * x * x * x;
            }
        }
    }
    if (a.moments && a.adc_hist) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = warp_sum(m[k]);
            if (lane == 0 && v != 0.0) atomicAdd(a.moments + k, v);
        }
    }
}

// (A warp-per-trace variant served traces longer than 64 bins until the compact layout above made the
//  tile kernel faster there too: 3.4 against 5.7 ms per 16384-event dark.yml cube,
//  profiles/r01/k3_compact_variants.txt.  It is in the history, not in the build.)

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
