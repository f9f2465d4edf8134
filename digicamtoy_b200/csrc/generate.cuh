// K1: fused stochastic trace generation -> digitisation -> coalesced int16 store.
//
// One thread owns one (event, pixel) trace.  A CTA owns NT consecutive traces of the
// [event][pixel][bin] cube, i.e. one contiguous run of NT*B int16 in HBM, which it writes with
// 16-byte streaming stores gathered from shared memory.
//
// Stages per trace (reference digicamtoy/tracegenerator/__init__.py, `next` :201-226):
//   NSB photo-electrons    generate_nsb :240-257      Poisson process by exponential gaps
//   crosstalk              generate_crosstalk :297-327 Borel cluster size per primary
//   gain smearing          generate_photon_smearing :329-348
//   pulse shaping          compute_analog_signal :350-362   only taps that land in kept bins
//   signal pulses          add_signal_photon :228-238
//   electronic noise       generate_electronic_smearing :364-371  kept bins only
//   digitisation           convert_to_digital :373-381  + 12-bit saturation (new)
//
// Shared memory: every trace has a private row of SB = (B | 1) floats (odd stride: lanes that touch
// the same bin hit 32 different banks).  The analog sums accumulate there; digitisation then packs
// the int16 result into the front of the same row, so no second staging buffer is needed.
#pragma once
#include "device_params.cuh"
#include "fused_stats.cuh"
#include "sampling.cuh"

namespace dct {

constexpr int GEN_THREADS = 128;        // default traces per CTA; 32 for very long traces

__host__ __device__ inline int row_stride(int B) { return B | 1; }

__host__ __device__ inline size_t gen_smem_bytes(int B, int n_table_entries, int nt) {
    return (size_t)n_table_entries * 16 + (size_t)nt * row_stride(B) * sizeof(float);
}

// shaping arithmetic shared by every float32 path; explicit intrinsics so that no two kernels
// contract it differently (the sweep and scatter kernels must agree bit for bit).
__device__ __forceinline__ float horner(const float4 c, float u) {
    return __fmaf_rn(__fmaf_rn(__fmaf_rn(c.w, u, c.z), u, c.y), u, c.x);
}

// Template value at y in [0, L] (inclusive both ends, like interp1d with fill_value=0,
// reference :141-143), any table; 0 outside.
__device__ __forceinline__ float template_at(const DevParams& p, const float4* coef, float y) {
    if (!(y >= 0.0f && y <= p.L)) return 0.0f;
    int m = (int)(y * p.inv_knot_dx);
    m = min(m, p.n_iv - 1);
    const float u = y - (float)m * p.knot_dx;
    return horner(coef[m], u);
}

// Digitise one sample: analog sum -> LSB:  clamp(rint_half_even(acc*gain + noise + baseline)).
// (The reference adds in the order :362, :371, :379 in float64; here two FFMA in float32 -- the
// stochastic kernels only have to agree with each other bit for bit, which they do because both
// call this function; the replay kernels keep the reference's operation order.)
// Rounding = float add of 1.5*2^23 (round-to-nearest-even in hardware, FMA pipe), no F2I (XU pipe).
__device__ __forceinline__ int digitise(float acc, float gain, float z, float sigma_e, float baseline,
                                        float lo, float hi) {
    float v = __fmaf_rn(acc, gain, __fmaf_rn(z, sigma_e, baseline));
    v = fminf(fmaxf(v, lo), hi);
    return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

struct GenArgs {
    DevParams p;
    uint32_t k0, k1;            // seed
    RoundKeys rk;               // its Philox key schedule (make_round_keys(k0, k1))
    uint64_t first_event;
    uint64_t n_traces;          // n_events * P
    int16_t* adc;
    float* sig_time;
    float* sig_charge;
    int table_in_smem;          // generic coef32 table staged in shared memory
    float* analog = nullptr;    // test hook: [n_traces, B] analog sums (p.e.-weighted template) before digitisation
    FusedStats st;              // on-device statistics taken before the store (st.adc_hist == nullptr: off)
    int pack12 = 0;             // host transport: `adc` is a byte stream of 12-bit samples (store_run_packed12)
};

// dynamic shared memory of a generation kernel: its own `base` bytes, then (statistics on) the
// fused-statistics scratch on a 16-byte boundary
__host__ __device__ inline size_t stats_offset(size_t base) { return (base + 15) & ~(size_t)15; }
inline size_t smem_with_stats(size_t base, const GenArgs& a, int nt) {
    return a.st.adc_hist ? stats_offset(base) + fused_stats_smem(nt, a.p.B) : base;
}

// test hook of every generation kernel: the analog sums a trace is digitised from
__device__ __forceinline__ void dump_analog(const GenArgs& a, uint64_t g, const float* row) {
    if (a.analog)
        for (int b = 0; b < a.p.B; ++b) a.analog[g * (uint64_t)a.p.B + b] = row[b];
}

// Identity of the trace a thread owns.
struct TraceId {
    uint64_t g;          // trace index within this launch
    int pixel;
    StreamKey key;
};

__device__ __forceinline__ TraceId trace_id(const GenArgs& a, uint64_t g) {
    TraceId t;
    t.g = g;
    const uint64_t ev_local = g / (uint64_t)a.p.P;
    t.pixel = (int)(g - ev_local * (uint64_t)a.p.P);
    const uint64_t ev = a.first_event + ev_local;
    t.key.k0 = a.k0; t.key.k1 = a.k1;
    t.key.ev_lo = (uint32_t)ev; t.key.ev_hi = (uint32_t)(ev >> 32);
    t.key.pixel = (uint32_t)t.pixel;
    return t;
}

// Moves the onset (cell*dt + off) forward by one exponential gap; returns how many cell
// boundaries were crossed.  Shared by both kernels so that they place every pe identically.
// The common cases (0 or 1 crossing) are predicated; long gaps take the floor path.
__device__ __forceinline__ int advance_onset(const DevParams& p, float gap, int& cell, float& off) {
    off += gap;
    int k = 0;
    if (off >= p.dt) {
        off -= p.dt;
        k = 1;
        if (off >= p.dt) {
            const float kf = floorf(off * p.inv_dt);
            off = fmaxf(__fmaf_rn(-kf, p.dt, off), 0.0f);
            k += (int)kf;
            if (off >= p.dt) { off -= p.dt; ++k; }
        }
    }
    cell += k;
    return k;
}

// Phase f (knot interval of the first tap) and local offset u of a pe that lies x0 in (0, dt]
// before a bin.  floor() by a round-toward-zero add of 2^23: FMA pipe, no F2I / I2F.
__device__ __forceinline__ void phase_of(const DevParams& p, float x0, int& f, float& u) {
    float r = __fadd_rz(x0 * p.inv_knot_dx, 8388608.0f);
    r = fminf(r, p.phase_cap);                       // 2^23 + n_phase - 1  (x0 == dt -> last phase)
    f = __float_as_int(r) & 0x7f;
    u = __fmaf_rn(8388608.0f - r, p.knot_dx, x0);    // x0 - f * knot_dx
}

// Per-pixel constants of the NSB photo-electron stream.
struct PixelNsb {
    float neg_ln2_over_rate;   // -ln2 / rate
    float xt;
    float4 borel_cdf;
    float sigma1;              // relative gain spread
};

// Random numbers of the continuous NSB stream: four photo-electron clusters per three Philox
// blocks (pe 4g .. 4g+3 of a trace <-> blocks 3g, 3g+1, 3g+2 of STREAM_NSB):
//   block 3g     .x .y .z .w -> exponential gap before pe 4g .. 4g+3            (generate_nsb :244-250)
//   block 3g + 1 .x .y .z .w -> Borel cluster size of pe 4g .. 4g+3            (crosstalk :297-327)
//   block 3g + 2 (.x, .y)    -> Box-Muller (radius, angle): cos -> pe 4g, sin -> pe 4g+1
//                (.z, .w)    -> the same for pe 4g+2 and 4g+3                 (gain smearing :329-348)
// Both normals of a Box-Muller pair are used (they are independent), so a cluster costs three random
// words instead of four, and half a logarithm / square root instead of one.
struct NsbGroup { u32x4 gap, bor, nrm; };

__device__ __forceinline__ NsbGroup draw_nsb_group(const RoundKeys& rk, const StreamKey& key, uint32_t g) {
    NsbGroup n;
    n.gap = draw_block_rk(rk, key, STREAM_NSB, 3u * g);
    n.bor = draw_block_rk(rk, key, STREAM_NSB, 3u * g + 1u);
    n.nrm = draw_block_rk(rk, key, STREAM_NSB, 3u * g + 2u);
    return n;
}

// Amplitude of a cluster of fn pe after gain smearing (:329-348):  A = n + sigma_1 sqrt(n) z.
// One definition for every kernel: they must agree bit for bit.
__device__ __forceinline__ float cluster_amplitude(float fn, float z, float sigma1) {
    return __fmaf_rn(sigma1, fast_sqrt(fn) * z, fn);
}

__device__ __forceinline__ float smeared_amplitude(int n, float z, float sigma1) {
    const float fn = (float)n;
    return __fmaf_rn(z * sigma1, fast_sqrt(fn), fn);
}

// On-grid NSB mode (sub_binning > 0, reference :259-267 + :306-312 + :329-348), one simulated bin, fast
// form: the total T of primaries and crosstalk progeny from ONE uniform through the pixel's
// generalised-Poisson table (sampling.cuh), then gain smearing.  Two bins share a Philox block of
// STREAM_SUBBIN_GP -- block bs / 2:  .x -> T of the even bin, .y -> T of the odd bin, (.z, .w) ->
// Box-Muller pair, cosine for the even and sine for the odd bin.  `r` is that block (callers that
// walk the bins in order draw it once per pair).  Returns false for an empty bin.
// cdf / stride: the pixel's table (global: stride 1; the grid kernel's shared copy: stride NT).
__device__ __forceinline__ bool subbin_total(const u32x4& r, uint32_t bs, float lam, const PixelNsb& q,
                                             const float* cdf, int stride, float& A) {
    const uint32_t w = (bs & 1u) ? r.y : r.x;
    const int total = gp_table_draw(u01_open1(w), cdf, stride, lam, q.xt);
    if (total == 0) return false;
    A = (float)total;
    if (q.sigma1 > 0.0f) {
        float z0, z1;
        normal_pair(r.z, r.w, z0, z1);
        A = smeared_amplitude(total, (bs & 1u) ? z1 : z0, q.sigma1);
    }
    return true;
}

// On-grid NSB mode (sub_binning > 0, reference :259-267 + :306-312 + :329-348), one simulated bin:
// n ~ Poisson(lam) pe on the sample, their crosstalk progeny, gain smearing.  Returns false for an
// empty bin.  Everything comes from Philox block `bs` of STREAM_SUBBIN in the common case, so the
// lanes of a warp stay in step:  .x -> n,  .y -> first crosstalk generation,  .z/.w -> the normal,
// and the 27 bits the uniforms do not use (low 9 of .x, .z, .w) -> second generation.  A cluster
// that is still alive then (2 % of the non-empty bins at 1 GHz, xt = 0.08) takes generations 3-6
// from block (bs + 1) << 16, and whatever is left goes to subbin_overflow.   p0 = exp(-lam).
// cdf (optional): this trace's poisson_table_fill(lam) with stride cdf_stride -- same counts, no loop.
__device__ __forceinline__ bool subbin_cluster(const StreamKey& key, uint32_t bs, float lam, float p0,
                                               const PixelNsb& q, float& A,
                                               const float* cdf = nullptr, int cdf_stride = 1) {
    const u32x4 r = draw_block(key, STREAM_SUBBIN, bs);
    int total = 0, gen = -1;
    uint32_t overflow_block = (bs + 1u) << 16;
    if (lam < 12.0f) {
        total = gen = cdf ? poisson_table_draw(u01_open1(r.x), cdf, cdf_stride, lam, p0)
                          : poisson_inv(u01_open1(r.x), lam, p0);
        if (total == 0) return false;
        if (q.xt > 0.0f) {
            auto next_generation = [&](uint32_t bits) {          // no-op once the cluster has died out
                const float l = q.xt * (float)gen;
                if (gen > 0 && l < 12.0f) {
                    gen = poisson_inv(u01_open1(bits), l, __expf(-l));
                    total += gen;
                }
            };
            if (q.xt * (float)gen < 12.0f) {
                next_generation(r.y);
                next_generation((r.x << 23) | ((r.z & 0x1ffu) << 14) | ((r.w & 0x1ffu) << 5));
                if (gen > 0 && q.xt * (float)gen < 12.0f) {
                    const u32x4 e = draw_block(key, STREAM_SUBBIN, overflow_block++);
                    next_generation(e.x);
                    next_generation(e.y);
                    next_generation(e.z);
                    next_generation(e.w);
                }
            }
        } else {
            gen = 0;
        }
    }
    if (gen != 0) {
        total = subbin_overflow(key.k0, key.k1, key.ev_lo, key.ev_hi, key.pixel, overflow_block, lam, q.xt, total, gen);
        if (total == 0) return false;
    }
    A = (float)total;
    if (q.sigma1 > 0.0f) A = smeared_amplitude(total, normal_one(r.z, r.w), q.sigma1);
    return true;
}

// Generic shaping of one cluster: bin `first_bin` (simulated index) sees template argument y0, the
// next one y0 + dt, ...  Inclusive at y == L.  Used for signal pulses (which may sit exactly on a
// sample) and for templates whose knot spacing does not divide the sampling period.
__device__ __forceinline__ void scatter_generic(const DevParams& p, const float4* coef, float* row,
                                                int first_bin, float y0, float A) {
    const int kb = first_bin - p.n_drop;
    for (int j = max(0, -kb);; ++j) {
        const float y = __fmaf_rn((float)j, p.dt, y0);
        if (y > p.L || kb + j >= p.B) break;
        row[kb + j] = __fmaf_rn(A, template_at(p, coef, y), row[kb + j]);
    }
}

// Signal pulses of one trace (add_signal_photon :228-238 + their crosstalk :319-324 + smearing),
// shaped into the accumulators; optional truth outputs (cherenkov_time / cherenkov_photon).
__device__ __forceinline__ void add_signal_pulses(const GenArgs& a, const TraceId& t, const float4* coef,
                                               float* row, float xt, float sigma1) {
    const DevParams& p = a.p;
    for (int k = 0; k < p.K; ++k) {
        const float npe = p.npe[t.pixel * p.K + k];
        const float jit = p.jitter[t.pixel * p.K + k];
        float ts = p.tsig[t.pixel * p.K + k];
        float A = 0.0f;
        const u32x4 r = draw_block(t.key, STREAM_SIGNAL, (uint32_t)k);
        if (jit > 0.0f) ts += (u01_open1(r.x) - 0.5f) * jit;             // U(-j/2, j/2)  (:233)
        if (npe > 0.0f) {
            int total;
            if (p.sig_cdf != nullptr && p.poisson && npe < GP_LAM_MAX) {
                // small pulses (every level of dark.yml, n_photon 1 and 10 of ac_dc.yml): primaries and
                // crosstalk progeny in one inversion of the block's fourth word through the pulse's
                // generalised-Poisson table -- same law as Poisson(n_photon) + one tree per primary
                total = gp_table_scan(u01_open1(r.w), p.sig_cdf + (size_t)(t.pixel * p.K + k) * GP_TABLE, npe, xt);
            } else {
                BlockReader rd(t.key, STREAM_BRANCH, (uint32_t)k << 20);
                const int n0 = p.poisson ? poisson(rd, npe) : (int)rintf(npe);
                total = branching_total(rd, n0, xt);
            }
            A = (float)total;
            if (total > 0 && sigma1 > 0.0f) A = smeared_amplitude(total, normal_one(r.y, r.z), sigma1);
        }
        if (a.sig_time) a.sig_time[t.g * p.K + k] = ts;
        if (a.sig_charge) a.sig_charge[t.g * p.K + k] = A;
        if (A != 0.0f) {
            const float s = (ts - p.t0f) + p.x_lo;                        // onset coordinate
            int b1 = (int)ceilf(s * p.inv_dt);
            float y0 = __fmaf_rn((float)b1, p.dt, -s);
            if (y0 < 0.0f) { y0 += p.dt; ++b1; }
            if (y0 >= p.dt) { y0 -= p.dt; --b1; }
            if (p.n_phase > 0) {
                // sampling period = whole number of knot steps: every tap of the pulse sits in the same
                // phase, at the same offset u of its knot interval (tap j <-> interval f + n_phase j):
                // one (f, u), then a load and four FMA per bin -- a third of scatter_generic's
                // instructions (10 % of the dark.yml kernel, profiles/r02/ncu_dark_sweep_r02.txt)
                int f; float u;
                phase_of(p, y0, f, u);
                const int kb = b1 - p.n_drop;
                const int j_hi = min(p.n_taps, p.B - kb);
                const float4* tab = p.poly + f;
                for (int j = max(0, -kb); j < j_hi; ++j)
                    row[kb + j] = __fmaf_rn(A, horner(__ldg(tab + j * p.n_phase), u), row[kb + j]);
                // The template's right end is inclusive (interp1d, reference :141-143): a bin exactly at
                // y = L belongs to the last knot interval, one past where the phase table puts it.  Not
                // a measure-zero case here: time_signal on a sample with jitter = 0 is the usual config.
                for (int j = p.n_taps - 1; j <= p.n_taps; ++j) {
                    const float y = __fmaf_rn((float)j, p.dt, y0);
                    if (y == p.L && f + p.n_phase * j >= p.n_iv && (unsigned)(kb + j) < (unsigned)p.B)
                        row[kb + j] = __fmaf_rn(A, template_at(p, coef, y), row[kb + j]);
                }
            } else {
                scatter_generic(p, coef, row, b1, y0, A);
            }
        }
    }
}

// Electronic noise (kept bins only) + digitisation of one trace, in place: the int16 results are
// packed into the front of the trace's own float row (sample b -> half-word b).  Reading runs
// ahead of writing (word b/2 <= b), so nothing unread is overwritten.
__device__ __forceinline__ void digitise_trace(const DevParams& p, const RoundKeys& rk, const TraceId& t, float* row) {
    const float gain = p.gain[t.pixel];
    const float sigma_e = p.sigma_e[t.pixel];
    const float baseline = p.baseline[t.pixel];
    const float lo = (float)p.adc_min, hi = (float)p.adc_max;
    const int B = p.B;
    uint32_t* packed = reinterpret_cast<uint32_t*>(row);
    const int n_full = B >> 2;                      // quads without a bounds check
#pragma unroll 2
    for (int q = 0; q < n_full; ++q) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (sigma_e > 0.0f) {
            const u32x4 r = draw_block_rk(rk, t.key, STREAM_NOISE, (uint32_t)q);
            normal_pair(r.x, r.y, z[0], z[1]);
            normal_pair(r.z, r.w, z[2], z[3]);
        }
        int d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = digitise(row[4 * q + i], gain, z[i], sigma_e, baseline, lo, hi);
        packed[2 * q] = (uint32_t)(d[0] & 0xffff) | ((uint32_t)d[1] << 16);
        packed[2 * q + 1] = (uint32_t)(d[2] & 0xffff) | ((uint32_t)d[3] << 16);
    }
    if (B & 3) {                                     // last, partial quad
        const int q = n_full, b = 4 * n_full;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (sigma_e > 0.0f) {
            const u32x4 r = draw_block_rk(rk, t.key, STREAM_NOISE, (uint32_t)q);
            normal_pair(r.x, r.y, z[0], z[1]);
            normal_pair(r.z, r.w, z[2], z[3]);
        }
        int d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            d[i] = (b + i < B) ? digitise(row[b + i], gain, z[i], sigma_e, baseline, lo, hi) : 0;
        packed[2 * q] = (uint32_t)(d[0] & 0xffff) | ((uint32_t)d[1] << 16);
        if (b + 2 < B) packed[2 * q + 1] = (uint32_t)(d[2] & 0xffff) | ((uint32_t)d[3] << 16);
    }
}

// Coalesced store of the CTA's contiguous run of the cube: 16-byte streaming stores gathered from
// the packed rows.  Output 32-bit word w of the run lives in row w / (B/2), word w % (B/2).
// 12-bit transport of the host pipeline (dct_generate_host): the launch's samples as one little-endian
// bit stream, sample s in bits [12 s, 12 s + 12) -- two samples per three bytes.  A CTA's run starts at
// trace0 = k NT (NT a multiple of 32), i.e. at a multiple of 48 bytes: units of 32 samples = three
// 16-byte stores; the last, partial unit goes out pair by pair.  Needs 0 <= ADC <= 4095 (the caller checks).
template <int NT>
__device__ __forceinline__ void store_run_packed12(const GenArgs& a, uint64_t trace0, const float* rows, const int tid) {
    const int B = a.p.B, SB = row_stride(B);
    const uint64_t left = a.n_traces - trace0;
    const int n_live = (int)(left < (uint64_t)NT ? left : (uint64_t)NT);
    const int n_s = n_live * B;
    unsigned char* dst = reinterpret_cast<unsigned char*>(a.adc) + (trace0 * (uint64_t)B / 2) * 3;
    const uint16_t* halves = reinterpret_cast<const uint16_t*>(rows);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(rows);
    const bool even = (B & 1) == 0;
    const float inv_B = 1.0f / (float)B;
    // samples s, s + 1 (s even) of the run as x0 | x1 << 12; a sample past the end reads as 0
    auto pair = [&](int s) -> uint32_t {
        int r = (int)(((float)s + 0.5f) * inv_B);            // exact: s < 2^16 (NT <= 256, B < 256)
        int c = s - r * B;
        if (even) {
            const uint32_t w = words[r * SB + (c >> 1)];
            return (w & 0xfffu) | ((w >> 4) & 0xfff000u);
        }
        const uint32_t x0 = halves[r * SB * 2 + c] & 0xfffu;
        if (++c == B) { c = 0; ++r; }
        const uint32_t x1 = (s + 1 < n_s) ? (halves[r * SB * 2 + c] & 0xfffu) : 0u;
        return x0 | (x1 << 12);
    };
    const int n32 = n_s >> 5;
    int4* dst4 = reinterpret_cast<int4*>(dst);
    for (int u = tid; u < n32; u += NT) {
        uint32_t w[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                         // 8 samples = four 24-bit pairs -> 3 words
            const int s = 32 * u + 8 * q;
            const uint32_t p0 = pair(s), p1 = pair(s + 2), p2 = pair(s + 4), p3 = pair(s + 6);
            w[3 * q + 0] = p0 | (p1 << 24);
            w[3 * q + 1] = (p1 >> 8) | (p2 << 16);
            w[3 * q + 2] = (p2 >> 16) | (p3 << 8);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            __stcs(dst4 + 3 * u + q, make_int4((int)w[4 * q], (int)w[4 * q + 1], (int)w[4 * q + 2], (int)w[4 * q + 3]));
    }
    for (int s = 32 * n32 + 2 * tid; s < n_s; s += 2 * NT) {
        const uint32_t p2 = pair(s);
        unsigned char* d = dst + (s >> 1) * 3;
        d[0] = (unsigned char)p2; d[1] = (unsigned char)(p2 >> 8);
        if (s + 1 < n_s) d[2] = (unsigned char)(p2 >> 16);      // (a dangling last sample ends in the low nibble of d[1])
    }
}

// (tid: index within the NT threads that store this run)
template <int NT>
__device__ __forceinline__ void store_run(const GenArgs& a, uint64_t trace0, const float* rows, const int tid) {
    if (a.pack12) { store_run_packed12<NT>(a, trace0, rows, tid); return; }
    const int B = a.p.B, SB = row_stride(B);
    const uint64_t left = a.n_traces - trace0;
    const int n_live = (int)(left < (uint64_t)NT ? left : (uint64_t)NT);
    int16_t* dst = a.adc + trace0 * (uint64_t)B;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(rows);
    if ((B & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int wpr = B >> 1;                                   // words per row
        const int n_words = n_live * wpr;
        const int n_vec = n_words >> 2;
        const float inv_wpr = 1.0f / (float)wpr;
        int4* dst4 = reinterpret_cast<int4*>(dst);
        for (int v = tid; v < n_vec; v += NT) {
            const int w0 = 4 * v;
            int r = (int)(((float)w0 + 0.5f) * inv_wpr);          // exact: w0 < 2^16
            int c = w0 - r * wpr;
            int x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = (int)words[r * SB + c];
                if (++c == wpr) { c = 0; ++r; }
            }
            __stcs(dst4 + v, make_int4(x[0], x[1], x[2], x[3]));
        }
        uint32_t* dst32 = reinterpret_cast<uint32_t*>(dst);
        for (int w = 4 * n_vec + tid; w < n_words; w += NT) {
            const int r = w / wpr;
            dst32[w] = words[r * SB + (w - r * wpr)];
        }
    } else {
        const int16_t* halves = reinterpret_cast<const int16_t*>(rows);
        const int n_el = n_live * B;
        for (int e = tid; e < n_el; e += NT) {
            const int r = e / B;
            dst[e] = halves[r * SB * 2 + (e - r * B)];
        }
    }
}

template <int NT>
__device__ __forceinline__ void store_run(const GenArgs& a, uint64_t trace0, const float* rows) {
    store_run<NT>(a, trace0, rows, (int)threadIdx.x);
}

__device__ __forceinline__ PixelNsb load_pixel_nsb(const DevParams& p, int pixel) {
    PixelNsb q;
    q.neg_ln2_over_rate = p.neg_ln2_over_rate[pixel];
    q.xt = p.xt[pixel];
    q.borel_cdf = p.borel_cdf[pixel];
    q.sigma1 = p.sigma1[pixel];
    return q;
}

template <bool POLY, int NT>
__global__ void __launch_bounds__(NT, (NT == 128 ? 5 : 1))
k_generate_scatter(const GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& p = a.p;
    const int tid = threadIdx.x;
    const int B = p.B;

    const int n_tab = POLY ? p.n_taps * p.n_phase : (a.table_in_smem ? p.n_iv : 0);
    float4* tab_s = reinterpret_cast<float4*>(smem_raw);
    float* rows = reinterpret_cast<float*>(tab_s + n_tab);
    {   // stage the template table
        const float4* src = POLY ? p.poly : p.coef32;
        for (int i = tid; i < n_tab; i += NT) tab_s[i] = src[i];
    }
    // generic coefficient table: shared copy if it fits, else straight from global (L1/L2)
    const float4* coef_g = (!POLY && a.table_in_smem) ? tab_s : p.coef32;

    float* row = rows + tid * row_stride(B);
    for (int b = 0; b < B; ++b) row[b] = 0.0f;
    __syncthreads();

    const uint64_t trace0 = (uint64_t)blockIdx.x * NT;
    if (trace0 + tid < a.n_traces) {
        const TraceId t = trace_id(a, trace0 + tid);
        const float rate = p.rate[t.pixel];
        const PixelNsb q = load_pixel_nsb(p, t.pixel);

        if (rate > 0.0f) {
            if (p.sub_binning <= 0) {
                // Poisson process on [t0, t_end): exponential gaps.  Same law as the reference's
                // Poisson count + iid uniform times (:244-250) without padding to max_photon.
                float elapsed = 0.0f;
                int cell = p.cell0;
                float off = p.off0;                                   // onset = cell*dt + off
                for (uint32_t g = 0;; ++g) {
                    const NsbGroup grp = draw_nsb_group(a.rk, t.key, g);
                    const uint32_t gw[4] = {grp.gap.x, grp.gap.y, grp.gap.z, grp.gap.w};
                    const uint32_t bw[4] = {grp.bor.x, grp.bor.y, grp.bor.z, grp.bor.w};
                    float z[4];
                    normal_pair(grp.nrm.x, grp.nrm.y, z[0], z[1]);
                    normal_pair(grp.nrm.z, grp.nrm.w, z[2], z[3]);
                    bool done = false;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (done) continue;
                        const float gap = exp_gap(gw[k], q.neg_ln2_over_rate);
                        elapsed += gap;
                        if (elapsed >= p.span) { done = true; continue; }
                        advance_onset(p, gap, cell, off);
                        const float A = cluster_amplitude((float)borel(bw[k], q.xt, q.borel_cdf), z[k], q.sigma1);
                        const float x0 = p.dt - off;                      // (0, dt]
                        if (POLY) {
                            int f; float u;
                            phase_of(p, x0, f, u);
                            const int kb = cell + 1 - p.n_drop;           // kept index of tap 0
                            const int j_hi = min(p.n_taps, B - kb);
                            for (int j = max(0, -kb); j < j_hi; ++j)
                                row[kb + j] = __fmaf_rn(A, horner(tab_s[j * p.n_phase + f], u), row[kb + j]);
                        } else {
                            // bin cell+1 sees y = x0.  A pe exactly on a sample is also seen by that
                            // sample at y = 0: probability ~2^-24 per pe, ignored by both kernels.
                            scatter_generic(p, coef_g, row, cell + 1, x0, A);
                        }
                    }
                    if (done) break;
                }
            } else {
                // on-grid NSB (:259-267): n ~ Poisson(rate dt) pe exactly on every simulated bin
                const float lam = rate * p.dt, p0 = __expf(-lam);
                const int n_sim = p.n_drop + B;
                const bool gp = lam < GP_LAM_MAX && p.gp_cdf != nullptr;
                for (int bs = 0; bs < n_sim; ++bs) {
                    float A;
                    if (gp) {
                        const u32x4 r = draw_block_rk(a.rk, t.key, STREAM_SUBBIN_GP, (uint32_t)bs >> 1);
                        if (!subbin_total(r, (uint32_t)bs, lam, q, p.gp_cdf + (size_t)t.pixel * GP_TABLE, 1, A)) continue;
                    } else if (!subbin_cluster(t.key, (uint32_t)bs, lam, p0, q, A)) {
                        continue;
                    }
                    for (int m = 0; m < p.grid_n; ++m) {
                        const int kb = bs + p.grid_m_lo + m - p.n_drop;
                        if (kb >= 0 && kb < B) row[kb] = __fmaf_rn(A, p.grid_tap[m], row[kb]);
                    }
                }
            }
        }
        add_signal_pulses(a, t, POLY ? p.coef32 : coef_g, row, q.xt, q.sigma1);
        dump_analog(a, t.g, row);
        digitise_trace(p, a.rk, t, row);
    }
    __syncthreads();
    if (DCT_FUSED_STATS && a.st.adc_hist)
        cta_fused_stats<NT, 1>(a.st, p.B, p.P, p.adc_min, p.adc_max, p.baseline, trace0, a.n_traces, rows,
                               smem_raw + stats_offset(gen_smem_bytes(B, n_tab, NT)));
    store_run<NT>(a, trace0, rows);
}

}  // namespace dct
