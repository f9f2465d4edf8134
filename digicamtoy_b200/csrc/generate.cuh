// K1: fused stochastic trace generation -> digitisation -> coalesced int16 store.
//
// One thread owns one (event, pixel) trace.  A CTA owns NT consecutive traces of the
// [event][pixel][bin] cube, i.e. one contiguous run of NT*B int16 in HBM, which it writes with
// 16-byte stores from a shared-memory staging buffer.
//
// Stages per trace (reference digicamtoy/tracegenerator/__init__.py, `next` :201-226):
//   NSB photo-electrons    generate_nsb :240-257      Poisson process by exponential gaps
//   crosstalk              generate_crosstalk :297-327 Borel cluster size per primary
//   gain smearing          generate_photon_smearing :329-348
//   pulse shaping          compute_analog_signal :350-362   only taps that land in kept bins
//   signal pulses          add_signal_photon :228-238
//   electronic noise       generate_electronic_smearing :364-371  kept bins only
//   digitisation           convert_to_digital :373-381  + 12-bit saturation (new)
#pragma once
#include "device_params.cuh"
#include "sampling.cuh"

namespace dct {

constexpr int GEN_THREADS = 128;        // default traces per CTA; 32 for very long traces

// shaping arithmetic shared by every float32 path (stochastic kernels and replay_f32); explicit
// intrinsics so that no two kernels contract it differently.
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

// Digitise one sample: analog sum -> LSB.  Same operation order as the reference
// (x gain :362, + noise :371, + baseline :379, round half-to-even :380), then saturate.
__device__ __forceinline__ int digitise(float acc, float gain, float noise, float baseline,
                                        int lo, int hi) {
    const float v = __fadd_rn(__fadd_rn(__fmul_rn(acc, gain), noise), baseline);
    return max(lo, min(hi, __float2int_rn(v)));
}

struct GenArgs {
    DevParams p;
    uint32_t k0, k1;            // seed
    uint64_t first_event;
    uint64_t n_traces;          // n_events * P
    int16_t* adc;
    float* sig_time;
    float* sig_charge;
    int table_in_smem;          // generic coef32 table staged in shared memory
};

// Dynamic shared memory layout (floats then halves):
//   V    [B][NT] float   analog accumulators of the kept bins, thread-minor => conflict-free
//   tab  float4[...]     polyphase table (or generic table)
//   out  [NT*B] int16    dense staging of the CTA's output run
__host__ __device__ inline size_t gen_smem_bytes(int B, int n_table_entries, int nt) {
    size_t v = (size_t)B * nt * sizeof(float);
    size_t t = (size_t)n_table_entries * 16;
    size_t o = ((size_t)nt * B * sizeof(int16_t) + 15) & ~(size_t)15;
    return v + t + o;
}

// Generic version: onset y0 in [0, dt) before... i.e. bin `first_bin` sees template argument y0,
// the next one y0 + dt, ...  Inclusive at y == L.  Used for signal pulses (which may sit exactly
// on a sample) and for templates whose knot spacing does not divide the sampling period.
template <int NT>
__device__ __forceinline__ void scatter_generic(const DevParams& p, const float4* coef, float* V,
                                                int first_bin, float y0, float A) {
    const int kb = first_bin - p.n_drop;
    int j = max(0, -kb);
    for (;; ++j) {
        const float y = __fmaf_rn((float)j, p.dt, y0);
        if (y > p.L || kb + j >= p.B) break;
        const float h = template_at(p, coef, y);
        float* v = V + (kb + j) * NT;
        *v = __fmaf_rn(A, h, *v);
    }
}

__device__ __forceinline__ float smeared_amplitude(int n, float z, float sigma1) {
    // A = n + N(0, sqrt(n) sigma_1 / gain)   (:331-348); zero spread -> exactly n
    const float fn = (float)n;
    return __fmaf_rn(z * sigma1, sqrtf(fn), fn);
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
__device__ __forceinline__ int advance_onset(const DevParams& p, float gap, int& cell, float& off) {
    off += gap;
    if (off < p.dt) return 0;
    const float kf = floorf(off * p.inv_dt);
    int k = (int)kf;
    off = fmaxf(__fmaf_rn(-kf, p.dt, off), 0.0f);
    if (off >= p.dt) { off -= p.dt; ++k; }
    cell += k;
    return k;
}

// Phase (which knot interval of the first tap) and local offset of a pe that lies x0 before a bin.
__device__ __forceinline__ void phase_of(const DevParams& p, float x0, int& f, float& u) {
    f = min((int)(x0 * p.inv_knot_dx), p.n_phase - 1);
    u = x0 - (float)f * p.knot_dx;
}

// One NSB photo-electron cluster from its Philox block: amplitude after crosstalk and smearing.
__device__ __forceinline__ float nsb_amplitude(const u32x4& r, float xt, float exp_neg_xt, float sigma1) {
    const int n = borel(r.y, xt, exp_neg_xt);
    return (sigma1 > 0.0f) ? smeared_amplitude(n, normal_one(r.z, r.w), sigma1) : (float)n;
}

// Signal pulses of one trace (add_signal_photon :228-238 + their crosstalk :319-324 + smearing),
// shaped into the accumulators V; optional truth outputs (cherenkov_time / cherenkov_photon).
template <int NT>
__device__ __forceinline__ void add_signal_pulses(const GenArgs& a, const TraceId& t, const float4* coef,
                                                  float* V, float xt, float sigma1) {
    const DevParams& p = a.p;
    for (int k = 0; k < p.K; ++k) {
        const float npe = p.npe[t.pixel * p.K + k];
        const float jit = p.jitter[t.pixel * p.K + k];
        float ts = p.tsig[t.pixel * p.K + k];
        float A = 0.0f;
        const u32x4 r = draw_block(t.key, STREAM_SIGNAL, (uint32_t)k);
        if (jit > 0.0f) ts += (u01_open1(r.x) - 0.5f) * jit;             // U(-j/2, j/2)  (:233)
        if (npe > 0.0f) {
            BlockReader rd(t.key, STREAM_BRANCH, (uint32_t)k << 20);
            const int n0 = p.poisson ? poisson(rd, npe) : (int)rintf(npe);
            const int total = branching_total(rd, n0, xt);
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
            scatter_generic<NT>(p, coef, V, b1, y0, A);
        }
    }
}

// Electronic noise (kept bins only) + digitisation of one trace into its row of the staging run.
template <int NT>
__device__ __forceinline__ void digitise_trace(const DevParams& p, const TraceId& t, const float* V,
                                               int16_t* row) {
    const float gain = p.gain[t.pixel];
    const float sigma_e = p.sigma_e[t.pixel];
    const float baseline = p.baseline[t.pixel];
    const int B = p.B;
    for (int q = 0; 4 * q < B; ++q) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (sigma_e > 0.0f) {
            const u32x4 r = draw_block(t.key, STREAM_NOISE, (uint32_t)q);
            normal_pair(r.x, r.y, z[0], z[1]);
            normal_pair(r.z, r.w, z[2], z[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = 4 * q + i;
            if (b < B)
                row[b] = (int16_t)digitise(V[b * NT], gain, z[i] * sigma_e, baseline, p.adc_min, p.adc_max);
        }
    }
}

// Coalesced store of the CTA's contiguous run of the cube: 16-byte streaming stores.
template <int NT>
__device__ __forceinline__ void store_run(const GenArgs& a, uint64_t trace0, const int16_t* out_s) {
    const int tid = threadIdx.x;
    const uint64_t left = a.n_traces - trace0;
    const size_t n_el = (size_t)(left < (uint64_t)NT ? left : (uint64_t)NT) * a.p.B;
    int16_t* dst = a.adc + trace0 * (uint64_t)a.p.B;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const size_t n_vec = n_el / 8;
        const int4* src4 = reinterpret_cast<const int4*>(out_s);
        int4* dst4 = reinterpret_cast<int4*>(dst);
        for (size_t i = tid; i < n_vec; i += NT) __stcs(dst4 + i, src4[i]);
        for (size_t i = n_vec * 8 + tid; i < n_el; i += NT) dst[i] = out_s[i];
    } else {
        for (size_t i = tid; i < n_el; i += NT) dst[i] = out_s[i];
    }
}

template <bool POLY, int NT>
__global__ void __launch_bounds__(NT)
k_generate_scatter(const GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& p = a.p;
    const int tid = threadIdx.x;
    const int B = p.B;

    float* Vall = reinterpret_cast<float*>(smem_raw);
    const int n_tab = POLY ? p.n_taps * p.n_phase : (a.table_in_smem ? p.n_iv : 0);
    float4* tab_s = reinterpret_cast<float4*>(smem_raw + (size_t)B * NT * sizeof(float));
    int16_t* out_s = reinterpret_cast<int16_t*>(tab_s + n_tab);

    {   // stage the template table
        const float4* src = POLY ? p.poly : p.coef32;
        for (int i = tid; i < n_tab; i += NT) tab_s[i] = src[i];
    }
    // generic coefficient table: shared copy if it fits, else straight from global (L1/L2)
    const float4* coef_g = (!POLY && a.table_in_smem) ? tab_s : p.coef32;

    float* V = Vall + tid;
    for (int b = 0; b < B; ++b) V[b * NT] = 0.0f;
    __syncthreads();

    const uint64_t trace0 = (uint64_t)blockIdx.x * NT;
    if (trace0 + tid < a.n_traces) {
        const TraceId t = trace_id(a, trace0 + tid);
        const float rate = p.rate[t.pixel];
        const float xt = p.xt[t.pixel];
        const float exp_neg_xt = p.exp_neg_xt[t.pixel];
        const float sigma1 = p.sigma1[t.pixel];

        if (rate > 0.0f) {
            if (p.sub_binning <= 0) {
                // Poisson process on [t0, t_end): exponential gaps.  Same law as the reference's
                // Poisson count + iid uniform times (:244-250) without padding to max_photon.
                const float inv_rate = p.inv_rate[t.pixel];
                float elapsed = 0.0f;
                int cell = p.cell0;
                float off = p.off0;                                   // onset = cell*dt + off
                for (uint32_t i = 0;; ++i) {
                    const u32x4 r = draw_block(t.key, STREAM_NSB, i);
                    const float gap = exp_gap(r.x, inv_rate);
                    elapsed += gap;
                    if (elapsed >= p.span) break;
                    advance_onset(p, gap, cell, off);
                    const float A = nsb_amplitude(r, xt, exp_neg_xt, sigma1);
                    const float x0 = p.dt - off;                      // (0, dt]
                    if (POLY) {
                        int f; float u;
                        phase_of(p, x0, f, u);
                        const int kb = cell + 1 - p.n_drop;           // kept index of tap 0
                        const int j_hi = min(p.n_taps, B - kb);
                        for (int j = max(0, -kb); j < j_hi; ++j) {
                            const float h = horner(tab_s[j * p.n_phase + f], u);
                            float* v = V + (kb + j) * NT;
                            *v = __fmaf_rn(A, h, *v);
                        }
                    } else {
                        // bin cell+1 sees y = x0.  A pe exactly on a sample is also seen by that
                        // sample at y = 0: probability ~2^-24 per pe, ignored by both kernels.
                        scatter_generic<NT>(p, coef_g, V, cell + 1, x0, A);
                    }
                }
            } else {
                // on-grid NSB (:259-267): n ~ Poisson(rate dt) pe exactly on every simulated bin
                BlockReader rd(t.key, STREAM_SUBBIN, 0);
                const float lam = rate * p.dt;
                const int n_sim = p.n_drop + B;
                for (int bs = 0; bs < n_sim; ++bs) {
                    const int n = poisson(rd, lam);
                    if (n <= 0) continue;
                    const int total = branching_total(rd, n, xt);
                    float A = (float)total;
                    if (sigma1 > 0.0f) {
                        const uint32_t r0 = rd.next(), r1 = rd.next();
                        A = smeared_amplitude(total, normal_one(r0, r1), sigma1);
                    }
                    for (int m = 0; m < p.grid_n; ++m) {
                        const int kb = bs + p.grid_m_lo + m - p.n_drop;
                        if (kb >= 0 && kb < B) {
                            float* v = V + kb * NT;
                            *v = __fmaf_rn(A, p.grid_tap[m], *v);
                        }
                    }
                }
            }
        }
        add_signal_pulses<NT>(a, t, POLY ? p.coef32 : coef_g, V, xt, sigma1);
        digitise_trace<NT>(p, t, V, out_s + tid * B);
    }
    __syncthreads();
    store_run<NT>(a, trace0, out_s);
}

}  // namespace dct
