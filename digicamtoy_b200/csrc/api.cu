// C ABI of libdigicamtoy_sm100.so (declared in include/digicamtoy_b200.h).
// Host-side only: builds the device parameter block, picks launch geometry, launches kernels.
#include "../../include/digicamtoy_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include <memory>

#include "device_params.cuh"
#include "generate.cuh"
#include "sweep.cuh"
#include "grid.cuh"
#include "tensor.cuh"
#include "replay.cuh"
#include "stats.cuh"
#include "selftest.cuh"
#include "dump.cuh"

#ifndef DCT_STATS_COMPACT_BINS
#define DCT_STATS_COMPACT_BINS 0        // above this trace length the tile kernel runs in its compact layout
                                        // (measured: 3.36 against 5.71 ms on the 92-bin dark cube, where the other
                                        //  layout leaves 2 CTAs per SM; 2.19 against 2.54 ms on the 50-bin C4 cube)
#endif
#ifndef DCT_STATS_COMPACT_HIST
#define DCT_STATS_COMPACT_HIST 1024     // ... with a CTA histogram of this many ADC values
#endif
#ifndef DCT_STATS_COMPACT_SINGLE
#define DCT_STATS_COMPACT_SINGLE 1      // ... and one tile buffer
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(DCT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

struct DeviceArena {
    // one allocation for all parameter arrays
    unsigned char* base = nullptr;
    size_t size = 0;
};

}  // namespace

#include "host_transport.h"

// 12-bit transport of dct_generate_host: SLOTS chunks in flight (generated, copied, being unpacked)
#ifndef DCT_STATS_WIDE_SHIFT
#define DCT_STATS_WIDE_SHIFT 5.0       // NSB pedestal shift [LSB] from which K3 skips its thread-private histogram columns
                                       // (measured, profiles/r02/stats_wide.txt: K3 1.92 -> 1.64 ms on C4, 1.49 -> 1.27 ms at 0.125 GHz)
#endif
#ifndef DCT_HOST_PACKED_DEFAULT
#define DCT_HOST_PACKED_DEFAULT 0      // dct_generate_host: 12-bit transport when the ADC range allows it
#endif
struct PackedPipe {
    static constexpr int SLOTS = 3;
    unsigned char* d_buf[SLOTS] = {nullptr, nullptr, nullptr};
    unsigned char* h_buf[SLOTS] = {nullptr, nullptr, nullptr};
    float* d_sig[SLOTS][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t ev_gen[SLOTS] = {nullptr, nullptr, nullptr}, ev_copy[SLOTS] = {nullptr, nullptr, nullptr};
    std::unique_ptr<dct::WorkerPool> pool;
    ~PackedPipe() {
        for (int i = 0; i < SLOTS; ++i) {
            if (d_buf[i]) cudaFree(d_buf[i]);
            if (h_buf[i]) cudaFreeHost(h_buf[i]);
            for (int j = 0; j < 2; ++j) if (d_sig[i][j]) cudaFree(d_sig[i][j]);
            if (ev_gen[i]) cudaEventDestroy(ev_gen[i]);
            if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
        }
    }
};

struct dct_handle {
    int device = 0;
    int sm_count = 148;
    int smem_optin = 0;                // largest dynamic shared memory of one CTA [bytes]
    dct::DevParams p{};
    DeviceArena arena;
    int kernel = DCT_KERNEL_SCATTER;
    float grid_tap[dct::GRID_TAPS] = {0};   // on-grid taps as kernel parameters (K1c)
    int gen_threads = dct::GEN_THREADS;
    bool poly = false;
    bool table_in_smem = false;
    int n_table_entries = 0;
    size_t gen_smem = 0;
    double mean_pe_per_trace = 0.0;
    double mean_true_baseline = 0.0;   // baseline + NSB shift, averaged over pixels (stats window)
    double mean_baseline = 0.0;
    double max_pulse_lsb = 0.0;        // tallest expected signal pulse over the pedestal (statistics window sizes)
    // tensor-core kernel (K1t): coefficient tiles live in the arena
    const __half* tc_coef = nullptr;
    int tc_n_cells = 0, tc_col_shift = 0, tc_col0 = 16, tc_n_cells32 = 0;
    int tc_mega_groups = 0, tc_mega_n_cells32 = 0;     // persistent kernel: groups per SM (0: CTA kernel), its N = 32 cells
    float tc_v_scale = 0.0f;
    int tc_check_span = 1;
    float tc_y0 = 0.0f, tc_y_end = 0.0f;
    double tc_err_lsb = 0.0;           // estimated rms of K1t's fp16 rounding in the analog sums [LSB]
    // host-output pipeline (dct_generate_host), created lazily
    cudaStream_t s_gen = nullptr, s_copy = nullptr;
    cudaEvent_t ev_gen[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    int16_t* d_chunk[2] = {nullptr, nullptr};
    float* d_sig[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int16_t* h_stage[2] = {nullptr, nullptr};
    uint32_t chunk_events = 0, chunk_unit = 0;      // largest chunk of the host pipeline, its unit [events]
    bool pipeline_ready = false;
    int host_transport = 0;            // 0: int16 over PCIe, 1: 12-bit samples, widened by host threads
    std::unique_ptr<PackedPipe> packed;
};

namespace {

template <typename T>
size_t arena_reserve(size_t& cursor, size_t n) {
    cursor = (cursor + 255) & ~(size_t)255;
    size_t at = cursor;
    cursor += n * sizeof(T);
    return at;
}

// evaluates the (already normalised) spline in double, inclusive ends
double spline_at(const dct_config* c, double x) {
    const int n_iv = (int)c->n_knots - 1;
    const double x_lo = c->knot_x[0], x_hi = c->knot_x[n_iv];
    if (!(x >= x_lo && x <= x_hi)) return 0.0;
    const double dx = (x_hi - x_lo) / n_iv;
    int m = (int)std::floor((x - x_lo) / dx);
    m = std::max(0, std::min(m, n_iv - 1));
    if (m > 0 && x < c->knot_x[m]) --m;
    else if (m < n_iv - 1 && x >= c->knot_x[m + 1]) ++m;
    const double u = x - c->knot_x[m];
    const double* k = c->coef + 4 * m;
    return ((k[3] * u + k[2]) * u + k[1]) * u + k[0];
}

int validate(const dct_config* c) {
    if (!c) return fail(DCT_E_ARG, "config is NULL");
    if (c->struct_size != sizeof(dct_config))
        return fail(DCT_E_ARG, "dct_config.struct_size %u != %zu (ABI mismatch)", c->struct_size,
                    sizeof(dct_config));
    if (c->n_pixels == 0 || c->n_pulses == 0 || c->n_bins == 0)
        return fail(DCT_E_ARG, "n_pixels, n_pulses and n_bins must be > 0");
    if (c->n_knots < 4) return fail(DCT_E_ARG, "template needs >= 4 knots");
    if (!(c->dt > 0.0) || !(c->t_end > c->t0)) return fail(DCT_E_ARG, "bad time grid");
    if (c->adc_min > c->adc_max || c->adc_min < -32768 || c->adc_max > 32767)
        return fail(DCT_E_ARG, "saturation range must fit int16");
    const void* ptrs[] = {c->nsb_rate, c->crosstalk, c->gain, c->sigma_1, c->sigma_e, c->baseline,
                          c->n_photon, c->time_signal, c->jitter, c->knot_x, c->coef};
    for (const void* q : ptrs)
        if (!q) return fail(DCT_E_ARG, "a required config array is NULL");
    const int n_iv = (int)c->n_knots - 1;
    const double dx = (c->knot_x[n_iv] - c->knot_x[0]) / n_iv;
    if (!(dx > 0.0)) return fail(DCT_E_ARG, "template knots must ascend");
    for (int i = 0; i <= n_iv; ++i)
        if (std::fabs(c->knot_x[i] - (c->knot_x[0] + i * dx)) > 1e-9 * std::max(1.0, std::fabs(dx)))
            return fail(DCT_E_UNSUPPORTED, "template knots must be uniformly spaced");
    for (uint32_t i = 0; i < c->n_pixels; ++i) {
        if (!(c->crosstalk[i] >= 0.0 && c->crosstalk[i] < 1.0))
            return fail(DCT_E_ARG, "crosstalk[%u]=%g outside [0,1)", i, c->crosstalk[i]);
        if (!(c->nsb_rate[i] >= 0.0)) return fail(DCT_E_ARG, "nsb_rate[%u] negative", i);
        if (!std::isfinite(c->gain[i]) || !std::isfinite(c->baseline[i]))
            return fail(DCT_E_ARG, "gain/baseline[%u] not finite", i);
    }
    return DCT_OK;
}

void free_pipeline(dct_handle* h) {
    for (int i = 0; i < 2; ++i) {
        if (h->d_chunk[i]) cudaFree(h->d_chunk[i]);
        if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
        for (int j = 0; j < 2; ++j)
            if (h->d_sig[i][j]) cudaFree(h->d_sig[i][j]);
        if (h->ev_gen[i]) cudaEventDestroy(h->ev_gen[i]);
        if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
        h->d_chunk[i] = nullptr; h->h_stage[i] = nullptr;
        h->d_sig[i][0] = h->d_sig[i][1] = nullptr;
        h->ev_gen[i] = h->ev_copy[i] = nullptr;
    }
    if (h->s_gen) cudaStreamDestroy(h->s_gen);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    h->s_gen = h->s_copy = nullptr;
    h->chunk_events = 0;
    h->pipeline_ready = false;
    h->packed.reset();
}

template <bool POLY, int NT>
int launch_scatter(dct_handle* h, const dct::GenArgs& a, cudaStream_t st) {
    auto kern = dct::k_generate_scatter<POLY, NT>;
    const size_t smem = dct::smem_with_stats(h->gen_smem, a, NT);
    if (smem > 227 * 1024) return fail(DCT_E_UNSUPPORTED, "trace too long for fused statistics");
    static thread_local size_t configured[16] = {0};
    if (smem > 48 * 1024 && configured[h->device & 15] < smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->device & 15] = smem;
    }
    const uint64_t blocks = (a.n_traces + NT - 1) / NT;
    if (blocks > 0x7fffffffull) return fail(DCT_E_ARG, "too many traces for one launch");
    kern<<<(unsigned)blocks, NT, smem, st>>>(a);
    CK(cudaGetLastError());
    return DCT_OK;
}

}  // namespace

extern "C" {

int dct_abi_version(void) { return DCT_ABI_VERSION; }
const char* dct_last_error(void) { return g_err; }

int dct_create(const dct_config* c, int device, dct_handle** out) {
    if (!out) return fail(DCT_E_ARG, "out is NULL");
    *out = nullptr;
    int rc = validate(c);
    if (rc) return rc;
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    if (n_dev <= 0) return fail(DCT_E_CUDA, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(DCT_E_ARG, "device %d out of range", device);
    CK(cudaSetDevice(device));

    dct_handle* h = new (std::nothrow) dct_handle();
    if (!h) return fail(DCT_E_NOMEM, "out of host memory");
    h->device = device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);

    const int P = (int)c->n_pixels, K = (int)c->n_pulses, n_iv = (int)c->n_knots - 1;
    dct::DevParams& p = h->p;
    p.P = P; p.K = K; p.B = (int)c->n_bins; p.n_drop = (int)c->n_drop; p.n_iv = n_iv;
    p.poisson = (int)c->poisson; p.sub_binning = (int)c->sub_binning;
    p.adc_min = c->adc_min; p.adc_max = c->adc_max;
    const double x_lo = c->knot_x[0], x_hi = c->knot_x[n_iv];
    const double dx = (x_hi - x_lo) / n_iv;
    p.dt = (float)c->dt; p.inv_dt = (float)(1.0 / c->dt);
    p.span = (float)(c->t_end - c->t0);
    p.x_lo = (float)x_lo; p.L = (float)(x_hi - x_lo);
    p.knot_dx = (float)dx; p.inv_knot_dx = (float)(1.0 / dx);
    p.t0f = (float)c->t0;
    const double cell0 = std::floor(x_lo / c->dt + 1e-12);
    p.cell0 = (int)cell0;
    p.off0 = (float)std::max(0.0, x_lo - cell0 * c->dt);
    p.t0_d = c->t0; p.dt_d = c->dt; p.x_lo_d = x_lo; p.x_hi_d = x_hi; p.inv_knot_dx_d = 1.0 / dx;

    // polyphase decomposition: possible when the sampling period is a whole number of knot steps
    const double ratio = c->dt / dx;
    const double nph = std::floor(ratio + 0.5);
    p.n_phase = 0; p.n_taps = 0;
    if (nph >= 1.0 && nph <= 64.0 && std::fabs(ratio - nph) < 1e-9 * ratio) {
        p.n_phase = (int)nph;
        p.n_taps = (n_iv + p.n_phase - 1) / p.n_phase;
    }
    p.phase_cap = 8388608.0f + (float)std::max(0, p.n_phase - 1);
    // on-grid taps
    const int m_lo = (int)std::ceil(x_lo / c->dt - 1e-9), m_hi = (int)std::floor(x_hi / c->dt + 1e-9);
    p.grid_m_lo = m_lo; p.grid_n = std::max(0, m_hi - m_lo + 1);

    // ---- host staging of every device array
    std::vector<float> f_rate(P), f_inv(P), f_xt(P), f_exp((size_t)P * 4), f_gain(P), f_s1(P), f_se(P), f_base(P);
    std::vector<double> d_gain(P), d_base(P);
    std::vector<float> f_npe((size_t)P * K), f_tsig((size_t)P * K), f_jit((size_t)P * K);
    double pe_sum = 0.0;
    for (int i = 0; i < P; ++i) {
        f_rate[i] = (float)c->nsb_rate[i];
        f_inv[i] = c->nsb_rate[i] > 0.0 ? (float)(-0.6931471805599453 / c->nsb_rate[i]) : 0.0f;
        f_xt[i] = (float)c->crosstalk[i];
        {   // Borel(mu): P(1) = e^-mu, P(n+1)/P(n) = mu e^-mu (1+1/n)^(n-1)
            const double mu = c->crosstalk[i];
            double pn = std::exp(-mu), cdf = pn;
            for (int n = 1; n <= 4; ++n) {
                f_exp[(size_t)i * 4 + (n - 1)] = (float)std::min(cdf, 1.0);
                pn *= mu * std::exp(-mu) * std::pow(1.0 + 1.0 / n, n - 1.0);
                cdf += pn;
            }
        }
        f_gain[i] = (float)c->gain[i];
        f_s1[i] = (float)c->sigma_1[i];
        f_se[i] = (float)c->sigma_e[i];
        f_base[i] = (float)c->baseline[i];
        d_gain[i] = c->gain[i];
        d_base[i] = c->baseline[i];
        pe_sum += c->nsb_rate[i] * (c->t_end - c->t0);
    }
    h->mean_pe_per_trace = pe_sum / P;
    {   // expected pedestal incl. the NSB shift: baseline + gain rate tau / (1 - xt), tau = integral of h
        double tau = 0.0;
        for (int m = 0; m < n_iv; ++m) {
            const double* k = c->coef + 4 * m;
            tau += ((k[3] * dx / 4 + k[2] / 3) * dx + k[1] / 2) * dx * dx + k[0] * dx;
        }
        double acc = 0.0;
        for (int i = 0; i < P; ++i)
            acc += c->baseline[i] + c->gain[i] * c->nsb_rate[i] * tau / (1.0 - c->crosstalk[i]);
        h->mean_true_baseline = acc / P;
        double tallest = 0.0, h_peak = 0.0;
        for (int m = 0; m < n_iv; ++m) h_peak = std::max(h_peak, std::fabs(c->coef[4 * m]));
        for (int i = 0; i < P; ++i)
            for (int k = 0; k < K; ++k)
                tallest = std::max(tallest, std::fabs(c->gain[i]) * h_peak * std::max(0.0, c->n_photon[(size_t)i * K + k]) /
                                                std::max(1e-6, 1.0 - c->crosstalk[i]));
        h->max_pulse_lsb = tallest;
        double bsum = 0.0;
        for (int i = 0; i < P; ++i) bsum += std::nearbyint(c->baseline[i]);
        h->mean_baseline = bsum / P;
    }
    for (size_t i = 0; i < (size_t)P * K; ++i) {
        f_npe[i] = (float)c->n_photon[i];
        f_tsig[i] = (float)c->time_signal[i];
        f_jit[i] = (float)c->jitter[i];
    }
    std::vector<float> f_coef((size_t)n_iv * 4);
    for (size_t i = 0; i < (size_t)n_iv * 4; ++i) f_coef[i] = (float)c->coef[i];
    std::vector<float> f_poly((size_t)p.n_taps * p.n_phase * 4, 0.0f);
    for (int j = 0; j < p.n_taps; ++j)
        for (int f = 0; f < p.n_phase; ++f) {
            const int m = f + p.n_phase * j;
            if (m < n_iv)
                for (int q = 0; q < 4; ++q)
                    f_poly[((size_t)j * p.n_phase + f) * 4 + q] = (float)c->coef[4 * m + q];
        }
    std::vector<float> f_grid((size_t)std::max(1, p.grid_n));
    for (int m = 0; m < p.grid_n; ++m) f_grid[m] = (float)spline_at(c, (m_lo + m) * c->dt);
    float grid_tap_host[dct::GRID_TAPS] = {0};
    for (int m = 0; m < p.grid_n && m < dct::GRID_TAPS; ++m) grid_tap_host[m] = f_grid[m];

    // K1t coefficient tiles: G[n][(f, k)] = coefficient k of tap n, phase f re-expanded in
    // v = 2u/dx - 1  (u = a (v + 1), a = dx / 2), split into fp16 hi + lo, in the canonical K-major
    // no-swizzle UMMA layout (dct::tc_tile_offset)
    // four tiles: hi, lo, and both again one row further down (for cells that start on an odd column)
    const size_t tc_tile = (size_t)dct::TC_N * dct::TC_K;
    std::vector<uint16_t> tc_coef(4 * tc_tile, 0);
    bool tensor_ok = dct::tensor_supported(p);
    if (tensor_ok) {
        const double a1 = 0.5 * dx, a2 = a1 * a1, a3 = a2 * a1;
        for (int n = 0; n < dct::TC_TAPS; ++n)
            for (int f = 0; f < dct::TC_PHASES; ++f) {
                const int m = f + dct::TC_PHASES * n;
                if (m >= n_iv) continue;
                const double* k = c->coef + 4 * m;
                const double g[4] = {k[0] + k[1] * a1 + k[2] * a2 + k[3] * a3,
                                     k[1] * a1 + 2.0 * k[2] * a2 + 3.0 * k[3] * a3,
                                     k[2] * a2 + 3.0 * k[3] * a3, k[3] * a3};
                for (int q = 0; q < 4; ++q) {
                    const __half hi = __float2half_rn((float)g[q]);
                    const __half lo = __float2half_rn((float)(g[q] - (double)__half2float(hi)));
                    // K index by j = 7 - f, the knot interval within the cell in time order (see tensor.cuh)
                    const int kk = (dct::TC_PHASES - 1 - f) * 4 + q;
                    const size_t at = (size_t)dct::tc_tile_offset(n, kk) / 2;
                    const size_t at1 = (size_t)dct::tc_tile_offset(n + 1, kk) / 2;
                    memcpy(&tc_coef[at], &hi, 2);
                    memcpy(&tc_coef[tc_tile + at], &lo, 2);
                    memcpy(&tc_coef[2 * tc_tile + at1], &hi, 2);
                    memcpy(&tc_coef[3 * tc_tile + at1], &lo, 2);
                }
            }
        const int n_time = (int)std::ceil(((double)p.off0 + (c->t_end - c->t0)) / c->dt) + 1;
        const int n_useful = p.B - 1 + p.n_drop - p.cell0;
        h->tc_n_cells = std::max(0, std::min(n_time, n_useful));
        const dct::TensorLayout lay = dct::tensor_layout(p.B, p.n_drop, p.cell0, h->tc_n_cells);
        tensor_ok = lay.ok;
        h->tc_col_shift = lay.col_shift; h->tc_col0 = lay.col0; h->tc_n_cells32 = lay.n_cells32;
        // persistent kernel: five groups per SM when the window fits 102 accumulator columns, else four
        const dct::TensorLayout lay5 = dct::tensor_layout(p.B, p.n_drop, p.cell0, h->tc_n_cells, dct::TC_MEGA_COLS5);
        const bool fits5 = lay5.ok && lay5.col0 + ((p.B + 15) / 16) * 16 <= dct::TC_MEGA_COLS5 &&
                           dct::tensor_mega_smem_bytes(p.B, 5, 5) <= (size_t)h->smem_optin;
        h->tc_mega_groups = fits5 ? 5 : (dct::tensor_mega_smem_bytes(p.B, 6, 4) <= (size_t)h->smem_optin ? 4 : 0);
        h->tc_mega_n_cells32 = fits5 ? lay5.n_cells32 : lay.n_cells32;
        const char* mega_env = getenv("DCT_TENSOR_MEGA");
        if (mega_env ? atoi(mega_env) == 0 : !DCT_TC_MEGA) h->tc_mega_groups = 0;
        if (mega_env && atoi(mega_env) == 4 && h->tc_mega_groups == 5) {       // A/B: four groups, six slots
            h->tc_mega_groups = 4; h->tc_mega_n_cells32 = lay.n_cells32;
        }
        h->tc_v_scale = (float)(2.0 / dx);
        // the cell test ends the pe stream by itself when the useful cells end a cell before the window
        h->tc_check_span = ((double)n_useful + 1.0) * c->dt <= (double)p.off0 + (c->t_end - c->t0) ? 0 : 1;
        // producer clock in knot steps from the start of cell0; the stream ends at the end of the
        // NSB window or after the last useful cell, whichever comes first
        h->tc_y0 = (float)((double)p.off0 / dx);
        h->tc_y_end = (float)std::min(((double)p.off0 + (c->t_end - c->t0)) / dx,
                                      (double)h->tc_n_cells * dct::TC_PHASES);
    }

    // generalised-Poisson tables of the on-grid mode, per pixel, in double (sampling.cuh: gp_table_draw)
    std::vector<float> f_gp;
    if (p.sub_binning > 0) {
        f_gp.assign((size_t)P * dct::GP_TABLE, 1.0f);
        for (int i = 0; i < P; ++i) {
            const double lam = c->nsb_rate[i] * c->dt, mu = c->crosstalk[i];
            if (!(lam > 0.0) || lam >= (double)dct::GP_LAM_MAX) continue;
            double cdf = 0.0;
            for (int k = 0; k < dct::GP_TABLE; ++k) {
                const double a = lam + k * mu;
                cdf += std::exp(std::log(lam) + (k - 1.0) * std::log(a) - a - std::lgamma(k + 1.0));
                f_gp[(size_t)i * dct::GP_TABLE + k] = (float)std::min(cdf, 1.0);
            }
        }
    }

    // ... and of the small signal pulses (Poisson(n_photon) primaries + their crosstalk progeny), per pixel and pulse
    std::vector<float> f_sig_gp;
    if (c->poisson) {
        bool any_small = false;
        for (size_t i = 0; i < (size_t)P * K; ++i) any_small |= c->n_photon[i] > 0.0 && c->n_photon[i] < (double)dct::GP_LAM_MAX;
        if (any_small) {
            f_sig_gp.assign((size_t)P * K * dct::GP_TABLE, 1.0f);
            for (size_t i = 0; i < (size_t)P * K; ++i) {
                const double lam = (double)(float)c->n_photon[i], mu = c->crosstalk[i / K];
                if (!(lam > 0.0) || lam >= (double)dct::GP_LAM_MAX) continue;
                double cdf = 0.0;
                for (int k = 0; k < dct::GP_TABLE; ++k) {
                    const double a = lam + k * mu;
                    cdf += std::exp(std::log(lam) + (k - 1.0) * std::log(a) - a - std::lgamma(k + 1.0));
                    f_sig_gp[i * dct::GP_TABLE + k] = (float)std::min(cdf, 1.0);
                }
            }
        }
    }

    // ---- one arena
    size_t cur = 0;
    const size_t o_rate = arena_reserve<float>(cur, P), o_inv = arena_reserve<float>(cur, P),
                 o_xt = arena_reserve<float>(cur, P), o_exp = arena_reserve<float>(cur, (size_t)P * 4),
                 o_gain = arena_reserve<float>(cur, P), o_s1 = arena_reserve<float>(cur, P),
                 o_se = arena_reserve<float>(cur, P), o_base = arena_reserve<float>(cur, P),
                 o_dgain = arena_reserve<double>(cur, P), o_dbase = arena_reserve<double>(cur, P),
                 o_npe = arena_reserve<float>(cur, (size_t)P * K),
                 o_tsig = arena_reserve<float>(cur, (size_t)P * K),
                 o_jit = arena_reserve<float>(cur, (size_t)P * K),
                 o_c32 = arena_reserve<float>(cur, (size_t)n_iv * 4),
                 o_c64 = arena_reserve<double>(cur, (size_t)n_iv * 4),
                 o_kn = arena_reserve<double>(cur, (size_t)n_iv + 1),
                 o_poly = arena_reserve<float>(cur, f_poly.size() + 4),
                 o_grid = arena_reserve<float>(cur, f_grid.size()),
                 o_tc = arena_reserve<uint16_t>(cur, tc_coef.size()),
                 o_gp = arena_reserve<float>(cur, f_gp.size()),
                 o_sgp = arena_reserve<float>(cur, f_sig_gp.size());
    h->arena.size = cur + 256;
    cudaError_t e = cudaMalloc(&h->arena.base, h->arena.size);
    if (e != cudaSuccess) {
        delete h;
        return fail(DCT_E_CUDA, "cudaMalloc(%zu) failed: %s", cur, cudaGetErrorString(e));
    }
    unsigned char* base = h->arena.base;
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpy(base + off, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t ue = cudaSuccess;
#define UP(off, vec) if (ue == cudaSuccess) ue = up(off, (vec).data(), (vec).size() * sizeof((vec)[0]))
    UP(o_rate, f_rate); UP(o_inv, f_inv); UP(o_xt, f_xt); UP(o_exp, f_exp); UP(o_gain, f_gain);
    UP(o_s1, f_s1); UP(o_se, f_se); UP(o_base, f_base); UP(o_dgain, d_gain); UP(o_dbase, d_base);
    UP(o_npe, f_npe); UP(o_tsig, f_tsig); UP(o_jit, f_jit); UP(o_c32, f_coef); UP(o_poly, f_poly);
    UP(o_grid, f_grid); UP(o_tc, tc_coef); UP(o_gp, f_gp); UP(o_sgp, f_sig_gp);
#undef UP
    if (ue == cudaSuccess) ue = up(o_c64, c->coef, (size_t)n_iv * 4 * sizeof(double));
    if (ue == cudaSuccess) ue = up(o_kn, c->knot_x, ((size_t)n_iv + 1) * sizeof(double));
    if (ue != cudaSuccess) {
        cudaFree(h->arena.base);
        delete h;
        return fail(DCT_E_CUDA, "parameter upload failed: %s", cudaGetErrorString(ue));
    }
    p.rate = (const float*)(base + o_rate); p.neg_ln2_over_rate = (const float*)(base + o_inv);
    p.xt = (const float*)(base + o_xt); p.borel_cdf = (const float4*)(base + o_exp);
    p.gain = (const float*)(base + o_gain); p.sigma1 = (const float*)(base + o_s1);
    p.sigma_e = (const float*)(base + o_se); p.baseline = (const float*)(base + o_base);
    p.gain_d = (const double*)(base + o_dgain); p.baseline_d = (const double*)(base + o_dbase);
    p.npe = (const float*)(base + o_npe); p.tsig = (const float*)(base + o_tsig);
    p.jitter = (const float*)(base + o_jit);
    p.coef32 = (const float4*)(base + o_c32); p.coef64 = (const double*)(base + o_c64);
    p.knots = (const double*)(base + o_kn); p.poly = (const float4*)(base + o_poly);
    p.grid_tap = (const float*)(base + o_grid);
    p.gp_cdf = f_gp.empty() ? nullptr : (const float*)(base + o_gp);
    p.sig_cdf = f_sig_gp.empty() ? nullptr : (const float*)(base + o_sgp);
    h->tc_coef = (const __half*)(base + o_tc);

    // ---- kernel selection and launch geometry
    h->poly = p.n_phase > 0 && p.sub_binning <= 0;
    h->table_in_smem = !h->poly && n_iv <= 4096;
    h->n_table_entries = h->poly ? p.n_taps * p.n_phase : (h->table_in_smem ? n_iv : 0);
    h->gen_threads = dct::GEN_THREADS;
    h->gen_smem = dct::gen_smem_bytes(p.B, h->n_table_entries, h->gen_threads);
    const size_t smem_cap = 227 * 1024;
    if (h->gen_smem > smem_cap) {
        h->gen_threads = 32;
        h->gen_smem = dct::gen_smem_bytes(p.B, h->n_table_entries, h->gen_threads);
    }
    if (h->gen_smem > smem_cap) {
        cudaFree(h->arena.base);
        delete h;
        return fail(DCT_E_UNSUPPORTED, "trace of %d bins does not fit shared memory", p.B);
    }
    const bool sweep_ok = dct::sweep_supported(p);
    const bool grid_ok = dct::grid_supported(p);
    for (int m = 0; m < dct::GRID_TAPS; ++m) h->grid_tap[m] = grid_tap_host[m];
    int want = (int)c->kernel;
    if (want == DCT_KERNEL_GRID && !grid_ok) {
        cudaFree(h->arena.base);
        delete h;
        return fail(DCT_E_UNSUPPORTED, "grid kernel needs sub_binning > 0 and at most %d on-grid template taps",
                    dct::GRID_TAPS);
    }
    if (want == DCT_KERNEL_SWEEP && !sweep_ok) {
        cudaFree(h->arena.base);
        delete h;
        return fail(DCT_E_UNSUPPORTED,
                    "sweep kernel needs the polyphase table with %d taps, continuous NSB and <= %d bins",
                    dct::SWEEP_TAPS, dct::SWEEP_MAX_BINS);
    }
    if (want == DCT_KERNEL_TENSOR && !(tensor_ok && h->tc_n_cells > 0)) {
        cudaFree(h->arena.base);
        delete h;
        return fail(DCT_E_UNSUPPORTED,
                    "tensor kernel needs the polyphase table with %d taps, continuous NSB and an even number of <= %d bins",
                    dct::TC_TAPS, dct::TC_MAX_BINS);
    }
    // K1t sums through fp16 moments: measured 2.6e-3 LSB rms on C4 (gain 2.5 LSB/pe, template peak 1,
    // 1 GHz = 88 overlapping pe), i.e. 1.1e-4 x gain x template peak x sqrt(overlapping pe).  AUTO only
    // takes it while that estimate stays below 0.01 LSB (the reference normalises the template by its
    // maximum INSIDE the simulated window, :144-149, so short or late windows can scale it up a lot).
    double tc_err_lsb = 0.0;
    {
        double h_max = 0.0, g_max = 0.0, r_max = 0.0;
        for (int m = 0; m < n_iv; ++m) h_max = std::max(h_max, std::fabs(c->coef[4 * m]));
        for (int i = 0; i < P; ++i) { g_max = std::max(g_max, std::fabs(c->gain[i])); r_max = std::max(r_max, c->nsb_rate[i]); }
        tc_err_lsb = 1.1e-4 * g_max * h_max * std::sqrt(std::max(1.0, r_max * (x_hi - x_lo)));
    }
    h->tc_err_lsb = tc_err_lsb;
    if (want == DCT_KERNEL_AUTO)
        want = grid_ok ? DCT_KERNEL_GRID
             : (tensor_ok && h->tc_n_cells > 0 && h->mean_pe_per_trace >= DCT_TC_MIN_MEAN_PE && tc_err_lsb <= 0.01) ? DCT_KERNEL_TENSOR
             : (sweep_ok && h->mean_pe_per_trace >= dct::SWEEP_MIN_MEAN_PE) ? DCT_KERNEL_SWEEP
                                                                            : DCT_KERNEL_SCATTER;
    h->kernel = want;
    g_err[0] = 0;
    *out = h;
    return DCT_OK;
}

void dct_destroy(dct_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_pipeline(h);
    if (h->arena.base) cudaFree(h->arena.base);
    delete h;
}

int dct_selected_kernel(const dct_handle* h) { return h ? h->kernel : DCT_E_ARG; }

// histogram windows of the statistics kernels: centred on the expected pedestal (true_baseline,
// host-computed); pulses only go up, so three quarters of a window lie above it
static int stats_win_lo(const dct_handle* h, int win) { return (int)std::floor(h->mean_true_baseline) - win / 4; }

static int generate_impl(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events, int16_t* adc,
                         float* sig_time, float* sig_charge, float* analog, void* stream,
                         const dct::FusedStats* stats = nullptr, bool pack12 = false) {
    if (!h || !adc) return fail(DCT_E_ARG, "handle or adc is NULL");
    if (n_events == 0) return DCT_OK;
    if ((first_event + n_events) >> 56) return fail(DCT_E_ARG, "event index exceeds 2^56");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    dct::GenArgs a;
    a.p = h->p;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.rk = dct::make_round_keys(a.k0, a.k1);
    a.first_event = first_event;
    a.n_traces = (uint64_t)n_events * (uint64_t)h->p.P;
    a.adc = adc; a.sig_time = sig_time; a.sig_charge = sig_charge;
    a.table_in_smem = h->table_in_smem ? 1 : 0;
    a.analog = analog;
    if (stats) a.st = *stats;
    a.pack12 = pack12 ? 1 : 0;
    if (h->kernel == DCT_KERNEL_TENSOR) {
        dct::TensorArgs ta;
        ta.g = a;
        ta.coef = h->tc_coef;
        ta.n_cells = h->tc_n_cells; ta.col_shift = h->tc_col_shift; ta.v_scale = h->tc_v_scale;
        ta.col0 = h->tc_col0; ta.n_cells32 = h->tc_n_cells32;
        ta.check_span = h->tc_check_span;
        ta.y0 = h->tc_y0; ta.y_end = h->tc_y_end;
        // the persistent kernel has no fused statistics: those launches take the CTA kernel
        if (h->tc_mega_groups && !(stats && stats->adc_hist)) {
            ta.n_cells32 = h->tc_mega_n_cells32;
            return dct::launch_tensor_mega(h->device, h->sm_count, h->tc_mega_groups, ta, st, g_err, sizeof(g_err));
        }
        return dct::launch_tensor(h->device, ta, st, g_err, sizeof(g_err));
    }
    if (h->kernel == DCT_KERNEL_GRID) return dct::launch_grid(h->device, a, h->grid_tap, st, g_err, sizeof(g_err));
    if (h->kernel == DCT_KERNEL_SWEEP) return dct::launch_sweep(h->device, h->sm_count, a, st, g_err, sizeof(g_err));
    if (h->gen_threads == 32)
        return h->poly ? launch_scatter<true, 32>(h, a, st) : launch_scatter<false, 32>(h, a, st);
    return h->poly ? launch_scatter<true, dct::GEN_THREADS>(h, a, st)
                   : launch_scatter<false, dct::GEN_THREADS>(h, a, st);
}

int dct_generate(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events, int16_t* adc,
                 float* sig_time, float* sig_charge, void* stream) {
    return generate_impl(h, seed, first_event, n_events, adc, sig_time, sig_charge, nullptr, stream);
}

int dct_generate_stats(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events, int16_t* adc,
                       float* sig_time, float* sig_charge, unsigned long long* adc_hist, double* bin_sum,
                       double* bin_sumsq, double* pixel_sum, double* pixel_sumsq,
                       unsigned long long* charge_hist, int32_t charge_lo, uint32_t n_charge_bins,
                       double* moments, void* stream) {
    if (!h) return fail(DCT_E_ARG, "handle is NULL");
    if (!adc_hist || !bin_sum || !bin_sumsq || !pixel_sum || !pixel_sumsq)
        return fail(DCT_E_ARG, "adc_hist, bin_sum, bin_sumsq, pixel_sum and pixel_sumsq are required");
    if (charge_hist && n_charge_bins == 0) return fail(DCT_E_ARG, "n_charge_bins must be > 0");
    const int n_hist = h->p.adc_max - h->p.adc_min + 1;
    if (n_hist > 16384) return fail(DCT_E_UNSUPPORTED, "ADC histogram wider than 16384 bins");
    dct::FusedStats st;
    st.adc_hist = adc_hist; st.bin_sum = bin_sum; st.bin_sumsq = bin_sumsq;
    st.pixel_sum = pixel_sum; st.pixel_sumsq = pixel_sumsq; st.charge_hist = charge_hist;
    st.moments = moments; st.charge_lo = charge_lo; st.n_charge_bins = (int)n_charge_bins;
    st.win_lo = stats_win_lo(h, dct::FS_WIN);
    st.hist_lo = st.win_lo - dct::FS_HIST / 8;
    return generate_impl(h, seed, first_event, n_events, adc, sig_time, sig_charge, nullptr, stream, &st);
}

int dct_generate_analog(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events, int16_t* adc,
                        float* analog, void* stream) {
    if (!analog) return fail(DCT_E_ARG, "analog is NULL");
    return generate_impl(h, seed, first_event, n_events, adc, nullptr, nullptr, analog, stream);
}

// chunk ends [events] of one dct_generate_host call; sizes in units: 1, 2, 4, 4, ... 4, (rest), 2, 1
static std::vector<uint32_t> host_chunk_schedule(uint32_t unit, uint32_t n_events) {
    const uint64_t U = ((uint64_t)n_events + unit - 1) / unit;
    std::vector<uint32_t> sizes, chunk_end;
    if (U <= 8) sizes.assign((size_t)U, 1);
    else {
        sizes = {1, 2};
        for (uint64_t left = U - 6; left > 0; ) { const uint32_t k = (uint32_t)std::min<uint64_t>(4, left); sizes.push_back(k); left -= k; }
        sizes.push_back(2); sizes.push_back(1);
    }
    uint64_t at = 0;
    for (uint32_t k : sizes) {
        const uint64_t next = std::min<uint64_t>(n_events, at + (uint64_t)k * unit);
        if (next > at) chunk_end.push_back((uint32_t)next);
        at = next;
    }
    return chunk_end;
}

// dct_generate_host with the 12-bit transport: the kernels write a chunk as a stream of 12-bit samples,
// the copy engine brings it into a pinned staging slot, host threads widen it into the caller's buffer
// (pinned or not).  Three slots: while one chunk is widened, the next two are generated and copied.
static int generate_host_packed(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                                int16_t* adc_host, float* sig_time_host, float* sig_charge_host) {
    PackedPipe& pp = *h->packed;
    constexpr uint32_t K = PackedPipe::SLOTS;
    const size_t trace_samples = (size_t)h->p.P * h->p.B;
    const size_t sig_elems = (size_t)h->p.P * h->p.K;
    const std::vector<uint32_t> chunk_end = host_chunk_schedule(h->chunk_unit, n_events);
    const uint32_t n_chunks = (uint32_t)chunk_end.size();
    int rc = DCT_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == DCT_OK) rc = fail(DCT_E_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    auto first_of = [&](uint32_t c) { return c ? chunk_end[c - 1] : 0u; };
    uint32_t issued = 0, widened = 0;
    while (widened < n_chunks && rc == DCT_OK) {
        // keep the device busy: every free slot gets the next chunk (generation, then its copies)
        while (issued < n_chunks && issued - widened < K && rc == DCT_OK) {
            const uint32_t slot = issued % K, e0 = first_of(issued), ne = chunk_end[issued] - e0;
            rc = generate_impl(h, seed, first_event + e0, ne, reinterpret_cast<int16_t*>(pp.d_buf[slot]),
                               sig_time_host ? pp.d_sig[slot][0] : nullptr,
                               sig_charge_host ? pp.d_sig[slot][1] : nullptr, nullptr, h->s_gen, nullptr, true);
            if (rc) break;
            const size_t bytes = (ne * trace_samples * 3 + 1) / 2;
            if (!cuda_ok(cudaEventRecord(pp.ev_gen[slot], h->s_gen), "cudaEventRecord")) break;
            if (!cuda_ok(cudaStreamWaitEvent(h->s_copy, pp.ev_gen[slot], 0), "cudaStreamWaitEvent")) break;
            if (!cuda_ok(cudaMemcpyAsync(pp.h_buf[slot], pp.d_buf[slot], bytes, cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
            if (sig_time_host &&
                !cuda_ok(cudaMemcpyAsync(sig_time_host + (size_t)e0 * sig_elems, pp.d_sig[slot][0],
                                         ne * sig_elems * sizeof(float), cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
            if (sig_charge_host &&
                !cuda_ok(cudaMemcpyAsync(sig_charge_host + (size_t)e0 * sig_elems, pp.d_sig[slot][1],
                                         ne * sig_elems * sizeof(float), cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
            if (!cuda_ok(cudaEventRecord(pp.ev_copy[slot], h->s_copy), "cudaEventRecord")) break;
            ++issued;
        }
        if (rc != DCT_OK || widened >= issued) break;
        // the oldest chunk in flight: wait for its copy, widen it
        const uint32_t slot = widened % K, e0 = first_of(widened), ne = chunk_end[widened] - e0;
        if (!cuda_ok(cudaEventSynchronize(pp.ev_copy[slot]), "cudaEventSynchronize")) break;
        dct::unpack12_parallel(*pp.pool, pp.h_buf[slot], adc_host + (size_t)e0 * trace_samples, ne * trace_samples);
        ++widened;
    }
    // nothing may still be running on the handle's streams (or writing to the caller's truth arrays) on return
    cuda_ok(cudaStreamSynchronize(h->s_copy), "cudaStreamSynchronize");
    cuda_ok(cudaStreamSynchronize(h->s_gen), "cudaStreamSynchronize");
    return rc;
}

int dct_generate_host(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                      int16_t* adc_host, float* sig_time_host, float* sig_charge_host) {
    if (!h || !adc_host) return fail(DCT_E_ARG, "handle or adc_host is NULL");
    if (n_events == 0) return DCT_OK;
    CK(cudaSetDevice(h->device));
    const size_t trace_bytes = (size_t)h->p.P * h->p.B * sizeof(int16_t);
    const size_t sig_elems = (size_t)h->p.P * h->p.K;
    if (!h->pipeline_ready) {
        // chunk ~64 MiB: large enough to amortise launches, small enough to pipeline
        free_pipeline(h);                      // nothing half-built survives a failed earlier attempt
        // Chunks are multiples of a unit of ~16 MiB, at most four units (dct_generate_host ramps the
        // sizes up and down again so that little of the first generation and of the last copy is exposed).
        uint32_t unit = (uint32_t)std::max<size_t>(1, ((size_t)16 << 20) / trace_bytes);
        if (h->kernel == DCT_KERNEL_TENSOR && h->tc_mega_groups) {
            // The persistent kernel hands every group of every SM the same number of 128-trace items
            // (+-1): a unit is what fills three rounds (517 instead of 438 C4 events per launch costs 9 %).
            const uint64_t slots = (uint64_t)h->sm_count * h->tc_mega_groups;
            unit = (uint32_t)std::max<uint64_t>(1, 3 * slots * dct::TC_TRACES / (uint64_t)h->p.P);
        }
        h->chunk_unit = unit;
        const uint32_t ce = 4 * unit;
        h->chunk_events = ce;
        // transport: 12-bit samples when every ADC value fits (DCT_HOST_TRANSPORT=plain / packed12 overrides)
        const char* tr = getenv("DCT_HOST_TRANSPORT");
        const bool fits12 = h->p.adc_min >= 0 && h->p.adc_max <= 4095;
        h->host_transport = (fits12 && (tr ? !strcmp(tr, "packed12") : DCT_HOST_PACKED_DEFAULT)) ? 1 : 0;
        cudaError_t e = cudaStreamCreateWithFlags(&h->s_gen, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking);
        if (e == cudaSuccess && h->host_transport == 1) {
            h->packed.reset(new (std::nothrow) PackedPipe());
            if (!h->packed) { free_pipeline(h); return fail(DCT_E_NOMEM, "out of host memory"); }
            PackedPipe& pp = *h->packed;
            const size_t bytes = ((size_t)ce * h->p.P * h->p.B * 3 + 1) / 2 + 64;
            for (int i = 0; i < PackedPipe::SLOTS && e == cudaSuccess; ++i) {
                e = cudaEventCreateWithFlags(&pp.ev_gen[i], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pp.ev_copy[i], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaMalloc(&pp.d_buf[i], bytes);
                if (e == cudaSuccess) e = cudaMallocHost(&pp.h_buf[i], bytes);
                if (e == cudaSuccess) e = cudaMalloc(&pp.d_sig[i][0], ce * sig_elems * sizeof(float));
                if (e == cudaSuccess) e = cudaMalloc(&pp.d_sig[i][1], ce * sig_elems * sizeof(float));
            }
            // host threads that widen the samples: a share of the cores, 2 .. 8 (DCT_HOST_THREADS overrides)
            int n_gpu = 1;
            cudaGetDeviceCount(&n_gpu);
            int n_thr = (int)std::thread::hardware_concurrency() / (2 * std::max(1, n_gpu));
            n_thr = std::min(8, std::max(2, n_thr));
            if (const char* t = getenv("DCT_HOST_THREADS")) n_thr = std::min(64, std::max(1, atoi(t)));
            pp.pool.reset(new dct::WorkerPool(n_thr));
        }
        for (int i = 0; i < 2 && e == cudaSuccess && h->host_transport == 0; ++i) {
            e = cudaEventCreateWithFlags(&h->ev_gen[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaMalloc(&h->d_chunk[i], ce * trace_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&h->d_sig[i][0], ce * sig_elems * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&h->d_sig[i][1], ce * sig_elems * sizeof(float));
        }
        if (e != cudaSuccess) {
            free_pipeline(h);
            return fail(DCT_E_CUDA, "host pipeline setup failed: %s", cudaGetErrorString(e));
        }
        h->pipeline_ready = true;
    }
    if (h->host_transport == 1)
        return generate_host_packed(h, seed, first_event, n_events, adc_host, sig_time_host, sig_charge_host);
    // Is the destination page-locked?  Then copy straight into it; otherwise go through pinned staging.
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, adc_host) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned && !h->h_stage[0])
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaMallocHost(&h->h_stage[i], h->chunk_events * trace_bytes);
            if (e != cudaSuccess) {
                free_pipeline(h);
                return fail(DCT_E_CUDA, "pinned staging buffer: %s", cudaGetErrorString(e));
            }
        }

    const std::vector<uint32_t> chunk_end = host_chunk_schedule(h->chunk_unit, n_events);
    const uint32_t n_chunks = (uint32_t)chunk_end.size();
    struct Pending { int16_t* dst; size_t bytes; bool live; } pend[2] = {{nullptr, 0, false}, {nullptr, 0, false}};
    // On any error: nothing may still be writing into the caller's buffer when we return.
    int rc = DCT_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == DCT_OK) rc = fail(DCT_E_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    for (uint32_t c = 0; c < n_chunks && rc == DCT_OK; ++c) {
        const int slot = (int)(c & 1);
        const uint32_t e0 = c ? chunk_end[c - 1] : 0, ne = chunk_end[c] - e0;
        if (ne == 0) continue;
        // the slot's previous copy must have drained before we overwrite its device buffer
        if (c >= 2) {
            if (pinned) {
                // the generation stream waits by itself: every launch of the call is queued ahead of time
                if (!cuda_ok(cudaStreamWaitEvent(h->s_gen, h->ev_copy[slot], 0), "cudaStreamWaitEvent")) break;
            } else {
                if (!cuda_ok(cudaEventSynchronize(h->ev_copy[slot]), "cudaEventSynchronize")) break;
                if (pend[slot].live) { memcpy(pend[slot].dst, h->h_stage[slot], pend[slot].bytes); pend[slot].live = false; }
            }
        }
#ifdef DCT_HOST_DEBUG          // diagnosis builds: DCT_HOST_SKIP=gen / copy leaves one stage of the pipeline out
        static const char* skip = getenv("DCT_HOST_SKIP");
        const bool skip_gen = skip && !strcmp(skip, "gen"), skip_copy = skip && !strcmp(skip, "copy");
#else
        constexpr bool skip_gen = false, skip_copy = false;
#endif
        if (!skip_gen)
            rc = dct_generate(h, seed, first_event + e0, ne, h->d_chunk[slot],
                              sig_time_host ? h->d_sig[slot][0] : nullptr,
                              sig_charge_host ? h->d_sig[slot][1] : nullptr, h->s_gen);
        if (rc) break;
        if (!cuda_ok(cudaEventRecord(h->ev_gen[slot], h->s_gen), "cudaEventRecord")) break;
        if (!cuda_ok(cudaStreamWaitEvent(h->s_copy, h->ev_gen[slot], 0), "cudaStreamWaitEvent")) break;
        int16_t* dst = adc_host + (size_t)e0 * h->p.P * h->p.B;
        if (skip_copy) {
        } else if (pinned) {
            if (!cuda_ok(cudaMemcpyAsync(dst, h->d_chunk[slot], ne * trace_bytes, cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
        } else {
            if (!cuda_ok(cudaMemcpyAsync(h->h_stage[slot], h->d_chunk[slot], ne * trace_bytes, cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
            pend[slot].dst = dst; pend[slot].bytes = ne * trace_bytes; pend[slot].live = true;
        }
        if (sig_time_host &&
            !cuda_ok(cudaMemcpyAsync(sig_time_host + (size_t)e0 * sig_elems, h->d_sig[slot][0],
                                     ne * sig_elems * sizeof(float), cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
        if (sig_charge_host &&
            !cuda_ok(cudaMemcpyAsync(sig_charge_host + (size_t)e0 * sig_elems, h->d_sig[slot][1],
                                     ne * sig_elems * sizeof(float), cudaMemcpyDeviceToHost, h->s_copy), "cudaMemcpyAsync")) break;
        if (!cuda_ok(cudaEventRecord(h->ev_copy[slot], h->s_copy), "cudaEventRecord")) break;
        // generation of the next chunk in this slot's sibling may start right away
    }
    // drain both streams whatever happened, then hand over what was staged
    const bool drained = cuda_ok(cudaStreamSynchronize(h->s_copy), "cudaStreamSynchronize") &
                         cuda_ok(cudaStreamSynchronize(h->s_gen), "cudaStreamSynchronize");
    if (drained)
        for (int s2 = 0; s2 < 2; ++s2)
            if (pend[s2].live) memcpy(pend[s2].dst, h->h_stage[s2], pend[s2].bytes);
    return rc;
}

static int replay_common(bool f64, dct_handle* h, uint32_t n_events, const int64_t* nsb_offsets,
                         const double* nsb_time, const double* nsb_amp, const double* sig_time,
                         const double* sig_amp, const double* e_noise, int16_t* adc,
                         unsigned long long* n_clamped, void* stream) {
    if (!h || !nsb_offsets || !sig_time || !sig_amp || !adc)
        return fail(DCT_E_ARG, "a required replay pointer is NULL");
    if (h->p.sub_binning > 0) { /* lists are explicit; nothing mode-specific */ }
    if (n_events == 0) return DCT_OK;
    CK(cudaSetDevice(h->device));
    dct::ReplayArgs a;
    a.p = h->p;
    a.n_traces = (uint64_t)n_events * h->p.P;
    a.nsb_offsets = nsb_offsets; a.nsb_time = nsb_time; a.nsb_amp = nsb_amp;
    a.sig_time = sig_time; a.sig_amp = sig_amp; a.e_noise = e_noise; a.adc = adc;
    a.n_clamped = n_clamped;
    const uint64_t blocks = (a.n_traces + dct::REPLAY_WARPS - 1) / dct::REPLAY_WARPS;
    if (blocks > 0x7fffffffull) return fail(DCT_E_ARG, "too many traces for one launch");
    const size_t smem = (size_t)h->p.n_iv * 4 * (f64 ? sizeof(double) : sizeof(float));
    cudaStream_t st = (cudaStream_t)stream;
    if (f64) {
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(dct::k_replay<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dct::k_replay<double><<<(unsigned)blocks, dct::REPLAY_WARPS * 32, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(dct::k_replay<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dct::k_replay<float><<<(unsigned)blocks, dct::REPLAY_WARPS * 32, smem, st>>>(a);
    }
    CK(cudaGetLastError());
    return DCT_OK;
}

int dct_replay_f64(dct_handle* h, uint32_t n_events, const int64_t* nsb_offsets, const double* nsb_time,
                   const double* nsb_amp, const double* sig_time, const double* sig_amp,
                   const double* e_noise, int16_t* adc, unsigned long long* n_clamped, void* stream) {
    return replay_common(true, h, n_events, nsb_offsets, nsb_time, nsb_amp, sig_time, sig_amp, e_noise,
                         adc, n_clamped, stream);
}

int dct_replay_f32(dct_handle* h, uint32_t n_events, const int64_t* nsb_offsets, const double* nsb_time,
                   const double* nsb_amp, const double* sig_time, const double* sig_amp,
                   const double* e_noise, int16_t* adc, unsigned long long* n_clamped, void* stream) {
    return replay_common(false, h, n_events, nsb_offsets, nsb_time, nsb_amp, sig_time, sig_amp, e_noise,
                         adc, n_clamped, stream);
}

int dct_stats_accumulate(dct_handle* h, const int16_t* adc, uint32_t n_events,
                         unsigned long long* adc_hist, double* bin_sum, double* bin_sumsq,
                         double* pixel_sum, double* pixel_sumsq, unsigned long long* charge_hist,
                         int32_t charge_lo, uint32_t n_charge_bins, double* moments, void* stream) {
    if (!h || !adc) return fail(DCT_E_ARG, "handle or adc is NULL");
    if (n_events == 0) return DCT_OK;
    CK(cudaSetDevice(h->device));
    dct::StatsArgs a;
    a.P = h->p.P; a.B = h->p.B; a.adc_min = h->p.adc_min; a.adc_max = h->p.adc_max;
    a.n_traces = (uint64_t)n_events * h->p.P;
    a.adc = adc; a.baseline = h->p.baseline;
    a.adc_hist = adc_hist; a.bin_sum = bin_sum; a.bin_sumsq = bin_sumsq;
    a.pixel_sum = pixel_sum; a.pixel_sumsq = pixel_sumsq; a.charge_hist = charge_hist;
    a.charge_lo = charge_lo; a.n_charge_bins = (int)n_charge_bins; a.moments = moments;
    const int n_hist = a.adc_max - a.adc_min + 1;
    if (adc_hist && n_hist > 16384) return fail(DCT_E_UNSUPPORTED, "ADC histogram wider than 16384 bins");
    if (charge_hist && (a.B > 32 * dct::STATS_BINS_PER_LANE || n_charge_bins == 0))
        return fail(DCT_E_UNSUPPORTED, "charge histogram needs <= 128 bins per trace and n_charge_bins > 0");
    // the tile kernel derives the power sums from the histogram, so it needs the histogram buffer;
    // every event's block of rows must start on a 16-byte boundary for cp.async
    const int wpr = a.B / 2;
    const bool fast = (a.B % 2 == 0) && a.B <= 128 && (reinterpret_cast<uintptr_t>(adc) & 15) == 0 &&
                      ((uint64_t)a.P * wpr) % 4 == 0 && ((uint64_t)dct::STATS_ROWS * wpr) % 4 == 0 &&
                      (adc_hist != nullptr || moments == nullptr) && n_hist <= 16384 &&
                      // 32-bit sums of squares: 32 rows per column sum, B samples per trace
                      std::max(std::abs(a.adc_min), std::abs(a.adc_max)) <= 11585 &&
                      (uint64_t)a.B * (uint64_t)std::max(std::abs(a.adc_min), std::abs(a.adc_max)) *
                              (uint64_t)std::max(std::abs(a.adc_min), std::abs(a.adc_max)) < (1ull << 32);
    if (fast) {
        // histogram windows: centred on the expected pedestal (true_baseline, host-computed) and on
        // the charge of a trace that holds nothing but that pedestal shift
        int win_lo = stats_win_lo(h, dct::STATS_WIN);
        // Cubes with a night-sky pedestal spread far beyond the 16 thread-private values (C4: sigma ~ 30 LSB)
        // would split every warp between the private columns and the CTA histogram, sample by sample:
        // there every sample goes to the CTA histogram (a window no value can reach), one path for the warp.
        const int ped_lo = win_lo;                     // where the pedestal is: the compact CTA histogram is placed around it
        if ((h->mean_true_baseline - h->mean_baseline) > DCT_STATS_WIDE_SHIFT) win_lo = 0x3fffffff;
        // Quiet cubes (no pulse taller than ~100 LSB and an NSB pedestal shift below 3 LSB: dark runs) keep everything within
        // a few hundred values of the pedestal: a 512-value histogram slice and a 256-bin charge window
        // leave room for one more CTA per SM (measured on dark.yml: 2.45 -> 2.20 ms per 16384 events; the
        // same sizes on C4, whose pulses reach the rail, cost 10 %: 1.91 -> 2.13 ms).
        const bool quiet = h->max_pulse_lsb < 100.0 && (h->mean_true_baseline - h->mean_baseline) < 3.0;   // and hardly any NSB
        const int n_charge_s = quiet ? std::min(256, dct::STATS_CHARGE_WIN) : dct::STATS_CHARGE_WIN;
        const int compact_hist = quiet ? std::min(512, DCT_STATS_COMPACT_HIST) : DCT_STATS_COMPACT_HIST;
        const double q0 = a.B * (h->mean_true_baseline - h->mean_baseline) - (double)charge_lo;
        int charge_win_lo = (int)std::floor(q0) - n_charge_s / 2;
        charge_win_lo = std::max(0, std::min(charge_win_lo, std::max(0, (int)n_charge_bins - n_charge_s)));
        // Long traces: the two tile buffers plus a full-range CTA histogram would leave two CTAs per
        // SM.  Compact layout: one tile buffer, CTA histogram over a window around the pedestal
        // (samples beyond it go to the global histogram one atomic each: bright pulses only).
        const bool compact = a.B > DCT_STATS_COMPACT_BINS && n_hist > compact_hist;
        const int single_tile = compact ? DCT_STATS_COMPACT_SINGLE : 0;
        const int n_hist_s = compact ? compact_hist : n_hist;
        int hist_lo = a.adc_min;
        if (compact) {
            hist_lo = ped_lo - compact_hist / 8;
            hist_lo = std::max(a.adc_min, std::min(hist_lo, a.adc_max + 1 - n_hist_s));
        }
        const size_t tile_words = ((size_t)dct::STATS_ROWS * wpr + 3) & ~(size_t)3;
        const size_t smem = ((single_tile ? 1 : 2) * tile_words + (size_t)n_hist_s + (size_t)n_charge_s +
                             (size_t)dct::STATS_PRIV_WORDS) * sizeof(unsigned int);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(dct::k_stats_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(220 * 1024) / smem));
        const uint64_t n_pb = ((uint64_t)a.P + dct::STATS_ROWS - 1) / dct::STATS_ROWS;
        uint64_t blocks = (uint64_t)h->sm_count * per_sm;
        // (pixel block, event slice) work items: ~2 per CTA for balance, never more slices than events
        uint64_t n_slices = (2 * blocks + n_pb - 1) / n_pb;
        n_slices = std::max<uint64_t>(1, std::min<uint64_t>(n_slices, n_events));
        blocks = std::min<uint64_t>(blocks, n_pb * n_slices);
        dct::k_stats_tile<<<(unsigned)blocks, dct::STATS_ROWS, smem, (cudaStream_t)stream>>>(
            a, win_lo, (int)n_slices, charge_win_lo, hist_lo, n_hist_s, single_tile, n_charge_s);
        CK(cudaGetLastError());
        return DCT_OK;
    }
    const size_t smem = adc_hist ? (size_t)n_hist * sizeof(unsigned int) : 0;
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(dct::k_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t warps_needed = a.n_traces;
    uint64_t blocks = (warps_needed + (dct::STATS_THREADS / 32) - 1) / (dct::STATS_THREADS / 32);
    blocks = std::min<uint64_t>(blocks, (uint64_t)h->sm_count * 8);
    dct::k_stats<<<(unsigned)blocks, dct::STATS_THREADS, smem, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return DCT_OK;
}

int dct_write_calibrate(int16_t* buf, size_t n, void* stream) {
    if (!buf) return fail(DCT_E_ARG, "buf is NULL");
    if (reinterpret_cast<uintptr_t>(buf) & 15) return fail(DCT_E_ARG, "buf must be 16-byte aligned");
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t n_vec = n / 8, n_tail = n - n_vec * 8;
    dct::k_write_calibrate<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<int4*>(buf), n_vec, buf + n_vec * 8, n_tail);
    CK(cudaGetLastError());
    return DCT_OK;
}

int dct_trace_pe(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events, uint32_t max_pe,
                 int32_t* n_pe, double* pe_time, float* pe_amp, float* noise, void* stream) {
    if (!h || !n_pe) return fail(DCT_E_ARG, "handle or n_pe is NULL");
    if (max_pe > 0 && (!pe_time || !pe_amp)) return fail(DCT_E_ARG, "pe_time / pe_amp is NULL");
    if (n_events == 0) return DCT_OK;
    CK(cudaSetDevice(h->device));
    dct::DumpArgs d;
    d.g.p = h->p;
    d.g.k0 = (uint32_t)seed; d.g.k1 = (uint32_t)(seed >> 32);
    d.g.rk = dct::make_round_keys(d.g.k0, d.g.k1);
    d.g.first_event = first_event;
    d.g.n_traces = (uint64_t)n_events * (uint64_t)h->p.P;
    d.g.adc = nullptr; d.g.sig_time = nullptr; d.g.sig_charge = nullptr; d.g.table_in_smem = 0;
    d.max_pe = max_pe; d.n_pe = n_pe; d.pe_time = pe_time; d.pe_amp = pe_amp; d.noise = noise;
    const uint64_t blocks = (d.g.n_traces + 127) / 128;
    if (blocks > 0x7fffffffull) return fail(DCT_E_ARG, "too many traces for one launch");
    dct::k_dump_pe<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d);
    CK(cudaGetLastError());
    return DCT_OK;
}

int dct_test_chunk_schedule(uint32_t unit, uint32_t n_events, uint32_t* ends, uint32_t cap) {
    if (!ends || unit == 0) { fail(DCT_E_ARG, "ends is NULL or unit is 0"); return -1; }
    const std::vector<uint32_t> e = host_chunk_schedule(unit, n_events);
    for (size_t i = 0; i < e.size() && i < cap; ++i) ends[i] = e[i];
    return (int)e.size();
}

int dct_test_unpack12(const unsigned char* packed, int16_t* out, uint64_t n_samples, uint32_t threads) {
    if (!packed || !out) return fail(DCT_E_ARG, "packed or out is NULL");
    if (threads == 0) { dct::unpack12_scalar(packed, out, (size_t)n_samples); return DCT_OK; }
    dct::WorkerPool pool((int)std::min<uint32_t>(threads, 64));
    dct::unpack12_parallel(pool, packed, out, (size_t)n_samples);
    return DCT_OK;
}

int dct_test_sample(uint32_t kind, double a, double b, uint64_t seed, uint32_t n, float* out, void* stream) {
    if (!out) return fail(DCT_E_ARG, "out is NULL");
    if (n == 0) return DCT_OK;
    dct::k_test_sample<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        kind, (float)a, (float)b, (uint32_t)seed, (uint32_t)(seed >> 32), n, out);
    CK(cudaGetLastError());
    return DCT_OK;
}

}  // extern "C"
