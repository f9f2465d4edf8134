/* digicamtoy_b200 -- C ABI of the B200 (sm_100a) trace-generation library.
 *
 * The reference (cta-sst-1m/digicamtoy) is pure Python and has no FFI layer; the interface it
 * exposes for this path is the class NTraceGenerator
 * (reference digicamtoy/tracegenerator/__init__.py:38-381).  This header is the boundary a
 * maintainer of the reference would bind with ctypes to replace the body of that class
 * (INTEGRATION.md shows the stub).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C, no C++ or torch types; every pointer argument says "host" or "device".
 *   - return value 0 = success, negative = error (DCT_E_*); text via dct_last_error()
 *     (thread-local).  Nothing throws or aborts.
 *   - a handle is immutable after dct_create and may be shared between threads; launches are
 *     asynchronous on the caller's stream unless the name ends in _host.
 *   - the library never falls back to a CPU implementation: without a CUDA device every call
 *     returns DCT_E_CUDA.
 *   - results are a pure function of (config, seed, global event index, pixel): independent of
 *     launch geometry, chunking and of how events are split over GPUs.
 */
#ifndef DIGICAMTOY_B200_H
#define DIGICAMTOY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCT_ABI_VERSION 3

#if defined(__GNUC__)
#define DCT_API __attribute__((visibility("default")))
#else
#define DCT_API
#endif

enum {
    DCT_OK = 0,
    DCT_E_ARG = -1,      /* bad argument / inconsistent config            */
    DCT_E_CUDA = -2,     /* CUDA runtime error (message has the detail)    */
    DCT_E_UNSUPPORTED = -3,
    DCT_E_NOMEM = -4
};

/* Kernel selection for dct_generate.  DCT_KERNEL_AUTO: the grid kernel for sub_binning > 0 (at most
 * 24 on-grid template taps); else, for the 22-tap / 8-phase polyphase table with continuous NSB, the
 * tensor kernel at high NSB rates and the sweep kernel below; else the scatter kernel. */
enum {
    DCT_KERNEL_AUTO = 0,
    DCT_KERNEL_SCATTER = 1,  /* shared-memory accumulators, any template / sampling           */
    DCT_KERNEL_SWEEP = 2,    /* time-ordered register window (default template at 4 ns sampling) */
    DCT_KERNEL_GRID = 3,     /* sub_binning > 0 only: fixed on-grid taps, register window          */
    DCT_KERNEL_TENSOR = 4    /* high NSB rates: per-cell pe moments x coefficient tile on tcgen05.mma
                                (default template at 4 ns sampling, even n_bins <= 100)              */
};

/* Everything NTraceGenerator.__init__ derives (reference :52-157), already transformed on the
 * host in float64: gain is the *effective* gain after the NSB model (:112-128), sigma_1 is
 * already divided by it (:136), the template knots are already normalised (:139-154).
 * All arrays are HOST pointers; dct_create copies them to the device. */
typedef struct dct_config {
    uint32_t struct_size;       /* = sizeof(dct_config), ABI guard                              */
    uint32_t n_pixels;          /* P                                                            */
    uint32_t n_pulses;          /* K  signal pulses per pixel (columns of n_photon, >= 1)       */
    uint32_t n_bins;            /* B  kept bins written per trace                               */
    uint32_t n_drop;            /* simulated bins before the first kept one: 40 // dt  (:375)   */
    uint32_t n_knots;           /* template knots                                               */
    uint32_t poisson;           /* 1: N ~ Poisson(n_photon) (:236), 0: N = round(n_photon)      */
    uint32_t sub_binning;       /* 0: continuous NSB (:242-257), >0: on-grid NSB (:259-267)     */
    int32_t  adc_min, adc_max;  /* saturation of the digitiser (new; reference has none, :381)  */
    uint32_t kernel;            /* DCT_KERNEL_*                                                 */
    uint32_t reserved;
    double   t0;                /* time of simulated bin 0 = time_start - 40 ns (:96)           */
    double   t_end;             /* end of the NSB window (:97)                                  */
    double   dt;                /* time_sampling (:98)                                          */
    const double* nsb_rate;     /* [P] GHz                                                      */
    const double* crosstalk;    /* [P]                                                          */
    const double* gain;         /* [P] LSB / p.e., effective                                    */
    const double* sigma_1;      /* [P] relative (sigma_1 / gain)                                */
    const double* sigma_e;      /* [P] LSB                                                      */
    const double* baseline;     /* [P] LSB                                                      */
    const double* n_photon;     /* [P,K]                                                        */
    const double* time_signal;  /* [P,K] ns                                                     */
    const double* jitter;       /* [P,K] ns, full width of the uniform jitter (:228-235)        */
    const double* knot_x;       /* [n_knots] ns, uniformly spaced, ascending                    */
    const double* coef;         /* [n_knots-1][4] c0..c3 of the not-a-knot cubic on each
                                   interval, local variable u = x - knot_x[i]  (:152-154)       */
} dct_config;

typedef struct dct_handle dct_handle;

/* Library identity. */
DCT_API int         dct_abi_version(void);
DCT_API const char* dct_last_error(void);

/* replaces NTraceGenerator.__init__ / reset (:40-169, :190-199): builds the device parameter
 * block, the float32 polyphase template table and the float64 table on CUDA device `device`. */
DCT_API int  dct_create(const dct_config* cfg, int device, dct_handle** out);
DCT_API void dct_destroy(dct_handle* h);

/* Which kernel dct_generate will use for this config (DCT_KERNEL_SCATTER / _SWEEP / _GRID / _TENSOR). */
DCT_API int dct_selected_kernel(const dct_handle* h);

/* replaces n_events calls of NTraceGenerator.next() (:201-226 -> stages :228-381) for global
 * events [first_event, first_event + n_events).
 *   adc        device int16 [n_events, P, B], values in [adc_min, adc_max]    (adc_count, :381)
 *   sig_time   device float [n_events, P, K] or NULL                          (cherenkov_time)
 *   sig_charge device float [n_events, P, K] or NULL                          (cherenkov_photon)
 *   stream     cudaStream_t passed as void* (NULL = default stream)                            */
DCT_API int dct_generate(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                 int16_t* adc, float* sig_time, float* sig_charge, void* stream);

/* Test hook (not on the product path): dct_generate that also writes the analog sums every sample
 * is digitised from -- sum over pe of amplitude x template, in p.e., before gain / noise / baseline
 * (the quantity of compute_analog_signal :350-362 divided by the gain).
 *   analog     device float [n_events, P, B]                                                   */
DCT_API int dct_generate_analog(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                        int16_t* adc, float* analog, void* stream);

/* Same, into HOST buffers (pageable or pinned): generates in chunks on an internal stream pair and
 * overlaps the device->host copy of one chunk with the generation of the next.  Synchronous. */
DCT_API int dct_generate_host(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                      int16_t* adc_host, float* sig_time_host, float* sig_charge_host);

/* replaces compute_analog_signal + generate_electronic_smearing + convert_to_digital (:350-381)
 * for photo-electron lists recorded from the reference itself (deterministic; parity gate).
 *   nsb_offsets device int64 [n_events*P + 1]  CSR offsets into nsb_time / nsb_amp
 *   nsb_time    device double [nnz]  pe times (ns), reference order
 *   nsb_amp     device double [nnz]  smeared amplitude x mask (:354)
 *   sig_time    device double [n_events, P, K]
 *   sig_amp     device double [n_events, P, K]
 *   e_noise     device double [n_events, P, B] noise of the kept bins, or NULL
 *   adc         device int16  [n_events, P, B]  clamped to [adc_min, adc_max]
 *   n_clamped   device uint64 [1], incremented once per sample whose unclamped value was outside
 *               the range (those are the samples on which reference and build differ by design)
 * _f64 evaluates and accumulates in double (expected: zero mismatches), _f32 in float (the
 * arithmetic the stochastic kernels use; expected: <= 1 LSB on rounding ties).                 */
DCT_API int dct_replay_f64(dct_handle* h, uint32_t n_events, const int64_t* nsb_offsets,
                   const double* nsb_time, const double* nsb_amp, const double* sig_time,
                   const double* sig_amp, const double* e_noise, int16_t* adc,
                   unsigned long long* n_clamped, void* stream);
DCT_API int dct_replay_f32(dct_handle* h, uint32_t n_events, const int64_t* nsb_offsets,
                   const double* nsb_time, const double* nsb_amp, const double* sig_time,
                   const double* sig_amp, const double* e_noise, int16_t* adc,
                   unsigned long long* n_clamped, void* stream);

/* Validation statistics of a cube (the quantities of reference nsb_baseline.py:21-32 and the ADC /
 * charge histograms), accumulated on the device so that sharded runs only exchange these.
 * Every output may be NULL; all are device pointers and are ADDED to (zero them first).
 *   adc_hist     uint64 [adc_max - adc_min + 1]
 *   bin_sum      double [B]      sum over events and pixels of adc, per bin
 *   bin_sumsq    double [B]
 *   pixel_sum    double [P]      sum over events and bins, per pixel
 *   pixel_sumsq  double [P]
 *   charge_hist  uint64 [n_charge_bins]  histogram of (sum_b adc - B*round(baseline_p)) per trace,
 *                bin = clamp(charge - charge_lo, 0, n_charge_bins-1)
 *   moments      double [4]  raw power sums  sum x, sum x^2, sum x^3, sum x^4 of all samples     */
DCT_API int dct_stats_accumulate(dct_handle* h, const int16_t* adc, uint32_t n_events,
                         unsigned long long* adc_hist, double* bin_sum, double* bin_sumsq,
                         double* pixel_sum, double* pixel_sumsq, unsigned long long* charge_hist,
                         int32_t charge_lo, uint32_t n_charge_bins, double* moments, void* stream);

/* dct_generate with the statistics of dct_stats_accumulate taken inside the generation kernel, from the
 * digitised samples in shared memory just before they are stored: same results bit for bit (power
 * sums: to float64 rounding), no second pass over the cube.  adc_hist, bin_sum, bin_sumsq, pixel_sum
 * and pixel_sumsq are required; charge_hist and moments may be NULL.  All buffers accumulate.
 * Replaces the host-side analysis of reference nsb_baseline.py:21-32 for cubes that are generated
 * and analysed in one go. */
DCT_API int dct_generate_stats(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                       int16_t* adc, float* sig_time, float* sig_charge,
                       unsigned long long* adc_hist, double* bin_sum, double* bin_sumsq,
                       double* pixel_sum, double* pixel_sumsq, unsigned long long* charge_hist,
                       int32_t charge_lo, uint32_t n_charge_bins, double* moments, void* stream);

/* Write-only calibration kernel: streams n int16 values to buf with the same 16-byte stores the
 * generator uses.  Its GB/s is the practical ceiling of the HBM-write roofline. */
DCT_API int dct_write_calibrate(int16_t* buf, size_t n, void* stream);

/* The photo-electron lists and noise values behind the traces of events [first_event, +n_events):
 * what the reference leaves in nsb_time / nsb_photon (x mask) after next() (:248-257, :345-348).
 * Counter-based RNG: this re-derives exactly what dct_generate consumed for the same
 * (seed, event, pixel).  Call once with max_pe = 0 to get the counts, then with max_pe >= max count.
 *   n_pe     device int32  [n_events, P]           clusters per trace (always written)
 *   pe_time  device double [n_events, P, max_pe]   absolute time in ns       (first n_pe valid)
 *   pe_amp   device float  [n_events, P, max_pe]   smeared cluster amplitude (first n_pe valid)
 *   noise    device float  [n_events, P, B]        electronic noise of the kept bins, or NULL   */
DCT_API int dct_trace_pe(dct_handle* h, uint64_t seed, uint64_t first_event, uint32_t n_events,
                         uint32_t max_pe, int32_t* n_pe, double* pe_time, float* pe_amp, float* noise,
                         void* stream);

/* Test hook (not on the product path): n draws of one device sampler into device float out[n].
 * kind: 0 normal, 1 exponential gap (a = mean), 2 Borel(a), 3 Poisson(a), 4 total progeny of
 * b primaries with Poisson(a) offspring, 5 uniform [0,1). */
DCT_API int dct_test_sample(uint32_t kind, double a, double b, uint64_t seed, uint32_t n, float* out,
                            void* stream);

/* Test hook, host only (no device needed): widens n_samples of the 12-bit transport stream of
 * dct_generate_host (little-endian, sample s in bits [12 s, 12 s + 12)) into int16, on `threads`
 * host threads (0: the scalar loop on the calling thread). */
DCT_API int dct_test_unpack12(const unsigned char* packed, int16_t* out, uint64_t n_samples, uint32_t threads);

/* Test hook, host only: the chunk ends [events] dct_generate_host would use for a call of n_events with a
 * chunk unit of `unit` events (sizes 1, 2, 4 ... 4, rest, 2, 1 units).  Writes at most `cap` ends to
 * `ends` and returns the number of chunks (negative on error). */
DCT_API int dct_test_chunk_schedule(uint32_t unit, uint32_t n_events, uint32_t* ends, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DIGICAMTOY_B200_H */
