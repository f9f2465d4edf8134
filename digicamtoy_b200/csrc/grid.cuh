// K1c: generation kernel for the on-grid NSB mode (sub_binning > 0).
//
// Reference digicamtoy/tracegenerator/__init__.py:259-267: n ~ Poisson(rate * dt) photo-electrons
// sit exactly on every simulated sample, so the pulse template is only ever needed at multiples of
// the sampling period: a fixed FIR (23 taps for the 88 ns template at 4 ns).  The scatter kernel
// does that as 23 dependent read-modify-writes of the shared-memory row per bin; here the taps are
// kernel parameters (constant-bank operands of the FMAs) and the partial sums of the next GRID_TAPS
// bins live in registers: after simulated bin bs has been drawn, W[0] is final, is retired to the
// row, and the window moves one bin -- the same for every lane, no predication.
//
// Random numbers and the order of the float32 FMAs into every output bin are those of the scatter
// kernel's sub_binning branch (generate.cuh), so the two kernels produce identical cubes.
#pragma once
#include "generate.cuh"

namespace dct {

constexpr int GRID_TAPS = 24;                // window length; templates with more on-grid taps use K1b
constexpr int GRID_THREADS = 128;

struct GridArgs {
    GenArgs g;
    float tap[GRID_TAPS];                    // h((grid_m_lo + j) * dt), zero beyond grid_n
};

__host__ __device__ inline size_t grid_smem_bytes(int B) {    // rows + one Poisson table per trace
    return gen_smem_bytes(B, 0, GRID_THREADS) + (size_t)GRID_THREADS * POISSON_TABLE * sizeof(float);
}

__host__ __device__ inline bool grid_supported(const DevParams& p) {
    return p.sub_binning > 0 && p.grid_n >= 1 && p.grid_n <= GRID_TAPS && p.grid_m_lo >= 0 &&
           grid_smem_bytes(p.B) <= 227 * 1024;
}

#ifndef DCT_GRID_MIN_BLOCKS
#define DCT_GRID_MIN_BLOCKS 5
#endif

template <int NT>
__global__ void __launch_bounds__(NT, DCT_GRID_MIN_BLOCKS)
k_generate_grid(const GridArgs ga) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GenArgs& a = ga.g;
    const DevParams& p = a.p;
    const int tid = threadIdx.x;
    const int B = p.B;
    float* rows = reinterpret_cast<float*>(smem_raw);
    float* row = rows + tid * row_stride(B);
    float* cdf = rows + NT * row_stride(B) + tid;            // this trace's Poisson table, stride NT
    for (int b = 0; b < B; ++b) row[b] = 0.0f;
    __syncthreads();

    const uint64_t trace0 = (uint64_t)blockIdx.x * NT;
    if (trace0 + tid < a.n_traces) {
        const TraceId t = trace_id(a, trace0 + tid);
        const float rate = p.rate[t.pixel];
        const PixelNsb q = load_pixel_nsb(p, t.pixel);
        if (rate > 0.0f) {
            float W[GRID_TAPS];                              // W[j] <-> simulated bin bs + grid_m_lo + j
#pragma unroll
            for (int j = 0; j < GRID_TAPS; ++j) W[j] = 0.0f;
            const float lam = rate * p.dt, p0 = __expf(-lam);
            if (lam < 12.0f) poisson_table_fill(cdf, NT, lam, p0);   // (only this thread reads it)
            const int n_sim = p.n_drop + B;
            int kb = p.grid_m_lo - p.n_drop;                 // kept index of W[0]
            for (int bs = 0; bs < n_sim; ++bs, ++kb) {
                float A;
                if (subbin_cluster(t.key, (uint32_t)bs, lam, p0, q, A, cdf, NT)) {
#pragma unroll
                    for (int j = 0; j < GRID_TAPS; ++j) W[j] = __fmaf_rn(A, ga.tap[j], W[j]);
                }
                if ((unsigned)kb < (unsigned)B) row[kb] = W[0];
#pragma unroll
                for (int j = 0; j < GRID_TAPS - 1; ++j) W[j] = W[j + 1];
                W[GRID_TAPS - 1] = 0.0f;
            }
            // what is left in the window belongs to bins past the end of the trace
        }
        add_signal_pulses(a, t, p.coef32, row, q.xt, q.sigma1);
        digitise_trace(p, a.rk, t, row);
    }
    __syncthreads();
    if (DCT_FUSED_STATS && a.st.adc_hist)
        cta_fused_stats<NT, 1>(a.st, p.B, p.P, p.adc_min, p.adc_max, p.baseline, trace0, a.n_traces, rows, smem_raw + stats_offset(grid_smem_bytes(B)));
    store_run<NT>(a, trace0, rows);
}

inline int launch_grid(int device, const GenArgs& a, const float* taps, cudaStream_t st, char* err, size_t err_len) {
    auto kern = k_generate_grid<GRID_THREADS>;
    const size_t smem = smem_with_stats(grid_smem_bytes(a.p.B), a, GRID_THREADS);
    static thread_local size_t configured[16] = {0};
    if (smem > 48 * 1024 && configured[device & 15] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { snprintf(err, err_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        configured[device & 15] = smem;
    }
    const uint64_t blocks = (a.n_traces + GRID_THREADS - 1) / GRID_THREADS;
    if (blocks > 0x7fffffffull) { snprintf(err, err_len, "too many traces for one launch"); return -1; }
    GridArgs ga;
    ga.g = a;
    for (int j = 0; j < GRID_TAPS; ++j) ga.tap[j] = taps[j];
    kern<<<(unsigned)blocks, GRID_THREADS, smem, st>>>(ga);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, err_len, "grid kernel launch failed: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace dct
