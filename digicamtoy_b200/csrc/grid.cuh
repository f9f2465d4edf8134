// K1c: generation kernel for the on-grid NSB mode (sub_binning > 0).
//
// Reference digicamtoy/tracegenerator/__init__.py:259-267: n ~ Poisson(rate * dt) photo-electrons
// sit exactly on every simulated sample, so the pulse template is only ever needed at multiples of
// the sampling period: a fixed FIR (23 taps for the 88 ns template at 4 ns).  The scatter kernel
// does that as 23 dependent read-modify-writes of the shared-memory row per bin; here the taps are
// kernel parameters (constant-bank operands of the FMAs) and the partial sums of the next GRID_TAPS
// bins live in registers: after simulated bin bs has been drawn its own partial sum is final and is
// retired to the row -- the same for every lane, no predication.  The loop over the bins is unrolled
// GRID_UNROLL times and the window moves by that many registers per pass; the number of pe on a bin
// (primaries + crosstalk progeny) comes from one table look-up, two bins share a Philox block.
//
// Random numbers and the order of the float32 FMAs into every output bin are those of the scatter
// kernel's sub_binning branch (generate.cuh), so the two kernels produce identical cubes.
#pragma once
#include "generate.cuh"

namespace dct {

constexpr int GRID_TAPS = 24;                // window length; templates with more on-grid taps use K1b
constexpr int GRID_THREADS = 128;
#ifndef DCT_GRID_UNROLL
#define DCT_GRID_UNROLL 2
#endif
constexpr int GRID_UNROLL = DCT_GRID_UNROLL;   // simulated bins per pass of the bin loop (even: two bins share a Philox block)

struct GridArgs {
    GenArgs g;
    float tap[GRID_TAPS];                    // h((grid_m_lo + j) * dt), zero beyond grid_n
};

__host__ __device__ inline size_t grid_smem_bytes(int B) {    // rows + one table of cumulative probabilities per trace
    return gen_smem_bytes(B, 0, GRID_THREADS) + (size_t)GRID_THREADS * GP_TABLE * sizeof(float);
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
            const float lam = rate * p.dt, p0 = __expf(-lam);
            const int n_sim = p.n_drop + B;
            const int kb0 = p.grid_m_lo - p.n_drop;          // kept index of the bin retired at simulated bin 0
            if (lam < GP_LAM_MAX && p.gp_cdf != nullptr) {
                // the pixel's generalised-Poisson table -> this trace's shared-memory column (stride NT)
                const float4* src = reinterpret_cast<const float4*>(p.gp_cdf + (size_t)t.pixel * GP_TABLE);
#pragma unroll
                for (int i = 0; i < GP_TABLE / 4; ++i) {
                    const float4 c4 = src[i];
                    cdf[(4 * i + 0) * NT] = c4.x; cdf[(4 * i + 1) * NT] = c4.y;
                    cdf[(4 * i + 2) * NT] = c4.z; cdf[(4 * i + 3) * NT] = c4.w;
                }
                // GRID_UNROLL bins per pass of the outer loop: bin bs0 + s adds its taps to V[s + m] (static
                // indices), then V[s] is final and goes to the row; after the pass the window moves down by
                // GRID_UNROLL registers at once (GRID_TAPS moves per GRID_UNROLL bins instead of per bin).
                // A pass is short enough for the instruction cache; unrolling all GRID_TAPS bins (pure
                // renaming, no moves) was measured: 39 % issue, stalled on instruction fetch.
                // Same FMAs in the same order into every output bin as the scatter kernel.
                float V[GRID_TAPS + GRID_UNROLL - 1];
#pragma unroll
                for (int j = 0; j < GRID_TAPS + GRID_UNROLL - 1; ++j) V[j] = 0.0f;
                u32x4 r = {0u, 0u, 0u, 0u};
                for (int bs0 = 0; bs0 < n_sim; bs0 += GRID_UNROLL) {
#pragma unroll
                    for (int s = 0; s < GRID_UNROLL; ++s) {
                        const int bs = bs0 + s;
                        if (bs < n_sim) {                    // (uniform across the CTA)
                            if ((s & 1) == 0) r = draw_block_rk(a.rk, t.key, STREAM_SUBBIN_GP, (uint32_t)bs >> 1);
                            float A = 0.0f;
                            subbin_total(r, (uint32_t)bs, lam, q, cdf, NT, A);
#pragma unroll
                            for (int m = 0; m < GRID_TAPS; ++m) V[s + m] = __fmaf_rn(A, ga.tap[m], V[s + m]);
                            const int kb = bs + kb0;
                            if ((unsigned)kb < (unsigned)B) row[kb] = V[s];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < GRID_TAPS + GRID_UNROLL - 1; ++j)
                        V[j] = (j + GRID_UNROLL < GRID_TAPS + GRID_UNROLL - 1) ? V[j + GRID_UNROLL] : 0.0f;
                }
            } else {
                // rate * dt >= 12 (3 GHz at 4 ns): primaries by PTRS, crosstalk generation by generation
                float W[GRID_TAPS];                          // partial sums of the next GRID_TAPS simulated bins
#pragma unroll
                for (int j = 0; j < GRID_TAPS; ++j) W[j] = 0.0f;
                int kb = kb0;
                for (int bs = 0; bs < n_sim; ++bs, ++kb) {
                    float A;
                    if (subbin_cluster(t.key, (uint32_t)bs, lam, p0, q, A)) {
#pragma unroll
                        for (int j = 0; j < GRID_TAPS; ++j) W[j] = __fmaf_rn(A, ga.tap[j], W[j]);
                    }
                    if ((unsigned)kb < (unsigned)B) row[kb] = W[0];
#pragma unroll
                    for (int j = 0; j < GRID_TAPS - 1; ++j) W[j] = W[j + 1];
                    W[GRID_TAPS - 1] = 0.0f;
                }
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
