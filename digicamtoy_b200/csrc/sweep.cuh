// K1, high-rate variant: time-ordered sweep with a register window.
//
// The NSB photo-electrons of a trace come out of the exponential-gap generator already sorted
// in time.  A pe whose onset lies in sampling cell c only touches bins c+1 .. c+TAPS, so a thread
// keeps exactly those TAPS partial sums in registers (statically indexed), adds each pe with one
// 16-byte shared-memory load + 4 FFMA per tap (polyphase table, immediate offsets), and when the
// onset moves to the next cell it retires the oldest bin to shared memory and slides the window.
// No per-tap address arithmetic, no read-modify-write of shared memory.
//
// Divergence control (ncu, profiles/r01): the one-cell slide is a predicated select executed by
// every lane every iteration (22 FSEL) instead of a branchy loop that ran with 7 of 32 lanes; only
// gaps longer than one cell (0.5 % of pe at 1 GHz) take the loop.  Optional (DCT_SWEEP_SKIP=1, off):
// taps that can only land in dropped or non-existent bins for every lane of the warp are skipped
// in groups of four using the warp-wide min / max of the cell index (REDUX) -- a win with scalar
// FFMA taps, a loss once the taps became FFMA2 (DESIGN.md tuning record).
//
// It consumes the same Philox blocks and performs the same floating-point operations per kept bin,
// in the same order, as k_generate_scatter: both kernels produce identical cubes (tests check it).
//
// Replaces the dense (pixel, pe, bin) spline evaluation of compute_analog_signal
// (reference digicamtoy/tracegenerator/__init__.py:350-362).
#pragma once
#include "generate.cuh"

namespace dct {

constexpr int SWEEP_TAPS = 22;            // 88 ns template / 4 ns sampling
constexpr int SWEEP_PHASES = 8;           // 4 ns sampling / 0.5 ns knots
constexpr int SWEEP_THREADS = 128;
constexpr int SWEEP_MAX_BINS = 256;
#ifndef DCT_SWEEP_GROUP
#define DCT_SWEEP_GROUP 4
#endif
// cost-attribution experiments (profiles/r01/cost_attribution.txt); all produce WRONG cubes
#ifndef DCT_EXP_NO_SLIDE
#define DCT_EXP_NO_SLIDE 0
#endif
#ifndef DCT_EXP_NO_TAPS
#define DCT_EXP_NO_TAPS 0
#endif
#ifndef DCT_SWEEP_SKIP
#define DCT_SWEEP_SKIP 0      // 1: skip tap groups no lane of the warp needs (REDUX min / max of the cell); measured slower once the taps became FFMA2
#endif
#ifndef DCT_SWEEP_MIN_BLOCKS
#define DCT_SWEEP_MIN_BLOCKS 5
#endif
constexpr int SWEEP_GROUP = DCT_SWEEP_GROUP;   // taps per skip group (even)
constexpr double SWEEP_MIN_MEAN_PE = 0.0;    // measured (scripts/kernel_crossover.py): sweep >= scatter at every rate, 0.7 .. 240 pe / trace

inline bool sweep_supported(const DevParams& p) {
    return p.n_phase == SWEEP_PHASES && p.n_taps == SWEEP_TAPS && p.sub_binning <= 0 &&
           p.B <= SWEEP_MAX_BINS &&
           gen_smem_bytes(p.B, SWEEP_TAPS * SWEEP_PHASES, SWEEP_THREADS) <= 227 * 1024;
}

// Packed FP32 FMA (Blackwell FFMA2): two IEEE fma.rn.f32 in one issue slot.  Bit-identical to two
// scalar __fmaf_rn, so the scatter kernel (scalar) and this kernel still agree exactly.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// W[j] <- W[j+1] (j < TAPS-1), W[TAPS-1] <- 0 for the lanes with `go`; a no-op for the others.
// Written as predicated moves on read-write operands so that every W[j] keeps its register:
// the C++ forms (select or branch) made ptxas keep two copies of the window and move all 22
// values back at the end of every iteration (profiles/r01: 22-44 extra MOV/FSEL per pe).
template <int TAPS>
__device__ __forceinline__ void slide_window(float (&W)[TAPS], bool go) {
    static_assert(TAPS == 22 && TAPS % 2 == 0, "slide_window is written out for 22 taps");
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %22, 0;\n\t"
        "@p mov.f32 %0, %1;\n\t@p mov.f32 %1, %2;\n\t@p mov.f32 %2, %3;\n\t@p mov.f32 %3, %4;\n\t"
        "@p mov.f32 %4, %5;\n\t@p mov.f32 %5, %6;\n\t@p mov.f32 %6, %7;\n\t@p mov.f32 %7, %8;\n\t"
        "@p mov.f32 %8, %9;\n\t@p mov.f32 %9, %10;\n\t@p mov.f32 %10, %11;\n\t@p mov.f32 %11, %12;\n\t"
        "@p mov.f32 %12, %13;\n\t@p mov.f32 %13, %14;\n\t@p mov.f32 %14, %15;\n\t@p mov.f32 %15, %16;\n\t"
        "@p mov.f32 %16, %17;\n\t@p mov.f32 %17, %18;\n\t@p mov.f32 %18, %19;\n\t@p mov.f32 %19, %20;\n\t"
        "@p mov.f32 %20, %21;\n\t@p mov.f32 %21, 0f00000000;\n\t}"
        : "+f"(W[0]), "+f"(W[1]), "+f"(W[2]), "+f"(W[3]), "+f"(W[4]), "+f"(W[5]), "+f"(W[6]), "+f"(W[7]),
          "+f"(W[8]), "+f"(W[9]), "+f"(W[10]), "+f"(W[11]), "+f"(W[12]), "+f"(W[13]), "+f"(W[14]),
          "+f"(W[15]), "+f"(W[16]), "+f"(W[17]), "+f"(W[18]), "+f"(W[19]), "+f"(W[20]), "+f"(W[21])
        : "r"((int)go));
}

// W[j] <- W[j+S] (j < TAPS-S), the last S <- 0, for the lanes with `go`: one predicated move per
// register, in place like slide_window (used for gaps of several cells: 2, 4, 8 and 16 at a time).
template <int S>
__device__ __forceinline__ void slide_by(float (&W)[22], bool go) {
    static_assert(S >= 1 && S < 22, "window of 22 taps");
    const int g = (int)go;
#pragma unroll
    for (int j = 0; j + S < 22; ++j)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p mov.f32 %0, %1;\n\t}" : "+f"(W[j]) : "f"(W[j + S]), "r"(g));
#pragma unroll
    for (int j = 22 - S; j < 22; ++j)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p mov.f32 %0, 0f00000000;\n\t}" : "+f"(W[j]) : "r"(g));
}

#ifndef DCT_SWEEP_BINARY_SLIDE
#define DCT_SWEEP_BINARY_SLIDE 1      // gaps of 2 .. 22 cells: slide by 2, 4, 8, 16 (binary digits of the gap) instead of cell by cell
#endif

template <int TAPS, int PHASES, int NT>
__global__ void __launch_bounds__(NT, DCT_SWEEP_MIN_BLOCKS)
k_generate_sweep(const GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& p = a.p;
    const int tid = threadIdx.x;
    const int B = p.B;

    float4* tab_s = reinterpret_cast<float4*>(smem_raw);
    float* rows = reinterpret_cast<float*>(tab_s + TAPS * PHASES);
    // pair-interleaved copy of the polyphase table for FFMA2: for tap pair m = (2m, 2m+1), phase f
    //   lo[m*PHASES + f] = (c0_a, c0_b, c1_a, c1_b)      hi[m*PHASES + f] = (c2_a, c2_b, c3_a, c3_b)
    float4* tab_hi = tab_s + (TAPS / 2) * PHASES;
    for (int i = tid; i < (TAPS / 2) * PHASES; i += NT) {
        const int m = i / PHASES, f = i - m * PHASES;
        const float4 ca = p.poly[(2 * m) * PHASES + f], cb = p.poly[(2 * m + 1) * PHASES + f];
        tab_s[i] = make_float4(ca.x, cb.x, ca.y, cb.y);
        tab_hi[i] = make_float4(ca.z, cb.z, ca.w, cb.w);
    }
    float* row = rows + tid * row_stride(B);
    for (int b = 0; b < B; ++b) row[b] = 0.0f;
    __syncthreads();

    const uint64_t trace0 = (uint64_t)blockIdx.x * NT;
    if (trace0 + tid < a.n_traces) {
        const TraceId t = trace_id(a, trace0 + tid);
        const float rate = p.rate[t.pixel];
        const PixelNsb q = load_pixel_nsb(p, t.pixel);

        if (rate > 0.0f) {
            float W[TAPS];                                   // W[j] <-> simulated bin cell+1+j
#pragma unroll
            for (int j = 0; j < TAPS; ++j) W[j] = 0.0f;
            float elapsed = 0.0f;
            int cell = p.cell0;
            float off = p.off0;
            const int kb_shift = 1 - p.n_drop;               // kept index of W[0] = cell + kb_shift
#if DCT_SWEEP_SKIP
            const int first_needed = p.n_drop - 1;           // tap j is kept iff j >= first_needed - cell
            const int last_needed = p.n_drop + B - 1;        //             and j <  last_needed  - cell
#endif
            // One photo-electron: slide the window to its cell, then add its taps.
            auto add_pe = [&](float gap, float A) {
                int kb = cell + kb_shift;
                int k = advance_onset(p, gap, cell, off);
                // slide by one cell: predicated, all lanes
                const bool one = k >= 1;
                if (one && (unsigned)kb < (unsigned)B) row[kb] = W[0];
#if !DCT_EXP_NO_SLIDE
                slide_window<TAPS>(W, one);
#endif
                if (k >= 2) {                                 // gap longer than a cell (rare at 1 GHz, the rule below 0.2 GHz)
                    ++kb;
                    if (k > TAPS) {
#pragma unroll
                        for (int j = 0; j < TAPS; ++j) {
                            if ((unsigned)(kb + j) < (unsigned)B) row[kb + j] = W[j];
                            W[j] = 0.0f;
                        }
                    } else {
#if DCT_SWEEP_BINARY_SLIDE
                        // k - 1 more cells, 1 .. 21 = binary digits: retire the S oldest bins, slide by S.
                        // At most five short steps per warp, where the cell-by-cell loop ran to the
                        // longest gap of the warp with two lanes active (profiles/r02/ncu_acdc_sweep_r02.txt).
                        const int m = k - 1;
#define DCT_SLIDE_STEP(S)                                                                   \
                        if (m & S) {                                                        \
                            _Pragma("unroll")                                               \
                            for (int j = 0; j < S; ++j)                                     \
                                if ((unsigned)(kb + j) < (unsigned)B) row[kb + j] = W[j];   \
                            slide_by<S>(W, true);                                           \
                            kb += S;                                                        \
                        }
                        DCT_SLIDE_STEP(1)
                        DCT_SLIDE_STEP(2)
                        DCT_SLIDE_STEP(4)
                        DCT_SLIDE_STEP(8)
                        DCT_SLIDE_STEP(16)
#undef DCT_SLIDE_STEP
#else
                        for (--k; k > 0; --k, ++kb) {
                            if ((unsigned)kb < (unsigned)B) row[kb] = W[0];
#pragma unroll
                            for (int j = 0; j < TAPS - 1; ++j) W[j] = W[j + 1];
                            W[TAPS - 1] = 0.0f;
                        }
#endif
                    }
                }
                int f; float u;
                phase_of(p, p.dt - off, f, u);
                const float4* tlo = tab_s + f;
                const float4* thi = tab_hi + f;
                const float2 uu = make_float2(u, u), AA = make_float2(A, A);
                // warp-uniform tap range: skip groups no lane needs
#if DCT_SWEEP_SKIP == 1
                const unsigned act = __activemask();
                const int j_lo = first_needed - __reduce_max_sync(act, cell);
                const int j_hi = last_needed - __reduce_min_sync(act, cell);
#elif DCT_SWEEP_SKIP == 2
                // only the end of the trace (where most of the unneeded taps are): one REDUX
                const int j_lo = 0;
                const int j_hi = last_needed - __reduce_min_sync(__activemask(), cell);
#else
                const int j_lo = 0, j_hi = TAPS;
#endif
#if !DCT_EXP_NO_TAPS
#pragma unroll
                for (int g0 = 0; g0 < TAPS; g0 += SWEEP_GROUP) {
                    if (g0 + SWEEP_GROUP > j_lo && g0 < j_hi) {
#pragma unroll
                        for (int j = g0; j < g0 + SWEEP_GROUP && j < TAPS; j += 2) {
                            const float4 lo = tlo[(j / 2) * PHASES], hi = thi[(j / 2) * PHASES];
                            float2 t = ffma2(make_float2(hi.z, hi.w), uu, make_float2(hi.x, hi.y));
                            t = ffma2(t, uu, make_float2(lo.z, lo.w));
                            t = ffma2(t, uu, make_float2(lo.x, lo.y));
                            const float2 w = ffma2(AA, t, make_float2(W[j], W[j + 1]));
                            W[j] = w.x;
                            W[j + 1] = w.y;
                        }
                    }
                }
#else
                W[0] += A * u + (float)(j_lo + j_hi) + tlo[0].x + thi[0].x + uu.x + AA.y;   // keep the inputs alive
#endif
            };
            // Four photo-electrons per three Philox blocks (NsbGroup), two per pass of the inner loop:
            // their samplers are independent and overlap; the window updates stay in order.
            bool more = true;
            for (uint32_t g = 0; more; ++g) {
                const NsbGroup grp = draw_nsb_group(a.rk, t.key, g);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ga = h ? grp.gap.z : grp.gap.x, gb = h ? grp.gap.w : grp.gap.y;
                    const uint32_t ba = h ? grp.bor.z : grp.bor.x, bb = h ? grp.bor.w : grp.bor.y;
                    float za, zb;
                    normal_pair(h ? grp.nrm.z : grp.nrm.x, h ? grp.nrm.w : grp.nrm.y, za, zb);
                    const float gap_a = exp_gap(ga, q.neg_ln2_over_rate);
                    const float gap_b = exp_gap(gb, q.neg_ln2_over_rate);
                    const float A_a = cluster_amplitude((float)borel(ba, q.xt, q.borel_cdf), za, q.sigma1);
                    const float A_b = cluster_amplitude((float)borel(bb, q.xt, q.borel_cdf), zb, q.sigma1);
                    elapsed += gap_a;
                    if (elapsed >= p.span) { more = false; break; }
                    add_pe(gap_a, A_a);
                    elapsed += gap_b;
                    if (elapsed >= p.span) { more = false; break; }
                    add_pe(gap_b, A_b);
                }
            }
            // retire what is left in the window
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
                const int kb = cell + kb_shift + j;
                if ((unsigned)kb < (unsigned)B) row[kb] = W[j];
            }
        }
        add_signal_pulses(a, t, p.coef32, row, q.xt, q.sigma1);
        dump_analog(a, t.g, row);
        digitise_trace(p, a.rk, t, row);
    }
    __syncthreads();
    if (DCT_FUSED_STATS && a.st.adc_hist)
        cta_fused_stats<NT, 1>(a.st, p.B, p.P, p.adc_min, p.adc_max, p.baseline, trace0, a.n_traces, rows,
                               smem_raw + stats_offset(gen_smem_bytes(B, TAPS * PHASES, NT)));
    store_run<NT>(a, trace0, rows);
}

inline int launch_sweep(int device, int /*sm_count*/, const GenArgs& a, cudaStream_t st, char* err, size_t err_len) {
    auto kern = k_generate_sweep<SWEEP_TAPS, SWEEP_PHASES, SWEEP_THREADS>;
    const size_t smem = smem_with_stats(gen_smem_bytes(a.p.B, SWEEP_TAPS * SWEEP_PHASES, SWEEP_THREADS), a, SWEEP_THREADS);
    static thread_local size_t configured[16] = {0};
    if (smem > 48 * 1024 && configured[device & 15] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { snprintf(err, err_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        configured[device & 15] = smem;
    }
    const uint64_t blocks = (a.n_traces + SWEEP_THREADS - 1) / SWEEP_THREADS;
    if (blocks > 0x7fffffffull) { snprintf(err, err_len, "too many traces for one launch"); return -1; }
    kern<<<(unsigned)blocks, SWEEP_THREADS, smem, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, err_len, "sweep launch failed: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace dct
