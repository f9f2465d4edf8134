// K1t: high-rate generation kernel with the pulse shaping on the 5th-generation tensor cores.
//
// compute_analog_signal (reference digicamtoy/tracegenerator/__init__.py:350-362) is
//     V[b] = sum_pe A_pe * h(t_b - t_pe),      h = cubic spline with knots every 0.5 ns.
// Within one knot interval h is a cubic, so a pe in sampling cell c (4 ns), knot phase f (0..7),
// local offset u contributes   A * (g0 + g1 v + g2 v^2 + g3 v^3)[tap j][phase f]   to bin c+1+j,
// v = 2u/dx - 1 in [-1, 1].  Summing the four "moments" (A, Av, Av^2, Av^3) of all pe of one
// (cell, phase) first turns the shaping of a cell into a small dense product
//     D[trace, c+1+j] += M_c[trace, (f, k)] * G[(f, k), j]          (128 traces x 32 x 32)
// with ONE constant coefficient tile G (the template is shift-invariant: cell c+1 uses the same
// tile one accumulator column further on).  That product runs on tcgen05.mma (kind::f16, fp32
// accumulators in TMEM); the CUDA cores are left with the random numbers, the samplers, four
// multiplies and one 8-byte shared-memory store per pe -- no per-tap loads or FMAs, no window.
//
// The photo-electron stream is the one of the sweep / scatter kernels, Philox block for Philox
// block (NsbGroup: same gaps, cluster sizes and amplitudes), so dct_trace_pe describes this kernel too;
// only the summation differs: moments are rounded to fp16 (11 significant bits, round to nearest,
// unbiased) before the product, the coefficients are split hi + lo (22 bits), accumulation is
// fp32.  The analog sum differs from the fp32 kernels by ~2e-3 LSB rms at 1 GHz
// (tests/test_tensor_gpu.py), far below the 1.25 LSB of electronic noise; signal pulses (up to
// thousands of pe) are NOT sent through fp16: they are shaped in fp32 by the code of generate.cuh.
//
// CTA = 128 traces = 4 producer warps (thread = trace) + warp 4 (MMA issue) + warp 5 (slot clearing)
// + 2 idle warps (a CTA must hold a multiple of four warps, see TC_THREADS).  Producers free-run
// through time in knot steps, four pe per look at the ring (one NsbGroup = three Philox blocks); the
// moments of cell c go to slot c mod R of a shared-memory ring (A operand, K-major, no swizzle).  A warp
// publishes the cell all its lanes are past (mailbox + a flip of the notify mbarrier); once all four
// warps are past cell c the MMA warp issues its 4 MMAs (2 k-steps x hi/lo coefficients) and commits
// them to the slot's mbarrier; the cleaner warp waits for that, zeroes the slot and releases it.  A
// lane that is R cells ahead of the slowest trace of the CTA idles until a slot is free.  Four CTAs per
// SM hold all 512 TMEM columns: TMEM caps the occupancy.
#pragma once
#include <cuda_fp16.h>

#include "generate.cuh"

namespace dct {

constexpr int TC_TRACES = 128;                 // MMA M
constexpr int TC_THREADS = 256;                // 4 producer warps + 4 helper warps (MMA issue, slot clearing, 2 idle).
                                               // A multiple of 4 warps on purpose: a warp may only touch the TMEM
                                               // lanes 32 (%warpid % 4) .., and with 5 warps per CTA the second CTA
                                               // of an SM faulted on its first tcgen05.st
constexpr int TC_N = 32;                       // accumulator columns per cell MMA (22 taps + padding)
constexpr int TC_TAPS = 22;
constexpr int TC_PHASES = 8;
constexpr int TC_K = TC_PHASES * 4;            // fp16 K per cell: 8 phases x 4 moments
constexpr int TC_SLOT_BYTES = TC_TRACES * TC_K * 2;      // 8 KiB
constexpr int TC_G_BYTES = TC_N * TC_K * 2;              // 2 KiB per coefficient tile
constexpr int TC_TMEM_COLS = 128;
constexpr int TC_MAX_BINS = 100;
// Accumulator column of kept bin 0.  A cell's MMA writes TC_N (32) columns from its own column on, so
// windows of up to 80 bins keep 16 spare columns in front and 32 behind the kept bins.  Longer ones
// (dark.yml: 92 bins) start at column 10 and let the last few cells -- whose taps 16.. land beyond the
// window anyway -- run with N = 16 so that nothing is written past column 127 (tensor_layout).
__host__ __device__ inline int tensor_col0(int B) { return B <= 80 ? 16 : 10; }

struct TensorLayout {
    int col0, col_shift;        // column of kept bin 0; column of tap 0 of cell c (relative to cell0) = c + col_shift
    int n_cells32;              // cells 0 .. n_cells32-1 use N = 32, the rest N = 16
    bool ok;
};

// Column plan for `n_cells` cells in an accumulator of `cols` columns; ok = every tap that lands on a
// kept bin is covered.
inline TensorLayout tensor_layout(int B, int n_drop, int cell0, int n_cells, int cols = TC_TMEM_COLS) {
    TensorLayout L;
    L.col0 = tensor_col0(B);
    L.col_shift = cell0 + 1 - n_drop + L.col0;
    L.n_cells32 = n_cells;
    L.ok = L.col_shift >= 0;
    const int last_col = L.col0 + B - 1;                       // column of the last kept bin
    for (int c = 0; c < n_cells; ++c) {
        const int col = c + L.col_shift, d = col & ~1;         // odd columns accumulate from the even one below
        if (d + TC_N <= cols) continue;                        // N = 32 fits
        if (c < L.n_cells32) L.n_cells32 = c;
        const int taps_needed = last_col - col + 1;            // taps 0 .. taps_needed-1 land on kept bins
        if (d + 16 > cols || (col - d) + taps_needed > 16) L.ok = false;
    }
    return L;
}
#ifndef DCT_TC_RING
#define DCT_TC_RING 5                          // 40 KiB of moment slots: four CTAs (all 512 TMEM columns) per SM
#endif
#ifndef DCT_TC_MIN_BLOCKS
#define DCT_TC_MIN_BLOCKS 4
#endif
#ifndef DCT_TC_POLL_NS
#define DCT_TC_POLL_NS 1000                    // suspend-time hint of the MMA warp's wait on the notify barrier (it re-reads the mailboxes after)
#endif
#ifndef DCT_TC_GROUPS
#define DCT_TC_GROUPS 1                        // NsbGroups (four pe each) a lane draws per look at the ring state
#endif
#ifndef DCT_TC_REFRESH_LIMIT
#define DCT_TC_REFRESH_LIMIT 0                 // 1: look at the released-cell counter before every store, not once per group
#endif
// mean pe per trace from which AUTO prefers this kernel over the sweep kernel: measured crossover at
// ~65 pe / trace (0.25 GHz on the 240 ns window = 60 pe: sweep 10.5 against 10.9 ms per 8192 events,
// 0.4 GHz = 96 pe: 14.6 against 12.1; profiles/r02/crossover_tensor.txt) -- below it the per-cell
// hand-over (59 cells whatever the rate) costs more than the taps it saves.
#ifndef DCT_TC_MIN_MEAN_PE
#define DCT_TC_MIN_MEAN_PE 66.0
#endif

struct TensorArgs {
    GenArgs g;
    // coefficient tiles [TC_N][TC_K] in the canonical K-major no-swizzle layout (tc_tile_offset):
    // hi, lo, then the same two with every tap one row (= accumulator column) further down.
    // tcgen05.mma faults ("misaligned address") on an odd accumulator column, so odd cells
    // accumulate from the even column below through the shifted tiles.
    const __half* coef;
    int n_cells;                 // cells cell0 .. cell0 + n_cells - 1 can reach a kept bin
    int col_shift;               // accumulator column of tap 0 of cell c (relative to cell0): c + col_shift
    int col0;                    // accumulator column of kept bin 0
    int n_cells32;               // cells below it issue N = 32 MMAs, the others N = 16 (end of a long window)
    float v_scale;               // 2 / knot_dx
    int check_span;              // 0: the cell test alone ends the pe stream (see k_generate_tensor)
    float y0, y_end;             // producer clock in knot steps from the start of cell0: first value, end of the stream
};

// Byte offset of element (row, k) in a canonical K-major, no-swizzle UMMA operand tile with
// K = 32 halves: 8-row x 16-byte core matrices, 128 B apart along K (LBO), 512 B apart along
// rows (SBO).
__host__ __device__ inline int tc_tile_offset(int row, int k) {
    return (row >> 3) * 512 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

inline bool tensor_supported(const DevParams& p) {
    return p.n_phase == TC_PHASES && p.n_taps == TC_TAPS && p.sub_binning <= 0 && p.B <= TC_MAX_BINS &&
           p.cell0 >= 0 && p.n_drop - 1 - p.cell0 <= tensor_col0(p.B) && (p.B & 1) == 0;
}

__host__ __device__ inline size_t tensor_smem_bytes(int B, int ring) {
    size_t ringb = (size_t)ring * TC_SLOT_BYTES;
    size_t rows = (size_t)TC_TRACES * row_stride(B) * sizeof(float);
#ifndef DCT_TC_PAD_SMEM
#define DCT_TC_PAD_SMEM 0      // experiment: extra bytes to limit the CTAs per SM
#endif
    return (ringb > rows ? ringb : rows) + 4 * TC_G_BYTES + 64 + 8 * (size_t)ring + DCT_TC_PAD_SMEM;
}

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 in
// [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    const uint64_t lo = (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16);
    const uint64_t hi = (uint64_t)(512u >> 4) | (1ull << 14);
    return lo | (hi << 32);
}

// instruction descriptor: D = F32, A = B = F16, both K-major, N = 32 (or 16), M = 128
__host__ __device__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TRACES >> 4) << 24); }

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate,
                                        uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// bounded wait: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(32);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
// one look at a phase with a hardware-suspended wait of up to `ns` nanoseconds; true = the phase is over
__device__ __forceinline__ bool mbar_try_wait_ns(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}

// The cleaner warp's wait for a phase.  DCT_TC_CLEAN_SLEEP_NS > 0: look, sleep, look (few instructions
// while waiting); 0: mbarrier.try_wait back to back (the hardware suspends each look for its own time limit).
#ifndef DCT_TC_CLEAN_SLEEP_NS
#define DCT_TC_CLEAN_SLEEP_NS 100            // 0 .. 1000 ns: no measurable difference (profiles/r02/tensor_mega.txt)
#endif
__device__ __forceinline__ void mbar_wait_lazy(uint32_t bar, uint32_t parity) {
    uint32_t looks = 0, done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (DCT_TC_CLEAN_SLEEP_NS) __nanosleep(DCT_TC_CLEAN_SLEEP_NS);
        if (++looks > (1u << 26)) __trap();          // bounded: a protocol error fails the launch
    }
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
}

// 16 consecutive accumulator columns of this warp's 32 lanes -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// zero 32 consecutive accumulator columns of this warp's 32 lanes
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        :: "r"(taddr), "r"(z) : "memory");
}

}  // namespace tc

// One group of 128 traces: its ring of moment slots, its mailboxes and its accumulator.  The CTA
// kernel has one group per CTA, the persistent kernel several per CTA.
//   ctl[0..3]: cell every lane of producer warp w has reached; ctl[4]: cells released; ctl[5]: TMEM
//   base (CTA kernel); ctl[10..11]: mbarrier a producer warp flips whenever it publishes a new
//   position (wakes the MMA warp); ctl[16 ..]: one mbarrier per ring slot that the MMAs of its cell
//   commit to.
struct TcGroup {
    unsigned char* ring;        // RING slots; the float rows of the epilogue afterwards
    uint32_t* ctl;
    uint32_t tmem;              // accumulator: lane 0, first column of the group
    uint32_t g_addr;            // coefficient tiles hi, lo, hi', lo' (shared address)
};
constexpr int TC_CTL_BYTES = 128;               // per group in the persistent kernel (64 + 8 RING, RING <= 8)

template <int RING>
__device__ __forceinline__ void tc_init_barriers(const TcGroup& G) {
    const uint32_t notify = tc::smem_u32(G.ctl + 10), done = tc::smem_u32(G.ctl + 16);
    tc::mbar_init(notify, 1);          // every arrival completes a phase
    for (int i = 0; i < RING; ++i) tc::mbar_init(done + 8u * (uint32_t)i, 1);
}

// MMA warp: issues the MMAs of a cell as soon as every producer warp is past it (it sleeps on the
// notify barrier in between).
template <int RING>
__device__ __forceinline__ void tc_issue_cells(const TensorArgs& ta, const TcGroup& G, int lane, uint32_t& notify_parity) {
    const uint32_t ring_addr = tc::smem_u32(G.ring);
    const uint32_t notify = tc::smem_u32(G.ctl + 10), done = tc::smem_u32(G.ctl + 16);
    const uint32_t* warp_pos = G.ctl;
    const int n_cells = ta.n_cells;
    int slot_i = 0;
    for (int c = 0; c < n_cells; ++c) {
        const uint32_t slot = ring_addr + (uint32_t)slot_i * TC_SLOT_BYTES;
        // A stale parity only costs one more look, a missed flip at most DCT_TC_POLL_NS.
        uint32_t spins = 0;
        while (true) {
            const uint32_t m = min(min(tc::ld_volatile(warp_pos), tc::ld_volatile(warp_pos + 1)),
                                   min(tc::ld_volatile(warp_pos + 2), tc::ld_volatile(warp_pos + 3)));
            if (m > (uint32_t)c) break;
            if (tc::mbar_try_wait_ns(notify, notify_parity, DCT_TC_POLL_NS)) notify_parity ^= 1u;
            if (++spins > (1u << 24)) __trap();
        }
        __threadfence_block();
        tc::fence_after();
        if (lane == 0) {
            const uint32_t col = (uint32_t)(c + ta.col_shift);
            const uint32_t d = G.tmem + (col & ~1u);
            const uint32_t g = G.g_addr + (col & 1u) * (2 * TC_G_BYTES);      // odd column: shifted tiles
            const uint64_t a0 = tc::smem_desc(slot), a1 = tc::smem_desc(slot + 256);
            const uint32_t idesc = c < ta.n_cells32 ? tc::idesc_n(TC_N) : tc::idesc_n(16);
            tc::mma_f16(d, a0, tc::smem_desc(g), 1u, idesc);
            tc::mma_f16(d, a1, tc::smem_desc(g + 256), 1u, idesc);
            tc::mma_f16(d, a0, tc::smem_desc(g + TC_G_BYTES), 1u, idesc);
            tc::mma_f16(d, a1, tc::smem_desc(g + TC_G_BYTES + 256), 1u, idesc);
            tc::mma_commit(done + 8u * (uint32_t)slot_i);
        }
        __syncwarp();
        if (++slot_i == RING) slot_i = 0;
    }
}

// Cleaner warp: waits for the MMAs of each cell (tcgen05.commit -> done[slot]), clears the slot and
// releases it.  It runs independently of the MMA warp, so the hand-over of cell c+1 does not queue
// behind the clearing of cell c.  `phases`: one parity bit per slot, carried from launch to end.
template <int RING>
__device__ __forceinline__ void tc_clean_cells(const TensorArgs& ta, const TcGroup& G, int lane, uint32_t& phases) {
    const uint32_t done = tc::smem_u32(G.ctl + 16);
    uint32_t* released = G.ctl + 4;
    const int n_cells = ta.n_cells;
    int slot_i = 0;
    for (int c = 0; c < n_cells; ++c) {
        tc::mbar_wait_lazy(done + 8u * (uint32_t)slot_i, (phases >> slot_i) & 1u);      // the MMAs of cell c have read the slot
        phases ^= 1u << slot_i;
        if (c + RING < n_cells) {
            // the slot will be used again: clear it (generic writes after the async reads finished)
            int4* s4 = reinterpret_cast<int4*>(G.ring + (size_t)slot_i * TC_SLOT_BYTES);
#pragma unroll
            for (int i = 0; i < TC_SLOT_BYTES / 16 / 32; ++i) s4[i * 32 + lane] = make_int4(0, 0, 0, 0);
            tc::fence_async_proxy();
            __syncwarp();
        }
        if (lane == 0) tc::st_volatile(released, (uint32_t)(c + 1));
        if (++slot_i == RING) slot_i = 0;
    }
}

// Producer warps: thread = trace (gtid = 0 .. 127 within the group, gwarp = gtid / 32).
// Time runs in knot steps (0.5 ns) from the start of cell0:  y += gap / dx.  Its integer part
// is the knot interval (8 per cell), gi = 8 cell + j, and with w = frac(y) the pe sits
// x0 = (8 - j - w) dx before the next bin: phase f = 7 - j, local offset u = (1 - w) dx,
// v = 2u/dx - 1 = 1 - 2w.  (The sweep kernel keeps (cell, offset in ns); the two placements
// agree to ~1e-4 ns, a 1e-4 LSB effect.  The moments are stored by j; the coefficient tile
// is permuted to match.)
//
// CHECK_SPAN: the pe stream ends where the elapsed time passes the NSB window (as in the other
// kernels).  When the last cell that can reach a kept bin ends a full cell before the window does
// (C4: cells 0..58 of 60), the cell test alone ends the stream at the same pe or earlier, with
// nothing lost, and the elapsed time need not be tracked.
template <int RING, bool CHECK_SPAN>
__device__ __forceinline__ void tc_produce(const TensorArgs& ta, const TcGroup& G, uint64_t trace0, int gtid, int gwarp,
                                           int lane) {
    const GenArgs& a = ta.g;
    const DevParams& p = a.p;
    const int n_cells = ta.n_cells;
    uint32_t* warp_pos = G.ctl;
    uint32_t* released = G.ctl + 4;
    const uint32_t notify = tc::smem_u32(G.ctl + 10);

    const bool live = trace0 + gtid < a.n_traces;
    const TraceId t = trace_id(a, live ? trace0 + gtid : 0);
    const float rate = live ? p.rate[t.pixel] : 0.0f;
    const PixelNsb q = load_pixel_nsb(p, t.pixel);
    const float gap_scale = q.neg_ln2_over_rate * p.inv_knot_dx;
    // row offset of this trace inside a slot (K-major core matrices, see tc_tile_offset)
    uint32_t row_off = tc::smem_u32(G.ring) + (uint32_t)((gtid >> 3) * 512 + (gtid & 7) * 16);
    asm volatile("" : "+r"(row_off));                // keep it in a register (ptxas rebuilt it per pe)
    auto nsb_block = [&](uint32_t block) { return draw_block_rk(a.rk, t.key, STREAM_NSB, block); };
    const uint32_t n8 = (uint32_t)n_cells * TC_PHASES;

    // Lane state: a queue of the pe of one (or two) NsbGroups (knot interval gq = 8 cell + j, amplitude,
    // offset), nq = the next one to store (NQ: queue empty, draw the next group).  A pe past the end
    // of the stream carries gq >= n8 and is never stored (the write limit never exceeds n8), so a
    // finished lane simply keeps its queue.
    float y = ta.y0;
    constexpr int NQ = 4 * DCT_TC_GROUPS;            // pe per look at the ring: one or two NsbGroups
    uint32_t gq[NQ];
    float Aq[NQ], vq[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) { gq[k] = n8; Aq[k] = 0.f; vq[k] = 0.f; }
    int nq = (rate > 0.0f) ? NQ : 0;
    uint32_t cur8 = (rate > 0.0f) ? 0u : n8;         // 8 x the cell this lane has reached
    uint32_t prev = 0xffffffffu;                     // interval of the last pe of the group before
    uint32_t warp_done = 0;
    uint32_t limit8 = min((uint32_t)RING * TC_PHASES, n8);   // intervals below it may be written: 8 (released + RING)
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    uint32_t grp = 0;
    uint32_t idle = 0;
    const float cdf1 = q.borel_cdf.x, cdf2 = q.borel_cdf.y;

    // one pe of the group: clock, interval, offset, amplitude
    auto make_pe = [&](uint32_t w_gap, uint32_t w_bor, float z, uint32_t& gq, float& Aq, float& vq) {
        y = fminf(__fmaf_rn(fast_lg2(u01_open0(w_gap)), gap_scale, y), 4194304.0f);
        const float yi = __fadd_rz(y, 8388608.0f);              // 2^23 + floor(y)
        uint32_t g = __float_as_uint(yi) & 0x007fffffu;
        if (CHECK_SPAN) g = (y < ta.y_end) ? g : 0x00800000u;    // past the NSB window
        gq = g;
        vq = __fmaf_rn(y - (yi - 8388608.0f), -2.0f, 1.0f);
        // Borel cluster size: the first threshold inline, the rest (0.8 % of the draws at
        // xt = 0.08) through the generic sampler -- same value as borel()
        const float u = u01_open1(w_bor);
        float fn = u >= cdf1 ? 2.0f : 1.0f;
        if (u >= cdf2) fn = (float)borel(w_bor, q.xt, p.borel_cdf[t.pixel]);
        Aq = cluster_amplitude(fn, z, q.sigma1);
    };
    // store pe k of the queue if it is this lane's next one and its slot is free
    auto store_pe = [&](int k, uint32_t gq, uint32_t before, float A, float v, uint32_t next) {
#if DCT_TC_REFRESH_LIMIT
        if (k > 0) limit8 = min((tc::ld_volatile(released) + (uint32_t)RING) * TC_PHASES, n8);
#endif
        if (nq == k && gq < limit8) {
            const float keep = gq == before ? 1.0f : 0.0f;   // same (cell, interval) as the pe before: sums carry on
            const float m1 = A * v, m2 = m1 * v, m3 = m2 * v;
            acc0 = __fmaf_rn(acc0, keep, A);
            acc1 = __fmaf_rn(acc1, keep, m1);
            acc2 = __fmaf_rn(acc2, keep, m2);
            acc3 = __fmaf_rn(acc3, keep, m3);
            const __half2 h01 = __floats2half2_rn(acc0, acc1);
            const __half2 h23 = __floats2half2_rn(acc2, acc3);
            // slot of the cell + place of interval j in the row: (j >> 1) 128 + (j & 1) 8 bytes,
            // from a byte table (PRMT) -- {0, 1, 16, 17, 32, 33, 48, 49} x 8
            const uint32_t joff8 = __byte_perm(0x11100100u, 0x31302120u, gq & 7u);
            const uint32_t addr = row_off + ((gq >> 3) % (uint32_t)RING) * TC_SLOT_BYTES + (joff8 << 3);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};"
                         :: "r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                            "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
            nq = k + 1;
            cur8 = next;
        }
    };

    while (true) {
        if (nq == NQ) {
            // queue empty: the next pe, four per three Philox blocks (gaps, cluster sizes, two Box-Muller pairs)
            prev = gq[NQ - 1];
#pragma unroll
            for (int gi = 0; gi < DCT_TC_GROUPS; ++gi) {
                const u32x4 rg = nsb_block(3u * grp);
                const u32x4 rb = nsb_block(3u * grp + 1u);
                const u32x4 rn = nsb_block(3u * grp + 2u);
                ++grp;
                float z0, z1, z2, z3;
                normal_pair(rn.x, rn.y, z0, z1);
                normal_pair(rn.z, rn.w, z2, z3);
                make_pe(rg.x, rb.x, z0, gq[4 * gi + 0], Aq[4 * gi + 0], vq[4 * gi + 0]);
                make_pe(rg.y, rb.y, z1, gq[4 * gi + 1], Aq[4 * gi + 1], vq[4 * gi + 1]);
                make_pe(rg.z, rb.z, z2, gq[4 * gi + 2], Aq[4 * gi + 2], vq[4 * gi + 2]);
                make_pe(rg.w, rb.w, z3, gq[4 * gi + 3], Aq[4 * gi + 3], vq[4 * gi + 3]);
            }
            nq = 0;
            cur8 = gq[0];
        }
#pragma unroll
        for (int k = 0; k < NQ; ++k)
            store_pe(k, gq[k], k ? gq[k - 1] : prev, Aq[k], vq[k], gq[k + 1 < NQ ? k + 1 : NQ - 1]);
        // publish the cell every lane of this warp has reached
        const uint32_t wmin = __reduce_min_sync(0xffffffffu, cur8) >> 3;
        if (wmin > warp_done) {
            tc::fence_async_proxy();
            __syncwarp();
            if (lane == 0) {
                tc::st_volatile(warp_pos + gwarp, min(wmin, (uint32_t)n_cells));
                tc::mbar_arrive(notify);
            }
            warp_done = wmin;
        }
        if (wmin >= (uint32_t)n_cells) break;
        limit8 = min((tc::ld_volatile(released) + (uint32_t)RING) * TC_PHASES, n8);
        if (!__any_sync(0xffffffffu, nq == NQ || cur8 < limit8)) {      // every lane waits for a slot
            __nanosleep(32);
            if (++idle > (1u << 26)) __trap();       // a protocol error fails the launch instead of hanging
        }
    }
    // every MMA of this group has completed once the last cell is released
    {
        uint32_t spins = 0;
        while (tc::ld_volatile(released) < (uint32_t)n_cells) {
            __nanosleep(32);
            if (++spins > (1u << 26)) __trap();
        }
    }
    __threadfence_block();
    tc::fence_after();
}

__device__ __forceinline__ void tc_bar_sync(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n_threads) : "memory");
}

// Epilogue of the producer warps, once the ring is dead for all four of them: accumulator row +
// signal pulses (fp32, as in the other kernels) -> noise, digitisation -> coalesced store.
// `bar`: a named barrier of the group's 128 producer threads.
template <bool FUSED>
__device__ __forceinline__ void tc_epilogue(const TensorArgs& ta, const TcGroup& G, uint64_t trace0, int gtid, int gwarp,
                                            uint32_t bar, unsigned char* stats_smem) {
    const GenArgs& a = ta.g;
    const DevParams& p = a.p;
    const int B = p.B;
    float* rows = reinterpret_cast<float*>(G.ring);
    float* row = rows + gtid * row_stride(B);
    const bool live = trace0 + gtid < a.n_traces;
    for (int b = 0; b < B; ++b) row[b] = 0.0f;
    if (live) {
        const TraceId t = trace_id(a, trace0 + gtid);
        add_signal_pulses(a, t, p.coef32, row, p.xt[t.pixel], p.sigma1[t.pixel]);
    }
    // NSB sums: accumulator columns col0 .. col0 + B - 1 of this thread's TMEM lane
    const uint32_t tw = G.tmem + ((uint32_t)(gwarp * 32) << 16) + (uint32_t)ta.col0;
    for (int b0 = 0; b0 < B; b0 += 16) {
        float d[16];
        tc::tmem_ld16(tw + b0, d);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (b0 + j < B) row[b0 + j] += d[j];
    }
    tc::fence_before();
    if (live) {
        const TraceId t = trace_id(a, trace0 + gtid);
        dump_analog(a, t.g, row);
        digitise_trace(p, a.rk, t, row);
    }
    tc_bar_sync(bar, TC_TRACES);
    if (FUSED && DCT_FUSED_STATS && a.st.adc_hist)
        cta_fused_stats<TC_TRACES, 1>(a.st, p.B, p.P, p.adc_min, p.adc_max, p.baseline, trace0, a.n_traces, rows, stats_smem);
    store_run<TC_TRACES>(a, trace0, rows, gtid);
}

// ---------------------------------------------------------------------------------------------------
// CTA kernel: one group per CTA, four CTAs per SM (all 512 TMEM columns)
template <int RING, bool CHECK_SPAN>
__global__ void __launch_bounds__(TC_THREADS, DCT_TC_MIN_BLOCKS)
k_generate_tensor(const __grid_constant__ TensorArgs ta) {
    static_assert(RING >= 2, "ring needs two slots");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GenArgs& a = ta.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.p.B;

    // ---- shared memory carve-up
    const size_t ring_bytes = (size_t)RING * TC_SLOT_BYTES;
    const size_t rows_bytes = (size_t)TC_TRACES * row_stride(B) * sizeof(float);
    unsigned char* g_tiles = smem_raw + (ring_bytes > rows_bytes ? ring_bytes : rows_bytes);   // hi, lo, hi', lo'
    TcGroup G;
    G.ring = smem_raw;
    G.ctl = reinterpret_cast<uint32_t*>(g_tiles + 4 * TC_G_BYTES);
    G.g_addr = tc::smem_u32(g_tiles);
    uint32_t* tmem_slot = G.ctl + 5;

    // ---- one-time setup
    {
        const int4* src = reinterpret_cast<const int4*>(ta.coef);
        for (int i = tid; i < 4 * TC_G_BYTES / 16; i += TC_THREADS) reinterpret_cast<int4*>(g_tiles)[i] = src[i];
        int4* r4 = reinterpret_cast<int4*>(G.ring);
        for (int i = tid; i < (int)(ring_bytes / 16); i += TC_THREADS) r4[i] = make_int4(0, 0, 0, 0);
        if (tid < 5) G.ctl[tid] = 0;          // not ctl[5]: tcgen05.alloc (warp 4) writes it concurrently
        if (tid == 0) {
            tc_init_barriers<RING>(G);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc::smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_async_proxy();           // zeroed ring + coefficient tiles -> visible to the tensor core
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tc::ld_volatile(tmem_slot);
    G.tmem = tmem;

    if (warp < 4) {
        // clear this warp's 32 accumulator lanes
        const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < TC_TMEM_COLS; c += 32) tc::tmem_zero32(tw + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();

    const uint64_t trace0 = (uint64_t)blockIdx.x * TC_TRACES;

    // Warps 6-7 only exist so that the CTA has a multiple of four warps (TMEM lane rule).
    if (warp == 4) {
        uint32_t notify_parity = 0;
        tc_issue_cells<RING>(ta, G, lane, notify_parity);
    } else if (warp == 5) {
        uint32_t phases = 0;
        tc_clean_cells<RING>(ta, G, lane, phases);
    } else if (warp < 4) {
        tc_produce<RING, CHECK_SPAN>(ta, G, trace0, tid, warp, lane);
    }
    // the ring is dead from here on: it becomes the float rows of the epilogue
    __syncthreads();

    if (warp < 4)
        tc_epilogue<true>(ta, G, trace0, tid, warp, 1u, smem_raw + stats_offset(tensor_smem_bytes(B, RING)));
    tc::fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Persistent kernel: ONE CTA per SM holding NG independent groups of 128 traces, each with its own
// four producer warps, MMA warp, cleaner warp, ring and accumulator; a group walks through its share
// of the 128-trace items by itself (named barriers per group, never a CTA-wide one between set-up
// and exit), so the groups drift apart and one group's epilogue overlaps the others' generation as
// separate CTAs would.  What the single CTA buys: one copy of the coefficient tiles (8 KiB) instead
// of one per CTA and the whole tensor memory in one allocation that can be cut into NG = 5 slices of
// 102 columns (a 50-bin window needs 98) where tcgen05.alloc only hands out powers of two -- five
// groups with five slots each per SM instead of four.
// Warps 0 .. 4 NG - 1: producers (group = warp / 4, so warp % 4 is the TMEM lane quarter);
// warps 4 NG + 2 g, + 1: MMA warp and cleaner warp of group g; the rest pad the CTA to 4 k warps.
__host__ __device__ constexpr int tensor_mega_threads(int ng) { return ((ng * 6 + 3) / 4) * 128; }
__host__ __device__ inline size_t tensor_mega_group_bytes(int B, int ring) {
    size_t ringb = (size_t)ring * TC_SLOT_BYTES;
    size_t rows = (size_t)TC_TRACES * row_stride(B) * sizeof(float);
    return ((ringb > rows ? ringb : rows) + 127) & ~(size_t)127;
}
__host__ __device__ inline size_t tensor_mega_smem_bytes(int B, int ring, int ng) {
    return (size_t)ng * tensor_mega_group_bytes(B, ring) + 4 * TC_G_BYTES + (size_t)ng * TC_CTL_BYTES + 16;
}

template <int RING, bool CHECK_SPAN, int NG>
__global__ void __launch_bounds__(tensor_mega_threads(NG), 1)
k_generate_tensor_mega(const __grid_constant__ TensorArgs ta, const int cols_per_group) {
    static_assert(RING >= 2 && RING <= 8, "ring: 2 .. 8 slots");
    static_assert(2 * NG + 1 <= 15, "two named barriers per group");
    constexpr int NT = tensor_mega_threads(NG);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GenArgs& a = ta.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.p.B;

    const size_t group_bytes = tensor_mega_group_bytes(B, RING);
    unsigned char* g_tiles = smem_raw + (size_t)NG * group_bytes;
    unsigned char* ctl0 = g_tiles + 4 * TC_G_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl0 + (size_t)NG * TC_CTL_BYTES);

    // role of this warp
    const bool producer = warp < 4 * NG;
    const int helper = warp - 4 * NG;                          // 2 g: MMA warp, 2 g + 1: cleaner
    const int g = producer ? (warp >> 2) : (helper >> 1);
    const bool member = g < NG;                                // padding warps belong to no group
    const int gtid = tid - g * TC_TRACES, gwarp = warp & 3;    // producers only
    TcGroup G;
    G.ring = smem_raw + (size_t)(member ? g : 0) * group_bytes;
    G.ctl = reinterpret_cast<uint32_t*>(ctl0 + (size_t)(member ? g : 0) * TC_CTL_BYTES);
    G.g_addr = tc::smem_u32(g_tiles);

    // ---- one-time setup
    {
        const int4* src = reinterpret_cast<const int4*>(ta.coef);
        for (int i = tid; i < 4 * TC_G_BYTES / 16; i += NT) reinterpret_cast<int4*>(g_tiles)[i] = src[i];
        int4* r4 = reinterpret_cast<int4*>(smem_raw);
        for (int i = tid; i < (int)((size_t)NG * group_bytes / 16); i += NT) r4[i] = make_int4(0, 0, 0, 0);
        if (tid < NG * 5) reinterpret_cast<uint32_t*>(ctl0 + (size_t)(tid / 5) * TC_CTL_BYTES)[tid % 5] = 0;
        if (tid < NG) {
            TcGroup Gi = G;
            Gi.ctl = reinterpret_cast<uint32_t*>(ctl0 + (size_t)tid * TC_CTL_BYTES);
            tc_init_barriers<RING>(Gi);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (warp == 4 * NG) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc::smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_async_proxy();           // zeroed rings + coefficient tiles -> visible to the tensor core
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tc::ld_volatile(tmem_slot);
    G.tmem = tmem + (uint32_t)((member ? g : 0) * cols_per_group);

    auto zero_accumulator = [&]() {       // this producer warp's 32 lanes of the group's columns
        const uint32_t tw = G.tmem + ((uint32_t)(gwarp * 32) << 16);
        for (int c = 0; c < cols_per_group; c += 32) tc::tmem_zero32(tw + min(c, cols_per_group - 32));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    if (producer) zero_accumulator();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();

    if (member) {
        const uint32_t bar_prod = 1u + (uint32_t)g, bar_group = 1u + (uint32_t)NG + (uint32_t)g;
        const uint32_t n_items = (uint32_t)((a.n_traces + TC_TRACES - 1) / TC_TRACES);
        uint32_t notify_parity = 0, phases = 0;
        for (uint32_t item = blockIdx.x * (uint32_t)NG + (uint32_t)g; item < n_items; item += gridDim.x * (uint32_t)NG) {
            const uint64_t trace0 = (uint64_t)item * TC_TRACES;
            if (producer) {
                tc_produce<RING, CHECK_SPAN>(ta, G, trace0, gtid, gwarp, lane);
                tc_bar_sync(bar_prod, TC_TRACES);            // the ring is dead: it becomes the float rows
                tc_epilogue<false>(ta, G, trace0, gtid, gwarp, bar_prod, nullptr);
                tc_bar_sync(bar_prod, TC_TRACES);            // rows stored
                // back to the state after set-up: ring, mailboxes and accumulator cleared
                int4* r4 = reinterpret_cast<int4*>(G.ring);
                for (int i = gtid; i < RING * TC_SLOT_BYTES / 16; i += TC_TRACES) r4[i] = make_int4(0, 0, 0, 0);
                if (gtid < 5) G.ctl[gtid] = 0;
                zero_accumulator();
                tc::fence_async_proxy();
                tc::fence_before();
            } else if ((helper & 1) == 0) {
                tc_issue_cells<RING>(ta, G, lane, notify_parity);
            } else {
                tc_clean_cells<RING>(ta, G, lane, phases);
            }
            tc_bar_sync(bar_group, TC_TRACES + 64);
            tc::fence_after();
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 4 * NG) {
        tc::fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
    }
}

#ifndef DCT_TC_MEGA
#define DCT_TC_MEGA 1          // default of the persistent kernel (run time: DCT_TENSOR_MEGA=0/1 at dct_create)
#endif
constexpr int TC_MEGA_COLS5 = 102;             // accumulator columns per group with five groups per SM

template <typename K>
inline int tc_configure(K kern, size_t smem, size_t& configured, char* err, size_t err_len) {
    if (configured >= smem) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { snprintf(err, err_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    configured = smem;
    return 0;
}

inline int launch_tensor(int device, const TensorArgs& ta, cudaStream_t st, char* err, size_t err_len) {
    auto kern = ta.check_span ? k_generate_tensor<DCT_TC_RING, true> : k_generate_tensor<DCT_TC_RING, false>;
    const size_t smem = smem_with_stats(tensor_smem_bytes(ta.g.p.B, DCT_TC_RING), ta.g, TC_TRACES);
    static thread_local size_t configured[16][2] = {{0}};
    if (int rc = tc_configure(kern, smem, configured[device & 15][ta.check_span], err, err_len)) return rc;
    const uint64_t blocks = (ta.g.n_traces + TC_TRACES - 1) / TC_TRACES;
    if (blocks > 0x7fffffffull) { snprintf(err, err_len, "too many traces for one launch"); return -1; }
    kern<<<(unsigned)blocks, TC_THREADS, smem, st>>>(ta);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, err_len, "tensor launch failed: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

// groups: 5 (102 columns each, `ta.n_cells32` planned for that width) or 4 (128 columns each)
inline int launch_tensor_mega(int device, int sm_count, int groups, const TensorArgs& ta, cudaStream_t st, char* err,
                              size_t err_len) {
    const int B = ta.g.p.B;
    const uint64_t items = (ta.g.n_traces + TC_TRACES - 1) / TC_TRACES;
    if (items > 0x7fffffffull) { snprintf(err, err_len, "too many traces for one launch"); return -1; }
    static thread_local size_t configured[16][2][2] = {{{0}}};
    size_t& conf = configured[device & 15][groups == 5][ta.check_span];
    const unsigned ctas = (unsigned)std::min<uint64_t>((uint64_t)sm_count, (items + groups - 1) / groups);
    if (groups == 5) {
        constexpr int R = 5;
        auto kern = ta.check_span ? k_generate_tensor_mega<R, true, 5> : k_generate_tensor_mega<R, false, 5>;
        const size_t smem = tensor_mega_smem_bytes(B, R, 5);
        if (int rc = tc_configure(kern, smem, conf, err, err_len)) return rc;
        kern<<<ctas, tensor_mega_threads(5), smem, st>>>(ta, TC_MEGA_COLS5);
    } else {
        constexpr int R = 6;
        auto kern = ta.check_span ? k_generate_tensor_mega<R, true, 4> : k_generate_tensor_mega<R, false, 4>;
        const size_t smem = tensor_mega_smem_bytes(B, R, 4);
        if (int rc = tc_configure(kern, smem, conf, err, err_len)) return rc;
        kern<<<ctas, tensor_mega_threads(4), smem, st>>>(ta, TC_TMEM_COLS);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, err_len, "tensor launch failed: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace dct
