// Debug / validation kernel: re-derives the photo-electron lists and noise values a trace is built
// from.  Because every random number is a pure function of (seed, event, pixel, stream, block), this
// kernel reproduces exactly what k_generate_* consumed without touching those kernels.  It is what
// the reference exposes as nsb_time / nsb_photon / mask after next()
// (reference digicamtoy/tracegenerator/__init__.py:248-257, :345-348), and it closes the parity loop
// in the other direction: GPU pe lists -> reference arithmetic (oracle) -> must equal the GPU ADC.
#pragma once
#include "generate.cuh"

namespace dct {

struct DumpArgs {
    GenArgs g;
    uint32_t max_pe;          // capacity per trace of pe_time / pe_amp (0: count only)
    int32_t* n_pe;            // [n_traces]
    double* pe_time;          // [n_traces, max_pe]  absolute time (ns)
    float* pe_amp;            // [n_traces, max_pe]  smeared cluster amplitude (p.e.)
    float* noise;             // [n_traces, B] electronic noise of the kept bins (LSB), or NULL
};

__global__ void k_dump_pe(const DumpArgs d) {
    const GenArgs& a = d.g;
    const DevParams& p = a.p;
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_traces) return;
    const TraceId t = trace_id(a, g);
    const float rate = p.rate[t.pixel];
    const PixelNsb q = load_pixel_nsb(p, t.pixel);
    uint32_t count = 0;
    auto emit = [&](double time, float amp) {
        if (count < d.max_pe) {
            d.pe_time[g * d.max_pe + count] = time;
            d.pe_amp[g * d.max_pe + count] = amp;
        }
        ++count;
    };
    if (rate > 0.0f) {
        if (p.sub_binning <= 0) {
            float elapsed = 0.0f;
            int cell = p.cell0;
            float off = p.off0;
            for (uint32_t grp_i = 0;; ++grp_i) {
                const NsbGroup grp = draw_nsb_group(a.rk, t.key, grp_i);
                const uint32_t gw[4] = {grp.gap.x, grp.gap.y, grp.gap.z, grp.gap.w};
                const uint32_t bw[4] = {grp.bor.x, grp.bor.y, grp.bor.z, grp.bor.w};
                float z[4];
                normal_pair(grp.nrm.x, grp.nrm.y, z[0], z[1]);
                normal_pair(grp.nrm.z, grp.nrm.w, z[2], z[3]);
                bool done = false;
                for (int k = 0; k < 4 && !done; ++k) {
                    const float gap = exp_gap(gw[k], q.neg_ln2_over_rate);
                    elapsed += gap;
                    if (elapsed >= p.span) { done = true; break; }
                    advance_onset(p, gap, cell, off);
                    const float A = cluster_amplitude((float)borel(bw[k], q.xt, q.borel_cdf), z[k], q.sigma1);
                    // onset coordinate s = cell*dt + off = (time - t0) + x_lo
                    emit((double)cell * p.dt_d + (double)off + p.t0_d - p.x_lo_d, A);
                }
                if (done) break;
            }
        } else {
            const float lam = rate * p.dt, p0 = __expf(-lam);
            const int n_sim = p.n_drop + p.B;
            const bool gp = lam < GP_LAM_MAX && p.gp_cdf != nullptr;
            for (int bs = 0; bs < n_sim; ++bs) {
                float A;
                if (gp) {
                    const u32x4 r = draw_block_rk(a.rk, t.key, STREAM_SUBBIN_GP, (uint32_t)bs >> 1);
                    if (!subbin_total(r, (uint32_t)bs, lam, q, p.gp_cdf + (size_t)t.pixel * GP_TABLE, 1, A)) continue;
                } else if (!subbin_cluster(t.key, (uint32_t)bs, lam, p0, q, A)) {
                    continue;
                }
                emit(p.t0_d + (double)bs * p.dt_d, A);
            }
        }
    }
    d.n_pe[g] = (int32_t)count;
    if (d.noise) {
        const float sigma_e = p.sigma_e[t.pixel];
        for (int qd = 0; 4 * qd < p.B; ++qd) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            if (sigma_e > 0.0f) {
                const u32x4 r = draw_block(t.key, STREAM_NOISE, (uint32_t)qd);
                normal_pair(r.x, r.y, z[0], z[1]);
                normal_pair(r.z, r.w, z[2], z[3]);
            }
            for (int i = 0; i < 4; ++i)
                if (4 * qd + i < p.B) d.noise[g * p.B + 4 * qd + i] = z[i] * sigma_e;
        }
    }
}

}  // namespace dct
