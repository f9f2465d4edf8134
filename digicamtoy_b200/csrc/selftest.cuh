// Test hook: draws n samples from one of the device samplers so the host tests can compare each
// distribution with its closed form.  Not used by the product path.
#pragma once
#include "sampling.cuh"

namespace dct {

enum : uint32_t { SAMPLE_NORMAL = 0, SAMPLE_EXP_GAP = 1, SAMPLE_BOREL = 2, SAMPLE_POISSON = 3,
                  SAMPLE_BRANCHING = 4, SAMPLE_UNIFORM = 5,
                  // the inversion samplers fed with the LARGEST uniforms, 1 - (i+1) 2^-23: the float sum
                  // of the masses may end below them and the search must still stop in the tail
                  SAMPLE_POISSON_TOP = 6, SAMPLE_BOREL_TOP = 7, SAMPLE_POISSON_INV_TOP = 8,
                  SAMPLE_POISSON_TABLE_TOP = 9,
                  SAMPLE_NORMAL_RADIUS_TOP = 10 };   // Box-Muller radius for the SMALLEST uniforms (r0 = i)

__global__ void k_test_sample(uint32_t kind, float a, float b, uint32_t k0, uint32_t k1, uint32_t n,
                              float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StreamKey key;
    key.k0 = k0; key.k1 = k1; key.ev_lo = i; key.ev_hi = 0; key.pixel = 7;
    const u32x4 r = draw_block(key, STREAM_NSB, 0);
    float v = 0.0f;
    switch (kind) {
        case SAMPLE_NORMAL: v = normal_one(r.z, r.w); break;
        case SAMPLE_EXP_GAP: v = exp_gap(r.x, NEG_LN2 * a); break;
        case SAMPLE_BOREL: {
            // same four thresholds dct_create computes on the host
            double pn = exp(-(double)a), cdf = pn;
            float c[4];
            for (int n = 1; n <= 4; ++n) {
                c[n - 1] = (float)fmin(cdf, 1.0);
                pn *= (double)a * exp(-(double)a) * pow(1.0 + 1.0 / n, n - 1.0);
                cdf += pn;
            }
            v = (float)borel(r.y, a, make_float4(c[0], c[1], c[2], c[3]));
        } break;
        case SAMPLE_POISSON: { BlockReader rd(key, STREAM_BRANCH, 0); v = (float)poisson(rd, a); } break;
        case SAMPLE_BRANCHING: { BlockReader rd(key, STREAM_BRANCH, 0); v = (float)branching_total(rd, (int)b, a); } break;
        case SAMPLE_UNIFORM: v = u01_open1(r.x); break;
        case SAMPLE_BOREL_TOP: {
            double pn = exp(-(double)a), cdf = pn;
            float c[4];
            for (int n = 1; n <= 4; ++n) {
                c[n - 1] = (float)fmin(cdf, 1.0);
                pn *= (double)a * exp(-(double)a) * pow(1.0 + 1.0 / n, n - 1.0);
                cdf += pn;
            }
            v = (float)borel(0xFFFFFFFFu - (i << 9), a, make_float4(c[0], c[1], c[2], c[3]));
        } break;
        case SAMPLE_POISSON_TOP: {
            // poisson() reads its uniform from a block reader: hand it one whose buffer is preset
            BlockReader rd(key, STREAM_BRANCH, 0);
            rd.buf.x = 0xFFFFFFFFu - (i << 9); rd.have = 1;
            v = (float)poisson(rd, a);
        } break;
        case SAMPLE_NORMAL_RADIUS_TOP: { float z0, z1; normal_pair(i, 0x80000000u, z0, z1); v = sqrtf(z0 * z0 + z1 * z1); } break;
        case SAMPLE_POISSON_INV_TOP: v = (float)poisson_inv(u01_open1(0xFFFFFFFFu - (i << 9)), a, __expf(-a)); break;
        case SAMPLE_POISSON_TABLE_TOP: {
            float tab[POISSON_TABLE];
            poisson_table_fill(tab, 1, a, __expf(-a));
            v = (float)poisson_table_draw(u01_open1(0xFFFFFFFFu - (i << 9)), tab, 1, a, __expf(-a));
        } break;
        default: break;
    }
    out[i] = v;
}

}  // namespace dct
