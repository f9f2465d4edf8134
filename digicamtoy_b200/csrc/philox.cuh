// Counter-based Philox4x32 (Salmon et al., SC'11) for the trace generator.
//
// The reference draws everything from NumPy's global MT19937 stream
// (reference digicamtoy/tracegenerator/__init__.py:233,236,245,248,266,271,337,342,369), which is
// inherently serial.  Here every random number is a pure function of
//     key     = 64-bit user seed
//     counter = (block index within a stream, stream id | event_hi, pixel, event_lo)
// so the ADC cube does not depend on grid size, chunking, or how events are split across GPUs.
#pragma once
#include <stdint.h>

namespace dct {

#ifndef DCT_PHILOX_ROUNDS
#define DCT_PHILOX_ROUNDS 10
#endif

struct u32x4 { uint32_t x, y, z, w; };

// Stream ids (upper byte of counter word 1).  One (event, pixel) owns all of them.
enum : uint32_t {
    STREAM_NSB      = 0u,   // block i -> photo-electron i of the continuous NSB process
    STREAM_SIGNAL   = 1u,   // block k -> jitter / smearing of signal pulse k
    STREAM_BRANCH   = 2u,   // variable-length: Poisson(n_photon) and crosstalk generations of the pulses
    STREAM_NOISE    = 3u,   // block q -> electronic noise of kept bins 4q..4q+3
    STREAM_SUBBIN   = 4u,   // on-grid NSB mode, rate * dt >= 12: block bs -> simulated bin bs; (bs+1)<<16.. overflow
    STREAM_SUBBIN_GP = 5u,  // on-grid NSB mode (sub_binning > 0): block m -> simulated bins 2m, 2m+1 (total pe, normals)
};

__host__ __device__ __forceinline__ u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < DCT_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x;
        const uint64_t p1 = (uint64_t)M1 * c.z;
        u32x4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// Round keys of one seed, computed once per thread and kept in registers where the budget allows
// (saves the 2 x ROUNDS key-schedule adds per block: ~20 of ~60 instructions).
struct PhiloxKeys { uint32_t k0[DCT_PHILOX_ROUNDS], k1[DCT_PHILOX_ROUNDS]; };

__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys ks;
#pragma unroll
    for (int r = 0; r < DCT_PHILOX_ROUNDS; ++r) {
        ks.k0[r] = k0 + (uint32_t)r * 0x9E3779B9u;
        ks.k1[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return ks;
}

__host__ __device__ __forceinline__ u32x4 philox4x32_keyed(u32x4 c, const PhiloxKeys& ks) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < DCT_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x;
        const uint64_t p1 = (uint64_t)M1 * c.z;
        u32x4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ ks.k0[r];
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ ks.k1[r];
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

// Identity of one (event, pixel) cell of the cube.
struct StreamKey {
    uint32_t k0, k1;        // seed
    uint32_t ev_lo, ev_hi;  // global event index (ev_hi < 2^24)
    uint32_t pixel;
};

// The key schedule (k + r W) depends only on the seed: kernels that take it as a parameter array
// XOR the round keys straight from the constant bank -- no key registers, no adds (20 of ~60
// instructions per block).  Bit-identical to philox4x32(c, k0, k1).
struct RoundKeys { uint32_t k[2 * DCT_PHILOX_ROUNDS]; };

__host__ __device__ inline RoundKeys make_round_keys(uint32_t k0, uint32_t k1) {
    RoundKeys rk;
    for (int r = 0; r < DCT_PHILOX_ROUNDS; ++r) {
        rk.k[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        rk.k[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return rk;
}

__host__ __device__ __forceinline__ u32x4 philox4x32_rk(u32x4 c, const RoundKeys& rk) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < DCT_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x;
        const uint64_t p1 = (uint64_t)M1 * c.z;
        u32x4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ rk.k[2 * r];
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ rk.k[2 * r + 1];
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

#ifndef DCT_USE_ROUND_KEYS
#define DCT_USE_ROUND_KEYS 1     // 0: key schedule computed per block (the round-1 form), for comparison
#endif
__host__ __device__ __forceinline__ u32x4 draw_block_rk(const RoundKeys& rk, const StreamKey& s, uint32_t stream,
                                                        uint32_t block) {
#if !DCT_USE_ROUND_KEYS
    return draw_block(s, stream, block);
#endif
    u32x4 c;
    c.x = block;
    c.y = (stream << 24) | (s.ev_hi & 0x00FFFFFFu);
    c.z = s.pixel;
    c.w = s.ev_lo;
    return philox4x32_rk(c, rk);
}

__host__ __device__ __forceinline__ u32x4 draw_block(const StreamKey& s, uint32_t stream, uint32_t block) {
    u32x4 c;
    c.x = block;
    c.y = (stream << 24) | (s.ev_hi & 0x00FFFFFFu);
    c.z = s.pixel;
    c.w = s.ev_lo;
    return philox4x32(c, s.k0, s.k1);
}

// Uniform in (0, 1]: never 0, so log() is always finite.  24-bit mantissa, built without I2F.
__device__ __forceinline__ float u01_open0(uint32_t r) {
    return __uint_as_float(0x3f800000u | (r >> 9)) - (1.0f - 5.9604645e-08f);   // (0,1]
}
// Uniform in (0, 1] on a 2^-32 grid near 0 (the float conversion keeps 24 significant bits, so the
// resolution is relative).  For Box-Muller radii: the 23-bit form stops at sqrt(-2 ln 2^-24) =
// 5.77 sigma -- eight samples in 1e9 too few beyond it; this one reaches 6.77 sigma (1.3e-11).
#ifndef DCT_NORMAL_TAIL32
#define DCT_NORMAL_TAIL32 1
#endif
__device__ __forceinline__ float u01_tail(uint32_t r) {
#if DCT_NORMAL_TAIL32
    return fmaf(__uint2float_rn(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (r + 0.5) 2^-32
#else
    return u01_open0(r);
#endif
}
// Uniform in [0, 1).
__device__ __forceinline__ float u01_open1(uint32_t r) {
    return __uint_as_float(0x3f800000u | (r >> 9)) - 1.0f;
}

// A sequential reader over one variable-length stream (rejection samplers need it).
struct BlockReader {
    StreamKey key;
    uint32_t stream;
    uint32_t block;
    u32x4 buf;
    int have;
    __device__ __forceinline__ BlockReader(const StreamKey& k, uint32_t s, uint32_t first_block)
        : key(k), stream(s), block(first_block), have(0) {}
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) { buf = draw_block(key, stream, block++); have = 4; }
        uint32_t r = buf.x;
        buf.x = buf.y; buf.y = buf.z; buf.z = buf.w;
        --have;
        return r;
    }
};

}  // namespace dct
