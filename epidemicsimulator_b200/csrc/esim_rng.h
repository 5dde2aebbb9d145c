// Counter-based random stream of the B200 step kernels.
//
// The reference draws from rand 0.8 `thread_rng()` (sim/src/simulator.rs:102,342; citizen.rs:42-45,242),
// which is neither seedable nor order-independent.  The replacement is Philox4x32-10
// (Salmon et al., SC'11; the generator behind cuRAND's CURAND_RNG_PSEUDO_PHILOX4_32_10):
//
//     key     = (seed_lo, seed_hi)
//     counter = (citizen global index | draw index, time_step, slot_pair, domain)
//
// so every draw is a pure function of (seed, citizen, step, slot) and the result of a step does not
// depend on thread order, on the number of GPUs or on how citizens are laid out.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ESIM_HD __host__ __device__ __forceinline__
#else
#define ESIM_HD inline
#endif

namespace esim {

// domains (counter word 3)
constexpr uint32_t DOM_BUILDING = 0;  // building trials: slot 0 = household, slot 1+j = j-th workplace/room trial
constexpr uint32_t DOM_PT       = 1;  // public transport: word 0 = shuffle key, words 2..3 = the bus trial
constexpr uint32_t DOM_VAX      = 2;  // vaccination candidates: counter word 0 = draw index

struct Philox4 { uint32_t v[4]; };

ESIM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

ESIM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    Philox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

// 52 random bits the way rand 0.8's Uniform<f64> consumes a u64: `next_u64() >> 12`.
ESIM_HD uint64_t u52_from(const Philox4& p, uint32_t parity) {
    // selects instead of a dynamic index keep the block in registers
    const uint32_t lo = parity ? p.v[2] : p.v[0];
    const uint32_t hi = parity ? p.v[3] : p.v[1];
    const uint64_t x = ((uint64_t)hi << 32) | (uint64_t)lo;
    return x >> 12;
}

// Uniform::<f64>::new_inclusive(0.0, 1.0).sample():  value0_1 * scale + low with
// value0_1 = m * 2^-52 and scale = 1 + 2^-52 (the largest scale for which the maximum stays <= 1.0).
ESIM_HD double u01_from_u52(uint64_t m) {
    return ((double)m * 0x1p-52) * (1.0 + 0x1p-52);
}

// The draw used by Citizen::expose (citizen.rs:242): slot -> (slot >> 1) selects the Philox block,
// (slot & 1) the 64-bit half.
ESIM_HD uint64_t trial_u52(uint64_t seed, uint32_t citizen, uint32_t step, uint32_t slot) {
    const Philox4 p = philox4x32_10(citizen, step, slot >> 1, DOM_BUILDING, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u52_from(p, slot & 1u);
}

// Vaccination candidate `draw` of step `step`: uniform index in [0, n)
ESIM_HD uint32_t vax_candidate(uint64_t seed, uint32_t draw, uint32_t step, uint32_t n) {
    const Philox4 p = philox4x32_10(draw, step, 0u, DOM_VAX, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t x = ((uint64_t)p.v[1] << 32) | (uint64_t)p.v[0];
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umul64hi(x, (uint64_t)n);
#else
    return (uint32_t)(((unsigned __int128)x * (unsigned __int128)n) >> 64);
#endif
}

}  // namespace esim
