// costas.cuh — per-symbol arithmetic of the Costas loop (PM/costas_loop.hpp:92-149) for sm_100a.
//
// The loop is a sequential second-order recurrence over symbols, but a "syncword_phase" tag resets its
// whole state (set_phase(): _phase = tag value, _freq = 0, PM/costas_loop.hpp:38-45), so the stretches
// between two such tags — one packet each in the receiver (PM/packet_receiver.hpp:125, 211-218) — are
// independent of each other: the GPU runs one thread per stretch.
//
// ARITHMETIC CONTRACT (mirrored bit-for-bit by oracle/oracle.hpp, CostasLoop with trig = Mirror):
//   * every add / sub / mul below is a separately rounded IEEE binary32 op, in the reference's
//     evaluation order (std::complex<float> product = (ac - bd, ad + bc), no contraction);
//   * std::cos / std::sin of the reference (libm, < 1 ulp) are replaced by b200_sincosf below:
//     Cody-Waite reduction by pi/2 (quotient by magic-number rounding, three fmaf steps), then the classic single-precision minimax
//     polynomials on |r| <= pi/4 (max abs error 1.2e-7 on [-4, 4], tests/test_oracle_golden.py).
//     The libm oracle and the mirror oracle agree to north_star's filter-output tolerance (rel-L2 < 1e-5).
#pragma once
#include <cuda_runtime.h>

namespace b200sync {

enum ClConstellation : int { kClPilot = 0, kClBpsk = 1, kClQpsk = 2 };

__device__ __forceinline__ void b200_sincosf(float x, float& s, float& c) {
    // nearest multiple of pi/2 by the add-and-subtract-1.5*2^23 idiom (two FADDs on the FMA pipe instead of
    // FRND + F2I on the quarter-rate conversion pipe, which sit on the loop's critical path); equal to
    // rintf() for |x| < 2^22 * pi/2, and the oracle mirrors this very formulation
    const float qb = __fadd_rn(__fmul_rn(x, 0.636619772367581343f), 12582912.0f);
    const float q = __fsub_rn(qb, 12582912.0f);
    float r = __fmaf_rn(q, -1.5703125f, x);
    r = __fmaf_rn(q, -4.837512969970703125e-4f, r);
    r = __fmaf_rn(q, -7.54978995489188216e-8f, r);
    const float z = __fmul_rn(r, r);
    float ps = __fmaf_rn(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, z, -1.6666654611e-1f);
    ps = __fmul_rn(ps, z);
    const float sr = __fmaf_rn(ps, r, r);
    float pc = __fmaf_rn(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, z, 4.166664568298827e-2f);
    pc = __fmul_rn(pc, __fmul_rn(z, z));
    const float cr = __fadd_rn(__fmaf_rn(-0.5f, z, 1.0f), pc);
    const int n = __float_as_int(qb) & 3;  // 12582912 = 0xC00000 is a multiple of 4: the low mantissa bits are q mod 4
    const float s0 = (n & 1) ? cr : sr;
    const float c0 = (n & 1) ? sr : cr;
    s = (n & 2) ? -s0 : s0;
    c = ((n + 1) & 2) ? -c0 : c0;
}

struct ClState {
    float phase, freq;
};

// one iteration of the loop body, PM/costas_loop.hpp:112-146
template <int CONSTELLATION>
__device__ __forceinline__ float2 costas_step(float2 x, ClState& st, float k1, float k2) {
    float sn, cs;
    b200_sincosf(st.phase, sn, cs);
    const float lo_re = cs, lo_im = -sn;  // :113-114
    float2 z;                              // :115 inSpan[j] * lo
    z.x = __fsub_rn(__fmul_rn(x.x, lo_re), __fmul_rn(x.y, lo_im));
    z.y = __fadd_rn(__fmul_rn(x.x, lo_im), __fmul_rn(x.y, lo_re));
    float error;
    if constexpr (CONSTELLATION == kClPilot) {
        error = z.y;  // :121
    } else if constexpr (CONSTELLATION == kClBpsk) {
        error = __fmul_rn(z.x, z.y);  // :125
    } else {
        error = __fadd_rn(z.x > 0.0f ? z.y : -z.y, z.y > 0.0f ? -z.x : z.x);  // :131-132
    }
    st.freq = __fadd_rn(st.freq, __fmul_rn(k2, error));                      // :138
    st.phase = __fadd_rn(st.phase, __fadd_rn(__fmul_rn(k1, error), st.freq));  // :139
    constexpr float kPi = 3.14159274101257324f, kTwoPi = 6.28318548202514648f;  // pi_v<float>, 2 * pi_v<float>
    if (st.phase >= kPi) st.phase = __fsub_rn(st.phase, kTwoPi);          // :140-144
    else if (st.phase < -kPi) st.phase = __fadd_rn(st.phase, kTwoPi);
    return z;
}

}  // namespace b200sync
