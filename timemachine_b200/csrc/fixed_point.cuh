// Fixed-point numerical contract of the force path.
//
// Every pair/bond term is rounded to a 64-bit fixed-point integer BEFORE it is accumulated, so sums are
// order-independent and exclusions cancel all-pairs terms bit-exactly. Scales follow the reference
// (timemachine/cpp/src/fixed_point.hpp:5-11): forces/energies 2^36, du/dq 2^36, du/dsig 2^37, du/deps 2^38,
// du/dw 2^36. Rounding is round-half-even (reference k_fixed_point.cuh:10-24 is algebraically llrintf inside the int64
// range; outside it - clashing atoms, |force| >= 2^27 kJ/mol/nm - it yields specific garbage, which round_to_i64(float)
// reproduces operation for operation so that even those values are the reference's).
#pragma once

#include "common.cuh"

namespace tmb {

constexpr u64 FIXED_EXPONENT = 0x1000000000ull;         // 2^36
constexpr u64 FIXED_EXPONENT_DU_DCHARGE = 0x1000000000ull;
constexpr u64 FIXED_EXPONENT_DU_DSIG = 0x2000000000ull; // 2^37
constexpr u64 FIXED_EXPONENT_DU_DEPS = 0x4000000000ull; // 2^38
constexpr u64 FIXED_EXPONENT_DU_DW = 0x1000000000ull;

// The reference's float -> int64 (real_to_int64, k_fixed_point.cuh:10-24), operation for operation: the high word by a
// truncating, SATURATING 32-bit conversion of x / 2^32, the low word from the remainder.  Inside the int64 range this is
// round-half-even; beyond it the high word sticks at INT_MAX / wraps from INT_MIN - 1, and those are the reference's bits.
__device__ __forceinline__ i64 round_to_i64(float x) {
    const float z = x * 0x1.0p-32f;
    int hi = __float2int_rz(z);
    const float delta = __fmaf_rn(-0x1.0p32f, static_cast<float>(hi), x); // x - 2^32 * hi (the product is exact)
    const int test = __float_as_uint(delta) > 0xbf000000u ? 1 : 0;        // remainder below -0.5
    unsigned int lo = __float2uint_rn(fabsf(delta));
    lo = test ? 0u - lo : lo;
    const unsigned int hi_u = static_cast<unsigned int>(hi) - static_cast<unsigned int>(test); // wraps like the reference's int
    return static_cast<i64>((static_cast<u64>(hi_u) << 32) | lo);
}
__device__ __forceinline__ i64 round_to_i64(double x) { return __double2ll_rn(x); }
// One instruction, equal to the above for |x| < 2^63: the tile kernel's hot path, which knows its terms are in range
__device__ __forceinline__ i64 round_to_i64_in_range(float x) { return __float2ll_rn(x); }
__device__ __forceinline__ i64 round_to_i64_in_range(double x) { return __double2ll_rn(x); }

// value * EXP is exact in binary floating point (power of two) barring overflow
template <u64 EXP, typename Real> __device__ __forceinline__ u64 to_fixed(Real v) {
    return static_cast<u64>(round_to_i64(v * static_cast<Real>(EXP)));
}
template <u64 EXP, typename Real> __device__ __forceinline__ u64 to_fixed_in_range(Real v) {
    return static_cast<u64>(round_to_i64_in_range(v * static_cast<Real>(EXP)));
}

template <typename Real> __device__ __forceinline__ u64 to_fixed_force(Real v) { return to_fixed<FIXED_EXPONENT>(v); }

// Energies: non-finite or out-of-int64 terms are pinned to LLONG_MAX so that a clash can only be cancelled by
// the matching exclusion term, and sums are carried in int128 so overflow is detectable
// (reference k_fixed_point.cuh:88-98).
template <typename Real> __device__ __forceinline__ i128 energy_to_fixed(Real u_real) {
    Real u = u_real * static_cast<Real>(FIXED_EXPONENT);
    // |u| < 2^63 is exactly the reference's "(int128)u strictly inside (LLONG_MIN, LLONG_MAX)" for float and double
    if (!(fabs(u) < static_cast<Real>(9223372036854775808.0))) {
        return static_cast<i128>(LLONG_MAX);
    }
    return static_cast<i128>(round_to_i64_in_range(u));
}

template <typename Real> __host__ __device__ __forceinline__ Real fixed_to_real(u64 v) {
    return static_cast<Real>(static_cast<i64>(v)) / static_cast<Real>(FIXED_EXPONENT);
}

} // namespace tmb
