// Fixed-point numerical contract of the force path.
//
// Every pair/bond term is rounded to a 64-bit fixed-point integer BEFORE it is accumulated, so sums are
// order-independent and exclusions cancel all-pairs terms bit-exactly. Scales follow the reference
// (timemachine/cpp/src/fixed_point.hpp:5-11): forces/energies 2^36, du/dq 2^36, du/dsig 2^37, du/deps 2^38,
// du/dw 2^36. Rounding is round-half-even (reference k_fixed_point.cuh:10-24 is algebraically llrintf).
#pragma once

#include "common.cuh"

namespace tmb {

constexpr u64 FIXED_EXPONENT = 0x1000000000ull;         // 2^36
constexpr u64 FIXED_EXPONENT_DU_DCHARGE = 0x1000000000ull;
constexpr u64 FIXED_EXPONENT_DU_DSIG = 0x2000000000ull; // 2^37
constexpr u64 FIXED_EXPONENT_DU_DEPS = 0x4000000000ull; // 2^38
constexpr u64 FIXED_EXPONENT_DU_DW = 0x1000000000ull;

__device__ __forceinline__ i64 round_to_i64(float x) { return __float2ll_rn(x); }
__device__ __forceinline__ i64 round_to_i64(double x) { return __double2ll_rn(x); }

// value * EXP is exact in binary floating point (power of two) barring overflow
template <u64 EXP, typename Real> __device__ __forceinline__ u64 to_fixed(Real v) {
    return static_cast<u64>(round_to_i64(v * static_cast<Real>(EXP)));
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
    return static_cast<i128>(round_to_i64(u));
}

template <typename Real> __host__ __device__ __forceinline__ Real fixed_to_real(u64 v) {
    return static_cast<Real>(static_cast<i64>(v)) / static_cast<Real>(FIXED_EXPONENT);
}

} // namespace tmb
