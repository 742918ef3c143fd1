// Index along a 3-D Hilbert curve (Butz's algorithm).
//
// The orientation/starting conventions are those of Doug Moore's hilbert_c2i(nDims=3, nBits, coord), which the
// reference vendors (timemachine/cpp/src/vendored/hilbert.cpp:196-237) and calls with nBits = 8 on a 128^3 grid
// (hilbert_sort.cu:17-33).  The permutation produced by the sort is only "bit-exact with the reference" if this
// function reproduces that mapping for every cell, which tests/test_hilbert.py checks exhaustively (2,097,152 cells)
// against the vendored routine compiled into oracle/_ref/libhilbert_ref.so.
//
// Plain C++ so the same source is compiled for the device (LUT kernel) and for the host-side test shim.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define TMB_HD __host__ __device__
#else
#define TMB_HD
#endif

namespace tmb {

TMB_HD inline uint32_t hilbert3d_index(uint32_t c0, uint32_t c1, uint32_t c2, int nbits) {
    uint32_t index = 0;
    uint32_t rot = 0;  // current cyclic rotation of the 3-bit digit, 0..2
    uint32_t flip = 0; // reflection mask applied before rotating
    uint32_t above = 0; // raw interleaved digit of the level above
    for (int level = nbits - 1; level >= 0; --level) {
        const uint32_t raw = (((c2 >> level) & 1u) << 2) | (((c1 >> level) & 1u) << 1) | ((c0 >> level) & 1u);
        // Gray-code difference with the level above, then undo the current reflection and rotation
        uint32_t digit = (raw ^ above) ^ flip;
        digit = ((digit >> rot) | (digit << (3u - rot))) & 7u;
        index = (index << 3) | digit;
        above = raw;
        flip = 1u << rot;
        // next rotation: advance by one plus the (1-based) position of the lowest set bit when it lies in the low
        // two bits of the digit
        uint32_t low = digit & (0u - digit) & 3u;
        rot += 1u + (low == 1u ? 1u : (low == 2u ? 2u : 0u));
        if (rot >= 3u) {
            rot -= 3u;
        }
        if (rot >= 3u) {
            rot -= 3u;
        }
    }
    const uint32_t total_bits = 3u * static_cast<uint32_t>(nbits);
    // every third bit set, shifted down one: ...100100100
    uint32_t every_third = 0;
    for (uint32_t b = 0; b < total_bits; b += 3u) {
        every_third |= 1u << b;
    }
    index ^= every_third >> 1;
    // Gray decode (prefix xor from the top)
    for (uint32_t d = 1; d < total_bits; d <<= 1) {
        index ^= index >> d;
    }
    return index;
}

constexpr int HILBERT_GRID_DIM = 128; // reference k_hilbert.cuh:6
constexpr int HILBERT_N_BITS = 8;     // reference k_hilbert.cuh:9

} // namespace tmb
