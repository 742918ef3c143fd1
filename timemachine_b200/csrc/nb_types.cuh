// Device-side data layout of the nonbonded working set.
//
// The f64 API buffers (coords[N,3], params[N,4]) are gathered ONCE per evaluation into Hilbert order and cast to the
// kernel's arithmetic type, packed for 128-bit loads:
//     xw [slot] = {x, y, z, w}        (w = 4D decoupling coordinate, params[:,3])
//     qse[slot] = {q, sig, eps, 0}
// (the reference keeps 56 B/atom of f64 and casts per tile, k_nonbonded.cuh:134-178; f32 here moves 32 B/atom).
// Casting before any subtraction is exactly the reference's rounding sequence (SURVEY.md §8a cheat-sheet item 1).
#pragma once

#include "common.cuh"

namespace tmb {

template <typename Real> struct alignas(16) Vec4 {
    Real x, y, z, w;
};

static_assert(sizeof(Vec4<float>) == 16, "float4 layout");
static_assert(sizeof(Vec4<double>) == 32, "double4 layout");

// The interaction tile list: tile t pairs the 32 row slots of block tile_rows[t] with the 32 column slots
// tile_cols[32 t .. 32 t + 31] (entries >= K are padding). 132 B per tile, same information as the reference's
// ixn_tiles / ixn_atoms (neighborlist.cu:22-28) but emitted in runs of equal row block.
struct TileList {
    unsigned int *count;     // [1]
    int *rows;               // [capacity]
    unsigned int *cols;      // [capacity * 32]
    unsigned int capacity;   // tiles
    unsigned int *overflow;  // [2]: [0] set when a build ran out of capacity (cleared by the next build);
                             //      [1] sticky: the largest tile count that did not fit since the host last looked
};

} // namespace tmb
