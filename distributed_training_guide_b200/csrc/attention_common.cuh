// Shared by the attention forward / backward kernels.
#pragma once
#include <cuda.h>

namespace dtg {
// TMA descriptor over a [B, S, heads, 128] bf16 tensor: dims {128, heads, S, B}, box {64, 1, rows, 1},
// 128-byte swizzle.  One [rows x 128] head tile = two loads (columns 0-63 and 64-127).
CUtensorMap make_tmap_heads(const void* base, int B, int S, int heads, int box_rows);
}  // namespace dtg
