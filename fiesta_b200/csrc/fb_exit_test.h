// fiesta_b200 -- the bit logic of k_wavefront's exit test (-DWF_EXIT_TEST=1, fb_esdf.cu), as plain host/device functions so
// that tests/test_exit_test_logic.py can check it on the CPU against a brute-force restatement.
//
// Geometry (fb_common.cuh): a visit stages a 12 x 12 x 16 (x, y, z) box of records; the tile is x, y in [2, 10), z in [4, 12).
// `cmp` holds, per z-row, the records changed by the visit as bits (bit = box z), in a 16 x 16 row array with a 2-row zero
// border: row (rx, ry) of the box is cmp[(rx + 2) * 16 + (ry + 2)], so every neighbour row can be read without bounds checks.
// Direction k uses the z-row grouped order of k_wavefront's kd[] table (same row dz = -2,-1,+1,+2 | rows x-1, x+1, y-1, y+1
// with dz = -1,0,+1 | the four xy diagonals | x-2, x+2, y-2, y+2); koff[k] is its offset in box words.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD static inline
#endif

#define FBX_BOX 12
#define FBX_BOXZ 16
#define FBX_HALO 2
#define FBX_ZPAD 4
#define FBX_TILE 8
#define FBX_CODE_MASK 0x7fffffffu
// dirs_ (parameters.h:55-68) in the z-row grouped order used by k_wavefront
#define FBX_KD_INIT                                                                                          \
  {{0, 0, -2}, {0, 0, -1}, {0, 0, 1}, {0, 0, 2}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 0}, {1, 0, 1},       \
   {0, -1, -1}, {0, -1, 0}, {0, -1, 1}, {0, 1, -1}, {0, 1, 0}, {0, 1, 1}, {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},   \
   {-2, 0, 0}, {2, 0, 0}, {0, -2, 0}, {0, 2, 0}}

// Halo voxels of box row (rx, ry) that have a changed record of the tile among their 24 neighbours (bit = box z, 2..13).
FB_HD uint32_t fbx_row_candidates(const uint32_t *cmp, int rx, int ry) {
  const int p = (rx + 2) * 16 + (ry + 2);
  const uint32_t m0 = cmp[p];
  const uint32_t f4 = cmp[p - 16] | cmp[p + 16] | cmp[p - 1] | cmp[p + 1];
  uint32_t act = (m0 << 1) | (m0 >> 1) | (m0 << 2) | (m0 >> 2) | f4 | (f4 << 1) | (f4 >> 1) |
                 cmp[p - 17] | cmp[p - 15] | cmp[p + 15] | cmp[p + 17] | cmp[p - 32] | cmp[p + 32] | cmp[p - 2] | cmp[p + 2];
  const bool inner = rx >= FBX_HALO && rx < FBX_HALO + FBX_TILE && ry >= FBX_HALO && ry < FBX_HALO + FBX_TILE;
  return act & (inner ? (0x3ffcu & ~(0xffu << FBX_ZPAD)) : 0x3ffcu);
}

// The 24-bit set of neighbours of box voxel (rx, ry, zb) that are changed records of the tile (bit k = direction k).
FB_HD uint32_t fbx_changed_neighbours(const uint32_t *cmp, int rx, int ry, int zb) {
  const int p = (rx + 2) * 16 + (ry + 2);
  const uint32_t m0 = cmp[p];
  return ((m0 >> (zb - 2)) & 3u) | (((m0 >> (zb + 1)) & 3u) << 2) |
         (((cmp[p - 16] >> (zb - 1)) & 7u) << 4) | (((cmp[p + 16] >> (zb - 1)) & 7u) << 7) |
         (((cmp[p - 1] >> (zb - 1)) & 7u) << 10) | (((cmp[p + 1] >> (zb - 1)) & 7u) << 13) |
         (((cmp[p - 17] >> zb) & 1u) << 16) | (((cmp[p - 15] >> zb) & 1u) << 17) |
         (((cmp[p + 15] >> zb) & 1u) << 18) | (((cmp[p + 17] >> zb) & 1u) << 19) |
         (((cmp[p - 32] >> zb) & 1u) << 20) | (((cmp[p + 32] >> zb) & 1u) << 21) |
         (((cmp[p - 2] >> zb) & 1u) << 22) | (((cmp[p + 2] >> zb) & 1u) << 23);
}

FB_HD unsigned fbx_dist2(uint32_t c, int x, int y, int z) {
  const int ox = (int)((c & FBX_CODE_MASK) >> 20) - 1 - x, oy = (int)((c >> 10) & 1023u) - y, oz = (int)(c & 1023u) - z;
  return (unsigned)(ox * ox + oy * oy + oz * oz);
}

// Would the border voxel at box position (rx, ry, zb) = grid (x, y, z), holding record `cy` (>= 1: observed), take one of
// the changed records selected by `m` (a subset of fbx_changed_neighbours)?  Strict improvement, ties to the smaller code.
FB_HD bool fbx_improves(const uint32_t *V, const int *koff, int rx, int ry, int zb, int x, int y, int z, uint32_t cy, uint32_t m) {
  const int sidx = (rx * FBX_BOX + ry) * FBX_BOXZ + zb;
  const unsigned dy = cy >= 2u ? fbx_dist2(cy, x, y, z) : 0xffffffffu;
  bool improves = false;
  while (m) {
#ifdef __CUDA_ARCH__
    const int k = __ffs((int)m) - 1;
#else
    const int k = __builtin_ctz(m);
#endif
    m &= m - 1u;
    const uint32_t c = V[sidx + koff[k]] & FBX_CODE_MASK;               // a changed record of this tile (interior: in bounds)
    if (c >= 2u && c != cy) {
      const unsigned d = fbx_dist2(c, x, y, z);
      improves = improves || d < dy || (d == dy && c < cy);
    }
  }
  return improves;
}

// Direction bit of the neighbour tile a halo voxel belongs to: (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1).
FB_HD int fbx_dir_bit(int rx, int ry, int zb) {
  const int ox = rx < FBX_HALO ? -1 : rx >= FBX_HALO + FBX_TILE ? 1 : 0, oy = ry < FBX_HALO ? -1 : ry >= FBX_HALO + FBX_TILE ? 1 : 0;
  const int oz = zb < FBX_ZPAD ? -1 : zb >= FBX_ZPAD + FBX_TILE ? 1 : 0;
  return (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1);
}
