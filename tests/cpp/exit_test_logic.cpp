// CPU unit test of fiesta_b200/csrc/fb_exit_test.h (the bit logic of k_wavefront's exit test): random 12x12x16 boxes with a
// random set of changed tile records; every helper is compared with a brute-force restatement over the 24 directions.
// Built and run by tests/test_exit_test_logic.py.  Prints "OK <trials> <voxels checked> <improving>" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "fb_exit_test.h"

static const int KD[24][3] = FBX_KD_INIT;

static uint32_t pack(int x, int y, int z) { return ((uint32_t)(x + 1) << 20) | ((uint32_t)y << 10) | (uint32_t)z; }

int main(int argc, char **argv) {
  const int trials = argc > 1 ? atoi(argv[1]) : 300;
  std::mt19937 rng(12345);
  int koff[24];
  for (int k = 0; k < 24; ++k) koff[k] = KD[k][0] * (FBX_BOX * FBX_BOXZ) + KD[k][1] * FBX_BOXZ + KD[k][2];
  long long checked = 0, improving = 0;
  for (int t = 0; t < trials; ++t) {
    const int x0 = 8 * (int)(rng() % 20) + 8, y0 = 8 * (int)(rng() % 20) + 8, z0 = 8 * (int)(rng() % 20) + 8;   // tile origin in the grid
    static uint32_t V[FBX_BOX * FBX_BOX * FBX_BOXZ];
    static unsigned char chg[FBX_BOX][FBX_BOX][FBX_BOXZ];
    uint32_t cmp[16 * 16];
    memset(cmp, 0, sizeof(cmp));
    memset(chg, 0, sizeof(chg));
    const int nobs = 1 + (int)(rng() % 6);
    uint32_t obs[6];
    for (int i = 0; i < nobs; ++i) obs[i] = pack(x0 - 6 + (int)(rng() % 20), y0 - 6 + (int)(rng() % 20), z0 - 6 + (int)(rng() % 20));
    const unsigned density = 1 + rng() % 60;                            // per cent of the tile records that changed
    for (int rx = 0; rx < FBX_BOX; ++rx)
      for (int ry = 0; ry < FBX_BOX; ++ry)
        for (int zb = 0; zb < FBX_BOXZ; ++zb) {
          const unsigned r = rng() % 100;
          uint32_t c = r < 8 ? 0u : r < 20 ? 1u : obs[rng() % nobs];   // unknown / no obstacle / some obstacle
          if (rng() % 7 == 0) c |= 0x80000000u;                         // stale FRESH bits must be ignored
          V[(rx * FBX_BOX + ry) * FBX_BOXZ + zb] = c;
          const bool interior = rx >= 2 && rx < 10 && ry >= 2 && ry < 10 && zb >= 4 && zb < 12;
          if (interior && (c & FBX_CODE_MASK) >= 2u && rng() % 100 < density) {   // a changed record always holds an obstacle
            chg[rx][ry][zb] = 1;
            cmp[(rx + 2) * 16 + (ry + 2)] |= 1u << zb;
          }
        }
    for (int rx = 0; rx < FBX_BOX; ++rx)
      for (int ry = 0; ry < FBX_BOX; ++ry) {
        const uint32_t act = fbx_row_candidates(cmp, rx, ry);
        for (int zb = 0; zb < FBX_BOXZ; ++zb) {
          const bool interior = rx >= 2 && rx < 10 && ry >= 2 && ry < 10 && zb >= 4 && zb < 12;
          const bool reach = zb >= 2 && zb <= 13;                        // z pad beyond the 2-voxel halo
          uint32_t want = 0;
          if (reach)
            for (int k = 0; k < 24; ++k) {
              const int nx = rx + KD[k][0], ny = ry + KD[k][1], nz = zb + KD[k][2];
              if (nx >= 2 && nx < 10 && ny >= 2 && ny < 10 && nz >= 4 && nz < 12 && chg[nx][ny][nz]) want |= 1u << k;
            }
          const bool listed = (act >> zb) & 1u;
          if (listed != (!interior && reach && want != 0u)) { printf("row_candidates mismatch trial %d at %d %d %d\n", t, rx, ry, zb); return 1; }
          if (!reach || interior) continue;
          const uint32_t got = fbx_changed_neighbours(cmp, rx, ry, zb);
          if (got != want) { printf("changed_neighbours mismatch trial %d at %d %d %d: %06x vs %06x\n", t, rx, ry, zb, got, want); return 1; }
          const uint32_t cy = V[(rx * FBX_BOX + ry) * FBX_BOXZ + zb] & FBX_CODE_MASK;
          if (cy == 0u) continue;
          const int x = x0 - 2 + rx, y = y0 - 2 + ry, z = z0 - 4 + zb;
          bool exp = false;
          const unsigned dy = cy >= 2u ? fbx_dist2(cy, x, y, z) : 0xffffffffu;
          for (int k = 0; k < 24; ++k)
            if ((want >> k) & 1u) {
              const uint32_t c = V[((rx + KD[k][0]) * FBX_BOX + ry + KD[k][1]) * FBX_BOXZ + zb + KD[k][2]] & FBX_CODE_MASK;
              if (c >= 2u && c != cy) { const unsigned d = fbx_dist2(c, x, y, z); if (d < dy || (d == dy && c < cy)) exp = true; }
            }
          // the kernel splits the mask between lanes: the OR over the slices must equal the whole
          const bool whole = fbx_improves(V, koff, rx, ry, zb, x, y, z, cy, got);
          const bool split = (got & 0x555555u ? fbx_improves(V, koff, rx, ry, zb, x, y, z, cy, got & 0x555555u) : false) ||
                             (got & 0xaaaaaau ? fbx_improves(V, koff, rx, ry, zb, x, y, z, cy, got & 0xaaaaaau) : false);
          if (whole != exp || split != exp) { printf("improves mismatch trial %d at %d %d %d\n", t, rx, ry, zb); return 1; }
          const int ox = rx < 2 ? -1 : rx > 9 ? 1 : 0, oy = ry < 2 ? -1 : ry > 9 ? 1 : 0, oz = zb < 4 ? -1 : zb > 11 ? 1 : 0;
          if (fbx_dir_bit(rx, ry, zb) != (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1)) { printf("dir_bit mismatch\n"); return 1; }
          ++checked;
          improving += exp;
        }
      }
  }
  printf("OK %d %lld %lld\n", trials, checked, improving);
  return 0;
}
