// fiesta_b200 -- host interface of the order-exact mode (fb_exact.cu).
#pragma once
#include "fb_common.cuh"

struct FbExactStats {
  unsigned long long expansions;      // == the reference's "Expanding N nodes" (ESDFMap.cpp:347,394)
  unsigned long long voxels_changed;  // accepted final writes over all generations
  unsigned generations, eval_rounds, reseed_rounds, dependants;
};

struct FbExact {
  unsigned long long *MB;      // per voxel: {queue position | behaviour | code} of its live entry in the current generation
  unsigned long long *LS;      // per voxel: time of the last relink into a dependant list
  unsigned long long *tkey;    // per voxel: serial time of the first pending observation
  uint32_t *touched;           // voxels with pending observations (unordered; ordered by tkey at integration)
  unsigned long long tclock;   // relink clock
  unsigned long long key_base; // observation clock
  unsigned *d_count, *d_flag, *h_count;
  uint4 *SUM;                  // per voxel offer summary of the current generation: {first ts, best d, best ts, best code}
  uint32_t *SUMg;              // per voxel: generation id for which SUM is valid (= the voxel is a push/pull target)
  unsigned gen_id;
  uint32_t *targets, *work; size_t cap_targets, cap_work;
  uint32_t *tdirty;            // per 8^3 tile: evaluation round in which its elements must be re-evaluated
  unsigned eval_clock;
  uint32_t *E[2]; size_t cap_E[2];
  uint32_t *slotv, *slotc, *sel; uint8_t *slotf; size_t cap_slotv, cap_slotc, cap_sel, cap_slotf;
  unsigned long long *k1, *k2, *k1b, *k2b; size_t cap_k1, cap_k2, cap_k1b, cap_k2b;
  uint32_t *dv, *idx[2], *deps, *nc[2]; size_t cap_dv, cap_idx[2], cap_deps, cap_nc[2];
  uint8_t *flags, *flags2; size_t cap_flags, cap_flags2;
  void *cub_tmp; size_t cub_bytes;
  char err[256];
};

cudaError_t fb_exact_init(FbExact *X, const FbGeom &g, cudaStream_t s);
void fb_exact_free(FbExact *X);
cudaError_t fb_exact_update_occupancy(FbExact *X, const FbGeom &g, unsigned n, unsigned long long *cnt, double *occ, uint32_t *cobs, uint32_t *occbits,
                                      uint32_t **ins, size_t *cap_ins, unsigned *n_ins, uint32_t **del, size_t *cap_del, unsigned *n_del,
                                      int global_map, const double L[5], cudaStream_t s, int *launches);
cudaError_t fb_exact_update_esdf(FbExact *X, const FbGeom &g, uint32_t *cobs, uint32_t *scratch, const double *occ, const uint32_t *occbits, double l_occ,
                                 const uint32_t *ins, unsigned n_ins, const uint32_t *del, unsigned n_del, cudaStream_t s, FbExactStats *st, int *launches);
