// fiesta_b200 -- ORDER-EXACT mode of UpdateOccupancy / UpdateESDF (FIESTA_MODE_EXACT).
//
// The reference result is a function of its sequential FIFO order (/root/reference/src/ESDFMap.cpp:273-398): seeds in
// insert_queue_/delete_queue_ order, dependants of a deleted obstacle in LIFO list order (:301-334), neighbours in dirs_
// order, strict improvement, and every queue element seeing the writes of all earlier elements.  This file reproduces
// that order with data-parallel kernels (the CPU model of exactly this formulation is oracle/exact_model.c, which matches
// the sequential reference voxel for voxel):
//
//  * Every FIFO generation is one list E of (voxel) entries in queue order.  Element i at direction k acts at the
//    timestamp ts = 32*i + k (its pull acts at 32*i + 24).
//  * A voxel's state "as seen at time T" is a pure function of the snapshot at generation start and of the behaviour
//    (dead / pulled code / pushes code) of the <= 25 elements that can write it: the lexicographic minimum (distance,
//    timestamp) over their offers with timestamp < T that beat the snapshot -- exactly what a sequence of strict `>` tests
//    in timestamp order leaves behind (x_state()).
//  * An element's behaviour depends only on states at its own pop time (k_x_eval); it is kept, together with the entry's
//    queue position, in one packed per-voxel word so that a reader learns everything about a potential writer with one load.  Starting from "everybody pushes its
//    snapshot code", the behaviours are re-evaluated until none changes; element i is right once all elements before it
//    are, so the fixpoint is the sequential execution (2-3 rounds in practice).
//  * The last accepted write to a voxel in a generation is the lexicographic minimum over ALL offers; those writes, in
//    timestamp order, are the live entries of the next generation (k_x_commit + ordered compaction).  Non-final accepted
//    writes only create entries the reference skips as stale (:345), so dropping them changes nothing.
//  * The doubly linked dependant lists are replaced by a per-voxel link time LS (time of the last relink; every accepted
//    write relinks at the list front, :24-42): dependants of deleted obstacles are found by a dense scan and ordered by
//    (position of the obstacle in delete_queue_, link time descending) = the order of the reference's list walk; their
//    re-seeding ("first valid neighbour in dirs_ order", :308-321, which sees earlier re-seeded dependants) is iterated to
//    its fixpoint the same way.
//  * occupancy_queue_ order = order of first observation: every observation carries its serial time (host event number,
//    or point index and position along the ray); the per-voxel minimum orders the integration, so insert_queue_ /
//    delete_queue_ come out in the reference's order.
#include <cub/cub.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include "fb_common.cuh"
#include "fb_exact.h"

#define XNONE 0xffffffffu
#define X_DEAD 0ull
#define X_PULL 1ull
#define X_PUSH 2ull

__constant__ int x_dirs[24][3] = {
    {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
    {-1, -1, 0}, {1, 1, 0}, {0, -1, -1}, {0, 1, 1}, {-1, 0, -1}, {1, 0, 1},
    {-1, 1, 0}, {1, -1, 0}, {0, -1, 1}, {0, 1, -1}, {1, 0, -1}, {-1, 0, 1},
    {-2, 0, 0}, {2, 0, 0}, {0, -2, 0}, {0, 2, 0}, {0, 0, -2}, {0, 0, 2}};

__device__ __forceinline__ void x_coords(const FbGeom &g, uint32_t ii, int &x, int &y, int &z) {
  z = ii % (unsigned)g.pz; const unsigned xy = ii / (unsigned)g.pz; y = xy % (unsigned)g.gy; x = xy / (unsigned)g.gy;
}
__device__ __forceinline__ unsigned x_d2(int x, int y, int z, uint32_t c) {
  int ox, oy, oz; fb_unpack(c, ox, oy, oz); ox -= x; oy -= y; oz -= z;
  return (unsigned)(ox * ox + oy * oy + oz * oz);
}
__device__ __forceinline__ unsigned x_dist_of(int x, int y, int z, uint32_t c) { return c < 2u ? 0xffffffffu : x_d2(x, y, z, c); }

struct XState { unsigned d; uint32_t c; unsigned ts; };

// Per-voxel packed word of the current generation: {queue position:27 | kind:2 | code:31}, all ones = no live entry here.
// One 8-byte load tells a reader everything about a potential writer.
#define XMB_NONE 0xffffffffffffffffull
__device__ __forceinline__ unsigned long long x_mb(unsigned i, unsigned long long kind, uint32_t code) {
  return ((unsigned long long)i << 33) | (kind << 31) | (unsigned long long)(code & FB_CODE_MASK);
}
__device__ __forceinline__ unsigned x_mb_idx(unsigned long long w) { return (unsigned)(w >> 33); }
__device__ __forceinline__ unsigned long long x_mb_kind(unsigned long long w) { return (w >> 31) & 3ull; }
__device__ __forceinline__ uint32_t x_mb_code(unsigned long long w) { return (uint32_t)(w & FB_CODE_MASK); }

// State of voxel (x,y,z) as seen at time T (exclusive) given the behaviours of this generation's elements.
__device__ __forceinline__ XState x_state(const FbGeom &g, const uint32_t *cobs, const unsigned long long *MB, int x, int y, int z, unsigned T) {
  XState s;
  const long long v = fb_ii(g, x, y, z);
  s.c = cobs[v] & FB_CODE_MASK; s.d = x_dist_of(x, y, z, s.c); s.ts = XNONE;
  const unsigned d0 = s.d;
  if (s.c == FB_UNKNOWN) return s;                             // never observed: distance_ = -10000 is never > tmp (:382)
  if (!fb_in_range(g, x, y, z)) return s;                     // pushes only go to voxels inside the update box (:378)
#pragma unroll 8
  for (int k = 0; k < 24; ++k) {
    const int qx = x - x_dirs[k][0], qy = y - x_dirs[k][1], qz = z - x_dirs[k][2];
    if (!fb_in_grid(g, qx, qy, qz)) continue;
    const unsigned long long w = MB[fb_ii(g, qx, qy, qz)];
    if (w == XMB_NONE || x_mb_kind(w) != X_PUSH) continue;
    const unsigned ts = x_mb_idx(w) * 32u + (unsigned)k;
    if (ts >= T) continue;
    const uint32_t c = x_mb_code(w);
    const unsigned d = x_d2(x, y, z, c);
    if (d < d0 && (d < s.d || (d == s.d && ts < s.ts))) { s.d = d; s.c = c; s.ts = ts; }
  }
  const unsigned long long w = MB[v];
  if (w != XMB_NONE && x_mb_kind(w) == X_PULL) {
    const unsigned ts = x_mb_idx(w) * 32u + 24u;
    if (ts < T) {
      const uint32_t c = x_mb_code(w);
      const unsigned d = x_d2(x, y, z, c);
      if (d < d0 && (d < s.d || (d == s.d && ts < s.ts))) { s.d = d; s.c = c; s.ts = ts; }
    }
  }
  return s;
}


// ---- per-voxel offer summaries ------------------------------------------------------------------------------------
// Every voxel that some element of the generation can write (its 24 neighbours and itself) is a TARGET.  Once per round a
// target's <= 25 offers are gathered into a summary {first = earliest timestamp of an offer that beats the snapshot,
// best = lexicographic minimum (distance, timestamp) over all offers}.  A reader at time T then gets the snapshot if
// T <= first, `best` if T > best.ts, and has to gather the offers itself only in between (about 1 % of the queries in the
// CPU model, oracle/exact_model.c); non-targets have no writer and read as the snapshot.
__device__ __forceinline__ uint4 x_summarize(const FbGeom &g, const uint32_t *cobs, const unsigned long long *MB, int x, int y, int z) {
  const XState f = x_state(g, cobs, MB, x, y, z, XNONE);
  unsigned first = XNONE;
  const long long v = fb_ii(g, x, y, z);
  const uint32_t c0 = cobs[v] & FB_CODE_MASK;
  if (c0 != FB_UNKNOWN && fb_in_range(g, x, y, z)) {
    const unsigned d0 = x_dist_of(x, y, z, c0);
#pragma unroll 8
    for (int k = 0; k < 24; ++k) {
      const int qx = x - x_dirs[k][0], qy = y - x_dirs[k][1], qz = z - x_dirs[k][2];
      if (!fb_in_grid(g, qx, qy, qz)) continue;
      const unsigned long long w = MB[fb_ii(g, qx, qy, qz)];
      if (w == XMB_NONE || x_mb_kind(w) != X_PUSH) continue;
      const unsigned ts = x_mb_idx(w) * 32u + (unsigned)k;
      if (ts < first && x_d2(x, y, z, x_mb_code(w)) < d0) first = ts;
    }
    const unsigned long long w = MB[v];
    if (w != XMB_NONE && x_mb_kind(w) == X_PULL) {
      const unsigned ts = x_mb_idx(w) * 32u + 24u;
      if (ts < first && x_d2(x, y, z, x_mb_code(w)) < d0) first = ts;
    }
  }
  return make_uint4(first, f.d, f.ts, f.c);
}
__device__ __forceinline__ XState x_state_sum(const FbGeom &g, const uint32_t *cobs, const unsigned long long *MB, const uint4 *SUM,
                                              const uint32_t *SUMg, unsigned gen, int x, int y, int z, unsigned T) {
  const long long v = fb_ii(g, x, y, z);
  XState s;
  bool snapshot = SUMg[v] != gen;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (!snapshot) { u = SUM[v]; snapshot = u.x == XNONE || T <= u.x; }
  if (snapshot) { s.c = cobs[v] & FB_CODE_MASK; s.d = x_dist_of(x, y, z, s.c); s.ts = XNONE; return s; }
  if (T > u.z) { s.d = u.y; s.ts = u.z; s.c = u.w; return s; }
  return x_state(g, cobs, MB, x, y, z, T);                    // first < T <= best.ts: gather
}
// Targets of the generation (deduplicated through SUMg).
__global__ void k_x_targets(FbGeom g, const uint32_t *E, unsigned n, uint32_t *SUMg, unsigned gen, uint32_t *targets, unsigned *ntargets) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z; x_coords(g, E[i], x, y, z);
  for (int k = 0; k < 25; ++k) {
    const int nx = k < 24 ? x + x_dirs[k][0] : x, ny = k < 24 ? y + x_dirs[k][1] : y, nz = k < 24 ? z + x_dirs[k][2] : z;
    if (!fb_in_range(g, nx, ny, nz) || !fb_in_grid(g, nx, ny, nz)) continue;
    const long long v = fb_ii(g, nx, ny, nz);
    if (SUMg[v] != gen && atomicExch(&SUMg[v], gen) != gen) targets[atomicAdd(ntargets, 1u)] = (uint32_t)v;
  }
}
__global__ void k_x_sum(FbGeom g, const uint32_t *targets, unsigned nt, const uint32_t *cobs, const unsigned long long *MB, uint4 *SUM,
                        const uint32_t *tdirty, unsigned stamp, int first_round) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  int x, y, z; x_coords(g, targets[i], x, y, z);
  if (!first_round && tdirty[((x >> 3) * g.ty + (y >> 3)) * g.tz + (z >> 3)] != stamp) return;
  SUM[targets[i]] = x_summarize(g, cobs, MB, x, y, z);
}

// ------------------------------------------------------------------ occupancy (ordered)
__global__ void k_x_gather_keys(const uint32_t *vox, unsigned n, const unsigned long long *tkey, unsigned long long *keys) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = tkey[vox[i]];
}

// ESDFMap::UpdateOccupancy (ESDFMap.cpp:235-271) over the queue in first-observation order; flags mark insert / delete pushes.
__global__ void k_x_integrate(FbGeom g, const uint32_t *vox, unsigned n, unsigned long long *cnt, double *occ, uint32_t *cobs,
                              uint32_t *occbits, unsigned long long *tkey, uint8_t *f_ins, uint8_t *f_del, int global_map, double l_hit,
                              double l_miss, double l_min, double l_max, double l_occ) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ii = vox[i];
  uint8_t pi = 0, pd = 0;
  const unsigned long long c = cnt[ii];
  const long long hit = (long long)(c >> 32), tot = (long long)(c & 0xffffffffull);
  cnt[ii] = 0ull;
  tkey[ii] = ~0ull;
  const double upd = (hit >= tot - hit) ? l_hit : l_miss;
  if (cobs[ii] == FB_UNKNOWN) cobs[ii] = FB_INF;
  double o = occ[ii];
  const bool was = o > l_occ;
  const bool skip = (upd >= 0 && o >= l_max) || (upd <= 0 && o <= l_min);
  if (!skip) {
    if (!global_map) { int x, y, z; x_coords(g, ii, x, y, z); if (!fb_in_last_range(g, x, y, z)) { o = 0; cobs[ii] = FB_INF; } }
    double s = o + upd;
    s = s > l_min ? s : l_min;
    s = s < l_max ? s : l_max;
    occ[ii] = s;
    const bool now = s > l_occ;
    if (now && !was) { pi = 1; atomicOr(&occbits[ii >> 5], 1u << (ii & 31)); }
    else if (!now && was) { pd = 1; atomicAnd(&occbits[ii >> 5], ~(1u << (ii & 31))); }
  }
  f_ins[i] = pi; f_del[i] = pd;
}

// ------------------------------------------------------------------ E1: insert seeds
__global__ void k_x_flag_exist(const uint32_t *list, unsigned n, const double *occ, double l_occ, uint8_t *flags, int want) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = ((occ[list[i]] > l_occ) ? 1 : 0) == want;
}
__global__ void k_x_apply_seed(FbGeom g, const uint32_t *E, unsigned n, uint32_t *cobs, unsigned long long *LS, unsigned long long t0) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ii = E[i];
  int x, y, z; x_coords(g, ii, x, y, z);
  cobs[ii] = fb_pack(x, y, z);                                 // closest_obstacle_ = self, distance_ = 0 (:286-287)
  LS[ii] = t0 + i;                                             // InsertIntoList(idx, idx) (:288)
}

// ------------------------------------------------------------------ E2: delete
__global__ void k_x_del_minpos(const uint32_t *del, unsigned n, const double *occ, double l_occ, uint32_t *scratch) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(occ[del[i]] > l_occ)) atomicMin(&scratch[del[i]], i);   // `if (!Exist(idx))` (:297); first occurrence wins
}
__global__ void k_x_del_flag(const uint32_t *del, unsigned n, const double *occ, double l_occ, const uint32_t *scratch, uint8_t *flags) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (!(occ[del[i]] > l_occ) && scratch[del[i]] == i) ? 1 : 0;
}
__global__ void k_x_del_rank(const uint32_t *sel, unsigned n, uint32_t *scratch) {   // sel = deleted obstacles in queue order
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scratch[sel[i]] = i;
}
// Dependants = voxels whose closest obstacle is a deleted one (the reference walks head_[idx] -> next_, :301).
__global__ void k_x_scan_deps(FbGeom g, const uint32_t *cobs, const uint32_t *rank, const unsigned long long *LS, unsigned long long *k1,
                              unsigned long long *k2, uint32_t *dv, unsigned *ndep, unsigned cap) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < g.ptotal; v += (long long)gridDim.x * blockDim.x) {
    const uint32_t c = cobs[v] & FB_CODE_MASK;
    bool dep = false;
    unsigned r = 0;
    if (c >= 2u) {
      int ox, oy, oz; fb_unpack(c, ox, oy, oz);
      r = rank[fb_ii(g, ox, oy, oz)];
      dep = r != XNONE;
    }
    const unsigned slot = fb_warp_append(ndep, dep);
    if (dep && slot < cap) { k1[slot] = r; k2[slot] = ~LS[v]; dv[slot] = (uint32_t)v; }   // ~LS: most recently linked first
  }
}
__global__ void k_x_iota(uint32_t *a, unsigned n) { const unsigned i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }
__global__ void k_x_gather64(const unsigned long long *src, const uint32_t *idx, unsigned n, unsigned long long *dst) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_x_gather32(const uint32_t *src, const uint32_t *idx, unsigned n, uint32_t *dst) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_x_set_ord(const uint32_t *deps, unsigned n, uint32_t *ord, uint32_t *nc) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { ord[deps[i]] = i; nc[i] = FB_INF; }
}
// One round of the re-seeding fixpoint: dependant i takes the closest obstacle of the FIRST neighbour in dirs_ order that
// has a valid one (:308-321); dependants processed earlier expose their new value, later ones their (deleted) old one.
__global__ void k_x_reseed(FbGeom g, const uint32_t *deps, unsigned n, const uint32_t *cobs, const uint32_t *ord, const uint32_t *occbits,
                           const uint32_t *nc_in, uint32_t *nc_out, unsigned *changed) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z; x_coords(g, deps[i], x, y, z);
  uint32_t res = FB_INF;
  for (int k = 0; k < 24; ++k) {
    const int nx = x + x_dirs[k][0], ny = y + x_dirs[k][1], nz = z + x_dirs[k][2];
    if (!fb_in_range(g, nx, ny, nz) || !fb_in_grid(g, nx, ny, nz)) continue;
    const long long nv = fb_ii(g, nx, ny, nz);
    const unsigned o = ord[nv];
    uint32_t c;
    if (o != XNONE) { if (o < i) c = nc_in[o]; else continue; }
    else c = cobs[nv] & FB_CODE_MASK;
    if (c >= 2u) {
      int ox, oy, oz; fb_unpack(c, ox, oy, oz);
      const long long oi = fb_ii(g, ox, oy, oz);
      if ((occbits[oi >> 5] >> (oi & 31)) & 1u) { res = c; break; }             // Exist(closest obstacle) (:312), then `break` (:319)
    }
  }
  nc_out[i] = res;
  if (res != nc_in[i]) *changed = 1u;
}
__global__ void k_x_apply_reseed(const uint32_t *deps, unsigned n, const uint32_t *nc, uint32_t *cobs, unsigned long long *LS,
                                 unsigned long long t0, uint8_t *flags) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cobs[deps[i]] = nc[i];
  LS[deps[i]] = t0 + i;                                        // InsertIntoList(new_obs_idx, obs_idx) (:333)
  flags[i] = nc[i] >= 2u;                                      // `if (distance < infinity_) update_queue_.push` (:329-331)
}
// ------------------------------------------------------------------ E3: relax, one FIFO generation at a time
// Initial guess of the behaviour fixpoint: every entry is live and pushes its snapshot code.
__global__ void k_x_init_beh(const uint32_t *E, unsigned n, const uint32_t *cobs, unsigned long long *MB) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) MB[E[i]] = x_mb(i, X_PUSH, cobs[E[i]]);
}
// One round of the behaviour fixpoint.  B is updated IN PLACE (elements evaluated later in the same round already see the
// new behaviour of earlier ones); an element is re-evaluated only in the first round of a generation or when an element
// within reach (<= 4 voxels: its own 8^3 tile or one of the 26 around it) changed its behaviour in the previous round.
// The loop ends with a round in which nothing changed, which reads a stable B: that state is the sequential execution.
__global__ void k_x_eval(FbGeom g, const uint32_t *E, unsigned n, const uint32_t *cobs, unsigned long long *MB, const uint4 *SUM,
                         const uint32_t *SUMg, unsigned gen, uint32_t *tdirty, unsigned stamp, int first_round, unsigned *changed) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = E[i];
  int x, y, z; x_coords(g, p, x, y, z);
  const int tx = x >> 3, ty = y >> 3, tz = z >> 3;
  if (!first_round && tdirty[(tx * g.ty + ty) * g.tz + tz] != stamp) return;
  const unsigned T0 = i * 32u;
  const XState s = x_state_sum(g, cobs, MB, SUM, SUMg, gen, x, y, z, T0);
  const uint32_t c0 = cobs[p] & FB_CODE_MASK;
  unsigned long long nb;
  if (s.d != x_dist_of(x, y, z, c0)) nb = x_mb(i, X_DEAD, 0);  // `xx.distance_ != distance_buffer_[idx]`: stale (:345)
  else {
    unsigned curd = s.d; uint32_t curc = s.c; bool ch = false;
    for (int k = 0; k < 24; ++k) {                             // pull phase (:349-367)
      const int nx = x + x_dirs[k][0], ny = y + x_dirs[k][1], nz = z + x_dirs[k][2];
      if (!fb_in_range(g, nx, ny, nz) || !fb_in_grid(g, nx, ny, nz)) continue;
      const XState sn = x_state_sum(g, cobs, MB, SUM, SUMg, gen, nx, ny, nz, T0);
      if (sn.c < 2u) continue;
      const unsigned t = x_d2(x, y, z, sn.c);
      if (curd > t) { curd = t; curc = sn.c; ch = true; }
    }
    nb = ch ? x_mb(i, X_PULL, curc) : x_mb(i, X_PUSH, s.c);
  }
  if (nb != MB[p]) {
    MB[p] = nb;
    *changed = 1u;
    for (int a = max(tx - 1, 0); a <= min(tx + 1, g.tx - 1); ++a)
      for (int b = max(ty - 1, 0); b <= min(ty + 1, g.ty - 1); ++b)
        for (int c = max(tz - 1, 0); c <= min(tz + 1, g.tz - 1); ++c) tdirty[(a * g.ty + b) * g.tz + c] = stamp + 1u;
  }
}
// The same evaluation with one WARP per element, for the rounds after the first: only the few elements in dirty tiles do
// any work there, so the round's duration is the latency of a single evaluation -- 25 state queries run on 25 lanes
// instead of one after the other.  The sequential pull loop "for k: if (dist > tmp) take" (:349-367) is the lexicographic
// minimum (tmp, k) over the neighbours that beat the own distance (warp reduction).
// Elements whose tile is dirty for this round -> work list (so that the evaluation kernel is not launched over millions
// of idle threads).
__global__ void k_x_collect(FbGeom g, const uint32_t *E, unsigned n, const uint32_t *tdirty, unsigned stamp, uint32_t *work, unsigned *nwork) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  bool on = false;
  if (i < n) { int x, y, z; x_coords(g, E[i], x, y, z); on = tdirty[((x >> 3) * g.ty + (y >> 3)) * g.tz + (z >> 3)] == stamp; }
  const unsigned slot = fb_warp_append(nwork, on);
  if (on) work[slot] = i;
}
__global__ void k_x_eval_warp(FbGeom g, const uint32_t *E, const uint32_t *work, const unsigned *nwork, const uint32_t *cobs, unsigned long long *MB,
                              const uint4 *SUM, const uint32_t *SUMg, unsigned gen, uint32_t *tdirty, unsigned stamp, unsigned *changed) {
  const unsigned lane = threadIdx.x & 31u, nw = *nwork;
  for (unsigned wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < nw; wi += (gridDim.x * blockDim.x) >> 5) {
  const unsigned i = work[wi];
  const uint32_t p = E[i];
  int x, y, z; x_coords(g, p, x, y, z);
  const int tx = x >> 3, ty = y >> 3, tz = z >> 3;
  const unsigned T0 = i * 32u;
  int qx = x, qy = y, qz = z;
  bool valid = lane == 24;
  if (lane < 24) { qx += x_dirs[lane][0]; qy += x_dirs[lane][1]; qz += x_dirs[lane][2]; valid = fb_in_range(g, qx, qy, qz) && fb_in_grid(g, qx, qy, qz); }
  XState st; st.d = 0xffffffffu; st.c = 0; st.ts = XNONE;
  if (valid) st = x_state_sum(g, cobs, MB, SUM, SUMg, gen, qx, qy, qz, T0);   // lanes 0..23: neighbour k at pop time; lane 24: the element itself
  const unsigned sd = __shfl_sync(0xffffffffu, st.d, 24);
  const uint32_t sc = __shfl_sync(0xffffffffu, st.c, 24);
  const uint32_t c0 = cobs[p] & FB_CODE_MASK;
  unsigned long long nb;
  if (sd != x_dist_of(x, y, z, c0)) nb = x_mb(i, X_DEAD, 0);
  else {
    unsigned long long key = ~0ull;
    if (lane < 24 && valid && st.c >= 2u) {
      const unsigned t = x_d2(x, y, z, st.c);
      if (t < sd) key = ((unsigned long long)t << 8) | lane;
    }
    unsigned long long best = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other < best ? other : best; }
    if (best == ~0ull) nb = x_mb(i, X_PUSH, sc);
    else nb = x_mb(i, X_PULL, __shfl_sync(0xffffffffu, st.c, (int)(best & 0xffu)));
  }
  if (lane == 0 && nb != MB[p]) {
    MB[p] = nb;
    *changed = 1u;
    for (int a = max(tx - 1, 0); a <= min(tx + 1, g.tx - 1); ++a)
      for (int b = max(ty - 1, 0); b <= min(ty + 1, g.ty - 1); ++b)
        for (int c = max(tz - 1, 0); c <= min(tz + 1, g.tz - 1); ++c) tdirty[(a * g.ty + b) * g.tz + c] = stamp + 1u;
  }
  }
}
// Final writes of the generation -> slots (timestamp order) of the next generation's queue.
__global__ void k_x_commit(FbGeom g, const uint32_t *E, unsigned n, const unsigned long long *MB, const uint4 *SUM, const uint32_t *SUMg, unsigned gen,
                           uint32_t *slotv, uint32_t *slotc, uint8_t *slotf, unsigned long long *expansions) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long b = XMB_NONE;
  if (i < n) b = MB[E[i]];
  const bool live = i < n && x_mb_kind(b) != X_DEAD;
  const unsigned nlive = __popc(__ballot_sync(0xffffffffu, live));
  if ((threadIdx.x & 31) == 0 && nlive) atomicAdd(expansions, (unsigned long long)nlive);   // `times++` (:347)
  if (!live) return;
  int x, y, z; x_coords(g, E[i], x, y, z);
  // The last accepted write to a voxel is its summary's `best`; the element/direction that made it owns the slot.
  if (x_mb_kind(b) == X_PUSH) {
    for (int k = 0; k < 24; ++k) {                             // push phase (:375-391)
      const int nx = x + x_dirs[k][0], ny = y + x_dirs[k][1], nz = z + x_dirs[k][2];
      if (!fb_in_range(g, nx, ny, nz) || !fb_in_grid(g, nx, ny, nz)) continue;
      const long long v = fb_ii(g, nx, ny, nz);
      const unsigned ts = i * 32u + (unsigned)k;
      if (SUMg[v] != gen) continue;
      const uint4 u = SUM[v];
      if (u.z == ts) { slotv[ts] = (uint32_t)v; slotc[ts] = u.w; slotf[ts] = 1; }
    }
  } else {
    const unsigned ts = i * 32u + 24u;
    const uint4 u = SUM[E[i]];
    if (SUMg[E[i]] == gen && u.z == ts) { slotv[ts] = E[i]; slotc[ts] = u.w; slotf[ts] = 1; }
  }
}
__global__ void k_x_clear_M(const uint32_t *E, unsigned n, unsigned long long *MB) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) MB[E[i]] = XMB_NONE;
}
__global__ void k_x_apply(const uint32_t *sel, unsigned n, const uint32_t *slotv, const uint32_t *slotc, uint32_t *cobs,
                          unsigned long long *LS, unsigned long long t0, uint32_t *Enext) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t s = sel[r], v = slotv[s];
  cobs[v] = slotc[s];
  LS[v] = t0 + s;                                              // every accepted write relinks the voxel at its list's front
  Enext[r] = v;
}
__global__ void k_x_fill32(uint32_t *a, size_t n, uint32_t val) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = val;
}
__global__ void k_x_fill64(unsigned long long *a, size_t n, unsigned long long val) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = val;
}

// ================================================================== host side
#define XCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { snprintf(X->err, sizeof(X->err), "%s: %s", #call, cudaGetErrorString(e__)); return e__; } } while (0)
static inline unsigned nblk(size_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

template <typename T>
static cudaError_t x_ensure(FbExact *X, T **p, size_t *cap, size_t need) {
  if (need <= *cap) return cudaSuccess;
  size_t nc = need + need / 2 + 4096;
  T *np = nullptr;
  XCK(cudaMalloc((void **)&np, nc * sizeof(T)));
  if (*p) cudaFree(*p);
  *p = np; *cap = nc;
  return cudaSuccess;
}
static cudaError_t x_tmp(FbExact *X, size_t bytes) {
  if (bytes <= X->cub_bytes) return cudaSuccess;
  if (X->cub_tmp) cudaFree(X->cub_tmp);
  X->cub_bytes = bytes + bytes / 2 + (1u << 20);
  XCK(cudaMalloc(&X->cub_tmp, X->cub_bytes));
  return cudaSuccess;
}
// ordered compaction: out[0..count) = in[i] for flags[i] != 0, order kept
static cudaError_t x_select(FbExact *X, const uint32_t *in, const uint8_t *flags, uint32_t *out, unsigned n, unsigned *count, cudaStream_t s) {
  *count = 0;
  if (n == 0) return cudaSuccess;
  size_t bytes = 0;
  XCK(cub::DeviceSelect::Flagged(nullptr, bytes, in, flags, out, X->d_count, (int)n, s));
  cudaError_t e = x_tmp(X, bytes); if (e) return e;
  XCK(cub::DeviceSelect::Flagged(X->cub_tmp, bytes, in, flags, out, X->d_count, (int)n, s));
  XCK(cudaMemcpyAsync(X->h_count, X->d_count, 4, cudaMemcpyDeviceToHost, s));
  XCK(cudaStreamSynchronize(s));
  *count = *X->h_count;
  return cudaSuccess;
}
static cudaError_t x_sort_pairs(FbExact *X, const unsigned long long *kin, unsigned long long *kout, const uint32_t *vin, uint32_t *vout, unsigned n, cudaStream_t s) {
  size_t bytes = 0;
  XCK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, 64, s));
  cudaError_t e = x_tmp(X, bytes); if (e) return e;
  XCK(cub::DeviceRadixSort::SortPairs(X->cub_tmp, bytes, kin, kout, vin, vout, (int)n, 0, 64, s));
  return cudaSuccess;
}
static cudaError_t x_flag(FbExact *X, cudaStream_t s, unsigned *out) {   // read-and-clear the device "changed" flag
  XCK(cudaMemcpyAsync(X->h_count, X->d_flag, 4, cudaMemcpyDeviceToHost, s));
  XCK(cudaMemsetAsync(X->d_flag, 0, 4, s));
  XCK(cudaStreamSynchronize(s));
  *out = *X->h_count;
  return cudaSuccess;
}

cudaError_t fb_exact_init(FbExact *X, const FbGeom &g, cudaStream_t s) {
  memset(X, 0, sizeof(*X));
  const size_t P = (size_t)g.ptotal;
  XCK(cudaMalloc((void **)&X->MB, P * 8)); XCK(cudaMalloc((void **)&X->LS, P * 8)); XCK(cudaMalloc((void **)&X->tkey, P * 8));
  XCK(cudaMalloc((void **)&X->touched, P * 4));
  XCK(cudaMalloc((void **)&X->SUM, P * 16)); XCK(cudaMalloc((void **)&X->SUMg, P * 4)); XCK(cudaMemsetAsync(X->SUMg, 0, P * 4, s));
  XCK(cudaMalloc((void **)&X->tdirty, (size_t)g.ntiles * 4)); XCK(cudaMemsetAsync(X->tdirty, 0, (size_t)g.ntiles * 4, s));
  XCK(cudaMalloc((void **)&X->d_count, 16)); XCK(cudaMalloc((void **)&X->d_flag, 16));
  XCK(cudaMallocHost((void **)&X->h_count, 16));
  XCK(cudaMemsetAsync(X->d_count, 0, 16, s)); XCK(cudaMemsetAsync(X->d_flag, 0, 16, s));
  k_x_fill64<<<148 * 8, 256, 0, s>>>(X->MB, P, XMB_NONE);
  k_x_fill64<<<148 * 8, 256, 0, s>>>(X->tkey, P, ~0ull);
  XCK(cudaMemsetAsync(X->LS, 0, P * 8, s));
  X->tclock = 1; X->key_base = 0; X->eval_clock = 1; X->gen_id = 0;
  return cudaGetLastError();
}
void fb_exact_free(FbExact *X) {
  void *p[] = {X->MB, X->LS, X->tkey, X->touched, X->tdirty, X->SUM, X->SUMg, X->targets, X->work, X->d_count, X->d_flag, X->E[0], X->E[1], X->slotv, X->slotc, X->slotf, X->sel,
               X->k1, X->k2, X->k1b, X->k2b, X->dv, X->idx[0], X->idx[1], X->deps, X->nc[0], X->nc[1], X->flags, X->flags2, X->cub_tmp};
  for (void *q : p) if (q) cudaFree(q);
  if (X->h_count) cudaFreeHost(X->h_count);
  memset(X, 0, sizeof(*X));
}

// UpdateOccupancy in first-observation order.  touched[0..n) = voxels with pending observations (any order).
cudaError_t fb_exact_update_occupancy(FbExact *X, const FbGeom &g, unsigned n, unsigned long long *cnt, double *occ, uint32_t *cobs, uint32_t *occbits,
                                      uint32_t **ins, size_t *cap_ins, unsigned *n_ins, uint32_t **del, size_t *cap_del, unsigned *n_del,
                                      int global_map, const double L[5], cudaStream_t s, int *launches) {
  if (n == 0) return cudaSuccess;
  cudaError_t e;
  if ((e = x_ensure(X, &X->k1, &X->cap_k1, n))) return e;
  if ((e = x_ensure(X, &X->k1b, &X->cap_k1b, n))) return e;
  if ((e = x_ensure(X, &X->deps, &X->cap_deps, n))) return e;
  if ((e = x_ensure(X, &X->flags, &X->cap_flags, n))) return e;
  if ((e = x_ensure(X, &X->flags2, &X->cap_flags2, n))) return e;
  k_x_gather_keys<<<nblk(n), 256, 0, s>>>(X->touched, n, X->tkey, X->k1);
  if ((e = x_sort_pairs(X, X->k1, X->k1b, X->touched, X->deps, n, s))) return e;
  k_x_integrate<<<nblk(n), 256, 0, s>>>(g, X->deps, n, cnt, occ, cobs, occbits, X->tkey, X->flags, X->flags2, global_map, L[0], L[1], L[2], L[3], L[4]);
  *launches += 3;
  // append to the queues, order kept
  {   // grow the queues if needed (contents kept)
    for (int q = 0; q < 2; ++q) {
      uint32_t **lst = q ? del : ins; size_t *cap = q ? cap_del : cap_ins; const unsigned have = q ? *n_del : *n_ins;
      if ((size_t)have + n > *cap) {
        size_t nc = (size_t)have + n + ((size_t)have + n) / 2 + 4096;
        uint32_t *np = nullptr;
        XCK(cudaMalloc((void **)&np, nc * 4));
        if (*lst && have) XCK(cudaMemcpyAsync(np, *lst, (size_t)have * 4, cudaMemcpyDeviceToDevice, s));
        XCK(cudaStreamSynchronize(s));
        if (*lst) cudaFree(*lst);
        *lst = np; *cap = nc;
      }
    }
  }
  unsigned c = 0;
  if ((e = x_select(X, X->deps, X->flags, *ins + *n_ins, n, &c, s))) return e;
  *n_ins += c;
  if ((e = x_select(X, X->deps, X->flags2, *del + *n_del, n, &c, s))) return e;
  *n_del += c;
  return cudaSuccess;
}

cudaError_t fb_exact_update_esdf(FbExact *X, const FbGeom &g, uint32_t *cobs, uint32_t *scratch, const double *occ, const uint32_t *occbits, double l_occ,
                                 const uint32_t *ins, unsigned n_ins, const uint32_t *del, unsigned n_del, cudaStream_t s, FbExactStats *st, int *launches) {
  cudaError_t e;
  const size_t P = (size_t)g.ptotal;
  memset(st, 0, sizeof(*st));
  if ((e = x_ensure(X, &X->flags, &X->cap_flags, (size_t)(n_ins > n_del ? n_ins : n_del) + 16))) return e;
  if ((e = x_ensure(X, &X->E[0], &X->cap_E[0], (size_t)n_ins + 16))) return e;
  // ---- E1: insert seeds in insert_queue_ order (:278-291)
  unsigned nE = 0;
  if (n_ins) {
    k_x_flag_exist<<<nblk(n_ins), 256, 0, s>>>(ins, n_ins, occ, l_occ, X->flags, 1);
    if ((e = x_select(X, ins, X->flags, X->E[0], n_ins, &nE, s))) return e;
    if (nE) k_x_apply_seed<<<nblk(nE), 256, 0, s>>>(g, X->E[0], nE, cobs, X->LS, X->tclock);
    X->tclock += nE;
    *launches += 2;
  }
  // ---- E2: deletes (:292-337)
  if (n_del) {
    if ((e = x_ensure(X, &X->sel, &X->cap_sel, (size_t)n_del + 16))) return e;
    k_x_fill32<<<148 * 8, 256, 0, s>>>(scratch, P, XNONE);
    k_x_del_minpos<<<nblk(n_del), 256, 0, s>>>(del, n_del, occ, l_occ, scratch);
    k_x_del_flag<<<nblk(n_del), 256, 0, s>>>(del, n_del, occ, l_occ, scratch, X->flags);
    unsigned nd = 0;
    if ((e = x_select(X, del, X->flags, X->sel, n_del, &nd, s))) return e;
    *launches += 3;
    if (nd) {
      k_x_fill32<<<148 * 8, 256, 0, s>>>(scratch, P, XNONE);
      k_x_del_rank<<<nblk(nd), 256, 0, s>>>(X->sel, nd, scratch);
      // dependants: the list is sized by a first counting attempt, then (rarely) re-run with more room
      unsigned ndep = 0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        size_t cap = X->cap_dv;
        XCK(cudaMemsetAsync(X->d_count, 0, 4, s));
        k_x_scan_deps<<<148 * 16, 256, 0, s>>>(g, cobs, scratch, X->LS, X->k1, X->k2, X->dv, X->d_count, (unsigned)((cap < X->cap_k1 ? cap : X->cap_k1) < X->cap_k2 ? (cap < X->cap_k1 ? cap : X->cap_k1) : X->cap_k2));
        XCK(cudaMemcpyAsync(X->h_count, X->d_count, 4, cudaMemcpyDeviceToHost, s));
        XCK(cudaStreamSynchronize(s));
        ndep = *X->h_count;
        *launches += 1;
        if (ndep <= cap && ndep <= X->cap_k1 && ndep <= X->cap_k2) break;
        if ((e = x_ensure(X, &X->dv, &X->cap_dv, ndep))) return e;
        if ((e = x_ensure(X, &X->k1, &X->cap_k1, ndep))) return e;
        if ((e = x_ensure(X, &X->k2, &X->cap_k2, ndep))) return e;
      }
      st->dependants = ndep;
      if (ndep) {
        if ((e = x_ensure(X, &X->k1b, &X->cap_k1b, ndep))) return e;
        if ((e = x_ensure(X, &X->k2b, &X->cap_k2b, ndep))) return e;
        if ((e = x_ensure(X, &X->idx[0], &X->cap_idx[0], ndep))) return e;
        if ((e = x_ensure(X, &X->idx[1], &X->cap_idx[1], ndep))) return e;
        if ((e = x_ensure(X, &X->deps, &X->cap_deps, ndep))) return e;
        if ((e = x_ensure(X, &X->nc[0], &X->cap_nc[0], ndep))) return e;
        if ((e = x_ensure(X, &X->nc[1], &X->cap_nc[1], ndep))) return e;
        if ((e = x_ensure(X, &X->flags, &X->cap_flags, ndep))) return e;
        // order = (obstacle's position in delete_queue_, link time descending): two stable radix sorts
        k_x_iota<<<nblk(ndep), 256, 0, s>>>(X->idx[0], ndep);
        if ((e = x_sort_pairs(X, X->k2, X->k2b, X->idx[0], X->idx[1], ndep, s))) return e;
        k_x_gather64<<<nblk(ndep), 256, 0, s>>>(X->k1, X->idx[1], ndep, X->k1b);
        if ((e = x_sort_pairs(X, X->k1b, X->k2b, X->idx[1], X->idx[0], ndep, s))) return e;
        k_x_gather32<<<nblk(ndep), 256, 0, s>>>(X->dv, X->idx[0], ndep, X->deps);
        k_x_fill32<<<148 * 8, 256, 0, s>>>(scratch, P, XNONE);
        k_x_set_ord<<<nblk(ndep), 256, 0, s>>>(X->deps, ndep, scratch, X->nc[0]);
        *launches += 8;
        int cur = 0;
        for (int it = 0; it < 100000; ++it) {
          k_x_reseed<<<nblk(ndep), 256, 0, s>>>(g, X->deps, ndep, cobs, scratch, occbits, X->nc[cur], X->nc[cur ^ 1], X->d_flag);
          *launches += 1;
          cur ^= 1;
          unsigned ch = 0;
          if ((e = x_flag(X, s, &ch))) return e;
          st->reseed_rounds++;
          if (!ch) break;
        }
        k_x_apply_reseed<<<nblk(ndep), 256, 0, s>>>(X->deps, ndep, X->nc[cur], cobs, X->LS, X->tclock, X->flags);
        X->tclock += ndep;
        if ((e = x_ensure(X, &X->E[1], &X->cap_E[1], (size_t)nE + ndep + 16))) return e;   // E[0] may be too small: rebuild in E[1]
        if (nE) XCK(cudaMemcpyAsync(X->E[1], X->E[0], (size_t)nE * 4, cudaMemcpyDeviceToDevice, s));
        unsigned nr = 0;
        if ((e = x_select(X, X->deps, X->flags, X->E[1] + nE, ndep, &nr, s))) return e;
        // keep the generation-0 list in E[0]
        if ((e = x_ensure(X, &X->E[0], &X->cap_E[0], (size_t)nE + nr + 16))) return e;
        XCK(cudaMemcpyAsync(X->E[0], X->E[1], (size_t)(nE + nr) * 4, cudaMemcpyDeviceToDevice, s));
        nE += nr;
        *launches += 3;
      }
    }
  }
  // ---- E3: relax (:338-392)
  static const bool xdbg = getenv("FIESTA_DEBUG_X") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  double t_eval = 0, t_commit = 0, t_select = 0, t_apply = 0, t_k[5] = {0, 0, 0, 0, 0}; auto t_start = now();
  auto lap = [&](int k, std::chrono::steady_clock::time_point &a) { if (xdbg) { cudaStreamSynchronize(s); auto b = now(); t_k[k] += ms(a, b); a = b; } };
  int cur = 0;
  XCK(cudaMemsetAsync(X->d_count + 2, 0, 8, s));               // expansions counter (u64 at d_count[2..3])
  while (nE) {
    st->generations++;
    const size_t nslots = (size_t)nE * 32;
    if (nE >= (1u << 27)) { snprintf(X->err, sizeof(X->err), "exact mode: generation with more than 2^27 entries"); return cudaErrorInvalidValue; }
    if ((e = x_ensure(X, &X->slotv, &X->cap_slotv, nslots))) return e;
    if ((e = x_ensure(X, &X->slotc, &X->cap_slotc, nslots))) return e;
    if ((e = x_ensure(X, &X->slotf, &X->cap_slotf, nslots))) return e;
    if ((e = x_ensure(X, &X->sel, &X->cap_sel, nslots))) return e;
    auto t0 = now();
    k_x_init_beh<<<nblk(nE), 256, 0, s>>>(X->E[cur], nE, cobs, X->MB);
    // targets of this generation
    ++X->gen_id;
    if ((e = x_ensure(X, &X->targets, &X->cap_targets, (size_t)nE * 25))) return e;
    if ((e = x_ensure(X, &X->work, &X->cap_work, (size_t)nE))) return e;
    XCK(cudaMemsetAsync(X->d_count, 0, 4, s));
    auto tl = now();
    k_x_targets<<<nblk(nE), 256, 0, s>>>(g, X->E[cur], nE, X->SUMg, X->gen_id, X->targets, X->d_count);
    XCK(cudaMemcpyAsync(X->h_count, X->d_count, 4, cudaMemcpyDeviceToHost, s));
    XCK(cudaStreamSynchronize(s));
    const unsigned nT = *X->h_count;
    *launches += 2;
    lap(0, tl);
    // Rounds are launched four at a time between host checks: a round after convergence finds no dirty tile and costs
    // next to nothing, and a batch that changed nothing proves that the last state survived a full round.
    for (int it = 0; it < 100000; it += 4) {
      for (int q = 0; q < 4; ++q) {
        ++X->eval_clock;
        const int first = it + q == 0;
        k_x_sum<<<nblk(nT), 256, 0, s>>>(g, X->targets, nT, cobs, X->MB, X->SUM, X->tdirty, X->eval_clock, first);
        lap(1, tl);
        if (first) { k_x_eval<<<nblk(nE, 128), 128, 0, s>>>(g, X->E[cur], nE, cobs, X->MB, X->SUM, X->SUMg, X->gen_id, X->tdirty, X->eval_clock, 1, X->d_flag); lap(2, tl); }
        else {
          XCK(cudaMemsetAsync(X->d_count + 1, 0, 4, s));
          k_x_collect<<<nblk(nE), 256, 0, s>>>(g, X->E[cur], nE, X->tdirty, X->eval_clock, X->work, X->d_count + 1);
          lap(3, tl);
          k_x_eval_warp<<<148 * 4, 256, 0, s>>>(g, X->E[cur], X->work, X->d_count + 1, cobs, X->MB, X->SUM, X->SUMg, X->gen_id, X->tdirty, X->eval_clock, X->d_flag);
          lap(4, tl);
        }
      }
      *launches += 10;
      unsigned ch = 0;
      if ((e = x_flag(X, s, &ch))) return e;
      st->eval_rounds += 4;
      if (!ch) break;
    }
    auto t1 = now();
    XCK(cudaMemsetAsync(X->slotf, 0, nslots, s));
    k_x_commit<<<nblk(nE, 128), 128, 0, s>>>(g, X->E[cur], nE, X->MB, X->SUM, X->SUMg, X->gen_id, X->slotv, X->slotc, X->slotf, (unsigned long long *)(X->d_count + 2));
    k_x_clear_M<<<nblk(nE), 256, 0, s>>>(X->E[cur], nE, X->MB);
    if (xdbg) cudaStreamSynchronize(s);
    auto t2 = now();
    unsigned n2 = 0;
    {
      size_t bytes = 0;
      cub::CountingInputIterator<uint32_t> it0(0);
      XCK(cub::DeviceSelect::Flagged(nullptr, bytes, it0, X->slotf, X->sel, X->d_count, (int)nslots, s));
      if ((e = x_tmp(X, bytes))) return e;
      XCK(cub::DeviceSelect::Flagged(X->cub_tmp, bytes, it0, X->slotf, X->sel, X->d_count, (int)nslots, s));
      XCK(cudaMemcpyAsync(X->h_count, X->d_count, 4, cudaMemcpyDeviceToHost, s));
      XCK(cudaStreamSynchronize(s));
      n2 = *X->h_count;
    }
    auto t3 = now();
    if ((e = x_ensure(X, &X->E[cur ^ 1], &X->cap_E[cur ^ 1], (size_t)n2 + 16))) return e;
    if (n2) k_x_apply<<<nblk(n2), 256, 0, s>>>(X->sel, n2, X->slotv, X->slotc, cobs, X->LS, X->tclock, X->E[cur ^ 1]);
    *launches += 5;
    if (xdbg) cudaStreamSynchronize(s);
    auto t4 = now();
    t_eval += ms(t0, t1); t_commit += ms(t1, t2); t_select += ms(t2, t3); t_apply += ms(t3, t4);
    X->tclock += nslots + 1;
    st->voxels_changed += n2;
    cur ^= 1;
    nE = n2;
  }
  XCK(cudaMemcpyAsync(X->h_count, X->d_count + 2, 8, cudaMemcpyDeviceToHost, s));
  XCK(cudaStreamSynchronize(s));
  st->expansions = *(unsigned long long *)X->h_count;
  if (xdbg) fprintf(stderr, "[x] gens %u rounds %u deps %u | relax %.1f ms: eval %.1f (targets %.1f sum %.1f eval1 %.1f collect %.1f evalw %.1f) commit %.1f select %.1f apply %.1f\n", st->generations, st->eval_rounds, st->dependants, ms(t_start, now()), t_eval, t_k[0], t_k[1], t_k[2], t_k[3], t_k[4], t_commit, t_select, t_apply);
  return cudaSuccess;
}
