// fiesta_b200 -- per-frame ray casting into the hit/miss counters.
//
// Replaces Fiesta::RaycastMultithread + RaycastProcess (/root/reference/include/Fiesta.h:194-303), Raycast()
// (/root/reference/src/raycast.cpp:56-158) and the counter half of ESDFMap::SetOccupancy
// (/root/reference/src/ESDFMap.cpp:401-437) with results identical to the reference's SERIAL mode
// (ray_cast_num_thread = 0, the only deterministic one).
//
// The serial loop is order dependent: a ray's back-to-front free-space marking stops at the first voxel an EARLIER
// ray already stamped this frame (set_free_, Fiesta.h:265-273), and a ray is skipped when an earlier point ended in the
// same voxel (set_occ_, Fiesta.h:227-230).  The device pipeline reproduces that exactly:
//   k_ray_endpoints : per point: transform, length gating, endpoint SetOccupancy (one 64-bit atomicAdd on
//                     {hit:32|total:32}), and "lowest point index wins" ownership of the endpoint voxel (atomicMax on a
//                     tagged stamp) == the set_occ_ dedupe.
//   k_ray_trace     : per surviving ray (one thread each; the DDA's tMax += tDelta recurrence is sequential fp64 and
//                     must round exactly like the reference): Amanatides-Woo traversal with the reference's quirks
//                     (direction from integer voxel deltas, tie order z>y>x, corner-based distance cut), run twice:
//                     count, then write the voxel list back-to-front, transposed ([step][ray]) so a warp reads
//                     contiguous memory.
//   k_ray_resolve   : persistent cooperative kernel.  Round r: every ray walks its list from the far end and stops at the
//                     first voxel whose round r-1 stamp belongs to a lower ray index, stamping what it passes with
//                     atomicMax(tag|~ray) into the round-r array.  Rays with lower indices than everything they meet are
//                     exact after round 1, and by induction on the ray index the iteration reaches the unique fixpoint =
//                     the serial result; it stops when no ray's reach changed.  A last walk adds the counts.
// All fp64 arithmetic is written in the reference's operation order and the library is compiled with -fmad=false.
#include <cooperative_groups.h>
#include "fb_common.cuh"

namespace cg = cooperative_groups;

__device__ __forceinline__ double fb_norm3(double ax, double ay, double az, const double *b) {
  const double x = ax - b[0], y = ay - b[1], z = az - b[2];
  return sqrt((x * x + y * y) + z * z);                       // Eigen squaredNorm order, then sqrt
}

// ESDFMap::SetOccupancy(Vector3d,int) index part (ESDFMap.cpp:401-421): returns false when pos is outside the map.
// ref_idx = the reference's linear index (may alias when pos lies exactly on the map's upper face).
__device__ __forceinline__ bool fb_pos_to_vox(const FbGeom &g, double px, double py, double pz, int &vx, int &vy, int &vz) {
  if (px < g.min_range[0] || py < g.min_range[1] || pz < g.min_range[2]) return false;   // PosInMap, ESDFMap.cpp:46-61
  if (px > g.max_range[0] || py > g.max_range[1] || pz > g.max_range[2]) return false;
  vx = (int)floor((px - g.origin[0]) / g.res);                                            // Pos2Vox, ESDFMap.cpp:74-77
  vy = (int)floor((py - g.origin[1]) / g.res);
  vz = (int)floor((pz - g.origin[2]) / g.res);
  return true;
}

// Resolve the voxel a SetOccupancy call addresses.  Returns the list/stamp element: class + device index.
// The reference indexes set_free_/set_occ_ with the LINEAR index, which aliases for an out-of-grid coordinate; aliasing
// is reproduced while the linear index stays inside the array and treated as "not in map" beyond it.
__device__ __forceinline__ bool fb_resolve_vox(const FbGeom &g, int vx, int vy, int vz, long long &ii, bool &in_range) {
  in_range = fb_in_range(g, vx, vy, vz);
  if (fb_in_grid(g, vx, vy, vz)) { ii = fb_ii(g, vx, vy, vz); return true; }
  const long long ri = (long long)vx * g.gyz + (long long)vy * g.gz + vz;                 // Vox2Idx, ESDFMap.cpp:91
  if (ri < 0 || ri >= g.total) return false;
  const int ax = (int)(ri / g.gyz), ay = (int)(ri % g.gyz / g.gz), az = (int)(ri % g.gz); // Idx2Vox, ESDFMap.cpp:113-115
  ii = fb_ii(g, ax, ay, az);
  in_range = false;                                           // VoxInRange is evaluated on the coordinates, not the alias
  return true;
}

// Point -> world endpoint with the reference's gating (Fiesta.h:200-215).  Returns 0 = skipped, 1 = hit, 2 = clipped miss.
__device__ __forceinline__ int fb_endpoint(const FbRayArgs &a, long long i, double &px, double &py, double &pz) {
  const float fx = a.xyz[3 * i], fy = a.xyz[3 * i + 1], fz = a.xyz[3 * i + 2];
  if (isnan(fx) || isnan(fy) || isnan(fz)) return 0;
  const double x = (double)fx, y = (double)fy, z = (double)fz;
  double w[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = ((a.T[4 * r] * x + a.T[4 * r + 1] * y) + a.T[4 * r + 2] * z) + a.T[4 * r + 3] * 1.0;
  px = w[0] / w[3]; py = w[1] / w[3]; pz = w[2] / w[3];
  const double len = fb_norm3(px, py, pz, a.org);
  if (len < a.min_len) return 0;
  if (len > a.max_len) {
    px = (px - a.org[0]) / len * a.max_len + a.org[0];
    py = (py - a.org[1]) / len * a.max_len + a.org[1];
    pz = (pz - a.org[2]) / len * a.max_len + a.org[2];
    return 2;
  }
  return 1;
}

__device__ __forceinline__ void fb_count(const FbRayArgs &a, long long ii, unsigned occ) {
  const unsigned long long old = atomicAdd(&a.cnt[ii], ((unsigned long long)occ << 32) | 1ull);
  const bool first = (unsigned)(old & 0xffffffffull) == 0u;   // num_miss_ == 1 -> occupancy_queue_.push (ESDFMap.cpp:426-435)
  const unsigned slot = fb_warp_append(&a.ctr->n_touched, first);
  if (first && slot < a.touched_cap) a.touched[slot] = (uint32_t)ii;
}

// ---------------------------------------------------------------- endpoints
__global__ void k_ray_endpoints(FbGeom g, FbRayArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double px, py, pz;
  const int kind = fb_endpoint(a, i, px, py, pz);
  int len = -1;                                               // -1: no ray (skipped point)
  if (kind) {
    len = 0;
    int vx, vy, vz;
    if (fb_pos_to_vox(g, px, py, pz, vx, vy, vz)) {
      long long ii; bool in_range;
      if (fb_resolve_vox(g, vx, vy, vz, ii, in_range)) {
        if (in_range) fb_count(a, ii, kind == 1 ? 1u : 0u);
        // set_occ_ ownership: lowest point index wins (Fiesta.h:227-230)
        atomicMax(&a.stamp[1][ii], (a.tag_base << FB_RAY_BITS) | (FB_RAY_MASK - (unsigned)i));
      }
    }
  }
  a.ray_len[i] = len;
}

// ---------------------------------------------------------------- DDA
struct FbDda {
  int c[3], e[3], step[3];
  double tmax[3], tdelta[3], maxd;
};

__device__ __forceinline__ double fb_intbound(double s, double ds) {   // raycast.cpp:10-23
  if (ds < 0) { s = -s; ds = -ds; }
  s = fmod(fmod(s, 1.0) + 1.0, 1.0);
  return (1 - s) / ds;
}

__device__ __forceinline__ bool fb_dda_init(FbDda &d, const double *start, double ex, double ey, double ez) {
  const double end[3] = {ex, ey, ez};
  double dd[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { d.c[k] = (int)floor(start[k]); d.e[k] = (int)floor(end[k]); dd[k] = end[k] - start[k]; }
  d.maxd = (dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double delta = (double)(d.e[k] - d.c[k]);
    d.step[k] = (delta == 0.0) ? 0 : (delta < 0.0 ? -1 : 1);
    d.tmax[k] = fb_intbound(start[k], delta);
    d.tdelta[k] = ((double)d.step[k]) / delta;
  }
  return d.step[0] != 0 || d.step[1] != 0 || d.step[2] != 0;
}

// Walks the reference loop (raycast.cpp:116-157).  emit(x,y,z) is called for every voxel Raycast() would push.
// Returns the number of pushed voxels, or -1 (reference throws, > 1500 voxels) or -2 (reference never returns: it has
// overshot the end voxel and left the box for good -- see oracle/esdf_oracle.c for the same guard).
template <typename F>
__device__ __forceinline__ int fb_dda_walk(FbDda d, const FbRayArgs &a, F emit) {
  int n = 0;
  for (;;) {
    if (d.c[0] >= a.bmin[0] && d.c[0] < a.bmax[0] && d.c[1] >= a.bmin[1] && d.c[1] < a.bmax[1] && d.c[2] >= a.bmin[2] && d.c[2] < a.bmax[2]) {
      emit(d.c[0], d.c[1], d.c[2], n);
      ++n;
      const double x = d.c[0] - a.start[0], y = d.c[1] - a.start[1], z = d.c[2] - a.start[2];
      if ((x * x + y * y) + z * z > d.maxd) return n;
      if (n > 1500) return -1;
    }
    if (d.c[0] == d.e[0] && d.c[1] == d.e[1] && d.c[2] == d.e[2]) break;
    bool overshot = false, gone = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if ((d.step[k] > 0 && d.c[k] > d.e[k]) || (d.step[k] < 0 && d.c[k] < d.e[k]) || (d.step[k] == 0 && d.c[k] != d.e[k])) overshot = true;
      if ((d.step[k] >= 0 && !(d.c[k] < a.bmax[k])) || (d.step[k] <= 0 && !(d.c[k] >= a.bmin[k]))) gone = true;
    }
    if (overshot && gone) return -2;
    if (d.tmax[0] < d.tmax[1]) {
      if (d.tmax[0] < d.tmax[2]) { d.c[0] += d.step[0]; d.tmax[0] += d.tdelta[0]; } else { d.c[2] += d.step[2]; d.tmax[2] += d.tdelta[2]; }
    } else {
      if (d.tmax[1] < d.tmax[2]) { d.c[1] += d.step[1]; d.tmax[1] += d.tdelta[1]; } else { d.c[2] += d.step[2]; d.tmax[2] += d.tdelta[2]; }
    }
  }
  return n;
}

// pass 0: count + ownership check; pass 1: write the list.
template <int PASS>
__global__ void k_ray_trace(FbGeom g, FbRayArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int len = a.ray_len[i];
  if (PASS == 0 ? (len < 0) : (len <= 0)) return;
  double px, py, pz;
  const int kind = fb_endpoint(a, i, px, py, pz);
  (void)kind;
  if (PASS == 0) {
    int vx, vy, vz;
    if (fb_pos_to_vox(g, px, py, pz, vx, vy, vz)) {
      long long ii; bool in_range;
      if (fb_resolve_vox(g, vx, vy, vz, ii, in_range)) {
        const unsigned own = __ldcg(&a.stamp[1][ii]);
        if ((own & FB_RAY_MASK) != (FB_RAY_MASK - (unsigned)i)) { a.ray_len[i] = -1; return; }   // an earlier point owns this voxel
      }
    }
  }
  FbDda d;
  const bool moving = fb_dda_init(d, a.start, px / g.res, py / g.res, pz / g.res);
  if (PASS == 0) {
    int n = 0;
    if (moving) n = fb_dda_walk(d, a, [](int, int, int, int) {});
    atomicAdd(&a.ctr->rays_cast, 1u);
    if (n < 0) { atomicAdd(&a.ctr->rays_dropped, 1u); if (n == -1) atomicExch(&a.ctr->ray_error, 1u); n = 0; }
    int L = n > 0 ? n - 1 : 0;                                // `for (i = output.size() - 2; ...)`: the last voxel is skipped
    if (L > a.cap) { L = 0; atomicExch(&a.ctr->ray_error, 2u); atomicAdd(&a.ctr->rays_dropped, 1u); }
    a.ray_len[i] = L;
    a.ray_reach[i] = -1;
    if (L) atomicAdd(&a.ctr->ray_voxels, (unsigned long long)L);
  } else {
    const int L = len;
    fb_dda_walk(d, a, [&](int x, int y, int z, int j) {
      if (j >= L) return;                                     // the last pushed voxel is never visited (Fiesta.h:239)
      const double cx = (x + 0.5) * g.res, cy = (y + 0.5) * g.res, cz = (z + 0.5) * g.res;   // Fiesta.h:240
      const double l = fb_norm3(cx, cy, cz, a.org);
      unsigned e;
      if (l < a.min_len) e = FB_CLS_STOP << 30;
      else if (l > a.max_len) e = FB_CLS_SKIP << 30;
      else {
        int vx, vy, vz; long long ii; bool in_range;
        if (fb_pos_to_vox(g, cx, cy, cz, vx, vy, vz) && fb_resolve_vox(g, vx, vy, vz, ii, in_range))
          e = ((in_range ? FB_CLS_COUNT : FB_CLS_STAMP) << 30) | (unsigned)ii;
        else e = FB_CLS_SKIP << 30;                           // SetOccupancy returned -10000 (Fiesta.h:253)
      }
      a.ray_list[(long long)(L - 1 - j) * a.n + i] = e;
    });
  }
}

// ---------------------------------------------------------------- stamp resolution + counting
__global__ void __launch_bounds__(256) k_ray_resolve(FbGeom g, FbRayArgs a) {
  cg::grid_group grid = cg::this_grid();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  unsigned round = 1;
  for (;; ++round) {
    const unsigned tag = a.tag_base + round, oldtag = tag - 1u;
    const uint32_t *so = a.stamp[(round - 1u) & 1u];
    uint32_t *sn = a.stamp[round & 1u];
    bool changed = false;
    for (long long i = tid; i < a.n; i += nthreads) {
      const int L = a.ray_len[i];
      if (L <= 0) continue;
      const unsigned me = (unsigned)i;
      int t = 0;
      for (; t < L; ++t) {
        const unsigned e = __ldcg(&a.ray_list[(long long)t * a.n + i]);
        const unsigned cls = e >> 30;
        if (cls == FB_CLS_STOP) break;
        if (cls == FB_CLS_SKIP) continue;
        const unsigned ii = e & FB_LIST_IDX_MASK;
        const unsigned o = __ldcg(&so[ii]);
        if ((o >> FB_RAY_BITS) == oldtag && (FB_RAY_MASK - (o & FB_RAY_MASK)) < me) break;   // set_free_[idx] == tt by an earlier ray
        atomicMax(&sn[ii], (tag << FB_RAY_BITS) | (FB_RAY_MASK - me));
      }
      if (t != a.ray_reach[i]) { a.ray_reach[i] = t; changed = true; }
    }
    if (changed) atomicExch(&a.ctr->ray_flag[round % 3u], 1u);
    grid.sync();
    const unsigned any = __ldcg(&a.ctr->ray_flag[round % 3u]);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctr->ray_flag[(round + 2u) % 3u] = 0u;   // next used two barriers from now
    if (!any || round >= a.max_rounds) break;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.ctr->ray_rounds = round;
    if (round >= a.max_rounds) a.ctr->ray_error = 3u;
  }
  // final walk: `SetOccupancy(tmp, 0)` for every visited voxel, including the one that stops the walk (Fiesta.h:248-268)
  const unsigned tag = a.tag_base + round;
  const uint32_t *sf = a.stamp[round & 1u];
  for (long long i = tid; i < a.n; i += nthreads) {
    const int L = a.ray_len[i];
    const unsigned me = (unsigned)i;
    for (int t = 0; t < L; ++t) {
      const unsigned e = __ldcg(&a.ray_list[(long long)t * a.n + i]);
      const unsigned cls = e >> 30;
      if (cls == FB_CLS_STOP) break;
      if (cls == FB_CLS_SKIP) continue;
      const unsigned ii = e & FB_LIST_IDX_MASK;
      if (cls == FB_CLS_COUNT) fb_count(a, ii, 0u);
      const unsigned o = __ldcg(&sf[ii]);
      if ((o >> FB_RAY_BITS) == tag && (FB_RAY_MASK - (o & FB_RAY_MASK)) < me) break;
    }
  }
}

int fb_ray_resolve_blocks(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ray_resolve, 256, 0) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

cudaError_t fb_ray_frame(const FbGeom &g, const FbRayArgs &a, int nblocks_resolve, cudaStream_t s, int *launches) {
  if (a.n <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((a.n + 127) / 128);
  k_ray_endpoints<<<blocks, 128, 0, s>>>(g, a);
  k_ray_trace<0><<<blocks, 128, 0, s>>>(g, a);
  k_ray_trace<1><<<blocks, 128, 0, s>>>(g, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  long long want = (a.n + 255) / 256;
  int nb = (int)(want < nblocks_resolve ? want : nblocks_resolve);
  if (nb < 1) nb = 1;
  void *args[] = {(void *)&g, (void *)&a};
  e = cudaLaunchCooperativeKernel((void *)k_ray_resolve, dim3(nb), dim3(256), args, 0, s);
  if (launches) *launches += 4;
  return e;
}
