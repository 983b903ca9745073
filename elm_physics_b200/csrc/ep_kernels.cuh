// Step kernels for sm_100a (compiled with -fmad=false; see ep_math.cuh).  One launch of each per
// simulate call, over ALL resident worlds:
//   k_place       Shape.placeIn of every moving body's shapes            (src/Internal/Shape.elm:94-110)
//   k_sort        stable gravity sort per world                          (src/Physics.elm:548-573)
//   k_bp_*        i<j bounding-sphere test, ordered ballot compaction    (src/Internal/BroadPhase.elm:54-215)
//   k_narrowphase shape matrix per candidate pair                        (src/Internal/NarrowPhase.elm, src/Collision/*)
//   k_equations   Jacobian rows + Spook scalars + warm-start lookup      (src/Internal/Equation.elm:116-450)
//   k_islands     union-find (min id root), per-island group lists       (src/Internal/Islands.elm)
//   k_solve       warm start + PGS per island, reference row order       (src/Internal/Solver.elm:39-880)
//   k_integrate   SolverBody.solved                                      (src/Internal/SolverBody.elm:97-225)
#pragma once
#include "ep_device.cuh"
#include "ep_narrowphase.cuh"

namespace ep {

constexpr double kMaxImpulse = 1000000.0;      // Equation.elm:459-461
constexpr double kWarmStartFactor = 0.85;      // Equation.elm:467-469

// The warm-start cache is keyed by bodyKey(min id, max id) (ContactId.elm:62-68, ContactCache.elm).  Ids are unique inside
// a world, so the key is equivalent to the unordered pair of body ROWS; every i<j pair of a world owns one entry of a dense
// table (same size as the pair-slot pool), which makes insert and lookup a single L2-resident access instead of a probe
// sequence through a hash table.
__device__ __forceinline__ int64_t pair_index(const Dev& d, int w, int ra, int rb) {
  const int r0 = d.world_body_off[w];
  const int64_t n = d.world_body_off[w + 1] - r0;
  const int64_t a = (ra < rb ? ra : rb) - r0, b = (ra < rb ? rb : ra) - r0;
  return d.world_pair_off[w] + a * n - a * (a + 1) / 2 + (b - a - 1);
}

// ---- k_place ------------------------------------------------------------------------------------------
__global__ void k_place(Dev d, int n_items) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_items; t += gridDim.x * blockDim.x) {
    int inst = d.pl_inst[t], code = d.pl_code[t];
    int rel = code >> 1;
    int body = d.inst_body[inst];
    const double* src = d.sh_f64 + d.sh_f64_off[d.inst_shape[inst]] + rel;
    double* dst = d.wf64 + d.inst_woff[inst] + rel;
    Q4 q; q.x = d.quat[4 * body]; q.y = d.quat[4 * body + 1]; q.z = d.quat[4 * body + 2]; q.w = d.quat[4 * body + 3];
    V3 v = ld3(src);
    st3(dst, (code & 1) ? rotate(q, v) : point_place(ld3(d.pos + 3 * body), q, v));
  }
}

// ---- k_sort : one CTA per world, rank sort == stable sort under the reference comparator ---------------
__global__ void k_sort(Dev d) {
  for (int w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
    int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    V3 g = ld3(d.gravity + 3 * w);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      V3 p = ld3(d.pos + 3 * (r0 + i));
      d.sort_key[r0 + i] = -(p.x * g.x + p.y * g.y + p.z * g.z);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double a = d.sort_key[r0 + i];
      int rank = 0;
      for (int j = 0; j < n; j++) {
        double b = d.sort_key[r0 + j];
        bool lt = b - a < 0, gt = b - a > 0;          // Physics.elm:563-571
        rank += (lt || (!gt && j < i)) ? 1 : 0;
      }
      d.sorted[r0 + rank] = r0 + i;
    }
    __syncthreads();
  }
}

// The same sort for worlds of hundreds to thousands of bodies (config 4: one world of 2005 -- the one-CTA rank sort above walks
// 2005 keys per body on ONE SM, 0.18 ms): the ranks of different bodies are independent, so `parts` CTAs share a world, 128 bodies
// each, every CTA with its own copy of the world's keys in shared memory (same expression, same bits).
__global__ void __launch_bounds__(128) k_sort_wide(Dev d, int parts) {
  extern __shared__ double s_key[];
  for (int wp = blockIdx.x; wp < d.n_worlds * parts; wp += gridDim.x) {
    const int w = wp / parts, part = wp - w * parts;
    const int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    if (part * 128 >= n) continue;                  // the whole CTA together
    const V3 g = ld3(d.gravity + 3 * w);
    for (int j = threadIdx.x; j < n; j += 128) {
      const V3 p = ld3(d.pos + 3 * (r0 + j));
      s_key[j] = -(p.x * g.x + p.y * g.y + p.z * g.z);
    }
    __syncthreads();
    const int i = part * 128 + threadIdx.x;
    if (i < n) {
      const double a = s_key[i];
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < n; j++) {
        const double b = s_key[j];
        const bool lt = b - a < 0, gt = b - a > 0;          // Physics.elm:563-571
        rank += (lt || (!gt && j < i)) ? 1 : 0;
      }
      d.sorted[r0 + rank] = r0 + i;
    }
    __syncthreads();
  }
}

// ---- k_broadphase -------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_from_index(int64_t p, int n, int& i, int& j) {
  if (n <= 1024) {      // everything below is exact in fp32 / int32 (t*t, 8p < 2^23); the two loops fix a sqrt off by one
    const int q = (int)p;
    const float t = 2.0f * n - 1.0f;
    int ii = (int)((t - sqrtf(t * t - 8.0f * (float)q)) * 0.5f);
    if (ii < 0) ii = 0;
    while (ii * n - ii * (ii + 1) / 2 > q) ii--;
    while ((ii + 1) * n - (ii + 1) * (ii + 2) / 2 <= q) ii++;
    i = ii;
    j = q - (ii * n - ii * (ii + 1) / 2) + ii + 1;
    return;
  }
  double t = 2.0 * n - 1.0;
  int ii = (int)((t - sqrt(t * t - 8.0 * (double)p)) * 0.5);
  if (ii < 0) ii = 0;
  while ((int64_t)ii * n - (int64_t)ii * (ii + 1) / 2 > p) ii--;
  while ((int64_t)(ii + 1) * n - (int64_t)(ii + 1) * (ii + 2) / 2 <= p) ii++;
  i = ii;
  j = (int)(p - ((int64_t)ii * n - (int64_t)ii * (ii + 1) / 2)) + ii + 1;
}
// constraintsBetween (BroadPhase.elm:144-186): +n forward registrations, -n flipped ones, 0 none
__device__ int constraints_between(const Dev& d, int r1, int r2) {
  if (d.n_constraints == 0) return 0;
  if (d.kind[r1] != 2 && d.kind[r2] != 2) return 0;
  int n = 0;
  for (int c = d.con_a_off[r1]; c < d.con_a_off[r1 + 1]; c++) n += d.con_b[c] == r2;
  if (n > 0) return n;
  for (int c = d.con_a_off[r2]; c < d.con_a_off[r2 + 1]; c++) n += d.con_b[c] == r1;
  return -n;
}
// bodiesMayContact (BroadPhase.elm:189-215)
__device__ __forceinline__ bool may_contact(const Dev& d, int w, int r0, int n, int r1, int r2) {
  double rr = d.radius[r1] + d.radius[r2];
  double dx = d.pos[3 * r2] - d.pos[3 * r1], dy = d.pos[3 * r2 + 1] - d.pos[3 * r1 + 1], dz = d.pos[3 * r2 + 2] - d.pos[3 * r1 + 2];
  double dist2 = dx * dx + dy * dy + dz * dz;
  if (!(rr * rr - dist2 > 0)) return false;
  if (!(d.kind[r1] == 2 || d.kind[r2] == 2)) return false;
  if (d.collide) return d.collide[d.collide_off[w] + (int64_t)(r1 - r0) * n + (r2 - r0)] != 0;
  return true;
}

// How deep the bounding spheres of a candidate pair overlap, 0 (just touching) .. kDepthBins-1: a proxy for the cost of its
// narrowphase items (separated hulls leave the SAT at an early axis, interpenetrating ones run every axis and the
// clipping).  Items are ordered by it inside a class so that the lanes of a warp do similar work.  Ordering only.
// The upper half of the bins holds the pairs that had contacts LAST frame (they will most likely run the whole SAT and the
// clipping again), the lower half the pairs that had none: those mostly leave at the axis that separated them last frame
// (ContactSink::hint), so whole warps of them cost one axis.
constexpr int kDepthBins = 8;
__device__ __forceinline__ int depth_bin(const Dev& d, int w, int r1, int r2) {
  double rr = d.radius[r1] + d.radius[r2];
  double dx = d.pos[3 * r2] - d.pos[3 * r1], dy = d.pos[3 * r2 + 1] - d.pos[3 * r1 + 1], dz = d.pos[3 * r2 + 2] - d.pos[3 * r1 + 2];
  double q = 1.0 - (dx * dx + dy * dy + dz * dz) / (rr * rr);
  if (d.quads) {                    // small batches: the narrowphase is the latency of its heaviest warp and the lookup only costs the broadphase
    const int b8 = (int)(q * kDepthBins);
    return b8 < 0 ? 0 : (b8 > kDepthBins - 1 ? kDepthBins - 1 : b8);
  }
  int b = (int)(q * (kDepthBins / 2));
  b = b < 0 ? 0 : (b > kDepthBins / 2 - 1 ? kDepthBins / 2 - 1 : b);
  const bool touched = d.prev.pair_slot[pair_index(d, w, r1, r2)] >= 0;
  return touched ? kDepthBins / 2 + b : b;
}

// Work-item classes: shape sub-kind = convex box 0, convex other 1, plane 2, sphere 3, capsule 4, particle 5;
// class = lo * 6 + hi of the unordered pair, so every lane of a warp in k_narrowphase runs the same Collision.* module.
// (one load: the table is filled by ep_upload_worlds from sh_kind / the hull's box flag -- three dependent loads per shape, and a
// pair of compound bodies walks ns1 x ns2 shape pairs in k_bp_count and again in k_bp_emit)
__device__ __forceinline__ int shape_subkind(const Dev& d, int inst) { return d.inst_sub[inst]; }
__device__ __forceinline__ int item_class(int ka, int kb) { return ka < kb ? ka * 6 + kb : kb * 6 + ka; }
// classes in descending expected cost: the work list starts the long items first
__constant__ int kClassOrder[kClasses] = {7, 1, 10, 0, 4, 9, 8, 3, 2, 22, 28, 21, 16, 11, 5, 29, 17, 23, 35, 14, 15, 20, 6, 12, 13, 18, 19, 24, 25, 26, 27, 30, 31, 32, 33, 34};

// The i<j pairs of a world are cut into chunks of kBpChunk consecutive pair indices (host table, ep_upload_worlds), so a
// world of any size spreads over many CTAs while the slots still come out in pair-enumeration order:
//   k_bp_count  one CTA per chunk: candidate pairs and shape pairs (work items) of the chunk
//   k_bp_scan   one warp per world: exclusive scan over the world's chunks, the world's slot / item ranges
//   k_bp_emit   one CTA per chunk: ordered ballot compaction into the pair slots, work items scattered into class bins
constexpr int kBpChunk = 4096;
__global__ void __launch_bounds__(128) k_bp_count(Dev d) {
  __shared__ int s_warp[4], s_warp2[4];
  __shared__ int s_hist[kClasses * kDepthBins];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kClasses * kDepthBins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int c = blockIdx.x; c < d.n_chunks; c += gridDim.x) {
    const int w = d.chunk_world[c];
    const int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    const int64_t P = (int64_t)n * (n - 1) / 2, p0 = d.chunk_p0[c], p1 = min(P, p0 + kBpChunk);
    int cnt = 0, icnt = 0;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      int i, j; pair_from_index(p, n, i, j);
      int r1 = d.sorted[r0 + i], r2 = d.sorted[r0 + j];
      bool may = may_contact(d, w, r0, n, r1, r2);
      cnt += (may || constraints_between(d, r1, r2) != 0) ? 1 : 0;
      if (may) {
        const int n1 = d.inst_off[r1 + 1] - d.inst_off[r1], n2 = d.inst_off[r2 + 1] - d.inst_off[r2];
        icnt += n1 * n2;
        const int db = depth_bin(d, w, r1, r2);
        for (int a = 0; a < n1; a++)
          for (int b = 0; b < n2; b++) {
            const int ka = shape_subkind(d, d.inst_off[r1] + a), kb = shape_subkind(d, d.inst_off[r2] + b);
            if ((ka == 2 && kb == 2) || (ka == 5 && kb == 5)) continue;          // NarrowPhase.elm:132-134, 235-237
            atomicAdd(&s_hist[item_class(ka, kb) * kDepthBins + db], 1);
          }
      }
    }
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); icnt += __shfl_down_sync(0xffffffffu, icnt, o); }
    if (lane == 0) { s_warp[wid] = cnt; s_warp2[wid] = icnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      d.chunk_cnt[c] = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
      d.chunk_icnt[c] = s_warp2[0] + s_warp2[1] + s_warp2[2] + s_warp2[3];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < kClasses * kDepthBins; i += blockDim.x) if (s_hist[i]) atomicAdd(&d.cur.counters[CNT_CLASS0 + i], s_hist[i]);
}
__global__ void __launch_bounds__(128) k_bp_scan(Dev d) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0) {
    // item array layout: classes in descending expected cost (kClassOrder), each starting at a multiple of 32 (warps are
    // homogeneous in class), deepest overlaps first inside a class.  One thread per class loads its depth bins (all loads in
    // flight at once), one thread chains the 36 class totals, then every class writes its own offsets.
    __shared__ int s_tot[kClasses], s_at[kClasses + 1];
    const int ci = threadIdx.x;
    int cnt[kDepthBins];
    if (ci < kClasses) {
      const int cls = kClassOrder[ci];
      int t = 0;
#pragma unroll
      for (int db = 0; db < kDepthBins; db++) { cnt[db] = d.cur.counters[CNT_CLASS0 + cls * kDepthBins + db]; t += cnt[db]; }
      s_tot[ci] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int at = 0;
      for (int k = 0; k < kClasses; k++) { s_at[k] = at; at = (at + s_tot[k] + 31) & ~31; }
      s_at[kClasses] = at;
    }
    __syncthreads();
    if (ci < kClasses) {
      const int cls = kClassOrder[ci];
      int at = s_at[ci];
      d.np_first[ci] = at;
#pragma unroll
      for (int db = kDepthBins - 1; db >= 0; db--) {
        const int bin = cls * kDepthBins + db;
        d.cls_off[bin] = at; d.cls_cur[bin] = 0;
        at += cnt[db];
      }
      d.np_end[ci] = at;
      if (ci == 0) d.np_first[kClasses] = s_at[kClasses];
    }
  }
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int w = warp; w < d.n_worlds; w += nwarps) {
    const int c0 = d.world_chunk_off[w], c1 = d.world_chunk_off[w + 1];
    int tot = 0, itot = 0;
    for (int c = c0 + lane; c < c1; c += 32) { tot += d.chunk_cnt[c]; itot += d.chunk_icnt[c]; }
    for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, o); itot += __shfl_xor_sync(0xffffffffu, itot, o); }
    int base = 0, ibase = 0, dead = 0;
    if (lane == 0) {
      base = atomicAdd(&d.cur.counters[CNT_CAND], tot);
      ibase = atomicAdd(&d.cur.counters[CNT_ITEMS], itot);
      if ((int64_t)base + tot > d.slot_cap) { d.flags[FLAG_PAIR_OVERFLOW] = 1; dead = 1; }
      if ((int64_t)ibase + itot > d.item_cap) { d.flags[FLAG_PAIR_OVERFLOW] = 5; dead = 1; }
      if (dead) { tot = 0; base = 0; ibase = 0; }
      d.cur.cand_base[w] = base; d.cur.cand_n[w] = tot;
      d.world_min_rem[w] = d.iterations[w];
    }
    base = __shfl_sync(0xffffffffu, base, 0); ibase = __shfl_sync(0xffffffffu, ibase, 0); dead = __shfl_sync(0xffffffffu, dead, 0);
    // exclusive scan over the chunks of the world, 32 at a time
    for (int cb = c0; cb < c1; cb += 32) {
      const int c = cb + lane;
      const int v = c < c1 ? d.chunk_cnt[c] : 0, iv = c < c1 ? d.chunk_icnt[c] : 0;
      int sc = v, isc = iv;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, sc, o), it = __shfl_up_sync(0xffffffffu, isc, o);
        if (lane >= o) { sc += t; isc += it; }
      }
      if (c < c1) { d.chunk_base[c] = dead ? -1 : base + sc - v; d.chunk_ibase[c] = dead ? -1 : ibase + isc - iv; }
      base += __shfl_sync(0xffffffffu, sc, 31); ibase += __shfl_sync(0xffffffffu, isc, 31);
    }
  }
}
__global__ void __launch_bounds__(128) k_bp_emit(Dev d) {
  __shared__ int s_warp[32], s_warp2[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c = blockIdx.x; c < d.n_chunks; c += gridDim.x) {
    const int w = d.chunk_world[c];
    const int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    const int64_t P = (int64_t)n * (n - 1) / 2, pc0 = d.chunk_p0[c], pc1 = min(P, pc0 + kBpChunk);
    int running = d.chunk_base[c], irunning = d.chunk_ibase[c];
    const bool dead = running < 0 || d.chunk_cnt[c] == 0;
    // ordered (enumeration order) ballot compaction of the slots; work items scattered into class bins
    for (int64_t p0 = pc0; p0 < pc1 && !dead; p0 += blockDim.x) {
      int64_t p = p0 + threadIdx.x;
      bool flag = false; int r1 = 0, r2 = 0, ncon = 0, nit = 0, ns2 = 0, dbin = 0; bool may = false;
      if (p < pc1) {
        int i, j; pair_from_index(p, n, i, j);
        r1 = d.sorted[r0 + i]; r2 = d.sorted[r0 + j];
        may = may_contact(d, w, r0, n, r1, r2);
        ncon = constraints_between(d, r1, r2);
        flag = may || ncon != 0;
        if (may) { ns2 = d.inst_off[r2 + 1] - d.inst_off[r2]; nit = (d.inst_off[r1 + 1] - d.inst_off[r1]) * ns2; dbin = depth_bin(d, w, r1, r2); }
      }
      unsigned bal = __ballot_sync(0xffffffffu, flag);
      int iscan = nit;                                            // inclusive warp scan of the item counts
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, iscan, o); if (lane >= o) iscan += v; }
      __syncthreads();
      if (lane == 0) s_warp[wid] = __popc(bal);
      if (lane == 31) s_warp2[wid] = iscan;
      __syncthreads();
      int before = 0, total = 0, ibefore = 0, itotal = 0;
      for (int k = 0; k < nwarps; k++) { int cc = s_warp[k], ic = s_warp2[k]; if (k < wid) { before += cc; ibefore += ic; } total += cc; itotal += ic; }
      int s = running + before + __popc(bal & ((1u << lane) - 1u));
      int item0 = irunning + ibefore + iscan - nit;
      if (flag) {
        d.cur.slot_row1[s] = r1; d.cur.slot_row2[s] = r2;
        d.cur.slot_ccount[s] = 0; d.cur.slot_cstart[s] = 0; d.cur.slot_jcount[s] = 0; d.cur.slot_jstart[s] = 0;
        d.cur.slot_ncon[s] = ncon; d.cur.slot_root[s] = -1;
        d.cur.slot_item0[s] = item0; d.cur.slot_nitems[s] = nit;
      }
      // scatter the shape pairs of this candidate into the class bins (warp-aggregated atomics per class)
      int maxit = nit;
      for (int o = 16; o > 0; o >>= 1) maxit = max(maxit, __shfl_xor_sync(0xffffffffu, maxit, o));
      for (int k = 0; k < maxit; k++) {
        bool act = k < nit;
        int ia = 0, ib = 0, cls = kClasses * kDepthBins;
        if (act) {
          ia = d.inst_off[r1] + k / ns2; ib = d.inst_off[r2] + k % ns2;
          int ka = shape_subkind(d, ia), kb = shape_subkind(d, ib);
          if ((ka == 2 && kb == 2) || (ka == 5 && kb == 5)) { d.raw_cnt[item0 + k] = 0; act = false; }   // NarrowPhase.elm:132-134, 235-237
          else cls = item_class(ka, kb) * kDepthBins + dbin;
        }
        unsigned peers = __match_any_sync(0xffffffffu, cls);
        if (act) {
          int leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
          int basec = 0;
          if (lane == leader) basec = d.cls_off[cls] + atomicAdd(&d.cls_cur[cls], __popc(peers));
          basec = __shfl_sync(peers, basec, leader);
          d.items[basec + rank] = make_int4(s, ia, ib, item0 + k);
        }
      }
      running += total; irunning += itotal;
    }
    __syncthreads();
  }
}

// ---- k_narrowphase : one lane per work item, warps homogeneous in class ---------------------------------------
__device__ __forceinline__ ShapeRef shape_ref(const Dev& d, int inst) {
  ShapeRef s; int sh = d.inst_shape[inst];
  s.kind = d.sh_kind[sh]; s.f = d.wf64 + d.inst_woff[inst]; s.ib = d.sh_i32 + d.sh_i32_off[sh];
  return s;
}

// Separating-axis hints (ContactSink::hint / sep) live in a dense table over the (world, i < j body pair)s, like the warm-start
// index: entry = shape-pair number inside the body pair << 20 | axis id + 1, 0 = none.  Persistent across frames; a pair of
// compound bodies keeps the hint of one of its shape pairs (the others simply find none).
__device__ __forceinline__ bool class_has_sat(int cls) { return cls == 0 || cls == 1 || cls == 7 || cls == 4 || cls == 10; }
__device__ __forceinline__ int hint_index_and_load(const Dev& d, const int4& it, ContactSink& sink, int& kitem) {
  const int s = it.x, r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
  const int idx = (int)pair_index(d, d.body_world[r1], r1, r2);
  kitem = it.w - d.cur.slot_item0[s];
  const int h = d.sep_hint[idx];
  sink.hint = (h != 0 && (h >> 20) == kitem) ? (h & 0xfffff) - 1 : -1;
  return idx;
}
__device__ __forceinline__ void hint_store(const Dev& d, int idx, int kitem, const ContactSink& sink) {
  if (sink.sep >= 0 && sink.sep < 0xfffff && kitem < 2048) { if (sink.sep != sink.hint) d.sep_hint[idx] = (kitem << 20) | (sink.sep + 1); }
  else if (sink.hint >= 0) d.sep_hint[idx] = 0;       // the pair touches now (or the hint no longer separates)
}

// L = items per warp (lanes 0..L-1 work): 32 packs the warps full; a smaller L spreads the same items over more warps, which
// cuts the serialisation of divergent lanes inside a warp -- it pays when there are fewer warps than the GPU has room for
// (small batches: the kernel time is the latency of its heaviest warp).
__global__ void __launch_bounds__(128) k_narrowphase(Dev d, int only_cls, int L, unsigned long long skip_mask) {
  TagPt pa[kMaxPoly], pb[kMaxPoly];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (d.flags[FLAG_PAIR_OVERFLOW]) return;      // a world's pairs / items did not fit: its items were counted but never emitted
  int ci = 0;                     // position in kClassOrder of the class this warp's chunk belongs to
  const int total = d.np_first[kClasses];
  for (int chunk = warp; chunk * L < total; chunk += nwarps) {
    while (ci + 1 < kClasses && chunk * L >= d.np_first[ci + 1]) ci++;
    const int cls = kClassOrder[ci];
    const int idx = chunk * L + lane;
    if (lane >= L || idx >= d.np_end[ci] || (only_cls >= 0 && cls != only_cls) || ((skip_mask >> cls) & 1ull)) continue;
    int4 it = d.items[idx];
    ContactSink sink; sink.buf = d.raw + (size_t)it.w * kRawStride; sink.n = 0; sink.cap = kRawStride; sink.overflow = false;
    bool polyOvf = false;
    const bool sat = class_has_sat(cls);
    int kitem = 0, hidx = 0;
    if (sat) hidx = hint_index_and_load(d, it, sink, kitem);
    shape_pair_contacts(shape_ref(d, it.y), shape_ref(d, it.z), sink, pa, pb, polyOvf);
    if (sat) hint_store(d, hidx, kitem, sink);
    // solver order = reverse emission inside a shape pair (Equation.elm:135-154)
    for (int x = 0, y = sink.n - 1; x < y; x++, y--) { RawContact t = sink.buf[x]; sink.buf[x] = sink.buf[y]; sink.buf[y] = t; }
    if (sink.overflow) d.flags[FLAG_PAIR_OVERFLOW] = 2;
    if (polyOvf) d.flags[FLAG_POLY_OVERFLOW] = 1;
    d.raw_cnt[it.w] = sink.n;
  }
}

// The same work as three kernels, one per GROUP of Collision.* modules (shape_pair_contacts_group), launched side by side:
// class (lo, hi) of sub-kinds 0 box, 1 convex, 2 plane, 3 sphere, 4 capsule, 5 particle -> group 0 hull-hull without boxes (class 7),
// group 3 box-box / box-hull (0, 1), group 1 capsule-hull (4, 10), group 2 the rest.  Each instance carries only its modules' code (the one kernel for everything is
// 26 k SASS lines and stalls on instruction fetch) and only group 0 (and the plane-hull path of group 2) the clip-polygon buffers.
__host__ __device__ constexpr int np_group_of_class(int cls) { return cls == 7 ? 0 : ((cls == 0 || cls == 1) ? 3 : ((cls == 4 || cls == 10) ? 1 : 2)); }
template <int G>
__global__ void __launch_bounds__(128) k_narrowphase_group(Dev d, int L) {
  TagPt pa[G == 1 ? 1 : kMaxPoly], pb[G == 1 ? 1 : kMaxPoly];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (d.flags[FLAG_PAIR_OVERFLOW]) return;
  int chunk0 = 0;                 // chunks of this group's classes in front of class position ci
  for (int ci = 0; ci < kClasses; ci++) {
    if (np_group_of_class(kClassOrder[ci]) != G) continue;
    const int first = d.np_first[ci], end = d.np_end[ci], nch = (end - first + L - 1) / L;
    // this warp's chunks of the class: global chunk index g = chunk0 + c, g % nwarps == warp
    int c = (warp - chunk0 % nwarps + nwarps) % nwarps;
    for (; c < nch; c += nwarps) {
      const int idx = first + c * L + lane;
      if (lane >= L || idx >= end) continue;
      int4 it = d.items[idx];
      ContactSink sink; sink.buf = d.raw + (size_t)it.w * kRawStride; sink.n = 0; sink.cap = kRawStride; sink.overflow = false;
      bool polyOvf = false;
      int kitem = 0, hidx = 0;
      if (G != 2) hidx = hint_index_and_load(d, it, sink, kitem);
      shape_pair_contacts_group<G>(shape_ref(d, it.y), shape_ref(d, it.z), sink, pa, pb, polyOvf);
      if (G != 2) hint_store(d, hidx, kitem, sink);
      // solver order = reverse emission inside a shape pair (Equation.elm:135-154)
      for (int x = 0, y = sink.n - 1; x < y; x++, y--) { RawContact t = sink.buf[x]; sink.buf[x] = sink.buf[y]; sink.buf[y] = t; }
      if (sink.overflow) d.flags[FLAG_PAIR_OVERFLOW] = 2;
      if (polyOvf) d.flags[FLAG_POLY_OVERFLOW] = 1;
      d.raw_cnt[it.w] = sink.n;
    }
    chunk0 += nch;
  }
}

// Classes 7 (hull-hull without boxes), 10 and 4 (capsule against a hull / a box) with a QUAD of lanes per item (convex_convex_quad,
// capsule_convex_quad, ep_narrowphase.cuh): 8 items per warp.  For small batches, where the kernel time is the latency of the
// heaviest warp; beside k_narrowphase(skip_mask = kQuadClasses).
constexpr unsigned long long kQuadClasses = (1ull << 7) | (1ull << 10) | (1ull << 4);
__global__ void __launch_bounds__(128) k_narrowphase_quad(Dev d) {
  TagPt pa[kMaxPoly], pb[kMaxPoly];
  const int lane = threadIdx.x & 31, q = lane & 3, sub = lane >> 2;
  const unsigned qm = 0xFu << (lane & ~3);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (d.flags[FLAG_PAIR_OVERFLOW]) return;
  int chunk0 = 0;                 // chunks of the quad classes in front of class position ci: the classes' chunks are dealt on over the warps
  for (int ci = 0; ci < kClasses; ci++) {
    if (!((kQuadClasses >> kClassOrder[ci]) & 1ull)) continue;
    const int first = d.np_first[ci], end = d.np_end[ci], nch = (end - first + 7) / 8;
    int c = (warp - chunk0 % nwarps + nwarps) % nwarps;
    chunk0 += nch;
    for (; c < nch; c += nwarps) {
      const int idx = first + c * 8 + sub;
      if (idx >= end) continue;                      // the four lanes of a quad together
      int4 it = d.items[idx];
      ContactSink sink; sink.buf = d.raw + (size_t)it.w * kRawStride; sink.n = 0; sink.cap = kRawStride; sink.overflow = false;
      bool polyOvf = false;
      int kitem = 0;
      const int hidx = hint_index_and_load(d, it, sink, kitem);
      sink.hint = __shfl_sync(qm, sink.hint, lane & ~3);          // one value for the quad (the table may change under it)
      const ShapeRef a = shape_ref(d, it.y), b = shape_ref(d, it.z);
      if (a.kind == SH_CONVEX && b.kind == SH_CONVEX) convex_convex_quad(make_cvx(a.f, a.ib), make_cvx(b.f, b.ib), sink, pa, pb, polyOvf, q, qm);
      else if (a.kind == SH_CONVEX && b.kind == SH_CAPSULE) capsule_convex_quad(true, as_capsule(b), make_cvx(a.f, a.ib), sink, q, qm);
      else if (a.kind == SH_CAPSULE && b.kind == SH_CONVEX) capsule_convex_quad(false, as_capsule(a), make_cvx(b.f, b.ib), sink, q, qm);
      if (q == 0) {
        // solver order = reverse emission inside a shape pair (Equation.elm:135-154)
        for (int x = 0, y = sink.n - 1; x < y; x++, y--) { RawContact t = sink.buf[x]; sink.buf[x] = sink.buf[y]; sink.buf[y] = t; }
        if (sink.overflow) d.flags[FLAG_PAIR_OVERFLOW] = 2;
        if (polyOvf) d.flags[FLAG_POLY_OVERFLOW] = 1;
        d.raw_cnt[it.w] = sink.n;
        hint_store(d, hidx, kitem, sink);
      }
    }
  }
}

// ---- k_slot_counts / k_equations / k_joint_equations : rows of every contact and constraint ---------------
struct BodyK { int kind; double invMass; M33 I; };
__device__ __forceinline__ BodyK load_bodyk(const Dev& d, int r) {
  BodyK b; b.kind = d.kind[r]; b.invMass = d.inv_mass[r];
  const double* m = d.iiw + 9 * r;
  b.I.m11 = m[0]; b.I.m21 = m[1]; b.I.m31 = m[2]; b.I.m12 = m[3]; b.I.m22 = m[4]; b.I.m32 = m[5]; b.I.m13 = m[6]; b.I.m23 = m[7]; b.I.m33 = m[8];
  return b;
}
struct BodyE { BodyK k; V3 x, v, w, f, tq; Q4 q; };
__device__ __forceinline__ BodyE load_bodye(const Dev& d, int r) {
  BodyE b; b.k = load_bodyk(d, r);
  b.x = ld3(d.pos + 3 * r); b.v = ld3(d.vel + 3 * r); b.w = ld3(d.angvel + 3 * r); b.f = ld3(d.force + 3 * r); b.tq = ld3(d.torque + 3 * r);
  b.q.x = d.quat[4 * r]; b.q.y = d.quat[4 * r + 1]; b.q.z = d.quat[4 * r + 2]; b.q.w = d.quat[4 * r + 3];
  return b;
}
// J = (wA, vB, wB), Equation.elm:26-36 ; wA = n x ri, vB = n, wB = rj x n (389-400)
__device__ __forceinline__ void contact_jac(V3 n, V3 ri, V3 rj, double* J) {
  J[0] = n.y * ri.z - n.z * ri.y; J[1] = n.z * ri.x - n.x * ri.z; J[2] = n.x * ri.y - n.y * ri.x;
  J[3] = n.x; J[4] = n.y; J[5] = n.z;
  J[6] = rj.y * n.z - rj.z * n.y; J[7] = rj.z * n.x - rj.x * n.z; J[8] = rj.x * n.y - rj.y * n.x;
}
// computeGW, Equation.elm:626-631
__device__ __forceinline__ double compute_gw(const BodyE& bi, const BodyE& bj, const double* J) {
  return -(J[3] * bi.v.x + J[4] * bi.v.y + J[5] * bi.v.z)
         + (J[0] * bi.w.x + J[1] * bi.w.y + J[2] * bi.w.z)
         + (J[3] * bj.v.x + J[4] * bj.v.y + J[5] * bj.v.z)
         + (J[6] * bj.w.x + J[7] * bj.w.y + J[8] * bj.w.z);
}
// computeGimgt, Equation.elm:612-621
__device__ __forceinline__ double compute_gimgt(const BodyE& bi, const BodyE& bj, const double* J) {
  const M33 &I = bi.k.I, &K = bj.k.I;
  return bi.k.invMass + bj.k.invMass
         + (J[0] * (I.m11 * J[0] + I.m12 * J[1] + I.m13 * J[2]))
         + (J[1] * (I.m21 * J[0] + I.m22 * J[1] + I.m23 * J[2]))
         + (J[2] * (I.m31 * J[0] + I.m32 * J[1] + I.m33 * J[2]))
         + (J[6] * (K.m11 * J[6] + K.m12 * J[7] + K.m13 * J[8]))
         + (J[7] * (K.m21 * J[6] + K.m22 * J[7] + K.m23 * J[8]))
         + (J[8] * (K.m31 * J[6] + K.m32 * J[7] + K.m33 * J[8]));
}
// computeGiMf, Equation.elm:583-607
__device__ __forceinline__ double compute_gimf(V3 gravity, const BodyE& bi, const BodyE& bj, const double* J) {
  V3 gi = bi.k.kind == 2 ? gravity : mk(0, 0, 0), gj = bj.k.kind == 2 ? gravity : mk(0, 0, 0);
  const M33 &I = bi.k.I, &K = bj.k.I;
  return -(J[3] * (bi.k.invMass * bi.f.x + gi.x) + J[4] * (bi.k.invMass * bi.f.y + gi.y) + J[5] * (bi.k.invMass * bi.f.z + gi.z))
         + (J[3] * (bj.k.invMass * bj.f.x + gj.x) + J[4] * (bj.k.invMass * bj.f.y + gj.y) + J[5] * (bj.k.invMass * bj.f.z + gj.z))
         + (J[0] * (I.m11 * bi.tq.x + I.m12 * bi.tq.y + I.m13 * bi.tq.z))
         + (J[1] * (I.m21 * bi.tq.x + I.m22 * bi.tq.y + I.m23 * bi.tq.z))
         + (J[2] * (I.m31 * bi.tq.x + I.m32 * bi.tq.y + I.m33 * bi.tq.z))
         + (J[6] * (K.m11 * bj.tq.x + K.m12 * bj.tq.y + K.m13 * bj.tq.z))
         + (J[7] * (K.m21 * bj.tq.x + K.m22 * bj.tq.y + K.m23 * bj.tq.z))
         + (J[8] * (K.m31 * bj.tq.x + K.m32 * bj.tq.y + K.m33 * bj.tq.z));
}
// computeContactB, Equation.elm:532-546
__device__ __forceinline__ double compute_contact_b(double spookA, double spookB, double e, V3 pi, V3 pj, V3 ni, const BodyE& bi, const BodyE& bj, const double* J) {
  double g = ((pj.x - pi.x) * ni.x) + ((pj.y - pi.y) * ni.y) + ((pj.z - pi.z) * ni.z);
  double gW = (e + 1) * (vdot(bj.v, ni) - vdot(bi.v, ni))
              + (bj.w.x * J[6] + bj.w.y * J[7] + bj.w.z * J[8])
              + (bi.w.x * J[0] + bi.w.y * J[1] + bi.w.z * J[2]);
  return -g * spookA - gW * spookB;
}
// invInertiaWorld * angular Jacobian of a row, as applyVelocityBody1/2 forms it (Solver.elm:505-550); zero for non-dynamic bodies
__device__ __forceinline__ void inertia_times(const BodyK& k, const double* w, double* out) {
  if (k.kind != 2) { out[0] = 0; out[1] = 0; out[2] = 0; return; }
  const M33& I = k.I;
  out[0] = I.m11 * w[0] + I.m12 * w[1] + I.m13 * w[2];
  out[1] = I.m21 * w[0] + I.m22 * w[1] + I.m23 * w[2];
  out[2] = I.m31 * w[0] + I.m32 * w[1] + I.m33 * w[2];
}
struct Spook { double a, b, eps; };
__device__ __forceinline__ Spook spook(double dt) {      // Equation.elm:369-376, 494-501
  Spook s;
  s.a = 4.0 / (dt * (1 + 4 * 3.0));
  s.b = (4.0 * 3.0) / (1 + 4 * 3.0);
  s.eps = 4.0 / (dt * dt * 10000000.0 * (1 + 4 * 3.0));
  return s;
}
// a row in the layout of the quad solver (QRow, ep_device.cuh): per lane role the dot coefficients a and the update coefficients c
__device__ __forceinline__ void store_qrow(QRow& q, const double* J, const double* Ia, const double* Ib, double B, double invC, double lam, double mu) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    q.a[k][0] = -J[3 + k]; q.a[k][1] = J[k]; q.a[k][2] = J[3 + k]; q.a[k][3] = J[6 + k];
    q.c[k][0] = Ia[k]; q.c[k][1] = Ib[k];
  }
  q.B = B; q.invC = invC; q.lam = lam; q.mu = mu;
  q.i0 = 0; q.i1 = 0; q.i2 = 0; q.i3 = 0;
}
__device__ __forceinline__ void joint_row(JointRow& r, double dt, V3 gravity, const BodyE& b1, const BodyE& b2, double velocityB, double eps) {
  r.B = velocityB - (dt * compute_gimf(gravity, b1, b2, r.J));     // computeSolverB 520-522
  r.invC = 1 / (compute_gimgt(b1, b2, r.J) + eps);                 // computeSolverInvC 527-529
  r.eps = eps; r.lam = 0;
  inertia_times(b1.k, r.J, r.Ia); inertia_times(b2.k, r.J + 6, r.Ib);
}
// point-to-point rows for basis x,y,z (Equation.elm:308-361); writes 3 rows in EMISSION order
__device__ void p2p_rows(JointRow* out, double dt, V3 gravity, const Spook& sp, const BodyE& b1, const BodyE& b2, V3 pivot1, V3 pivot2) {
  V3 ri = rotate(b1.q, pivot1), rj = rotate(b2.q, pivot2);
  for (int a = 0; a < 3; a++) {
    V3 ni = mk(a == 0 ? 1.0 : 0.0, a == 1 ? 1.0 : 0.0, a == 2 ? 1.0 : 0.0);
    V3 pi = vadd(b1.x, ri), pj = vadd(b2.x, rj);
    contact_jac(ni, ri, rj, out[a].J);
    joint_row(out[a], dt, gravity, b1, b2, compute_contact_b(sp.a, sp.b, 0, pi, pj, ni, b1, b2, out[a].J), sp.eps);
  }
}
// rotational row (Equation.elm:272-305, 549-565)
__device__ void rot_row(JointRow& r, double dt, V3 gravity, const Spook& sp, const BodyE& b1, const BodyE& b2, V3 ni, V3 nj) {
  r.J[0] = nj.y * ni.z - nj.z * ni.y; r.J[1] = nj.z * ni.x - nj.x * ni.z; r.J[2] = nj.x * ni.y - nj.y * ni.x;
  r.J[3] = 0; r.J[4] = 0; r.J[5] = 0;
  r.J[6] = ni.y * nj.z - ni.z * nj.y; r.J[7] = ni.z * nj.x - ni.x * nj.z; r.J[8] = ni.x * nj.y - ni.y * nj.x;
  double g = 0 - vdot(ni, nj);
  double gW = compute_gw(b1, b2, r.J);
  joint_row(r, dt, gravity, b1, b2, -g * sp.a - gW * sp.b, sp.eps);
}
__device__ __forceinline__ int constraint_rows(int kind) { return kind == 0 ? 3 : (kind == 1 ? 5 : (kind == 2 ? 6 : 1)); }

// contactEquations (Equation.elm:364-450) for one contact; lam/ct1 = cached normal lambda and friction-1 direction.
// One row at a time, every row stored as soon as it is known (the records live in global memory): the thread never holds
// more than one 9-double Jacobian next to the two bodies.  Q: rows in the quad solver's layout (three QRow at qrows) instead
// of the lane solver's RowRec.  Returns the friction-1 direction used.
template <bool Q>
__device__ __forceinline__ V3 contact_row(RowRec* rec, QRow* qrows, V3 ni, V3 pi, V3 pj, double bounciness, double friction, double lam, V3 ct1,
                                          double dt, V3 gravity, const Spook& sp, const BodyE& b1, const BodyE& b2) {
  V3 ri = vsub(pi, b1.x), rj = vsub(pj, b2.x);
  V3 t1, t2; stable_tangents(ct1, ni, t1, t2);
  double J[9], ia[3], ib[3];
  {
    contact_jac(ni, ri, rj, J);
    const double B = compute_contact_b(sp.a, sp.b, bounciness, pi, pj, ni, b1, b2, J) - (dt * compute_gimf(gravity, b1, b2, J));
    const double invC = 1 / (compute_gimgt(b1, b2, J) + sp.eps);
    inertia_times(b1.k, J, ia); inertia_times(b2.k, J + 6, ib);
    if (Q) store_qrow(qrows[0], J, ia, ib, B, invC, lam * kWarmStartFactor, 0.0);
    else {
      RowRec& r = *rec;
      r.nB = B; r.nInvC = invC;
#pragma unroll
      for (int k = 0; k < 9; k++) r.nJ[k] = J[k];
#pragma unroll
      for (int k = 0; k < 3; k++) { r.nIa[k] = ia[k]; r.nIb[k] = ib[k]; }
    }
  }
  {
    contact_jac(t1, ri, rj, J);
    const double B = (-compute_gw(b1, b2, J) * sp.b) - (dt * compute_gimf(gravity, b1, b2, J));
    const double invC = 1 / (compute_gimgt(b1, b2, J) + sp.eps);
    inertia_times(b1.k, J, ia); inertia_times(b2.k, J + 6, ib);
    if (Q) store_qrow(qrows[1], J, ia, ib, B, invC, 0.0, friction);
    else {
      RowRec& r = *rec;
      r.f1B = B; r.f1InvC = invC;
#pragma unroll
      for (int k = 0; k < 9; k++) r.f1J[k] = J[k];
#pragma unroll
      for (int k = 0; k < 3; k++) { r.f1Ia[k] = ia[k]; r.f1Ib[k] = ib[k]; }
    }
  }
  {
    contact_jac(t2, ri, rj, J);
    const double B = (-compute_gw(b1, b2, J) * sp.b) - (dt * compute_gimf(gravity, b1, b2, J));
    const double invC = 1 / (compute_gimgt(b1, b2, J) + sp.eps);
    inertia_times(b1.k, J, ia); inertia_times(b2.k, J + 6, ib);
    if (Q) store_qrow(qrows[2], J, ia, ib, B, invC, 0.0, friction);
    else {
      RowRec& r = *rec;
      r.f2B = B; r.f2InvC = invC;
#pragma unroll
      for (int k = 0; k < 9; k++) r.f2J[k] = J[k];
#pragma unroll
      for (int k = 0; k < 3; k++) { r.f2Ia[k] = ia[k]; r.f2Ib[k] = ib[k]; }
    }
  }
  if (!Q) { RowRec& r = *rec; r.mu = friction; r.lamN = lam * kWarmStartFactor; r.lamF1 = 0; r.lamF2 = 0; r.dN = 0; r.dF1 = 0; r.dF2 = 0; }
  return t1;
}
// addConstraintEquations (Equation.elm:157-361) for the pair (r1, r2) (global body rows; ncon < 0 == flipped registration):
// rows are emitted per constraint in list order and consed, so the solver order is the reverse of emission:
// row e of E lands at jstart + E-1-e.
__device__ __forceinline__ int joint_row_count(const Dev& d, int r1, int r2, int ncon) {
  const bool flipped = ncon < 0;
  const int ra = flipped ? r2 : r1, rb = flipped ? r1 : r2;
  int E = 0;
  for (int c = d.con_a_off[ra]; c < d.con_a_off[ra + 1]; c++) if (d.con_b[c] == rb) E += constraint_rows(d.con_kind[c]);
  return E;
}
// the E rows land at [j0, j0 + E) (allocated by k_slot_counts)
__device__ void joint_rows_for_pair(const Dev& d, int r1, int r2, int ncon, const BodyE& b1, const BodyE& b2, double dt, V3 gravity,
                                    const Spook& sp, int j0, int E) {
  bool flipped = ncon < 0;
  int ra = flipped ? r2 : r1, rb = flipped ? r1 : r2;
  int e = 0;
  for (int c = d.con_a_off[ra]; c < d.con_a_off[ra + 1] && E > 0; c++) {
    if (d.con_b[c] != rb) continue;
    const double* A = d.con_data + 24 * c;      // registering body's side
    const double* Bd = A + 12;                   // partner's side
    const double* s1 = flipped ? Bd : A;         // relativeToCenterOfMassFlipped, Constraint.elm:64-95
    const double* s2 = flipped ? A : Bd;
    int kind = d.con_kind[c];
    JointRow tmp[6];
    int nr = constraint_rows(kind);
    if (kind == 3) {                              // Distance, Equation.elm:175-228
      double half = A[0] / 2;
      V3 ni = vdirection(b2.x, b1.x);
      V3 ri = vscale(half, ni), rj = vscale(-half, ni);
      V3 pi = vadd(ri, b1.x), pj = vadd(rj, b2.x);
      contact_jac(ni, ri, rj, tmp[0].J);
      joint_row(tmp[0], dt, gravity, b1, b2, compute_contact_b(sp.a, sp.b, 0, pi, pj, ni, b1, b2, tmp[0].J), sp.eps);
    } else {
      p2p_rows(tmp, dt, gravity, sp, b1, b2, ld3(s1), ld3(s2));
      if (kind == 1) {                            // Hinge, Equation.elm:231-242
        V3 worldAxis2 = rotate(b2.q, ld3(s2 + 3));
        V3 n1, n2; tangents(rotate(b1.q, ld3(s1 + 3)), n1, n2);
        rot_row(tmp[3], dt, gravity, sp, b1, b2, n1, worldAxis2);
        rot_row(tmp[4], dt, gravity, sp, b1, b2, n2, worldAxis2);
      } else if (kind == 2) {                     // Lock, Equation.elm:245-269
        V3 x1 = rotate(b1.q, ld3(s1 + 3)), y1 = rotate(b1.q, ld3(s1 + 6)), z1 = rotate(b1.q, ld3(s1 + 9));
        V3 x2 = rotate(b2.q, ld3(s2 + 3)), y2 = rotate(b2.q, ld3(s2 + 6)), z2 = rotate(b2.q, ld3(s2 + 9));
        rot_row(tmp[3], dt, gravity, sp, b1, b2, x1, y2);
        rot_row(tmp[4], dt, gravity, sp, b1, b2, y1, z2);
        rot_row(tmp[5], dt, gravity, sp, b1, b2, z1, x2);
      }
    }
    for (int k = 0; k < nr; k++, e++) {
      if (d.quads) store_qrow(d.jqrows[j0 + (E - 1 - e)], tmp[k].J, tmp[k].Ia, tmp[k].Ib, tmp[k].B, tmp[k].invC, 0.0, 0.0);
      else { tmp[k].dl = 0; tmp[k].pad[0] = 0; tmp[k].pad[1] = 0; tmp[k].pad[2] = 0; tmp[k].pad[3] = 0; d.jrows[j0 + (E - 1 - e)] = tmp[k]; }
    }
  }
}

// k_slot_counts: one thread per pair slot.  Contact and joint-row ranges of every pair group (the two pool allocations are
// warp-aggregated: one atomic per warp), the warm-start cache index of this frame, and for every contact of the frame where
// it comes from (slot, raw contact) -- so that k_equations runs one thread per CONTACT (a thread per group ran 9 of 32 lanes:
// groups hold 1..8 contacts) and k_islands, which needs only the counts, runs beside it.
// Pair slots of the frame.  After a pair-slot / work-item overflow (a world's slots were counted but never emitted, and the
// narrowphase did not run) there are none: the frame is void and ep_sync reports the overflow.
__device__ __forceinline__ int slot_total(const Dev& d) {
  return d.flags[FLAG_PAIR_OVERFLOW] ? 0 : min(d.cur.counters[CNT_CAND], (int)d.slot_cap);
}
__global__ void __launch_bounds__(256) k_slot_counts(Dev d) {
  const int total = slot_total(d);
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  for (int s0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; s0 < total; s0 += stride) {
    const int s = s0 + lane;
    int nc = 0, E = 0, ncon = 0, it0 = 0, nit = 0, r1 = 0, r2 = 0;
    if (s < total) {
      ncon = d.cur.slot_ncon[s]; it0 = d.cur.slot_item0[s]; nit = d.cur.slot_nitems[s];
      for (int k = 0; k < nit; k++) nc += d.raw_cnt[it0 + k];
      if (nc > 0 || ncon != 0) { r1 = d.cur.slot_row1[s]; r2 = d.cur.slot_row2[s]; }
      if (ncon != 0) E = joint_row_count(d, r1, r2, ncon);
    }
    int cs = nc, js = E;                                          // inclusive warp scans
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, cs, o), b = __shfl_up_sync(0xffffffffu, js, o);
      if (lane >= o) { cs += a; js += b; }
    }
    int cbase = 0, jbase = 0;
    if (lane == 31) {
      if (cs > 0) cbase = atomicAdd(&d.cur.counters[CNT_CONTACTS], cs);
      if (js > 0) jbase = atomicAdd(&d.cur.counters[CNT_JROWS], js);
    }
    cbase = __shfl_sync(0xffffffffu, cbase, 31); jbase = __shfl_sync(0xffffffffu, jbase, 31);
    const int ctot = __shfl_sync(0xffffffffu, cs, 31), jtot = __shfl_sync(0xffffffffu, js, 31);
    if (cbase + ctot > d.contact_cap) { if (nc > 0) d.flags[FLAG_POOL_OVERFLOW] = 1; nc = 0; }
    if (jbase + jtot > d.jrow_cap) { if (E > 0) d.flags[FLAG_JOINT_OVERFLOW] = 1; E = 0; }
    if (nc > 0) {
      const int c0 = cbase + cs - nc;
      d.cur.slot_cstart[s] = c0; d.cur.slot_ccount[s] = nc;
      d.cur.pair_slot[pair_index(d, d.body_world[r1], r1, r2)] = s;
      int out = c0;
      for (int k = 0; k < nit; k++) {                    // shape1 outer, shape2 inner (NarrowPhase.elm:33-59)
        const int n = d.raw_cnt[it0 + k];
        for (int x = 0; x < n; x++) d.contact_src[out++] = make_int2(s, (it0 + k) * kRawStride + x);
      }
    }
    if (ncon != 0) { d.cur.slot_jstart[s] = E > 0 ? jbase + js - E : 0; d.cur.slot_jcount[s] = E; }
  }
}

// The identity and geometry of contact `out` of the frame (ContactRec without lam / t1) and its warm-start seed: the cached
// normal lambda and friction-1 direction of the same (bodyKey, shapeKey, featureKey) in the previous frame, or zeros.
__device__ __forceinline__ void gather_contact(const Dev& d, int out, ContactRec& c, double& lam, V3& ct1, int& r1, int& r2, int& w) {
  const int2 src = d.contact_src[out];
  const int s = src.x;
  r1 = d.cur.slot_row1[s]; r2 = d.cur.slot_row2[s];
  w = d.body_world[r1];
  const int k = src.y / kRawStride - d.cur.slot_item0[s];
  const int i1b = d.inst_off[r1], i2b = d.inst_off[r2], ns2 = d.inst_off[r2 + 1] - i2b;
  const int a = i1b + k / ns2, b = i2b + k % ns2;
  const int sk = (k / ns2 + 1) * 256 + (k % ns2 + 1);     // ContactId.elm:75-77
  const double f1 = d.inst_friction[a], f2 = d.inst_friction[b], e1 = d.inst_bounciness[a], e2 = d.inst_bounciness[b];
  const RawContact rc = d.raw[src.y];
  c.fk = rc.fk; c.sk = sk; c.slot = s;
  c.ni[0] = rc.ni.x; c.ni[1] = rc.ni.y; c.ni[2] = rc.ni.z;
  c.pi[0] = rc.pi.x; c.pi[1] = rc.pi.y; c.pi[2] = rc.pi.z;
  c.pj[0] = rc.pj.x; c.pj[1] = rc.pj.y; c.pj[2] = rc.pj.z;
  c.friction = sqrt(f1 * f2);                                 // Material.elm:23-25
  c.bounciness = (e1 + e2 + eabs(e1 - e2)) * 0.5;             // Material.elm:31-33
  // Cache.getGroup (bodyKey id1 id2), Equation.elm:122-128; Cache.lookup: first match in pairGroup.contacts order == LAST
  // match in solver order (ContactCache.elm:76-87)
  lam = 0; ct1 = mk(0, 0, 0);
  const int ps = d.prev.pair_slot[pair_index(d, w, r1, r2)];
  if (ps >= 0) {
    const int pc0 = d.prev.slot_cstart[ps], pcn = d.prev.slot_ccount[ps];
    // the keys of the last (up to) 8 cached contacts in one batch of independent loads (a loop with an early exit serialises one
    // L2 round trip per candidate), then the match in registers; longer groups continue with the loop
    int hit = -1;
    {
      int64_t fk8[8]; int sk8[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int m = pcn - 1 - j;
        const ContactRec& p = d.prev.contacts[pc0 + (m >= 0 ? m : 0)];
        fk8[j] = p.fk; sk8[j] = p.sk;
      }
#pragma unroll
      for (int j = 7; j >= 0; j--) if (pcn - 1 - j >= 0 && (fk8[j] - rc.fk == 0) && (sk8[j] - sk == 0)) hit = pcn - 1 - j;      // ends with the LARGEST matching m
    }
    if (hit < 0)
      for (int m = pcn - 9; m >= 0; m--) {
        const ContactRec& p = d.prev.contacts[pc0 + m];
        if ((p.fk - rc.fk == 0) && (p.sk - sk == 0)) { hit = m; break; }
      }
    if (hit >= 0) { const ContactRec& p = d.prev.contacts[pc0 + hit]; lam = p.lam; ct1 = mk(p.t1[0], p.t1[1], p.t1[2]); }
  }
}
// k_equations: one thread per contact (contactEquations), rows to the per-contact records (RowRec / QRow)
#ifndef EP_EQ_OCC
#define EP_EQ_OCC 3
#endif
template <bool Q>
__global__ void __launch_bounds__(128, EP_EQ_OCC) k_equations(Dev d) {
  if (d.flags[FLAG_POOL_OVERFLOW]) return;       // contact_src has holes then; ep_sync reports the overflow
  const int ncont = min(d.cur.counters[CNT_CONTACTS], d.contact_cap);
  for (int out = blockIdx.x * blockDim.x + threadIdx.x; out < ncont; out += gridDim.x * blockDim.x) {
    ContactRec c; double lam; V3 ct1; int r1, r2, w;
    gather_contact(d, out, c, lam, ct1, r1, r2, w);
    const double dt = d.dt[w]; const V3 gravity = ld3(d.gravity + 3 * w);
    const BodyE b1 = load_bodye(d, r1), b2 = load_bodye(d, r2);
    const V3 t1 = contact_row<Q>(Q ? nullptr : d.rows + out, Q ? d.qrows + 3 * (size_t)out : nullptr, mk(c.ni[0], c.ni[1], c.ni[2]), mk(c.pi[0], c.pi[1], c.pi[2]),
                                 mk(c.pj[0], c.pj[1], c.pj[2]), c.bounciness, c.friction, lam, ct1, dt, gravity, spook(dt), b1, b2);
    c.lam = lam * kWarmStartFactor; c.t1[0] = t1.x; c.t1[1] = t1.y; c.t1[2] = t1.z;
    d.cur.contacts[out] = c;
  }
}
// addConstraintEquations: one thread per constrained pair (launched only when the batch has constraints)
__global__ void __launch_bounds__(128, 2) k_joint_equations(Dev d) {
  const int total = slot_total(d);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
    const int ncon = d.cur.slot_ncon[s];
    if (ncon == 0) continue;
    const int E = d.cur.slot_jcount[s];
    if (E <= 0) continue;
    const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s], w = d.body_world[r1];
    const double dt = d.dt[w];
    joint_rows_for_pair(d, r1, r2, ncon, load_bodye(d, r1), load_bodye(d, r2), dt, ld3(d.gravity + 3 * w), spook(dt), d.cur.slot_jstart[s], E);
  }
}

// ---- k_islands : one thread per world -----------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {       // Islands.elm:125-145 (path halving)
  for (;;) {
    int p = parent[x];
    if (p == x) return x;
    int g = parent[p];
    if (g == p) return p;
    parent[x] = g;
    x = g;
  }
}
// island size class: row count (3 per contact + joint rows) on a log2 scale; class kBins-1 runs first
// The class also bounds the island's pair groups (<= 2^class; hence its dynamic bodies <= 2^class + 1: a component of n dynamic
// bodies needs n - 1 groups), which is what the quad solver sizes its shared-memory tables by (ep_solve_quads.cuh).
__device__ __forceinline__ int island_bin(int rows, int groups) {
  int b = 0;
  while (b < kBins - 1 && (rows > (3 << b) || groups > (1 << b))) b++;
  return b;
}
// Sort key inside a size class, most significant first: row count (16 sub-classes of the class's range: the lanes of a tile
// then run about the same number of rows), iterations the island used last frame (8 classes), group count (fewer body-pair
// switches diverge).  Larger = runs longer = scheduled first.
__device__ __forceinline__ int island_key(int bin, int rows, int pred_iters, int groups) {
  const int lo = bin == 0 ? 0 : (3 << (bin - 1)), hi = 3 << bin;
  const int sz = bin >= kBins - 1 ? 15 : min(15, max(0, (rows - lo - 1) * 16 / max(1, hi - lo)));
  const int it = min(7, max(0, pred_iters) / 3);
  const int gk = groups <= 1 ? 0 : (groups == 2 ? 1 : (groups <= 4 ? 2 : 3));
  return ((bin * 16 + sz) * 8 + it) * kGroupKeys + gk;
}
// One CTA per world (one warp for small worlds).  Union-find with lock-free hooking: the root of a set is always its
// member with the LOWEST body id (Islands.elm:39-46 "lower root wins"), so the final roots do not depend on the order of
// the unions and the parallel version yields exactly the reference's roots.  Islands of a world are numbered by ascending
// body row of the root; the groups of an island stay in pair-enumeration order (Islands.elm:51-119).
__device__ __forceinline__ int uf_root(volatile int* parent, int x) {
  for (;;) { int p = parent[x]; if (p == x) return x; x = p; }
}
// The same with path halving, for the union phase only: a non-root never becomes a root again and the CAS below only ever
// changes roots, so pointing x at its grandparent (still an ancestor) races with nothing that matters; the component's root stays
// its minimum id.  Without it the one giant island of config 4 (1900 bodies linked by minimum id) grows chains hundreds of
// hops deep and every find walks them (k_islands 0.36 ms, most of it here).  NOT for the flatten pass, whose result must be a root.
__device__ __forceinline__ int uf_root_halving(volatile int* parent, int x) {
  for (;;) {
    const int p = parent[x]; if (p == x) return x;
    const int gp = parent[p]; if (gp == p) return p;
    parent[x] = gp; x = gp;
  }
}
__global__ void k_islands(Dev d, int use_smem) {
  extern __shared__ int s_isl[];           // [3][nb_max] parent / count / fill when the largest world fits
  __shared__ int s_scan[33];
  __shared__ int s_nisl, s_valid, s_ibase, s_gbase;
  constexpr int kWalk = 512;
  __shared__ int s_wroot[kWalk], s_wcc[kWalk], s_wjc[kWalk], s_wslot[kWalk];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarps = nt >> 5;
  for (int w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
    const int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    int* parent = use_smem ? s_isl : d.uf_parent + r0;
    int* count = use_smem ? s_isl + d.nb_max : d.uf_count + r0;
    int* fill = use_smem ? s_isl + 2 * d.nb_max : d.uf_fill + r0;
    const int s0 = d.cur.cand_base[w], sn = d.cur.cand_n[w];
    for (int l = tid; l < n; l += nt) { parent[l] = l; count[l] = 0; fill[l] = 0; d.body_local[r0 + l] = 0; }
    if (tid == 0) { s_nisl = 0; s_valid = 0; }
    __syncthreads();
    // Islands.connect for every dynamic-dynamic pair group (Solver.elm:201-206)
    for (int s = s0 + tid; s < s0 + sn; s += nt) {
      if (d.cur.slot_ccount[s] <= 0 && d.cur.slot_jcount[s] <= 0) continue;
      const int ra = d.cur.slot_row1[s], rb = d.cur.slot_row2[s];
      if (d.kind[ra] != 2 || d.kind[rb] != 2) continue;
      int a = ra - r0, b = rb - r0;
      for (;;) {
        a = uf_root_halving(parent, a); b = uf_root_halving(parent, b);
        if (a == b) break;
        const int ia = d.body_id[r0 + a], ib = d.body_id[r0 + b];
        const int win = (ia - ib < 0) ? a : b, lose = (ia - ib < 0) ? b : a;
        if (atomicCAS(&parent[lose], lose, win) == lose) break;
      }
    }
    __syncthreads();
    for (int l = tid; l < n; l += nt) parent[l] = uf_root(parent, l);      // flatten (reads may see already-flattened entries: still roots)
    __syncthreads();
    // annotate (Islands.elm:68-89): island of a group = root of its dynamic body; groups / rows / contacts / joint rows per root
    int valid = 0;
    for (int s = s0 + tid; s < s0 + sn; s += nt) {
      if (d.cur.slot_ccount[s] <= 0 && d.cur.slot_jcount[s] <= 0) { d.cur.slot_root[s] = -1; continue; }
      const int ra = d.cur.slot_row1[s], rb = d.cur.slot_row2[s];
      const int root = parent[(d.kind[ra] == 2 ? ra : rb) - r0];
      d.cur.slot_root[s] = r0 + root;
      atomicAdd(&count[root], 1);
      valid++;
    }
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    if (lane == 0 && valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    // number the islands by ascending root row: block-wide exclusive scan of (count > 0)
    int run = 0;
    for (int base = 0; base < n; base += nt) {
      const int l = base + tid;
      const int flag = (l < n && count[l] > 0) ? 1 : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      if (lane == 0) s_scan[wid] = __popc(bal);
      __syncthreads();
      int before = run;
      for (int k = 0; k < wid; k++) before += s_scan[k];
      int total = 0;
      for (int k = 0; k < nwarps; k++) total += s_scan[k];
      if (flag) fill[l] = before + __popc(bal & ((1u << lane) - 1u));       // island number inside the world (for now)
      run += total;
      __syncthreads();
    }
    if (tid == 0) {
      s_nisl = run;
      d.world_n_islands[w] = run;
      s_ibase = run ? atomicAdd(&d.cur.counters[CNT_ISLANDS], run) : 0;
      s_gbase = run ? atomicAdd(&d.cur.counters[CNT_ISL_GROUPS], s_valid) : 0;
    }
    __syncthreads();
    if (s_nisl == 0) { __syncthreads(); continue; }
    const int ibase = s_ibase, gbase = s_gbase;
    // group offsets of the islands: exclusive scan of count[] over the roots in island order.  Islands are few per
    // world in the batched configs and the scan is over bodies, so a single warp does it.
    if (wid == 0) {
      int goff = gbase;
      for (int base = 0; base < n; base += 32) {
        const int l = base + lane;
        const int c = (l < n) ? count[l] : 0;
        int sc = c;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += v; }
        if (c > 0) {
          const int isl = ibase + fill[l];
          d.isl_off[isl] = goff + sc - c; d.isl_cnt[isl] = c; d.isl_world[isl] = w; d.isl_root[isl] = r0 + l;
          d.isl_nc[isl] = 0; d.isl_nj[isl] = 0; d.isl_nb[isl] = 0;
          count[l] = isl + 1;                       // root -> island index + 1 from here on
          fill[l] = goff + sc - c;                  // next free group slot of the island
        }
        goff += __shfl_sync(0xffffffffu, sc, 31);
      }
    }
    __syncthreads();
    // groups of an island in enumeration order: one warp walks the slots 32 at a time, lanes with the same root rank
    // themselves with match_any; contacts / joint rows accumulate per island.  The walk is serial, so what it reads is staged
    // in shared memory by the whole CTA, kWalk slots at a time (a large world -- config 4: 20 k candidate slots -- spent
    // 0.5 ms here in 630 dependent L2 round trips of one warp)
    // Only the slots that carry a group are staged (ordered ballot compaction: most candidate pairs of a large world end without
    // contacts -- config 4: 5.3 k groups in 20 k slots -- and the one walking warp paid ~600 cycles per 32 slots, empty or not).
    for (int cb = 0; cb < sn; cb += kWalk) {
      const int cm = min(kWalk, sn - cb);
      int filled = 0;
      for (int i0 = 0; i0 < cm; i0 += nt) {
        const int i = i0 + tid, s = s0 + cb + i;
        const int root = i < cm ? d.cur.slot_root[s] : -1;
        const unsigned bal = __ballot_sync(0xffffffffu, root >= 0);
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int before = filled, total = 0;
        for (int k = 0; k < nwarps; k++) { const int c = s_scan[k]; if (k < wid) before += c; total += c; }
        if (root >= 0) {
          const int at = before + __popc(bal & ((1u << lane) - 1u));
          s_wroot[at] = root; s_wslot[at] = s; s_wcc[at] = max(d.cur.slot_ccount[s], 0); s_wjc[at] = max(d.cur.slot_jcount[s], 0);
        }
        filled += total;
        __syncthreads();
      }
      if (wid == 0) {
        for (int base = 0; base < filled; base += 32) {
          const int s = (base + lane < filled) ? s_wslot[base + lane] : 0;
          const int root = (base + lane < filled) ? s_wroot[base + lane] : -1;
          const unsigned peers = __match_any_sync(0xffffffffu, root);
          if (root >= 0) {
            const int l = root - r0;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            const int at = fill[l] + rank;
            d.isl_groups[at] = s;
            const int isl = count[l] - 1;
            if (s_wcc[base + lane]) atomicAdd(&d.isl_nc[isl], s_wcc[base + lane]);
            if (s_wjc[base + lane]) atomicAdd(&d.isl_nj[isl], s_wjc[base + lane]);
            __syncwarp(peers);
            if (lane == leader) fill[l] += __popc(peers);
          }
          __syncwarp();
        }
      }
      __syncthreads();
    }
    // dynamic bodies of an island = its union-find component
    for (int l = tid; l < n; l += nt) {
      if (d.kind[r0 + l] != 2) continue;
      const int c = count[parent[l]];
      if (c > 0) d.body_local[r0 + l] = 1 + atomicAdd(&d.isl_nb[c - 1], 1);      // any order: the index only names the body inside its island
    }
    __syncthreads();
    // size class and sort key of each island (k_island_sort)
    for (int l = tid; l < n; l += nt) {
      if (parent[l] != l || count[l] <= 0) continue;
      const int isl = count[l] - 1;
      const int rows = 3 * d.isl_nc[isl] + d.isl_nj[isl];
      const int bin = island_bin(rows, d.isl_cnt[isl]);
      const int key = island_key(bin, rows, d.body_iters[r0 + l], d.isl_cnt[isl]);
      d.isl_key[isl] = key;
      atomicAdd(&d.key_cnt[key], 1);
      atomicAdd(&d.cur.counters[CNT_BIN0 + bin], 1);
    }
    __syncthreads();
  }
}

// Work-list order: every block recomputes the (tiny) exclusive prefix of the key histogram in DESCENDING key order, then
// scatters its share of the islands.  The order inside one key is arbitrary and never influences results.
__global__ void __launch_bounds__(256) k_island_sort(Dev d) {
  __shared__ int s_start[kKeys];
  __shared__ int s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kPer = (kKeys + 255) / 256;
  int local[kPer], sum = 0;
  for (int j = 0; j < kPer; j++) {
    int pos = tid * kPer + j;                       // position in descending key order
    local[j] = pos < kKeys ? d.key_cnt[kKeys - 1 - pos] : 0;
    sum += local[j];
  }
  int scan = sum;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, scan, o); if (lane >= o) scan += v; }
  if (lane == 31) s_warp[wid] = scan;
  __syncthreads();
  int before = 0;
  for (int k2 = 0; k2 < wid; k2++) before += s_warp[k2];
  int run = before + scan - sum;
  for (int j = 0; j < kPer; j++) {
    int pos = tid * kPer + j;
    if (pos < kKeys) s_start[kKeys - 1 - pos] = run;
    run += local[j];
  }
  __syncthreads();
  const int n = d.cur.counters[CNT_ISLANDS];
  for (int isl = blockIdx.x * blockDim.x + tid; isl < n; isl += gridDim.x * blockDim.x) {
    int key = d.isl_key[isl];
    d.isl_sorted[s_start[key] + atomicAdd(&d.key_cnt[kKeys + key], 1)] = isl;
  }
}

// ---- solver-body deltas and the row dot product shared by the island solvers (ep_solve.cuh, ep_solve_quads.cuh) ------
struct SB { double vX, vY, vZ, wX, wY, wZ; };
__device__ __forceinline__ SB load_sb(const double* p) { SB s; s.vX = p[0]; s.vY = p[1]; s.vZ = p[2]; s.wX = p[3]; s.wY = p[4]; s.wZ = p[5]; return s; }
__device__ __forceinline__ void store_sb(double* p, const SB& s) { p[0] = s.vX; p[1] = s.vY; p[2] = s.vZ; p[3] = s.wX; p[4] = s.wY; p[5] = s.wZ; }
__device__ __forceinline__ double gw_lambda(const double* J, const SB& b1, const SB& b2) {
  return -(J[3] * b1.vX + J[4] * b1.vY + J[5] * b1.vZ)
         + (J[0] * b1.wX + J[1] * b1.wY + J[2] * b1.wZ)
         + (J[3] * b2.vX + J[4] * b2.vY + J[5] * b2.vZ)
         + (J[6] * b2.wX + J[7] * b2.wY + J[8] * b2.wZ);
}
__device__ __forceinline__ int class_start(const Dev& d, int bin) {
  int at = 0;
  for (int b = kBins - 1; b > bin; b--) at += d.cur.counters[CNT_BIN0 + b];
  return at;
}
__device__ __forceinline__ int island_from_worklist(const Dev& d, int t, int bin_lo, int bin_hi) {
  return d.isl_sorted[class_start(d, bin_hi) + t];
}
// ---- k_integrate : one thread per body (SolverBody.elm:97-225) ---------------------------------------------
__global__ void k_integrate(Dev d) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < d.n_bodies; r += gridDim.x * blockDim.x) {
    int kind = d.kind[r];
    double* sb = d.sbv + 6 * r;
    SB s = load_sb(sb);
    sb[0] = 0; sb[1] = 0; sb[2] = 0; sb[3] = 0; sb[4] = 0; sb[5] = 0;
    V3 x = ld3(d.pos + 3 * r);
    Q4 q; q.x = d.quat[4 * r]; q.y = d.quat[4 * r + 1]; q.z = d.quat[4 * r + 2]; q.w = d.quat[4 * r + 3];
    st3(d.pre_pos + 3 * r, x);
    d.pre_quat[4 * r] = q.x; d.pre_quat[4 * r + 1] = q.y; d.pre_quat[4 * r + 2] = q.z; d.pre_quat[4 * r + 3] = q.w;
    if (kind == 1) continue;
    int w = d.body_world[r];
    double dt = d.dt[w];
    V3 v = ld3(d.vel + 3 * r), om = ld3(d.angvel + 3 * r);
    V3 step_v, step_w;
    if (kind == 3) { step_v = v; step_w = om; }
    else {
      V3 g = ld3(d.gravity + 3 * w), f = ld3(d.force + 3 * r), tq = ld3(d.torque + 3 * r);
      double im = d.inv_mass[r], ld = d.ldp[r], ad = d.adp[r];
      V3 ll = ld3(d.lin_lock + 3 * r), al = ld3(d.ang_lock + 3 * r);
      const double* m = d.iiw + 9 * r;   // m11,m21,m31,m12,m22,m32,m13,m23,m33
      V3 nv = mk(((g.x + f.x * im) * dt + v.x * ld + s.vX) * ll.x, ((g.y + f.y * im) * dt + v.y * ld + s.vY) * ll.y, ((g.z + f.z * im) * dt + v.z * ld + s.vZ) * ll.z);
      double vl = vlen(nv), br = d.radius[r];
      step_v = ((vl == 0) || (br == 0) || (vl * dt - br < 0)) ? nv : vscale(br / (vl * dt), nv);
      V3 nw = mk(((m[0] * tq.x + m[3] * tq.y + m[6] * tq.z) * dt + om.x * ad + s.wX) * al.x,
                 ((m[1] * tq.x + m[4] * tq.y + m[7] * tq.z) * dt + om.y * ad + s.wY) * al.y,
                 ((m[2] * tq.x + m[5] * tq.y + m[8] * tq.z) * dt + om.z * ad + s.wZ) * al.z);
      step_w = nw;
      st3(d.vel + 3 * r, nv); st3(d.angvel + 3 * r, nw);
    }
    // rotateBy (Transform3d.elm:368-376), translateBy, normalize (256-262)
    V3 a = mk(step_w.x * dt, step_w.y * dt, step_w.z * dt);
    double nqx = q.x + (a.x * q.w + a.y * q.z - a.z * q.y) * 0.5;
    double nqy = q.y + (a.y * q.w + a.z * q.x - a.x * q.z) * 0.5;
    double nqz = q.z + (a.z * q.w + a.x * q.y - a.y * q.x) * 0.5;
    double nqw = q.w + (-a.x * q.x - a.y * q.y - a.z * q.z) * 0.5;
    V3 nx = vadd(mk(step_v.x * dt, step_v.y * dt, step_v.z * dt), x);
    double len = sqrt(nqx * nqx + nqy * nqy + nqz * nqz + nqw * nqw);
    Q4 nq; nq.x = nqx / len; nq.y = nqy / len; nq.z = nqz / len; nq.w = nqw / len;
    st3(d.pos + 3 * r, nx);
    d.quat[4 * r] = nq.x; d.quat[4 * r + 1] = nq.y; d.quat[4 * r + 2] = nq.z; d.quat[4 * r + 3] = nq.w;
    if (kind == 2) {
      M33 I = inv_inertia_world(nq, ld3(d.inv_inertia + 3 * r));
      double* m = d.iiw + 9 * r;
      m[0] = I.m11; m[1] = I.m21; m[2] = I.m31; m[3] = I.m12; m[4] = I.m22; m[5] = I.m32; m[6] = I.m13; m[7] = I.m23; m[8] = I.m33;
    }
    d.force[3 * r] = 0; d.force[3 * r + 1] = 0; d.force[3 * r + 2] = 0;
    d.torque[3 * r] = 0; d.torque[3 * r + 1] = 0; d.torque[3 * r + 2] = 0;
  }
}

// warm-start upload helper: (re)build the pair table of a frame from its slots
__global__ void k_pair_slots(Dev d, FrameBuf fb, int n_slots) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    if (fb.slot_ccount[s] <= 0) continue;
    int r1 = fb.slot_row1[s], r2 = fb.slot_row2[s];
    if (r1 == r2) continue;
    fb.pair_slot[pair_index(d, d.body_world[r1], r1, r2)] = s;
  }
}

}  // namespace ep
