// Spook / Projected Gauss-Seidel island solver (src/Internal/Solver.elm:345-880).
//
// Islands are independent (Islands.elm:3-7); inside one the reference's order is kept exactly: per iteration sweep 1
// (joint rows, then contact normals) over the groups in pair-enumeration order, then sweep 2 (friction1, friction2 per
// contact); warm start in reverse group order; sum|dlambda| association of `step` (two sums added) vs `solve2Body` (one
// running sum).  Two mappings, chosen by island size class (island_bin, ep_kernels.cuh):
//
//   classes 0..kLaneBinMax  one LANE per island, one WARP per tile of 32 islands (k_tile_build + k_solve_lanes).  The work
//       list is sorted by (class, iterations the island used last frame, group count), so the 32 lanes of a tile run about
//       the same rows for about the same number of iterations.  k_tile_build gathers each tile ONCE into a lane-transposed
//       image (word k of lane l at [k][l]: every access of the warp is one coalesced 256 B request / conflict-free LDS.64):
//       Jacobian rows, Spook scalars, lambdas, the islands' dynamic bodies (constants + solver-body delta velocities) and
//       per-row descriptors (which two bodies a row couples).  k_solve_lanes<2> stages the whole image in shared memory
//       (small classes), k_solve_lanes<1> stages bodies + descriptors and streams the rows from L2 one row ahead of their
//       use; HBM sees each row once per step instead of once per iteration.
//   classes above           one CTA per island, the pair groups of one dependency level in parallel (k_solve_team).
#pragma once
#include "ep_kernels.cuh"

namespace ep {

constexpr int kLaneBinMax = 5;   // islands of up to 32 contacts (96 rows) run lane-per-island
// Tile image of 32 islands, every array lane-transposed ([k][lane]); in this order:
//   bodies  [maxB][16] f64   invMass, I (m11 m21 m31 m12 ...), solver-body delta v/w.  Local body 0 is the shared stand-in of
//                            every non-dynamic body: their solver bodies are zero and never modified (Solver.elm:507,531,796,815)
//   ints    [maxB] body row | [maxR] sweep-1 descriptors | [maxC] sweep-2 descriptors | [maxC] global contact index
//   nrows   [maxR][12] f64   sweep-1 rows in solver order (per group: joint rows, then contact normals): J[9], B, invC, lambda
//   frows   [maxC][25] f64   sweep-2 rows: f1J[9], f2J[9], f1B, f1InvC, f2B, f2InvC, mu, lambdaF1, lambdaF2
// descriptor = local body1 | local body2 << 8 | (sweep 1: joint flag << 16, first-row-of-group flag << 17)
//                                              | (sweep 2: index of the contact's normal row << 16)
constexpr int kBodyF64 = 16, kNRowF64 = 12, kFRowF64 = 25;
enum { DESC_JOINT = 1 << 16, DESC_GROUP_START = 1 << 17 };
struct TileDims { int maxB, maxR, maxC; };
__host__ __device__ inline int64_t tile_head_words(TileDims t) { return 32 * (int64_t)(kBodyF64 * t.maxB) + 16 * (int64_t)(t.maxB + t.maxR + 2 * t.maxC); }
__host__ __device__ inline int64_t tile_words(TileDims t) { return tile_head_words(t) + 32 * (int64_t)(kNRowF64 * t.maxR + kFRowF64 * t.maxC); }   // 8-byte words

struct Tile {
  double* bd; int* ti; double* rd; int lane; TileDims t;     // bodies, ints, rows (each may live in shared or global memory)
  __device__ __forceinline__ double& body(int b, int f) const { return bd[(b * kBodyF64 + f) * 32 + lane]; }
  __device__ __forceinline__ int& body_row(int b) const { return ti[b * 32 + lane]; }
  __device__ __forceinline__ int& desc1(int r) const { return ti[(t.maxB + r) * 32 + lane]; }
  __device__ __forceinline__ int& desc2(int c) const { return ti[(t.maxB + t.maxR + c) * 32 + lane]; }
  __device__ __forceinline__ int& contact_index(int c) const { return ti[(t.maxB + t.maxR + t.maxC + c) * 32 + lane]; }
  __device__ __forceinline__ double& nrow(int r, int f) const { return rd[(r * kNRowF64 + f) * 32 + lane]; }
  __device__ __forceinline__ double& frow(int c, int f) const { return rd[(t.maxR * kNRowF64 + c * kFRowF64 + f) * 32 + lane]; }
};
__device__ __forceinline__ void tile_bind(Tile& T, double* head, double* rows) {
  T.bd = head; T.ti = (int*)(head + 32 * (int64_t)(kBodyF64 * T.t.maxB)); T.rd = rows;
}

// Tiles of the lane classes are numbered class kLaneBinMax first; tile t of class b holds sorted islands
// [class_start(b) + 32 t, ...).  Returns false past the last tile of class 0.
struct TileRef { int bin, first, count; };     // first sorted index, islands in the tile (1..32)
__device__ __forceinline__ bool lane_tile(const Dev& d, int gtile, TileRef& r) {
  int base = 0;
  for (int b = kBins - 1; b > kLaneBinMax; b--) base += d.cur.counters[CNT_BIN0 + b];
  for (int b = kLaneBinMax; b >= 0; b--) {
    int cnt = d.cur.counters[CNT_BIN0 + b], nt = (cnt + 31) >> 5;
    if (gtile < nt) { r.bin = b; r.first = base + gtile * 32; r.count = min(32, cnt - gtile * 32); return true; }
    gtile -= nt; base += cnt;
  }
  return false;
}
__device__ __forceinline__ int warp_max(int v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- k_tile_build : one warp per tile, one lane per island ---------------------------------------------------------------
// Residency mode of a tile, decided here from the dynamic shared memory of the launches (Dev::tile_smem): 2 = whole image
// staged in shared memory, 1 = bodies + descriptors staged and the rows streamed from L2 one row ahead, 0 = in place.
__global__ void __launch_bounds__(128) k_tile_build(Dev d) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int gtile = warp;; gtile += nwarps) {
    TileRef tr;
    if (!lane_tile(d, gtile, tr)) break;
    const int isl = lane < tr.count ? d.isl_sorted[tr.first + lane] : -1;
    const int g0 = isl >= 0 ? d.isl_off[isl] : 0, gn = isl >= 0 ? d.isl_cnt[isl] : 0;
    const int nc = isl >= 0 ? d.isl_nc[isl] : 0, nj = isl >= 0 ? d.isl_nj[isl] : 0;
    TileDims td;
    td.maxB = warp_max(isl >= 0 ? d.isl_nb[isl] : 0) + 1;
    td.maxR = warp_max(nc + nj);
    td.maxC = warp_max(nc);
    const int64_t words = tile_words(td), head = tile_head_words(td);
    int mode = 0;
    if (td.maxB < 256 && td.maxR < 256) {
      if (tr.bin <= 2) mode = words * 8 <= d.tile_smem[tr.bin] ? 2 : (head * 8 <= d.tile_smem[3] ? 1 : 0);
      else mode = head * 8 <= d.tile_smem[tr.bin] ? 1 : 0;
    } else td.maxC = -1;                      // cannot happen for classes <= kLaneBinMax (rows <= 96); refuse rather than corrupt
    unsigned long long off = 0;
    if (lane == 0) {
      off = atomicAdd(d.tile_top, (unsigned long long)words);
      if ((int64_t)(off + words) > d.tile_arena_cap || td.maxC < 0) { d.flags[FLAG_POOL_OVERFLOW] = 2; td.maxC = -1; off = 0; }
      d.tile_dir[gtile] = make_int4((int)(off & 0xffffffffu), (int)(off >> 32), td.maxB | (td.maxR << 16), (td.maxC & 0xfffffff) | (mode << 28));
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    td.maxC = __shfl_sync(0xffffffffu, td.maxC, 0);
    if (isl < 0 || td.maxC < 0) continue;
    Tile T; T.t = td; T.lane = lane;
    tile_bind(T, d.tile_arena + off, d.tile_arena + off + head);
    // local body table: 0 = the non-dynamic stand-in, dynamic bodies 1.. in order of first appearance
    T.body_row(0) = -1;
#pragma unroll
    for (int k = 0; k < kBodyF64; k++) T.body(0, k) = 0.0;
    int nB = 1, r = 0, c = 0;
    for (int gi = 0; gi < gn; gi++) {
      const int s = d.isl_groups[g0 + gi];
      const int rr[2] = {d.cur.slot_row1[s], d.cur.slot_row2[s]};
      int lb[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int br = rr[e];
        int m = 0;
        if (d.kind[br] == 2) {
          m = d.body_mark[br];
          if (m == 0) {
            m = nB++; d.body_mark[br] = m; T.body_row(m) = br;
            T.body(m, 0) = d.inv_mass[br];
            const double* im = d.iiw + 9 * (size_t)br;
#pragma unroll
            for (int k = 0; k < 9; k++) T.body(m, 1 + k) = im[k];
            const double* sb = d.sbv + 6 * (size_t)br;
#pragma unroll
            for (int k = 0; k < 6; k++) T.body(m, 10 + k) = sb[k];
          }
        }
        lb[e] = m;
      }
      const int pair = lb[0] | (lb[1] << 8);
      int first = DESC_GROUP_START;
      const int j0 = d.cur.slot_jstart[s], jn = d.cur.slot_jcount[s];
      for (int k = 0; k < jn; k++, r++) {
        const JointRow& jr = d.jrows[j0 + k];
#pragma unroll
        for (int f = 0; f < 9; f++) T.nrow(r, f) = jr.J[f];
        T.nrow(r, 9) = jr.B; T.nrow(r, 10) = jr.invC; T.nrow(r, 11) = jr.lam;
        T.desc1(r) = pair | DESC_JOINT | first; first = 0;
      }
      const int c0 = d.cur.slot_cstart[s], cn = max(d.cur.slot_ccount[s], 0);
      for (int k = 0; k < cn; k++, r++, c++) {
        const double2* src = (const double2*)(d.cur.rows + c0 + k);     // RowRec is 64 B aligned
        double2 v[18];
#pragma unroll
        for (int f = 0; f < 18; f++) v[f] = src[f];
        const double last = ((const double*)src)[36];
#pragma unroll
        for (int f = 0; f < 6; f++) { T.nrow(r, 2 * f) = v[f].x; T.nrow(r, 2 * f + 1) = v[f].y; }
#pragma unroll
        for (int f = 6; f < 18; f++) { T.frow(c, 2 * f - 12) = v[f].x; T.frow(c, 2 * f - 11) = v[f].y; }
        T.frow(c, 24) = last;
        T.desc1(r) = pair | first; first = 0;
        T.desc2(c) = pair | (r << 16);
        T.contact_index(c) = c0 + k;
      }
    }
  }
}

// Rows of the tile column of this lane, loaded into registers one row AHEAD of their use so that the (shared-memory or
// L2) latency of row c+1 hides behind the dependent FP64 chain of row c.
struct RowN { double J[9], B, invC, lam; int desc; };
struct RowF { double e1[9], e2[9], f1B, f1InvC, f2B, f2InvC, mu, l1, l2; int desc; };
__device__ __forceinline__ void load_row_n(const Tile& T, int r, RowN& o) {
#pragma unroll
  for (int k = 0; k < 9; k++) o.J[k] = T.nrow(r, k);
  o.B = T.nrow(r, 9); o.invC = T.nrow(r, 10); o.lam = T.nrow(r, 11); o.desc = T.desc1(r);
}
__device__ __forceinline__ void load_row_f(const Tile& T, int c, RowF& o) {
#pragma unroll
  for (int k = 0; k < 9; k++) { o.e1[k] = T.frow(c, k); o.e2[k] = T.frow(c, 9 + k); }
  o.f1B = T.frow(c, 18); o.f1InvC = T.frow(c, 19); o.f2B = T.frow(c, 20); o.f2InvC = T.frow(c, 21);
  o.mu = T.frow(c, 22); o.l1 = T.frow(c, 23); o.l2 = T.frow(c, 24); o.desc = T.desc2(c);
}
// The two solver bodies of the current row, held in registers across consecutive rows that share them (pairs are
// enumerated body1-major, so body1 usually stays).  Local body 0 never touches memory.
struct Pair { int b1, b2; BodyK k1, k2; SB s1, s2; };
__device__ __forceinline__ void tile_store_sb(const Tile& T, int b, const SB& s) {
  T.body(b, 10) = s.vX; T.body(b, 11) = s.vY; T.body(b, 12) = s.vZ; T.body(b, 13) = s.wX; T.body(b, 14) = s.wY; T.body(b, 15) = s.wZ;
}
__device__ __forceinline__ void pair_load_body(const Tile& T, int b, BodyK& k, SB& s) {
  if (b == 0) {
    k.kind = 1; k.invMass = 0; k.I.m11 = k.I.m21 = k.I.m31 = k.I.m12 = k.I.m22 = k.I.m32 = k.I.m13 = k.I.m23 = k.I.m33 = 0;
    s.vX = s.vY = s.vZ = s.wX = s.wY = s.wZ = 0;
  } else {
    k.kind = 2; k.invMass = T.body(b, 0);
    k.I.m11 = T.body(b, 1); k.I.m21 = T.body(b, 2); k.I.m31 = T.body(b, 3); k.I.m12 = T.body(b, 4); k.I.m22 = T.body(b, 5);
    k.I.m32 = T.body(b, 6); k.I.m13 = T.body(b, 7); k.I.m23 = T.body(b, 8); k.I.m33 = T.body(b, 9);
    s.vX = T.body(b, 10); s.vY = T.body(b, 11); s.vZ = T.body(b, 12); s.wX = T.body(b, 13); s.wY = T.body(b, 14); s.wZ = T.body(b, 15);
  }
}
__device__ __forceinline__ void pair_enter(const Tile& T, Pair& P, int desc) {
  const int nb1 = desc & 0xff, nb2 = (desc >> 8) & 0xff;
  if (nb1 == P.b1 && nb2 == P.b2) return;
  // a body that changes sides goes through the tile: stores first, loads after
  if (nb1 != P.b1 && P.b1 > 0) tile_store_sb(T, P.b1, P.s1);
  if (nb2 != P.b2 && P.b2 > 0) tile_store_sb(T, P.b2, P.s2);
  if (nb1 != P.b1) { pair_load_body(T, nb1, P.k1, P.s1); P.b1 = nb1; }
  if (nb2 != P.b2) { pair_load_body(T, nb2, P.k2, P.s2); P.b2 = nb2; }
}
__device__ __forceinline__ void pair_flush(const Tile& T, Pair& P) {
  if (P.b1 > 0) tile_store_sb(T, P.b1, P.s1);
  if (P.b2 > 0) tile_store_sb(T, P.b2, P.s2);
  P.b1 = -1; P.b2 = -1;
}

// The whole solve of the island of this lane on its tile column, then the scatter of the results.  Each sweep walks ONE
// flat row sequence (the lanes of the warp stay in lock step on the row index whatever their group structure is);
// changing the body pair is the only divergent section.
__device__ __forceinline__ void solve_lane(const Dev& d, const Tile& T, int isl) {
  const int gn = d.isl_cnt[isl], w = d.isl_world[isl], C = d.isl_nc[isl], R = C + d.isl_nj[isl];
  Pair P; P.b1 = -1; P.b2 = -1;
  // warm start (applyContactsWarmStart, Solver.elm:102-119): groups in REVERSE enumeration order (buildAndWarmStart walks
  // pairGroups, Solver.elm:122-208), the seeded normal impulses of a group in solver order
  for (int hi = R - 1; hi >= 0;) {
    int lo = hi;
    while (!(T.desc1(lo) & DESC_GROUP_START)) lo--;
    for (int r = lo; r <= hi; r++) {
      const int desc = T.desc1(r);
      if (desc & DESC_JOINT) continue;
      const double lam = T.nrow(r, 11);
      if (lam == 0) continue;                        // Solver.elm:41
      pair_enter(T, P, desc);
      double J[9];
#pragma unroll
      for (int k = 0; k < 9; k++) J[k] = T.nrow(r, k);
      apply_row(lam, J, P.k1, P.k2, P.s1, P.s2);
    }
    hi = lo - 1;
  }
  pair_flush(T, P);
  const int iterations = d.iterations[w];
  int remaining = iterations, rem;
  const double eps = spook(d.dt[w]).eps;
  for (;;) {
    // sweep 1: velocityNonFrictionGroup (Solver.elm:850-864) group by group: joint rows (572-619), then contact normals (625-672)
    double p1 = 0;
    {
      RowN nxt;
      if (R > 0) load_row_n(T, 0, nxt);
      for (int r = 0; r < R; r++) {
        RowN cur = nxt;
        if (r + 1 < R) load_row_n(T, r + 1, nxt);
        pair_enter(T, P, cur.desc);
        const double lo = (cur.desc & DESC_JOINT) ? -kMaxImpulse : 0.0;
        double dl = row_solve(cur.J, cur.B, cur.invC, eps, lo, kMaxImpulse, cur.lam, P.k1, P.k2, P.s1, P.s2);
        T.nrow(r, 11) = cur.lam;
        p1 = p1 + eabs(dl);
      }
      pair_flush(T, P);
    }
    // sweep 2: velocityFrictionGroup (Solver.elm:869-880); solve2Body keeps ONE running sum through both phases
    // (Solver.elm:473-489), step adds two sums (345-358)
    double p2 = (gn == 1) ? p1 : 0;
    {
      RowF nxt;
      if (C > 0) load_row_f(T, 0, nxt);
      for (int c = 0; c < C; c++) {
        RowF cur = nxt;
        if (c + 1 < C) load_row_f(T, c + 1, nxt);
        const double lamN = T.nrow((cur.desc >> 16) & 0xff, 11);
        pair_enter(T, P, cur.desc);
        double a1, a2;
        p2 = friction_pair(cur.e1, cur.e2, cur.f1B, cur.f1InvC, cur.f2B, cur.f2InvC, eps, cur.mu, lamN, cur.l1, cur.l2,
                           P.k1, P.k2, P.s1, P.s2, p2, a1, a2);
        T.frow(c, 23) = cur.l1; T.frow(c, 24) = cur.l2;
      }
      pair_flush(T, P);
    }
    double deltaTot = (gn == 1) ? p2 : p1 + p2;
    if (remaining == 1) { rem = 0; break; }
    if (deltaTot - kSolverTolerance < 0) { rem = remaining - 1; break; }
    remaining--;
  }
  atomicMin(&d.world_min_rem[w], rem);
  d.body_iters[d.isl_root[isl]] = iterations - rem;
  // scatter: lambdas back to the slot-order rows (next frame's warm start, Solver.elm:320-335), delta velocities to the bodies
  for (int c = 0; c < C; c++) {
    RowRec& o = d.cur.rows[T.contact_index(c)];
    o.lamN = T.nrow((T.desc2(c) >> 16) & 0xff, 11); o.lamF1 = T.frow(c, 23); o.lamF2 = T.frow(c, 24);
  }
  const int nB = d.isl_nb[isl] + 1;
  for (int b = 1; b < nB; b++) {
    double* sb = d.sbv + 6 * (size_t)T.body_row(b);
#pragma unroll
    for (int k = 0; k < 6; k++) sb[k] = T.body(b, 10 + k);
  }
}

// One warp per tile.  MODE 2: whole image staged in dynamic shared memory; MODE 1: bodies + descriptors staged, rows read
// from the image in place (L2) one row ahead; MODE 0: everything in place.  Processes the tiles of classes [bin_lo, bin_hi]
// whose residency mode (k_tile_build) is MODE.
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? 128 : 32) k_solve_lanes(Dev d, int bin_lo, int bin_hi) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  int t0 = 0;                     // first global tile of class bin_hi
  for (int b = kLaneBinMax; b > bin_hi; b--) t0 += (d.cur.counters[CNT_BIN0 + b] + 31) >> 5;
  int nt = 0;
  for (int b = bin_hi; b >= bin_lo; b--) nt += (d.cur.counters[CNT_BIN0 + b] + 31) >> 5;
  for (int t = warp; t < nt; t += nwarps) {
    const int gtile = t0 + t;
    const int4 e = d.tile_dir[gtile];
    const int mode = (e.w >> 28) & 3;
    int maxC = e.w & 0xfffffff; if (maxC & 0x8000000) maxC = -1;
    if (mode != MODE || maxC < 0) continue;
    TileRef tr; lane_tile(d, gtile, tr);
    Tile T; T.lane = lane; T.t.maxB = e.z & 0xffff; T.t.maxR = (e.z >> 16) & 0xffff; T.t.maxC = maxC;
    double* img = d.tile_arena + (((unsigned long long)(unsigned)e.y << 32) | (unsigned)e.x);
    const int head = (int)tile_head_words(T.t);
    if (MODE == 0) {
      tile_bind(T, img, img + head);
    } else {
      double* sm = (double*)tile_smem;
      const int n = MODE == 2 ? (int)tile_words(T.t) : head;
      __syncwarp();
      for (int i = lane; i < n; i += 32) sm[i] = img[i];
      __syncwarp();
      tile_bind(T, sm, MODE == 2 ? sm + head : img + head);
    }
    if (lane < tr.count) solve_lane(d, T, d.isl_sorted[tr.first + lane]);
  }
}

// ---- k_solve_team : one CTA per island, the pair groups of a dependency level in parallel ------------------------------
// PGS is sequential inside an island, but two pair groups that share no DYNAMIC body commute exactly: a group reads and
// writes only the solver-body deltas of its own two bodies, and static / kinematic solver bodies are never modified
// (Solver.elm:507,531,796,815).  level(g) = 1 + max(level of the previous group touching either dynamic body); the groups
// of one level are mutually independent, and executing the levels in ascending order keeps, for every body, the
// reference's order of the groups that touch it (descending order for the warm start, which walks the groups backwards,
// Solver.elm:122-208).  Every group computes bit-for-bit what the sequential sweep computes.
// sum|dlambda| decides the early exit (Solver.elm:360-367).  All terms are >= 0, so ANY summation order is within
// (rows+2)*2^-53 relative of the reference's left-to-right sum; the team reduces in parallel and only when that sum lands
// within 1e-9 relative of the tolerance does thread 0 redo it in the reference's association from the per-row |dlambda|.
template <int T>
__global__ void __launch_bounds__(T) k_solve_team(Dev d, int bin_lo, int bin_hi) {
  __shared__ int s_nlev, s_decide;
  __shared__ double s_part[T / 32];
  const int tid = threadIdx.x;
  int total = 0;
  for (int b = bin_lo; b <= bin_hi; b++) total += d.cur.counters[CNT_BIN0 + b];
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int isl = island_from_worklist(d, item, bin_lo, bin_hi);
    const int g0 = d.isl_off[isl], gn = d.isl_cnt[isl], w = d.isl_world[isl];
    const int r0 = d.world_body_off[w];
    int* last = d.uf_fill + r0;            // per-body scratch; islands of a world own disjoint dynamic bodies
    int* lvl = d.sched_lvl + g0; int* sched = d.sched + g0; int* loff = d.lvl_off + g0; int* lcur = d.lvl_cur + g0;
    for (int gi = tid; gi < gn; gi += T) {
      int s = d.isl_groups[g0 + gi];
      int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
      if (d.kind[r1] == 2) last[r1 - r0] = 0;
      if (d.kind[r2] == 2) last[r2 - r0] = 0;
      loff[gi] = 0;
    }
    __syncthreads();
    if (tid == 0) {
      int nlev = 0;
      for (int gi = 0; gi < gn; gi++) {
        int s = d.isl_groups[g0 + gi];
        int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
        bool d1 = d.kind[r1] == 2, d2 = d.kind[r2] == 2;
        int l = 0;
        if (d1) l = last[r1 - r0];
        if (d2) l = max(l, last[r2 - r0]);
        lvl[gi] = l;
        l++;
        if (d1) last[r1 - r0] = l;
        if (d2) last[r2 - r0] = l;
        nlev = max(nlev, l);
      }
      s_nlev = nlev;
    }
    __syncthreads();
    const int nlev = s_nlev;
    for (int gi = tid; gi < gn; gi += T) atomicAdd(&loff[lvl[gi]], 1);
    __syncthreads();
    if (tid == 0) { int run = 0; for (int l = 0; l < nlev; l++) { int c = loff[l]; loff[l] = run; lcur[l] = run; run += c; } }
    __syncthreads();
    for (int gi = tid; gi < gn; gi += T) sched[atomicAdd(&lcur[lvl[gi]], 1)] = d.isl_groups[g0 + gi];
    __syncthreads();
    // warm start, levels descending
    for (int l = nlev - 1; l >= 0; l--) {
      const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
      for (int i = b + tid; i < e; i += T) warm_group(d, sched[i]);
      __syncthreads();
    }
    const int iterations = d.iterations[w];
    int remaining = iterations, rem;
    const double eps = spook(d.dt[w]).eps;
    for (;;) {
      double part = 0;
      for (int l = 0; l < nlev; l++) {
        const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
        for (int i = b + tid; i < e; i += T) part += sweep_nonfriction(d, sched[i], eps, 0.0);
        __syncthreads();
      }
      for (int l = 0; l < nlev; l++) {
        const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
        for (int i = b + tid; i < e; i += T) part += sweep_friction(d, sched[i], eps, 0.0);
        __syncthreads();
      }
      if (remaining == 1) { rem = 0; break; }
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) s_part[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        double S = 0;
        for (int k = 0; k < T / 32; k++) S += s_part[k];
        int below;
        if (S < kSolverTolerance * (1.0 - 1e-9)) below = 1;
        else if (S > kSolverTolerance * (1.0 + 1e-9)) below = 0;
        else {      // the reference's association: sweep 1 then sweep 2 left to right (`step` adds the two sums, `solve2Body` runs one through)
          double p1 = 0;
          for (int gi = 0; gi < gn; gi++) {
            int s = d.isl_groups[g0 + gi];
            for (int k = 0, j0 = d.cur.slot_jstart[s]; k < d.cur.slot_jcount[s]; k++) p1 = p1 + d.jrows[j0 + k].dl;
            for (int k = 0, c0 = d.cur.slot_cstart[s]; k < d.cur.slot_ccount[s]; k++) p1 = p1 + d.cur.rows[c0 + k].dN;
          }
          double p2 = (gn == 1) ? p1 : 0;
          for (int gi = 0; gi < gn; gi++) {
            int s = d.isl_groups[g0 + gi];
            for (int k = 0, c0 = d.cur.slot_cstart[s]; k < d.cur.slot_ccount[s]; k++) p2 = p2 + d.cur.rows[c0 + k].dF1 + d.cur.rows[c0 + k].dF2;
          }
          double deltaTot = (gn == 1) ? p2 : p1 + p2;
          below = (deltaTot - kSolverTolerance < 0) ? 1 : 0;
        }
        s_decide = below;
      }
      __syncthreads();
      const int below = s_decide;
      __syncthreads();
      if (below) { rem = remaining - 1; break; }
      remaining--;
    }
    if (tid == 0) { atomicMin(&d.world_min_rem[w], rem); d.body_iters[d.isl_root[isl]] = iterations - rem; }
    __syncthreads();
  }
}

}  // namespace ep
