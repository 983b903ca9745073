// Spook / Projected Gauss-Seidel island solver (src/Internal/Solver.elm:345-880).
//
// Islands are independent (Islands.elm:3-7); inside one the reference's order is kept exactly: per iteration sweep 1
// (joint rows, then contact normals) over the groups in pair-enumeration order, then sweep 2 (friction1, friction2 per
// contact); warm start in reverse group order; sum|dlambda| association of `step` (two sums added) vs `solve2Body` (one
// running sum).  Two mappings, chosen by island size class (island_bin, ep_kernels.cuh):
//
//   classes 0..kLaneBinMax  one LANE per island, one WARP per tile of 32 islands (k_solve_lanes).  The work list is sorted
//       by (class, iterations the island used last frame, group count), so the 32 lanes of a tile run about the same rows
//       for about the same number of iterations.  The warp gathers its tile ONCE into a lane-transposed image (word k of
//       lane l at [k][l]: every access of the warp is one coalesced 256 B request / conflict-free LDS.64): Jacobian rows,
//       Spook scalars, lambdas, the islands' dynamic bodies (constants + solver-body delta velocities) and per-row
//       descriptors (which two bodies a row couples) -- in shared memory where it fits -- and iterates there; HBM sees
//       each row once per step instead of once per iteration.
//   classes above           one CTA per island, the pair groups of one dependency level in parallel (k_solve_team).
#pragma once
#include "ep_kernels.cuh"

namespace ep {

constexpr int kLaneBinMax = 5;   // islands of up to 32 contacts (96 rows) run lane-per-island
// Tile image of 32 islands, every array lane-transposed ([k][lane]); in this order:
//   bodies  [maxB][8] f64    invMass, solver-body delta v/w (6), pad.  Local body 0 is the shared stand-in of every non-dynamic
//                            body: invMass 0, deltas 0 (their solver bodies are never modified, Solver.elm:507,531,796,815)
//   ints    [maxB] body row | [maxR] sweep-1 descriptors | [maxC] sweep-2 descriptors | [maxC] global contact index
//   nrows   [maxR][18] f64   sweep-1 rows in solver order (per group: joint rows, then contact normals):
//                            J[9], I1.wA (3), I2.wB (3), B, invC, lambda
//   frows   [maxC][37] f64   sweep-2 rows: f1J[9], I1.wA, I2.wB, f2J[9], I1.wA, I2.wB, f1B, f1InvC, f2B, f2InvC, mu, lambdaF1, lambdaF2
// I.w = invInertiaWorld x the row's angular Jacobian: the reference evaluates (I.w) * deltaLambda on every row of every
// iteration (Solver.elm:505-550); I and w are constant during the solve, so the product is formed ONCE here, with the same
// operations in the same order, and the solver neither holds the 3x3 matrices in registers nor redoes the 15 flops.
// For the stand-in body I = 0 and invMass = 0, so its update adds +-0 to +0: it stays +0 without any branch on the body kind.
// descriptor = local body1 | local body2 << 8 | (sweep 1: joint flag << 16, first-row-of-group flag << 17)
//                                              | (sweep 2: index of the contact's normal row << 16)
constexpr int kBodyF64 = 8, kNRowF64 = 18, kFRowF64 = 37;
enum { NR_IA = 9, NR_IB = 12, NR_B = 15, NR_INVC = 16, NR_LAM = 17 };
enum { FR_E1 = 0, FR_E1IA = 9, FR_E1IB = 12, FR_E2 = 15, FR_E2IA = 24, FR_E2IB = 27, FR_F1B = 30, FR_F1INVC = 31, FR_F2B = 32, FR_F2INVC = 33,
       FR_MU = 34, FR_L1 = 35, FR_L2 = 36 };
enum { DESC_JOINT = 1 << 16, DESC_GROUP_START = 1 << 17 };
// Islands per tile (= active lanes of its warp) by size class.  Measured on B200: putting fewer islands of the large
// classes in a tile so that the whole image fits shared memory does NOT pay (the tile's latency is the FP64 dependency
// chain, ~1900 cycles per contact-iteration wherever the rows live, and 8x more tiles then queue for the same SM slots),
// so every class uses full warps and the large classes stream their rows from L2 one row ahead.
__host__ __device__ constexpr int class_lanes(int bin) { return 32; }
struct TileDims { int maxB, maxR, maxC; };
__host__ __device__ inline int64_t tile_head_words(TileDims t, int L) { return L * (int64_t)(kBodyF64 * t.maxB) + (L / 2) * (int64_t)(t.maxB + t.maxR + 2 * t.maxC); }
__host__ __device__ inline int64_t tile_words(TileDims t, int L) { return tile_head_words(t, L) + L * (int64_t)(kNRowF64 * t.maxR + kFRowF64 * t.maxC); }   // 8-byte words

// LT > 0: compile-time lane stride (the staged kernels); LT == 0: run-time stride (the in-place fallback)
template <int LT>
struct TileT {
  double* bd; int* ti; double* rd; int lane; int Lrt; TileDims t;     // bodies, ints, rows
  __device__ __forceinline__ int L() const { return LT ? LT : Lrt; }
  __device__ __forceinline__ double& body(int b, int f) const { return bd[(b * kBodyF64 + f) * L() + lane]; }
  __device__ __forceinline__ int& body_row(int b) const { return ti[b * L() + lane]; }
  __device__ __forceinline__ int& desc1(int r) const { return ti[(t.maxB + r) * L() + lane]; }
  __device__ __forceinline__ int& desc2(int c) const { return ti[(t.maxB + t.maxR + c) * L() + lane]; }
  __device__ __forceinline__ int& contact_index(int c) const { return ti[(t.maxB + t.maxR + t.maxC + c) * L() + lane]; }
  __device__ __forceinline__ double& nrow(int r, int f) const { return rd[(r * kNRowF64 + f) * L() + lane]; }
  __device__ __forceinline__ double& frow(int c, int f) const { return rd[(t.maxR * kNRowF64 + c * kFRowF64 + f) * L() + lane]; }
  // head = bodies + ints (may be staged in shared memory on its own), rows behind it in the image
  __device__ __forceinline__ void bind(double* head, double* rows) {
    bd = head; ti = (int*)(head + L() * (int64_t)(kBodyF64 * t.maxB)); rd = rows;
  }
};
// a row's Jacobian followed by its two I.w triples (computed with the row, k_equations)
template <class TT>
__device__ __forceinline__ void tile_write_jac(const TT& T, bool frow, int idx, int f0, const double* J, const double* Ia, const double* Ib) {
#pragma unroll
  for (int k = 0; k < 9; k++) { if (frow) T.frow(idx, f0 + k) = J[k]; else T.nrow(idx, f0 + k) = J[k]; }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (frow) { T.frow(idx, f0 + 9 + k) = Ia[k]; T.frow(idx, f0 + 12 + k) = Ib[k]; }
    else { T.nrow(idx, f0 + 9 + k) = Ia[k]; T.nrow(idx, f0 + 12 + k) = Ib[k]; }
  }
}

// Tiles of the lane classes are numbered class kLaneBinMax first; tile t of class b holds sorted islands
// [class_start(b) + L_b t, ...).  Returns false past the last tile of class 0.
struct TileRef { int bin, first, count; };     // first sorted index, islands in the tile (1..class_lanes(bin))
__device__ __forceinline__ int class_tiles(const Dev& d, int b) { return (d.cur.counters[CNT_BIN0 + b] + class_lanes(b) - 1) / class_lanes(b); }
__device__ __forceinline__ bool lane_tile(const Dev& d, int gtile, TileRef& r) {
  int base = 0;
  for (int b = kBins - 1; b > kLaneBinMax; b--) base += d.cur.counters[CNT_BIN0 + b];
  for (int b = kLaneBinMax; b >= 0; b--) {
    const int cnt = d.cur.counters[CNT_BIN0 + b], L = class_lanes(b), nt = (cnt + L - 1) / L;
    if (gtile < nt) { r.bin = b; r.first = base + gtile * L; r.count = min(L, cnt - gtile * L); return true; }
    gtile -= nt; base += cnt;
  }
  return false;
}
__device__ __forceinline__ int warp_max(int v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Rows of the tile column of this lane, loaded into registers one row AHEAD of their use so that the load latency of
// row c+1 hides behind the dependent FP64 chain of row c.
struct RowN { double J[9], ia[3], ib[3], B, invC, lam; int desc; };
struct RowF { double e1[9], e1a[3], e1b[3], e2[9], e2a[3], e2b[3], f1B, f1InvC, f2B, f2InvC, mu, l1, l2, nl; int desc; };
// one address computation per row, then constant offsets
template <class TT>
__device__ __forceinline__ void load_row_n(const TT& T, int r, RowN& o) {
  const int L = T.L();
  const double* p = T.rd + ((size_t)r * kNRowF64 * L + T.lane);
#pragma unroll
  for (int k = 0; k < 9; k++) o.J[k] = p[k * L];
#pragma unroll
  for (int k = 0; k < 3; k++) { o.ia[k] = p[(NR_IA + k) * L]; o.ib[k] = p[(NR_IB + k) * L]; }
  o.B = p[NR_B * L]; o.invC = p[NR_INVC * L]; o.lam = p[NR_LAM * L]; o.desc = T.desc1(r);
}
template <class TT>
__device__ __forceinline__ void load_row_f(const TT& T, int c, RowF& o) {
  const int L = T.L();
  const double* p = T.rd + (((size_t)T.t.maxR * kNRowF64 + (size_t)c * kFRowF64) * L + T.lane);
#pragma unroll
  for (int k = 0; k < 9; k++) { o.e1[k] = p[(FR_E1 + k) * L]; o.e2[k] = p[(FR_E2 + k) * L]; }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.e1a[k] = p[(FR_E1IA + k) * L]; o.e1b[k] = p[(FR_E1IB + k) * L]; o.e2a[k] = p[(FR_E2IA + k) * L]; o.e2b[k] = p[(FR_E2IB + k) * L];
  }
  o.f1B = p[FR_F1B * L]; o.f1InvC = p[FR_F1INVC * L]; o.f2B = p[FR_F2B * L]; o.f2InvC = p[FR_F2INVC * L];
  o.mu = p[FR_MU * L]; o.l1 = p[FR_L1 * L]; o.l2 = p[FR_L2 * L]; o.desc = T.desc2(c);
  o.nl = T.nrow((o.desc >> 16) & 0xff, NR_LAM);     // the normal lambda of this contact: sweep 1 is complete when sweep 2 loads
}
// The two solver bodies of the current row (inverse mass + delta v/w), held in registers across consecutive rows that
// share them (pairs are enumerated body1-major, so body1 usually stays).  Local body 0 never touches memory.
struct Pair { int b1, b2; double im1, im2; SB s1, s2; };
template <class TT>
__device__ __forceinline__ void tile_store_sb(const TT& T, int b, const SB& s) {
  const int L = T.L();
  double* p = T.bd + ((size_t)b * kBodyF64 * L + T.lane);
  p[L] = s.vX; p[2 * L] = s.vY; p[3 * L] = s.vZ; p[4 * L] = s.wX; p[5 * L] = s.wY; p[6 * L] = s.wZ;
}
template <class TT>
__device__ __forceinline__ void pair_load_body(const TT& T, int b, double& im, SB& s) {
  const int L = T.L();
  const double* p = T.bd + ((size_t)b * kBodyF64 * L + T.lane);
  im = p[0]; s.vX = p[L]; s.vY = p[2 * L]; s.vZ = p[3 * L]; s.wX = p[4 * L]; s.wY = p[5 * L]; s.wZ = p[6 * L];
}
// Changing the body pair is decided by the WARP: when any active lane's row couples other bodies than its previous row,
// every active lane writes its two bodies back to its tile column and reads the two of the new row -- straight-line code,
// no per-lane branches (ncu: the per-lane version spent 35 % of the solver's samples in these four divergent sections).
// A lane that keeps its pair rewrites and rereads the same values; local body 0 (the stand-in of every non-dynamic body) is
// an ordinary column of zeros: +0 written back, +0 and inverse mass 0 read.
template <class TT>
__device__ __forceinline__ void pair_enter(const TT& T, Pair& P, int desc) {
  const int nb1 = desc & 0xff, nb2 = (desc >> 8) & 0xff;
  if (!__any_sync(__activemask(), (nb1 != P.b1) | (nb2 != P.b2))) return;
  tile_store_sb(T, P.b1, P.s1);
  tile_store_sb(T, P.b2, P.s2);
  pair_load_body(T, nb1, P.im1, P.s1); P.b1 = nb1;
  pair_load_body(T, nb2, P.im2, P.s2); P.b2 = nb2;
}
__device__ __forceinline__ void pair_reset(Pair& P) {      // both sides hold the stand-in
  P.b1 = 0; P.b2 = 0; P.im1 = 0; P.im2 = 0;
  P.s1.vX = P.s1.vY = P.s1.vZ = P.s1.wX = P.s1.wY = P.s1.wZ = 0;
  P.s2.vX = P.s2.vY = P.s2.vZ = P.s2.wX = P.s2.wY = P.s2.wZ = 0;
}
template <class TT>
__device__ __forceinline__ void pair_flush(const TT& T, Pair& P) {
  tile_store_sb(T, P.b1, P.s1);
  tile_store_sb(T, P.b2, P.s2);
  pair_reset(P);
}
// applyVelocityBody1/2 (Solver.elm:505-550) with the precomputed I.w; branch-free (see the stand-in note above)
__device__ __forceinline__ void apply_pre(double dl, const double* J, const double* ia, const double* ib, Pair& P) {
  const double k1 = dl * P.im1, k2 = dl * P.im2;
  P.s1.vX = P.s1.vX - k1 * J[3]; P.s1.vY = P.s1.vY - k1 * J[4]; P.s1.vZ = P.s1.vZ - k1 * J[5];
  P.s1.wX = P.s1.wX + ia[0] * dl; P.s1.wY = P.s1.wY + ia[1] * dl; P.s1.wZ = P.s1.wZ + ia[2] * dl;
  P.s2.vX = P.s2.vX + k2 * J[3]; P.s2.vY = P.s2.vY + k2 * J[4]; P.s2.vZ = P.s2.vZ + k2 * J[5];
  P.s2.wX = P.s2.wX + ib[0] * dl; P.s2.wY = P.s2.wY + ib[1] * dl; P.s2.wZ = P.s2.wZ + ib[2] * dl;
}

// The whole solve of the island of this lane on its tile column, then the scatter of the results.  Each sweep walks ONE
// flat row sequence: ALL 32 lanes of the warp run the same loops (to the largest row count / iteration count of the tile,
// predicated off where a lane has fewer rows, has converged, or holds no island) and re-converge with __syncwarp after
// every row, so changing the body pair is the only divergent section and it never outlives its row.
// the [first, first + bytes) block of the tile image towards L1, lines shared out over the lanes
template <int BYTES>
__device__ __forceinline__ void prefetch_block(const double* first, int lane) {
  const char* p = (const char*)first + lane * 128;
#pragma unroll
  for (int o = 0; o < BYTES; o += 32 * 128)
    if (o + 32 * 128 <= BYTES || o + lane * 128 < BYTES) asm volatile("prefetch.global.L1 [%0];" :: "l"(p + o));
}

template <int PREFETCH, class TT>      // 0: rows read where they are used (staged tiles); 1: next row prefetched to L1 (rows in L2)
__device__ __forceinline__ void solve_lane(const Dev& d, const TT& T, int isl) {
  const bool have = isl >= 0;
  const int gn = have ? d.isl_cnt[isl] : 0, w = have ? d.isl_world[isl] : 0, C = have ? d.isl_nc[isl] : 0, R = have ? C + d.isl_nj[isl] : 0;
  const int Rmax = warp_max(R), Cmax = warp_max(C);
  Pair P; pair_reset(P);
  // warm start (applyContactsWarmStart, Solver.elm:102-119): groups in REVERSE enumeration order (buildAndWarmStart walks
  // pairGroups, Solver.elm:122-208), the seeded normal impulses of a group in solver order
  for (int hi = R - 1; hi >= 0;) {
    int lo = hi;
    while (!(T.desc1(lo) & DESC_GROUP_START)) lo--;
    for (int r = lo; r <= hi; r++) {
      const int desc = T.desc1(r);
      if (desc & DESC_JOINT) continue;
      const double lam = T.nrow(r, NR_LAM);
      if (lam == 0) continue;                        // Solver.elm:41
      pair_enter(T, P, desc);
      double J[9], ia[3], ib[3];
#pragma unroll
      for (int k = 0; k < 9; k++) J[k] = T.nrow(r, k);
#pragma unroll
      for (int k = 0; k < 3; k++) { ia[k] = T.nrow(r, NR_IA + k); ib[k] = T.nrow(r, NR_IB + k); }
      apply_pre(lam, J, ia, ib, P);
    }
    hi = lo - 1;
  }
  pair_flush(T, P);
  __syncwarp();
  const int iterations = have ? d.iterations[w] : 1;
  int remaining = iterations, rem = 0;
  bool done = !have;
  const double eps = have ? spook(d.dt[w]).eps : 0.0;
  while (__any_sync(0xffffffffu, !done)) {
    // sweep 1: velocityNonFrictionGroup (Solver.elm:850-864) group by group: joint rows (572-619), then contact normals (625-672)
    double p1 = 0;
    {
      if (PREFETCH) prefetch_block<kNRowF64 * 32 * 8>(T.rd, T.lane);
      for (int r = 0; r < Rmax; r++) {
        if (PREFETCH && r + 1 < Rmax) prefetch_block<kNRowF64 * 32 * 8>(T.rd + (size_t)(r + 1) * kNRowF64 * 32, T.lane);
        if (!done && r < R) {
          RowN cur;
          load_row_n(T, r, cur);
          pair_enter(T, P, cur.desc);
          const double lo = (cur.desc & DESC_JOINT) ? -kMaxImpulse : 0.0;
          const double prev = cur.invC * (cur.B - gw_lambda(cur.J, P.s1, P.s2) - eps * cur.lam);
          const double dl = (cur.lam + prev - lo < 0) ? lo - cur.lam : ((cur.lam + prev - kMaxImpulse > 0) ? kMaxImpulse - cur.lam : prev);
          apply_pre(dl, cur.J, cur.ia, cur.ib, P);
          T.nrow(r, NR_LAM) = cur.lam + dl;
          p1 = p1 + eabs(dl);
        }
        __syncwarp();
      }
      if (!done) pair_flush(T, P);
      __syncwarp();
    }
    // sweep 2: velocityFrictionGroup (Solver.elm:869-880, 678-844): friction1 then friction2, cone +-mu*lambda_n; solve2Body
    // keeps ONE running sum through both phases (Solver.elm:473-489), step adds two sums (345-358)
    double p2 = (gn == 1) ? p1 : 0;
    {
      const double* fbase = T.rd + (size_t)T.t.maxR * kNRowF64 * 32;
      if (PREFETCH) prefetch_block<kFRowF64 * 32 * 8>(fbase, T.lane);
      for (int c = 0; c < Cmax; c++) {
        if (PREFETCH && c + 1 < Cmax) prefetch_block<kFRowF64 * 32 * 8>(fbase + (size_t)(c + 1) * kFRowF64 * 32, T.lane);
        if (!done && c < C) {
          RowF cur;
          load_row_f(T, c, cur);
          pair_enter(T, P, cur.desc);
          const double cap = cur.mu * cur.nl;
          const double dPrev1 = cur.f1InvC * (cur.f1B - gw_lambda(cur.e1, P.s1, P.s2) - eps * cur.l1);
          const double d1 = (cur.l1 + dPrev1 + cap < 0) ? -cap - cur.l1 : ((cur.l1 + dPrev1 - cap > 0) ? cap - cur.l1 : dPrev1);
          apply_pre(d1, cur.e1, cur.e1a, cur.e1b, P);
          const double dPrev2 = cur.f2InvC * (cur.f2B - gw_lambda(cur.e2, P.s1, P.s2) - eps * cur.l2);
          const double d2 = (cur.l2 + dPrev2 + cap < 0) ? -cap - cur.l2 : ((cur.l2 + dPrev2 - cap > 0) ? cap - cur.l2 : dPrev2);
          apply_pre(d2, cur.e2, cur.e2a, cur.e2b, P);
          T.frow(c, FR_L1) = cur.l1 + d1; T.frow(c, FR_L2) = cur.l2 + d2;
          p2 = p2 + eabs(d1) + eabs(d2);
        }
        __syncwarp();
      }
      if (!done) pair_flush(T, P);
      __syncwarp();
    }
    if (!done) {
      const double deltaTot = (gn == 1) ? p2 : p1 + p2;
      if (remaining == 1) { rem = 0; done = true; }
      else if (deltaTot - kSolverTolerance < 0) { rem = remaining - 1; done = true; }
      else remaining--;
    }
  }
  if (!have) return;
  atomicMin(&d.world_min_rem[w], rem);
  d.body_iters[d.isl_root[isl]] = iterations - rem;
  atomicAdd(&d.work[0], (unsigned long long)C * (iterations - rem));
  if (R > C) atomicAdd(&d.work[1], (unsigned long long)(R - C) * (iterations - rem));
  // scatter: lambdas back to the slot-order rows (next frame's warm start, Solver.elm:320-335), delta velocities to the bodies
  for (int c = 0; c < C; c++) {
    d.cur.contacts[T.contact_index(c)].lam = T.nrow((T.desc2(c) >> 16) & 0xff, NR_LAM);      // next frame's warm start (Solver.elm:320-335)
  }
  const int nB = d.isl_nb[isl] + 1;
  for (int b = 1; b < nB; b++) {
    double* sb = d.sbv + 6 * (size_t)T.body_row(b);
#pragma unroll
    for (int k = 0; k < 6; k++) sb[k] = T.body(b, 1 + k);
  }
}

// ---- k_solve_lanes : tile build + solve in one kernel, one warp per tile -------------------------------------------------
// The warp of a tile
//   (A) walks the pair groups of its 32 islands: body tables (k_islands' body_local numbering; local body 0 is the stand-in),
//       row descriptors, joint rows copied from k_joint_equations' records,
//   (B) gathers the rows of every contact from k_equations' per-contact records (RowRec) into the lane-transposed image --
//       shared memory (MODE 2: whole image; these tiles have no image in HBM at all), or the arena for the row part
//       (MODE 1: bodies + descriptors in shared memory, rows written once with whole-line stores and streamed back from L2
//       by the same SM one row ahead; MODE 0: everything in the arena, four warps per CTA),
//   (C) solves (solve_lane).
// Residency of a tile follows from the dynamic shared memory of the launch that serves its class (Dev::tile_smem); every
// launch visits the tiles of its classes and takes those of its own MODE.  Round 1 did (A) and (B) in a kernel of their own
// (k_tile_build) in front of the solver: 0.29 ms at 8192 worlds during which nothing else ran; inside the solver's warps the
// gather overlaps the solves of the other tiles (step 2.87 -> 2.82 ms at 8192 worlds, 1.98 -> 1.90 at 4096), and the tiles of
// classes 0-2 never exist in HBM.  Forming the rows here as well (contactEquations per lane, no RowRec at all) was built,
// bit-exact, and measured: the 1300 instructions per contact at the solver's occupancy cost what k_equations + k_tile_build
// cost (2.92 ms), so k_equations stays a thread-per-contact kernel beside k_islands (profiles/r02_solver_notes.md).
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? 128 : 32, MODE == 0 ? 3 : 12) k_solve_lanes(Dev d, int bin_lo, int bin_hi) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (d.flags[FLAG_POOL_OVERFLOW]) return;       // the contact pool overflowed: the frame is void; ep_sync reports it
  int t0 = 0;                     // first global tile of class bin_hi
  for (int b = kLaneBinMax; b > bin_hi; b--) t0 += class_tiles(d, b);
  int nt = 0;
  for (int b = bin_hi; b >= bin_lo; b--) nt += class_tiles(d, b);
  for (int t = warp; t < nt; t += nwarps) {
    TileRef tr; lane_tile(d, t0 + t, tr);
    const int isl = lane < tr.count ? d.isl_sorted[tr.first + lane] : -1;
    const int g0 = isl >= 0 ? d.isl_off[isl] : 0, gn = isl >= 0 ? d.isl_cnt[isl] : 0;
    const int nc = isl >= 0 ? d.isl_nc[isl] : 0, nj = isl >= 0 ? d.isl_nj[isl] : 0;
    TileT<32> T; T.lane = lane; T.Lrt = 32;
    T.t.maxB = warp_max(isl >= 0 ? d.isl_nb[isl] : 0) + 1;
    T.t.maxR = warp_max(nc + nj);
    T.t.maxC = warp_max(nc);
    const int64_t words = tile_words(T.t, 32), head = tile_head_words(T.t, 32);
    const int mode = tr.bin <= 2 ? (words * 8 <= d.tile_smem[tr.bin] ? 2 : (head * 8 <= d.tile_smem[3] ? 1 : 0)) : (head * 8 <= d.tile_smem[tr.bin] ? 1 : 0);
    if (mode != MODE) continue;
    if (T.t.maxB >= 256 || T.t.maxR >= 256) { if (lane == 0) d.flags[FLAG_POOL_OVERFLOW] = 2; continue; }     // cannot happen for classes <= kLaneBinMax (rows <= 96); refuse rather than corrupt
    const int64_t need = MODE == 2 ? 0 : (MODE == 1 ? words - head : words);
    unsigned long long off = 0;
    if (MODE != 2) {
      if (lane == 0) {
        off = atomicAdd(d.tile_top, (unsigned long long)need);
        if ((int64_t)(off + need) > d.tile_arena_cap) { d.flags[FLAG_POOL_OVERFLOW] = 2; off = ~0ull; }
      }
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off == ~0ull) continue;
    }
    double* sm = (double*)tile_smem;
    __syncwarp();
    if (MODE == 2) T.bind(sm, sm + head);
    else if (MODE == 1) T.bind(sm, d.tile_arena + off);
    else T.bind(d.tile_arena + off, d.tile_arena + off + head);
    // ---- (A) body table, descriptors, joint rows -----------------------------------------------------------------------------
    if (isl >= 0) {
      T.body_row(0) = -1;
#pragma unroll
      for (int k = 0; k < kBodyF64; k++) T.body(0, k) = 0.0;
    }
    const int gmax = warp_max(gn);
    int r = 0, c = 0;
    for (int gi = 0; gi < gmax; gi++) {
      if (gi >= gn) continue;
      const int s = d.isl_groups[g0 + gi];
      const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
      const int lb1 = d.kind[r1] == 2 ? d.body_local[r1] : 0, lb2 = d.kind[r2] == 2 ? d.body_local[r2] : 0;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int lb = e ? lb2 : lb1, br = e ? r2 : r1;
        if (lb == 0) continue;                       // several groups may write the same body: same values
        T.body_row(lb) = br;
        T.body(lb, 0) = d.inv_mass[br];
#pragma unroll
        for (int k = 1; k < kBodyF64; k++) T.body(lb, k) = 0.0;       // the solver-body deltas start at zero (SolverBody.elm:28-38)
      }
      const int pair = lb1 | (lb2 << 8);
      int first = DESC_GROUP_START;
      const int j0 = d.cur.slot_jstart[s], jn = max(d.cur.slot_jcount[s], 0);
      for (int k = 0; k < jn; k++, r++) {
        const JointRow& jr = d.jrows[j0 + k];
        tile_write_jac(T, false, r, 0, jr.J, jr.Ia, jr.Ib);
        T.nrow(r, NR_B) = jr.B; T.nrow(r, NR_INVC) = jr.invC; T.nrow(r, NR_LAM) = jr.lam;
        T.desc1(r) = pair | DESC_JOINT | first; first = 0;
      }
      const int c0 = d.cur.slot_cstart[s], cn = max(d.cur.slot_ccount[s], 0);
      for (int k = 0; k < cn; k++, r++, c++) {
        T.desc1(r) = pair | first; first = 0;
        T.desc2(c) = pair | (r << 16);
        T.contact_index(c) = c0 + k;
      }
    }
    __syncwarp();
    // ---- (B) contact rows, all lanes in lock step on the contact index so that every store of the warp covers whole 128 B lines
    // where the rows go to the arena (a lane-by-lane copy in group order writes the same lines with partial masks: twice the
    // DRAM traffic, measured)
    for (int cc = 0; cc < T.t.maxC; cc++) {
      if (cc >= nc) continue;
      const int rowN = (T.desc2(cc) >> 16) & 0xff;
      // the record in two halves (15 + 14 independent 16 B loads in flight each), then the transposed stores; RowRec field
      // offsets in doubles: nJ 0, nB 9, nInvC 10, lamN 11, f1J 12, f2J 21, f1B 30, f1InvC 31, f2B 32, f2InvC 33, mu 34,
      // lamF1 35, lamF2 36, d* 37-39, nIa 40, nIb 43, f1Ia 46, f1Ib 49, f2Ia 52, f2Ib 55
      const double2* sp = (const double2*)(d.rows + T.contact_index(cc));
      {
        double2 v[15];
#pragma unroll
        for (int f = 0; f < 15; f++) v[f] = sp[f];
        const double* src = (const double*)v;       // doubles 0..29
#pragma unroll
        for (int k = 0; k < 9; k++) { T.nrow(rowN, k) = src[k]; T.frow(cc, FR_E1 + k) = src[12 + k]; T.frow(cc, FR_E2 + k) = src[21 + k]; }
        T.nrow(rowN, NR_B) = src[9]; T.nrow(rowN, NR_INVC) = src[10]; T.nrow(rowN, NR_LAM) = src[11];
      }
      {
        double2 v[14];
#pragma unroll
        for (int f = 0; f < 14; f++) v[f] = sp[15 + f];
        const double* src = (const double*)v - 30;  // doubles 30..57
        T.frow(cc, FR_F1B) = src[30]; T.frow(cc, FR_F1INVC) = src[31]; T.frow(cc, FR_F2B) = src[32]; T.frow(cc, FR_F2INVC) = src[33];
        T.frow(cc, FR_MU) = src[34]; T.frow(cc, FR_L1) = src[35]; T.frow(cc, FR_L2) = src[36];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          T.nrow(rowN, NR_IA + k) = src[40 + k]; T.nrow(rowN, NR_IB + k) = src[43 + k];
          T.frow(cc, FR_E1IA + k) = src[46 + k]; T.frow(cc, FR_E1IB + k) = src[49 + k];
          T.frow(cc, FR_E2IA + k) = src[52 + k]; T.frow(cc, FR_E2IB + k) = src[55 + k];
        }
      }
    }
    __syncwarp();
    solve_lane<MODE == 1 ? 1 : 0>(d, T, isl);
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
// Data: the solver-body deltas and inverse masses of the island's world sit in shared memory (SMEMB; else in place in
// global memory), the level walk runs on shared-memory copies of the groups' body pairs, every group's record
// (bodies, row ranges) is laid out in schedule order, and a group's rows are read one row ahead of their use.
struct TeamRec { int b1, b2, c0, cn, j0, jn, pad0, pad1; };      // world-local dynamic body index or -1; contact / joint row ranges

template <int T, bool SMEMB>
__global__ void __launch_bounds__(T) k_solve_team(Dev d, int bin_lo, int bin_hi, int force_resum) {
  constexpr int CH = 8 * T;                 // groups per pass of the level walk
  extern __shared__ __align__(16) unsigned char team_smem[];
  __shared__ int s_nlev, s_decide;
  __shared__ double s_part[T / 32];
  const int tid = threadIdx.x;
  int* s_a = (int*)team_smem; int* s_b = s_a + CH; int* s_lv = s_b + CH;         // chunk of the level walk
  int* s_last = s_lv + CH;                                                        // [nb_max]
  double* s_im = (double*)(s_last + ((d.nb_max + 1) & ~1));                       // [nb_max]
  double* s_sb = s_im + d.nb_max;                                                 // [nb_max][6]
  int total = 0;
  for (int b = bin_lo; b <= bin_hi; b++) total += d.cur.counters[CNT_BIN0 + b];
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int isl = island_from_worklist(d, item, bin_lo, bin_hi);
    const int g0 = d.isl_off[isl], gn = d.isl_cnt[isl], w = d.isl_world[isl];
    const int r0 = d.world_body_off[w], nbw = d.world_body_off[w + 1] - r0;
    int* last = SMEMB ? s_last : d.uf_fill + r0;
    int* lvl = d.sched_lvl + g0; int* sched = d.sched + g0; int* loff = d.lvl_off + g0; int* lcur = d.lvl_cur + g0;
    TeamRec* recs = (TeamRec*)d.sched_rec + g0;
    auto sbp = [&](int l) -> double* { return SMEMB ? s_sb + 6 * l : d.sbv + 6 * (size_t)(r0 + l); };
    auto imass = [&](int l) -> double { return SMEMB ? s_im[l] : d.inv_mass[r0 + l]; };
    if (SMEMB) {
      // the whole world's bodies are copied (entries of other islands are never used nor written back)
      for (int l = tid; l < nbw; l += T) {
        last[l] = 0;
        s_im[l] = d.inv_mass[r0 + l];
        const double* sb = d.sbv + 6 * (size_t)(r0 + l);
#pragma unroll
        for (int k = 0; k < 6; k++) s_sb[6 * l + k] = sb[k];
      }
    }
    for (int gi = tid; gi < gn; gi += T) {
      loff[gi] = 0;
      if (!SMEMB) {      // global scratch is shared with the other islands of the world: touch this island's bodies only
        const int s = d.isl_groups[g0 + gi];
        const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
        if (d.kind[r1] == 2) last[r1 - r0] = 0;
        if (d.kind[r2] == 2) last[r2 - r0] = 0;
      }
    }
    if (tid == 0) s_nlev = 0;
    __syncthreads();
    // dependency levels: a serial walk in enumeration order, on shared-memory copies of the body pairs, CH groups per pass
    for (int base = 0; base < gn; base += CH) {
      const int m = min(CH, gn - base);
      for (int i = tid; i < m; i += T) {
        const int s = d.isl_groups[g0 + base + i];
        const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
        s_a[i] = d.kind[r1] == 2 ? r1 - r0 : -1; s_b[i] = d.kind[r2] == 2 ? r2 - r0 : -1;
      }
      __syncthreads();
      if (tid == 0) {
        int nlev = s_nlev;
        for (int i = 0; i < m; i++) {
          const int a = s_a[i], b = s_b[i];
          int l = 0;
          if (a >= 0) l = last[a];
          if (b >= 0) l = max(l, last[b]);
          s_lv[i] = l;
          l++;
          if (a >= 0) last[a] = l;
          if (b >= 0) last[b] = l;
          nlev = max(nlev, l);
        }
        s_nlev = nlev;
      }
      __syncthreads();
      for (int i = tid; i < m; i += T) { lvl[base + i] = s_lv[i]; atomicAdd(&loff[s_lv[i]], 1); }
      __syncthreads();
    }
    const int nlev = s_nlev;
    if (tid == 0) { int run = 0; for (int l = 0; l < nlev; l++) { int c = loff[l]; loff[l] = run; lcur[l] = run; run += c; } }
    __syncthreads();
    for (int gi = tid; gi < gn; gi += T) {
      const int s = d.isl_groups[g0 + gi];
      const int at = atomicAdd(&lcur[lvl[gi]], 1);
      sched[at] = s;
      const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
      TeamRec rc;
      rc.b1 = d.kind[r1] == 2 ? r1 - r0 : -1; rc.b2 = d.kind[r2] == 2 ? r2 - r0 : -1;
      rc.c0 = d.cur.slot_cstart[s]; rc.cn = max(d.cur.slot_ccount[s], 0); rc.j0 = d.cur.slot_jstart[s]; rc.jn = max(d.cur.slot_jcount[s], 0);
      rc.pad0 = 0; rc.pad1 = 0;
      recs[at] = rc;
    }
    __syncthreads();
    // one pair group of a level: load its bodies, run `rows`, store the dynamic bodies back
    auto load_pair = [&](const TeamRec& rc, Pair& P) {
      P.b1 = rc.b1; P.b2 = rc.b2;
      if (rc.b1 >= 0) { const double* p = sbp(rc.b1); P.im1 = imass(rc.b1); P.s1.vX = p[0]; P.s1.vY = p[1]; P.s1.vZ = p[2]; P.s1.wX = p[3]; P.s1.wY = p[4]; P.s1.wZ = p[5]; }
      else { P.im1 = 0; P.s1.vX = P.s1.vY = P.s1.vZ = P.s1.wX = P.s1.wY = P.s1.wZ = 0; }
      if (rc.b2 >= 0) { const double* p = sbp(rc.b2); P.im2 = imass(rc.b2); P.s2.vX = p[0]; P.s2.vY = p[1]; P.s2.vZ = p[2]; P.s2.wX = p[3]; P.s2.wY = p[4]; P.s2.wZ = p[5]; }
      else { P.im2 = 0; P.s2.vX = P.s2.vY = P.s2.vZ = P.s2.wX = P.s2.wY = P.s2.wZ = 0; }
    };
    auto store_pair = [&](const Pair& P) {
      if (P.b1 >= 0) { double* p = sbp(P.b1); p[0] = P.s1.vX; p[1] = P.s1.vY; p[2] = P.s1.vZ; p[3] = P.s1.wX; p[4] = P.s1.wY; p[5] = P.s1.wZ; }
      if (P.b2 >= 0) { double* p = sbp(P.b2); p[0] = P.s2.vX; p[1] = P.s2.vY; p[2] = P.s2.vZ; p[3] = P.s2.wX; p[4] = P.s2.wY; p[5] = P.s2.wZ; }
    };
    // warm start (applyContactsWarmStart, Solver.elm:102-119), levels descending
    for (int l = nlev - 1; l >= 0; l--) {
      const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
      for (int i = b + tid; i < e; i += T) {
        const TeamRec rc = recs[i];
        if (rc.cn <= 0) continue;
        Pair P; load_pair(rc, P);
        for (int k = 0; k < rc.cn; k++) {
          const RowRec& r = d.rows[rc.c0 + k];
          if (r.lamN == 0) continue;                      // Solver.elm:41
          apply_pre(r.lamN, r.nJ, r.nIa, r.nIb, P);
        }
        store_pair(P);
      }
      __syncthreads();
    }
    const int iterations = d.iterations[w];
    int remaining = iterations, rem;
    const double eps = spook(d.dt[w]).eps;
    for (;;) {
      double part = 0;
      // sweep 1: joint rows (Solver.elm:572-619), then contact normals (625-672) of every group, level by level
      for (int l = 0; l < nlev; l++) {
        const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
        for (int i = b + tid; i < e; i += T) {
          const TeamRec rc = recs[i];
          Pair P; load_pair(rc, P);
          for (int k = 0; k < rc.jn; k++) {
            JointRow& r = d.jrows[rc.j0 + k];
            const double lam = r.lam;
            const double prev = r.invC * (r.B - gw_lambda(r.J, P.s1, P.s2) - r.eps * lam);
            const double dl = (lam + prev - (-kMaxImpulse) < 0) ? (-kMaxImpulse) - lam : ((lam + prev - kMaxImpulse > 0) ? kMaxImpulse - lam : prev);
            apply_pre(dl, r.J, r.Ia, r.Ib, P);
            r.lam = lam + dl; r.dl = eabs(dl);
            part += eabs(dl);
          }
          if (rc.cn > 0) {
            RowRec* rows = d.rows + rc.c0;
            double J[9], ia[3], ib[3], B, invC, lam;
            auto ld = [&](const RowRec& r) {
#pragma unroll
              for (int q = 0; q < 9; q++) J[q] = r.nJ[q];
#pragma unroll
              for (int q = 0; q < 3; q++) { ia[q] = r.nIa[q]; ib[q] = r.nIb[q]; }
              B = r.nB; invC = r.nInvC; lam = r.lamN;
            };
            ld(rows[0]);
            for (int k = 0; k < rc.cn; k++) {
              double cJ[9], cia[3], cib[3]; const double cB = B, cInvC = invC, cLam = lam;
#pragma unroll
              for (int q = 0; q < 9; q++) cJ[q] = J[q];
#pragma unroll
              for (int q = 0; q < 3; q++) { cia[q] = ia[q]; cib[q] = ib[q]; }
              if (k + 1 < rc.cn) ld(rows[k + 1]);            // next row in flight while this one computes
              const double prev = cInvC * (cB - gw_lambda(cJ, P.s1, P.s2) - eps * cLam);
              const double dl = (cLam + prev - 0.0 < 0) ? 0.0 - cLam : ((cLam + prev - kMaxImpulse > 0) ? kMaxImpulse - cLam : prev);
              apply_pre(dl, cJ, cia, cib, P);
              rows[k].lamN = cLam + dl; rows[k].dN = eabs(dl);
              part += eabs(dl);
            }
          }
          store_pair(P);
        }
        __syncthreads();
      }
      // sweep 2: friction1 then friction2 of every contact (Solver.elm:678-844), level by level
      for (int l = 0; l < nlev; l++) {
        const int b = loff[l], e = (l + 1 < nlev) ? loff[l + 1] : gn;
        for (int i = b + tid; i < e; i += T) {
          const TeamRec rc = recs[i];
          if (rc.cn <= 0) continue;
          Pair P; load_pair(rc, P);
          RowRec* rows = d.rows + rc.c0;
          for (int k = 0; k < rc.cn; k++) {
            RowRec& r = rows[k];
            if (k + 1 < rc.cn) {        // next row towards L1 while this one computes (464 B = up to five 128 B lines)
              const char* nx = (const char*)(rows + k + 1);
#pragma unroll
              for (int q = 0; q < 5; q++) asm volatile("prefetch.global.L1 [%0];" :: "l"(nx + (q < 4 ? 128 * q : 463)));
            }
            const double cap = r.mu * r.lamN;
            const double l1 = r.lamF1, l2 = r.lamF2;
            const double dPrev1 = r.f1InvC * (r.f1B - gw_lambda(r.f1J, P.s1, P.s2) - eps * l1);
            const double d1 = (l1 + dPrev1 + cap < 0) ? -cap - l1 : ((l1 + dPrev1 - cap > 0) ? cap - l1 : dPrev1);
            apply_pre(d1, r.f1J, r.f1Ia, r.f1Ib, P);
            const double dPrev2 = r.f2InvC * (r.f2B - gw_lambda(r.f2J, P.s1, P.s2) - eps * l2);
            const double d2 = (l2 + dPrev2 + cap < 0) ? -cap - l2 : ((l2 + dPrev2 - cap > 0) ? cap - l2 : dPrev2);
            apply_pre(d2, r.f2J, r.f2Ia, r.f2Ib, P);
            r.lamF1 = l1 + d1; r.lamF2 = l2 + d2; r.dF1 = eabs(d1); r.dF2 = eabs(d2);
            part += eabs(d1) + eabs(d2);
          }
          store_pair(P);
        }
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
        if (!force_resum && S < kSolverTolerance * (1.0 - 1e-9)) below = 1;
        else if (!force_resum && S > kSolverTolerance * (1.0 + 1e-9)) below = 0;
        else {
          atomicAdd(&d.work[2], 1ull);      // the reference's association: sweep 1 then sweep 2 left to right (`step` adds the two sums, `solve2Body` runs one through)
          double p1 = 0;
          for (int gi = 0; gi < gn; gi++) {
            int s = d.isl_groups[g0 + gi];
            for (int k = 0, j0 = d.cur.slot_jstart[s]; k < d.cur.slot_jcount[s]; k++) p1 = p1 + d.jrows[j0 + k].dl;
            for (int k = 0, c0 = d.cur.slot_cstart[s]; k < d.cur.slot_ccount[s]; k++) p1 = p1 + d.rows[c0 + k].dN;
          }
          double p2 = (gn == 1) ? p1 : 0;
          for (int gi = 0; gi < gn; gi++) {
            int s = d.isl_groups[g0 + gi];
            for (int k = 0, c0 = d.cur.slot_cstart[s]; k < d.cur.slot_ccount[s]; k++) p2 = p2 + d.rows[c0 + k].dF1 + d.rows[c0 + k].dF2;
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
    if (tid == 0) {
      atomicMin(&d.world_min_rem[w], rem); d.body_iters[d.isl_root[isl]] = iterations - rem;
      atomicAdd(&d.work[0], (unsigned long long)d.isl_nc[isl] * (iterations - rem));
      if (d.isl_nj[isl]) atomicAdd(&d.work[1], (unsigned long long)d.isl_nj[isl] * (iterations - rem));
    }
    for (int i = tid; i < gn; i += T) {      // solved normal lambdas to the contacts: the next frame's warm start (Solver.elm:320-335)
      const TeamRec rc = recs[i];
      for (int k = 0; k < rc.cn; k++) d.cur.contacts[rc.c0 + k].lam = d.rows[rc.c0 + k].lamN;
    }
    if (SMEMB) {     // the island's dynamic bodies back to the solver-body array (other islands of the world own other bodies)
      for (int i = tid; i < gn; i += T) {
        const TeamRec rc = recs[i];
        for (int e2 = 0; e2 < 2; e2++) {
          const int l = e2 ? rc.b2 : rc.b1;
          if (l < 0) continue;
          double* o = d.sbv + 6 * (size_t)(r0 + l);
#pragma unroll
          for (int k = 0; k < 6; k++) o[k] = s_sb[6 * l + k];
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace ep
