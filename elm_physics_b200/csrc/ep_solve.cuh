// Spook / Projected Gauss-Seidel island solver (src/Internal/Solver.elm:345-880), one LANE per island, one WARP per tile of
// 32 islands of the same size class.  The whole working set of a tile — Jacobian rows, Spook scalars, lambdas, the
// islands' body constants and their solver-body delta velocities — is gathered ONCE into shared memory in a
// lane-transposed layout (word k of lane l at [k][l]: conflict-free, every LDS.64 of the warp is two 128 B wavefronts),
// iterated there for all solver iterations, and scattered back once.  HBM sees each row once per step instead of once per
// iteration; the inner loop is bound by the FP64 pipe (non-FMA DADD/DMUL), not by memory.
//
// The reference's order is kept exactly: per island, groups in pair-enumeration order; per iteration sweep 1 (joint rows,
// then contact normals) over all groups, then sweep 2 (friction1, friction2 per contact) over all groups; warm start in
// reverse group order; sum|dlambda| association of `step` (two sums added) vs `solve2Body` (one running sum).
// Islands that do not fit a tile (size classes above the launch's, or more bodies/groups than the tile has) run the
// global-memory path solve_island_global (ep_kernels.cuh) on their lane.
#pragma once
#include "ep_kernels.cuh"

namespace ep {

constexpr int kRowF64 = 38;      // RowRec without the pad: nJ f1J f2J | nB nInvC f1B f1InvC f2B f2InvC eps mu | lamN lamF1 lamF2
constexpr int kBodyF64 = 16;     // invMass, I (9, m11 m21 m31 m12 ...), solver-body delta v/w (6)
constexpr int kBodyI32 = 2;      // global body row, kind
constexpr int kTileGroupI32 = 6;     // local body1, local body2, contact offset (tile), contact count, joint start (global), joint count
struct TileDims { int maxB, maxG, maxC; };
__host__ __device__ inline TileDims tile_dims(int bin) { TileDims t; t.maxC = 1 << bin; t.maxG = (1 << bin); t.maxB = (1 << bin) + 1; return t; }
__host__ __device__ inline size_t tile_bytes(TileDims t) {
  return 32 * (size_t)(8 * (kBodyF64 * t.maxB + kRowF64 * t.maxC) + 4 * (kBodyI32 * t.maxB + (kTileGroupI32 + 1) * t.maxG));
}

struct Tile {
  double* fd; int* fi; int lane; TileDims t;
  __device__ __forceinline__ double& body(int b, int f) const { return fd[(b * kBodyF64 + f) * 32 + lane]; }
  __device__ __forceinline__ double& row(int c, int f) const { return fd[((t.maxB * kBodyF64) + c * kRowF64 + f) * 32 + lane]; }
  __device__ __forceinline__ int& bodyi(int b, int f) const { return fi[(b * kBodyI32 + f) * 32 + lane]; }
  __device__ __forceinline__ int& group(int g, int f) const { return fi[(t.maxB * kBodyI32 + g * (kTileGroupI32 + 1) + f) * 32 + lane]; }
};

__device__ __forceinline__ BodyK tile_bodyk(const Tile& T, int b) {
  BodyK k; k.kind = T.bodyi(b, 1); k.invMass = T.body(b, 0);
  k.I.m11 = T.body(b, 1); k.I.m21 = T.body(b, 2); k.I.m31 = T.body(b, 3); k.I.m12 = T.body(b, 4); k.I.m22 = T.body(b, 5);
  k.I.m32 = T.body(b, 6); k.I.m13 = T.body(b, 7); k.I.m23 = T.body(b, 8); k.I.m33 = T.body(b, 9);
  return k;
}
__device__ __forceinline__ SB tile_sb(const Tile& T, int b) {
  SB s; s.vX = T.body(b, 10); s.vY = T.body(b, 11); s.vZ = T.body(b, 12); s.wX = T.body(b, 13); s.wY = T.body(b, 14); s.wZ = T.body(b, 15);
  return s;
}
__device__ __forceinline__ void tile_store_sb(const Tile& T, int b, const SB& s) {
  T.body(b, 10) = s.vX; T.body(b, 11) = s.vY; T.body(b, 12) = s.vZ; T.body(b, 13) = s.wX; T.body(b, 14) = s.wY; T.body(b, 15) = s.wZ;
}
__device__ __forceinline__ void tile_J(const Tile& T, int c, int f0, double* J) {
#pragma unroll
  for (int k = 0; k < 9; k++) J[k] = T.row(c, f0 + k);
}

// Gather the island of this lane into the tile; false when it does not fit (then nothing of the tile is used by the lane).
__device__ bool tile_gather(const Dev& d, const Tile& T, int isl, int& nB_out, int& nG_out) {
  const int g0 = d.isl_off[isl], gn = d.isl_cnt[isl];
  if (gn > T.t.maxG) return false;
  int nB = 0, coff = 0;
  for (int gi = 0; gi < gn; gi++) {
    int s = d.isl_groups[g0 + gi];
    int rr[2] = {d.cur.slot_row1[s], d.cur.slot_row2[s]};
    int lb[2];
    for (int e = 0; e < 2; e++) {
      int b = 0;
      while (b < nB && T.bodyi(b, 0) != rr[e]) b++;
      if (b == nB) {
        if (nB == T.t.maxB) return false;
        T.bodyi(b, 0) = rr[e];
        nB++;
      }
      lb[e] = b;
    }
    int cn = d.cur.slot_ccount[s];
    if (coff + cn > T.t.maxC) return false;
    T.group(gi, 0) = lb[0]; T.group(gi, 1) = lb[1]; T.group(gi, 2) = coff; T.group(gi, 3) = cn;
    T.group(gi, 4) = d.cur.slot_jstart[s]; T.group(gi, 5) = d.cur.slot_jcount[s]; T.group(gi, 6) = d.cur.slot_cstart[s];
    coff += cn;
  }
  for (int b = 0; b < nB; b++) {
    int r = T.bodyi(b, 0);
    T.bodyi(b, 1) = d.kind[r];
    T.body(b, 0) = d.inv_mass[r];
    const double* m = d.iiw + 9 * (size_t)r;
#pragma unroll
    for (int k = 0; k < 9; k++) T.body(b, 1 + k) = m[k];
    const double* sb = d.sbv + 6 * (size_t)r;
#pragma unroll
    for (int k = 0; k < 6; k++) T.body(b, 10 + k) = sb[k];
  }
  for (int gi = 0; gi < gn; gi++) {
    int c0 = T.group(gi, 6), cn = T.group(gi, 3), off = T.group(gi, 2);
    for (int k = 0; k < cn; k++) {
      const double* src = (const double*)(d.cur.rows + c0 + k);
#pragma unroll 2
      for (int f = 0; f < kRowF64; f++) T.row(off + k, f) = src[f];
    }
  }
  nB_out = nB; nG_out = gn;
  return true;
}

// velocityNonFrictionGroup (Solver.elm:850-864) on the tile
__device__ __forceinline__ double tile_sweep_nonfriction(const Dev& d, const Tile& T, int g, double tot) {
  const int b1 = T.group(g, 0), b2 = T.group(g, 1);
  BodyK k1 = tile_bodyk(T, b1), k2 = tile_bodyk(T, b2);
  SB s1 = tile_sb(T, b1), s2 = tile_sb(T, b2);
  const int j0 = T.group(g, 4), jn = T.group(g, 5);
  for (int k = 0; k < jn; k++) {
    JointRow& r = d.jrows[j0 + k];
    double lam = r.lam;
    double dl = row_solve(r.J, r.B, r.invC, r.eps, -kMaxImpulse, kMaxImpulse, lam, k1, k2, s1, s2);
    r.lam = lam;
    tot = tot + eabs(dl);
  }
  const int c0 = T.group(g, 2), cn = T.group(g, 3);
  for (int k = 0; k < cn; k++) {
    const int c = c0 + k;
    double J[9]; tile_J(T, c, 0, J);
    double lam = T.row(c, 35);
    double dl = row_solve(J, T.row(c, 27), T.row(c, 28), T.row(c, 33), 0.0, kMaxImpulse, lam, k1, k2, s1, s2);
    T.row(c, 35) = lam;
    tot = tot + eabs(dl);
  }
  if (k1.kind == 2) tile_store_sb(T, b1, s1);
  if (k2.kind == 2) tile_store_sb(T, b2, s2);
  return tot;
}
// velocityFrictionGroup (Solver.elm:869-880) on the tile
__device__ __forceinline__ double tile_sweep_friction(const Tile& T, int g, double tot) {
  const int c0 = T.group(g, 2), cn = T.group(g, 3);
  if (cn <= 0) return tot;
  const int b1 = T.group(g, 0), b2 = T.group(g, 1);
  BodyK k1 = tile_bodyk(T, b1), k2 = tile_bodyk(T, b2);
  SB s1 = tile_sb(T, b1), s2 = tile_sb(T, b2);
  for (int k = 0; k < cn; k++) {
    const int c = c0 + k;
    double e1[9], e2[9]; tile_J(T, c, 9, e1); tile_J(T, c, 18, e2);
    double l1 = T.row(c, 36), l2 = T.row(c, 37);
    tot = friction_pair(e1, e2, T.row(c, 29), T.row(c, 30), T.row(c, 31), T.row(c, 32), T.row(c, 33), T.row(c, 34), T.row(c, 35),
                        l1, l2, k1, k2, s1, s2, tot);
    T.row(c, 36) = l1; T.row(c, 37) = l2;
  }
  if (k1.kind == 2) tile_store_sb(T, b1, s1);
  if (k2.kind == 2) tile_store_sb(T, b2, s2);
  return tot;
}

// One warp per CTA; dynamic shared memory = tile_bytes(tile_dims(tile_bin)).  Processes size classes [bin_lo, bin_hi].
__global__ void __launch_bounds__(32) k_solve_tiles(Dev d, int bin_lo, int bin_hi, int tile_bin) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  Tile T; T.t = tile_dims(tile_bin); T.lane = threadIdx.x;
  T.fd = (double*)tile_smem; T.fi = (int*)(T.fd + 32 * (kBodyF64 * T.t.maxB + kRowF64 * T.t.maxC));
  int total = 0;
  for (int b = bin_lo; b <= bin_hi; b++) total += d.cur.counters[CNT_BIN0 + b];
  const int n_tiles = (total + 31) >> 5;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int t = tile * 32 + threadIdx.x;
    const int isl = t < total ? island_from_worklist(d, t, bin_lo, bin_hi) : -1;
    if (isl < 0) continue;
    int nB = 0, gn = 0;
    if (!tile_gather(d, T, isl, nB, gn)) { solve_island_global(d, isl); continue; }
    const int w = d.isl_world[isl];
    // warm start in reverse enumeration order (buildAndWarmStart walks pairGroups, Solver.elm:122-208)
    for (int gi = gn - 1; gi >= 0; gi--) {
      const int cn = T.group(gi, 3);
      if (cn <= 0) continue;
      const int b1 = T.group(gi, 0), b2 = T.group(gi, 1), c0 = T.group(gi, 2);
      BodyK k1 = tile_bodyk(T, b1), k2 = tile_bodyk(T, b2);
      SB s1 = tile_sb(T, b1), s2 = tile_sb(T, b2);
      for (int k = 0; k < cn; k++) {
        double lam = T.row(c0 + k, 35);
        if (lam == 0) continue;                        // Solver.elm:41
        double J[9]; tile_J(T, c0 + k, 0, J);
        apply_row(lam, J, k1, k2, s1, s2);
      }
      if (k1.kind == 2) tile_store_sb(T, b1, s1);
      if (k2.kind == 2) tile_store_sb(T, b2, s2);
    }
    int remaining = d.iterations[w], rem;
    for (;;) {
      double p1 = 0, p2;
      for (int gi = 0; gi < gn; gi++) p1 = tile_sweep_nonfriction(d, T, gi, p1);
      // solve2Body keeps ONE running sum through both phases (Solver.elm:473-489); step adds two sums (345-358)
      p2 = (gn == 1) ? p1 : 0;
      for (int gi = 0; gi < gn; gi++) p2 = tile_sweep_friction(T, gi, p2);
      double deltaTot = (gn == 1) ? p2 : p1 + p2;
      if (remaining == 1) { rem = 0; break; }
      if (deltaTot - kSolverTolerance < 0) { rem = remaining - 1; break; }
      remaining--;
    }
    atomicMin(&d.world_min_rem[w], rem);
    // scatter: lambdas back to the slot-order rows (next frame's warm start, Solver.elm:320-335), delta velocities to the bodies
    for (int gi = 0; gi < gn; gi++) {
      int c0 = T.group(gi, 6), cn = T.group(gi, 3), off = T.group(gi, 2);
      for (int k = 0; k < cn; k++) {
        RowRec& r = d.cur.rows[c0 + k];
        r.lamN = T.row(off + k, 35); r.lamF1 = T.row(off + k, 36); r.lamF2 = T.row(off + k, 37);
      }
    }
    for (int b = 0; b < nB; b++) {
      if (T.bodyi(b, 1) != 2) continue;
      double* sb = d.sbv + 6 * (size_t)T.bodyi(b, 0);
#pragma unroll
      for (int k = 0; k < 6; k++) sb[k] = T.body(b, 10 + k);
    }
  }
}

}  // namespace ep
