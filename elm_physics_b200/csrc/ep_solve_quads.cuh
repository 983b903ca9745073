// The QUAD solver: the second family of island solvers, chosen per batch at ep_upload_worlds (Dev::quads) for SMALL batches --
// a shard of a batch that is split over several GPUs -- where the step is bound by the latency of its longest island, not by
// throughput (profiles/r02_solver_notes.md: at 1024 worlds per GPU the class-5 launch takes 0.46 ms against 0.72 ms for the lane
// tiles and no k_tile_build is needed; at 8192 worlds per GPU the lane tiles win by 2.5x and stay in charge).
//
// A row is solved by a QUAD of lanes -- one lane per (body of the pair) x (linear, angular) half, see QRow in ep_device.cuh:
// each lane forms a 3-term partial of G.W, the four partials are exchanged by shuffles and combined in the reference's
// association ((p0 + p1) + p2) + p3 (Solver.elm:642-646), every lane derives the same delta-lambda and updates its own three
// solver-body components (Solver.elm:505-550).  Everything else -- row order, warm start in reverse group order, the two
// sum|dlambda| associations, the early exit -- is as in ep_solve.cuh.
//
//   classes 0..kLaneBinMax  one quad per island, KQ islands per warp (k_solve_quads).  The warp first STAGES its islands in
//       shared memory: the rows of every pair group are copied from the per-contact records k_equations wrote (cp.async, 16 B
//       chunks, no register round trip) into a quad-interleaved image in which every access of the warp is one contiguous,
//       conflict-free request; the islands' dynamic bodies (inverse mass + solver-body deltas, which start at zero,
//       SolverBody.elm:28-38) get a table next to it.  Then it iterates there: HBM / L2 see each row once per step.
//   classes above           one warp / CTA per island, the pair groups of a dependency level in parallel, a quad per group
//       (k_solve_qteam).
#pragma once
#include "ep_solve.cuh"

namespace ep {

// ---- shared-memory image of the KQ islands of a warp ----------------------------------------------------------------------
// Row slot i of the tile = KQ * 192 bytes, field-major and quad-interleaved so that the 32 lanes of one access touch one
// contiguous run:   a[f]: f * KQ*32 + k*32 + q*8        (f = 0..2; 8 B per lane, 32 B per quad)
//                   c[f]: KQ*96 + f * KQ*16 + k*16 + (q>>1)*8     (angular lanes only)
//                   (B, invC): KQ*144 + k*16   (lam, mu): KQ*160 + k*16   ints (desc, gidx, aux, -): KQ*176 + k*16
// Sweep-1 rows (joint rows and contact normals, solver order) fill the slots upwards from 0, sweep-2 rows (friction 1,
// friction 2 per contact) downwards from cap-1: an island of R rows never needs more than R slots whatever the split.
// ints of a sweep-1 slot: desc = local body1 | local body2 << 8 | DESC_JOINT | DESC_GROUP_START, gidx = global contact index
// (-1 for a joint row), aux = index of the row that the warm start visits at this position (reverse group order,
// Solver.elm:122-208).  ints of a sweep-2 slot: desc = the pair, aux = slot of the contact's normal row.
// Body table: entry b = 64 B: (v0 v1 v2 invMass | w0 w1 w2 1.0); entry 0 is the all-zero stand-in of every non-dynamic body
// of the island: inverse mass 0 and I w = 0 make its update add +-0 to +0 -- it stays +0 without a branch on the body kind,
// exactly as the reference never modifies static / kinematic solver bodies (Solver.elm:507,531,796,815).
template <int KQ>
struct QuadTile {
  unsigned char* rows; unsigned char* bodies; int* body_row; int4* groups;
  int cap_rows, cap_bodies, cap_groups;
  static constexpr int kRowBytes = KQ * 192;
  __device__ __forceinline__ unsigned char* slot(int i) const { return rows + (size_t)i * kRowBytes; }
  __device__ __forceinline__ unsigned char* body(int k, int b) const { return bodies + ((size_t)k * cap_bodies + b) * 64; }
  __host__ __device__ static size_t bytes(int cap_rows, int cap_bodies, int cap_groups) {
    return (size_t)cap_rows * kRowBytes + (size_t)KQ * cap_bodies * 68 + (size_t)KQ * cap_groups * 16 + 64;
  }
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// tiles of a launch that serves classes [bin_lo, bin_hi] are numbered class bin_hi first; tile t of class b holds sorted
// islands [class_start(b) + KQ t, ...)
template <int KQ>
__device__ __forceinline__ bool quad_tile(const Dev& d, int t, int bin_lo, int bin_hi, TileRef& r) {
  int base = class_start(d, bin_hi);
  for (int b = bin_hi; b >= bin_lo; b--) {
    const int cnt = d.cur.counters[CNT_BIN0 + b], nt = (cnt + KQ - 1) / KQ;
    if (t < nt) { r.bin = b; r.first = base + t * KQ; r.count = min(KQ, cnt - t * KQ); return true; }
    t -= nt; base += cnt;
  }
  return false;
}
// One row of a quad: the lane's view.
struct LaneRow { double a0, a1, a2, c0, c1, c2, B, invC, lam, mu; int desc, aux; };

template <int KQ>
__global__ void __launch_bounds__(32) k_solve_quads(Dev d, int bin_lo, int bin_hi, int cap_rows, int cap_bodies, int cap_groups) {
  extern __shared__ __align__(16) unsigned char q_smem[];
  __shared__ int s_goff[9], s_g0[8];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x, k = lane >> 2, q = lane & 3, qb = lane & ~3;
  const int kk = k < KQ ? k : 0;                 // lanes beyond the tile's quads alias quad 0 for loads; they never store
  QuadTile<KQ> T;
  T.cap_rows = cap_rows; T.cap_bodies = cap_bodies; T.cap_groups = cap_groups;
  T.rows = q_smem;
  T.bodies = T.rows + (size_t)cap_rows * QuadTile<KQ>::kRowBytes;
  T.body_row = (int*)(T.bodies + (size_t)KQ * cap_bodies * 64);
  T.groups = (int4*)(((uintptr_t)(T.body_row + KQ * cap_bodies) + 15) & ~(uintptr_t)15);
  // byte offsets of this lane inside a row slot
  const int offA = kk * 32 + q * 8, offC = KQ * 96 + kk * 16 + (q >> 1) * 8, offS = KQ * 144 + kk * 16;
  const bool angular = (q & 1) != 0;
  // staging: this lane copies chunks q, q + 4, q + 8 of a 192 B record (chunk 11, the ints, is written by the walk instead)
  int dst_off[3];
#pragma unroll
  for (int x = 0; x < 3; x++) {
    const int j = q + 4 * x;
    dst_off[x] = j < 6 ? (j >> 1) * KQ * 32 + kk * 32 + (j & 1) * 16 : (j < 9 ? KQ * 96 + (j - 6) * KQ * 16 + kk * 16 : KQ * 144 + (j - 9) * KQ * 16 + kk * 16);
  }
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    TileRef tr;
    if (!quad_tile<KQ>(d, tile, bin_lo, bin_hi, tr)) break;
    const bool have = k < KQ && k < tr.count;
    const int isl = have ? d.isl_sorted[tr.first + k] : -1;
    const int g0 = have ? d.isl_off[isl] : 0, gn = have ? d.isl_cnt[isl] : 0, w = have ? d.isl_world[isl] : 0;
    const int C = have ? d.isl_nc[isl] : 0, R1 = have ? C + d.isl_nj[isl] : 0, R2 = 2 * C;
    const int nB = have ? d.isl_nb[isl] + 1 : 1;
    bool bad = have && (R1 + R2 > cap_rows || gn > cap_groups || nB > cap_bodies);      // cannot happen (island_bin); refuse rather than corrupt
    if (bad) d.flags[FLAG_POOL_OVERFLOW] = 3;
    const bool live = have && !bad;
    // ---- stage 1: the pair groups of all islands of the tile, flat over the 32 lanes ------------------------------------------
    {
      int run = 0;
#pragma unroll
      for (int x = 0; x < KQ; x++) {
        const int g = __shfl_sync(full, live ? gn : 0, x * 4), b = __shfl_sync(full, g0, x * 4);
        if (lane == 0) { s_goff[x] = run; s_g0[x] = b; }
        run += g;
      }
      if (lane == 0) s_goff[KQ] = run;
      __syncwarp();
      const int total = run;
      if (live && q == 0) {                          // entry 0 of the island's body table: the stand-in
        double* z = (double*)T.body(k, 0);
#pragma unroll
        for (int x = 0; x < 8; x++) z[x] = 0.0;
        T.body_row[k * cap_bodies] = -1;
      }
      for (int t = lane; t < total; t += 32) {
        int x = 0;
        while (x + 1 < KQ && t >= s_goff[x + 1]) x++;
        const int gi = t - s_goff[x];
        const int s = d.isl_groups[s_g0[x] + gi];
        const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
        const int lb1 = d.kind[r1] == 2 ? d.body_local[r1] : 0, lb2 = d.kind[r2] == 2 ? d.body_local[r2] : 0;
        const int jn = max(d.cur.slot_jcount[s], 0), cn = max(d.cur.slot_ccount[s], 0);
        T.groups[x * cap_groups + gi] = make_int4(lb1 | (lb2 << 8) | (jn << 16), d.cur.slot_cstart[s], cn, d.cur.slot_jstart[s]);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int lb = e ? lb2 : lb1, r = e ? r2 : r1;
          if (lb == 0) continue;
          double* z = (double*)T.body(x, lb);        // several groups may write the same body: same values
          z[0] = 0; z[1] = 0; z[2] = 0; z[3] = d.inv_mass[r]; z[4] = 0; z[5] = 0; z[6] = 0; z[7] = 1.0;
          T.body_row[x * cap_bodies + lb] = r;
        }
      }
      __syncwarp();
    }
    // ---- stage 2: every quad walks its island's groups; the four lanes copy each row's record chunk-wise ---------------------
    {
      const int gmax = warp_max(live ? gn : 0);
      int i1 = 0, i2 = 0;
      for (int gi = 0; gi < gmax; gi++) {
        if (live && gi < gn) {
          const int4 g = T.groups[k * cap_groups + gi];
          const int pair = g.x & 0xffff, jn = g.x >> 16, c0 = g.y, cn = g.z, j0 = g.w;
          const int b1 = i1, wpos = R1 - (i1 + jn + cn);            // the group's rows sit at warm-start positions [wpos, ...)
          int first = DESC_GROUP_START;
          for (int x = 0; x < jn + cn; x++, i1++) {
            const bool joint = x < jn;
            const unsigned char* src = joint ? (const unsigned char*)(d.jqrows + j0 + x) : (const unsigned char*)(d.qrows + 3 * (size_t)(c0 + x - jn));
            unsigned char* dst = T.slot(i1);
            cp_async16(dst + dst_off[0], src + q * 16);
            cp_async16(dst + dst_off[1], src + (q + 4) * 16);
            if (q < 3) cp_async16(dst + dst_off[2], src + (q + 8) * 16);
            if (q == 3) {
              int* ints = (int*)(dst + KQ * 176 + k * 16);
              ints[0] = pair | (joint ? DESC_JOINT : 0) | first; ints[1] = joint ? -1 : c0 + x - jn;
              ((int*)(T.slot(wpos + (i1 - b1)) + KQ * 176 + k * 16))[2] = i1;
            }
            first = 0;
            if (!joint) {
#pragma unroll
              for (int f = 0; f < 2; f++) {
                const unsigned char* fs = src + (1 + f) * sizeof(QRow);
                unsigned char* fd = T.slot(cap_rows - 1 - (i2 + f));
                cp_async16(fd + dst_off[0], fs + q * 16);
                cp_async16(fd + dst_off[1], fs + (q + 4) * 16);
                if (q < 3) cp_async16(fd + dst_off[2], fs + (q + 8) * 16);
                if (q == 3) { int* ints = (int*)(fd + KQ * 176 + k * 16); ints[0] = pair; ints[1] = -1; ints[2] = i1; }
              }
              i2 += 2;
            }
          }
        }
        __syncwarp();
      }
      cp_async_wait_all();
      __syncwarp();
    }
    // ---- the solve ----------------------------------------------------------------------------------------------------------
    const int R1max = warp_max(live ? R1 : 0), R2max = warp_max(live ? R2 : 0);
    const unsigned char* mybody0 = T.bodies + (size_t)kk * cap_bodies * 64 + (q & 1) * 32;     // this lane's half of entry 0
    int cur = 0;                                  // local body this lane holds (0: the stand-in)
    double s0 = 0, s1 = 0, s2 = 0, m = 0;
    auto load_row = [&](const unsigned char* row, LaneRow& o) {
      o.a0 = *(const double*)(row + offA); o.a1 = *(const double*)(row + offA + KQ * 32); o.a2 = *(const double*)(row + offA + 2 * KQ * 32);
      o.c0 = o.a0; o.c1 = o.a1; o.c2 = o.a2;
      if (angular) { o.c0 = *(const double*)(row + offC); o.c1 = *(const double*)(row + offC + KQ * 16); o.c2 = *(const double*)(row + offC + 2 * KQ * 16); }
      const double2 bi = *(const double2*)(row + offS), lm = *(const double2*)(row + offS + KQ * 16);
      const int4 it = *(const int4*)(row + offS + KQ * 32);
      o.B = bi.x; o.invC = bi.y; o.lam = lm.x; o.mu = lm.y; o.desc = it.x; o.aux = it.z;
    };
    // the body pair of a row: lanes whose body changes write theirs back and read the new one (through shared memory: a body
    // may move between the two sides of the quad)
    auto enter = [&](bool on, int desc) {
      const int mine = (q < 2) ? (desc & 0xff) : ((desc >> 8) & 0xff);
      const bool sw = on && mine != cur;
      if (sw && cur != 0) { double* p = (double*)(mybody0 + cur * 64); *(double2*)p = make_double2(s0, s1); p[2] = s2; }
      __syncwarp();
      if (sw) {
        const double* p = (const double*)(mybody0 + mine * 64);
        const double2 x = *(const double2*)p, y = *(const double2*)(p + 2);
        s0 = x.x; s1 = x.y; s2 = y.x; m = y.y; cur = mine;
      }
    };
    // G.W of the row from the four partials of the quad (Solver.elm:642-646)
    auto quad_gw = [&](double a0, double a1, double a2) {
      const double p = (a0 * s0 + a1 * s1) + a2 * s2;
      const double p0 = __shfl_sync(full, p, qb), p1 = __shfl_sync(full, p, qb + 1), p2 = __shfl_sync(full, p, qb + 2), p3 = __shfl_sync(full, p, qb + 3);
      return ((p0 + p1) + p2) + p3;
    };
    // warm start (applyContactsWarmStart, Solver.elm:102-119): groups in REVERSE enumeration order (buildAndWarmStart walks
    // pairGroups, Solver.elm:122-208), the seeded normal impulses of a group in solver order; rows with lambda 0 are skipped (41)
    for (int t = 0; t < R1max; t++) {
      const bool in = live && t < R1;
      const int i = in ? ((const int*)(T.slot(t) + offS + KQ * 32))[2] : 0;
      LaneRow r; load_row(T.slot(i), r);
      const bool on = in && !(r.desc & DESC_JOINT) && r.lam != 0;
      enter(on, r.desc);
      if (on) { const double tm = r.lam * m; s0 = s0 + tm * r.c0; s1 = s1 + tm * r.c1; s2 = s2 + tm * r.c2; }
    }
    const int iterations = live ? d.iterations[w] : 1;
    int remaining = iterations, rem = 0;
    bool done = !live;
    const double eps = live ? spook(d.dt[w]).eps : 0.0;
    while (__any_sync(full, !done)) {
      // sweep 1: velocityNonFrictionGroup (Solver.elm:850-864) group by group: joint rows (572-619), then contact normals (625-672).
      // The row after the current one is read while this one computes, and what does not hang on the solver bodies -- the bounds,
      // lambda's share of the right-hand side, the two clamped candidates -- exists before G.W does: the chain of a row is the dot,
      // the shuffles, the sum, two compares and selects, the update (no branch; the quad's first lane stores under a predicate)
      double p1 = 0;
      {
        auto step1 = [&](int i, const LaneRow& r) {
          const bool on = !done && i < R1;
          const double lo = (r.desc & DESC_JOINT) ? -kMaxImpulse : 0.0;
          const double el = eps * r.lam, dlo = lo - r.lam, dhi = kMaxImpulse - r.lam;
          enter(on, r.desc);
          const double gw = quad_gw(r.a0, r.a1, r.a2);
          const double prev = r.invC * (r.B - gw - el);
          const double sum = r.lam + prev;
          const bool under = sum - lo < 0, over = sum - kMaxImpulse > 0;
          const double dl = under ? dlo : (over ? dhi : prev);
          const double tm = dl * m;
          const double n0 = s0 + tm * r.c0, n1 = s1 + tm * r.c1, n2 = s2 + tm * r.c2;
          s0 = on ? n0 : s0; s1 = on ? n1 : s1; s2 = on ? n2 : s2;
          if (on && q == 0) *(double*)(T.slot(i) + offS + KQ * 16) = r.lam + dl;
          p1 = p1 + (on ? fabs(dl) : 0.0);
        };
        LaneRow ra, rb;
        if (R1max > 0) load_row(T.slot(0), ra);
        int i = 0;
        for (; i + 1 < R1max; i += 2) {
          load_row(T.slot(i + 1), rb);
          step1(i, ra);
          load_row(T.slot(min(i + 2, R1max - 1)), ra);
          step1(i + 1, rb);
        }
        if (i < R1max) step1(i, ra);
      }
      __syncwarp();
      // sweep 2: velocityFrictionGroup (Solver.elm:869-880, 678-844): friction1 then friction2, cone +-mu*lambda_n; solve2Body
      // keeps ONE running sum through both phases (Solver.elm:473-489), step adds two sums (345-358)
      double p2 = (gn == 1) ? p1 : 0;
      {
        // a friction row and the normal lambda of its contact (sweep 1 is complete)
        auto load2 = [&](int j, LaneRow& r, double& nl) {
          load_row(T.slot(cap_rows - 1 - j), r);
          nl = *(const double*)(T.slot(j < R2 ? r.aux : 0) + offS + KQ * 16);
        };
        auto step2 = [&](int j, const LaneRow& r, double nl) {
          const bool on = !done && j < R2;
          const double cap = r.mu * nl;
          const double el = eps * r.lam, dlo = -cap - r.lam, dhi = cap - r.lam;
          enter(on, r.desc);
          const double gw = quad_gw(r.a0, r.a1, r.a2);
          const double prev = r.invC * (r.B - gw - el);
          const double sum = r.lam + prev;
          const bool under = sum + cap < 0, over = sum - cap > 0;
          const double dl = under ? dlo : (over ? dhi : prev);
          const double tm = dl * m;
          const double n0 = s0 + tm * r.c0, n1 = s1 + tm * r.c1, n2 = s2 + tm * r.c2;
          s0 = on ? n0 : s0; s1 = on ? n1 : s1; s2 = on ? n2 : s2;
          if (on && q == 0) *(double*)(T.slot(cap_rows - 1 - j) + offS + KQ * 16) = r.lam + dl;
          p2 = p2 + (on ? fabs(dl) : 0.0);
        };
        LaneRow ra, rb; double na = 0, nb2 = 0;
        if (R2max > 0) load2(0, ra, na);
        int j = 0;
        for (; j + 1 < R2max; j += 2) {
          load2(j + 1, rb, nb2);
          step2(j, ra, na);
          load2(min(j + 2, R2max - 1), ra, na);
          step2(j + 1, rb, nb2);
        }
        if (j < R2max) step2(j, ra, na);
      }
      __syncwarp();
      if (!done) {
        const double deltaTot = (gn == 1) ? p2 : p1 + p2;
        if (remaining == 1) { rem = 0; done = true; }
        else if (deltaTot - kSolverTolerance < 0) { rem = remaining - 1; done = true; }
        else remaining--;
      }
    }
    // the bodies still held in registers go back to the table
    if (live && cur != 0) { double* p = (double*)(mybody0 + cur * 64); p[0] = s0; p[1] = s1; p[2] = s2; }
    __syncwarp();
    if (live) {
      if (q == 0) {
        atomicMin(&d.world_min_rem[w], rem);
        d.body_iters[d.isl_root[isl]] = iterations - rem;
        atomicAdd(&d.work[0], (unsigned long long)C * (iterations - rem));
        if (R1 > C) atomicAdd(&d.work[1], (unsigned long long)(R1 - C) * (iterations - rem));
      }
      // solved normal lambdas back to the contacts (next frame's warm start, Solver.elm:320-335), delta velocities to the bodies
      for (int i = q; i < R1; i += 4) {
        const unsigned char* row = T.slot(i);
        const int gidx = ((const int*)(row + KQ * 176 + k * 16))[1];
        if (gidx >= 0) d.cur.contacts[gidx].lam = *(const double*)(row + KQ * 160 + k * 16);
      }
      for (int b = 1 + (q >> 1); b < nB; b += 2) {
        const double* p = (const double*)(T.body(k, b) + (q & 1) * 32);
        double* sb = d.sbv + 6 * (size_t)T.body_row[k * cap_bodies + b] + (q & 1) * 3;
        sb[0] = p[0]; sb[1] = p[1]; sb[2] = p[2];
      }
    }
    __syncwarp();
  }
}

// ---- k_solve_qteam : one CTA per island, the pair groups of a dependency level in parallel, a QUAD of lanes per group ----
// PGS is sequential inside an island, but two pair groups that share no DYNAMIC body commute exactly (see k_solve_team in
// ep_solve.cuh for the argument and the level schedule: level(g) = 1 + max(level of the previous group touching either dynamic
// body); levels ascending for the sweeps, descending for the warm start).  Inside a level the quads of the CTA take the pair
// groups round-robin; the body pair is fixed for a whole group, so the row loop is the bare row, next row loaded while this
// one computes.  NW = 1: a warp per island (__syncwarp between levels); NW > 1: islands of hundreds of groups.
// sum|dlambda| decides the early exit (Solver.elm:360-367): reduced in parallel, redone in the reference's association from
// the per-row |dlambda| only when the parallel sum lands within 1e-9 relative of the tolerance (all terms are >= 0).
// SB: body table (64 B entries: v | invMass | w | 1; entry 0 = the all-zero stand-in of every non-dynamic body) in shared
// memory, bodies numbered by their index inside the island (body_local); else in global scratch, numbered inside the world.
// SR: the island's group records, its rows (staged once per step by cp.async, schedule order) and their |dlambda| in shared
// memory; else group records and rows are read in place (global memory / L2) and lambda is updated in place.
struct QTeamRec { int lb1, lb2, c0, cn, j0, jn, row0, enum_pos; };      // 32 B, schedule order; row0 = first row of the group in the island's row numbering

template <int NW, bool SB, bool SR>
__global__ void __launch_bounds__(NW * 32) k_solve_qteam(Dev d, int bin_lo, int bin_hi, int force_resum, int cap_bodies, int cap_rows, int cap_groups) {
  constexpr int T = NW * 32, NQ = NW * 8, CH = NW == 1 ? 128 : 1024, LCAP = NW == 1 ? 128 : 1024;       // CH: groups per pass of the level walk
  extern __shared__ __align__(16) unsigned char qt_smem[];
  __shared__ int s_nlev, s_decide;
  __shared__ double s_part[NW];
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, q = lane & 3, qb = lane & ~3, quad = tid >> 2;
  const bool angular = (q & 1) != 0;
  // shared memory: level-walk chunk (3 x CH ints) | level offsets | last level per body | body table | [SR: group records | rows | |dlambda|]
  int* s_a = (int*)qt_smem; int* s_b = s_a + CH; int* s_lv = s_b + CH;
  int* s_loff = s_lv + CH;                                          // [LCAP + 1]
  int* s_last = s_loff + LCAP + 1;                                  // [cap_bodies] (SB) -- else the walk uses global scratch
  double* s_body = (double*)(((uintptr_t)(s_last + (SB ? cap_bodies : 0)) + 15) & ~(uintptr_t)15);     // [cap_bodies][8]
  QTeamRec* s_recs = (QTeamRec*)(s_body + (SB ? (size_t)cap_bodies * 8 : 0));                          // [cap_groups]
  QRow* s_rows = (QRow*)(s_recs + (SR ? cap_groups : 0));                                                // [cap_rows]
  double* s_dl = (double*)(s_rows + (SR ? cap_rows : 0));                                                // [cap_rows]
  double* jdl = d.row_dl + 3 * (size_t)d.contact_cap;
  auto sync = [&]() { if (NW == 1) __syncwarp(); else __syncthreads(); };
  int total = 0;
  for (int b = bin_lo; b <= bin_hi; b++) total += d.cur.counters[CNT_BIN0 + b];
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int isl = island_from_worklist(d, item, bin_lo, bin_hi);
    const int g0 = d.isl_off[isl], gn = d.isl_cnt[isl], w = d.isl_world[isl];
    const int r0 = d.world_body_off[w];
    const int nB = d.isl_nb[isl] + 1;
    int* last = SB ? s_last : d.uf_fill + r0 - 1;       // !SB: indexed by world-local body + 1 (entry of the stand-in unused)
    int* lvl = d.sched_lvl + g0; int* gloff = d.lvl_off + g0; int* lcur = d.lvl_cur + g0; int* epos = d.sched + g0;
    QTeamRec* recs = SR ? s_recs : (QTeamRec*)d.sched_rec + g0;
    double* body = SB ? s_body : d.team_body + 8 * (size_t)(r0 + w);     // !SB: a world owns (its bodies + 1) entries, entry 0 the stand-in, body l at l + 1
    if ((SB && nB > cap_bodies) || (SR && (3 * d.isl_nc[isl] + d.isl_nj[isl] > cap_rows || gn > cap_groups))) {      // island_bin / the launch's capacity rule exclude it; refuse rather than corrupt
      if (tid == 0) d.flags[FLAG_POOL_OVERFLOW] = 4;
      continue;
    }
    // table index of a body row: inside the island (SB) or inside the world (+1; 0 = the stand-in of every non-dynamic body)
    auto tab = [&](int r) -> int { return d.kind[r] == 2 ? (SB ? d.body_local[r] : r - r0 + 1) : 0; };
    if (tid < 8) body[tid] = 0.0;
    for (int gi = tid; gi < gn; gi += T) {
      gloff[gi] = 0;
      const int s = d.isl_groups[g0 + gi];
      const int rr[2] = {d.cur.slot_row1[s], d.cur.slot_row2[s]};
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int b = tab(rr[e]);
        if (b == 0) continue;
        double* z = body + 8 * (size_t)b;              // several groups may write the same body: same values
        z[0] = 0; z[1] = 0; z[2] = 0; z[3] = d.inv_mass[rr[e]]; z[4] = 0; z[5] = 0; z[6] = 0; z[7] = 1.0;
        last[b] = 0;
      }
    }
    if (tid == 0) s_nlev = 0;
    sync();
    // dependency levels: a serial walk in enumeration order, on shared-memory copies of the body pairs, CH groups per pass
    for (int base = 0; base < gn; base += CH) {
      const int mm = min(CH, gn - base);
      for (int i = tid; i < mm; i += T) {
        const int s = d.isl_groups[g0 + base + i];
        s_a[i] = tab(d.cur.slot_row1[s]); s_b[i] = tab(d.cur.slot_row2[s]);
      }
      sync();
      if (tid == 0) {
        int nlev = s_nlev;
        for (int i = 0; i < mm; i++) {
          const int a = s_a[i], b = s_b[i];
          int l = 0;
          if (a > 0) l = last[a];
          if (b > 0) l = max(l, last[b]);
          s_lv[i] = l;
          l++;
          if (a > 0) last[a] = l;
          if (b > 0) last[b] = l;
          nlev = max(nlev, l);
        }
        s_nlev = nlev;
      }
      sync();
      for (int i = tid; i < mm; i += T) { lvl[base + i] = s_lv[i]; atomicAdd(&gloff[s_lv[i]], 1); }
      sync();
    }
    const int nlev = s_nlev;
    const bool lsm = nlev <= LCAP;                     // level offsets in shared memory (else read in place)
    if (tid == 0) { int run = 0; for (int l = 0; l < nlev; l++) { int c = gloff[l]; gloff[l] = run; lcur[l] = run; if (lsm) s_loff[l] = run; run += c; } if (lsm) s_loff[nlev] = gn; }
    sync();
    for (int gi = tid; gi < gn; gi += T) {
      const int s = d.isl_groups[g0 + gi];
      const int at = atomicAdd(&lcur[lvl[gi]], 1);
      epos[gi] = at;
      QTeamRec rc;
      rc.lb1 = tab(d.cur.slot_row1[s]); rc.lb2 = tab(d.cur.slot_row2[s]);
      rc.c0 = d.cur.slot_cstart[s]; rc.cn = max(d.cur.slot_ccount[s], 0); rc.j0 = d.cur.slot_jstart[s]; rc.jn = max(d.cur.slot_jcount[s], 0);
      rc.row0 = 0; rc.enum_pos = gi;
      recs[at] = rc;
    }
    sync();
    auto level_begin = [&](int l) -> int { return lsm ? s_loff[l] : (l < nlev ? gloff[l] : gn); };
    // SR: row numbering in schedule order (per group: joint rows, contact normals, then friction 1 / friction 2 per contact) and the
    // rows themselves, copied once from the records k_equations wrote -- 12 chunks of 16 B per row, no register round trip
    if (SR) {
      if (tid == 0) { int run = 0; for (int i = 0; i < gn; i++) { recs[i].row0 = run; run += recs[i].jn + 3 * recs[i].cn; } }
      sync();
      for (int i = quad; i < gn; i += NQ) {
        const QTeamRec rc = recs[i];
        const int nr = rc.jn + 3 * rc.cn;
        for (int x = 0; x < nr; x++) {
          const QRow* src = x < rc.jn ? d.jqrows + rc.j0 + x : (x < rc.jn + rc.cn ? d.qrows + 3 * (size_t)(rc.c0 + x - rc.jn) : d.qrows + 3 * (size_t)(rc.c0 + ((x - rc.jn - rc.cn) >> 1)) + 1 + ((x - rc.jn - rc.cn) & 1));
          unsigned char* dst = (unsigned char*)(s_rows + rc.row0 + x);
#pragma unroll
          for (int ch = 0; ch < 3; ch++) cp_async16(dst + (q + 4 * ch) * 16, (const unsigned char*)src + (q + 4 * ch) * 16);
        }
      }
      cp_async_wait_all();
      sync();
    }
    // this lane's half of a table entry: (v | invMass) for the linear lanes, (w | 1) for the angular ones
    auto load_half = [&](int b, double& s0, double& s1, double& s2, double& m) {
      const double2* p = (const double2*)(body + 8 * (size_t)b + (q & 1) * 4);
      const double2 x = p[0], y = p[1];
      s0 = x.x; s1 = x.y; s2 = y.x; m = y.y;
    };
    auto store_half = [&](int b, double s0, double s1, double s2) {
      if (b == 0) return;
      double* p = body + 8 * (size_t)b + (q & 1) * 4;
      *(double2*)p = make_double2(s0, s1); p[2] = s2;
    };
    // the row records of a group: sweep 1 = joint rows then contact normals, sweep 2 = friction 1, friction 2 per contact
    auto row_ptr = [&](const QTeamRec& rc, int sweep, int x) -> QRow* {
      if (SR) return s_rows + rc.row0 + (sweep == 1 ? x : rc.jn + rc.cn + x);
      if (sweep == 2) return d.qrows + 3 * (size_t)(rc.c0 + (x >> 1)) + 1 + (x & 1);
      return x < rc.jn ? d.jqrows + rc.j0 + x : d.qrows + 3 * (size_t)(rc.c0 + x - rc.jn);
    };
    auto dl_ref = [&](const QTeamRec& rc, int sweep, int x) -> double* {
      if (SR) return s_dl + rc.row0 + (sweep == 1 ? x : rc.jn + rc.cn + x);
      if (sweep == 2) return d.row_dl + 3 * (size_t)(rc.c0 + (x >> 1)) + 1 + (x & 1);
      return x < rc.jn ? jdl + rc.j0 + x : d.row_dl + 3 * (size_t)(rc.c0 + x - rc.jn);
    };
    // a row as the lane sees it; everything that does not hang on the solver bodies -- the bounds, lambda's share of the right-hand
    // side, the two clamped candidates of delta-lambda, the address of |dlambda| -- is formed when the row is fetched, off the chain
    struct TRow { double a0, a1, a2, c0, c1, c2, B, invC, lam, lo, hi, el, dlo, dhi; QRow* at; double* dlp; };
    auto load_trow = [&](QRow* R, TRow& o) {
      o.a0 = R->a[0][q]; o.a1 = R->a[1][q]; o.a2 = R->a[2][q];
      o.c0 = o.a0; o.c1 = o.a1; o.c2 = o.a2;
      if (angular) { o.c0 = R->c[0][q >> 1]; o.c1 = R->c[1][q >> 1]; o.c2 = R->c[2][q >> 1]; }
      const double2 bi = *(const double2*)&R->B, lm = *(const double2*)&R->lam;
      o.B = bi.x; o.invC = bi.y; o.lam = lm.x; o.lo = lm.y /* mu, until the caller turns it into the bounds */; o.hi = 0; o.at = R; o.dlp = nullptr;
      o.el = 0; o.dlo = 0; o.dhi = 0;
    };
    QTeamRec none; none.lb1 = 0; none.lb2 = 0; none.c0 = 0; none.cn = 0; none.j0 = 0; none.jn = 0; none.row0 = 0; none.enum_pos = 0;
    // warm start (applyContactsWarmStart, Solver.elm:102-119), levels descending; joint rows and rows with lambda 0 are skipped (41)
    for (int l = nlev - 1; l >= 0; l--) {
      const int lb = level_begin(l), le = level_begin(l + 1);
      for (int base = lb; base < le; base += NQ) {
        const int i = base + quad;
        const QTeamRec rc = i < le ? recs[i] : none;
        if (rc.cn <= 0) continue;
        const int mine = (q < 2) ? rc.lb1 : rc.lb2;
        double s0, s1, s2, m; load_half(mine, s0, s1, s2, m);
        for (int x = 0; x < rc.cn; x++) {
          TRow r; load_trow(row_ptr(rc, 1, rc.jn + x), r);
          if (r.lam == 0) continue;
          const double tm = r.lam * m; s0 = s0 + tm * r.c0; s1 = s1 + tm * r.c1; s2 = s2 + tm * r.c2;
        }
        store_half(mine, s0, s1, s2);
      }
      sync();
    }
    const int iterations = d.iterations[w];
    int remaining = iterations, rem;
    const double eps = spook(d.dt[w]).eps;
    for (;;) {
      double part = 0;
#pragma unroll 1
      for (int sweep = 1; sweep <= 2; sweep++) {
        for (int l = 0; l < nlev; l++) {
          const int lb = level_begin(l), le = level_begin(l + 1);
          for (int base = lb; base < le; base += NQ) {
            const int i = base + quad;
            const QTeamRec rc = i < le ? recs[i] : none;
            const int n = sweep == 1 ? rc.jn + rc.cn : 2 * rc.cn;
            const int nmax = __reduce_max_sync(full, n);
            if (nmax == 0) continue;
            const int mine = (q < 2) ? rc.lb1 : rc.lb2;
            double s0, s1, s2, m; load_half(mine, s0, s1, s2, m);
            // one row: G.W from the quad's four partials (Solver.elm:642-646), clamp, update of this lane's half -- without a branch:
            // selects between candidates that exist before G.W does, a predicated store by the quad's first lane
            auto solve_row = [&](int x, const TRow& r) {
              const bool on = x < n;
              const double p = (r.a0 * s0 + r.a1 * s1) + r.a2 * s2;
              const double p0 = __shfl_sync(full, p, qb), p1 = __shfl_sync(full, p, qb + 1), p2 = __shfl_sync(full, p, qb + 2), p3 = __shfl_sync(full, p, qb + 3);
              const double gw = ((p0 + p1) + p2) + p3;
              const double prev = r.invC * (r.B - gw - r.el);
              const double sum = r.lam + prev;
              const bool under = sum - r.lo < 0, over = sum - r.hi > 0;
              const double dl = under ? r.dlo : (over ? r.dhi : prev);
              const double tm = dl * m;
              const double n0 = s0 + tm * r.c0, n1 = s1 + tm * r.c1, n2 = s2 + tm * r.c2;
              s0 = on ? n0 : s0; s1 = on ? n1 : s1; s2 = on ? n2 : s2;
              const double ad = fabs(dl);
              const bool wr = on && q == 0;
              part += wr ? ad : 0.0;
              if (wr) { r.at->lam = r.lam + dl; *r.dlp = ad; }
            };
            // a row and, for a friction row, the normal lambda of its contact (sweep 1 is complete when sweep 2 runs)
            auto fetch = [&](int x, TRow& o) {
              load_trow(row_ptr(rc, sweep, x), o);
              const double mu = o.lo;
              if (sweep == 2) { const double cap = mu * row_ptr(rc, 1, rc.jn + (x >> 1))->lam; o.lo = -cap; o.hi = cap; }
              else { o.lo = x < rc.jn ? -kMaxImpulse : 0.0; o.hi = kMaxImpulse; }
              o.el = eps * o.lam; o.dlo = o.lo - o.lam; o.dhi = o.hi - o.lam;
              o.dlp = dl_ref(rc, sweep, x);
            };
            TRow ra, rb;
            fetch(0, ra);             // a quad without rows reads row 0 of an empty record: a valid address, never used
            rb = ra;
            int x = 0;
            for (; x + 1 < nmax; x += 2) {
              if (x + 1 < n) fetch(x + 1, rb);
              solve_row(x, ra);
              if (x + 2 < n) fetch(x + 2, ra);
              solve_row(x + 1, rb);
            }
            if (x < nmax) solve_row(x, ra);
            store_half(mine, s0, s1, s2);
          }
          sync();
        }
      }
      if (remaining == 1) { rem = 0; break; }
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(full, part, o);
      if (lane == 0) s_part[tid >> 5] = part;
      sync();
      if (tid == 0) {
        double S = 0;
        for (int kx = 0; kx < NW; kx++) S += s_part[kx];
        int below;
        if (!force_resum && S < kSolverTolerance * (1.0 - 1e-9)) below = 1;
        else if (!force_resum && S > kSolverTolerance * (1.0 + 1e-9)) below = 0;
        else {      // the reference's association: sweep 1 then sweep 2 left to right (`step` adds the two sums, `solve2Body` runs one through)
          atomicAdd(&d.work[2], 1ull);
          double p1 = 0;
          for (int gi = 0; gi < gn; gi++) {
            const QTeamRec rc = recs[epos[gi]];
            for (int x = 0; x < rc.jn + rc.cn; x++) p1 = p1 + *dl_ref(rc, 1, x);
          }
          double p2 = (gn == 1) ? p1 : 0;
          for (int gi = 0; gi < gn; gi++) {
            const QTeamRec rc = recs[epos[gi]];
            for (int x = 0; x < 2 * rc.cn; x++) p2 = p2 + *dl_ref(rc, 2, x);
          }
          const double deltaTot = (gn == 1) ? p2 : p1 + p2;
          below = (deltaTot - kSolverTolerance < 0) ? 1 : 0;
        }
        s_decide = below;
      }
      sync();
      const int below = s_decide;
      sync();
      if (below) { rem = remaining - 1; break; }
      remaining--;
    }
    if (tid == 0) {
      atomicMin(&d.world_min_rem[w], rem); d.body_iters[d.isl_root[isl]] = iterations - rem;
      atomicAdd(&d.work[0], (unsigned long long)d.isl_nc[isl] * (iterations - rem));
      if (d.isl_nj[isl]) atomicAdd(&d.work[1], (unsigned long long)d.isl_nj[isl] * (iterations - rem));
    }
    // solved normal lambdas to the contacts (next frame's warm start, Solver.elm:320-335); the dynamic bodies' deltas to the solver-body array
    for (int i = tid; i < gn; i += T) {
      const QTeamRec rc = recs[i];
      for (int x = 0; x < rc.cn; x++) d.cur.contacts[rc.c0 + x].lam = row_ptr(rc, 1, rc.jn + x)->lam;
      const int s = d.isl_groups[g0 + rc.enum_pos];
      const int rr[2] = {d.cur.slot_row1[s], d.cur.slot_row2[s]};
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int b = e ? rc.lb2 : rc.lb1;
        if (b == 0) continue;
        const double* z = body + 8 * (size_t)b;
        double* o = d.sbv + 6 * (size_t)rr[e];
        o[0] = z[0]; o[1] = z[1]; o[2] = z[2]; o[3] = z[4]; o[4] = z[5]; o[5] = z[6];
      }
    }
    sync();
  }
}

}  // namespace ep
