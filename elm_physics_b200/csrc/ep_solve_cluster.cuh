// k_solve_cluster : ONE GIANT ISLAND ON A THREAD-BLOCK CLUSTER (BASELINE config 4's pile: 8.4 k contacts in one island; config 2's
// 100-box pyramid: 762).  k_solve_team<256> ran such an island on one SM, a thread per pair group: 125 threads each pulling their own
// 464 B row records through the L1 of ONE SM every level of every iteration (9.8 of the pile's 11.1 ms per step, 147 SMs idle).
//
// Here the island runs on a cluster of up to 16 CTAs (16 SMs, co-scheduled; hardware cluster barrier ~400 cycles):
//   * the level schedule is the one of k_solve_team (level(g) = 1 + max(level of the previous group touching either dynamic body):
//     groups of a level commute exactly; levels ascending for the sweeps, descending for the warm start) and is built by CTA 0;
//   * a pair group is solved by a QUAD of lanes (lane role = (body of the pair) x (linear, angular), partial dots combined by
//     shuffles in the reference's association ((p0 + p1) + p2) + p3, Solver.elm:642-646 -- the arithmetic of ep_solve_quads.cuh,
//     read straight from k_equations' RowRec: lane 0 negates vB, which negates its partial exactly);
//   * the groups of a level are dealt over the cluster's warps first (one busy quad per warp while a level has fewer groups than
//     the cluster has warps), so the row traffic of a level is spread over 16 L1s / LSUs instead of one;
//   * PRODUCER / CONSUMER warps: a lone warp issues one instruction per ~4 cycles, so the time of a row is its instruction
//     count (first version, every quad fetching its own rows from RowRec: ~840 cycles per row, C2 8.0 ms, C4 6.9 ms).  Each
//     consumer warp therefore has a producer warp that, one step ahead (a step = one level of one sweep), reads the rows of the
//     consumer's next groups from k_equations' records (L2), lays them out per lane role -- coefficients with the sign folded
//     in, bounds (+-mu lambda_n for friction rows), the address of the row's lambda -- in a double-buffered shared-memory
//     window; the consumer's row is six LDS.128, the arithmetic, two stores.  The L2 round trip of the rows -- what bounded both
//     team kernels -- is off the chain; the group records run two steps ahead in registers;
//   * the solver-body deltas live in a table DISTRIBUTED over the shared memories of the cluster (body b of the island in CTA
//     b mod NC; ld / st.shared::cluster, ~215 cycles against ~700 measured for the L2 round trip), and the consumer writes a
//     row's new lambda and |dlambda| back into its window slot, from where the PRODUCER flushes them to the row records one step
//     later: the consumer has no global store in flight when it arrives at the cluster barrier, so the barrier's release fence
//     (MEMBAR.ALL.GPU) does not wait for an L2 acknowledgement on the critical path.  Islands whose bodies do not fit the tables
//     keep them in the global solver-body array (ld.cg / st.cg);
//   * CLUSTER mode (islands of thousands of groups): a step ends with barrier.cluster arrive(release) / wait(acquire).  SOLO mode
//     (islands of <= 1024 groups: C2's pyramid): one CTA, one __syncthreads per step.
// sum|dlambda| : per-CTA partial sums are posted to CTA 0's shared memory (DSMEM) in front of the last barrier of an iteration,
// every thread of the cluster adds the same partials in the same order -> the same decision everywhere without a broadcast; the
// canonical-order re-summation when the sum lands within 1e-9 relative of the tolerance is the team kernel's.
#pragma once
#include <cstddef>
#include "ep_solve.cuh"

namespace ep {

__device__ __forceinline__ unsigned cl_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_count() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ long long cl_clock() { long long x; asm volatile("mov.u64 %0, %%clock64;" : "=l"(x) :: "memory"); return x; }
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared-memory address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ unsigned cl_map(const void* p, unsigned rank) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p); unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_f64(unsigned addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double cl_ld_f64(unsigned addr) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ double2 cl_ld_v2f64(unsigned addr) { double2 v; asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void cl_st_v2f64(unsigned addr, double a, double b) { asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(a), "d"(b) : "memory"); }
__device__ __forceinline__ int cl_ld_s32(unsigned addr) { int v; asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_drain() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kRowChunks = (int)(sizeof(RowRec) / 16);      // 29 chunks of 16 B
static_assert(sizeof(RowRec) % 16 == 0 && sizeof(TeamRec) == 32, "row / group records are copied in 16 B chunks");
// double offsets inside RowRec / JointRow
constexpr int kR_nJ = offsetof(RowRec, nJ) / 8, kR_nB = offsetof(RowRec, nB) / 8, kR_nInvC = offsetof(RowRec, nInvC) / 8, kR_lamN = offsetof(RowRec, lamN) / 8;
constexpr int kR_f1J = offsetof(RowRec, f1J) / 8, kR_f2J = offsetof(RowRec, f2J) / 8, kR_f1B = offsetof(RowRec, f1B) / 8, kR_f1InvC = offsetof(RowRec, f1InvC) / 8;
constexpr int kR_f2B = offsetof(RowRec, f2B) / 8, kR_f2InvC = offsetof(RowRec, f2InvC) / 8, kR_mu = offsetof(RowRec, mu) / 8;
constexpr int kR_lamF1 = offsetof(RowRec, lamF1) / 8, kR_lamF2 = offsetof(RowRec, lamF2) / 8;
constexpr int kR_dN = offsetof(RowRec, dN) / 8, kR_dF1 = offsetof(RowRec, dF1) / 8, kR_dF2 = offsetof(RowRec, dF2) / 8;
constexpr int kR_nIa = offsetof(RowRec, nIa) / 8, kR_nIb = offsetof(RowRec, nIb) / 8, kR_f1Ia = offsetof(RowRec, f1Ia) / 8, kR_f1Ib = offsetof(RowRec, f1Ib) / 8;
constexpr int kR_f2Ia = offsetof(RowRec, f2Ia) / 8, kR_f2Ib = offsetof(RowRec, f2Ib) / 8;
constexpr int kJ_J = offsetof(JointRow, J) / 8, kJ_B = offsetof(JointRow, B) / 8, kJ_invC = offsetof(JointRow, invC) / 8, kJ_eps = offsetof(JointRow, eps) / 8;
constexpr int kJ_lam = offsetof(JointRow, lam) / 8, kJ_dl = offsetof(JointRow, dl) / 8, kJ_Ia = offsetof(JointRow, Ia) / 8, kJ_Ib = offsetof(JointRow, Ib) / 8;

// shared-memory window slot of one row-solve: per lane role q 48 B (a0 a1 a2 c0 c1 c2) at 48 q, then (B invC | lambda lo | hi, address of
// the row's lambda); the consumer overwrites (lambda lo) with (new lambda, |dlambda|), which the producer flushes to the row record
constexpr int kSlotBytes = 240, kSlotShared = 192;
constexpr int kClusterWarps = 4;         // consumer warps per CTA (one per SM sub-partition), each with its producer warp
constexpr int kClusterThreads = 2 * kClusterWarps * 32;
constexpr int kSoloGroupsMax = 1024;     // islands up to this many pair groups whose dynamic bodies fit the table run SOLO
// shared memory of a CTA: NW consumer warps x 2 windows x ws slots (the schedule build of CTA 0 aliases them), then the CTA's share
// of the body table (64 B per body: v | invMass | w | 1) and the bodies' global rows
__host__ __device__ inline size_t cluster_smem_bytes(int nw, int ws, int cap_bodies) {
  const size_t win = (size_t)nw * 2 * ws * kSlotBytes, walk = 3 * 2048 * sizeof(int) + 4096;
  return (win > walk ? win : walk) + (size_t)cap_bodies * 68 + 32;
}

template <int NW, bool SOLO>
__global__ void __launch_bounds__(2 * NW * 32, 1) k_solve_cluster(Dev d, int bin_lo, int bin_hi, int force_resum, int ws, int cap_bodies, int solo_bodies) {
  constexpr int T = 2 * NW * 32, CH = 2048, LCAP = 2048;
  static_assert(NW <= 8, "s_part / s_desc sizes");
  extern __shared__ __align__(16) unsigned char cl_smem[];
  __shared__ int s_nlev, s_decide, s_changed;
  __shared__ int s_ps[T];               // partial sums of the schedule build's scan
  __shared__ double s_part[NW];
  __shared__ double s_cpart[16];        // CTA 0's copy receives the partial sum|dlambda| of every CTA of the cluster
  __shared__ int s_loff[LCAP + 1];
  __shared__ int4 s_desc[2][NW][8];     // producer -> consumer, per step parity: (first window slot | -1, row-solves, body 1, body 2) of each quad's group
  __shared__ int s_lvn[2][NW];          // ... and the number of groups of the step's level
  __shared__ int s_wflush[2][NW], s_wdlt[2][NW];      // slots of a window whose results wait for the flush, offset of |dlambda| from lambda
  __shared__ int s_c0[2][NW][64];          // producer: first contact of the group whose window slots start at [slot]
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, qb = lane & ~3, k = lane >> 2;
  const bool producer = warp >= NW;
  const int cw = producer ? warp - NW : warp;           // the consumer warp this warp is, or serves
  const int rank = SOLO ? 0 : (int)cl_rank(), NC = SOLO ? 1 : (int)cl_size(), cid = SOLO ? (int)blockIdx.x : (int)cl_id(), ncl = SOLO ? (int)gridDim.x : (int)cl_count();
  const int ncs = 31 - __clz(NC);             // NC is a power of two
  const int NWT = NW * NC;                    // consumer warps of the cluster
  const int wg = cw * NC + rank;              // this one among them: consecutive groups of a level go to different CTAs
  const int slot0 = k * NWT + wg;             // this quad's group of a level (round 0); further rounds follow at stride 8 NWT
  const int stride = 8 * NWT;
  const size_t win_all = (size_t)NW * 2 * ws * kSlotBytes, walk_b = 3 * (size_t)CH * sizeof(int) + 4096;
  const size_t front = ((win_all > walk_b ? win_all : walk_b) + 15) & ~(size_t)15;
  int* s_a = (int*)cl_smem; int* s_b = s_a + CH; int* s_lv = s_b + CH; int* s_last = s_lv + CH;      // schedule build (CTA 0), before the windows are in use
  const int last_cap = (int)((front - 3 * (size_t)CH * sizeof(int)) / sizeof(int));
  unsigned char* win = cl_smem + (size_t)cw * 2 * ws * kSlotBytes;
  double* s_body = (double*)(cl_smem + front);                 // [cap_bodies][8]: this CTA's share of the island's body table
  int* s_brow = (int*)(s_body + (size_t)cap_bodies * 8);       // [cap_bodies] their global rows
  // the lane's view of a row record: dot coefficients a (lane 0: -vB), update coefficients c, the sign folded in at load time
  const double sg = q == 0 ? -1.0 : 1.0;
  const int ja = q == 1 ? 0 : (q == 3 ? 6 : 3);        // J = (wA | vB | wB)
  const bool angular = (q & 1) != 0;
  auto step_wait = [&]() { if (SOLO) __syncthreads(); else cl_wait(); };
  // The release of barrier.cluster.arrive is MEMBAR.ALL.GPU: it waits for every outstanding memory operation of the warp, LOADS
  // included.  The consumers need it (their body stores must be visible in the other CTAs).  A producer publishes nothing outside
  // its CTA during a step -- the window is read by its own consumer, the flushed results by itself -- so it orders its stores with
  // a CTA-scope fence BEFORE it starts the next step's row loads and arrives relaxed: the loads stay in flight across the barrier
  // (ncu before: 37 % of all warp samples of the pile's launch sat in the arrive's fence).
  auto step_arrive = [&]() { if (!SOLO) { if (producer) cl_arrive_relaxed(); else cl_arrive(); } };
  auto all_sync = [&]() { if (SOLO) __syncthreads(); else { cl_arrive(); cl_wait(); } };
  int total = 0;
  for (int b = bin_lo; b <= bin_hi; b++) total += d.cur.counters[CNT_BIN0 + b];
  for (int item = cid; item < total; item += ncl) {
    const int isl = island_from_worklist(d, item, bin_lo, bin_hi);
    const int g0 = d.isl_off[isl], gn = d.isl_cnt[isl], w = d.isl_world[isl], nb = d.isl_nb[isl];
    const int r0 = d.world_body_off[w];
    const bool solo_ok = gn <= kSoloGroupsMax && nb + 1 <= solo_bodies;       // the SOLO launch takes these, the cluster launch the others
    if (solo_ok != SOLO) continue;
#ifdef EP_CL_TRACE
    const long long tr_i0 = cl_clock();
#endif
    // bodies in the distributed table (numbered inside the island, 0 = the stand-in of every non-dynamic body), or, when they do
    // not fit, in the global solver-body array (global row, -1)
    const bool tab = nb + 1 <= NC * cap_bodies;
    const bool sl = nb + 1 <= last_cap;              // the level walk's per-body array in shared memory (else global scratch)
    int* lvl = d.sched_lvl + g0; int* sched = d.sched + g0; int* loff = d.lvl_off + g0; int* lcur = d.lvl_cur + g0;
    TeamRec* recs = (TeamRec*)d.sched_rec + g0;
    auto body_ref = [&](int r) -> int { return d.kind[r] == 2 ? (tab ? d.body_local[r] : r) : (tab ? 0 : -1); };
    // ---- the schedule, by CTA 0 (as k_solve_team; the walk's per-body "last level" sits in shared memory) ---------------------
    // level(g) = max over the two dynamic bodies of g of (1 + level of the previous group touching that body), 0 without one: a
    // longest-path problem with at most two predecessors per group.  In parallel when the tables fit the front of the shared memory:
    // the groups of every body (count / scan / fill, unordered), every group's predecessor per body = the largest earlier group in
    // that body's list, then rounds of lv[g] = max(lv[pred] + 1) until nothing changes (one round per level at most; values only
    // grow towards the unique fixpoint, so racing reads only delay).  The one-thread walk below did this at ~80 cycles per group
    // (config 4: 5.2 k groups, 0.2 ms per step) and stays as the fallback.
    const bool sched_par = !d.sched_serial && ((size_t)3 * gn + 2 * ((size_t)nb + 2)) * sizeof(int) <= front;
    if (rank == 0 && sched_par) {
      volatile int* p_lv = (volatile int*)cl_smem;      // [gn]
      int* p_pred = (int*)cl_smem + gn;                 // [2 gn] predecessor of group g through its body 1 / body 2, -1 = none
      int* p_off = p_pred + 2 * gn;                     // [nb + 2] group count, then list offset of body 1 .. nb (0 = the stand-in)
      int* p_cur = p_off + nb + 2;                      // [nb + 2]
      int* g_ab = (int*)recs;                           // global scratch inside the island's (not yet written) records: [2 gn] body refs
      int* g_adj = g_ab + 2 * gn;                       //   [2 gn] the bodies' group lists
      auto walk_ref = [&](int r) -> int { return d.kind[r] == 2 ? d.body_local[r] : 0; };
      for (int x = tid; x < nb + 2; x += T) { p_off[x] = 0; p_cur[x] = 0; }
      for (int gi = tid; gi < gn; gi += T) { loff[gi] = 0; p_lv[gi] = 0; }
      if (tid == 0) s_nlev = 0;
      __syncthreads();
      for (int gi = tid; gi < gn; gi += T) {
        const int s = d.isl_groups[g0 + gi];
        const int a = walk_ref(d.cur.slot_row1[s]), b = walk_ref(d.cur.slot_row2[s]);
        g_ab[2 * gi] = a; g_ab[2 * gi + 1] = b;
        if (a) atomicAdd(&p_off[a], 1);
        if (b) atomicAdd(&p_off[b], 1);
      }
      __syncthreads();
      {   // exclusive scan of the counts: a contiguous run of bodies per thread, the T partial sums by warp 0
        const int per = (nb + 2 + T - 1) / T, x0 = min(nb + 2, tid * per), x1 = min(nb + 2, x0 + per);
        int sum = 0;
        for (int x = x0; x < x1; x++) sum += p_off[x];
        s_ps[tid] = sum;
        __syncthreads();
        if (tid < 32) {
          constexpr int E = T / 32;
          int v[E], tot = 0;
#pragma unroll
          for (int k2 = 0; k2 < E; k2++) { v[k2] = s_ps[tid * E + k2]; tot += v[k2]; }
          int incl = tot;
          for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += u; }
          int run = incl - tot;
#pragma unroll
          for (int k2 = 0; k2 < E; k2++) { s_ps[tid * E + k2] = run; run += v[k2]; }
        }
        __syncthreads();
        int run = s_ps[tid];
        for (int x = x0; x < x1; x++) { const int c = p_off[x]; p_off[x] = run; run += c; }
      }
      __syncthreads();
      for (int gi = tid; gi < gn; gi += T) {
        const int a = g_ab[2 * gi], b = g_ab[2 * gi + 1];
        if (a) g_adj[p_off[a] + atomicAdd(&p_cur[a], 1)] = gi;
        if (b) g_adj[p_off[b] + atomicAdd(&p_cur[b], 1)] = gi;
      }
      __syncthreads();
      for (int gi = tid; gi < gn; gi += T) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int x = g_ab[2 * gi + e];
          int pred = -1;
          if (x) {
            const int lo = p_off[x], hi = lo + p_cur[x];
#pragma unroll 4
            for (int k2 = lo; k2 < hi; k2++) { const int o = g_adj[k2]; if (o < gi && o > pred) pred = o; }
          }
          p_pred[2 * gi + e] = pred;
        }
      }
      __syncthreads();
      for (;;) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        bool ch = false;
        for (int gi = tid; gi < gn; gi += T) {
          const int pa = p_pred[2 * gi], pb = p_pred[2 * gi + 1];
          const int la = pa >= 0 ? p_lv[pa] + 1 : 0, lb = pb >= 0 ? p_lv[pb] + 1 : 0;
          const int l = max(la, lb);
          if (l != p_lv[gi]) { p_lv[gi] = l; ch = true; }
        }
        if (ch) s_changed = 1;
        __syncthreads();
        const int again = s_changed;
        __syncthreads();
        if (!again) break;
      }
      int mx = 0;
      for (int gi = tid; gi < gn; gi += T) { const int l = p_lv[gi]; lvl[gi] = l; atomicAdd(&loff[l], 1); mx = max(mx, l + 1); }
      atomicMax(&s_nlev, mx);
      __syncthreads();
    }
    if (rank == 0 && !sched_par) {
      int* last = sl ? s_last : d.uf_fill + r0 - 1;      // 1-based: the body's index inside the island, or inside the world
      auto walk_ref = [&](int r) -> int { return d.kind[r] == 2 ? (sl ? d.body_local[r] : r - r0 + 1) : 0; };
      for (int gi = tid; gi < gn; gi += T) {
        loff[gi] = 0;
        const int s = d.isl_groups[g0 + gi];
        const int a = walk_ref(d.cur.slot_row1[s]), b = walk_ref(d.cur.slot_row2[s]);
        if (a) last[a] = 0;
        if (b) last[b] = 0;
      }
      if (tid == 0) s_nlev = 0;
      __syncthreads();
      for (int base = 0; base < gn; base += CH) {
        const int m = min(CH, gn - base);
        for (int i = tid; i < m; i += T) {
          const int s = d.isl_groups[g0 + base + i];
          s_a[i] = walk_ref(d.cur.slot_row1[s]); s_b[i] = walk_ref(d.cur.slot_row2[s]);
        }
        __syncthreads();
        if (tid == 0) {      // the serial part: one dependent shared-memory round trip per group (the next pair is read ahead)
          int nlev = s_nlev;
          int a = s_a[0], b = s_b[0];
          for (int i = 0; i < m; i++) {
            const int na = s_a[min(i + 1, m - 1)], nb2 = s_b[min(i + 1, m - 1)];
            const int la = a ? last[a] : 0, lb = b ? last[b] : 0;
            const int l = max(la, lb);
            s_lv[i] = l;
            if (a) last[a] = l + 1;
            if (b) last[b] = l + 1;
            nlev = max(nlev, l + 1);
            a = na; b = nb2;
          }
          s_nlev = nlev;
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) { lvl[base + i] = s_lv[i]; atomicAdd(&loff[s_lv[i]], 1); }
        __syncthreads();
      }
    }
    if (rank == 0) {
      const int nl = s_nlev;
      if (tid == 0) { int run = 0; for (int l = 0; l < nl; l++) { const int c = loff[l]; loff[l] = run; lcur[l] = run; run += c; } }
      __syncthreads();
      for (int gi = tid; gi < gn; gi += T) {
        const int s = d.isl_groups[g0 + gi];
        const int at = atomicAdd(&lcur[lvl[gi]], 1);
        sched[at] = s;
        const int r1 = d.cur.slot_row1[s], r2 = d.cur.slot_row2[s];
        TeamRec rc;
        rc.b1 = body_ref(r1); rc.b2 = body_ref(r2);
        rc.c0 = d.cur.slot_cstart[s]; rc.cn = max(d.cur.slot_ccount[s], 0); rc.j0 = d.cur.slot_jstart[s]; rc.jn = max(d.cur.slot_jcount[s], 0);
        rc.pad0 = r1; rc.pad1 = r2;
        recs[at] = rc;
      }
      __threadfence();
    }
    __syncthreads();
    if (!SOLO) { cl_arrive(); cl_wait(); }
    const int nlev = SOLO ? s_nlev : cl_ld_s32(cl_map(&s_nlev, 0));
    const bool lsm = nlev <= LCAP;
#ifdef EP_CL_TRACE
    const long long tr_i1 = cl_clock();
#endif
    if (lsm) { for (int l = tid; l <= nlev; l += T) s_loff[l] = l < nlev ? __ldcg(loff + l) : gn; }
    if (tid < 2 * NW) { s_wflush[tid / NW][tid % NW] = 0; s_wdlt[tid / NW][tid % NW] = 0; }
    if (tab) {       // this CTA's share of the body table: body b of the island lives in CTA b mod NC, entry b / NC
      for (int i = tid; i < gn; i += T) {
        const int4 x = __ldcg((const int4*)(recs + i)), y = __ldcg((const int4*)(recs + i) + 1);
        const int bb[2] = {x.x, x.y}, rr[2] = {y.z, y.w};
#pragma unroll
        for (int e = 0; e < 2; e++) {
          if (bb[e] == 0 || (bb[e] & (NC - 1)) != rank) continue;
          double* z = s_body + 8 * (size_t)(bb[e] >> ncs); const double* sb = d.sbv + 6 * (size_t)rr[e];      // several groups may write the same body: same values
          z[0] = sb[0]; z[1] = sb[1]; z[2] = sb[2]; z[3] = d.inv_mass[rr[e]]; z[4] = sb[3]; z[5] = sb[4]; z[6] = sb[5]; z[7] = 1.0;
          s_brow[bb[e] >> ncs] = rr[e];
        }
      }
    }
    __syncthreads();
    auto level_begin = [&](int l) -> int { return lsm ? s_loff[l] : (l < nlev ? __ldcg(loff + l) : gn); };
    // a group's results reach its row records one step after it was solved (the producer's flush) and its rows are read two steps
    // before it is solved again (the producer's loads run a step ahead of its stores): with four levels or more the flush is always
    // a whole step ahead of the loads
    const bool use_win = nlev >= 4 && ws > 0;
    TeamRec none; none.b1 = tab ? 0 : -1; none.b2 = none.b1; none.c0 = 0; none.cn = 0; none.j0 = 0; none.jn = 0; none.pad0 = 0; none.pad1 = 0;
    auto rec_at = [&](int l, int slot) -> TeamRec {
      const int lb = level_begin(l), n = level_begin(l + 1) - lb;
      if (slot >= n) return none;
      const int4* p = (const int4*)(recs + lb + slot);
      const int4 x = __ldcg(p), y = __ldcg(p + 1);
      TeamRec rc; rc.b1 = x.x; rc.b2 = x.y; rc.c0 = x.z; rc.cn = x.w; rc.j0 = y.x; rc.jn = y.y; rc.pad0 = 0; rc.pad1 = 0;
      return rc;
    };
    // steps: the warm start walks the levels downwards (phase 0), then sweep 1 / sweep 2 (phases 1 / 2) upwards, alternating
    auto advance = [&](int& ph, int& l) {
      if (ph == 0) { if (l == 0) ph = 1; else l--; }
      else if (l + 1 < nlev) l++;
      else { l = 0; ph = 3 - ph; }
    };
    auto rows_of = [&](int ph, const TeamRec& rc) -> int { return ph == 0 ? rc.cn : (ph == 1 ? rc.jn + rc.cn : 2 * rc.cn); };
    const double eps_c = spook(d.dt[w]).eps;
    double part = 0;
#ifdef EP_CL_TRACE
    long long tr_rows = 0, tr_body = 0, tr_loop = 0, tr_x = 0, tr_y = 0, tr_z = 0, tr_p1 = 0, tr_p2 = 0, tr_q = 0, tr_a = 0, tr_r0 = 0, tr_r1 = 0, tr_nr = 0, tr_slow = 0;
#endif
    // the lane's half of a solver body: (v | invMass) for the linear lanes, (w | 1) for the angular ones
    auto load_half = [&](int b, bool need, double& s0, double& s1, double& s2, double& m) {
      s0 = 0; s1 = 0; s2 = 0; m = 0;
      if (tab) {
        if (b != 0 && need) {
          const double* p = s_body + 8 * (size_t)(b >> ncs) + (angular ? 4 : 0);
          double2 x, y;
          if (SOLO) { x = ((const double2*)p)[0]; y = ((const double2*)p)[1]; }
          else { const unsigned a = cl_map(p, (unsigned)(b & (NC - 1))); x = cl_ld_v2f64(a); y = cl_ld_v2f64(a + 16); }
          s0 = x.x; s1 = x.y; s2 = y.x; m = y.y;
        }
      } else if (b >= 0 && need) {
        const double* p = d.sbv + 6 * (size_t)b + (angular ? 3 : 0);
        m = angular ? 1.0 : d.inv_mass[b];
        s0 = __ldcg(p); s1 = __ldcg(p + 1); s2 = __ldcg(p + 2);
      }
    };
    auto store_half = [&](int b, bool need, double s0, double s1, double s2) {
      if (!need) return;
      if (tab) {
        if (b == 0) return;
        double* p = s_body + 8 * (size_t)(b >> ncs) + (angular ? 4 : 0);
        if (SOLO) { *(double2*)p = make_double2(s0, s1); p[2] = s2; }
        else { const unsigned a = cl_map(p, (unsigned)(b & (NC - 1))); cl_st_v2f64(a, s0, s1); cl_st_f64(a + 16, s2); }
      } else if (b >= 0) {
        double* p = d.sbv + 6 * (size_t)b + (angular ? 3 : 0);
        __stcg(p, s0); __stcg(p + 1, s1); __stcg(p + 2, s2);
      }
    };
    // ---- consumer, window path: the row-solves of this quad's group from its slots -------------------------------------------
    struct WRow { double a0, a1, a2, c0, c1, c2, B, invC, lam, lo, hi, el, dlo, dhi; unsigned char* at; };
    auto run_group_win = [&](int ph, int n, int b, unsigned char* slots) {
      const int nmax = __reduce_max_sync(full, n);
      if (nmax == 0) return;
#ifdef EP_CL_TRACE
      tr_rows += nmax; tr_x = cl_clock();
#endif
      double s0, s1, s2, m; load_half(b, n > 0, s0, s1, s2, m);
#ifdef EP_CL_TRACE
      if (s0 == 12345.678) tr_rows++;
      tr_y = cl_clock(); tr_body += tr_y - tr_x;
#endif
      const int xl = max(n - 1, 0);                  // reads past the group's rows are clamped to its last slot (a quad without rows reads slot 0 of the window): never used
      // everything of a row that does not hang on the solver bodies is formed here, off the chain
      auto fetch = [&](int x, WRow& o) {
        unsigned char* at = slots + (size_t)min(x, xl) * kSlotBytes;
        const double2* p = (const double2*)(at + q * 48);
        const double2* s = (const double2*)(at + kSlotShared);
        const double2 u = p[0], v = p[1], z = p[2], e = s[0], f = s[1];
        o.a0 = u.x; o.a1 = u.y; o.a2 = v.x; o.c0 = v.y; o.c1 = z.x; o.c2 = z.y;
        o.B = e.x; o.invC = e.y; o.lam = f.x; o.lo = f.y; o.hi = *(const double*)(at + kSlotShared + 32); o.at = at;
        o.el = eps_c * o.lam; o.dlo = o.lo - o.lam; o.dhi = o.hi - o.lam;
      };
      if (ph == 0) {         // applyContactsWarmStart (Solver.elm:102-119); rows with lambda 0 are skipped (41)
        for (int x = 0; x < nmax; x++) {
          WRow t; fetch(x, t);
          if (x < n && t.lam != 0) { const double tm = t.lam * m; s0 = s0 + tm * t.c0; s1 = s1 + tm * t.c1; s2 = s2 + tm * t.c2; }
        }
        store_half(b, n > 0, s0, s1, s2);
        return;
      }
      // one row, branch-free: the candidates of the clamp exist before G.W does, the lane's half is updated by selects, lane 0 of
      // the quad leaves (new lambda, |dlambda|) in the slot under a predicate (every lane of the quad adds |dlambda|; lane 0's sum is kept)
      auto solve_row = [&](int x, const WRow& t) {
        const bool on = x < n;
        const double p = (t.a0 * s0 + t.a1 * s1) + t.a2 * s2;
        const double p0 = __shfl_sync(full, p, qb), p1 = __shfl_sync(full, p, qb + 1), p2 = __shfl_sync(full, p, qb + 2), p3 = __shfl_sync(full, p, qb + 3);
        const double gw = ((p0 + p1) + p2) + p3;
        const double prev = t.invC * (t.B - gw - t.el);
        const double sum = t.lam + prev;
        const bool under = sum - t.lo < 0, over = sum - t.hi > 0;
        const double dl = under ? t.dlo : (over ? t.dhi : prev);
        const double tm = dl * m;
        const double n0 = s0 + tm * t.c0, n1 = s1 + tm * t.c1, n2 = s2 + tm * t.c2;
        s0 = on ? n0 : s0; s1 = on ? n1 : s1; s2 = on ? n2 : s2;
        const double ad = fabs(dl);
        part += on ? ad : 0.0;
        if (on && q == 0) *(double2*)(t.at + kSlotShared + 16) = make_double2(t.lam + dl, ad);
      };
      WRow ra, rb;
      fetch(0, ra);
      int x = 0;
      for (; x + 1 < nmax; x += 2) {
        fetch(x + 1, rb);
        solve_row(x, ra);
        fetch(x + 2, ra);
        solve_row(x + 1, rb);
      }
      if (x < nmax) solve_row(x, ra);
#ifdef EP_CL_TRACE
      if (s0 == 12345.678) tr_rows++;
      tr_z = cl_clock(); tr_loop += tr_z - tr_y;
#endif
      store_half(b, n > 0, s0, s1, s2);
    };
    // ---- consumer, in-place path: rows straight from k_equations' records (groups that are not in the window) ------------------
    struct TRow { double a0, a1, a2, c0, c1, c2, B, invC, lam, eps, lo, hi; double* wl; double* wd; };
    auto run_group = [&](int ph, const TeamRec& rc) {
      const int n = rows_of(ph, rc);
      const int nmax = __reduce_max_sync(full, n);
      if (nmax == 0) return;
      const int b = q < 2 ? rc.b1 : rc.b2;
      double s0, s1, s2, m; load_half(b, n > 0, s0, s1, s2, m);
      RowRec* G0 = d.rows + rc.c0;
      auto fetch = [&](int x, TRow& o) {
        if (ph == 1 && x < rc.jn) {
          double* J = (double*)(d.jrows + rc.j0 + x);
          o.a0 = sg * J[kJ_J + ja]; o.a1 = sg * J[kJ_J + ja + 1]; o.a2 = sg * J[kJ_J + ja + 2];
          o.c0 = o.a0; o.c1 = o.a1; o.c2 = o.a2;
          if (angular) { const int ci = q == 1 ? kJ_Ia : kJ_Ib; o.c0 = J[ci]; o.c1 = J[ci + 1]; o.c2 = J[ci + 2]; }
          o.B = J[kJ_B]; o.invC = J[kJ_invC]; o.eps = J[kJ_eps]; o.lam = __ldcg(J + kJ_lam); o.lo = -kMaxImpulse; o.hi = kMaxImpulse;
          o.wl = J + kJ_lam; o.wd = J + kJ_dl;
          return;
        }
        if (ph <= 1) {
          double* R = (double*)(G0 + (ph == 0 ? x : x - rc.jn));
          o.a0 = sg * R[kR_nJ + ja]; o.a1 = sg * R[kR_nJ + ja + 1]; o.a2 = sg * R[kR_nJ + ja + 2];
          o.c0 = o.a0; o.c1 = o.a1; o.c2 = o.a2;
          if (angular) { const int ci = q == 1 ? kR_nIa : kR_nIb; o.c0 = R[ci]; o.c1 = R[ci + 1]; o.c2 = R[ci + 2]; }
          o.B = R[kR_nB]; o.invC = R[kR_nInvC]; o.lam = __ldcg(R + kR_lamN); o.eps = eps_c; o.lo = 0.0; o.hi = kMaxImpulse;
          o.wl = R + kR_lamN; o.wd = R + kR_dN;
          return;
        }
        const int f = x & 1;
        double* R = (double*)(G0 + (x >> 1));
        const int jb = f ? kR_f2J : kR_f1J;
        o.a0 = sg * R[jb + ja]; o.a1 = sg * R[jb + ja + 1]; o.a2 = sg * R[jb + ja + 2];
        o.c0 = o.a0; o.c1 = o.a1; o.c2 = o.a2;
        if (angular) { const int ci = f ? (q == 1 ? kR_f2Ia : kR_f2Ib) : (q == 1 ? kR_f1Ia : kR_f1Ib); o.c0 = R[ci]; o.c1 = R[ci + 1]; o.c2 = R[ci + 2]; }
        o.B = R[f ? kR_f2B : kR_f1B]; o.invC = R[f ? kR_f2InvC : kR_f1InvC]; o.lam = __ldcg(R + (f ? kR_lamF2 : kR_lamF1)); o.eps = eps_c;
        const double cap = R[kR_mu] * __ldcg(R + kR_lamN);
        o.lo = -cap; o.hi = cap;
        o.wl = R + (f ? kR_lamF2 : kR_lamF1); o.wd = R + (f ? kR_dF2 : kR_dF1);
      };
      auto solve_row = [&](int x, const TRow& t) {
        const bool on = x < n;
        if (ph == 0) {
          if (on && t.lam != 0) { const double tm = t.lam * m; s0 = s0 + tm * t.c0; s1 = s1 + tm * t.c1; s2 = s2 + tm * t.c2; }
          return;
        }
        const double p = (t.a0 * s0 + t.a1 * s1) + t.a2 * s2;
        const double p0 = __shfl_sync(full, p, qb), p1 = __shfl_sync(full, p, qb + 1), p2 = __shfl_sync(full, p, qb + 2), p3 = __shfl_sync(full, p, qb + 3);
        const double gw = ((p0 + p1) + p2) + p3;
        const double prev = t.invC * (t.B - gw - t.eps * t.lam);
        const double dl = (t.lam + prev - t.lo < 0) ? t.lo - t.lam : ((t.lam + prev - t.hi > 0) ? t.hi - t.lam : prev);
        if (on) {
          const double tm = dl * m;
          s0 = s0 + tm * t.c0; s1 = s1 + tm * t.c1; s2 = s2 + tm * t.c2;
          if (q == 0) { *t.wl = t.lam + dl; *t.wd = eabs(dl); part += eabs(dl); }
        }
      };
      TRow ra;
      ra.a0 = ra.a1 = ra.a2 = ra.c0 = ra.c1 = ra.c2 = ra.B = ra.invC = ra.lam = ra.eps = ra.lo = ra.hi = 0; ra.wl = nullptr; ra.wd = nullptr;
      for (int x = 0; x < nmax; x++) {
        if (x < n) fetch(x, ra);
        solve_row(x, ra);
      }
      store_half(b, n > 0, s0, s1, s2);
    };
    // ---- producer: the results the consumer left in window `buf` (new lambda, |dlambda| per slot) to the row records ------------
    auto flush = [&](int buf) {
      const int nf = s_wflush[buf][cw], dlt = s_wdlt[buf][cw];
      const unsigned char* wb = win + (size_t)buf * ws * kSlotBytes;
      for (int i = lane; i < nf; i += 32) {
        const double2 r = *(const double2*)(wb + (size_t)i * kSlotBytes + kSlotShared + 16);
        double* wl = (double*)__double_as_longlong(*(const double*)(wb + (size_t)i * kSlotBytes + kSlotShared + 40));
        wl[0] = r.x; wl[dlt] = r.y;
      }
      __syncwarp();
    };
    // ---- producer: the consumer's round-0 groups of one step: window slots in quad order while they fit (groups with joint rows and
    // what does not fit are solved in place), the rows laid out per lane role, and the consumer's descriptors.  In two halves a step
    // apart: `issue` allots the slots and LOADS the rows into registers (the L2 round trip, ~700 cycles, runs beside the producer's
    // other work and the barrier), `commit` forms the slots from them and stores slots and descriptors into the window ----------
    constexpr int PF = 2;                                    // slots per lane that travel in registers (8 PF per warp); more are loaded at commit time
    struct Pend { int ph, n_lvl, tot, par; unsigned Ml, Mh; int4 desc; int slot[PF]; const double* R[PF]; double v[PF][11]; } pd;
    pd.ph = 0; pd.n_lvl = 0; pd.tot = 0; pd.par = 0; pd.Ml = 0; pd.Mh = 0; pd.desc = make_int4(-1, 0, 0, 0);
    // row record and field offsets of window slot i of a step (M: the slots at which a group starts; c0 table of the step)
    auto slot_src = [&](int ph, unsigned Ml, unsigned Mh, const int* c0tab, int i, int& jb, int& ci, int& li, int& bi, int& ici) -> const double* {
      const unsigned mh = i >= 32 ? (Mh & ((2u << (i - 32)) - 1u)) : 0u, ml = i >= 32 ? Ml : (Ml & ((2u << i) - 1u));
      const int sb = mh ? 63 - __clz(mh) : 31 - __clz(ml);      // the group this slot belongs to starts at slot sb
      const int x = i - sb;
      const int f = ph == 2 ? (x & 1) : 0;
      jb = ph == 2 ? (f ? kR_f2J : kR_f1J) : kR_nJ;
      ci = !angular ? jb + ja : (ph == 2 ? (f ? (q == 1 ? kR_f2Ia : kR_f2Ib) : (q == 1 ? kR_f1Ia : kR_f1Ib)) : (q == 1 ? kR_nIa : kR_nIb));
      li = ph == 2 ? (f ? kR_lamF2 : kR_lamF1) : kR_lamN;
      bi = ph == 2 ? (f ? kR_f2B : kR_f1B) : kR_nB; ici = ph == 2 ? (f ? kR_f2InvC : kR_f1InvC) : kR_nInvC;
      return (const double*)(d.rows + c0tab[sb] + (ph == 2 ? x >> 1 : x));
    };
    // every load of a slot is issued before any is used: one L2 round trip (all four lanes read the row's scalars, lane 0 keeps them)
    auto slot_load = [&](const double* R, int jb, int ci, int li, int bi, int ici, double* v) {
      v[0] = R[jb + ja]; v[1] = R[jb + ja + 1]; v[2] = R[jb + ja + 2];
      v[3] = R[ci]; v[4] = R[ci + 1]; v[5] = R[ci + 2];
      v[6] = R[bi]; v[7] = R[ici]; v[8] = __ldcg(R + li); v[9] = R[kR_mu]; v[10] = __ldcg(R + kR_lamN);
    };
    auto slot_store = [&](int ph, unsigned char* wb, int i, const double* v, const double* lamp) {
      const double a0 = sg * v[0], a1 = sg * v[1], a2 = sg * v[2];
      const double c0 = angular ? v[3] : a0, c1 = angular ? v[4] : a1, c2 = angular ? v[5] : a2;
      double2* o = (double2*)(wb + (size_t)i * kSlotBytes + q * 48);
      o[0] = make_double2(a0, a1); o[1] = make_double2(a2, c0); o[2] = make_double2(c1, c2);
      const double cap = v[9] * v[10];
      const double lo = ph == 2 ? -cap : 0.0, hi = ph == 2 ? cap : kMaxImpulse;
      if (q == 0) {
        double2* s = (double2*)(wb + (size_t)i * kSlotBytes + kSlotShared);
        s[0] = make_double2(v[6], v[7]); s[1] = make_double2(v[8], lo); s[2] = make_double2(hi, __longlong_as_double((long long)lamp));
      }
    };
    auto issue = [&](int ph, int l, const TeamRec& rc, int par) {
      const int n = rows_of(ph, rc);
      const bool stg = use_win && n > 0 && !(ph == 1 && rc.jn > 0);
      int incl = (q == 0 && stg) ? n : 0;            // inclusive scan over the quad leaders (lanes 0, 4, .., 28)
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) { const int up = __shfl_up_sync(full, incl, o); if (lane >= o) incl += up; }
      incl = __shfl_sync(full, incl, qb);
      const bool fits = stg && incl <= ws;
      const int base = incl - n;
      pd.ph = ph; pd.par = par;
      pd.desc = make_int4(fits ? base : -1, n, rc.b1, rc.b2);
      pd.Ml = __reduce_or_sync(full, (q == 0 && fits && base < 32) ? (1u << base) : 0u);      // slots at which a group starts (ws <= 64)
      pd.Mh = __reduce_or_sync(full, (q == 0 && fits && base >= 32) ? (1u << (base - 32)) : 0u);
      pd.tot = __reduce_max_sync(full, fits ? incl : 0);
      pd.n_lvl = level_begin(l + 1) - level_begin(l);
      if (q == 0 && fits) s_c0[par][cw][base] = rc.c0;
      __syncwarp();
#ifdef EP_CL_TRACE
      tr_rows += (pd.tot + 7) / 8;
#endif
#pragma unroll
      for (int j = 0; j < PF; j++) {
        const int i = k + 8 * j;
        pd.slot[j] = -1;
        if (i < pd.tot) {
          int jb, ci, li, bi, ici;
          const double* R = slot_src(ph, pd.Ml, pd.Mh, s_c0[par][cw], i, jb, ci, li, bi, ici);
          slot_load(R, jb, ci, li, bi, ici, pd.v[j]);
          pd.slot[j] = i; pd.R[j] = R + li;
        }
      }
    };
    auto commit = [&](int buf) {
      if (q == 0) s_desc[buf][cw][k] = pd.desc;
      if (lane == 0) {
        s_lvn[buf][cw] = pd.n_lvl;
        s_wflush[buf][cw] = pd.ph == 0 ? 0 : pd.tot;
        s_wdlt[buf][cw] = pd.ph == 1 ? kR_dN - kR_lamN : kR_dF1 - kR_lamF1;
      }
      unsigned char* wb = win + (size_t)buf * ws * kSlotBytes;
#pragma unroll
      for (int j = 0; j < PF; j++) if (pd.slot[j] >= 0) slot_store(pd.ph, wb, pd.slot[j], pd.v[j], pd.R[j]);
      for (int i = k + 8 * PF; i < pd.tot; i += 8) {          // what did not travel in registers
        int jb, ci, li, bi, ici; double v[11];
        const double* R = slot_src(pd.ph, pd.Ml, pd.Mh, s_c0[pd.par][cw], i, jb, ci, li, bi, ici);
        slot_load(R, jb, ci, li, bi, ici, v);
        slot_store(pd.ph, wb, i, v, R + li);
      }
      __syncwarp();
    };
    const int iterations = d.iterations[w];
    int remaining = iterations, rem = 0;
    int ph_c = 0, l_c = nlev - 1;                    // the current step (consumer) ...
    int ph_i = ph_c, l_i = l_c, ph_n = 0, l_n = 0;   // ... the step the producer issues next (two ahead of the consumer), and the one after it
    TeamRec rc_i = none, rc_n = none;
    if (producer) {                                  // the first step's window and descriptors in front of the first barrier, the second step's rows on their way
      issue(ph_i, l_i, rec_at(l_i, slot0), 0);
      commit(0);
      __threadfence_block();
      advance(ph_i, l_i);
      issue(ph_i, l_i, rec_at(l_i, slot0), 1);
      advance(ph_i, l_i);
      rc_i = rec_at(l_i, slot0);
      ph_n = ph_i; l_n = l_i; advance(ph_n, l_n);
    }
    int t = 0;
    bool decide = false;
    __syncwarp();
    step_arrive();
#ifdef EP_CL_TRACE
    long long tr_pre = 0, tr_wait = 0, tr_work = 0, tr_t0 = cl_clock(), tr_b;
#endif
    for (;;) {
#ifdef EP_CL_TRACE
      tr_a = cl_clock();
#endif
      if (producer) rc_n = rec_at(l_n, slot0);        // in flight until the step after the next is issued
      __syncwarp();
#ifdef EP_CL_TRACE
      tr_b = cl_clock(); tr_pre += tr_b - tr_a;
#endif
      step_wait();                                    // the previous step of every warp is complete and visible
#ifdef EP_CL_TRACE
      tr_a = cl_clock(); tr_wait += tr_a - tr_b;
#endif
      if (decide) {                                   // an iteration has just ended: early exit? (Solver.elm:360-367)
        decide = false;
        if (remaining == 1) { rem = 0; break; }
        double S = 0;
        if (SOLO) { for (int c = 0; c < NW; c++) S += s_part[c]; }
        else { for (int c = 0; c < NC; c++) S += cl_ld_f64(cl_map(&s_cpart[c], 0)); }
        int below;
        if (!force_resum && S < kSolverTolerance * (1.0 - 1e-9)) below = 1;
        else if (!force_resum && S > kSolverTolerance * (1.0 + 1e-9)) below = 0;
        else {
          if (producer) flush((t + 1) & 1);           // the last step's results are still in its window
          __syncwarp();
          all_sync();
          if (rank == 0 && tid == 0) {      // the reference's association: sweep 1 then sweep 2 left to right (`step` adds the two sums, `solve2Body` runs one through)
            atomicAdd(&d.work[2], 1ull);
            double p1 = 0;
            for (int gi = 0; gi < gn; gi++) {
              const int s = d.isl_groups[g0 + gi];
              for (int x = 0, j0 = d.cur.slot_jstart[s]; x < d.cur.slot_jcount[s]; x++) p1 = p1 + __ldcg(&d.jrows[j0 + x].dl);
              for (int x = 0, c0 = d.cur.slot_cstart[s]; x < d.cur.slot_ccount[s]; x++) p1 = p1 + __ldcg(&d.rows[c0 + x].dN);
            }
            double p2 = (gn == 1) ? p1 : 0;
            for (int gi = 0; gi < gn; gi++) {
              const int s = d.isl_groups[g0 + gi];
              for (int x = 0, c0 = d.cur.slot_cstart[s]; x < d.cur.slot_ccount[s]; x++) p2 = p2 + __ldcg(&d.rows[c0 + x].dF1) + __ldcg(&d.rows[c0 + x].dF2);
            }
            const double deltaTot = (gn == 1) ? p2 : p1 + p2;
            s_decide = (deltaTot - kSolverTolerance < 0) ? 1 : 0;
          }
          __syncwarp();
          all_sync();
          below = SOLO ? s_decide : cl_ld_s32(cl_map(&s_decide, 0));
        }
        if (below) { rem = remaining - 1; break; }
        remaining--;
      }
      if (producer) {
        flush((t + 1) & 1);                           // step t-1's results; then that window receives step t+1, whose rows were loaded a step ago,
        commit((t + 1) & 1);
        __threadfence_block();
        issue(ph_i, l_i, rc_i, t & 1);                // and the loads of step t+2 start
        ph_i = ph_n; l_i = l_n; rc_i = rc_n; advance(ph_n, l_n);
      } else {
        // ---- the step: this quad's group of level l_c (round 0) from the window; what is not there and further rounds in place ----
        const int4 ds = s_desc[t & 1][cw][k];
        const int n_lvl = s_lvn[t & 1][cw];
        const bool slow = __any_sync(full, ds.y > 0 && ds.x < 0);
#ifdef EP_CL_TRACE
        if (n_lvl == -12345) tr_rows++;
        tr_q = cl_clock(); tr_p1 += tr_q - tr_a;
#endif
        if (!slow)
          run_group_win(ph_c, ds.x >= 0 ? ds.y : 0, q < 2 ? ds.z : ds.w, win + ((size_t)(t & 1) * ws + max(ds.x, 0)) * kSlotBytes);
        else {       // a group of the warp is not in the window: the whole warp solves in place, and what was staged must not be flushed
          if (lane == 0) s_wflush[t & 1][cw] = 0;
#ifdef EP_CL_TRACE
          tr_slow++;
#endif
          run_group(ph_c, rec_at(l_c, slot0));
        }
#ifdef EP_CL_TRACE
        tr_x = cl_clock(); tr_r0 += tr_x - tr_q;
#endif
        for (int s = wg + stride; s < n_lvl; s += stride)      // warp-uniform: quad 0 holds the warp's smallest slot
          run_group(ph_c, rec_at(l_c, s + k * NWT));
#ifdef EP_CL_TRACE
        tr_r1 += cl_clock() - tr_x; if (wg + stride < n_lvl) tr_nr++;
#endif
      }
      if (ph_c == 2 && l_c == nlev - 1) {             // last step of an iteration: post this CTA's sum|dlambda|
        if (!producer) {
          if (q != 0) part = 0;                         // the window path adds |dlambda| in all four lanes of a quad
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(full, part, o);
          if (lane == 0) s_part[warp] = part;
        }
        if (!SOLO) {
          __syncthreads();
          if (tid == 0) { double S = 0; for (int x = 0; x < NW; x++) S += s_part[x]; cl_st_f64(cl_map(&s_cpart[rank], 0), S); }
        }
        part = 0;
        decide = true;
      }
      __syncwarp();
#ifdef EP_CL_TRACE
      tr_b = cl_clock(); tr_work += tr_b - tr_a;
#endif
      step_arrive();
      advance(ph_c, l_c);
      t++;
    }
#ifdef EP_CL_TRACE
    if (tid == 0 && rank == 0) printf("[cl] build %lld cycles, to loop start %lld, slow-path steps of warp 0: %lld\n", tr_i1 - tr_i0, tr_t0 - tr_i0, tr_slow);
    if (lane == 0 && rank == 0) printf("[cl] isl %d gn %d nlev %d warp %d %s steps %d rows %lld: pre %lld wait %lld work %lld cycles per step; body %lld loop %lld p1 %lld p2 %lld r0 %lld rounds %lld (in %lld steps) per step; total %lld\n", isl, gn, nlev, warp, producer ? "producer" : "consumer", t, tr_rows, tr_pre / max(t, 1), tr_wait / max(t, 1), tr_work / max(t, 1), tr_body / max(t, 1), tr_loop / max(t, 1), tr_p1 / max(t, 1), tr_p2 / max(t, 1), tr_r0 / max(t, 1), tr_r1 / max(t, 1), tr_nr, (long long)(cl_clock() - tr_t0));
#endif
    if (producer) flush((t + 1) & 1);                 // the last step's results
    __syncthreads();
    if (!SOLO) { cl_arrive(); cl_wait(); }
    if (rank == 0 && tid == 0) {
      atomicMin(&d.world_min_rem[w], rem); d.body_iters[d.isl_root[isl]] = iterations - rem;
      atomicAdd(&d.work[0], (unsigned long long)d.isl_nc[isl] * (iterations - rem));
      if (d.isl_nj[isl]) atomicAdd(&d.work[1], (unsigned long long)d.isl_nj[isl] * (iterations - rem));
    }
    // solved normal lambdas to the contacts: the next frame's warm start (Solver.elm:320-335)
    for (int i = rank * T + tid; i < gn; i += NC * T) {
      const int4 x = __ldcg((const int4*)(recs + i));
      for (int c = 0; c < x.w; c++) d.cur.contacts[x.z + c].lam = __ldcg(&d.rows[x.z + c].lamN);
    }
    if (tab) {         // this CTA's share of the solver-body deltas back to the global array
      for (int i = tid; (i << ncs) + rank <= nb; i += T) {
        if ((i << ncs) + rank == 0) continue;
        const double* z = s_body + 8 * (size_t)i; double* o = d.sbv + 6 * (size_t)s_brow[i];
        o[0] = z[0]; o[1] = z[1]; o[2] = z[2]; o[3] = z[4]; o[4] = z[5]; o[5] = z[6];
      }
    }
    __syncthreads();
    if (!SOLO) { cl_arrive(); cl_wait(); }       // nobody reads another CTA's shared memory or the island's scratch past this point
  }
}

}  // namespace ep
