// Fused front end for batches of SMALL worlds (BASELINE configs 3 and 5): one CTA steps one world from body state
// to solver rows entirely in shared memory, so HBM sees only the algorithmic traffic (body state in, contacts +
// rows + island lists out) and nothing intermediate (placed hulls, sort keys, candidate pairs, raw contacts).
//   phase 0  load the world's bodies                                   (src/Internal/Body.elm:36-61)
//   phase 1  Shape.placeIn of every shape into shared memory           (src/Internal/Shape.elm:94-110)
//   phase 2  stable gravity sort (rank sort)                           (src/Physics.elm:548-573)
//   phase 3  i<j bounding-sphere test, ordered ballot compaction       (src/Internal/BroadPhase.elm:54-215)
//   phase 4  narrowphase shape matrix, hulls read from shared memory   (src/Internal/NarrowPhase.elm, src/Collision/*)
//   phase 5  pair groups compacted in enumeration order; rows + warm start (src/Internal/Equation.elm:116-450)
//   phase 6  union-find islands, island work list binned by size       (src/Internal/Islands.elm)
// The solver (ep_solve.cuh) and the integrator then run over all worlds' islands / bodies.
#pragma once
#include "ep_kernels.cuh"

namespace ep {

constexpr int kFusedThreads = 128;

// upper bound of contacts one shape pair can emit (sizes the per-candidate reservation in the raw pool)
__device__ __forceinline__ int max_shape_pair_contacts(int ka, bool boxa, int kb, bool boxb) {
  if (ka == SH_PLANE && kb == SH_PLANE) return 0;
  if (ka == SH_PARTICLE && kb == SH_PARTICLE) return 0;
  if (ka == SH_CONVEX && kb == SH_CONVEX) return 4;
  if (ka == SH_PLANE || kb == SH_PLANE) {
    int o = ka == SH_PLANE ? kb : ka; bool ob = ka == SH_PLANE ? boxb : boxa;
    return o == SH_CONVEX ? (ob ? 8 : 4) : (o == SH_CAPSULE ? 2 : 1);
  }
  if ((ka == SH_CAPSULE && kb == SH_CONVEX) || (ka == SH_CONVEX && kb == SH_CAPSULE)) return 3;
  return 1;
}

struct FusedSmem {
  // bodies (local index = position in the user list)
  double *pos, *quat, *vel, *angvel, *force, *torque, *iiw, *inv_mass, *radius, *key;
  double* wf;             // placed shape blocks
  RawContact* raw;        // raw contact pool
  int *kind, *sorted, *woff, *parent, *icount, *ifill;
  int *c_raw, *c_gidx, *c_coff;         // per candidate: raw start, compact group index, contact offset
  short *c_cnt, *c_ncon;                // per candidate: contact count (-1: no narrowphase), constraint count
  unsigned char *c_i, *c_j;             // per candidate: local body1 / body2
  unsigned short* raw_sk;               // per raw contact: shapeKey
  unsigned short* prev_key;             // per previous-frame group of this world: min local << 8 | max local
};

__host__ __device__ inline size_t fused_smem_bytes(int nb, int ninst, int wf64, int cand, int raw) {
  size_t b = 0;
  b += (size_t)(3 + 4 + 3 + 3 + 3 + 3 + 9 + 1 + 1 + 1) * nb * 8;   // body fields + key
  b += (size_t)wf64 * 8;
  b += (size_t)raw * sizeof(RawContact);
  b += (size_t)(5 * nb + ninst + 1) * 4;                           // kind, sorted, parent, icount, ifill, woff
  b += (size_t)cand * (3 * 4 + 2 * 2 + 2 + 2);                     // c_raw, c_gidx, c_coff, c_cnt, c_ncon, c_i, c_j, prev_key
  b += (size_t)raw * 2;
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ FusedSmem fused_carve(unsigned char* base, int nb, int ninst, int wf64, int cand, int raw) {
  FusedSmem s;
  double* f = (double*)base;
  s.pos = f; f += 3 * nb; s.quat = f; f += 4 * nb; s.vel = f; f += 3 * nb; s.angvel = f; f += 3 * nb;
  s.force = f; f += 3 * nb; s.torque = f; f += 3 * nb; s.iiw = f; f += 9 * nb; s.inv_mass = f; f += nb;
  s.radius = f; f += nb; s.key = f; f += nb;
  s.wf = f; f += wf64;
  s.raw = (RawContact*)f; f = (double*)(s.raw + raw);
  int* i = (int*)f;
  s.kind = i; i += nb; s.sorted = i; i += nb; s.parent = i; i += nb; s.icount = i; i += nb; s.ifill = i; i += nb;
  s.woff = i; i += ninst + 1;
  s.c_raw = i; i += cand; s.c_gidx = i; i += cand; s.c_coff = i; i += cand;
  short* h = (short*)i;
  s.c_cnt = h; h += cand; s.c_ncon = h; h += cand;
  s.prev_key = (unsigned short*)h; h += cand;
  s.raw_sk = (unsigned short*)h; h += raw;
  unsigned char* c = (unsigned char*)h;
  s.c_i = c; c += cand; s.c_j = c; c += cand;
  return s;
}

__device__ __forceinline__ BodyE load_bodye_s(const FusedSmem& s, int l) {
  BodyE b;
  b.k.kind = s.kind[l]; b.k.invMass = s.inv_mass[l];
  const double* m = s.iiw + 9 * l;
  b.k.I.m11 = m[0]; b.k.I.m21 = m[1]; b.k.I.m31 = m[2]; b.k.I.m12 = m[3]; b.k.I.m22 = m[4]; b.k.I.m32 = m[5]; b.k.I.m13 = m[6]; b.k.I.m23 = m[7]; b.k.I.m33 = m[8];
  b.x = ld3(s.pos + 3 * l); b.v = ld3(s.vel + 3 * l); b.w = ld3(s.angvel + 3 * l); b.f = ld3(s.force + 3 * l); b.tq = ld3(s.torque + 3 * l);
  b.q.x = s.quat[4 * l]; b.q.y = s.quat[4 * l + 1]; b.q.z = s.quat[4 * l + 2]; b.q.w = s.quat[4 * l + 3];
  return b;
}

__global__ void __launch_bounds__(kFusedThreads, 3) k_world_front(Dev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_w, s_cn, s_rawn, s_ng, s_nc, s_gbase, s_cbase, s_ok;
  __shared__ int s_warp[kFusedThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int nwarps = kFusedThreads / 32;
  const int NB = d.f_nb_max, CAND = d.f_cand_cap, RAW = d.f_raw_cap;
  FusedSmem S = fused_carve(smem_raw, NB, d.f_inst_max, d.f_wf64_max, CAND, RAW);
  TagPt pa[kMaxPoly], pb[kMaxPoly];
  long long tprev = 0;
#define EP_PHASE(i) do { if (d.phase_cycles && tid == 0) { long long tn_ = clock64(); atomicAdd(&d.phase_cycles[i], (unsigned long long)(tn_ - tprev)); tprev = tn_; } } while (0)

  for (;;) {
    __syncthreads();                                   // previous world's shared memory is dead
    if (tid == 0) { s_w = atomicAdd(&d.cur.counters[CNT_WORK], 1); s_rawn = 0; s_ok = 1; }
    __syncthreads();
    const int w = s_w;
    if (w >= d.n_worlds) break;
    if (d.phase_cycles && tid == 0) tprev = clock64();
    const int r0 = d.world_body_off[w], n = d.world_body_off[w + 1] - r0;
    const int i0 = d.inst_off[r0], ninst = d.inst_off[r0 + n] - i0;
    const double dt = d.dt[w]; const V3 gravity = ld3(d.gravity + 3 * w);

    // ---- phase 0: bodies -> shared memory (coalesced field copies) -------------------------------------------
    for (int k = tid; k < 3 * n; k += kFusedThreads) {
      S.pos[k] = d.pos[3 * (size_t)r0 + k]; S.vel[k] = d.vel[3 * (size_t)r0 + k]; S.angvel[k] = d.angvel[3 * (size_t)r0 + k];
      S.force[k] = d.force[3 * (size_t)r0 + k]; S.torque[k] = d.torque[3 * (size_t)r0 + k];
    }
    for (int k = tid; k < 4 * n; k += kFusedThreads) S.quat[k] = d.quat[4 * (size_t)r0 + k];
    for (int k = tid; k < 9 * n; k += kFusedThreads) S.iiw[k] = d.iiw[9 * (size_t)r0 + k];
    for (int k = tid; k < n; k += kFusedThreads) { S.inv_mass[k] = d.inv_mass[r0 + k]; S.radius[k] = d.radius[r0 + k]; S.kind[k] = d.kind[r0 + k]; }
    const int64_t wbase = ninst > 0 ? d.inst_woff[i0] : 0;
    for (int k = tid; k < ninst; k += kFusedThreads) S.woff[k] = (int)(d.inst_woff[i0 + k] - wbase);
    // previous frame's groups of this world, keyed by local body pair (bodyKey is symmetric, ContactId.elm:62-68)
    const int pg0 = d.prev.cand_base[w], pgn = min(d.prev.cand_n[w], CAND);
    for (int k = tid; k < pgn; k += kFusedThreads) {
      int a = d.prev.slot_row1[pg0 + k] - r0, b = d.prev.slot_row2[pg0 + k] - r0;
      S.prev_key[k] = d.prev.slot_ccount[pg0 + k] > 0 ? (unsigned short)(a < b ? (a << 8 | b) : (b << 8 | a)) : (unsigned short)0xFFFF;
    }
    __syncthreads();
    EP_PHASE(0);

    // ---- phase 1: copy the templates, then place every vector (one warp per shape instance) -------------------
    for (int k = wid; k < ninst; k += nwarps) {
      int sh = d.inst_shape[i0 + k];
      const double* src = d.sh_f64 + d.sh_f64_off[sh];
      int len = d.sh_f64_off[sh + 1] - d.sh_f64_off[sh];
      double* dst = S.wf + S.woff[k];
      for (int x = lane; x < len; x += 32) dst[x] = src[x];
      __syncwarp();
      int body = d.inst_body[i0 + k] - r0;
      Q4 q; q.x = S.quat[4 * body]; q.y = S.quat[4 * body + 1]; q.z = S.quat[4 * body + 2]; q.w = S.quat[4 * body + 3];
      V3 o = ld3(S.pos + 3 * body);
      for (int x = d.sh_pl_off[sh] + lane; x < d.sh_pl_off[sh + 1]; x += 32) {
        int code = d.sh_pl_code[x], rel = code >> 1;
        V3 v = ld3(src + rel);
        st3(dst + rel, (code & 1) ? rotate(q, v) : point_place(o, q, v));
      }
    }
    __syncthreads();
    EP_PHASE(1);
    // ---- phase 2: gravity sort (Physics.elm:548-573): rank sort == stable sort under the reference comparator --
    for (int i = tid; i < n; i += kFusedThreads) {
      V3 p = ld3(S.pos + 3 * i);
      S.key[i] = -(p.x * gravity.x + p.y * gravity.y + p.z * gravity.z);
    }
    __syncthreads();
    for (int i = tid; i < n; i += kFusedThreads) {
      double a = S.key[i];
      int rank = 0;
      for (int j = 0; j < n; j++) {
        double b = S.key[j];
        bool lt = b - a < 0, gt = b - a > 0;
        rank += (lt || (!gt && j < i)) ? 1 : 0;
      }
      S.sorted[rank] = i;
    }
    __syncthreads();

    EP_PHASE(2);
    // ---- phase 3: broadphase, candidates compacted in enumeration order ---------------------------------------
    {
      const int P = n * (n - 1) / 2;
      int running = 0;
      for (int p0 = 0; p0 < P; p0 += kFusedThreads) {
        int p = p0 + tid;
        bool flag = false, may = false; int l1 = 0, l2 = 0, ncon = 0;
        if (p < P) {
          int i, j; pair_from_index(p, n, i, j);
          l1 = S.sorted[i]; l2 = S.sorted[j];
          double rr = S.radius[l1] + S.radius[l2];
          double dx = S.pos[3 * l2] - S.pos[3 * l1], dy = S.pos[3 * l2 + 1] - S.pos[3 * l1 + 1], dz = S.pos[3 * l2 + 2] - S.pos[3 * l1 + 2];
          double dist2 = dx * dx + dy * dy + dz * dz;
          may = (rr * rr - dist2 > 0) && (S.kind[l1] == 2 || S.kind[l2] == 2);          // bodiesMayContact, BroadPhase.elm:189-215
          if (may && d.collide) may = d.collide[d.collide_off[w] + (int64_t)l1 * n + l2] != 0;
          ncon = constraints_between(d, r0 + l1, r0 + l2);
          flag = may || ncon != 0;
        }
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < nwarps; k++) { int c = s_warp[k]; if (k < wid) before += c; total += c; }
        if (flag) {
          int c = running + before + __popc(bal & ((1u << lane) - 1u));
          if (c < CAND) { S.c_i[c] = (unsigned char)l1; S.c_j[c] = (unsigned char)l2; S.c_cnt[c] = may ? 0 : -1; S.c_ncon[c] = (short)ncon; S.c_raw[c] = 0; }
        }
        running += total;
        __syncthreads();
      }
      if (tid == 0) {
        if (running > CAND) { d.flags[FLAG_PAIR_OVERFLOW] = 3; running = 0; s_ok = 0; }
        s_cn = running;
      }
    }
    __syncthreads();
    const int cn = s_cn;
    EP_PHASE(3);

    // ---- phase 4: narrowphase, one thread per candidate pair ----------------------------------------------------
    for (int c = tid; c < cn; c += kFusedThreads) {
      if (S.c_cnt[c] < 0) { S.c_cnt[c] = 0; continue; }
      int l1 = S.c_i[c], l2 = S.c_j[c];
      int a0 = d.inst_off[r0 + l1] - i0, a1 = d.inst_off[r0 + l1 + 1] - i0, b0 = d.inst_off[r0 + l2] - i0, b1 = d.inst_off[r0 + l2 + 1] - i0;
      int bound = 0;
      for (int a = a0; a < a1; a++) {
        int sa = d.inst_shape[i0 + a]; int ka = d.sh_kind[sa]; bool boxa = ka == SH_CONVEX && d.sh_i32[d.sh_i32_off[sa]] != 0;
        for (int b = b0; b < b1; b++) {
          int sb = d.inst_shape[i0 + b]; int kb = d.sh_kind[sb]; bool boxb = kb == SH_CONVEX && d.sh_i32[d.sh_i32_off[sb]] != 0;
          bound += max_shape_pair_contacts(ka, boxa, kb, boxb);
        }
      }
      if (bound == 0) continue;
      int start = atomicAdd(&s_rawn, bound);
      if (start + bound > RAW) { d.flags[FLAG_PAIR_OVERFLOW] = 4; continue; }
      ContactSink sink; sink.buf = S.raw + start; sink.n = 0; sink.cap = bound; sink.overflow = false;
      bool polyOvf = false;
      for (int a = a0; a < a1; a++) {
        ShapeRef sa; int sha = d.inst_shape[i0 + a];
        sa.kind = d.sh_kind[sha]; sa.f = S.wf + S.woff[a]; sa.ib = d.sh_i32 + d.sh_i32_off[sha];
        for (int b = b0; b < b1; b++) {
          ShapeRef sb; int shb = d.inst_shape[i0 + b];
          sb.kind = d.sh_kind[shb]; sb.f = S.wf + S.woff[b]; sb.ib = d.sh_i32 + d.sh_i32_off[shb];
          int n0 = sink.n;
          shape_pair_contacts(sa, sb, sink, pa, pb, polyOvf);
          // solver order = reverse emission inside a shape pair (Equation.elm:135-154)
          for (int x = n0, y = sink.n - 1; x < y; x++, y--) { RawContact t = sink.buf[x]; sink.buf[x] = sink.buf[y]; sink.buf[y] = t; }
          unsigned short sk = (unsigned short)((a - a0 + 1) * 256 + (b - b0 + 1));
          for (int x = n0; x < sink.n; x++) S.raw_sk[start + x] = sk;
        }
      }
      if (sink.overflow) d.flags[FLAG_PAIR_OVERFLOW] = 2;
      if (polyOvf) d.flags[FLAG_POLY_OVERFLOW] = 1;
      S.c_raw[c] = start; S.c_cnt[c] = (short)sink.n;
    }
    __syncthreads();

    EP_PHASE(4);
    // ---- phase 5a: compact the pair groups (contacts or constraints) in enumeration order -----------------------
    {
      const int per = (cn + kFusedThreads - 1) / kFusedThreads;
      const int cb = min(tid * per, cn), ce = min(cb + per, cn);
      int g = 0, k = 0;
      for (int c = cb; c < ce; c++) { int cnt = S.c_cnt[c]; if (cnt > 0 || S.c_ncon[c] != 0) { g++; k += cnt; } }
      int packed = (g << 16) | k;                              // <= 65535 contacts / groups per world
      int incl = packed;
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) s_warp[wid] = incl;
      __syncthreads();
      int before = 0, total = 0;
      for (int q = 0; q < nwarps; q++) { int v = s_warp[q]; if (q < wid) before += v; total += v; }
      int excl = before + incl - packed;
      g = excl >> 16; k = excl & 0xFFFF;
      for (int c = cb; c < ce; c++) {
        int cnt = S.c_cnt[c];
        if (cnt > 0 || S.c_ncon[c] != 0) { S.c_gidx[c] = g; S.c_coff[c] = k; g++; k += cnt; } else S.c_gidx[c] = -1;
      }
      if (tid == 0) {
        int ng = total >> 16, nc = total & 0xFFFF;
        int gbase = atomicAdd(&d.cur.counters[CNT_GROUPS], ng);
        int cbase = atomicAdd(&d.cur.counters[CNT_CONTACTS], nc);
        atomicAdd(&d.cur.counters[CNT_CAND], cn);
        if ((int64_t)gbase + ng > d.slot_cap) { d.flags[FLAG_PAIR_OVERFLOW] = 1; ng = 0; nc = 0; gbase = 0; s_ok = 0; }
        if ((int64_t)cbase + nc > d.contact_cap) { d.flags[FLAG_POOL_OVERFLOW] = 1; ng = 0; nc = 0; cbase = 0; s_ok = 0; }
        s_ng = ng; s_nc = nc; s_gbase = gbase; s_cbase = cbase;
        d.cur.cand_base[w] = gbase; d.cur.cand_n[w] = ng;
        d.world_min_rem[w] = d.iterations[w];
      }
    }
    __syncthreads();
    const int ng = s_ng, gbase = s_gbase, cbase = s_cbase;
    const bool ok = s_ok != 0;
    EP_PHASE(5);

    // ---- phase 6 (thread 0) || phase 5b (warps 1..): islands while the other warps build the rows ---------------
    if (tid == 0) {
      int nisl = 0;
      if (ok && ng > 0) {
        for (int l = 0; l < n; l++) { S.parent[l] = l; S.icount[l] = 0; S.ifill[l] = 0; }
        for (int c = cn - 1; c >= 0; c--) {           // Islands.connect inside buildAndWarmStart (Solver.elm:201-206)
          if (S.c_gidx[c] < 0) continue;
          int la = S.c_i[c], lb = S.c_j[c];
          if (S.kind[la] == 2 && S.kind[lb] == 2) {
            int a = uf_find(S.parent, la), b = uf_find(S.parent, lb);
            int ia = d.body_id[r0 + a], ib = d.body_id[r0 + b];       // lower id becomes root (Islands.elm:39-46)
            if (ia - ib < 0) S.parent[b] = a; else if (ia - ib > 0) S.parent[a] = b;
          }
        }
        for (int c = 0; c < cn; c++) {                // annotate (Islands.elm:68-89)
          int g = S.c_gidx[c];
          if (g < 0) continue;
          int la = S.c_i[c], lb = S.c_j[c];
          int root = uf_find(S.parent, S.kind[la] == 2 ? la : lb);
          d.cur.slot_root[gbase + g] = r0 + root;
          S.icount[root]++;
          S.ifill[root] += 3 * S.c_cnt[c] + (S.c_ncon[c] != 0 ? 6 : 0);     // size class only: joint rows counted as 6
        }
        for (int l = 0; l < n; l++) nisl += S.icount[l] > 0;
        int ibase = atomicAdd(&d.cur.counters[CNT_ISLANDS], nisl);
        int k = 0, goff = gbase;
        for (int l = 0; l < n; l++) if (S.icount[l] > 0) {
          int isl = ibase + k;
          d.isl_off[isl] = goff; d.isl_cnt[isl] = S.icount[l]; d.isl_world[isl] = w;
          int bin = island_bin(S.ifill[l]);
          int at = atomicAdd(&d.cur.counters[CNT_BIN0 + bin], 1);
          d.bin_items[(size_t)bin * d.n_bodies + at] = isl;
          S.ifill[l] = goff; goff += S.icount[l]; k++;
        }
        for (int c = 0; c < cn; c++) {                // groups of an island stay in enumeration order
          int g = S.c_gidx[c];
          if (g < 0) continue;
          int la = S.c_i[c], lb = S.c_j[c];
          int root = uf_find(S.parent, S.kind[la] == 2 ? la : lb);
          d.isl_groups[S.ifill[root]++] = gbase + g;
        }
      }
      d.world_n_islands[w] = nisl;
      EP_PHASE(6);
    } else if (ok && tid >= 32) {
      const Spook sp = spook(dt);
      for (int c = tid - 32; c < cn; c += kFusedThreads - 32) {
        int g = S.c_gidx[c];
        if (g < 0) continue;
        int l1 = S.c_i[c], l2 = S.c_j[c], cnt = S.c_cnt[c], ncon = S.c_ncon[c];
        int slot = gbase + g, c0 = cbase + S.c_coff[c];
        d.cur.slot_row1[slot] = r0 + l1; d.cur.slot_row2[slot] = r0 + l2;
        d.cur.slot_cstart[slot] = c0; d.cur.slot_ccount[slot] = cnt;
        d.cur.slot_ncon[slot] = ncon; d.cur.slot_jstart[slot] = 0; d.cur.slot_jcount[slot] = 0;
        BodyE b1 = load_bodye_s(S, l1), b2 = load_bodye_s(S, l2);
        if (cnt > 0) {
          // Cache.getGroup (bodyKey id1 id2), Equation.elm:122-128: previous-frame group of the same body pair
          unsigned short key = (unsigned short)(l1 < l2 ? (l1 << 8 | l2) : (l2 << 8 | l1));
          int pc0 = 0, pcn = 0;
          for (int m = 0; m < pgn; m++) if (S.prev_key[m] == key) { pc0 = d.prev.slot_cstart[pg0 + m]; pcn = d.prev.slot_ccount[pg0 + m]; break; }
          int a0 = d.inst_off[r0 + l1], b0 = d.inst_off[r0 + l2];
          int rs = S.c_raw[c];
          for (int k = 0; k < cnt; k++) {
            const RawContact& rc = S.raw[rs + k];
            int sk = S.raw_sk[rs + k];
            int ia = a0 + (sk >> 8) - 1, ib = b0 + (sk & 255) - 1;
            double f1 = d.inst_friction[ia], f2 = d.inst_friction[ib], e1 = d.inst_bounciness[ia], e2 = d.inst_bounciness[ib];
            ContactRec cr;
            cr.fk = rc.fk; cr.sk = sk; cr.slot = slot;
            cr.ni[0] = rc.ni.x; cr.ni[1] = rc.ni.y; cr.ni[2] = rc.ni.z;
            cr.pi[0] = rc.pi.x; cr.pi[1] = rc.pi.y; cr.pi[2] = rc.pi.z;
            cr.pj[0] = rc.pj.x; cr.pj[1] = rc.pj.y; cr.pj[2] = rc.pj.z;
            cr.friction = sqrt(f1 * f2);                          // Material.elm:23-25
            cr.bounciness = (e1 + e2 + eabs(e1 - e2)) * 0.5;      // Material.elm:31-33
            d.cur.contacts[c0 + k] = cr;
            // Cache.lookup: first match in pairGroup.contacts order == LAST match in solver order (ContactCache.elm:76-87)
            double lam = 0; V3 ct1 = mk(0, 0, 0);
            for (int m = pcn - 1; m >= 0; m--) {
              const ContactRec& p = d.prev.contacts[pc0 + m];
              if ((p.fk - rc.fk == 0) && (p.sk - sk == 0)) { const RowRec& pr = d.prev.rows[pc0 + m]; lam = pr.lamN; ct1 = mk(pr.f1J[3], pr.f1J[4], pr.f1J[5]); break; }
            }
            RowRec r;
            contact_row(r, rc.ni, rc.pi, rc.pj, cr.bounciness, cr.friction, lam, ct1, dt, gravity, sp, b1, b2);
            d.cur.rows[c0 + k] = r;
          }
        }
        if (ncon != 0) {
          int j0, E;
          joint_rows_for_pair(d, r0 + l1, r0 + l2, ncon, b1, b2, dt, gravity, sp, j0, E);
          d.cur.slot_jstart[slot] = j0; d.cur.slot_jcount[slot] = E;
        }
      }
    }
    if (d.phase_cycles) { __syncthreads(); EP_PHASE(7); }
  }
#undef EP_PHASE
}

}  // namespace ep
