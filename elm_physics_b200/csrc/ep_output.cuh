// Device-side marshalling of a frame's contacts into the layout of ep_contacts (include/elm_physics_b200.h): the valid pair
// slots of every world compacted in pair-enumeration order, contacts renumbered group by group in solver order.  The host
// then only issues one D2H copy per output array (no per-contact host loop).
#pragma once
#include "ep_kernels.cuh"

namespace ep {

// one warp per world: valid pair groups and their contacts
__global__ void __launch_bounds__(128) k_out_count(Dev d, FrameBuf f) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int w = warp; w < d.n_worlds; w += nwarps) {
    const int s0 = f.cand_base[w], sn = f.cand_n[w];
    int g = 0, c = 0;
    for (int s = s0 + lane; s < s0 + sn; s += 32) {
      const int cn = f.slot_ccount[s];
      const bool valid = cn > 0 || f.slot_jcount[s] > 0;
      g += valid ? 1 : 0; c += valid ? max(cn, 0) : 0;
    }
    for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane == 0) { d.out.wg_cnt[w] = g; d.out.wc_cnt[w] = c; }
  }
}

// single block: exclusive scans over the worlds
__global__ void __launch_bounds__(1024) k_out_scan(Dev d) {
  __shared__ int s_g[32], s_c[32];
  __shared__ int s_run_g, s_run_c;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { s_run_g = 0; s_run_c = 0; }
  __syncthreads();
  for (int base = 0; base < d.n_worlds; base += 1024) {
    const int w = base + tid;
    const int g = w < d.n_worlds ? d.out.wg_cnt[w] : 0, c = w < d.n_worlds ? d.out.wc_cnt[w] : 0;
    int sg = g, sc = c;
    for (int o = 1; o < 32; o <<= 1) {
      int vg = __shfl_up_sync(0xffffffffu, sg, o), vc = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) { sg += vg; sc += vc; }
    }
    if (lane == 31) { s_g[wid] = sg; s_c[wid] = sc; }
    __syncthreads();
    int bg = s_run_g, bc = s_run_c;
    for (int k = 0; k < wid; k++) { bg += s_g[k]; bc += s_c[k]; }
    if (w < d.n_worlds) { d.out.world_group_off[w] = bg + sg - g; d.out.wc_off[w] = bc + sc - c; }
    __syncthreads();
    if (tid == 1023) { s_run_g = bg + sg; s_run_c = bc + sc; }
    __syncthreads();
  }
  if (tid == 0) {
    d.out.world_group_off[d.n_worlds] = s_run_g;
    d.out.totals[0] = s_run_g; d.out.totals[1] = s_run_c;
    if (s_run_g <= d.out.group_cap) d.out.group_contact_off[s_run_g] = s_run_c;
  }
}

// one warp per world: group records in enumeration order + the slot -> group map
// early != 0: launched in the middle of the step (after the rows, beside the solver): the pre-step transform of body1 is the
// CURRENT one and the per-world solver results do not exist yet (the caller does not read them)
__global__ void __launch_bounds__(128) k_out_groups(Dev d, FrameBuf f, int early) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (d.out.totals[0] > d.out.group_cap) return;
  for (int w = warp; w < d.n_worlds; w += nwarps) {
    const int s0 = f.cand_base[w], sn = f.cand_n[w];
    int g = d.out.world_group_off[w], c = d.out.wc_off[w];
    for (int b = 0; b < sn; b += 32) {
      const int s = s0 + b + lane;
      const bool in = b + lane < sn;
      const int cn = in ? f.slot_ccount[s] : 0;
      const bool valid = in && (cn > 0 || f.slot_jcount[s] > 0);
      const int cv = valid ? max(cn, 0) : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      int scan = cv;
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, scan, o); if (lane >= o) scan += v; }
      const int tot = __shfl_sync(0xffffffffu, scan, 31);
      if (in) {
        const int gi = g + __popc(bal & ((1u << lane) - 1u));
        f.slot_group[s] = valid ? gi : -1;
        if (valid) {
          const int r1 = f.slot_row1[s], r2 = f.slot_row2[s], root = f.slot_root[s], nc = f.slot_ncon[s];
          d.out.group_id1[gi] = d.body_id[r1]; d.out.group_id2[gi] = d.body_id[r2];
          d.out.group_ncon[gi] = nc < 0 ? -nc : nc;
          d.out.group_island[gi] = root >= 0 ? d.body_id[root] : -1;
          d.out.group_contact_off[gi] = c + scan - cv;
#pragma unroll
          for (int k = 0; k < 3; k++) d.out.group_pre_pos1[3 * (size_t)gi + k] = (early ? d.pos : d.pre_pos)[3 * (size_t)r1 + k];
#pragma unroll
          for (int k = 0; k < 4; k++) d.out.group_pre_quat1[4 * (size_t)gi + k] = (early ? d.quat : d.pre_quat)[4 * (size_t)r1 + k];
        }
      }
      g += __popc(bal); c += tot;
    }
    if (lane == 0 && !early) {
      const int nbw = d.world_body_off[w + 1] - d.world_body_off[w];
      const int used = d.iterations[w] - d.world_min_rem[w];
      d.out.iterations[w] = nbw == 0 ? 0 : (1 - used > 0 ? 1 : used);      // Solver.elm:225-231, 275-276
      d.out.world_n_islands[w] = d.world_n_islands[w];
    }
  }
}

// one thread per contact of the frame
// early != 0: the solver is still running on the rows: lambda / t1 are not exported (the caller does not read them)
__global__ void __launch_bounds__(256) k_out_contacts(Dev d, FrameBuf f, int early) {
  if (d.out.totals[0] > d.out.group_cap || d.out.totals[1] > d.out.contact_cap) return;
  const int n = min(f.counters[CNT_CONTACTS], d.contact_cap);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const ContactRec& r = f.contacts[c];
    const int s = r.slot, g = f.slot_group[s];
    if (g < 0) continue;
    const size_t o = (size_t)d.out.group_contact_off[g] + (c - f.slot_cstart[s]);
    d.out.shape_key[o] = r.sk; d.out.feature_key[o] = r.fk;
#pragma unroll
    for (int k = 0; k < 3; k++) { d.out.ni[3 * o + k] = r.ni[k]; d.out.pi[3 * o + k] = r.pi[k]; d.out.pj[3 * o + k] = r.pj[k]; }
    d.out.friction[o] = r.friction; d.out.bounciness[o] = r.bounciness;
    if (early) continue;
    d.out.lambda[o] = r.lam;
    d.out.t1[3 * o] = r.t1[0]; d.out.t1[3 * o + 1] = r.t1[1]; d.out.t1[3 * o + 2] = r.t1[2];
  }
}

}  // namespace ep
