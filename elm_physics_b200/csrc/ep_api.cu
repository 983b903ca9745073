// C ABI of the stepping engine (include/elm_physics_b200.h): context, uploads, the per-step launch
// sequence, downloads.  Host code here is marshalling + launch plumbing only; all physics is in
// ep_kernels.cu.  There is no CPU fallback: without a usable CUDA device ep_create fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <mutex>
#include <chrono>
#include <vector>

#include "../../include/elm_physics_b200.h"
#include "ep_device.cuh"

#include "ep_kernels.cuh"
#include "ep_solve.cuh"
#include "ep_solve_quads.cuh"
#include "ep_solve_cluster.cuh"
#include "ep_output.cuh"
#include "ep_raycast.cuh"

using namespace ep;

struct ep_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<void*> shape_allocs, world_allocs;
  Dev d{};
  bool have_shapes = false, have_worlds = false;
  // host mirrors needed for marshalling
  std::vector<int32_t> h_sh_kind, h_sh_f64_off, h_sh_i32_off, h_sh_i32;
  std::vector<double> h_sh_f64;
  std::vector<int32_t> h_pl_off, h_pl_code;
  std::vector<int32_t> h_world_body_off, h_body_id, h_iterations;
  int sm_count = 148;
  int64_t launches = 0;
  int64_t opt[EP_OPT_COUNT] = {0};
  // solver: tiled size classes run on side streams next to the big-island kernel
  static constexpr int kTileBins = 4;
  static constexpr int kAux = 12;                  // side streams of the solver's concurrent launches
  cudaStream_t aux[kAux] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
  // ep_step_frame's copy streams: the bulk of the H2D runs beside the front kernels, the contact read-back beside the solver
  cudaStream_t cin = nullptr, cout = nullptr;
  cudaEvent_t ev_in = nullptr, ev_mid = nullptr, ev_out_done = nullptr;
  bool hook_wait_in = false, hook_early_out = false;     // set by ep_step_frame around its one step
  cudaEvent_t ev_block = nullptr;                        // blocking-sync event of wait_stream
  int tile_grid[kTileBins] = {0, 0, 0, 0};
  int tile_smem[6] = {0, 0, 0, 0, 0, 0};        // dynamic shared memory of the launch that serves each lane class
  int head_grid[6] = {0, 0, 0, 0, 0, 0};         // resident CTAs of the staged launch of each class
  int lanes_grid = 0;
  size_t team_smem_max = 48 * 1024;
  int cluster_nc = 0, cluster_ws = 64, cluster_count = 1, solo_bodies = 1024, cluster_bodies = 512;      // k_solve_cluster: CTAs per cluster (0 = unavailable), window slots per warp, clusters per launch, body table of the SOLO launch
  // quad solver (small batches): launches of k_solve_quads -- classes [lo, hi] with kq islands per warp, shared memory sized
  // for class hi (ep_solve_quads.cuh) -- and the largest batch (bodies) that takes this family (EP_SOLVER=lanes|quads overrides)
  struct QuadLaunch { int lo, hi, kq, cap_rows, cap_bodies, cap_groups; size_t smem; int resident; };
  std::vector<QuadLaunch> quad_plan;
  int64_t quads_max_bodies = 80000;             // measured crossover on B200: C3 at 2048 worlds (67 k bodies) still wins with quads, 4096 loses
  int quads_max_world = 64;                     // worlds that can hold an island of hundreds of contacts stay with the lane / team family            // device opt-in maximum of dynamic shared memory per CTA (set in ep_create)
  // per-kernel CUDA-event timing (ep_profile): pending (kernel, start, end) triples resolved at sync
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Pending { int kernel; cudaEvent_t a, b; };
  std::vector<Pending> ev_pending, np_pending;
  double np_ms[kClasses] = {0};
  double k_ms[EP_K_COUNT] = {0};
  int64_t k_launches[EP_K_COUNT] = {0};
  int64_t stats[EP_STAT_COUNT] = {0};
  // ep_raycast scratch (device), grown on demand
  void* ray_buf = nullptr; size_t ray_cap = 0;
  bool placed_current = false;      // the world-placed shape copies match the bodies' current transforms (ep_raycast)
  // EP_TRACE_FRAME=1: host-side phases of ep_step_frame_fields accumulated per context, printed by ep_destroy (development aid)
  double tr_us[6] = {0}; int64_t tr_frames = 0;
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                     \
      return EP_ECUDA;                                                                                   \
    }                                                                                                    \
  } while (0)

template <typename T>
static int dev_alloc(ep_ctx* ctx, std::vector<void*>& pool, T** out, size_t n) {
  void* p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  pool.push_back(p);
  *out = (T*)p;
  return EP_OK;
}
template <typename T>
static int dev_upload(ep_ctx* ctx, std::vector<void*>& pool, const T** out, const T* src, size_t n) {
  T* p = nullptr;
  int rc = dev_alloc(ctx, pool, &p, n);
  if (rc) return rc;
  if (n) CK(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *out = p;
  return EP_OK;
}
template <typename T>
static int dev_upload_rw(ep_ctx* ctx, std::vector<void*>& pool, T** out, const T* src, size_t n) {
  const T* p = nullptr;
  int rc = dev_upload(ctx, pool, &p, src, n);
  *out = (T*)p;
  return rc;
}
static void free_pool(std::vector<void*>& pool) {
  for (void* p : pool) cudaFree(p);
  pool.clear();
}
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

extern "C" int ep_create(ep_ctx** out, int device) {
  if (!out) return EP_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return EP_ENODEVICE;
  ep_ctx* ctx = new ep_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return EP_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  bool ok = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->cin, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->cout, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ctx->ev_mid, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->ev_out_done, cudaEventDisableTiming) == cudaSuccess;
  // side streams in launch order of the solver (longest class first): earlier ones get the higher priority so that the
  // long tiles are scheduled onto SMs before the short ones fill them
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < ep_ctx::kAux && ok; i++)
    ok = cudaStreamCreateWithPriority(&ctx->aux[i], cudaStreamNonBlocking, i < 4 ? prio_hi : prio_lo) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  // Shared memory of the lane-solver launches, sized for the typical tile of each class: classes 0..2 stage the whole image
  // (stand-in + 2 dynamic bodies, 2^class rows); classes 3..5 stage bodies + descriptors only (2^(class-1)+1 dynamic bodies,
  // 2^class rows) and stream the rows.  Tiles that need more fall to the next residency mode (k_solve_lanes).
  for (int i = 0; i < 3; i++) { TileDims t; t.maxB = 3; t.maxR = 1 << i; t.maxC = 1 << i; ctx->tile_smem[i] = (int)(tile_words(t, 32) * 8); }
  for (int i = 3; i < 6; i++) { TileDims t; t.maxB = (1 << (i - 1)) + 2; t.maxR = 1 << i; t.maxC = 1 << i; ctx->tile_smem[i] = (int)(tile_head_words(t, 32) * 8); }
  ok = ok && cudaFuncSetAttribute(k_solve_lanes<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->tile_smem[2]) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_solve_lanes<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->tile_smem[5]) == cudaSuccess;
  for (int i = 0; i < 6 && ok; i++) {
    int per_sm = 0;
    if (i < 3) ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_lanes<2>, 32, ctx->tile_smem[i]) == cudaSuccess && per_sm >= 1;
    else ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_lanes<1>, 32, ctx->tile_smem[i]) == cudaSuccess && per_sm >= 1;
    ctx->head_grid[i] = per_sm * ctx->sm_count;
  }
  if (const char* hg = getenv("EP_HEAD_GRID")) {      // tuning aid: resident CTAs per SM of each lane class, "c0,c1,c2,c3,c4,c5" (0 keeps the occupancy maximum)
    int v[6] = {0, 0, 0, 0, 0, 0};
    sscanf(hg, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
    for (int i = 0; i < 6; i++) if (v[i] > 0) ctx->head_grid[i] = std::min(ctx->head_grid[i], v[i] * ctx->sm_count);
  }
  if (ok) {
    int per_sm = 0;
    ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_lanes<0>, 128, 0) == cudaSuccess && per_sm >= 1;
    ctx->lanes_grid = per_sm * ctx->sm_count;
  }
  // The team kernels' dynamic shared memory depends on the largest world of a batch.  The attribute is per function and per
  // DEVICE, i.e. shared by every context of the process: it is set once, here, to the device's opt-in maximum, and a batch
  // that needs more falls back to the variant that keeps the bodies in global memory (ep_step_resident).
  int smem_optin = 0;
  ok = ok && cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess;
  smem_optin -= 1024;                           // room for the kernels' static shared memory
  ctx->team_smem_max = (size_t)std::max(smem_optin, 0);
  ok = ok && cudaFuncSetAttribute(k_solve_team<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_solve_team<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) == cudaSuccess;
  // Quad solver plan: which classes share a launch and how many islands a warp takes (EP_QUAD_PLAN="lo-hi:kq,..." overrides).
  // Shared memory of a launch is sized for its largest class: 3 * 2^c row slots, 2^c pair groups, 2^c + 2 bodies per island.
  {
    if (getenv("EP_QUADS_MAX_BODIES")) ctx->quads_max_bodies = atoll(getenv("EP_QUADS_MAX_BODIES"));
    const char* plan = getenv("EP_QUAD_PLAN");
    std::string ps = plan ? plan : "0-1:8,2-2:8,3-3:8,4-4:4,5-5:2";
    size_t pos = 0;
    while (pos < ps.size() && ok) {
      int lo = 0, hi = 0, kq = 0, used = 0;
      if (sscanf(ps.c_str() + pos, "%d-%d:%d%n", &lo, &hi, &kq, &used) != 3 || lo < 0 || hi < lo || hi > kLaneBinMax || (kq != 8 && kq != 4 && kq != 2)) { ok = false; break; }
      ep_ctx::QuadLaunch q{lo, hi, kq, 3 << hi, (1 << hi) + 2, 1 << hi, 0, 0};
      const void* fn = kq == 8 ? (const void*)k_solve_quads<8> : (kq == 4 ? (const void*)k_solve_quads<4> : (const void*)k_solve_quads<2>);
      q.smem = kq == 8 ? QuadTile<8>::bytes(q.cap_rows, q.cap_bodies, q.cap_groups) : (kq == 4 ? QuadTile<4>::bytes(q.cap_rows, q.cap_bodies, q.cap_groups) : QuadTile<2>::bytes(q.cap_rows, q.cap_bodies, q.cap_groups));
      int per_sm = 0;
      ok = ok && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
      ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32, q.smem) == cudaSuccess && per_sm >= 1;
      q.resident = per_sm * ctx->sm_count;
      ctx->quad_plan.push_back(q);
      pos += used;
      if (pos < ps.size() && ps[pos] == ',') pos++;
    }
    int covered = 0;
    for (auto& q : ctx->quad_plan) for (int b = q.lo; b <= q.hi; b++) covered |= 1 << b;
    ok = ok && covered == (1 << (kLaneBinMax + 1)) - 1 && (int)ctx->quad_plan.size() + 5 <= ep_ctx::kAux;
    const int optin = (int)ctx->team_smem_max;
    ok = ok && cudaFuncSetAttribute(k_solve_qteam<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_solve_qteam<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_solve_qteam<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_solve_qteam<16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess;
  }
  // Giant islands (class 9) run on k_solve_cluster (ep_solve_cluster.cuh): SOLO (one CTA, bodies in shared memory) for islands of
  // <= 1024 pair groups, else a thread-block cluster -- the largest this device co-schedules with the kernel's shared memory,
  // 16 (non-portable size) or 8 CTAs.  EP_CLUSTER=0 keeps them on k_solve_team<256>; EP_CLUSTER=n forces n; EP_SOLO=0: no SOLO launch.
  if (ok) {
    const char* ce = getenv("EP_CLUSTER");
    const int want = ce ? atoi(ce) : 16;
    if (getenv("EP_CLUSTER_WS")) ctx->cluster_ws = std::max(0, std::min(64, atoi(getenv("EP_CLUSTER_WS"))));
    if (getenv("EP_SOLO") && atoi(getenv("EP_SOLO")) == 0) ctx->solo_bodies = 0;
    if (getenv("EP_CLUSTER_BODIES")) ctx->cluster_bodies = std::max(1, std::min(512, atoi(getenv("EP_CLUSTER_BODIES"))));      // test aid: a small table forces the global-array fallback
    const size_t sm = cluster_smem_bytes(kClusterWarps, ctx->cluster_ws, ctx->cluster_bodies), sm_solo = cluster_smem_bytes(kClusterWarps, ctx->cluster_ws, ctx->solo_bodies);
    if (want > 0 && sm_solo <= ctx->team_smem_max &&
        cudaFuncSetAttribute(k_solve_cluster<kClusterWarps, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) == cudaSuccess &&
        cudaFuncSetAttribute(k_solve_cluster<kClusterWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_solo) == cudaSuccess) {
      cudaFuncSetAttribute(k_solve_cluster<kClusterWarps, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      for (int nc = std::min(want, 16); nc >= 1 && ctx->cluster_nc == 0; nc >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nc); cfg.blockDim = dim3(kClusterThreads); cfg.dynamicSmemBytes = sm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k_solve_cluster<kClusterWarps, false>, &cfg) == cudaSuccess && ncl >= 1) { ctx->cluster_nc = nc; ctx->cluster_count = std::min(ncl, 4); }
      }
      cudaGetLastError();       // a size the device refuses is not an error of the context
    }
  }
  if (!ok) { ep_destroy(ctx); return EP_ENODEVICE; }
  *out = ctx;
  return EP_OK;
}

extern "C" void ep_destroy(ep_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->tr_frames > 0)
    fprintf(stderr, "[ep] frame trace, %lld frames, mean us: enqueue H2D %.0f | enqueue step %.0f | enqueue D2H %.0f | (unused %.0f) | totals wait + contact copies + wait %.0f | flags + wait main %.0f\n",
            (long long)ctx->tr_frames, ctx->tr_us[0] / ctx->tr_frames, ctx->tr_us[1] / ctx->tr_frames, ctx->tr_us[2] / ctx->tr_frames,
            ctx->tr_us[3] / ctx->tr_frames, ctx->tr_us[4] / ctx->tr_frames, ctx->tr_us[5] / ctx->tr_frames);
  free_pool(ctx->world_allocs);
  free_pool(ctx->shape_allocs);
  if (ctx->ray_buf) cudaFree(ctx->ray_buf);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < ep_ctx::kAux; i++) { if (ctx->aux[i]) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); } if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->cin) { cudaStreamSynchronize(ctx->cin); cudaStreamDestroy(ctx->cin); }
  if (ctx->cout) { cudaStreamSynchronize(ctx->cout); cudaStreamDestroy(ctx->cout); }
  if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
  if (ctx->ev_out_done) cudaEventDestroy(ctx->ev_out_done);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int ep_configure(ep_ctx* ctx, int32_t option, int64_t value) {
  if (!ctx || option < 0 || option >= EP_OPT_COUNT) return EP_EINVAL;
  ctx->opt[option] = value;
  return EP_OK;
}
extern "C" const char* ep_last_error(const ep_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* ep_stream(ep_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int ep_upload_shapes(ep_ctx* ctx, const ep_shapes* S) {
  if (!ctx || !S || S->n_shapes < 0) return EP_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  free_pool(ctx->shape_allocs);
  int ns = S->n_shapes;
  ctx->h_sh_kind.assign(S->kind, S->kind + ns);
  ctx->h_sh_f64_off.assign(S->f64_off, S->f64_off + ns + 1);
  ctx->h_sh_i32_off.assign(S->i32_off, S->i32_off + ns + 1);
  ctx->h_sh_f64.assign(S->f64, S->f64 + S->f64_off[ns]);
  ctx->h_sh_i32.assign(S->i32, S->i32 + S->i32_off[ns]);
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_kind, ctx->h_sh_kind.data(), ns));
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_f64_off, ctx->h_sh_f64_off.data(), ns + 1));
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_i32_off, ctx->h_sh_i32_off.data(), ns + 1));
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_f64, ctx->h_sh_f64.data(), ctx->h_sh_f64.size()));
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_i32, ctx->h_sh_i32.data(), ctx->h_sh_i32.size()));
  // placement program of each template: which f64 triples are points (0) / directions (1) to be placed every step
  std::vector<int32_t> pl_off(ns + 1, 0), pl_code;
  for (int s = 0; s < ns; s++) {
    const int32_t* ib = ctx->h_sh_i32.data() + ctx->h_sh_i32_off[s];
    auto item = [&](int rel, int dir) { pl_code.push_back(rel * 2 + dir); };
    switch (ctx->h_sh_kind[s]) {
      case EP_SPHERE: case EP_PARTICLE: item(0, 0); break;
      case EP_CAPSULE: case EP_PLANE: item(0, 0); item(3, 1); break;
      case EP_CONVEX: {
        int nv = ib[1], ng = ib[2];
        item(0, 0);
        if (ib[0]) { item(3, 1); item(6, 1); item(9, 1); }
        for (int v = 0; v < nv; v++) item(EP_CONVEX_F64_HEADER + 3 * v, 0);
        for (int g = 0; g < ng; g++) {
          int base = EP_CONVEX_F64_HEADER + 3 * nv + EP_GROUP_F64 * g;
          item(base, 1);
          if (ib[EP_CONVEX_I32_HEADER + EP_GROUP_I32 * g]) item(base + 4, 1);
        }
        break;
      }
      default: ctx->err = "unknown shape kind"; return EP_EINVAL;
    }
    pl_off[s + 1] = (int32_t)pl_code.size();
  }
  ctx->h_pl_off = pl_off; ctx->h_pl_code = pl_code;
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_pl_off, pl_off.data(), pl_off.size()));
  TRY(dev_upload(ctx, ctx->shape_allocs, &ctx->d.sh_pl_code, pl_code.data(), pl_code.size()));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_shapes = true;
  return EP_OK;
}

static int alloc_frame(ep_ctx* ctx, FrameBuf& f, const Dev& d) {
  auto& pool = ctx->world_allocs;
  TRY(dev_alloc(ctx, pool, &f.contacts, d.contact_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_row1, d.slot_cap)); TRY(dev_alloc(ctx, pool, &f.slot_row2, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_cstart, d.slot_cap)); TRY(dev_alloc(ctx, pool, &f.slot_ccount, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_jstart, d.slot_cap)); TRY(dev_alloc(ctx, pool, &f.slot_jcount, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_ncon, d.slot_cap)); TRY(dev_alloc(ctx, pool, &f.slot_root, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_group, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.slot_item0, d.slot_cap)); TRY(dev_alloc(ctx, pool, &f.slot_nitems, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.pair_slot, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &f.cand_base, d.n_worlds)); TRY(dev_alloc(ctx, pool, &f.cand_n, d.n_worlds));
  TRY(dev_alloc(ctx, pool, &f.counters, CNT_COUNT));
  CK(cudaMemsetAsync(f.pair_slot, 0xFF, (size_t)d.slot_cap * sizeof(int32_t), ctx->stream));
  CK(cudaMemsetAsync(f.counters, 0, CNT_COUNT * sizeof(int32_t), ctx->stream));
  CK(cudaMemsetAsync(f.cand_base, 0, std::max(d.n_worlds, 1) * sizeof(int32_t), ctx->stream));
  CK(cudaMemsetAsync(f.cand_n, 0, std::max(d.n_worlds, 1) * sizeof(int32_t), ctx->stream));
  return EP_OK;
}

static int grid_for(const ep_ctx* ctx, int64_t n, int block, int per_sm) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

extern "C" int ep_upload_worlds(ep_ctx* ctx, const ep_bodies* B, const ep_params* P, const ep_contacts* warm) {
  if (!ctx || !B || !P) return EP_EINVAL;
  if (!ctx->have_shapes) { ctx->err = "ep_upload_shapes must be called first"; return EP_ESTATE; }
  if (B->n_worlds != P->n_worlds) { ctx->err = "bodies/params world count mismatch"; return EP_EINVAL; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  free_pool(ctx->world_allocs);
  ctx->have_worlds = false;
  auto& pool = ctx->world_allocs;
  Dev& d = ctx->d;
  const int nw = B->n_worlds, nb = B->n_bodies, ni = B->n_insts;
  d.n_worlds = nw; d.n_bodies = nb; d.n_insts = ni;
  for (int w = 0; w < nw; w++)
    if (P->iterations[w] < 1) { ctx->err = "solverIterations must be >= 1"; return EP_EINVAL; }
  ctx->h_world_body_off.assign(B->world_body_off, B->world_body_off + nw + 1);
  ctx->h_body_id.assign(B->body_id, B->body_id + nb);
  ctx->h_iterations.assign(P->iterations, P->iterations + nw);
  // bodies
  TRY(dev_upload(ctx, pool, &d.world_body_off, B->world_body_off, nw + 1));
  TRY(dev_upload(ctx, pool, &d.body_id, B->body_id, nb));
  TRY(dev_upload(ctx, pool, &d.kind, B->kind, nb));
  std::vector<int32_t> body_world(nb), inst_body(ni), pair_off(std::max(nw, 1), 0);
  int64_t slot_cap = 0;
  for (int w = 0; w < nw; w++) {
    int64_t n = B->world_body_off[w + 1] - B->world_body_off[w];
    pair_off[w] = (int32_t)std::min<int64_t>(slot_cap, (1ll << 31) - 1);
    slot_cap += n * (n - 1) / 2;
    for (int r = B->world_body_off[w]; r < B->world_body_off[w + 1]; r++) body_world[r] = w;
  }
  if (slot_cap >= (int64_t)1 << 31) { ctx->err = "too many body pairs for 32-bit pair slots"; return EP_ECAPACITY; }
  {   // the dense warm-start table stands for bodyKey(id1, id2) (ContactId.elm:62-68) only while ids are unique in a world
    std::vector<int32_t> ids;
    for (int w = 0; w < nw; w++) {
      ids.assign(B->body_id + B->world_body_off[w], B->body_id + B->world_body_off[w + 1]);
      std::sort(ids.begin(), ids.end());
      if (std::adjacent_find(ids.begin(), ids.end()) != ids.end()) { ctx->err = "duplicate body id inside a world (run AssignIds.assignIds first)"; return EP_EINVAL; }
    }
  }
  for (int r = 0; r < nb; r++) {
    if (B->body_id[r] < 0 || B->body_id[r] >= 262144) { ctx->err = "body id out of range [0, 2^18)"; return EP_EINVAL; }
    for (int k = B->inst_off[r]; k < B->inst_off[r + 1]; k++) inst_body[k] = r;
    if (B->inst_off[r + 1] - B->inst_off[r] > 255) { ctx->err = "more than 255 shapes on a body"; return EP_EINVAL; }
  }
  TRY(dev_upload(ctx, pool, &d.body_world, body_world.data(), nb));
  TRY(dev_upload(ctx, pool, &d.world_pair_off, pair_off.data(), nw));
  {   // broadphase chunk table
    std::vector<int32_t> cw, cp, wco(nw + 1, 0);
    for (int w = 0; w < nw; w++) {
      int64_t n = B->world_body_off[w + 1] - B->world_body_off[w], P = n * (n - 1) / 2;
      for (int64_t p0 = 0; p0 < P; p0 += kBpChunk) { cw.push_back(w); cp.push_back((int32_t)p0); }
      wco[w + 1] = (int32_t)cw.size();
    }
    d.n_chunks = (int32_t)cw.size();
    TRY(dev_upload(ctx, pool, &d.chunk_world, cw.data(), cw.size())); TRY(dev_upload(ctx, pool, &d.chunk_p0, cp.data(), cp.size()));
    TRY(dev_upload(ctx, pool, &d.world_chunk_off, wco.data(), wco.size()));
    TRY(dev_alloc(ctx, pool, &d.chunk_cnt, cw.size())); TRY(dev_alloc(ctx, pool, &d.chunk_icnt, cw.size()));
    TRY(dev_alloc(ctx, pool, &d.chunk_base, cw.size())); TRY(dev_alloc(ctx, pool, &d.chunk_ibase, cw.size()));
  }
  TRY(dev_upload_rw(ctx, pool, &d.pos, B->pos, 3 * (size_t)nb)); TRY(dev_upload_rw(ctx, pool, &d.quat, B->quat, 4 * (size_t)nb));
  TRY(dev_upload_rw(ctx, pool, &d.vel, B->vel, 3 * (size_t)nb)); TRY(dev_upload_rw(ctx, pool, &d.angvel, B->angvel, 3 * (size_t)nb));
  TRY(dev_upload_rw(ctx, pool, &d.force, B->force, 3 * (size_t)nb)); TRY(dev_upload_rw(ctx, pool, &d.torque, B->torque, 3 * (size_t)nb));
  TRY(dev_upload_rw(ctx, pool, &d.iiw, B->inv_inertia_world, 9 * (size_t)nb));
  TRY(dev_upload(ctx, pool, &d.inv_mass, B->inv_mass, nb)); TRY(dev_upload(ctx, pool, &d.inv_inertia, B->inv_inertia, 3 * (size_t)nb));
  TRY(dev_upload(ctx, pool, &d.ldp, B->lin_damp_pow, nb)); TRY(dev_upload(ctx, pool, &d.adp, B->ang_damp_pow, nb));
  TRY(dev_upload(ctx, pool, &d.lin_lock, B->lin_lock, 3 * (size_t)nb)); TRY(dev_upload(ctx, pool, &d.ang_lock, B->ang_lock, 3 * (size_t)nb));
  TRY(dev_upload(ctx, pool, &d.radius, B->bounding_radius, nb));
  TRY(dev_upload(ctx, pool, &d.inst_off, B->inst_off, nb + 1));
  TRY(dev_upload(ctx, pool, &d.inst_shape, B->inst_shape, ni));
  TRY(dev_upload(ctx, pool, &d.inst_body, inst_body.data(), ni));
  TRY(dev_upload(ctx, pool, &d.inst_friction, B->inst_friction, ni));
  TRY(dev_upload(ctx, pool, &d.inst_bounciness, B->inst_bounciness, ni));
  TRY(dev_alloc(ctx, pool, &d.pre_pos, 3 * (size_t)nb)); TRY(dev_alloc(ctx, pool, &d.pre_quat, 4 * (size_t)nb));
  CK(cudaMemcpyAsync(d.pre_pos, d.pos, 3 * (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d.pre_quat, d.quat, 4 * (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  // world-placed shape blocks + placement work items (moving bodies first)
  const int ns = (int)ctx->h_sh_kind.size();
  std::vector<int64_t> woff(ni);
  std::vector<int32_t> inst_sub(std::max(ni, 1));
  int64_t wtotal = 0;
  for (int k = 0; k < ni; k++) {
    int s = B->inst_shape[k];
    if (s < 0 || s >= ns) { ctx->err = "shape index out of range"; return EP_EINVAL; }
    const int kind = ctx->h_sh_kind[s];
    inst_sub[k] = kind == SH_CONVEX ? (ctx->h_sh_i32[ctx->h_sh_i32_off[s]] != 0 ? 0 : 1) : kind + 1;
    woff[k] = wtotal;
    wtotal += ctx->h_sh_f64_off[s + 1] - ctx->h_sh_f64_off[s];
  }
  {
    int nb_max = 1;
    for (int w = 0; w < nw; w++) nb_max = std::max(nb_max, B->world_body_off[w + 1] - B->world_body_off[w]);
    d.nb_max = nb_max;
  }
  std::vector<double> wf(std::max<int64_t>(wtotal, 1));
  std::vector<int32_t> pl_inst, pl_code;
  int n_moving = 0;
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < ni; k++) {
      bool moving = B->kind[inst_body[k]] != EP_STATIC;
      if (moving != (pass == 0)) continue;
      int s = B->inst_shape[k];
      const double* f = ctx->h_sh_f64.data() + ctx->h_sh_f64_off[s];
      std::copy(f, f + (ctx->h_sh_f64_off[s + 1] - ctx->h_sh_f64_off[s]), wf.begin() + woff[k]);
      for (int x = ctx->h_pl_off[s]; x < ctx->h_pl_off[s + 1]; x++) { pl_inst.push_back(k); pl_code.push_back(ctx->h_pl_code[x]); }
    }
    if (pass == 0) n_moving = (int)pl_inst.size();
  }
  d.n_pl_moving = n_moving; d.n_pl_all = (int)pl_inst.size();
  TRY(dev_upload(ctx, pool, &d.inst_woff, woff.data(), ni));
  TRY(dev_upload(ctx, pool, &d.inst_sub, inst_sub.data(), ni));
  TRY(dev_upload_rw(ctx, pool, &d.wf64, wf.data(), wf.size()));
  TRY(dev_upload(ctx, pool, &d.pl_inst, pl_inst.data(), pl_inst.size()));
  TRY(dev_upload(ctx, pool, &d.pl_code, pl_code.data(), pl_code.size()));
  // params
  TRY(dev_upload(ctx, pool, &d.gravity, P->gravity, 3 * (size_t)nw));
  TRY(dev_upload(ctx, pool, &d.dt, P->dt, nw));
  TRY(dev_upload(ctx, pool, &d.iterations, P->iterations, nw));
  d.collide = nullptr; d.collide_off = nullptr;
  if (P->collide && nw > 0) {
    if (!P->collide_off) { ctx->err = "collide without collide_off"; return EP_EINVAL; }
    std::vector<int64_t> off(P->collide_off, P->collide_off + nw);
    int64_t n_last = B->world_body_off[nw] - B->world_body_off[nw - 1];
    size_t total = (size_t)(off[nw - 1] + n_last * n_last);
    TRY(dev_upload(ctx, pool, &d.collide, P->collide, total));
    TRY(dev_upload(ctx, pool, &d.collide_off, off.data(), nw));
  }
  // constraint registrations -> CSR over the registering body row, stable in list order
  const int ncn = P->n_constraints;
  d.n_constraints = ncn;
  {
    std::vector<int> order(ncn);
    for (int c = 0; c < ncn; c++) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return P->con_body_a[x] < P->con_body_a[y]; });
    std::vector<int32_t> a_off(nb + 1, 0), cb(ncn), ck(ncn);
    std::vector<double> cd(24 * (size_t)ncn);
    for (int c = 0; c < ncn; c++) {
      int a = P->con_body_a[order[c]], b = P->con_body_b[order[c]];
      if (a < 0 || a >= nb || b < 0 || b >= nb || body_world[a] != body_world[b]) { ctx->err = "constraint between bodies of different worlds"; return EP_EINVAL; }
      a_off[a + 1]++;
      cb[c] = b; ck[c] = P->con_kind[order[c]];
      std::copy(P->con_data + 24 * (size_t)order[c], P->con_data + 24 * (size_t)order[c] + 24, cd.begin() + 24 * (size_t)c);
    }
    for (int r = 0; r < nb; r++) a_off[r + 1] += a_off[r];
    TRY(dev_upload(ctx, pool, &d.con_a_off, a_off.data(), nb + 1));
    TRY(dev_upload(ctx, pool, &d.con_b, cb.data(), ncn));
    TRY(dev_upload(ctx, pool, &d.con_kind, ck.data(), ncn));
    TRY(dev_upload(ctx, pool, &d.con_data, cd.data(), cd.size()));
  }
  // capacities
  d.slot_cap = std::max<int64_t>(slot_cap, 1);
  int64_t ccap = std::max<int64_t>(1 << 14, 16 * (int64_t)nb);
  if (ctx->opt[EP_OPT_CONTACT_CAP] > 0) ccap = std::max<int64_t>(ctx->opt[EP_OPT_CONTACT_CAP], 64);
  if (warm && warm->world_group_off) ccap = std::max<int64_t>(ccap, warm->group_contact_off[warm->world_group_off[nw]]);
  d.contact_cap = (int32_t)std::min<int64_t>(ccap, (1ll << 31) - 64);
  d.jrow_cap = 6 * ncn + 8;
  // scratch
  TRY(dev_alloc(ctx, pool, &d.sorted, nb)); TRY(dev_alloc(ctx, pool, &d.sort_key, nb));
  TRY(dev_alloc(ctx, pool, &d.sbv, 6 * (size_t)nb));
  CK(cudaMemsetAsync(d.sbv, 0, 6 * (size_t)std::max(nb, 1) * sizeof(double), ctx->stream));
  TRY(dev_alloc(ctx, pool, &d.uf_parent, nb)); TRY(dev_alloc(ctx, pool, &d.uf_count, nb)); TRY(dev_alloc(ctx, pool, &d.uf_fill, nb));
  TRY(dev_alloc(ctx, pool, &d.isl_off, nb)); TRY(dev_alloc(ctx, pool, &d.isl_cnt, nb)); TRY(dev_alloc(ctx, pool, &d.isl_world, nb));
  TRY(dev_alloc(ctx, pool, &d.isl_groups, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &d.sched, d.slot_cap)); TRY(dev_alloc(ctx, pool, &d.sched_lvl, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &d.lvl_off, d.slot_cap)); TRY(dev_alloc(ctx, pool, &d.lvl_cur, d.slot_cap));
  TRY(dev_alloc(ctx, pool, &d.sched_rec, 2 * (size_t)d.slot_cap));
  TRY(dev_alloc(ctx, pool, &d.isl_key, nb)); TRY(dev_alloc(ctx, pool, &d.isl_nc, nb)); TRY(dev_alloc(ctx, pool, &d.isl_nb, nb)); TRY(dev_alloc(ctx, pool, &d.isl_nj, nb));
  TRY(dev_alloc(ctx, pool, &d.isl_root, nb)); TRY(dev_alloc(ctx, pool, &d.isl_sorted, nb));
  TRY(dev_alloc(ctx, pool, &d.key_cnt, 2 * kKeys));
  TRY(dev_alloc(ctx, pool, &d.body_iters, nb)); TRY(dev_alloc(ctx, pool, &d.body_local, nb));
  CK(cudaMemsetAsync(d.body_iters, 0, std::max(nb, 1) * sizeof(int32_t), ctx->stream));
  TRY(dev_alloc(ctx, pool, &d.tile_top, 1));
  for (int i = 0; i < 6; i++) d.tile_smem[i] = ctx->tile_smem[i];
  {
    int64_t icap = std::max<int64_t>(1 << 16, 8 * (int64_t)ni);
    if (ctx->opt[EP_OPT_ITEM_CAP] > 0) icap = std::max<int64_t>(ctx->opt[EP_OPT_ITEM_CAP], 64);
    d.item_cap = (int32_t)std::min<int64_t>(icap, (1ll << 27));   // item * kRawStride stays below 2^31
    TRY(dev_alloc(ctx, pool, &d.items, (size_t)d.item_cap + 32 * kClasses));
    TRY(dev_alloc(ctx, pool, &d.cls_off, kClasses * 8)); TRY(dev_alloc(ctx, pool, &d.cls_cur, kClasses * 8));
    TRY(dev_alloc(ctx, pool, &d.np_first, kClasses + 1)); TRY(dev_alloc(ctx, pool, &d.np_end, kClasses));
    TRY(dev_alloc(ctx, pool, &d.raw, (size_t)d.item_cap * kRawStride));
    TRY(dev_alloc(ctx, pool, &d.contact_src, (size_t)d.contact_cap));
    TRY(dev_alloc(ctx, pool, &d.sep_hint, (size_t)d.slot_cap));
    CK(cudaMemsetAsync(d.sep_hint, 0, (size_t)d.slot_cap * sizeof(int32_t), ctx->stream));
    TRY(dev_alloc(ctx, pool, &d.raw_cnt, d.item_cap));
  }
  TRY(dev_alloc(ctx, pool, &d.world_min_rem, nw)); TRY(dev_alloc(ctx, pool, &d.world_n_islands, nw));
  CK(cudaMemcpyAsync(d.world_min_rem, P->iterations, nw * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d.world_n_islands, 0, std::max(nw, 1) * sizeof(int32_t), ctx->stream));
  {   // solver family of this batch: quads for small batches (latency-bound shards), lane tiles otherwise
    const char* sv = getenv("EP_SOLVER");
    d.quads = sv && !strcmp(sv, "quads") ? 1 : (sv && !strcmp(sv, "lanes") ? 0 : ((nb <= ctx->quads_max_bodies && d.nb_max <= ctx->quads_max_world) ? 1 : 0));
  }
  d.sched_serial = getenv("EP_SCHED_SERIAL") && atoi(getenv("EP_SCHED_SERIAL")) != 0 ? 1 : 0;
  d.rows = nullptr; d.jrows = nullptr; d.qrows = nullptr; d.jqrows = nullptr; d.row_dl = nullptr; d.team_body = nullptr;
  if (d.quads) {
    TRY(dev_alloc(ctx, pool, &d.qrows, 3 * (size_t)d.contact_cap)); TRY(dev_alloc(ctx, pool, &d.jqrows, d.jrow_cap));
    TRY(dev_alloc(ctx, pool, &d.row_dl, 3 * (size_t)d.contact_cap + d.jrow_cap));
    TRY(dev_alloc(ctx, pool, &d.team_body, 8 * ((size_t)nb + nw + 1)));
  } else {
    TRY(dev_alloc(ctx, pool, &d.rows, d.contact_cap)); TRY(dev_alloc(ctx, pool, &d.jrows, d.jrow_cap));
  }
  // tile images: 37 doubles per contact + 16 per dynamic body, with slack for the max-over-lanes padding of a tile
  d.tile_arena_cap = (int64_t)d.contact_cap * 72 + (int64_t)d.jrow_cap * 24 + (int64_t)nb * 24 + 4096;
  TRY(dev_alloc(ctx, pool, &d.tile_arena, (size_t)d.tile_arena_cap));
  TRY(dev_alloc(ctx, pool, &d.work, 4));
  CK(cudaMemsetAsync(d.work, 0, 4 * sizeof(unsigned long long), ctx->stream));
  TRY(dev_alloc(ctx, pool, &d.flags, FLAG_COUNT));
  CK(cudaMemsetAsync(d.flags, 0, FLAG_COUNT * sizeof(int32_t), ctx->stream));
  TRY(alloc_frame(ctx, d.cur, d));
  TRY(alloc_frame(ctx, d.prev, d));
  {   // device image of ep_contacts (ep_output.cuh)
    OutBuf& o = d.out;
    o.group_cap = (int32_t)std::min<int64_t>(d.slot_cap, (1ll << 31) - 2); o.contact_cap = d.contact_cap;
    const size_t g = (size_t)o.group_cap + 1, c = (size_t)o.contact_cap;
    TRY(dev_alloc(ctx, pool, &o.world_group_off, (size_t)nw + 1)); TRY(dev_alloc(ctx, pool, &o.group_id1, g)); TRY(dev_alloc(ctx, pool, &o.group_id2, g));
    TRY(dev_alloc(ctx, pool, &o.group_ncon, g)); TRY(dev_alloc(ctx, pool, &o.group_contact_off, g)); TRY(dev_alloc(ctx, pool, &o.group_island, g));
    TRY(dev_alloc(ctx, pool, &o.iterations, nw)); TRY(dev_alloc(ctx, pool, &o.world_n_islands, nw));
    TRY(dev_alloc(ctx, pool, &o.group_pre_pos1, 3 * g)); TRY(dev_alloc(ctx, pool, &o.group_pre_quat1, 4 * g));
    TRY(dev_alloc(ctx, pool, &o.shape_key, c)); TRY(dev_alloc(ctx, pool, &o.feature_key, c));
    TRY(dev_alloc(ctx, pool, &o.ni, 3 * c)); TRY(dev_alloc(ctx, pool, &o.pi, 3 * c)); TRY(dev_alloc(ctx, pool, &o.pj, 3 * c));
    TRY(dev_alloc(ctx, pool, &o.friction, c)); TRY(dev_alloc(ctx, pool, &o.bounciness, c)); TRY(dev_alloc(ctx, pool, &o.lambda, c)); TRY(dev_alloc(ctx, pool, &o.t1, 3 * c));
    TRY(dev_alloc(ctx, pool, &o.wg_cnt, nw)); TRY(dev_alloc(ctx, pool, &o.wc_cnt, nw)); TRY(dev_alloc(ctx, pool, &o.wc_off, nw));
    TRY(dev_alloc(ctx, pool, &o.totals, 2));
  }
  // previous frame's contacts = warm-start cache (Physics.elm:528-529, Equation.elm:50-58)
  if (warm && warm->world_group_off && warm->world_group_off[nw] > 0) {
    int ng = warm->world_group_off[nw];
    int nc = warm->group_contact_off[ng];
    if (ng > d.slot_cap) { ctx->err = "warm-start groups exceed pair capacity"; return EP_ECAPACITY; }
    std::vector<int32_t> row1(ng), row2(ng), cstart(ng), ccount(ng), zero(ng, 0), neg(ng, -1);
    std::vector<ContactRec> crec(std::max(nc, 1));
    memset(crec.data(), 0, crec.size() * sizeof(ContactRec));
    for (int w = 0; w < nw; w++) {
      std::unordered_map<int, int> id2row;
      for (int r = B->world_body_off[w]; r < B->world_body_off[w + 1]; r++) id2row[B->body_id[r]] = r;
      for (int g = warm->world_group_off[w]; g < warm->world_group_off[w + 1]; g++) {
        auto a = id2row.find(warm->group_id1[g]), b = id2row.find(warm->group_id2[g]);
        cstart[g] = warm->group_contact_off[g];
        ccount[g] = warm->group_contact_off[g + 1] - warm->group_contact_off[g];
        if (a == id2row.end() || b == id2row.end()) { row1[g] = 0; row2[g] = 0; ccount[g] = 0; continue; }   // body gone: entries never match
        row1[g] = a->second; row2[g] = b->second;
        for (int c = cstart[g]; c < cstart[g] + ccount[g]; c++) {
          crec[c].fk = warm->feature_key[c]; crec[c].sk = warm->shape_key[c]; crec[c].slot = g;
          crec[c].lam = warm->lambda[c];
          crec[c].t1[0] = warm->t1[3 * c]; crec[c].t1[1] = warm->t1[3 * c + 1]; crec[c].t1[2] = warm->t1[3 * c + 2];
        }
      }
    }
    FrameBuf& f = d.prev;
    std::vector<int32_t> wg_base(nw), wg_n(nw);
    for (int w = 0; w < nw; w++) { wg_base[w] = warm->world_group_off[w]; wg_n[w] = warm->world_group_off[w + 1] - warm->world_group_off[w]; }
    CK(cudaMemcpyAsync(f.cand_base, wg_base.data(), nw * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.cand_n, wg_n.data(), nw * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.slot_row1, row1.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.slot_row2, row2.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.slot_cstart, cstart.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.slot_ccount, ccount.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.contacts, crec.data(), nc * sizeof(ContactRec), cudaMemcpyHostToDevice, ctx->stream));
    k_pair_slots<<<grid_for(ctx, ng, 256, 8), 256, 0, ctx->stream>>>(d, f, ng); CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
  }
  // place every shape once (static bodies are never re-placed)
  if (d.n_pl_all > 0) {
    k_place<<<grid_for(ctx, d.n_pl_all, 256, 16), 256, 0, ctx->stream>>>(d, d.n_pl_all);
    CK(cudaGetLastError());
    // static items sit after the moving ones, so the per-step launch covers [0, n_pl_moving)
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_worlds = true;
  ctx->placed_current = false;
  return EP_OK;
}

static cudaEvent_t next_event(ep_ctx* ctx) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[ctx->ev_used++];
}
// Launch bracket: when profiling, every kernel gets its own start/end event on the context's stream.
struct Timed {
  ep_ctx* ctx; int kernel; cudaStream_t st; cudaEvent_t a = nullptr;
  Timed(ep_ctx* c, int k, cudaStream_t s = nullptr, bool count = true) : ctx(c), kernel(k), st(s ? s : c->stream) {
    if (count) ctx->launches++;
    ctx->k_launches[k]++;
    if (ctx->profiling && ctx->ev_pending.size() < (1u << 16)) { a = next_event(ctx); if (a) cudaEventRecord(a, st); }
  }
  ~Timed() {
    if (!a) return;
    cudaEvent_t b = next_event(ctx);
    if (b) { cudaEventRecord(b, st); ctx->ev_pending.push_back({kernel, a, b}); }
  }
};
static void resolve_events(ep_ctx* ctx) {     // stream must be idle
  for (auto& p : ctx->ev_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->k_ms[p.kernel] += ms;
  }
  ctx->ev_pending.clear();
  for (auto& p : ctx->np_pending) { float ms = 0; if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->np_ms[p.kernel] += ms; }
  if (!ctx->np_pending.empty()) {
    int32_t cnt[CNT_COUNT];
    cudaMemcpy(cnt, ctx->d.prev.counters, sizeof cnt, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[ep] narrowphase per class (lo*6+hi; 0 box 1 convex 2 plane 3 sphere 4 capsule 5 particle): total ms over the profiled steps, items in the last step\n");
    for (int c = 0; c < kClasses; c++) {
      int n = 0; for (int b = 0; b < 8; b++) n += cnt[CNT_CLASS0 + c * 8 + b];
      if (n) fprintf(stderr, "[ep]   class %2d (%d,%d): %9.3f ms  %8d items\n", c, c / 6, c % 6, ctx->np_ms[c], n);
    }
    for (int c = 0; c < kClasses; c++) ctx->np_ms[c] = 0;
  }
  ctx->np_pending.clear();
  ctx->ev_used = 0;
}

extern "C" int ep_profile(ep_ctx* ctx, int enable) {
  if (!ctx) return EP_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  resolve_events(ctx);
  ctx->profiling = enable != 0;
  for (int i = 0; i < EP_K_COUNT; i++) { ctx->k_ms[i] = 0; ctx->k_launches[i] = 0; }
  return EP_OK;
}
extern "C" int ep_kernel_times(ep_ctx* ctx, double* ms, int64_t* launches, int32_t n) {
  if (!ctx) return EP_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  resolve_events(ctx);
  for (int i = 0; i < n && i < EP_K_COUNT; i++) { if (ms) ms[i] = ctx->k_ms[i]; if (launches) launches[i] = ctx->k_launches[i]; }
  return EP_OK;
}
extern "C" const char* ep_kernel_name(int32_t k) {
  static const char* names[EP_K_COUNT] = {"k_place", "k_sort", "k_broadphase", "k_narrowphase", "k_equations", "k_islands", "k_solve", "k_integrate", "k_solve_cluster",
                                          "k_solve_lanes_b0", "k_solve_lanes_b1", "k_solve_lanes_b2", "k_solve_lanes_b3", "k_solve_lanes_b4", "k_solve_lanes_b5", "k_solve_lanes_inplace", "k_solve_team_warp", "k_solve_team_cta"};
  return (k >= 0 && k < EP_K_COUNT) ? names[k] : "";
}

static void launch_out_kernels(ep_ctx* ctx, const FrameBuf& f, cudaStream_t sx, int early);

extern "C" int ep_step_resident(ep_ctx* ctx, int32_t n_steps) {
  if (!ctx || n_steps < 0) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds"; return EP_ESTATE; }
  CK(cudaSetDevice(ctx->device));
  Dev& d = ctx->d;
  cudaStream_t st = ctx->stream;
  if (d.n_bodies == 0) return EP_OK;
  const int gw = std::max(1, std::min(d.n_worlds, ctx->sm_count * 16));
  const int gwide = ctx->sm_count * 8;
  if (n_steps > 0) ctx->placed_current = false;
  for (int s = 0; s < n_steps; s++) {
    CK(cudaMemsetAsync(d.cur.counters, 0, CNT_COUNT * sizeof(int32_t), st));
    CK(cudaMemsetAsync(d.key_cnt, 0, 2 * kKeys * sizeof(int32_t), st));
    CK(cudaMemsetAsync(d.tile_top, 0, sizeof(unsigned long long), st));
    {
      CK(cudaMemsetAsync(d.cur.pair_slot, 0xFF, (size_t)d.slot_cap * sizeof(int32_t), st));
      if (d.n_pl_moving > 0) { Timed t(ctx, EP_K_PLACE); k_place<<<grid_for(ctx, d.n_pl_moving, 256, 16), 256, 0, st>>>(d, d.n_pl_moving); }
      { Timed t(ctx, EP_K_SORT);
        if (d.nb_max > 256 && (size_t)d.nb_max * sizeof(double) <= 48 * 1024) {      // large worlds: several CTAs per world
          const int parts = (d.nb_max + 127) / 128;
          k_sort_wide<<<(int)std::min<int64_t>((int64_t)d.n_worlds * parts, (int64_t)ctx->sm_count * 64), 128, (size_t)d.nb_max * sizeof(double), st>>>(d, parts);
        } else k_sort<<<gw, d.nb_max <= 128 ? 128 : (d.nb_max <= 512 ? 256 : 1024), 0, st>>>(d); }
      { Timed t(ctx, EP_K_BROADPHASE);
        const int gc = std::max(1, std::min(d.n_chunks, ctx->sm_count * 16));
        k_bp_count<<<gc, 128, 0, st>>>(d);
        k_bp_scan<<<grid_for(ctx, (int64_t)d.n_worlds * 32, 128, 16), 128, 0, st>>>(d);
        k_bp_emit<<<gc, 128, 0, st>>>(d);
        ctx->launches += 2; }
      static const int np_lanes_env = getenv("EP_NP_LANES") ? atoi(getenv("EP_NP_LANES")) : 0;
      // small batches (the quad solver's regime): 16 items per warp -- the kernel time is the latency of the heaviest warp, and half-full
      // warps serialise fewer divergent lanes (0.32 -> 0.28 ms at 1024 worlds; at 8192 worlds full warps are 30 % faster)
      const int np_lanes = (np_lanes_env == 8 || np_lanes_env == 16 || np_lanes_env == 32 || np_lanes_env == 4) ? np_lanes_env : (d.quads ? 16 : 32);
      static const bool np_serial = getenv("EP_NP_SERIAL") != nullptr;    // profiling aid: one launch per work-item class
      if (np_serial && ctx->profiling) {
        for (int c = 0; c < kClasses; c++) {
          cudaEvent_t a = next_event(ctx), b = next_event(ctx);
          cudaEventRecord(a, st);
          k_narrowphase<<<gwide * (32 / np_lanes), 128, 0, st>>>(d, c, np_lanes, 0ull);
          cudaEventRecord(b, st);
          ctx->np_pending.push_back({c, a, b});
        }
      } else
      {
        // large batches: one kernel per module group (8192 worlds: 1.00 -> 0.84 ms); small batches: the single kernel (equal within noise)
        // EP_NP_MODE: 0 one kernel for every class, 1 one kernel per module group side by side, 2 a quad of lanes per item for the SAT
        // classes (hull-hull, capsule-hull) beside one kernel for the rest.  Measured on B200 (C3): mode 2 wins up to 4096 worlds per
        // GPU (k_narrowphase 0.57 -> 0.44 ms at 4096, 0.27 -> 0.245 at 1024), mode 1 at 8192 (0.74 against 0.78)
        static const int np_mode_env = getenv("EP_NP_MODE") ? atoi(getenv("EP_NP_MODE")) : -1;
        const int np_mode = (np_mode_env >= 0 && np_mode_env <= 2) ? np_mode_env : (d.n_bodies <= 200000 ? 2 : 1);
        const int np_split = np_mode == 1;
        const bool np_quad = np_mode == 2;
        static const int g12 = getenv("EP_NP_GRID12") ? atoi(getenv("EP_NP_GRID12")) : 1;
        Timed t(ctx, EP_K_NARROWPHASE);
        if (np_quad) {
          CK(cudaEventRecord(ctx->ev_fork, st));
          CK(cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0));
          k_narrowphase_quad<<<gwide * 2, 128, 0, ctx->aux[0]>>>(d);
          k_narrowphase<<<gwide * (32 / np_lanes), 128, 0, st>>>(d, -1, np_lanes, kQuadClasses);
          CK(cudaEventRecord(ctx->ev_join[0], ctx->aux[0]));
          CK(cudaStreamWaitEvent(st, ctx->ev_join[0], 0));
          ctx->launches += 1;
        } else if (!np_split) k_narrowphase<<<gwide * (32 / np_lanes), 128, 0, st>>>(d, -1, np_lanes, 0ull);
        else {      // one kernel per module group, side by side: hull-hull on the main stream, capsule-hull and the rest beside it
          CK(cudaEventRecord(ctx->ev_fork, st));
          CK(cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0)); CK(cudaStreamWaitEvent(ctx->aux[1], ctx->ev_fork, 0)); CK(cudaStreamWaitEvent(ctx->aux[2], ctx->ev_fork, 0));
          static const int l0_env = getenv("EP_NP_LANES_G0") ? atoi(getenv("EP_NP_LANES_G0")) : 0;
          const int l0 = (l0_env == 4 || l0_env == 8 || l0_env == 16 || l0_env == 32) ? l0_env : np_lanes;
          k_narrowphase_group<0><<<gwide * (32 / l0), 128, 0, st>>>(d, l0);
          static const int l1_env = getenv("EP_NP_LANES_G1") ? atoi(getenv("EP_NP_LANES_G1")) : 0;
          const int l1 = (l1_env == 4 || l1_env == 8 || l1_env == 16 || l1_env == 32) ? l1_env : np_lanes;
          k_narrowphase_group<1><<<gwide * g12 * (32 / l1), 128, 0, ctx->aux[0]>>>(d, l1);
          k_narrowphase_group<2><<<gwide * g12 * (32 / np_lanes), 128, 0, ctx->aux[1]>>>(d, np_lanes);
          k_narrowphase_group<3><<<gwide * g12 * (32 / np_lanes), 128, 0, ctx->aux[2]>>>(d, np_lanes);
          CK(cudaEventRecord(ctx->ev_join[0], ctx->aux[0])); CK(cudaEventRecord(ctx->ev_join[1], ctx->aux[1])); CK(cudaEventRecord(ctx->ev_join[2], ctx->aux[2]));
          CK(cudaStreamWaitEvent(st, ctx->ev_join[0], 0)); CK(cudaStreamWaitEvent(st, ctx->ev_join[1], 0)); CK(cudaStreamWaitEvent(st, ctx->ev_join[2], 0));
          ctx->launches += 3;
        }
      }
      // k_equations and k_islands both consume k_slot_counts' result and nothing of each other.  The light one (islands: a
      // warp per small world, latency-bound) goes first on the main stream with a grid that leaves CTA slots and registers
      // free, the heavy one (equations, 168 registers) on a side stream beside it.
      static const int isl_per_sm = getenv("EP_ISL_CTAS") ? atoi(getenv("EP_ISL_CTAS")) : 12;

      if (ctx->hook_wait_in) CK(cudaStreamWaitEvent(st, ctx->ev_in, 0));     // velocities, forces, masses: first read by k_equations
      { Timed t(ctx, EP_K_ISLANDS);        // brackets k_slot_counts + k_islands + k_island_sort
        k_slot_counts<<<grid_for(ctx, d.slot_cap, 256, 8), 256, 0, st>>>(d);
        CK(cudaEventRecord(ctx->ev_fork, st));
        // one CTA per world: a warp for small worlds, more for large ones; the union-find arrays sit in shared memory when they fit
        const int it = d.nb_max <= 64 ? 32 : (d.nb_max <= 512 ? 128 : (d.nb_max <= 1024 ? 256 : 1024));
        const size_t ism = 3 * (size_t)d.nb_max * sizeof(int);
        const int use_smem = ism <= 40 * 1024;
        k_islands<<<std::max(1, std::min(d.n_worlds, ctx->sm_count * isl_per_sm)), it, use_smem ? ism : 0, st>>>(d, use_smem);
        k_island_sort<<<grid_for(ctx, d.n_bodies / 2 + 1, 256, 4), 256, 0, st>>>(d); ctx->launches += 2; }
      CK(cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0));
      { Timed t(ctx, EP_K_EQUATIONS, ctx->aux[0]);
        if (d.quads) k_equations<true><<<gwide, 128, 0, ctx->aux[0]>>>(d); else k_equations<false><<<gwide, 128, 0, ctx->aux[0]>>>(d);
        if (d.n_constraints > 0) { k_joint_equations<<<gwide, 128, 0, ctx->aux[0]>>>(d); ctx->launches++; } }
      CK(cudaEventRecord(ctx->ev_join[0], ctx->aux[0]));
      CK(cudaStreamWaitEvent(st, ctx->ev_join[0], 0));
      if (ctx->hook_early_out) {     // the frame's pair groups and contacts are complete: marshal them beside the solver
        CK(cudaEventRecord(ctx->ev_mid, st));
        CK(cudaStreamWaitEvent(ctx->cout, ctx->ev_mid, 0));
        launch_out_kernels(ctx, d.cur, ctx->cout, 1);
        CK(cudaEventRecord(ctx->ev_out_done, ctx->cout));
      }
    }
    {   // islands by size class, all launches concurrent: classes 0..3 lane-per-island from shared-memory tiles, 4..6 one warp
        // per island, 7.. one CTA per island (level-scheduled, ep_solve.cuh).  EP_K_SOLVE brackets the whole fork/join.
      Timed t(ctx, EP_K_SOLVE, nullptr, false);
      static const bool serial = getenv("EP_SOLVE_SERIAL") != nullptr;     // profiling aid: one class after the other
      const char* fr = getenv("EP_FORCE_RESUM");      // test hook, read per step: always decide the early exit from the canonical-order sum
      const int force_resum = fr && atoi(fr) != 0;
      CK(cudaEventRecord(ctx->ev_fork, st));
      int ax = 0;
      auto side = [&](int kind, auto&& launch) -> int {
        cudaStream_t sx = serial ? st : ctx->aux[ax];
        CK(cudaStreamWaitEvent(sx, ctx->ev_fork, 0));
        { Timed tt(ctx, kind, sx); launch(sx); }
        CK(cudaEventRecord(ctx->ev_join[ax], sx));
        ax++;
        return EP_OK;
      };
      const int tiles_cap = d.n_bodies / 32 + 2;
      if (d.quads) {
        // QUAD family (small batches, ep_solve_quads.cuh), longest first: giant islands (16 warps, rows in place), class 8 (4 warps),
        // a warp per island for classes 7 and 6, then the quad-per-island classes from the largest down
        auto team_smem = [](int nw, int cap_bodies, int cap_rows, int cap_groups, bool sb, bool sr) {
          const int ch = nw == 1 ? 128 : 1024;
          return (size_t)3 * ch * 4 + (size_t)(ch + 1) * 4 + (sb ? (size_t)cap_bodies * 4 + 16 + (size_t)cap_bodies * 64 : 0) +
                 (sr ? (size_t)cap_groups * 32 + (size_t)cap_rows * (192 + 8) : 0) + 64;
        };
        // the two long poles first (they get the high-priority streams): a warp per island for classes 7 and 6
        static const int team_lo = getenv("EP_TEAM_LO") ? std::max(3, std::min(kLaneBinMax + 1, atoi(getenv("EP_TEAM_LO")))) : kLaneBinMax + 1;
        for (int c = 7; c >= team_lo; c--) {
          const int cbod = (1 << c) + 2, crow = 3 << c, cgrp = 1 << c;
          TRY(side(EP_K_SOLVE_TEAM_WARP, [&](cudaStream_t sx) { k_solve_qteam<1, true, true><<<ctx->sm_count * 4, 32, team_smem(1, cbod, crow, cgrp, true, true), sx>>>(d, c, c, force_resum, cbod, crow, cgrp); }));
        }
        // classes 8 and 9 hardly occur in worlds of <= 64 bodies: small grids (each CTA loops over the islands), so that their
        // large shared-memory footprint never waits for room beside the other launches
        const int cb = d.nb_max + 2;
        const size_t need = team_smem(16, cb, 0, 0, true, false);
        if (need <= ctx->team_smem_max) TRY(side(EP_K_SOLVE_TEAM_CTA, [&](cudaStream_t sx) { k_solve_qteam<16, true, false><<<8, 512, need, sx>>>(d, kBins - 1, kBins - 1, force_resum, cb, 0, 0); }));
        else TRY(side(EP_K_SOLVE_TEAM_CTA, [&](cudaStream_t sx) { k_solve_qteam<16, false, false><<<8, 512, team_smem(16, 0, 0, 0, false, false), sx>>>(d, kBins - 1, kBins - 1, force_resum, 0, 0, 0); }));
        const int c8b = (1 << 8) + 2, c8r = 3 << 8, c8g = 1 << 8;
        TRY(side(EP_K_SOLVE_TEAM_CTA, [&](cudaStream_t sx) { k_solve_qteam<4, true, true><<<8, 128, team_smem(4, c8b, c8r, c8g, true, true), sx>>>(d, 8, 8, force_resum, c8b, c8r, c8g); }));
        for (int i = (int)ctx->quad_plan.size() - 1; i >= 0; i--) {
          const ep_ctx::QuadLaunch& q = ctx->quad_plan[i];
          if (q.lo >= team_lo) continue;                       // served by the team-of-quads launches above
          const int grid = std::max(1, std::min(q.resident, d.n_bodies / q.kq + (q.hi - q.lo + 2)));
          TRY(side(EP_K_SOLVE_LANES_B0 + q.hi, [&](cudaStream_t sx) {
            if (q.kq == 8) k_solve_quads<8><<<grid, 32, q.smem, sx>>>(d, q.lo, q.hi, q.cap_rows, q.cap_bodies, q.cap_groups);
            else if (q.kq == 4) k_solve_quads<4><<<grid, 32, q.smem, sx>>>(d, q.lo, q.hi, q.cap_rows, q.cap_bodies, q.cap_groups);
            else k_solve_quads<2><<<grid, 32, q.smem, sx>>>(d, q.lo, q.hi, q.cap_rows, q.cap_bodies, q.cap_groups);
          }));
        }
      } else {
        // longest first: CTA-per-island teams, warp-per-island teams, then the lane classes from the largest down
        // team kernels: shared memory = level-walk chunk (3 x 8T ints) + per-body last level / inverse mass / deltas of the largest world
        const size_t body_b = (size_t)((d.nb_max + 1) & ~1) * 4 + (size_t)d.nb_max * 56;
        const bool smemb = body_b + 3 * 8 * 256 * 4 <= std::min<size_t>(200 * 1024, ctx->team_smem_max);
        const size_t sm_cta = 3 * 8 * 256 * 4 + (smemb ? body_b : 0), sm_warp = 3 * 8 * 32 * 4 + (smemb ? body_b : 0);
        // giant islands (class 9) on a thread-block cluster, first: it is the longest chain of the step
        // (only for batches whose worlds are large enough to hold one: beside the 8192 small worlds of config 3 the two launches --
        // 64 + 16 CTAs with ~100 KB of shared memory each, first in line -- delayed the lane classes by 0.3 ms although they exit at once)
        const bool use_cluster = ctx->cluster_nc > 0 && d.nb_max > 64;
        const int team_hi = use_cluster ? kBins - 2 : kBins - 1;
        if (use_cluster) {
          TRY(side(EP_K_SOLVE_CLUSTER, [&](cudaStream_t sx) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ctx->cluster_nc * ctx->cluster_count); cfg.blockDim = dim3(kClusterThreads);
            cfg.dynamicSmemBytes = cluster_smem_bytes(kClusterWarps, ctx->cluster_ws, ctx->cluster_bodies); cfg.stream = sx;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ctx->cluster_nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_solve_cluster<kClusterWarps, false>, d, kBins - 1, kBins - 1, force_resum, ctx->cluster_ws, ctx->cluster_bodies, ctx->solo_bodies);
          }));
          if (ctx->solo_bodies > 0)
            TRY(side(EP_K_SOLVE_CLUSTER, [&](cudaStream_t sx) {
              k_solve_cluster<kClusterWarps, true><<<16, kClusterThreads, cluster_smem_bytes(kClusterWarps, ctx->cluster_ws, ctx->solo_bodies), sx>>>(d, kBins - 1, kBins - 1, force_resum, ctx->cluster_ws, ctx->solo_bodies, ctx->solo_bodies);
            }));
        }
        if (smemb) {
          TRY(side(EP_K_SOLVE_TEAM_CTA, [&](cudaStream_t sx) { k_solve_team<256, true><<<ctx->sm_count * 2, 256, sm_cta, sx>>>(d, 8, team_hi, force_resum); }));
          TRY(side(EP_K_SOLVE_TEAM_WARP, [&](cudaStream_t sx) { k_solve_team<32, true><<<ctx->sm_count * 16, 32, sm_warp, sx>>>(d, kLaneBinMax + 1, 7, force_resum); }));
        } else {
          TRY(side(EP_K_SOLVE_TEAM_CTA, [&](cudaStream_t sx) { k_solve_team<256, false><<<ctx->sm_count * 2, 256, sm_cta, sx>>>(d, 8, team_hi, force_resum); }));
          TRY(side(EP_K_SOLVE_TEAM_WARP, [&](cudaStream_t sx) { k_solve_team<32, false><<<ctx->sm_count * 16, 32, sm_warp, sx>>>(d, kLaneBinMax + 1, 7, force_resum); }));
        }
        // the lane classes: every warp builds its own tile (bodies, descriptors, rows gathered from k_equations' records) and solves it
        TRY(side(EP_K_SOLVE_LANES_B5, [&](cudaStream_t sx) { k_solve_lanes<1><<<std::min(ctx->head_grid[5], tiles_cap), 32, ctx->tile_smem[5], sx>>>(d, 5, 5); }));
        TRY(side(EP_K_SOLVE_LANES_B4, [&](cudaStream_t sx) { k_solve_lanes<1><<<std::min(ctx->head_grid[4], tiles_cap), 32, ctx->tile_smem[4], sx>>>(d, 4, 4); }));
        TRY(side(EP_K_SOLVE_LANES_B3, [&](cudaStream_t sx) { k_solve_lanes<1><<<std::min(ctx->head_grid[3], tiles_cap), 32, ctx->tile_smem[3], sx>>>(d, 0, 3); }));
        for (int i = 2; i >= 0; i--)
          TRY(side(EP_K_SOLVE_LANES_B0 + i, [&](cudaStream_t sx) { k_solve_lanes<2><<<std::min(ctx->head_grid[i], tiles_cap), 32, ctx->tile_smem[i], sx>>>(d, i, i); }));
        { Timed tt(ctx, EP_K_SOLVE_LANES_INPLACE, st); k_solve_lanes<0><<<std::min(ctx->lanes_grid, tiles_cap / 4 + 1), 128, 0, st>>>(d, 0, kLaneBinMax); }
      }
      CK(cudaGetLastError());       // a solver launch that failed must not be followed by the integration of unsolved islands
      for (int i = 0; i < ax; i++) CK(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
    }
    if (ctx->hook_early_out) CK(cudaStreamWaitEvent(st, ctx->ev_out_done, 0));     // k_out_groups reads the pre-step transforms
    { Timed t(ctx, EP_K_INTEGRATE); k_integrate<<<grid_for(ctx, d.n_bodies, 256, 16), 256, 0, st>>>(d); }
    CK(cudaGetLastError());
    std::swap(d.cur, d.prev);     // the frame just computed becomes the warm-start cache
  }
  return EP_OK;
}

// Wait for a stream without spinning a host core (EP_BLOCKING_SYNC=0: spin): a host that drives one context per GPU of an
// 8-GPU box from 2 threads each has as many waiting threads as cores.
static int wait_stream(ep_ctx* ctx, cudaStream_t sx) {
  static const bool blocking = !(getenv("EP_BLOCKING_SYNC") && atoi(getenv("EP_BLOCKING_SYNC")) == 0);
  if (!blocking) { CK(cudaStreamSynchronize(sx)); return EP_OK; }
  if (!ctx->ev_block) CK(cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  CK(cudaEventRecord(ctx->ev_block, sx));
  CK(cudaEventSynchronize(ctx->ev_block));
  return EP_OK;
}

static int check_flags(ep_ctx* ctx) {
  int32_t fl[FLAG_COUNT];
  CK(cudaMemcpyAsync(fl, ctx->d.flags, sizeof fl, cudaMemcpyDeviceToHost, ctx->stream));
  TRY(wait_stream(ctx, ctx->stream));
  if (fl[FLAG_POOL_OVERFLOW]) { ctx->err = "contact pool overflow (raise EP_CONTACT_CAP)"; return EP_ECAPACITY; }
  if (fl[FLAG_PAIR_OVERFLOW] == 5) { ctx->err = "narrowphase work-item pool overflow (raise EP_OPT_ITEM_CAP)"; return EP_ECAPACITY; }
  if (fl[FLAG_PAIR_OVERFLOW]) { ctx->err = "pair slot / per-pair contact buffer overflow"; return EP_ECAPACITY; }
  if (fl[FLAG_POLY_OVERFLOW]) { ctx->err = "clip polygon buffer overflow (face with too many vertices)"; return EP_ECAPACITY; }
  if (fl[FLAG_JOINT_OVERFLOW]) { ctx->err = "joint row pool overflow"; return EP_ECAPACITY; }
  return EP_OK;
}

extern "C" int ep_sync(ep_ctx* ctx) {
  if (!ctx) return EP_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (!ctx->have_worlds) return EP_OK;
  return check_flags(ctx);
}

extern "C" int ep_download_bodies(ep_ctx* ctx, ep_bodies* B) {
  if (!ctx || !B) return EP_EINVAL;
  if (!ctx->have_worlds || B->n_bodies != ctx->d.n_bodies) { ctx->err = "body count mismatch"; return EP_ESTATE; }
  CK(cudaSetDevice(ctx->device));
  const Dev& d = ctx->d;
  size_t nb = d.n_bodies;
  cudaStream_t st = ctx->stream;
  const uint32_t down = EP_F_ALL;
  auto D2H = [&](double* dst, const double* src, size_t n) { return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st); };
  if (down & EP_F_POS) CK(D2H(B->pos, d.pos, 3 * nb));
  if (down & EP_F_QUAT) CK(D2H(B->quat, d.quat, 4 * nb));
  if (down & EP_F_VEL) CK(D2H(B->vel, d.vel, 3 * nb));
  if (down & EP_F_ANGVEL) CK(D2H(B->angvel, d.angvel, 3 * nb));
  if (down & EP_F_FORCE) CK(D2H(B->force, d.force, 3 * nb));          // zeros for every non-static body (SolverBody.elm:139-141, 221-223)
  if (down & EP_F_TORQUE) CK(D2H(B->torque, d.torque, 3 * nb));
  if (down & EP_F_INV_INERTIA_WORLD) CK(D2H(B->inv_inertia_world, d.iiw, 9 * nb));
  return check_flags(ctx);
}

// Compact a frame on the device into the layout of ep_contacts (ep_output.cuh)
static void launch_out_kernels(ep_ctx* ctx, const FrameBuf& f, cudaStream_t sx, int early) {
  const Dev& d = ctx->d;
  const int nw = d.n_worlds;
  k_out_count<<<grid_for(ctx, (int64_t)nw * 32, 128, 16), 128, 0, sx>>>(d, f);
  k_out_scan<<<1, 1024, 0, sx>>>(d);
  k_out_groups<<<grid_for(ctx, (int64_t)nw * 32, 128, 16), 128, 0, sx>>>(d, f, early);
  k_out_contacts<<<grid_for(ctx, d.contact_cap, 256, 8), 256, 0, sx>>>(d, f, early);
  ctx->launches += 4;
}
// ... and copy it into the caller's ep_contacts (one D2H per array the caller holds) on stream sx
static int copy_out_contacts(ep_ctx* ctx, const FrameBuf& f, ep_contacts* out, cudaStream_t st) {
  const Dev& d = ctx->d;
  const int nw = d.n_worlds;
  int32_t tot[2], cnt[CNT_COUNT];
  CK(cudaMemcpyAsync(tot, d.out.totals, sizeof tot, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(cnt, f.counters, sizeof cnt, cudaMemcpyDeviceToHost, st));
  TRY(wait_stream(ctx, st));
  const size_t ng = (size_t)tot[0], nc = (size_t)tot[1];
  if ((int64_t)ng > d.out.group_cap || (int64_t)nc > d.out.contact_cap) { ctx->err = "device contact image overflow"; return EP_ECAPACITY; }
  if ((int64_t)ng > out->group_cap || (int64_t)nc > out->contact_cap) { ctx->err = "ep_contacts capacity too small"; return EP_ECAPACITY; }
  cudaError_t e = cudaSuccess;
  auto D2H = [&](void* dst, const void* src, size_t bytes) { if (dst && bytes && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st); };
  const OutBuf& o = d.out;
  out->n_worlds = nw;
  D2H(out->world_group_off, o.world_group_off, ((size_t)nw + 1) * 4);
  D2H(out->group_id1, o.group_id1, ng * 4); D2H(out->group_id2, o.group_id2, ng * 4);
  D2H(out->group_n_constraints, o.group_ncon, ng * 4); D2H(out->group_contact_off, o.group_contact_off, (ng + 1) * 4);
  D2H(out->group_pre_pos1, o.group_pre_pos1, ng * 24); D2H(out->group_pre_quat1, o.group_pre_quat1, ng * 32);
  D2H(out->shape_key, o.shape_key, nc * 4); D2H(out->feature_key, o.feature_key, nc * 8);
  D2H(out->ni, o.ni, nc * 24); D2H(out->pi, o.pi, nc * 24); D2H(out->pj, o.pj, nc * 24);
  D2H(out->friction, o.friction, nc * 8); D2H(out->bounciness, o.bounciness, nc * 8);
  D2H(out->lambda, o.lambda, nc * 8); D2H(out->t1, o.t1, nc * 24);
  D2H(out->iterations, o.iterations, (size_t)nw * 4); D2H(out->world_n_islands, o.world_n_islands, (size_t)nw * 4);
  D2H(out->group_island, o.group_island, ng * 4);
  if (e != cudaSuccess) { ctx->err = std::string("download: ") + cudaGetErrorString(e); return EP_ECUDA; }
  ctx->stats[EP_STAT_CANDIDATE_PAIRS] = cnt[CNT_CAND]; ctx->stats[EP_STAT_GROUPS] = (int64_t)ng; ctx->stats[EP_STAT_CONTACTS] = (int64_t)nc;
  ctx->stats[EP_STAT_ISLANDS] = cnt[CNT_ISLANDS];
  return EP_OK;
}
// Compact the last computed frame on the device and copy it into the caller's ep_contacts (one D2H per array).
static int download_contacts_async(ep_ctx* ctx, ep_contacts* out) {
  launch_out_kernels(ctx, ctx->d.prev, ctx->stream, 0);
  CK(cudaGetLastError());
  return copy_out_contacts(ctx, ctx->d.prev, out, ctx->stream);
}

extern "C" int ep_download_contacts(ep_ctx* ctx, ep_contacts* out) {
  if (!ctx || !out) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds"; return EP_ESTATE; }
  CK(cudaSetDevice(ctx->device));
  TRY(check_flags(ctx));
  TRY(download_contacts_async(ctx, out));
  CK(cudaStreamSynchronize(ctx->stream));
  return EP_OK;
}

extern "C" void* ep_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
extern "C" void ep_host_free(void* p) { if (p) cudaFreeHost(p); }

// one gate per device, shared by every context of the process (see ep_step_frame)
struct FrameGate { std::mutex mu; cudaEvent_t ev = nullptr; };
static FrameGate& frame_gate(int device) {
  static FrameGate gates[64];
  return gates[device < 0 || device >= 64 ? 0 : device];
}

extern "C" int ep_step_frame(ep_ctx* ctx, ep_bodies* B, ep_contacts* out) { return ep_step_frame_fields(ctx, B, out, EP_F_ALL, EP_F_ALL); }

extern "C" int ep_step_frame_fields(ep_ctx* ctx, ep_bodies* B, ep_contacts* out, uint32_t up, uint32_t down) {
  if (!ctx || !B) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds (ep_upload_worlds first)"; return EP_ESTATE; }
  Dev& d = ctx->d;
  if (B->n_bodies != d.n_bodies || B->n_worlds != d.n_worlds || B->n_insts != d.n_insts) { ctx->err = "ep_step_frame: topology differs from the resident batch"; return EP_ESTATE; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nb = d.n_bodies;
  if (nb == 0) return EP_OK;
  static const bool trace = getenv("EP_TRACE_FRAME") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t_last = now();
  auto lap = [&](int k) { if (trace) { auto t = now(); ctx->tr_us[k] += std::chrono::duration<double, std::micro>(t - t_last).count(); t_last = t; } };
  // The frame as a pipeline over three streams (EP_FRAME_PIPELINE=0: everything on the main stream):
  //   main     H2D of what the front kernels read (pos, quat, bounding radius) | place .. narrowphase | rows, islands | solver | integrate | D2H of the state
  //   copy-in  H2D of everything else (80 % of the bytes), awaited by the first kernel that reads it (k_equations)
  //   copy-out when the caller does not ask for solver results in `out` (lambda, t1, iterations, world_n_islands NULL: all
  //            that Physics.contactPoints needs is known once the rows are built): marshalling + D2H of the contacts beside the solver
  static const bool pipelined = !(getenv("EP_FRAME_PIPELINE") && atoi(getenv("EP_FRAME_PIPELINE")) == 0);
  const bool early_out = pipelined && out && !out->lambda && !out->t1 && !out->iterations && !out->world_n_islands;
  cudaStream_t sin = pipelined ? ctx->cin : st;
  auto H2D = [&](cudaStream_t sx, const double* dst, const double* src, size_t n) { return cudaMemcpyAsync((void*)dst, src, n * sizeof(double), cudaMemcpyHostToDevice, sx); };
  // what the front kernels read goes first, on the main stream; the rest on the copy-in stream beside them.  Arrays the caller
  // did not touch since the last frame are not sent: the resident copy is what the last frame returned (`up`, EP_F_*)
  if (up & EP_F_POS) CK(H2D(st, d.pos, B->pos, 3 * nb));
  if (up & EP_F_QUAT) CK(H2D(st, d.quat, B->quat, 4 * nb));
  if (up & EP_F_CONSTANTS) CK(H2D(st, d.radius, B->bounding_radius, nb));
  if (up & EP_F_VEL) CK(H2D(sin, d.vel, B->vel, 3 * nb));
  if (up & EP_F_ANGVEL) CK(H2D(sin, d.angvel, B->angvel, 3 * nb));
  if (up & EP_F_FORCE) CK(H2D(sin, d.force, B->force, 3 * nb));
  if (up & EP_F_TORQUE) CK(H2D(sin, d.torque, B->torque, 3 * nb));
  if (up & EP_F_INV_INERTIA_WORLD) CK(H2D(sin, d.iiw, B->inv_inertia_world, 9 * nb));
  if (up & EP_F_CONSTANTS) {
    CK(H2D(sin, d.inv_mass, B->inv_mass, nb)); CK(H2D(sin, d.inv_inertia, B->inv_inertia, 3 * nb)); CK(H2D(sin, d.ldp, B->lin_damp_pow, nb)); CK(H2D(sin, d.adp, B->ang_damp_pow, nb));
    CK(H2D(sin, d.lin_lock, B->lin_lock, 3 * nb)); CK(H2D(sin, d.ang_lock, B->ang_lock, 3 * nb));
  }
  if (pipelined) CK(cudaEventRecord(ctx->ev_in, ctx->cin));
  lap(0);
  {
    // Several contexts on one device (sub-batches driven from their own host threads) take turns on the SMs when the
    // contact read-back FOLLOWS the step: the step of a context waits for the step of the context before it, so that it
    // runs beside the other contexts' copies instead of beside their steps (left alone the contexts fall into lock step --
    // all stepping, then all copying -- measured 49.6 -> 56.0 M body-steps/s for 2 x 4096 worlds).  When the read-back
    // already runs beside the frame's own solver the gate only serialises: 68.1 M without, 59.5 M with.
    static const int gate_env = getenv("EP_FRAME_GATE") ? atoi(getenv("EP_FRAME_GATE")) : -1;
    const bool gated = gate_env < 0 ? !early_out : gate_env != 0;
    FrameGate& g = frame_gate(ctx->device);
    std::unique_lock<std::mutex> lock(g.mu, std::defer_lock);
    if (gated) {
      lock.lock();
      if (!g.ev) CK(cudaEventCreateWithFlags(&g.ev, cudaEventDisableTiming));
      else CK(cudaStreamWaitEvent(st, g.ev, 0));
    }
    // static bodies may have been moved by the caller between frames: place every shape, not only the moving ones
    if (d.n_pl_all > d.n_pl_moving) { k_place<<<grid_for(ctx, d.n_pl_all, 256, 16), 256, 0, st>>>(d, d.n_pl_all); ctx->launches++; }
    ctx->hook_wait_in = pipelined; ctx->hook_early_out = early_out;
    const int rc = ep_step_resident(ctx, 1);
    ctx->hook_wait_in = false; ctx->hook_early_out = false;
    if (rc != EP_OK) return rc;
    if (gated) CK(cudaEventRecord(g.ev, st));
  }
  lap(1);
  auto D2H = [&](double* dst, const double* src, size_t n) { return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st); };
  if (down & EP_F_POS) CK(D2H(B->pos, d.pos, 3 * nb));
  if (down & EP_F_QUAT) CK(D2H(B->quat, d.quat, 4 * nb));
  if (down & EP_F_VEL) CK(D2H(B->vel, d.vel, 3 * nb));
  if (down & EP_F_ANGVEL) CK(D2H(B->angvel, d.angvel, 3 * nb));
  if (down & EP_F_FORCE) CK(D2H(B->force, d.force, 3 * nb));          // zeros for every non-static body (SolverBody.elm:139-141, 221-223)
  if (down & EP_F_TORQUE) CK(D2H(B->torque, d.torque, 3 * nb));
  if (down & EP_F_INV_INERTIA_WORLD) CK(D2H(B->inv_inertia_world, d.iiw, 9 * nb));
  lap(2);
  if (out) {
    if (early_out) {       // marshalled in the middle of the step on the copy-out stream (the frame is d.prev by now)
      TRY(copy_out_contacts(ctx, d.prev, out, ctx->cout));
      TRY(wait_stream(ctx, ctx->cout));
    } else TRY(download_contacts_async(ctx, out));
  }
  lap(4);
  const int rc_flags = check_flags(ctx);
  lap(5);
  if (trace) ctx->tr_frames++;
  return rc_flags;
}

extern "C" int ep_step(ep_ctx* ctx, ep_bodies* bodies, const ep_params* params, const ep_contacts* warm_in, ep_contacts* out, int32_t n_steps) {
  TRY(ep_upload_worlds(ctx, bodies, params, warm_in));
  TRY(ep_step_resident(ctx, n_steps));
  TRY(ep_download_bodies(ctx, bodies));
  if (out) TRY(ep_download_contacts(ctx, out));
  return EP_OK;
}

static int raycast_impl(ep_ctx* ctx, int32_t n_rays, const int32_t* ray_ref, const double* from, const double* direction,
                        int32_t* hit_body, double* hit_distance, double* hit_point, double* hit_normal, int attached) {
  if (!ctx || n_rays < 0 || (n_rays > 0 && (!ray_ref || !from || !direction || !hit_body || !hit_distance || !hit_point || !hit_normal))) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds"; return EP_ESTATE; }
  if (n_rays == 0) return EP_OK;
  Dev& d = ctx->d;
  const int limit = attached ? d.n_bodies : d.n_worlds;
  for (int i = 0; i < n_rays; i++)
    if (ray_ref[i] < 0 || ray_ref[i] >= limit) { ctx->err = attached ? "ray_body out of range" : "ray_world out of range"; return EP_EINVAL; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)n_rays, bytes = n * (4 + 24 + 24 + 4 + 8 + 24 + 24) + 256;
  if (bytes > ctx->ray_cap) {
    CK(cudaStreamSynchronize(st));
    if (ctx->ray_buf) cudaFree(ctx->ray_buf);
    ctx->ray_buf = nullptr; ctx->ray_cap = 0;
    CK(cudaMalloc(&ctx->ray_buf, bytes * 2));
    ctx->ray_cap = bytes * 2;
  }
  // layout: doubles first (from, direction, distance, point, normal), then the two int arrays
  double* d_from = (double*)ctx->ray_buf; double* d_dir = d_from + 3 * n; double* d_dist = d_dir + 3 * n;
  double* d_point = d_dist + n; double* d_normal = d_point + 3 * n;
  int* d_world = (int*)(d_normal + 3 * n); int* d_hit = d_world + n;
  CK(cudaMemcpyAsync(d_from, from, 24 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_dir, direction, 24 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_world, ray_ref, 4 * n, cudaMemcpyHostToDevice, st));
  // worldShapesWithMaterials follow the bodies' CURRENT transforms (SolverBody.elm:212): the placed copies are refreshed at
  // the start of a step, so they are refreshed here for the state the last step left -- once: further casts against the
  // same state (the four wheels of a car, every frame) reuse them
  if (d.n_pl_all > 0 && !ctx->placed_current) { k_place<<<grid_for(ctx, d.n_pl_all, 256, 16), 256, 0, st>>>(d, d.n_pl_all); ctx->launches++; }
  ctx->placed_current = true;
  k_raycast<<<grid_for(ctx, n_rays, 128, 16), 128, 0, st>>>(d, n_rays, d_world, d_from, d_dir, d_hit, d_dist, d_point, d_normal, attached);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(hit_body, d_hit, 4 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hit_distance, d_dist, 8 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hit_point, d_point, 24 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hit_normal, d_normal, 24 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return EP_OK;
}
extern "C" int ep_raycast(ep_ctx* ctx, int32_t n_rays, const int32_t* ray_world, const double* from, const double* direction,
                          int32_t* hit_body, double* hit_distance, double* hit_point, double* hit_normal) {
  return raycast_impl(ctx, n_rays, ray_world, from, direction, hit_body, hit_distance, hit_point, hit_normal, 0);
}
extern "C" int ep_raycast_attached(ep_ctx* ctx, int32_t n_rays, const int32_t* ray_body, const double* local_from, const double* local_direction,
                                   int32_t* hit_body, double* hit_distance, double* hit_point, double* hit_normal) {
  return raycast_impl(ctx, n_rays, ray_body, local_from, local_direction, hit_body, hit_distance, hit_point, hit_normal, 1);
}

// The 22 doubles per body that a step writes (SolverBody.elm:97-225: transform3d 7, velocity 3, angularVelocity 3,
// invInertiaWorld 9), packed [n_bodies][22] into DEVICE memory of the caller on stream `stream` (NULL: the context's): the
// send buffer of the one collective of a sharded batch, the final state gather (ncclAllGather), without host staging.
namespace ep {
__global__ void k_pack_state(Dev d, double* out) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < d.n_bodies; r += gridDim.x * blockDim.x) {
    double* o = out + 22 * (size_t)r;
#pragma unroll
    for (int k = 0; k < 3; k++) { o[k] = d.pos[3 * (size_t)r + k]; o[7 + k] = d.vel[3 * (size_t)r + k]; o[10 + k] = d.angvel[3 * (size_t)r + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) o[3 + k] = d.quat[4 * (size_t)r + k];
#pragma unroll
    for (int k = 0; k < 9; k++) o[13 + k] = d.iiw[9 * (size_t)r + k];
  }
}
}  // namespace ep
extern "C" int ep_pack_state_device(ep_ctx* ctx, void* dst_device, void* stream) {
  if (!ctx || !dst_device) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds"; return EP_ESTATE; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t sx = stream ? (cudaStream_t)stream : ctx->stream;
  if (ctx->d.n_bodies == 0) return EP_OK;
  if (sx != ctx->stream) {     // order behind the steps already enqueued on the context's stream
    CK(cudaEventRecord(ctx->ev_mid, ctx->stream));
    CK(cudaStreamWaitEvent(sx, ctx->ev_mid, 0));
  }
  k_pack_state<<<grid_for(ctx, ctx->d.n_bodies, 256, 16), 256, 0, sx>>>(ctx->d, (double*)dst_device);
  ctx->launches++;
  CK(cudaGetLastError());
  return EP_OK;
}

extern "C" int ep_apply(ep_ctx* ctx, int32_t n_ops, const int32_t* kind, const int32_t* body_row, const double* a, const double* b) {
  if (!ctx || n_ops < 0 || (n_ops > 0 && (!kind || !body_row || !a || !b))) return EP_EINVAL;
  if (!ctx->have_worlds) { ctx->err = "no resident worlds"; return EP_ESTATE; }
  if (n_ops == 0) return EP_OK;
  Dev& d = ctx->d;
  for (int k = 0; k < n_ops; k++)
    if (body_row[k] < 0 || body_row[k] >= d.n_bodies || kind[k] < 0 || kind[k] > 5) { ctx->err = "ep_apply: body row or kind out of range"; return EP_EINVAL; }
  // stable sort by body: operations on one body keep their list order and form one run
  std::vector<int> order(n_ops);
  for (int k = 0; k < n_ops; k++) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return body_row[x] < body_row[y]; });
  std::vector<int32_t> h_kind(n_ops), h_row(n_ops), run_off;
  std::vector<double> h_a(3 * (size_t)n_ops), h_b(3 * (size_t)n_ops);
  for (int k = 0; k < n_ops; k++) {
    const int s = order[k];
    h_kind[k] = kind[s]; h_row[k] = body_row[s];
    for (int q = 0; q < 3; q++) { h_a[3 * (size_t)k + q] = a[3 * (size_t)s + q]; h_b[3 * (size_t)k + q] = b[3 * (size_t)s + q]; }
    if (k == 0 || h_row[k] != h_row[k - 1]) run_off.push_back(k);
  }
  const int n_runs = (int)run_off.size();
  run_off.push_back(n_ops);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)n_ops, bytes = n * (24 + 24 + 4 + 4) + (size_t)(n_runs + 1) * 4 + 256;
  if (bytes > ctx->ray_cap) {
    CK(cudaStreamSynchronize(st));
    if (ctx->ray_buf) cudaFree(ctx->ray_buf);
    ctx->ray_buf = nullptr; ctx->ray_cap = 0;
    CK(cudaMalloc(&ctx->ray_buf, bytes * 2));
    ctx->ray_cap = bytes * 2;
  }
  double* d_a = (double*)ctx->ray_buf; double* d_b = d_a + 3 * n;
  int* d_kind = (int*)(d_b + 3 * n); int* d_row = d_kind + n; int* d_run = d_row + n;
  CK(cudaMemcpyAsync(d_a, h_a.data(), 24 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_b, h_b.data(), 24 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_kind, h_kind.data(), 4 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_row, h_row.data(), 4 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_run, run_off.data(), 4 * (size_t)(n_runs + 1), cudaMemcpyHostToDevice, st));
  k_apply<<<grid_for(ctx, n_runs, 128, 16), 128, 0, st>>>(d, n_runs, d_run, d_kind, d_row, d_a, d_b);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));     // the host vectors go out of scope
  return EP_OK;
}

extern "C" int ep_stats(ep_ctx* ctx, int64_t* out, int32_t n) {
  if (!ctx || !out) return EP_EINVAL;
  ctx->stats[EP_STAT_KERNEL_LAUNCHES] = ctx->launches;
  if (ctx->have_worlds) {
    unsigned long long wk[4] = {0, 0, 0, 0};
    cudaStreamSynchronize(ctx->stream);
    if (cudaMemcpy(wk, ctx->d.work, sizeof wk, cudaMemcpyDeviceToHost) == cudaSuccess) {
      ctx->stats[EP_STAT_CONTACT_ITERATIONS] = (int64_t)wk[0]; ctx->stats[EP_STAT_JOINT_ROW_ITERATIONS] = (int64_t)wk[1];
      ctx->stats[EP_STAT_RESUMS] = (int64_t)wk[2];
    }
    int32_t cnt[CNT_COUNT];      // counts of the last computed frame (the group count comes with a contact download only)
    if (cudaMemcpy(cnt, ctx->d.prev.counters, sizeof cnt, cudaMemcpyDeviceToHost) == cudaSuccess) {
      ctx->stats[EP_STAT_CANDIDATE_PAIRS] = cnt[CNT_CAND]; ctx->stats[EP_STAT_CONTACTS] = std::min(cnt[CNT_CONTACTS], ctx->d.contact_cap);
      ctx->stats[EP_STAT_ISLANDS] = cnt[CNT_ISLANDS];
    }
  }
  for (int i = 0; i < n && i < EP_STAT_COUNT; i++) out[i] = ctx->stats[i];
  return EP_OK;
}
