#!/usr/bin/env python
"""bench.py — body-steps/s of the hot path behind Physics.simulate on BASELINE.json config 3: the batch of 8192
independent worlds x (32 mixed dynamic bodies + a plane), sharded by world index over the N GPUs (strong scaling: the
headline; the weak-scaling figure, 8192 worlds per GPU, is reported beside it for N > 1).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement on all host threads

A "step" is one simulate call over every resident world (sort, broadphase, narrowphase, rows + warm start,
islands, PGS, integrate).  `value` is the device-timed whole-job throughput with state resident in HBM;
`e2e` is the same metric through the host-buffer C-ABI call (ep_step_frame: H2D of the bodies' state, one step,
D2H of the state + the frame's contacts, every step).  The oracle (oracle/, test infrastructure) is used here
only as the cpu_baseline / --impl reference leg and as the parity checker of the timed run.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "body_steps_per_sec"
UNIT = "body-steps/s"
BODIES_PER_WORLD = 32
WORKLOAD = "C3: 8192 independent worlds x (32 mixed dynamic bodies + plane), 20 iterations, dt 1/60, sharded by world index"


def env_int(name, default):
    return int(os.environ.get(name, default))


# ---- algorithmic HBM bytes per launch of each kernel (DESIGN.md "Kernels"; SURVEY §8d per-unit figures) ----------
def kernel_bytes(name, c):
    nb, nd, K, P, G = c["bodies"], c["dynamic_bodies"], c["contacts"], c["candidate_pairs"], c["groups"]
    table = {
        # read pos+quat (56 B) per moving body, write the placed shape block (avg `placed_f64` doubles) per body
        "k_place": nd * (56 + 8 * c["placed_f64_per_body"]),
        # read pos (24 B) + write key (8) + read keys + write perm (4)
        "k_sort": nb * (24 + 8 + 8 + 4),
        # per body: pos, radius, kind (36 B); per candidate: 8 slot words (32 B)
        "k_broadphase": nb * 36 + P * 32,
        # per candidate: 2 placed shape blocks read; per contact: ContactRec written (104 B)
        "k_narrowphase": P * 2 * 8 * c["placed_f64_per_body"] + K * 104,
        # per group: 2 bodies x 41 doubles; per contact: raw contact read (88) + warm-start read (56) + ContactRec (104) and
        # RowRec (464: the three rows + their I.w products) written
        "k_equations": G * 2 * 328 + K * (88 + 56 + 104 + 464),
        # per candidate slot: rows/kinds/counts (20 B) read, root written (4), island list (4)
        "k_islands": P * 28 + nb * 12,
        # per contact: the row record read once (464 B, gathered by the solving warp itself into its tile) + 3 lambdas written
        # (24 B); per dynamic body: inverse mass read (8 B), delta-v written (48 B) -- SURVEY §8d counts rows as resident in
        # shared memory / registers across the iterations
        "k_solve": K * 488 + nd * 56,
        # 512 B per body-step (336 read + 176 written)
        "k_integrate": nb * 512,
        # giant islands on a thread-block cluster (inside k_solve's bracket; its bytes are counted there)
        "k_solve_cluster": 0,
    }
    return table[name]


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/), scaled to
# this workload where the capture was taken on a 2048-world batch; None where no capture exists yet.
TRAFFIC_NCU = {}
try:
    TRAFFIC_NCU = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic_per_launch.json")))["bytes_per_launch_8192_worlds"]
    TRAFFIC_NCU["k_narrowphase"] = TRAFFIC_NCU.get("k_narrowphase", TRAFFIC_NCU.get("k_narrowphase_group"))
except Exception:  # noqa: BLE001
    pass


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self._stop_evt = [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as ex:  # noqa: BLE001
            self.err = str(ex)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_run(shapes, bodies, params, warm, steps, threads):
    """`steps` oracle steps over all worlds of `bodies`, worlds split evenly over `threads` OS threads
    (ctypes releases the GIL; the restatement is single-threaded per world like the reference).
    Returns (seconds, [(Bodies, Contacts)] per chunk)."""
    from elm_physics_b200 import abi, pack
    from oracle import binding as oracle
    import helpers as H
    oracle.lib()
    nw = bodies.n_worlds
    threads = max(1, min(threads, nw))
    cuts = [nw * i // threads for i in range(threads + 1)]
    chunks = []
    for i in range(threads):
        b = bodies.slice_worlds(cuts[i], cuts[i + 1])
        p = abi.slice_params(params, cuts[i], cuts[i + 1])
        w = None
        if warm is not None:
            w = pack.contacts_from_world_dicts([pack.contacts_world_dict(warm, k) for k in range(cuts[i], cuts[i + 1])])
        chunks.append([b, p, w, H.contacts_for(b)])
    def work(c):
        oracle.step(shapes, c[0], c[1], c[2], c[3], steps)
    ts = [threading.Thread(target=work, args=(c,)) for c in chunks]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, chunks, cuts


def run_reference(args, rank, world_size):
    """--impl reference: the CPU restatement of the reference's own path (the reference is pure Elm and cannot be
    built here: no elm/node), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from elm_physics_b200 import scenes
    cores = host_threads()
    sample_worlds = args.worlds                      # the same batch as the GPU arm (same generator, same seeds)
    shapes, bodies, params = scenes.mixed_worlds(sample_worlds)
    import helpers as H
    from elm_physics_b200 import pack
    # pre-roll on the CPU (untimed) so the timed steps see the same settled regime as the GPU arm
    _, chunks, cuts = oracle_run(shapes, bodies, params, None, args.preroll, cores)
    state = [(c[0], c[1], c[3]) for c in chunks]

    spare = [H.contacts_for(b) for b, _, _ in state]      # ping-pong output buffers: no allocation in the timed loop

    def pass_(n):
        from oracle import binding as oracle
        def work(i):
            b, p, w = state[i]
            oracle.step(shapes, b, p, w, spare[i], n)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(state))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        for i in range(len(state)):
            b, p, w = state[i]
            state[i], spare[i] = (b, p, spare[i]), w
        return dt

    for _ in range(args.warmup):
        pass_(1)
    times = [pass_(1) for _ in range(args.steps)]
    total = float(sum(times))
    body_steps = sample_worlds * BODIES_PER_WORLD * args.steps
    value = body_steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "ms_per_step_median": 1e3 * float(np.median(times)),
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "worlds_total": sample_worlds, "bodies_per_world": BODIES_PER_WORLD,
                   "preroll_steps": args.preroll},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_worlds} worlds x {args.steps} steps after {args.preroll} pre-roll steps, "
                                   f"{cores} threads (C++ restatement of the Elm path; elm/node absent)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Keep the process's stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner comes from
    C code, whatever NCCL_DEBUG says on the box) goes to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def e2e_contexts_for(worlds_per_gpu, requested=0):
    """Sub-batches (contexts + host threads) one rank drives its shard with in the end-to-end leg: `requested` if given, else two
    above 4096 worlds per GPU and one below (profiles/r02_e2e_contexts_sweep.txt); never more than the shard has worlds."""
    k = requested if requested > 0 else (2 if worlds_per_gpu > 4096 else 1)
    return max(1, min(k, worlds_per_gpu))


def host_waits_spin(worlds_total, world_size, local_ranks, cpu_count):
    """Spin in the host waits of a frame when every driving thread of the box can have four cores to itself."""
    per_gpu = -(-worlds_total // max(world_size, 1))
    return 4 * local_ranks * e2e_contexts_for(per_gpu) <= cpu_count


def bind_to_gpu_numa_node(device_index):
    """Run this rank (and the pinned buffers it allocates: first touch) on the CPUs next to its GPU, so that the host<->device
    copies of an 8-GPU box do not cross the socket interconnect.  Best effort; EP_BENCH_NO_BIND=1 disables."""
    if os.environ.get("EP_BENCH_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:  # noqa: BLE001
        return None


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=8192, help="worlds of the whole batch (sharded over the GPUs)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling extra (--worlds per GPU)")
    ap.add_argument("--no-cars", action="store_true", help="skip the C5 block (65536 car worlds + 4 rays per world per step)")
    ap.add_argument("--car-worlds", type=int, default=65536)
    ap.add_argument("--car-steps", type=int, default=20)
    ap.add_argument("--preroll", type=int, default=120, help="untimed steps that drop and pile the bodies first")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-contexts", type=int, default=0,
                    help="sub-batches (contexts + host threads) of the end-to-end measurement; 0 = by shard size (2 above 4096 worlds per GPU, else 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pile", action="store_true", help="skip the second half of BASELINE's metric (ms/step of the 2k-body convex pile)")
    ap.add_argument("--pile-preroll", type=int, default=600)
    ap.add_argument("--pile-steps", type=int, default=60)
    args = ap.parse_args()
    rank, world_size, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world_size)
        return

    import torch
    import torch.distributed as dist
    from elm_physics_b200 import abi, engine as eng, pack, scenes, sharding
    import helpers as H

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    # How the host waits for a frame (EP_BLOCKING_SYNC, read once by the library): sleeping on an event costs a wake-up of 30-50 us
    # per wait, three waits per frame -- 10 % of a 1024-world frame (26.8 -> 29.5 M body-steps/s end to end, gpurun_out/s4) and
    # nothing at 8192 worlds.  Spin when every driving thread of the box can have four cores to itself; otherwise (many ranks on
    # few cores) keep the library's default, the sleeping wait that the 8-GPU box of round 1 needed.
    local_ranks = env_int("LOCAL_WORLD_SIZE", world_size)
    if "EP_BLOCKING_SYNC" not in os.environ and host_waits_spin(args.worlds, world_size, local_ranks, os.cpu_count() or 1):
        os.environ["EP_BLOCKING_SYNC"] = "0"
    host_wait = "spin" if os.environ.get("EP_BLOCKING_SYNC") == "0" else "sleep"
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    WT = args.worlds                                                         # worlds of the whole batch
    w0, w1 = sharding.shard_range(WT, rank, world_size)                      # shard = contiguous world-index range
    W = w1 - w0                                                              # worlds of this rank
    shapes, bodies, params = scenes.mixed_worlds(W, first_world=w0)
    e = eng.Engine(local_rank)
    e.upload_shapes(shapes)
    e.upload_worlds(bodies, params, None)
    e.step_resident(args.preroll)
    e.sync()
    for _ in range(args.warmup):
        e.step_resident(1)
    e.sync()
    # snapshot of the state the timed region starts from (for the CPU baseline + parity check)
    start_bodies = bodies.copy()
    e.download_bodies(start_bodies)
    start_contacts = H.contacts_for(bodies)
    e.download_contacts(start_contacts)
    counts0 = e.stats()
    work0 = (counts0["contact_iterations"], counts0["joint_row_iterations"])

    stream = torch.cuda.ExternalStream(e.stream_handle(), device=torch.device("cuda", local_rank))
    # Timing rule: inputs larger than L2 or an L2 flush between timed iterations.  At 8192 worlds per GPU a step streams 2.7 GB
    # through HBM (no flush needed); a shard of a few thousand worlds or fewer would sit in the 126 MB L2 from step to step,
    # so its steps are timed one by one with a 256 MB write between them (outside the event brackets).
    flush_l2 = W < 4096 and not os.environ.get("EP_BENCH_NO_FLUSH")
    sampler = ClockSampler(local_rank)
    launches0 = e.stats()["kernel_launches"]
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    if flush_l2:
        scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        with torch.cuda.stream(stream):
            for a, b in evs:
                scratch.zero_()
                a.record(stream)
                e.step_resident(1)
                b.record(stream)
        e.sync()
        torch.cuda.synchronize()
        my_ms = sum(a.elapsed_time(b) for a, b in evs)
        del scratch
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        e.step_resident(args.steps)
        ev1.record(stream)
        e.sync()
        torch.cuda.synchronize()
        my_ms = ev0.elapsed_time(ev1)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(my_ms)
    launches = e.stats()["kernel_launches"] - launches0
    end_bodies = bodies.copy()
    e.download_bodies(end_bodies)
    end_contacts = H.contacts_for(bodies)
    e.download_contacts(end_contacts)
    counts = e.stats()

    # the only collective of the path: the final gather of the integrated state -- ncclAllGather of the 22 output doubles per
    # body, packed on the device (ep_pack_state_device) straight into the send buffer: no host staging.  Device-timed.
    gather = None
    if world_size > 1:
        nb_max = int(max_over_ranks(float(bodies.n_bodies)))
        send = recv = None
        cur = torch.cuda.current_stream()
        times = []
        for it in range(4):                      # the first pass creates the communicator's channels: untimed
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(cur)
            recv, send = sharding.gather_state_device(e, nb_max, send=send, recv=None if recv is None else recv.view(-1, 22))
            g1.record(cur)
            torch.cuda.synchronize()
            if it:
                times.append(max_over_ranks(g0.elapsed_time(g1)))
        got = recv[rank, :bodies.n_bodies].cpu().numpy()
        ok = got[:, 0:3].tobytes() == end_bodies.pos.tobytes() and got[:, 13:22].tobytes() == end_bodies.inv_inertia_world.reshape(-1, 9).tobytes()
        gbytes = world_size * nb_max * 22 * 8
        gather = {"ms": float(np.median(times)), "bytes_gathered_per_rank": int(gbytes), "GBps_per_rank": gbytes / (np.median(times) * 1e-3) / 1e9,
                  "how": "ep_pack_state_device + ncclAllGather (torch.distributed.all_gather_into_tensor), device to device, median of 3",
                  "matches_download": bool(ok)}
        del send, recv

    n_dyn = W * BODIES_PER_WORLD
    n_dyn_all = WT * BODIES_PER_WORLD
    body_steps_all = n_dyn_all * args.steps
    value = body_steps_all / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call, one frame per call --------------------------------------
    # ep_step_frame = what a `simulate` loop does per frame: H2D of every per-body array from pinned host memory, one
    # step warm-started by the previous frame's (resident) contacts, D2H of the integrated state + the frame's contacts.
    pool = eng.PinnedPool()
    eb = end_bodies.copy(alloc=pool.empty)
    cap = H.contacts_for(eb)
    eo = abi.Contacts(eb.n_worlds, cap.group_cap, cap.contact_cap, alloc=pool.empty)
    del cap
    h2d = d2h = 0
    barrier()
    e.step_frame(eb, eo)                  # untimed warm-up of the host path
    t_e2e = 0.0
    for _ in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        e.step_frame(eb, eo)
        t_e2e += max_over_ranks(time.perf_counter() - t0)
        h2d = sum(getattr(eb, n).nbytes for n, _ in abi.BODY_F64_FIELDS)
        d2h = eb.nbytes_out() + eo.nbytes_out()
    e2e_single = n_dyn_all * args.e2e_steps / t_e2e
    e2e_parity = None
    if world_size == 1 and not args.no_cpu_baseline:
        # the frames above continue the timed run: check the host-buffer path against the oracle over the same frames
        sample = min(W, 32)
        ob = end_bodies.slice_worlds(0, sample)
        owarm = pack.contacts_from_world_dicts([pack.contacts_world_dict(end_contacts, k) for k in range(sample)])
        oout = H.contacts_for(ob)
        from oracle import binding as oracle_binding
        oracle_binding.step(shapes, ob, abi.slice_params(params, 0, sample), owarm, oout, args.e2e_steps + 1)
        r1 = int(bodies.world_body_off[sample])
        e2e_parity = all(getattr(eb, n)[:r1].tobytes() == getattr(ob, n).tobytes() for n in abi.BODY_INOUT)

    # The same call on K contexts of W/K worlds each, one host thread per context (ctypes releases the GIL): the H2D of one
    # sub-batch, the step of another and the D2H of a third overlap (PCIe is full duplex, two copy engines), which is how a
    # host that owns several independent world batches would drive one GPU.  This is the headline e2e.
    # Measured (gpurun_out/s4/ksweep.txt, B200): a shard of 1024 / 4096 worlds runs 28.7 / 52.8 M body-steps/s as ONE context against
    # 26.8 / 49.7 M as two (both halves sit on the step's latency floor, so splitting doubles it); at 8192 worlds two contexts win (77 vs 67 M).
    K = e2e_contexts_for(W, args.e2e_contexts)

    def pipelined(fields, up=None, down=None):
        subs = []
        for k in range(K):
            w0k, w1k = sharding.shard_range(W, k, K)
            bk = end_bodies.slice_worlds(w0k, w1k).copy(alloc=pool.empty)
            pk = abi.slice_params(params, w0k, w1k)
            ek = eng.Engine(local_rank)
            ek.upload_shapes(shapes)
            ek.upload_worlds(bk, pk, None)
            capk = H.contacts_for(bk)
            ok = abi.Contacts(bk.n_worlds, capk.group_cap, capk.contact_cap, alloc=pool.empty, fields=fields)
            del capk
            for _ in range(3):
                ek.step_frame(bk, ok, up, down)         # untimed: warm start builds up, host path warms
            subs.append((ek, bk, ok))

        def frames(sub):
            ek, bk, ok = sub
            for _ in range(args.e2e_steps):
                ek.step_frame(bk, ok, up, down)

        barrier()
        ths = [threading.Thread(target=frames, args=(sub,)) for sub in subs]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        t_pipe = max_over_ranks(time.perf_counter() - t0)
        um, dm = (abi.EP_F_ALL if up is None else up), (abi.EP_F_ALL if down is None else down)
        h2d = sum(sum(getattr(bk, n).nbytes for n, _ in abi.BODY_F64_FIELDS if abi.BODY_FIELD_MASK[n] & um) for _, bk, _ in subs)
        d2h = sum(sum(getattr(bk, n).nbytes for n in abi.BODY_INOUT if abi.BODY_FIELD_MASK[n] & dm) + ok.nbytes_out() for _, bk, ok in subs)
        for ek, _, _ in subs:
            ek.close()
        return n_dyn_all * args.e2e_steps / t_pipe, sum_over_ranks(h2d), sum_over_ranks(d2h)

    # headline: the host reads back what the public API exposes of a frame -- the integrated bodies and what
    # Physics.contactPoints consumes (src/Physics.elm:1158-1207); the rest of `Contacts` is the next frame's warm start
    # and stays resident.  Beside it: the same with EVERY array of ep_contacts copied back.
    # Per-body arrays: the dynamic state goes up and comes back (EP_F_STATE: transform, velocities -- 104 B per body each way);
    # constants stay where ep_upload_worlds put them, forces / torques travel only on frames that applied some (none here) and
    # are zero after a step, invInertiaWorld follows from the returned orientation (ep_step_frame_fields).
    e2e_value, h2d, d2h = pipelined(abi.CONTACT_POINTS_FIELDS, abi.EP_F_STATE, abi.EP_F_STATE)
    e2e_full, h2d_full, d2h_full = pipelined(None)

    # the same K steps once more with a CUDA-event pair around every launch (ep_profile): the per-kernel durations behind
    # `roofline` / `kernels`.  Kept out of the pass above because ~60 event records per step cost ~3 % of it.
    counts_p0 = e.stats()
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.profile(True)
    pv0.record(stream)
    e.step_resident(args.steps)
    pv1.record(stream)
    e.sync()
    torch.cuda.synchronize()
    ms_profiled = pv0.elapsed_time(pv1)
    ktimes = e.kernel_times()
    e.profile(False)
    counts_prof = e.stats()

    # ---- N > 1: the weak-scaling figure beside the headline (every GPU gets a whole 8192-world batch of its own) -----------
    weak = None
    if world_size > 1 and not args.no_weak:
        ws, wb, wp = scenes.mixed_worlds(WT, first_world=rank * WT)
        ew = eng.Engine(local_rank)
        ew.upload_shapes(ws)
        ew.upload_worlds(wb, wp, None)
        ew.step_resident(args.preroll + args.warmup)
        ew.sync()
        wstream = torch.cuda.ExternalStream(ew.stream_handle(), device=torch.device("cuda", local_rank))
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        x0.record(wstream)
        ew.step_resident(args.steps)
        x1.record(wstream)
        ew.sync()
        torch.cuda.synchronize()
        wms = max_over_ranks(x0.elapsed_time(x1))
        weak = {"scaling": "weak", "worlds_per_gpu": WT, "worlds_total": WT * world_size, "ms_per_step": wms / args.steps,
                "value": WT * world_size * BODIES_PER_WORLD * args.steps / (wms * 1e-3), "unit": UNIT}
        ew.close()
        del ws, wb, wp

    # ---- BASELINE config 5: car worlds (chassis + 4 hinged wheels + ramp + 6 hollow crates + tether) sharded like the headline,
    # with the RaycastCar pattern between the steps: 4 rays per car from the chassis, placed on the device (ep_raycast_attached),
    # hits copied back to host buffers every step.  Wall clock around the whole loop (the ray results are a host read).
    cars = None
    if not args.no_cars:
        CT = args.car_worlds
        c0w, c1w = sharding.shard_range(CT, rank, world_size)
        cs, cb, cp = scenes.car_batch(CT)
        cb_r, cp_r, _ = scenes.slice_batch(cb, cp, range(c0w, c1w)) if world_size > 1 else (cb, cp, None)
        del cb, cp
        ec = eng.Engine(local_rank)
        ec.upload_shapes(cs)
        ec.upload_worlds(cb_r, cp_r, None)
        nwc = cb_r.n_worlds
        chassis = (cb_r.world_body_off[:nwc] + 2).astype(np.int32)
        corners = np.array([[-1.0, 1.9, -0.25], [1.0, 1.9, -0.25], [-1.0, -1.9, -0.25], [1.0, -1.9, -0.25]])
        # the host side of the casts on page-locked memory, like the frame buffers of the e2e leg (pageable arrays go through the
        # driver's staging buffer at a fifth of the PCIe rate: 3.1 ms per 262144 rays)
        ray_body, lf, ld = pool.empty(4 * nwc, np.int32), pool.empty((4 * nwc, 3), np.float64), pool.empty((4 * nwc, 3), np.float64)
        ray_body[:] = np.repeat(chassis, 4)
        lf[:] = np.tile(corners, (nwc, 1))
        ld[:] = np.array([0.0, 0.0, -1.0])
        ec.ray_alloc = pool.empty
        for _ in range(30):
            ec.step_resident(1)
            ec.raycast_attached(ray_body, lf, ld)
        ec.sync()
        dyn_per_world = int((cb_r.kind[:int(cb_r.world_body_off[1])] == abi.EP_DYNAMIC).sum())
        barrier()
        t0 = time.perf_counter()
        hits = 0
        for _ in range(args.car_steps):
            ec.step_resident(1)
            hit = ec.raycast_attached(ray_body, lf, ld)[0]
            hits += int((hit >= 0).sum())
        ec.sync()
        tc = max_over_ranks(time.perf_counter() - t0)
        # device time of the steps alone (no rays), same state
        cstream = torch.cuda.ExternalStream(ec.stream_handle(), device=torch.device("cuda", local_rank))
        y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y0.record(cstream)
        ec.step_resident(args.car_steps)
        y1.record(cstream)
        ec.sync()
        torch.cuda.synchronize()
        cms = max_over_ranks(y0.elapsed_time(y1))
        cars = {"workload": "C5: car worlds (chassis compound + 4 hinged wheels + ramp + 6 hollow crates + tether; 5 constraints, 21 joint rows), "
                            "4 attached rays per world per step, sharded by world index",
                "worlds_total": CT, "worlds_per_gpu": CT // world_size, "dynamic_bodies_per_world": dyn_per_world, "steps": args.car_steps,
                "body_steps_per_sec_with_rays": CT * dyn_per_world * args.car_steps / tc,
                "rays_per_sec": 4.0 * CT * args.car_steps / tc,
                "ms_per_step_with_rays": 1e3 * tc / args.car_steps,
                "body_steps_per_sec_steps_only_device_time": CT * dyn_per_world * args.car_steps / (cms * 1e-3),
                "ms_per_step_steps_only": cms / args.car_steps,
                "ray_hit_fraction": sum_over_ranks(hits) / (4.0 * CT * args.car_steps),
                "ray_io": "ep_raycast_attached: pinned host buffers in (13.6 MB per 262144 rays) and out (15.7 MB), synchronous, every step"}
        ec.close()
        del cs, cb_r, cp_r

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "worlds_total": WT, "worlds_per_gpu": WT // world_size, "bodies_per_world": BODIES_PER_WORLD,
                   "preroll_steps": args.preroll,
                   "sharding": "contiguous world-index range per rank, no data-path collective; one ncclAllGather of the final state",
                   "l2": ("flushed between timed steps (256 MB write outside the per-step event brackets): the shard's state would fit the 126 MB L2"
                          if flush_l2 else "no flush: a step streams ~0.33 MB per world through HBM (2.7 GB at 8192 worlds, profiles/), far more than the 126 MB L2")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": args.e2e_steps,
                "call": "ep_step_frame_fields(pinned host buffers): H2D of the bodies' dynamic state (transform + velocities; pos / quat "
                        "ahead of the front kernels, the rest beside them), one step warm-started by the resident previous frame, D2H "
                        "of the integrated state + the contact data Physics.contactPoints reads (pair ids, pre-step body1 frames, "
                        "contact points; marshalled and copied beside the solver); each rank drives its shard as %d context(s) from %d "
                        "host thread(s) (two above 4096 worlds per GPU, so that copies and steps of different sub-batches overlap; one "
                        "below, where a half shard would sit on the same latency floor as the whole); host waits: %s" % (K, K, host_wait),
                "host_wait": host_wait,
                "contexts": K, "every_array_value": e2e_full, "every_array_h2d_bytes_per_step": int(h2d_full),
                "every_array_d2h_bytes_per_step": int(d2h_full),
                "single_context_every_array_value": e2e_single, "bit_exact_vs_oracle_32_worlds": e2e_parity},
        "gpu_launches": int(launches), "ms_per_step_with_per_kernel_events": ms_profiled / args.steps,
        "clocks": clocks,
        "final_state_gather_ms": gather["ms"] if gather else None, "final_state_gather": gather,
        "weak_scaling": weak, "cars": cars,
    }
    if rank == 0:
        # ---- roofline of the dominant kernel (CUDA events around every launch of the timed region) ------------
        # the solver's per-size-class launches run concurrently inside the k_solve bracket: reported, not summed
        sub = {k: v for k, v in ktimes.items() if k.startswith("k_solve_")}
        ktimes = {k: v for k, v in ktimes.items() if not k.startswith("k_solve_") and v[1] > 0}
        dom = max(ktimes, key=lambda k: ktimes[k][0])
        total_k = sum(v[0] for v in ktimes.values())
        c = {"bodies": bodies.n_bodies, "dynamic_bodies": n_dyn, "contacts": (counts0["contacts"] + counts["contacts"]) / 2,
             "candidate_pairs": (counts0["candidate_pairs"] + counts["candidate_pairs"]) / 2,
             "groups": (counts0["groups"] + counts["groups"]) / 2,
             "placed_f64_per_body": float(shapes.f64_off[-1] - shapes.f64_off[1]) / 4.0}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dur_ms = ktimes[dom][0] / max(ktimes[dom][1], 1)
        ach = kernel_bytes(dom, c) / (dur_ms * 1e-3) / 1e9
        hbm_view = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": (TRAFFIC_NCU.get(dom) * W / 8192.0) if TRAFFIC_NCU.get(dom) else None,
                    "traffic_source": "profiles/r02_ncu_traffic_per_launch.json (ncu --set full at 8192 worlds), scaled by worlds per GPU / 8192",
                    "peak_source": "measured" if peaks else "fallback",
                    "avg_launch_ms": dur_ms, "share_of_step": ktimes[dom][0] / total_k if total_k else None}
        line["roofline"] = hbm_view
        # The solver is not an HBM kernel (each row crosses HBM once per step): its roof is the FP64 pipe without FMA
        # (tools/fp64_peak.cu: profiles/r01_fp64_peak.json) and, in practice, the dependent-issue latency of the PGS chain.
        # Algorithmic fp64 ops: 270 per contact-iteration, 90 per joint-row-iteration, counted on the device.
        ci = counts_prof["contact_iterations"] - counts_p0["contact_iterations"]        # of the profiled pass, like ktimes
        ji = counts_prof["joint_row_iterations"] - counts_p0["joint_row_iterations"]
        fp64_peak = 18.52
        try:
            fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))["dadd_tops"])
        except Exception:  # noqa: BLE001
            pass
        if "k_solve" in ktimes and ktimes["k_solve"][0] > 0:
            tops = (270.0 * ci + 90.0 * ji) / (ktimes["k_solve"][0] * 1e-3) / 1e12
            fp64_view = {"bound": "fp64", "kernel": "k_solve", "achieved": tops, "peak": fp64_peak, "unit": "T fp64 op/s (no FMA)",
                         "frac": tops / fp64_peak, "contact_iterations_per_step": ci / args.steps,
                         "mean_iterations_per_contact": ci / max(1.0, args.steps * 0.5 * (counts_p0["contacts"] + counts_prof["contacts"])),
                         "traffic": (TRAFFIC_NCU.get("k_solve") * W / 8192.0) if TRAFFIC_NCU.get("k_solve") else None, "avg_launch_ms": ktimes["k_solve"][0] / max(ktimes["k_solve"][1], 1),
                         "share_of_step": ktimes["k_solve"][0] / total_k if total_k else None,
                         "peak_source": "tools/fp64_peak.cu measured on this pool's B200 (profiles/r01_fp64_peak.json)"}
            line["roofline_fp64"] = fp64_view
            if dom == "k_solve":       # the dominant kernel's roof is the FP64 pipe: that is the headline roofline; the HBM view beside it
                line["roofline"] = fp64_view
                line["roofline_hbm"] = hbm_view
        line["kernels"] = {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1], "share": v[0] / total_k if total_k else None,
                               "algorithmic_GBps": kernel_bytes(k, c) / (max(v[0], 1e-9) / max(v[1], 1) * 1e-3) / 1e9}
                           for k, v in ktimes.items()}
        line["solve_classes_ms"] = {k: v[0] / max(v[1], 1) for k, v in sub.items() if v[1]}
        line["workload_counts"] = {k: c[k] for k in ("bodies", "contacts", "candidate_pairs", "groups")}
        line["workload_counts"]["islands"] = counts["islands"]
        # ---- CPU baseline on a bounded sample of the same timed steps + parity of the timed run (N=1 only) -----
        if world_size == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            sample = min(W, max(cores * 8, 64))
            sb = start_bodies.slice_worlds(0, sample)
            sp = abi.slice_params(params, 0, sample)
            sw = pack.contacts_from_world_dicts([pack.contacts_world_dict(start_contacts, k) for k in range(sample)])
            secs, chunks, cuts = oracle_run(shapes, sb, sp, sw, args.steps, cores)
            cpu_value = sample * BODIES_PER_WORLD * args.steps / secs
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"worlds [0,{sample}) of the timed state x {args.steps} steps, {cores} threads"}
            exact = True
            for i, ch in enumerate(chunks):
                r0, r1 = int(bodies.world_body_off[cuts[i]]), int(bodies.world_body_off[cuts[i + 1]])
                for name in abi.BODY_INOUT:
                    exact = exact and getattr(end_bodies, name)[r0:r1].tobytes() == getattr(ch[0], name).tobytes()
            line["parity"] = {"worlds": sample, "steps": args.steps, "bit_exact_vs_oracle": bool(exact)}
        # ---- second half of BASELINE's metric: ms/step of the 2k-body convex pile (config 4; one world, one GPU), and config 2
        # (100 boxes: the 91-box pyramid + a 9-box stack, one world) beside it.  Both are ONE giant island: the thread-block-cluster
        # solver's case (ep_solve_cluster.cuh).  The restatement's single-thread ms/step from the same settled state rides along.
        if world_size == 1 and not args.no_pile:
            def single_world(name, worlds, preroll, steps, cpu_steps):
                ps, pb, pp = pack.pack_worlds(worlds)
                ep = eng.Engine(local_rank)
                ep.upload_shapes(ps)
                ep.upload_worlds(pb, pp, None)
                ep.step_resident(preroll)
                ep.sync()
                cpu_ms = None
                if not args.no_cpu_baseline and cpu_steps > 0:      # the checker, timed from the GPU's settled state (bodies + warm-start contacts)
                    from oracle import binding as oracle_binding
                    sb = pb.copy()
                    ep.download_bodies(sb)
                    sw = abi.Contacts(1, 40 * pb.n_bodies, 64 * pb.n_bodies)
                    ep.download_contacts(sw)
                    so = abi.Contacts(1, 40 * pb.n_bodies, 64 * pb.n_bodies)
                    t0 = time.perf_counter()
                    oracle_binding.step(ps, sb, pp, sw, so, cpu_steps)
                    cpu_ms = (time.perf_counter() - t0) * 1e3 / cpu_steps
                pstream = torch.cuda.ExternalStream(ep.stream_handle(), device=torch.device("cuda", local_rank))
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record(pstream)
                ep.step_resident(steps)
                q1.record(pstream)
                ep.sync()
                torch.cuda.synchronize()
                pst = ep.stats()
                out = {"workload": name, "ms_per_step": q0.elapsed_time(q1) / steps, "steps": steps, "preroll_steps": preroll,
                       "bodies": int(pb.n_bodies), "contacts": pst["contacts"], "islands": pst["islands"], "unit": "ms/step",
                       "higher_is_better": False, "cpu_restatement_ms_per_step_1_thread": cpu_ms}
                ep.close()
                return out
            line["pile"] = single_world("C4: 2000 convex hulls / compounds dropped into a walled pit, one world, 20 iterations",
                                        scenes.pile_scene(2000), args.pile_preroll, args.pile_steps, 5)
            line["boxes"] = single_world("C2: 100 boxes (91-box pyramid + 9-box stack) on a plane, one world, 20 iterations",
                                         scenes.boxes_scene(), 300, args.pile_steps, 30)
        emit(line)
    pool.close()
    e.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
