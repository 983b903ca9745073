"""ctypes binding of the product library `csrc/libelm_physics_b200.so` (C ABI: include/elm_physics_b200.h)
plus the `simulate` glue that mirrors `Physics.simulate` (reference: src/Physics.elm:522-583).

There is NO CPU path here: if the CUDA library is missing or no GPU is usable, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libelm_physics_b200.so")
_LIB = None

EXPORTS = ("ep_create", "ep_destroy", "ep_last_error", "ep_upload_shapes", "ep_upload_worlds", "ep_step_resident",
           "ep_sync", "ep_download_bodies", "ep_download_contacts", "ep_step", "ep_stats", "ep_stream",
           "ep_profile", "ep_kernel_times", "ep_kernel_name", "ep_configure", "ep_step_frame", "ep_host_alloc",
           "ep_host_free", "ep_raycast", "ep_apply", "ep_raycast_attached", "ep_step_frame_fields", "ep_pack_state_device")


class EngineError(RuntimeError):
    pass


def load_library():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.ep_last_error.restype = C.c_char_p
        lib.ep_stream.restype = C.c_void_p
        lib.ep_kernel_name.restype = C.c_char_p
        lib.ep_destroy.restype = None
        lib.ep_host_alloc.restype = C.c_void_p
        lib.ep_host_alloc.argtypes = [C.c_size_t]
        lib.ep_host_free.restype = None
        lib.ep_host_free.argtypes = [C.c_void_p]
        for name in EXPORTS:
            getattr(lib, name)
        _LIB = lib
    return _LIB


class PinnedPool:
    """numpy arrays on page-locked host memory (ep_host_alloc); pass `pool.empty` as the `alloc` argument of
    abi.Bodies / abi.Contacts so that ep_step_frame copies at full PCIe rate."""

    def __init__(self):
        self.lib = load_library()
        self._ptrs = []

    def empty(self, shape, dtype):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        p = self.lib.ep_host_alloc(max(nbytes, 1))
        if not p:
            raise EngineError(f"ep_host_alloc({nbytes}) failed")
        self._ptrs.append(p)
        buf = (C.c_char * max(nbytes, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def close(self):
        for p in self._ptrs:
            self.lib.ep_host_free(p)
        self._ptrs = []


class Engine:
    """One ep_ctx on one CUDA device."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.ep_create(C.byref(self.ctx), C.c_int(device))
        if rc != 0:
            raise EngineError(f"ep_create(device={device}) failed with {rc}: no usable CUDA device (no CPU fallback)")
        self._keep = []

    def close(self):
        if self.ctx:
            self.lib.ep_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed with {rc}: {self.lib.ep_last_error(self.ctx).decode()}")

    def configure(self, contact_cap=None, item_cap=None):
        """ep_configure, read by the next upload_worlds: `contact_cap` = device contact pool (contacts per step over all
        worlds), `item_cap` = narrowphase work items (shape pairs of candidate body pairs) per step."""
        if contact_cap is not None:
            self._ck(self.lib.ep_configure(self.ctx, abi.EP_OPT_CONTACT_CAP, C.c_int64(contact_cap)), "ep_configure")
        if item_cap is not None:
            self._ck(self.lib.ep_configure(self.ctx, abi.EP_OPT_ITEM_CAP, C.c_int64(item_cap)), "ep_configure")

    def upload_shapes(self, shapes):
        self._ck(self.lib.ep_upload_shapes(self.ctx, C.byref(shapes.c)), "ep_upload_shapes")

    def upload_worlds(self, bodies, params, warm=None):
        self._ck(self.lib.ep_upload_worlds(self.ctx, C.byref(bodies.c), C.byref(params.c),
                                           C.byref(warm.c) if warm is not None else None), "ep_upload_worlds")

    def step_resident(self, n_steps=1):
        self._ck(self.lib.ep_step_resident(self.ctx, C.c_int32(n_steps)), "ep_step_resident")

    def sync(self):
        self._ck(self.lib.ep_sync(self.ctx), "ep_sync")

    def download_bodies(self, bodies):
        self._ck(self.lib.ep_download_bodies(self.ctx, C.byref(bodies.c)), "ep_download_bodies")

    def download_contacts(self, contacts):
        self._ck(self.lib.ep_download_contacts(self.ctx, C.byref(contacts.c)), "ep_download_contacts")

    def step(self, bodies, params, warm=None, out=None, n_steps=1):
        self._ck(self.lib.ep_step(self.ctx, C.byref(bodies.c), C.byref(params.c),
                                  C.byref(warm.c) if warm is not None else None,
                                  C.byref(out.c) if out is not None else None, C.c_int32(n_steps)), "ep_step")

    def step_frame(self, bodies, out=None, upload=None, download=None):
        """ep_step_frame / ep_step_frame_fields: one frame of the resident batch through host buffers (state in, state +
        contacts out).  `upload` / `download`: abi.EP_F_* masks of the per-body arrays to move (None = every array)."""
        if upload is None and download is None:
            self._ck(self.lib.ep_step_frame(self.ctx, C.byref(bodies.c), C.byref(out.c) if out is not None else None), "ep_step_frame")
            return
        up = abi.EP_F_ALL if upload is None else upload
        down = abi.EP_F_ALL if download is None else download
        self._ck(self.lib.ep_step_frame_fields(self.ctx, C.byref(bodies.c), C.byref(out.c) if out is not None else None,
                                               C.c_uint32(up), C.c_uint32(down)), "ep_step_frame_fields")

    def pack_state_device(self, dst_ptr, stream=None):
        """ep_pack_state_device: the 22 output doubles per body into caller-owned DEVICE memory ([n_bodies][22] f64), on `stream`
        (a cudaStream_t handle; None = the context's): the send buffer of the final NCCL state gather."""
        self._ck(self.lib.ep_pack_state_device(self.ctx, C.c_void_p(dst_ptr), C.c_void_p(stream) if stream else None), "ep_pack_state_device")

    def raycast(self, ray_world, origins, directions):
        """ep_raycast (Physics.raycast, src/Physics.elm:802-841) over the resident worlds.  Returns
        (hit_body rows or -1, distance, point (n,3), normalised normal (n,3))."""
        rw = np.ascontiguousarray(ray_world, dtype=np.int32)
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        dr = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        n = len(rw)
        hit, dist, point, normal = np.zeros(n, np.int32), np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
        pf, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._ck(self.lib.ep_raycast(self.ctx, C.c_int32(n), rw.ctypes.data_as(pi), o.ctypes.data_as(pf), dr.ctypes.data_as(pf),
                                     hit.ctypes.data_as(pi), dist.ctypes.data_as(pf), point.ctypes.data_as(pf),
                                     normal.ctypes.data_as(pf)), "ep_raycast")
        return hit, dist, point, normal

    def raycast_attached(self, ray_body, local_origins, local_directions):
        """ep_raycast_attached: rays given in the centre-of-mass frame of body rows `ray_body`, placed with the bodies' current
        transforms on the device, cast against the other bodies of each body's world (the RaycastCar pattern)."""
        rb = np.ascontiguousarray(ray_body, dtype=np.int32)
        o = np.ascontiguousarray(local_origins, dtype=np.float64).reshape(-1, 3)
        dr = np.ascontiguousarray(local_directions, dtype=np.float64).reshape(-1, 3)
        n = len(rb)
        if not hasattr(self, "_ray_out") or len(self._ray_out[0]) != n:
            # result buffers of the casts, reused from call to call; `ray_alloc` (e.g. PinnedPool.empty) puts them on page-locked
            # memory: from pageable arrays the 30 MB of a 262144-ray cast travel at ~10 GB/s through the driver's staging buffer
            alloc = getattr(self, "ray_alloc", None) or (lambda shape, dtype: np.zeros(shape, dtype))
            self._ray_out = (alloc(n, np.int32), alloc(n, np.float64), alloc((n, 3), np.float64), alloc((n, 3), np.float64))
        hit, dist, point, normal = self._ray_out
        pf, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._ck(self.lib.ep_raycast_attached(self.ctx, C.c_int32(n), rb.ctypes.data_as(pi), o.ctypes.data_as(pf), dr.ctypes.data_as(pf),
                                              hit.ctypes.data_as(pi), dist.ctypes.data_as(pf), point.ctypes.data_as(pf),
                                              normal.ctypes.data_as(pf)), "ep_raycast_attached")
        return hit, dist, point, normal

    def apply(self, kind, body_row, a, b):
        """ep_apply: body mutators on the resident state (src/Physics.elm:857-1017), operations in list order."""
        k = np.ascontiguousarray(kind, dtype=np.int32)
        r = np.ascontiguousarray(body_row, dtype=np.int32)
        va = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
        vb = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 3)
        pf, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._ck(self.lib.ep_apply(self.ctx, C.c_int32(len(k)), k.ctypes.data_as(pi), r.ctypes.data_as(pi),
                                   va.ctypes.data_as(pf), vb.ctypes.data_as(pf)), "ep_apply")

    def stats(self):
        out = np.zeros(abi.EP_STAT_COUNT, np.int64)
        self._ck(self.lib.ep_stats(self.ctx, out.ctypes.data_as(C.POINTER(C.c_int64)), abi.EP_STAT_COUNT), "ep_stats")
        return dict(kernel_launches=int(out[0]), candidate_pairs=int(out[1]), groups=int(out[2]),
                    contacts=int(out[3]), islands=int(out[4]), contact_iterations=int(out[5]),
                    joint_row_iterations=int(out[6]), resums=int(out[7]))

    def profile(self, enable=True):
        self._ck(self.lib.ep_profile(self.ctx, C.c_int(1 if enable else 0)), "ep_profile")

    def kernel_times(self):
        """{kernel name: (accumulated ms, launches)} since the last profile() call."""
        n = abi.EP_K_COUNT
        ms, cnt = np.zeros(n, np.float64), np.zeros(n, np.int64)
        self._ck(self.lib.ep_kernel_times(self.ctx, ms.ctypes.data_as(C.POINTER(C.c_double)),
                                          cnt.ctypes.data_as(C.POINTER(C.c_int64)), n), "ep_kernel_times")
        return {self.lib.ep_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def stream_handle(self):
        return self.lib.ep_stream(self.ctx)

    # ---- Physics.simulate (src/Physics.elm:522-583) for one world ---------------------------------
    def simulate(self, config, bodies_with_ids):
        from . import pack
        from . import physics as P
        with_ids, _max_id = P.assign_ids(list(bodies_with_ids))
        if not with_ids:
            return [], P.Contacts(packed=None, iterations=0, bodies=[])
        shapes, bodies, params = pack.pack_worlds([(config, with_ids)])
        warm = None
        if config.contacts is not None and config.contacts.packed is not None:
            warm = pack.contacts_from_world_dicts([config.contacts.packed])
        n = bodies.n_bodies
        out = abi.Contacts(1, n * (n - 1) // 2 + 1, max(1 << 14, 16 * n))      # the device pool's default size (ep_upload_worlds)
        self.upload_shapes(shapes)
        self.step(bodies, params, warm, out, 1)
        new_bodies = pack.unpack_bodies(bodies, [(config, with_ids)])[0]
        packed = pack.contacts_world_dict(out, 0)
        return new_bodies, P.Contacts(packed=packed, iterations=packed["iterations"], bodies=new_bodies)


_DEFAULT = None


def default_engine():
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Engine(0)
    return _DEFAULT
