"""Shared test helpers: bit-exact comparison of packed state / contacts, randomized pair scenes."""
import numpy as np

from elm_physics_b200 import abi
from elm_physics_b200 import pack
from elm_physics_b200 import physics as P


def contacts_for(bodies, per_world_pairs=None):
    """Output buffer big enough for any frame of `bodies` (worst case all pairs, 24 contacts each is never reached)."""
    nw = bodies.n_worlds
    n = np.diff(bodies.world_body_off).astype(np.int64)
    pairs = int((n * (n - 1) // 2).sum())
    return abi.Contacts(nw, pairs + 1, 8 * pairs + 64)


def assert_bodies_equal(a, b, what=""):
    for name in abi.BODY_INOUT:
        x, y = getattr(a, name), getattr(b, name)
        if x.tobytes() != y.tobytes():
            bad = np.argwhere(x.view(np.uint64) != y.view(np.uint64))
            r = int(bad[0][0])
            raise AssertionError(f"{what} body field {name} differs at row {r}: {x[r]!r} vs {y[r]!r} "
                                 f"({len(bad)} words differ)")


def assert_contacts_equal(a, b, what=""):
    ca, cb = a.canonical(), b.canonical()
    for k in ca:
        x, y = ca[k], cb[k]
        if x.shape != y.shape or x.tobytes() != y.tobytes():
            msg = f"{what} contacts field {k} differs: shapes {x.shape} vs {y.shape}"
            if x.shape == y.shape:
                bad = np.argwhere(x != y)
                i = tuple(bad[0]) if len(bad) else None
                msg += f", first at {i}: {x[i] if i is not None else ''} vs {y[i] if i is not None else ''}"
            raise AssertionError(msg)


def max_rel_dev(a, b):
    """Largest relative deviation over the in/out state (reported, not asserted: the bar is bit-exact)."""
    worst = 0.0
    for name in abi.BODY_INOUT:
        x, y = getattr(a, name), getattr(b, name)
        d = np.abs(x - y) / np.maximum(1.0, np.maximum(np.abs(x), np.abs(y)))
        worst = max(worst, float(d.max()) if d.size else 0.0)
    return worst


def lockstep(oracle, engine, shapes, bodies, params, steps, check_every=1, what=""):
    """Step oracle and GPU one simulate call at a time, each fed its own previous contacts through the
    host-buffer ABI (ep_step), and compare bit for bit every `check_every` steps."""
    bo, bg = bodies.copy(), bodies.copy()
    wo = wg = None
    engine.upload_shapes(shapes)
    for s in range(steps):
        oo, og = contacts_for(bodies), contacts_for(bodies)
        oracle.step(shapes, bo, params, wo, oo, 1)
        engine.step(bg, params, wg, og, 1)
        if s % check_every == 0 or s == steps - 1:
            assert_contacts_equal(oo, og, f"{what} step {s}")
            assert_bodies_equal(bo, bg, f"{what} step {s}")
        wo, wg = oo, og
    return bo, bg, wo, wg


def resident_vs_oracle(oracle, engine, shapes, bodies, params, steps, what=""):
    """N resident GPU steps (contacts fed back on device) against N oracle steps."""
    bo, bg = bodies.copy(), bodies.copy()
    oo, og = contacts_for(bodies), contacts_for(bodies)
    oracle.step(shapes, bo, params, None, oo, steps)
    engine.upload_shapes(shapes)
    engine.upload_worlds(bg, params, None)
    engine.step_resident(steps)
    engine.download_bodies(bg)
    engine.download_contacts(og)
    assert_contacts_equal(oo, og, what)
    assert_bodies_equal(bo, bg, what)
    return bo, bg, oo, og


# ---- randomized two-body worlds covering the 5x5 narrowphase matrix ------------------------------------
def _octo():
    he = 0.6
    return P.shape_unsafe_convex(
        [(2, 1, 0), (0, 5, 2), (1, 2, 4), (3, 0, 1), (2, 5, 4), (4, 3, 1), (5, 0, 3), (3, 4, 5)],
        [(0, 0, he), (0, he, 0), (he, 0, 0), (-he, 0, 0), (0, 0, -he), (0, -he, 0)])


def pair_templates():
    m = P.wood
    return [
        ("block", lambda: P.block((1.0, 0.8, 0.6), m)),
        ("sphere", lambda: P.sphere(0.5, P.rubber)),
        ("capsule", lambda: P.capsule(0.3, 1.0, P.steel)),
        ("cylinder", lambda: P.cylinder(0.5, 1.0, P.ice)),
        ("cone", lambda: P.cone(0.5, 1.0, P.plastic)),
        ("octo", lambda: P.dynamic([(_octo(), m)])),
        ("compound", lambda: P.dynamic([(P.shape_block((0.6, 0.6, 0.6), center=(-0.4, 0, 0)), m),
                                        (P.shape_sphere(0.35, center=(0.4, 0, 0)), P.rubber),
                                        (P.shape_capsule(0.2, 0.6, center=(0, 0.5, 0)), m)])),
        ("particle", lambda: P.point_mass((0, 0, 0), 1.5, P.steel)),
        ("plane", lambda: P.plane(P.wood)),
        ("static_block", lambda: P.static([(P.shape_block((2.0, 2.0, 0.5)), P.wood)])),
        ("kinematic_block", lambda: P.set_velocity_to((0.1, 0, 0.2), P.kinematic([(P.shape_block((1.0, 1.0, 0.4)), P.steel)]))),
    ]


def pair_worlds(seed, per_combo=6, gravity=(0.0, 0.0, -9.80665), spread=0.9):
    """One 2-body world per (template A, template B, trial): bodies placed `spread`-close with random
    orientation and velocity so that most pairs touch."""
    rng = np.random.RandomState(seed)
    tmpl = [(n, f()) for n, f in pair_templates()]
    worlds = []
    for na, a in tmpl:
        for nb_, b in tmpl:
            for _ in range(per_combo):
                def place(t, origin):
                    if t.kind == 1 and t.shapes[0][0].kind == 1:      # plane stays put
                        return t
                    q = rng.normal(size=4)
                    q = tuple(float(c) for c in q / np.linalg.norm(q))
                    if t.shapes[0][0].kind == 4:                       # point mass: no orientation
                        return P.move_to(origin, t)
                    out = P.place_quaternion(origin, q, t)
                    if t.kind == 2:
                        out = P.set_velocity_to(tuple(rng.uniform(-1, 1, 3)), out)
                        out = P.set_angular_velocity_to(tuple(rng.uniform(-2, 2, 3)), out)
                    return out
                oa = tuple(rng.uniform(-0.2, 0.2, 3) + np.array([0, 0, 0.45]))
                ob = tuple(np.array(oa) + rng.normal(size=3) * spread * 0.6)
                cfg = P.on_earth()
                cfg.gravity = gravity
                bodies, _ = P.assign_ids([("a", place(a, oa)), ("b", place(b, ob))])
                worlds.append((cfg, bodies))
    return worlds


def assert_world_contacts_equal(a, wa, b, wb, what=""):
    """World `wa` of contact block `a` == world `wb` of block `b`, bit for bit (groups, contacts, islands, iterations):
    the comparison behind the full-size tests, where `b` is the CPU oracle run on a SAMPLE of the worlds of `a`."""
    ga0, ga1 = int(a.world_group_off[wa]), int(a.world_group_off[wa + 1])
    gb0, gb1 = int(b.world_group_off[wb]), int(b.world_group_off[wb + 1])
    assert ga1 - ga0 == gb1 - gb0, f"{what}: world {wa}: {ga1 - ga0} vs {gb1 - gb0} pair groups"
    ca0, ca1 = int(a.group_contact_off[ga0]), int(a.group_contact_off[ga1])
    cb0, cb1 = int(b.group_contact_off[gb0]), int(b.group_contact_off[gb1])
    assert ca1 - ca0 == cb1 - cb0, f"{what}: world {wa}: {ca1 - ca0} vs {cb1 - cb0} contacts"
    assert ((a.group_contact_off[ga0:ga1 + 1] - ca0) == (b.group_contact_off[gb0:gb1 + 1] - cb0)).all(), f"{what}: world {wa}: group_contact_off"
    for name in ("group_id1", "group_id2", "group_n_constraints", "group_pre_pos1", "group_pre_quat1", "group_island"):
        x, y = getattr(a, name)[ga0:ga1], getattr(b, name)[gb0:gb1]
        assert x.tobytes() == y.tobytes(), f"{what}: world {wa}: {name} differs"
    for name, _, _ in abi.CONTACT_FIELDS:
        x, y = getattr(a, name)[ca0:ca1], getattr(b, name)[cb0:cb1]
        assert x.tobytes() == y.tobytes(), f"{what}: world {wa}: contact field {name} differs"
    assert a.iterations[wa] == b.iterations[wb] and a.world_n_islands[wa] == b.world_n_islands[wb], f"{what}: world {wa}: iterations / islands"


def sample_vs_oracle(oracle, engine, shapes, bodies, params, steps, sample, group_cap, contact_cap, what=""):
    """`steps` resident GPU steps of the WHOLE batch; the CPU oracle steps only the worlds in `sample` (as a batch of their
    own: worlds are independent, src/Physics.elm:522-525); those worlds must agree bit for bit -- bodies and contacts."""
    from elm_physics_b200 import scenes
    bs, ps, rows = scenes.slice_batch(bodies, params, sample)
    os_ = contacts_for(bs)
    oracle.step(shapes, bs, ps, None, os_, steps)
    bg = bodies.copy()
    og = abi.Contacts(bodies.n_worlds, group_cap, contact_cap)
    engine.upload_shapes(shapes)
    engine.upload_worlds(bg, params, None)
    engine.step_resident(steps)
    engine.sync()
    engine.download_bodies(bg)
    engine.download_contacts(og)
    for name in abi.BODY_INOUT:
        x, y = getattr(bg, name)[rows], getattr(bs, name)
        assert x.tobytes() == y.tobytes(), f"{what}: body field {name} differs in the sampled worlds"
    for k, w in enumerate(sample):
        assert_world_contacts_equal(og, int(w), os_, k, what)
    return bg, og, bs, os_
