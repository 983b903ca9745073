"""GPU vs oracle at the sizes BASELINE.json names (configs[2..4]): the whole batch runs on the GPU through the C ABI, the
CPU oracle steps a seeded SAMPLE of its worlds (worlds are independent `simulate` calls, src/Physics.elm:522-525) or, for
the single large scene, the whole world from a saved settled state.  Bit-exact, like every parity test here."""
import os

import numpy as np
import pytest

from elm_physics_b200 import abi
from elm_physics_b200 import pack
from elm_physics_b200 import scenes

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from elm_physics_b200 import engine as eng
    e = eng.Engine(0)
    yield e
    e.close()


def test_c3_8192_worlds_sample_vs_oracle(engine, oracle):
    """BASELINE configs[2]: 8192 worlds x (plane + 32 mixed bodies), 150 steps from the drop (fall, first impacts, piles
    forming: every island size class of the batched solver); 64 seeded worlds of the batch against the oracle."""
    shapes, bodies, params = scenes.mixed_worlds(8192)
    sample = np.sort(np.random.RandomState(2026).choice(8192, 64, replace=False))
    bg, og, _, os_ = H.sample_vs_oracle(oracle, engine, shapes, bodies, params, 150, sample, 64 * 8192, 160 * 8192, what="c3x8192")
    assert os_.n_contacts > 64 * 40 and og.n_contacts > 8192 * 40
    assert np.isfinite(bg.pos).all()


def test_c5_65536_car_worlds_sample_vs_oracle(engine, oracle):
    """BASELINE configs[4]: 65536 car worlds (chassis + 4 hinged wheels + ramp + 6 hollow crates + tether: 5 constraints,
    21 joint rows per world), every world different, 60 steps; 64 seeded worlds against the oracle, plus 4 downward rays
    per world (the RaycastCar pattern, examples/src/RaycastCar/Car.elm:116-128) for the sampled worlds."""
    nw = 65536
    shapes, bodies, params = scenes.car_batch(nw)
    assert params.n_constraints == 5 * nw
    sample = np.sort(np.random.RandomState(65536).choice(nw, 64, replace=False))
    bg, og, bs, os_ = H.sample_vs_oracle(oracle, engine, shapes, bodies, params, 60, sample, 40 * nw, 64 * nw, what="c5x65536")
    assert int(og.group_n_constraints[:og.n_groups].sum()) == 5 * nw
    # rays from the four chassis corners straight down, against the resident batch and against the oracle's sample
    ps = scenes.slice_batch(bodies, params, sample)[1]
    chassis = np.array([int(bodies.world_body_off[w]) + 2 for w in sample])
    corners = np.array([[-1.0, 2.0, 0.0], [1.0, 2.0, 0.0], [-1.0, -2.0, 0.0], [1.0, -2.0, 0.0]])
    frm = (bg.pos[chassis][:, None, :] + corners[None, :, :]).reshape(-1, 3)
    drc = np.tile(np.array([[0.0, 0.0, -1.0]]), (len(frm), 1))
    hit_g = engine.raycast(np.repeat(sample, 4).astype(np.int32), frm, drc)
    hit_o = oracle.raycast(shapes, bs, np.repeat(np.arange(64), 4).astype(np.int32), frm, drc)
    rows = scenes.slice_batch(bodies, params, sample)[2]
    assert (np.where(hit_g[0] >= 0, hit_g[0], -1) == np.where(hit_o[0] >= 0, rows[np.maximum(hit_o[0], 0)], -1)).all()
    for k in (1, 2, 3):
        assert hit_g[k].tobytes() == hit_o[k].tobytes()
    assert (hit_g[0] >= 0).sum() > 200


def test_c4_2000_body_pile_lockstep_from_settled_state(engine, oracle):
    """BASELINE configs[3] at full size: plane + pit + 2000 hulls / compounds.  The GPU drops and piles them (700 resident
    steps), the settled state (bodies + contacts) is saved, and from it GPU and oracle run 100 frames in lock step, each fed
    its own previous contacts: the big island (thousands of pair groups, dozens of dependency levels) goes through the
    multi-CTA island solver every frame.  The canonical-order re-summation of sum|dlambda| is forced on (EP_FORCE_RESUM,
    read per step) for the second half of the frames, so both decision paths of the early exit are compared."""
    shapes, bodies, params = pack.pack_worlds(scenes.pile_scene(2000))
    nb = bodies.n_bodies
    engine.upload_shapes(shapes)
    engine.configure(contact_cap=1 << 17)
    b = bodies.copy()
    engine.upload_worlds(b, params, None)
    engine.step_resident(700)
    engine.sync()
    engine.download_bodies(b)
    warm = abi.Contacts(1, 40 * nb, 64 * nb)
    engine.download_contacts(warm)
    assert warm.n_contacts > 6000, warm.n_contacts
    biggest = np.bincount(warm.group_island[:warm.n_groups][warm.group_island[:warm.n_groups] >= 0]).max()
    assert biggest > 2000, biggest          # one island holds most of the pile
    bo, bg = b.copy(), b.copy()
    wo = wg = warm
    try:
        for s in range(100):
            if s == 50:
                os.environ["EP_FORCE_RESUM"] = "1"
            oo, og = abi.Contacts(1, 40 * nb, 64 * nb), abi.Contacts(1, 40 * nb, 64 * nb)
            oracle.step(shapes, bo, params, wo, oo, 1)
            engine.step(bg, params, wg, og, 1)
            H.assert_contacts_equal(oo, og, f"pile2000 frame {s}")
            H.assert_bodies_equal(bo, bg, f"pile2000 frame {s}")
            wo, wg = oo, og
    finally:
        os.environ.pop("EP_FORCE_RESUM", None)
        engine.configure(contact_cap=0)
    st = engine.stats()
    assert st.get("resums", 1) > 0
