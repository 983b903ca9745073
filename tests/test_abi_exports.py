"""The C-ABI shared library loads without a GPU and exports every function include/elm_physics_b200.h
declares; without a CUDA device ep_create fails loudly (EP_ENODEVICE) — there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "elm_physics_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ep_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("ep_create", "ep_destroy", "ep_upload_shapes", "ep_upload_worlds", "ep_step_resident", "ep_step",
                 "ep_download_bodies", "ep_download_contacts", "ep_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from elm_physics_b200 import engine
    lib = engine.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(engine.EXPORTS) <= set(declared_functions())


def test_struct_layouts_match_the_header():
    """ctypes mirrors (abi.py) have one field per header member, in order."""
    from elm_physics_b200 import abi
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for struct, mirror in (("ep_shapes", abi.EpShapes), ("ep_bodies", abi.EpBodies), ("ep_params", abi.EpParams), ("ep_contacts", abi.EpContacts)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, re.S).group(1)
        members = [re.findall(r"(\w+)\s*$", decl.strip())[0] for decl in body.split(";") if decl.strip()]
        mirrored = [n.rstrip("_") for n, _ in mirror._fields_]
        assert members == mirrored, (struct, members, mirrored)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from elm_physics_b200 import abi, engine
    lib = engine.load_library()
    ctx = C.c_void_p()
    assert lib.ep_create(C.byref(ctx), 0) == abi.EP_ENODEVICE
    with pytest.raises(engine.EngineError):
        engine.Engine(0)
