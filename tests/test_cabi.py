"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/nvsf_b200.h declares; the Python operator module has the reference's surface."""
import ctypes
import inspect
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "nvsf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nvsf_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_ten_reference_entry_points():
    names = _header_functions()
    # reference nvsf/nerf/raymarching/src/bindings.cpp:7-20
    for ref in ["packbits", "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert",
                "march_rays_train", "composite_rays_train_forward",
                "composite_rays_train_backward", "march_rays", "composite_rays"]:
        assert f"nvsf_{ref}" in names


def test_library_exports_every_declared_symbol(pkg):
    handle = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(handle, name), f"{name} declared in nvsf_b200.h but not exported"
    handle.nvsf_abi_version.restype = ctypes.c_int
    assert handle.nvsf_abi_version() == pkg._lib.ABI_VERSION


def test_python_binding_covers_header(pkg):
    assert sorted(pkg._lib.PROTOTYPES) == _header_functions()


def test_status_strings(pkg):
    L = pkg._lib.lib()
    assert L.nvsf_status_string(0) == b"ok"
    assert b"invalid" in L.nvsf_status_string(-1)
    assert b"workspace" in L.nvsf_status_string(-2)


def test_workspace_size_is_monotone(pkg):
    L = pkg._lib.lib()
    sizes = [L.nvsf_march_rays_train_workspace_bytes(n) for n in (0, 1, 128, 129, 4096, 529408)]
    assert sizes == sorted(sizes) and sizes[-1] >= 529408 * 4


def test_operator_module_surface(pkg):
    rm = pkg.raymarching
    # parameter lists of the reference wrappers (raymarching.py:18,54,87,113,139,174-191,295,
    # 370-388,466-479), ctx excluded
    expect = {
        "_near_far_from_aabb": ["rays_o", "rays_d", "aabb", "min_near"],
        "_sph_from_ray": ["rays_o", "rays_d", "radius"],
        "_morton3D": ["coords"],
        "_morton3D_invert": ["indices"],
        "_packbits": ["grid", "thresh", "bitfield"],
        "_march_rays_train": ["rays_o", "rays_d", "bound", "density_bitfield", "C", "H", "nears",
                              "fars", "step_counter", "mean_count", "perturb", "align",
                              "force_all_rays", "dt_gamma", "max_steps"],
        "_composite_rays_train": ["sigmas", "rgbs", "deltas", "rays", "T_thresh"],
        "_march_rays": ["n_alive", "n_step", "rays_alive", "rays_t", "rays_o", "rays_d", "bound",
                        "density_bitfield", "C", "H", "near", "far", "align", "perturb",
                        "dt_gamma", "max_steps"],
        "_composite_rays": ["n_alive", "n_step", "rays_alive", "rays_t", "sigmas", "rgbs",
                            "deltas", "weights_sum", "depth", "image", "T_thresh"],
    }
    for cls, params in expect.items():
        fwd = getattr(rm, cls).forward
        got = list(inspect.signature(fwd).parameters)[1:]
        assert got[:len(params)] == params, (cls, got)
    defaults = inspect.signature(rm._march_rays_train.forward).parameters
    assert defaults["max_steps"].default == 1024 and defaults["dt_gamma"].default == 0
    assert inspect.signature(rm._composite_rays_train.forward).parameters["T_thresh"].default == 1e-4
    assert inspect.signature(rm._composite_rays.forward).parameters["T_thresh"].default == 1e-2
    assert inspect.signature(rm._near_far_from_aabb.forward).parameters["min_near"].default == 0.2
    for name in rm.__all__:
        assert callable(getattr(rm, name))


def test_missing_library_fails_loudly(pkg, monkeypatch):
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libnvsf_b200.so")
    with pytest.raises(ImportError):
        pkg._lib.lib()


def test_option_scope_is_per_model_and_restores():
    """Tuning options are process-wide in the C ABI; the host mirror applies a model's own overrides only around
    that model's launches (one re-entrant lock) and restores the previous values — also when the body raises."""
    import importlib
    import threading
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    L = pkg._lib.lib()
    base = L.nvsf_get_option(b"heads_tc"), L.nvsf_get_option(b"mlp_bwd_tc")
    with pkg._lib.option_scope({"heads_tc": 2, "mlp_bwd_tc": 0}):
        assert (L.nvsf_get_option(b"heads_tc"), L.nvsf_get_option(b"mlp_bwd_tc")) == (2, 0)
        with pkg._lib.option_scope({"heads_tc": 4}):      # re-entrant (render -> autograd backward)
            assert L.nvsf_get_option(b"heads_tc") == 4
        assert L.nvsf_get_option(b"heads_tc") == 2
    assert (L.nvsf_get_option(b"heads_tc"), L.nvsf_get_option(b"mlp_bwd_tc")) == base
    with pytest.raises(pkg._lib.NvsfError):
        with pkg._lib.option_scope({"heads_tc": 2, "no_such_option": 1}):
            pass
    assert L.nvsf_get_option(b"heads_tc") == base[0]      # the first override was rolled back
    with pytest.raises(RuntimeError):
        with pkg._lib.option_scope({"heads_tc": 3}):
            raise RuntimeError("body failed")
    assert L.nvsf_get_option(b"heads_tc") == base[0]

    # another thread cannot observe (or disturb) the overrides of a scope in progress
    seen, inside, go = [], threading.Event(), threading.Event()

    def other():
        inside.wait()
        with pkg._lib.option_scope():          # what every NeRFNetwork entry point does
            seen.append(L.nvsf_get_option(b"heads_tc"))

    th = threading.Thread(target=other)
    th.start()
    with pkg._lib.option_scope({"heads_tc": 1}):
        inside.set()
        go.wait(0.2)                            # the other thread is blocked on the lock meanwhile
        assert seen == []
    th.join()
    assert seen == [base[0]]


def test_option_scope_wait_is_bounded(monkeypatch):
    """A thread that cannot get the option lock raises instead of waiting forever (e.g. .backward() inside a
    hand-opened scope: the autograd thread needs the same lock)."""
    import importlib
    import threading
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    monkeypatch.setattr(pkg._lib, "_OPTION_LOCK_TIMEOUT_S", 0.2)
    err = []

    def other():
        try:
            with pkg._lib.option_scope():
                pass
        except pkg._lib.NvsfError as e:
            err.append(str(e))

    with pkg._lib.option_scope({"heads_tc": 6}):
        th = threading.Thread(target=other)
        th.start()
        th.join()
    assert len(err) == 1 and "option lock" in err[0]
    with pkg._lib.option_scope():      # and the lock is free again afterwards
        pass
