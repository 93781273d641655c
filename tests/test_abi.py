"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/nunerf.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nunerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nunerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "nu_nerf_b200", "libnunerf_b200.so")
    assert os.path.exists(path), "build the engine first (make / __graft_entry__.build())"
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nunerf.h but not exported"
    lib.nunerf_version.restype = ctypes.c_int
    assert lib.nunerf_version() >= 100
    lib.nunerf_last_error.restype = ctypes.c_char_p
    assert lib.nunerf_last_error() is not None


def test_python_binding_covers_the_header():
    from nu_nerf_b200 import _lib
    assert sorted(_lib.ALL_SYMBOLS) == declared_symbols()


def test_argument_validation_without_gpu():
    """Error convention: negative return code + message, no CUDA call needed for bad arguments."""
    from nu_nerf_b200 import _lib
    p = _lib.LinearT()
    rc = _lib.lib.nunerf_linear(ctypes.byref(p), None)
    assert rc < 0 and b"linear" in _lib.lib.nunerf_last_error()
    rc = _lib.lib.nunerf_upsample(None, None, None, None, 4, 300, 16, None, 1.0, None, None, None, None, None, None)
    assert rc < 0


def test_chain_argument_validation_without_gpu():
    """nunerf_mlp_chain rejects malformed programs before any CUDA call (same error convention)."""
    from nu_nerf_b200 import _lib
    a = _lib.MlpChainT()
    assert _lib.lib.nunerf_mlp_chain(ctypes.byref(a), None) < 0                      # no input
    buf = (ctypes.c_char * 4096)()
    a.x, a.ldx, a.K0, a.M, a.n_layers = ctypes.addressof(buf), 384, 384, 16, 1       # input wider than 320 columns
    rc = _lib.lib.nunerf_mlp_chain(ctypes.byref(a), None)
    assert rc < 0 and b"mlp_chain" in _lib.lib.nunerf_last_error()
    a.ldx, a.K0, a.n_layers = 64, 64, 11                                              # too many layers
    assert _lib.lib.nunerf_mlp_chain(ctypes.byref(a), None) < 0


def test_nonzero_thickness_entry_points_validate_without_gpu():
    """The entry points added for network/renderer.py: same error convention, checked before any CUDA call -- null pointers /
    empty launches of the shell bounce, null operands of the shading encode with the new frequency / sphere_direction
    fields set."""
    from nu_nerf_b200 import _lib
    lib = _lib.lib
    n = None
    assert lib.nunerf_shell_bounce(n, n, n, n, n, n, 4, 0, n, n, n, n, n, n, n) < 0
    assert b"shell_bounce" in lib.nunerf_last_error()
    buf = (ctypes.c_char * 4096)()
    a = ctypes.addressof(buf)
    assert lib.nunerf_shell_bounce(a, a, a, a, a, a, 0, 0, a, a, a, a, a, a, n) < 0                     # M = 0
    assert lib.nunerf_shell_bounce_bwd(a, a, a, a, a, a, n, 4, 1, a, a, a, a, a, a, a, a, a, a, n) < 0    # no pass mask
    assert b"shell_bounce_bwd" in lib.nunerf_last_error()
    se = _lib.ShadeEncodeT()                     # (the frequency / row-width checks sit behind the constant-table upload)
    se.M, se.pos_freq, se.refrac_freq, se.sphere_direction = 8, 8, 2, 1
    rc = lib.nunerf_shade_encode_fwd(ctypes.byref(se), None)
    assert rc < 0 and b"shade_encode_fwd" in lib.nunerf_last_error()
    assert ctypes.sizeof(_lib.ShadeEncodeT) % 8 == 0 and ctypes.sizeof(_lib.ShadeMixT) % 8 == 0
