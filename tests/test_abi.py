"""CPU: the C-ABI shared library loads and exports every symbol include/dynamorph_b200.h declares; host-only
entry points (layout queries, error paths) behave.  No compute call is made without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from dynamorph_b200 import _lib
from dynamorph_b200._lib import DmbModel

HEADER = os.path.join(ROOT, "include", "dynamorph_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmb_[a-z0-9_]+)\s*\(", src)))


def _spec(**kw):
    s = DmbModel()
    s.arch, s.num_inputs, s.num_hiddens, s.num_residual_hiddens = 0, 2, 16, 32
    s.num_residual_layers, s.num_embeddings, s.height, s.width = 2, 64, 128, 128
    s.commitment_cost, s.weight_recon, s.weight_commitment, s.bn_eps, s.bn_momentum = 0.25, 1.0, 1.0, 1e-5, 0.1
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), name
    # and every binding the Python side declares exists in the header
    for name in _lib.SIGNATURES:
        assert name in syms, name
    assert lib.dmb_abi_version() == 1


def test_parameter_layout_matches_reference_state_dict():
    """Offsets by state_dict key: `params` is the concatenation of the trainable tensors in reference order."""
    from oracle import vqvae_oracle as O
    lib = _lib.load()
    for arch, name in ((0, "z16"), (1, "z32")):
        st = O.default_state(name)
        s = _spec(arch=arch)
        n_p, n_b, n_bn = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.call("dmb_param_count", C.byref(s), C.byref(n_p), C.byref(n_b), C.byref(n_bn))
        keys = O.trainable_keys(st)
        assert n_p.value == sum(st[k].numel() for k in keys)
        off = 0
        for k in keys:
            which, o, n = C.c_int32(), C.c_int64(), C.c_int64()
            _lib.call("dmb_param_lookup", C.byref(s), k.encode(), C.byref(which), C.byref(o), C.byref(n))
            assert (which.value, o.value, n.value) == (0, off, st[k].numel()), k
            off += st[k].numel()
        boff = 0
        for k in st:
            if k.endswith(("running_mean", "running_var")):
                which, o, n = C.c_int32(), C.c_int64(), C.c_int64()
                _lib.call("dmb_param_lookup", C.byref(s), k.encode(), C.byref(which), C.byref(o), C.byref(n))
                assert (which.value, o.value, n.value) == (1, boff, st[k].numel()), k
                boff += st[k].numel()
        assert n_b.value == boff
        d, lh, lw = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.call("dmb_latent_shape", C.byref(s), C.byref(d), C.byref(lh), C.byref(lw))
        assert (d.value, lh.value, lw.value) == ((16, 16, 16) if arch == 0 else (16, 32, 32))


def test_errors_are_reported_not_thrown():
    lib = _lib.load()
    with pytest.raises(_lib.DmbError, match="num_hiddens"):
        _lib.call("dmb_param_count", C.byref(_spec(num_hiddens=12)), None, None, None)
    with pytest.raises(_lib.DmbError, match="no such state_dict key"):
        _lib.call("dmb_param_lookup", C.byref(_spec()), b"enc.99.weight", None, None, None)
    with pytest.raises(_lib.DmbError, match="width"):
        _lib.call("dmb_workspace_bytes", C.byref(_spec(width=100)), 4, 0, 0, C.byref(C.c_size_t()))
    nbytes = C.c_size_t()
    _lib.call("dmb_workspace_bytes", C.byref(_spec()), 256, 1, 1, C.byref(nbytes))
    assert nbytes.value > 256 * 131072


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under dynamorph_b200/ may reference it."""
    pkg = os.path.join(ROOT, "dynamorph_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(root, f)
                assert "/root/reference" not in txt.replace("/root/reference/", "REFDOC/") or f.endswith(".py"), f
