"""The C-ABI library loads without a GPU and exports every symbol include/coordx_b200.h declares
(no compute calls here); argument errors are reported through the error channel, not by crashing."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'coordx_b200.h')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as G
    G.build()
    from coordx_b200 import _lib
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(coordx_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported(lib):
    from coordx_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f'{n} is declared in include/coordx_b200.h but not exported'
    # the ctypes binding table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == names


def test_struct_layout_matches_header(lib):
    from coordx_b200 import _lib
    src = open(HEADER).read()
    assert int(re.search(r'#define COORDX_MAX_BRANCHES (\d+)', src).group(1)) == _lib.MAX_BRANCHES
    assert int(re.search(r'#define COORDX_MAX_LAYERS (\d+)', src).group(1)) == _lib.MAX_LAYERS
    # 4 + 4*4 + ... : int32 block, float, then pointers (8-byte aligned)
    n_i32 = 1 + _lib.MAX_BRANCHES + 5 + 2 * _lib.MAX_LAYERS + 4
    ptrs = 2 * _lib.MAX_BRANCHES + 2 * _lib.MAX_LAYERS + 1
    expect = (n_i32 * 4 + 4 + 7) // 8 * 8 + ptrs * 8
    assert ctypes.sizeof(_lib.Net) == expect
    assert ctypes.sizeof(_lib.Grads) == ptrs * 8
    assert lib.coordx_abi_version() == int(re.search(r'#define COORDX_ABI_VERSION (\d+)', src).group(1))


def test_argument_errors_without_gpu(lib):
    from coordx_b200 import _lib
    net = _lib.Net()                      # all zero: n_branches = 0 is out of range
    assert lib.coordx_trunk_tc_supported(net, _lib.PREC_BF16) == 0
    assert lib.coordx_branch_workspace_bytes(net, 0, 16, 0) == 0
    rc = lib.coordx_branch_forward(net, 0, None, 16, None, None, 0, None)
    assert rc == -1 and b'n_branches' in lib.coordx_last_error()
    net.n_branches, net.n_layers, net.n_pre, net.hidden, net.fuse_width = 2, 4, 2, 256, 256
    net.fusion_op = 7                     # the reference silently ignores unknown operators (modules.py:336-341)
    rc = lib.coordx_trunk_forward(net, None, None, 0, 0, None, 0, None, 0, None)
    assert rc == -1 and b'fusion' in lib.coordx_last_error()


def test_product_never_imports_oracle():
    """only tests/, smoke() and bench.py's cpu legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'coordx_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f), errors='replace').read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_forward_with_loss_validates_its_arguments_and_never_falls_back_to_a_cpu_path():
    """host logic of SingleBVPNet.forward_with_loss (the loss fused into the trunk's last layer): an unknown loss is a
    ValueError before anything is launched, CPU tensors raise like every other hot-path entry (no CPU fallback), and the
    loss hook never stays set on the module."""
    import pytest
    import torch
    from coordx_b200 import modules as M, _lib
    m = M.SingleBVPNet(type='sine', out_features=1, hidden_features=64, split_mlp=True, fusion_operator='prod', quiet=True)
    x = torch.linspace(-1, 1, 8).reshape(1, 8, 1, 1)
    y = torch.linspace(-1, 1, 6).reshape(1, 1, 6, 1)
    gt = torch.zeros(1, 48, 1)
    with pytest.raises(ValueError):
        m.forward_with_loss({'coords_split': [x, y]}, gt, 'huber')
    with pytest.raises(_lib.CoordxError):
        m.forward_with_loss({'coords_split': [x, y]}, gt, 'mse')
    assert m.net.__dict__.get('_loss_hook') is None
