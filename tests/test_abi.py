"""The C-ABI library loads without a GPU and exports every symbol include/remhos_b200.h declares;
without a CUDA device the context constructor fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'remhos_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rmh_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import remhos_b200 as rb
    lib = ctypes.CDLL(rb.LIB_PATH)
    names = declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.rmh_version() >= 100


def test_no_cpu_fallback():
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    import remhos_b200 as rb
    m = rb.Mesh.cartesian([3, 3], [1.0, 1.0], periodic=True).set_curvature(2)
    maps = m.dof_maps(1)
    with pytest.raises(rb.RmhError, match='no CUDA device'):
        rb.Context(dim=2, order=1, mesh_order=2, exec_mode=0, bounds_type=0, nodes=m.nodes(),
                   nbr_dof=maps['nbr_dof'], lat=maps['lat'], n_ent=maps['n_ent'],
                   nbr_elem=maps['nbr_elem'], vel_nodes=np.ones_like(m.nodes()))


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, 'remhos_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.hpp', '.h')):
                txt = open(os.path.join(base, f)).read()
                if 'remhos_oracle' in txt or 'oracle/' in txt:
                    bad.append(f)
    assert not bad, bad
