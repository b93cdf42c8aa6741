"""The C-ABI library loads and exports every symbol include/pccgeo.h declares (no compute calls here), and the
ctypes signature table covers exactly the declared functions."""
import os
import re

from pcc_geo_cnn_v2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'pccgeo.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(pccgeo_[a-z0-9_]+)\s*\(', src))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    h = _lib.lib()
    for n in names:
        assert hasattr(h, n), f'libpccgeo.so does not export {n}'
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    assert h.pccgeo_version() >= 100
    assert h.pccgeo_reduce_ws_doubles() > 0


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, 'include', 'pccgeo.h')).read()
    assert src.count('src/model_types.py') >= 4 and src.count('src/model_transforms.py') >= 2
    assert 'patch_gaussian_conditional.py' in src and 'focal_loss.py' in src


def test_errors_are_reported_not_crashes():
    h = _lib.lib()
    rc = h.pccgeo_conv3d_f32(None, None, None, None, None, 1, 1, 8, 8, 8, 1, 3, 1, 0, 0, None)
    assert rc == -1 and b'null' in h.pccgeo_last_error()
    assert h.pccgeo_set_option(b'no_such_option', 1) == -1
