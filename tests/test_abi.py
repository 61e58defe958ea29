"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from lidar_snow_sim_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, 'include')):
        if fn.endswith('.h'):
            txt = open(os.path.join(ROOT, 'include', fn)).read()
            names += re.findall(r'LSS_API[^;(]*?\b(lss_\w+)\s*\(', txt)
    return sorted(set(names))


def test_exports_every_declared_symbol(lib):
    from lidar_snow_sim_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 16
    bound = {s[0] for s in _lib.SIGNATURES}
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ but not exported'
        assert n in bound, f'{n} declared in include/ but not bound in _lib.SIGNATURES'
    assert bound <= set(names)


def test_version_and_strings(lib):
    assert lib.lss_version() >= 100
    assert lib.lss_status_string(0) == b'ok'
    assert b'120 m' in lib.lss_status_string(4)


def test_range_grid_matches_numpy(lib, oracle):
    R = np.zeros(1230)
    assert lib.lss_debug_range_grid(ctypes.c_void_p(R.ctypes.data)) == 0
    assert np.array_equal(R, oracle.range_grid())


def test_no_cpu_fallback(lib):
    """Without a GPU the engine must fail loudly, never fall back to host code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    assert lib.lss_create(0, ctypes.byref(h)) != 0
    from lidar_snow_sim_b200.engine import SnowfallEngine
    with pytest.raises(RuntimeError):
        SnowfallEngine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'lidar_snow_sim_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), f'{f} imports the oracle'
                assert 'liboracle' not in txt
                assert '/root/reference' not in txt


def test_focal_offset_matches_python():
    from lidar_snow_sim_b200.calib.hdl64e_s3 import sensor_arrays
    fd = sensor_arrays()[0]
    for v in fd:
        t = 1 - (v * 100) / 13100
        assert (1 - v * 100 / 13100) ** 2 == t * t          # simulation.py:76: pow(x, 2) == x*x for these values


def _header_params(name):
    txt = open(os.path.join(ROOT, 'include', 'lidar_snow_sim.h')).read()
    m = re.search(r'LSS_API[^;(]*?\b' + name + r'\s*\(([^;]*?)\)\s*;', txt, re.S)
    assert m, name
    return [p.strip() for p in m.group(1).split(',') if p.strip()]


def _call_args(text, call):
    """Arguments of the first `call(...)` in text (nesting aware, comments stripped)."""
    i = text.index(call + '(') + len(call) + 1
    depth, args, cur = 1, [], ''
    body = re.sub(r'#[^\n]*|/\*.*?\*/', '', text[i:], flags=re.S)
    for ch in body:
        if ch in '([':
            depth += 1
        elif ch in ')]':
            depth -= 1
            if depth == 0:
                break
        if ch == ',' and depth == 1:
            args.append(cur.strip())
            cur = ''
        else:
            cur += ch
    args.append(cur.strip())
    return [a for a in args if a]


def test_signatures_and_documented_calls_have_the_header_arity():
    """ctypes argtypes and the raw C-ABI examples of INTEGRATION.md follow the header's parameter lists."""
    from lidar_snow_sim_b200 import _lib
    for name, _, argtypes in _lib.SIGNATURES:
        params = _header_params(name)
        n = 0 if params == ['void'] else len(params)
        assert n == len(argtypes), f'{name}: header has {n} parameters, _lib.SIGNATURES {len(argtypes)}'
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    for call, name in (('L.lss_snowfall_batch', 'lss_snowfall_batch'),
                       ('\nlss_snowfall_batch_host_submit', 'lss_snowfall_batch_host_submit')):
        assert len(_call_args(doc, call)) == len(_header_params(name)), f'INTEGRATION.md: {name} example is out of date'
