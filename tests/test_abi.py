"""The C-ABI library: loads, exports every symbol include/ncb.h declares, and refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from nanocompore_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    with open(os.path.join(ROOT, 'include', 'ncb.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ncb_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.ncb_abi_version() == L.NCB_ABI_VERSION


def test_default_opts_are_the_reference_defaults():
    o = L.default_opts()        # config.py:196-214
    assert o.struct_size == C.sizeof(L.NcbOpts)
    assert o.tests == L.NCB_TEST_GMM | L.NCB_TEST_KS
    assert (o.n_samples, o.min_coverage, o.cluster_counts) == (4, 30, L.NCB_COUNTS_HARD)
    assert (o.em_max_iter, o.em_tol, o.reg_covar) == (100, 1e-3, 1e-6)


def test_enums_match_header():
    with open(os.path.join(ROOT, 'include', 'ncb.h')) as f:
        text = f.read()
    for name, table in (('PV', L.PV), ('ST', L.ST), ('DI', L.DI), ('DF', L.DF)):
        for key, val in table.items():
            m = re.search(rf'NCB_{name}_{key}\s*=\s*(\d+)', text)
            assert m and int(m.group(1)) == val, (name, key)
    for key in ('PV', 'ST', 'DI', 'DF'):
        m = re.search(rf'NCB_{key}_ROWS\s*=\s*(\d+)', text)
        assert int(m.group(1)) == getattr(L, f'NCB_{key}_ROWS')


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = L.load()
    assert lib.ncb_create(None, 0, None) == L.NCB_ERR_INVALID
    o = L.default_opts()
    o.struct_size = 3
    h = C.c_void_p()
    assert lib.ncb_create(C.byref(h), 0, C.byref(o)) == L.NCB_ERR_INVALID
    for field, value in (('em_tol', float('nan')), ('reg_covar', -1.0), ('min_coverage', -1),
                         ('em_max_iter', 0), ('n_samples', -1), ('depleted_condition', 2)):
        o = L.default_opts()
        setattr(o, field, value)
        assert lib.ncb_create(C.byref(h), 0, C.byref(o)) == L.NCB_ERR_INVALID and not h.value, field
    assert lib.ncb_compare(None, None, None, None) == L.NCB_ERR_INVALID
    assert lib.ncb_kernel_launches(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load()
    h = C.c_void_p()
    o = L.default_opts()
    assert lib.ncb_create(C.byref(h), 0, C.byref(o)) == L.NCB_ERR_CUDA and not h.value
    from nanocompore_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'nanocompore_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), fn


def test_struct_sizes_of_the_binding_and_of_the_documented_stub():
    """The library reports sizeof(ncb_opts / ncb_batch / ncb_results); the package's ctypes structures and the ones
    of the stub printed in INTEGRATION.md section 2 (executed here up to its first GPU call) have those sizes --
    ncb_default_opts writes a whole ncb_opts into the caller's struct."""
    import ctypes as C
    import re
    lib = L.load()
    assert lib.ncb_sizeof_opts() == C.sizeof(L.NcbOpts)
    assert lib.ncb_sizeof_batch() == C.sizeof(L.NcbBatch)
    assert lib.ncb_sizeof_results() == C.sizeof(L.NcbResults)
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r"```python\n# nanocompore/_ncb.py\n(.*?)```", text, flags=re.S).group(1)
    code = code.replace("C.CDLL('libncb.so')", f"C.CDLL({L.LIB_PATH!r})")
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)        # class definitions, CDLL, the size assertion
    assert C.sizeof(ns['Opts']) == lib.ncb_sizeof_opts() and C.sizeof(ns['Batch']) == lib.ncb_sizeof_batch()
    o = ns['Opts']()
    ns['lib'].ncb_default_opts(C.byref(o))
    assert o.struct_size == C.sizeof(ns['Opts']) and o.em_max_iter == 100
