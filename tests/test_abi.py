"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/bcg.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess
import pytest
from conftest import ROOT, PKG_DIR

LIB = os.path.join(PKG_DIR, 'bayesiancoresets_b200', 'lib', 'libbcg_b200.so')


def declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'bcg.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(bcg_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built_lib():
  if not os.path.exists(LIB):
    subprocess.check_call(['make', '-C', PKG_DIR])
  return LIB


def test_header_symbols_all_exported(built_lib):
  names = declared_symbols()
  assert len(names) >= 30
  L = ctypes.CDLL(built_lib)
  missing = [n for n in names if not hasattr(L, n)]
  assert missing == []


def test_binding_covers_header(built_lib):
  import bayesiancoresets_b200._native as nat
  assert sorted(nat.EXPORTED_SYMBOLS) == declared_symbols()
  assert nat.lib().bcg_abi_version() == 1


def test_no_cpu_fallback(built_lib):
  """without a CUDA device the product path must raise, never compute"""
  import bayesiancoresets_b200._native as nat
  n = ctypes.c_int(-1)
  rc = nat.lib().bcg_device_count(ctypes.byref(n))
  if rc == 0 and n.value > 0:
    pytest.skip('a GPU is present')
  import numpy as np
  import bayesiancoresets_b200 as bc
  with pytest.raises(nat.BcgError) as ei:
    bc.snnls.GIGA(np.random.randn(5, 20), np.ones(5))
  assert ei.value.code == 3   # BCG_ERR_NO_DEVICE
  src = open(os.path.join(PKG_DIR, 'bayesiancoresets_b200', '_native.py')).read()
  assert 'oracle' not in re.sub(r'""".*?"""', '', src, flags=re.S)


def test_pinned_allocation_needs_a_device_too(built_lib):
  """bc.pinned_empty is backed by bcg_host_alloc (cudaHostAlloc): without a CUDA device it raises like everything else"""
  import bayesiancoresets_b200._native as nat
  n = ctypes.c_int(-1)
  rc = nat.lib().bcg_device_count(ctypes.byref(n))
  if rc == 0 and n.value > 0:
    pytest.skip('a GPU is present')
  import bayesiancoresets_b200 as bc
  with pytest.raises(nat.BcgError) as ei:
    bc.pinned_empty((4, 3))
  assert ei.value.code == 3
  assert nat.lib().bcg_host_free(None) == 0


def test_product_package_never_imports_oracle():
  for dirpath, _, files in os.walk(os.path.join(PKG_DIR, 'bayesiancoresets_b200')):
    for f in files:
      if f.endswith('.py'):
        code = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle', code, flags=re.M), f
