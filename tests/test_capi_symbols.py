"""CPU tier: the C-ABI library loads and exports every symbol include/btkb200.h declares; argument checks
that need no device.  No compute calls here."""
import os
import re

import pytest

import btk_b200
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "btkb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(btkb200_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = btk_b200.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/btkb200.h but not exported"
    assert sorted(btk_b200._capi.SYMBOLS) == names


def test_no_cpu_fallback_without_device():
    if btk_b200.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(btk_b200.BtkError) as e:
        btk_b200.Plan(256, 4, 1, 8)
    assert e.value.code == btk_b200._capi.ECUDA


def test_unsupported_geometry_is_reported():
    with pytest.raises(btk_b200.BtkError) as e:
        btk_b200.Plan(200, 2, 1, 2)
    assert e.value.code == btk_b200._capi.EUNSUPPORTED


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "distantspeechrecognition-mirror_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "btk_oracle" not in src and "libbtk_ref" not in src and "oracle/" not in src, f
