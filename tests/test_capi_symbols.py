"""CPU test: the C-ABI library loads and exports every symbol include/acs.h declares (no device work)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "acs.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(str(ROOT / "aircombat_selfplay_b200" / "libacs.so"))
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_field_tables_are_consistent():
    from aircombat_selfplay_b200 import capi
    names = capi.state_field_names()
    assert len(names) == len(set(names)) == capi.lib().acs_n_state_fields()
    assert "q0" in names and "fcs/throttle-cmd-norm" in names
    out = capi.output_field_names()
    assert out[:3] == ["lon_deg", "lat_geod_deg", "h_sl_ft"]


def test_no_cpu_fallback():
    import torch
    from aircombat_selfplay_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.AcsError):
        capi.FdmBatch(4, 1)
