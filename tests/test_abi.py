"""The C-ABI libraries load without a GPU, export every symbol the headers declare, agree with the ctypes mirror on
struct layouts, and the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

from epidemicsimulator_b200 import _abi
from epidemicsimulator_b200._lib import CUDA_LIB, HOST_LIB, cuda_lib, host_lib

ROOT = Path(__file__).resolve().parent.parent


def declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//.*", "", text)
    return sorted(set(re.findall(r"\b(esim_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    names = sorted(set(declared_functions(ROOT / "include" / "esim.h") + declared_functions(ROOT / "include" / "esim_popgen_device.h")))
    assert len(names) >= 40
    lib = cuda_lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_library_exports_every_declared_symbol():
    names = declared_functions(ROOT / "include" / "esim_popgen.h")
    assert len(names) >= 10
    lib = host_lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "esim.h"
#include "esim_popgen.h"
int main(void) {
  printf("EsimConfig %zu %zu %zu\n", sizeof(EsimConfig), offsetof(EsimConfig, exposed_time), offsetof(EsimConfig, seed));
  printf("EsimPopulationSoA %zu %zu %zu\n", sizeof(EsimPopulationSoA), offsetof(EsimPopulationSoA, home_bldg), offsetof(EsimPopulationSoA, room_bldg));
  printf("EsimStepStats %zu %zu %zu\n", sizeof(EsimStepStats), offsetof(EsimStepStats, lockdown_hours), offsetof(EsimStepStats, vaccinated_now));
  printf("EsimStateView %zu %zu %zu\n", sizeof(EsimStateView), offsetof(EsimStateView, timer), offsetof(EsimStateView, vax_eligible));
  printf("EsimTimings %zu %zu %zu\n", sizeof(EsimTimings), offsetof(EsimTimings, k_update), offsetof(EsimTimings, steps));
  printf("EsimPopgenParams %zu %zu %zu\n", sizeof(EsimPopgenParams), offsetof(EsimPopgenParams, p_student), offsetof(EsimPopgenParams, initial_infected));
  return 0;
}''')
    exe = tmp_path / "sizes"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {ln.split()[0]: tuple(int(x) for x in ln.split()[1:]) for ln in out if ln}
    def lay(t, a, b):
        return (C.sizeof(t), getattr(t, a).offset, getattr(t, b).offset)
    assert got["EsimConfig"] == lay(_abi.EsimConfig, "exposed_time", "seed")
    assert got["EsimPopulationSoA"] == lay(_abi.EsimPopulationSoA, "home_bldg", "room_bldg")
    assert got["EsimStepStats"] == lay(_abi.EsimStepStats, "lockdown_hours", "vaccinated_now")
    assert got["EsimStateView"] == lay(_abi.EsimStateView, "timer", "vax_eligible")
    assert got["EsimTimings"] == lay(_abi.EsimTimings, "k_update", "steps")
    assert got["EsimPopgenParams"] == lay(_abi.EsimPopgenParams, "p_student", "initial_infected")


def test_default_config_is_disease_model_covid():
    cfg = _abi.EsimConfig()
    assert cuda_lib().esim_default_config(C.byref(cfg)) == 0
    from oracle.oracle_py import default_config
    ref = default_config()
    for name, _ in _abi.EsimConfig._fields_:
        assert getattr(cfg, name) == getattr(ref, name), name


def test_no_cpu_fallback_without_a_device():
    from tests.conftest import has_gpu
    if has_gpu():
        pytest.skip("a CUDA device is present")
    from epidemicsimulator_b200.simulator import Simulator
    with pytest.raises(_abi.SimError) as e:
        Simulator()
    assert e.value.code == _abi.ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "epidemicsimulator_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        text = path.read_text()
        assert "oracle" not in text.lower(), "%s mentions the oracle" % path
