"""CPU suite: the C-ABI library loads and exports every symbol include/bskenv.h declares
(no compute calls -- there is no GPU here), and refuses to run without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bskenv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bskenv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from basilisk_env_b200 import _native
    lib = C.CDLL(_native.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bskenv.h but not exported by libbskenv.so"
    assert sorted(_native.EXPORTS + _native.OPNAV_EXPORTS) == syms


def test_config_struct_matches_header_defaults():
    from basilisk_env_b200 import _native
    cfg = _native.default_config()
    assert cfg.abi_version == 1
    assert (cfg.dynRate, cfg.fswRate, cfg.step_duration) == (0.1, 1.0, 180.0)
    assert (cfg.mass, cfg.width, cfg.depth, cfg.height) == (330.0, 1.38, 1.04, 1.58)
    assert list(cfg.nHat_B) == [0.0, -1.0, 0.0] and list(cfg.sigma_R0N) == [1.0, 0.0, 0.0]
    assert (cfg.K, cfg.Ki, cfg.P) == (7.0, -1.0, 35.0)
    assert (cfg.thrForceSign, cfg.maxCounterValue, cfg.max_length, cfg.auto_reset) == (1, 4, 540, 0)
    assert (cfg.wheel_limit_rpm, cfg.power_max, cfg.failure_penalty) == (3000.0, 20.0, 1.0)
    assert cfg.use_j2 == 0 and cfg.hill_cel_pun == 0


def test_config_struct_layout_matches_c_compiler(tmp_path):
    """The ctypes mirror of bskenv_config has the size/offsets gcc gives the header's struct."""
    import subprocess
    from basilisk_env_b200 import _native
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bskenv.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(bskenv_config), offsetof(bskenv_config, mass),'
                   'offsetof(bskenv_config, nHat_B), offsetof(bskenv_config, thrForceSign), offsetof(bskenv_config, wheel_limit_rpm),'
                   'offsetof(bskenv_config, use_j2));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    Cfg = _native.Config
    want = [C.sizeof(Cfg), Cfg.mass.offset, Cfg.nHat_B.offset, Cfg.thrForceSign.offset, Cfg.wheel_limit_rpm.offset, Cfg.use_j2.offset]
    assert got == want


def test_opnav_config_struct_layout_matches_c_compiler(tmp_path):
    import subprocess
    from basilisk_env_b200 import _native
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bskenv.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(bskenv_opnav_config), offsetof(bskenv_opnav_config, max_length),'
                   'offsetof(bskenv_opnav_config, pixel_noise_std), offsetof(bskenv_opnav_config, noise_seed),'
                   'offsetof(bskenv_opnav_config, reserved));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    Cfg = _native.OpNavConfig
    assert got == [C.sizeof(Cfg), Cfg.max_length.offset, Cfg.pixel_noise_std.offset, Cfg.noise_seed.offset, Cfg.reserved.offset]
    cfg = _native.opnav_default_config()
    assert (cfg.dynRate, cfg.fswRate, cfg.step_duration_min, cfg.max_length, cfg.numModes) == (1.0, 1.0, 50.0, 40, 50)
    assert (cfg.nav_noise, cfg.camera_reenable, cfg.sample_orbit, cfg.auto_reset) == (1, 0, 0, 0)
    assert _native.opnav_state_field("filter_sBar") == (45, False) and _native.opnav_state_field("n_meas") == (9, True)


def test_state_field_table():
    from basilisk_env_b200 import _native
    assert _native.state_field("r_BN_N") == (0, False)
    assert _native.state_field("tick") == (0, True)
    with pytest.raises(KeyError):
        _native.state_field("no_such_field")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from basilisk_env_b200 import _native, BskEnvError, LeoPowerAttVecEnv
    L = _native.lib()
    h = C.c_void_p()
    cfg = _native.default_config()
    rc = L.bskenv_create(C.byref(cfg), 0, 4, 0, C.byref(h))
    assert rc == -3 and b"no CUDA device" in L.bskenv_last_error(None)
    with pytest.raises(BskEnvError):
        LeoPowerAttVecEnv(4)
    ocfg = _native.opnav_default_config()
    rc = L.bskenv_opnav_create(C.byref(ocfg), 0, 4, 0, C.byref(h))
    assert rc == -3 and b"no CUDA device" in L.bskenv_opnav_last_error(None)
    from basilisk_env_b200.opnav_env import OpNavVecEnv
    with pytest.raises(BskEnvError):
        OpNavVecEnv(4)


def test_product_does_not_import_oracle():
    """Nothing under basilisk_env_b200/ may import, link or execute oracle/ (or the host-compiled core)."""
    pkg = os.path.join(ROOT, "basilisk_env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "bsk_oracle" not in text, f
                assert "hostcore" not in text or f in ("leo_core.cuh", "opnav_core.cuh"), f
