"""memcheck substitute (compute-sanitizer is closed on the GPU pool, profiles/r2a_sanitizer.txt): the kernels' per-environment
programs, compiled for the host with AddressSanitizer + UndefinedBehaviorSanitizer, run reset / wrapped step / unwrapped step /
debug forward for every kernel variant's model.  The scratch block is a heap allocation of exactly `smem_floats` floats, so any
access outside an environment's shared-memory slice, any table read past its end and any signed overflow / misaligned access in
the table logic aborts the run.  (Warp-level races are what the GPU invariance test covers.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

DRIVER = r'''
import sys, os, ctypes
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import emu, common
emu._LIB = ASAN_LIB
emu.build = lambda force=False: ASAN_LIB
from backends import EmuBackend
for name in ("rodent", "fly_free", "fly_tethered", "rodent_pair"):
    m, cfg, clip, tables = common.setup(name, 6)
    b = EmuBackend(tables)
    n = 3
    keys = common.jax_keys(n, seed=7)
    st, out = b.reset(keys)
    first = {k: v.copy() for k, v in st.items()}
    fo, fi = out["obs"].copy(), out["info_i"].copy()
    acts = common.actions(8, n, m.nu, seed=8, scale=0.5)
    for t in range(8):                       # episode_length 6: the truncation / auto-reset path runs too
        b.step(st, out, first, fo, fi, acts[t])
    b.physics_step(st, acts[0], 2)
    b.pipeline_init(st)
    b.reward_obs(st, out, acts[1])
    b.forward_debug(st, acts[2], 0)
    st2, _ = b.reset(keys, fixed_start_frame=249)
    assert np.isfinite(out["obs"]).all()
    print("ok", name)
'''


@pytest.mark.timeout(900)
def test_programs_under_asan_and_ubsan(tmp_path):
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    libubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(libasan) and os.path.exists(libasan)):
        pytest.skip("libasan is not installed")
    lib = str(tmp_path / "libbt_emu_asan.so")
    csrc = os.path.join(ROOT, "brax_tracking_b200", "csrc")
    subprocess.run(["g++", "-O1", "-g", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                    "-o", lib, os.path.join(HERE, "host_emu", "bt_emu.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=f"{libasan}:{libubsan}", ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:exitcode=23",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    code = DRIVER.replace("ROOT", repr(ROOT)).replace("ASAN_LIB", repr(lib))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    bad = [l for l in r.stderr.splitlines() if "AddressSanitizer" in l or "runtime error" in l]
    assert r.returncode == 0 and not bad, (r.returncode, bad[:5], r.stderr[-2000:])
    assert r.stdout.count("ok ") == 4
