#!/usr/bin/env python3
"""Run the kernel-source emulation tests with the emulation compiled under AddressSanitizer + UndefinedBehaviorSanitizer
(development tool, CPU only): every index into the per-environment Work record, every lane map and every variant
(jvrc_walk / h1 / jvrc_step / terrain, fp64 and fp32, NaN and runaway states included) is then bounds- and UB-checked on the
same source the GPU executes.  usage: python tools/emu_sanitize.py [--oracle] [-DLHW_X_<candidate>=1 ...]   (exit code 0 = no finding;
--oracle also rebuilds oracle/sim_oracle.c under the sanitizers for the run and adds the oracle's own tests)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tests", "emu", "_build", "libsim_emu.SANITIZE_1.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.check_call(["g++", "-O1", "-g", "-fPIC", "-shared", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
                      + [a for a in sys.argv[1:] if a.startswith("-D")] + ["-o", lib, os.path.join(ROOT, "tests", "emu", "sim_emu.cpp")])
asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
env = dict(os.environ, LHW_EMU_DEFINES="SANITIZE=1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
tests = ["tests/test_kernel_source_emulation.py", "tests/test_h1_oracle.py", "tests/test_step_oracle.py"]
olib = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
if "--oracle" in sys.argv:      # the C restatement too: it is the checker, an out-of-bounds read there would bend every parity claim
    sys.path.insert(0, ROOT)
    from oracle.oracle import build as build_oracle
    build_oracle()
    os.replace(olib, olib + ".plain")
    subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-shared", "-fopenmp", "-std=gnu11", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", "-o", olib, os.path.join(ROOT, "oracle", "sim_oracle.c"), "-lm"])
    tests += ["tests/test_oracle_golden.py", "tests/test_physics_invariants.py"]
try:
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", "-k", "not flagged_candidate"] + tests, cwd=ROOT, env=env,
                       capture_output=True, text=True)
finally:
    os.remove(lib)
    if os.path.exists(olib + ".plain"):
        os.replace(olib + ".plain", olib)
        os.utime(olib)
findings = [l for l in (r.stdout + r.stderr).splitlines() if "runtime error" in l or "AddressSanitizer" in l]
print((r.stdout.strip().splitlines() or [""])[-1])
print(f"{len(findings)} sanitizer finding(s)")
for l in findings[:20]:
    print(l)
sys.exit(1 if findings or r.returncode else 0)
