"""The public headers are C: a user-style C99 program (tests/c_api_smoke.c) compiles against include/ and
links with the built library (CPU only; it evaluates a custom host force through the dispatcher)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c99_user_program_compiles_links_and_runs(built, tmp_path):
    exe = str(tmp_path / "c_api_smoke")
    lib = os.path.join(ROOT, "reboundx_b200", "lib")
    shim = os.path.join(ROOT, "reboundx_b200", "rebound_shim")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1",
           "-I", os.path.join(ROOT, "include"), "-I", shim, os.path.join(ROOT, "tests", "c_api_smoke.c"),
           "-L", lib, "-lreboundx_b200", "-L", shim, "-lrebound_shim", "-lm",
           f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{shim}", "-o", exe]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout)
    assert "c api ok" in run.stdout
