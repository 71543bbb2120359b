"""The C++ IDBGAligner-shaped shim compiles and reproduces a reference golden through the C-ABI
(linked against the host-emulation build so that it runs without a GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_roundtrip(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    emu_dir = os.path.join(ROOT, "tests", "emu", "build")
    exe = str(tmp_path / "test_shim")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                           "-o", exe, "-L" + emu_dir, "-lmgb_emu", "-Wl,-rpath," + emu_dir, "-fopenmp"])
    dbg = os.path.join(ROOT, "tests", "golden", "example_graphs", "test_DNA_graph.dbg")
    out = subprocess.run([exe, dbg], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "q1\tAGCTNCGAGGCCAA\t4=1X9=\t24" in out.stdout
    assert "dbg\t36=" in out.stdout
