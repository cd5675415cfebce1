"""GPU: a plain C host program (the reference example's calling pattern, host malloc arrays) linked against the CUDA
library prints the reference's known answers - the drop-in boundary exercised from C, no Python in the data path."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_relinks_and_reproduces_known_answers(tmp_path):
    libdir = os.path.join(ROOT, "stochqn_b200", "lib")
    exe = str(tmp_path / "dropin")
    subprocess.run(["gcc", "-std=c99", "-O1", os.path.join(ROOT, "tests", "dropin_sqn_rosen.c"), "-I" + os.path.join(ROOT, "include"),
                    "-L" + libdir, "-lstochqn_b200_f64", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    got = dict(line.split(" ", 1) for line in out if line and not line.startswith("it "))
    its = {int(line.split()[1]): line.split()[2] for line in out if line.startswith("it ")}
    assert got["f0"] == "266.6000"
    assert its[10] == "0.6755" and its[20] == "0.6633" and its[50] == "0.6303" and its[100] == "0.5798"
    assert its[150] == "0.5337" and its[200] == "0.4916"
    assert got["final"] == "0.4908"
    assert got["x"] == "1.048826 1.094597 1.253735 1.539749"
