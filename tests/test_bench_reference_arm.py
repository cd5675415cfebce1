"""CPU: `bench.py --impl reference` (the reference's own C library timed on the host cores, oracle/_ref) prints the
contract's JSON line.  Needs the compiled reference (oracle/build_ref.py, part of __graft_entry__.build())."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rosen_harness_f64")):
        if not os.path.exists("/root/reference/src/stochqn.c"):
            pytest.skip("oracle/_ref not built and the reference sources are not here")
        sys.path.insert(0, ROOT)
        from oracle import build_ref
        build_ref.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-budget", "4"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" not in line
    assert line["unit"] == "steps/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["n"] == 1 << 27 and line["gpu_launches"] == 0
    # the timed iterations run with the pair memory full (warm-up raised to mem_size + 2), and the final iterate's probes
    # travel with the line so that the CUDA arm can be compared with them
    assert line["warmup_run"] == 12 and line["check"]["mem_used"] == 10 and line["check"]["iterations"] == 14
    assert len(line["check"]["probes"]) == 7 and line["check"]["x_norm"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_dict(1 << 27, 1)
