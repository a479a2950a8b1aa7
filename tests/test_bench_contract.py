"""bench.py's contract with the driver, as far as it can be checked without a GPU: exactly one JSON line on stdout."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_carries_only_the_json_line_even_if_a_library_writes_to_fd_1():
    code = textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        import bench
        def arm(args):
            os.write(1, b"NCCL version 2.28.9+cuda12.9\\n")   # what NCCL does under torchrun
            print('{"metric": "m", "value": 1.0}')
            return 0
        bench.run_b200_arm = arm
        sys.argv = ["bench.py", "--gpus", "2"]
        sys.exit(bench.main())
    ''' % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"metric": "m", "value": 1.0}
    assert "NCCL version" in r.stderr


def test_reference_arm_prints_one_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "bins/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
