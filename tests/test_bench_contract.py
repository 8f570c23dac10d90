"""bench.py's reference arm runs on the CPU (the reference algorithm on the host cores): its JSON line must carry the keys
the driver reads.  (The GPU arm prints the same keys plus roofline / clocks / gpu_launches; it needs a B200.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--batch", "8", "--morph", "3d_hopper_3_shin"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "SET TD3 update samples/sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["value"] > 0 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("3d_hopper_3_shin TD3 Agent.update")
    cb = line["cpu_baseline"]
    from oracle import ref_loader
    # the live reference when it is on this machine (/root/reference/src, baseline/_ref/src, $SGRL_REF), else the oracle port
    assert cb["kind"] == ("reference" if ref_loader.find_reference() else "port")
    assert (cb["kind"] == "reference") == ("port_over_reference_time" in cb)
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_falls_back_to_the_port_without_the_reference_tree(tmp_path):
    env = dict(os.environ, SGRL_REF="", SGRL_REF_DISABLE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--batch", "8", "--morph", "3d_hopper_3_shin"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
