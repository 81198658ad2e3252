"""bench.py pieces that run without a GPU: the reference arm's JSON line (the CPU oracle port on a bounded sample) and the
helper behind roofline.traffic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys_and_names_its_sample():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                      # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "clips/sec, r21d_byol pretrain step, 16x112x112"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm runs a bounded sample of the batch-60 workload and says so in its config
    assert d["config"]["per_gpu_batch"] == 60 and d["config"]["sampled_batch"] == 2 and d["config"]["same_config"] is False


def test_roofline_traffic_comes_from_a_committed_capture_with_its_provenance():
    sys.path.insert(0, ROOT)
    import bench
    per_launch, src = bench.ncu_traffic("conv_halo_kernel", 60)
    assert src.startswith("profiles/r02_ncu_conv_halo_b60_dram.csv") and "commit" in src
    assert 1.5e9 < per_launch < 2.5e9          # ~1.88 GB per launch = the algorithmic bytes of the layers it runs
    assert bench.ncu_traffic("conv_halo_kernel", 7) == (None, None)


def test_watchdog_prints_the_measured_line_when_a_sub_record_hangs():
    """bench.py at N > 1: a config-5 sub-record that never returns (a lost peer) must not lose the batch-60 line."""
    code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
            "line = {'metric': 'm', 'value': 1.0}\n"
            "bench.run_with_watchdog(0.3, lambda: time.sleep(30), lambda: bench.bail_out(line, 'config5', 'too slow'))\n"
            "print('not reached')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.0, "config5": {"error": "too slow"}}
    # and a sub-record that returns in time passes its result through, the timer cancelled
    sys.path.insert(0, ROOT)
    import bench
    assert bench.run_with_watchdog(5.0, lambda: {"ok": 1}, lambda: (_ for _ in ()).throw(AssertionError("fired"))) == {"ok": 1}
