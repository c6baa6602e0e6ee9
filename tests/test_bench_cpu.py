"""bench.py contract pieces that run without a GPU: the recipe-style clock sampler (against a stand-in nvidia-smi) and the
reference arm's JSON line (the oracle timed on the host cores)."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clock_sampler_selects_samples_inside_timed_regions(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text('#!/bin/bash\nwhile true; do echo "$(date +"%Y/%m/%d %H:%M:%S.%3N"), 1905, 1965, Not Active, Not Active, Not Active, Active"; sleep 0.03; done\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0, period_ms=30)
    s.start(); time.sleep(0.2)
    s.mark_begin(); time.sleep(0.25); s.mark_end()
    time.sleep(0.15)
    s.mark_begin(); time.sleep(0.1); s.mark_end()
    s.stop()
    out = s.summary()
    assert out["window"] == "timed regions" and 4 <= out["samples"] < len(s.rows)      # only the samples inside the two windows
    assert out["sm_mhz"] == 1905.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    # no nvidia-smi at all: empty summary, no exception
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    s2 = bench.ClockSampler(0); s2.start(); s2.mark_begin(); s2.mark_end(); s2.stop()
    assert s2.summary()["samples"] == 0


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--pool", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.strip().splitlines()) == 1            # stdout carries the JSON line and nothing else
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "lidar_points_registered_per_s" and line["unit"] == "points/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_stdout_is_reserved_for_the_result_line():
    """claim_stdout(): whatever a library writes to file descriptor 1 during the run (NCCL prints its version banner there when
    torch.distributed creates the communicator) ends up on stderr; emit() still reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'banner from a library\\n'); "
            "print('a stray print'); bench.emit({'metric': 'x', 'value': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines() == ['{"metric": "x", "value": 1}']
    assert "banner from a library" in r.stderr and "a stray print" in r.stderr
