"""bench.py's safety net (no GPU needed): a section that misses its deadline must end the process with the partial line on
stdout -- marked `incomplete` -- instead of waiting for a collective timeout, and a section that finishes must leave no timer
behind."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code):
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, cwd=ROOT, timeout=120)


def test_deadline_prints_the_partial_line_and_exits_zero():
    r = _run('''
        import time, bench
        d = bench.Deadline(0)
        d.partial = {"metric": "audio-sec/sec", "value": 1.0, "e2e": None}
        d.arm(0.3, "end-to-end timing")
        time.sleep(30)
        print("not reached")
    ''')
    assert r.returncode == 0 and "not reached" not in r.stdout
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and line["e2e"] is None and "end-to-end timing" in line["incomplete"]
    assert "did not finish" in r.stderr


def test_deadline_without_a_measurement_fails_and_other_ranks_leave_quietly():
    r = _run('''
        import time, bench
        d = bench.Deadline(0)
        d.arm(0.3, "device-resident timing")
        time.sleep(30)
    ''')
    assert r.returncode == 3 and r.stdout.strip() == ""
    r = _run('''
        import time, bench
        d = bench.Deadline(5)
        d.partial = {}
        d.arm(0.3, "configs[4] sweep")
        time.sleep(30)
    ''')
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_disarmed_deadline_does_nothing():
    r = _run('''
        import time, bench
        d = bench.Deadline(0)
        d.partial = {"value": 2.0}
        d.arm(0.3, "x")
        d.arm(60, "y")          # re-arming cancels the first timer
        d.disarm()
        time.sleep(1.0)
        print("done")
    ''')
    assert r.returncode == 0 and r.stdout.strip() == "done"


def test_exception_after_the_headline_still_prints_the_line():
    r = _run('''
        import sys, bench
        def boom(args, rank, local_rank, world):
            d = bench.Deadline(rank)
            d.partial = {"metric": "audio-sec/sec", "value": 3.0, "e2e": None}
            d.arm(60, "end-to-end timing")
            raise RuntimeError("copy failed")
        bench.run_ours = boom
        sys.argv = ["bench.py"]
        bench.main()
        print("not reached")
    ''')
    assert r.returncode == 0 and "not reached" not in r.stdout
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 3.0 and "exception" in line["incomplete"]
    assert "RuntimeError: copy failed" in r.stderr


def test_exception_before_any_measurement_is_the_result():
    r = _run('''
        import sys, bench
        def boom(args, rank, local_rank, world):
            bench.Deadline(rank)
            raise RuntimeError("no device")
        bench.run_ours = boom
        sys.argv = ["bench.py"]
        bench.main()
    ''')
    assert r.returncode != 0 and r.stdout.strip() == "" and "RuntimeError: no device" in r.stderr
