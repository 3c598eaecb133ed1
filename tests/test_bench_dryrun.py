"""bench.py's control flow, run for real on a CPU-only host with the device faked (tests/bench_fakes.py): every section
executes, the line carries every key of the contract, a failing side measurement costs only its own key, and the N = 2 flow
(weak-scaling line + the configs[4] sweep with both gather forms) completes over gloo.  No number here means anything."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    os.environ["QVC_BENCH_NO_SAMPLER"] = "1"
    import bench_fakes
    bench_fakes.install()
    import bench
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
''').format(root=ROOT)

SMALL = ["--steps", "3", "--warmup", "1", "--batch", "2", "--frames", "20", "--mel-frames", "140", "--sweep-utts", "6",
         "--cpu-baseline-budget-s", "3"]

CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline")


def _line(stdout):
    lines = [l for l in stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout                       # stdout carries exactly one JSON line
    return json.loads(lines[0])


def test_single_gpu_flow_produces_every_key(tmp_path):
    script = tmp_path / "dry.py"
    script.write_text(DRIVER)
    r = subprocess.run([sys.executable, str(script)] + SMALL, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = _line(r.stdout)
    for k in CONTRACT + ("cpu_baseline", "tail_roofline", "bf16_mode", "fp16_mode", "fp32_strict_mode", "decoder_only_b256",
                         "latency_5s_clip_ms", "latency_0p5s_chunk_ms", "ragged_batch", "mel_frontend", "sweep_4096_bf16"):
        assert k in line, k
    assert "failed_sections" not in line and "incomplete" not in line
    assert line["n_gpus"] == 1 and line["scaling"] == "weak" and line["higher_is_better"] is True
    assert line["gpu_launches"] == 123 * 3 and line["gpu_launches_per_step"] == 123
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == (2 * 256 * 20 + 80 * 140) * 4 and e2e["d2h_bytes_per_step"] == 2 * 320 * 20 * 4
    roof = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof, k
    assert roof["bound"] == "tensor" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0
    sweep = line["sweep_4096_bf16"]
    assert sweep["scaling"] == "strong" and sweep["utterances"] == 6 and "device_gather_form" in sweep
    assert set(line["decoder_only_b256"]["modes"]) == {"tf32", "fp16", "bf16"}
    for name in ("latency_5s_clip_ms", "latency_0p5s_chunk_ms"):
        for prec in ("tf32", "fp16", "bf16"):
            assert line[name][prec]["calls"] == 1000 and "cached_speaker_cuda_graph" in line[name][prec]


def test_a_failing_side_measurement_costs_only_its_key(tmp_path):
    script = tmp_path / "dry.py"
    script.write_text(DRIVER)
    env = dict(os.environ, QVC_FAKE_FAIL="tail")
    r = subprocess.run([sys.executable, str(script)] + SMALL + ["--no-cpu-baseline"], capture_output=True, text=True, cwd=ROOT,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = _line(r.stdout)
    assert "tail_roofline" in line["failed_sections"] and "injected failure" in line["failed_sections"]["tail_roofline"]
    assert "tail_roofline" not in line and "injected failure" in r.stderr
    for k in CONTRACT + ("decoder_only_b256", "ragged_batch", "mel_frontend", "latency_5s_clip_ms"):
        assert k in line, k


def test_two_rank_flow_over_gloo(tmp_path):
    script = tmp_path / "dry.py"
    script.write_text(DRIVER)
    port = 29500 + os.getpid() % 400
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script), "--gpus", "2"] + SMALL,
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = _line(r.stdout)                                # rank 0 alone prints
    for k in CONTRACT:
        assert k in line, k
    assert line["n_gpus"] == 2 and "incomplete" not in line
    sweep = line["sweep_4096_bf16"]
    assert sweep["n_gpus"] == 2 and sweep["utterances_per_gpu"] == 3 and sweep["device_gather_form"]["value"] > 0


def test_a_rank_that_fails_after_the_headline_leaves_a_partial_line(tmp_path):
    """Rank 1 dies in the end-to-end section: it leaves quietly, rank 0 waits in the next collective until its deadline and
    then prints the headline measured so far, marked incomplete -- the run ends with a line and rc 0, not a watchdog abort."""
    script = tmp_path / "dry.py"
    script.write_text(DRIVER)
    port = 29500 + (os.getpid() + 7) % 400
    env = dict(os.environ, QVC_FAKE_FAIL="e2e_rank1", QVC_BENCH_DEADLINE_SCALE="0.02")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script), "--gpus", "2"] + SMALL,
                       capture_output=True, text=True, cwd=ROOT, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = _line(r.stdout)
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"] is None
    # over gloo the surviving rank's collective raises when its peer is gone (exception path); over NCCL it waits and the
    # deadline fires: either way the partial line is what ends the run
    assert "end-to-end timing" in line["incomplete"] or "exception" in line["incomplete"]
    assert "injected failure on rank 1" in r.stderr


def test_a_rank_that_hangs_ends_with_the_partial_line_at_the_deadline(tmp_path):
    """The 8-GPU incident of round 2 (profiles/r02_n8_incident.log), replayed: one rank never leaves the end-to-end section.
    Every rank gives up at the section's deadline, rank 0 with the headline it had already measured."""
    script = tmp_path / "dry.py"
    script.write_text(DRIVER)
    port = 29500 + (os.getpid() + 13) % 400
    env = dict(os.environ, QVC_FAKE_FAIL="hang_rank1", QVC_BENCH_DEADLINE_SCALE="0.03")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script), "--gpus", "2"] + SMALL,
                       capture_output=True, text=True, cwd=ROOT, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = _line(r.stdout)
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"] is None
    assert "end-to-end timing" in line["incomplete"] and "did not finish" in r.stderr
