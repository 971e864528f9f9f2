"""N > 1 host logic on CPU (gloo, world_size 2): stream sharding, max-over-ranks protocol, and the
reference arm's "rank 0 prints, the others exit 0" rule."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(nproc, *args, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), *args]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=300)


def test_stream_range_partition():
    from lewton_b200.sharding import owner_of, stream_range
    for n in (0, 1, 7, 512, 4096, 4097):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = stream_range(n, w, r)
                assert 0 <= lo <= hi <= n and hi - lo in (n // w, n // w + 1)
                seen += list(range(lo, hi))
                for s in (lo, hi - 1):
                    if lo < hi:
                        assert owner_of(s, n, w) == r
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        stream_range(4, 2, 2)


@pytest.mark.parametrize("n_streams", [4096, 5])
def test_two_ranks_gloo(n_streams):
    r = torchrun(2, os.path.join("tests", "rank_worker.py"), str(n_streams))
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"RANKS_OK 2 {n_streams}" in r.stdout


def test_reference_arm_under_torchrun_prints_once():
    r = torchrun(2, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", port=29612)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["n_gpus"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
